"""Multi-band frame structure comparison (tuning aid): per launch structure (set_fused 11 / 12) and frames in flight,
device time per frame through enqueue/wait and through a batch lap."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import capi, rigs
rig = sys.argv[1] if len(sys.argv) > 1 else "c3"
Ks, Rs, spec = rigs.cameras(rig); n = spec["n_used"]
comp = sv.Compositor((spec["W"], spec["H"]), Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], gains=spec["gain_values"])
sets = [[capi.DeviceImage.from_torch(torch.from_numpy(rigs.frame(rig, s, i)).cuda()) for i in range(n)] for s in range(8)]
modes = [int(m) for m in sys.argv[2].split(',')] if len(sys.argv) > 2 else [11, 12]
for mode in modes:
    comp.set_fused(mode)
    for it in range(3):
        recs = comp.profile_frame(sets[it % 4])
    print("mode", mode, "total %.1f us:" % (sum(r["ms"] for r in recs) * 1e3), ", ".join("%s %.1f" % (r["name"], r["ms"] * 1e3) for r in recs))
    for depth in (1, 4, 8):
        comp.set_depth(depth)
        def run(k):
            slots = []
            for i in range(k):
                if i >= depth: comp.wait(slots[i - depth])
                slots.append(comp.enqueue(sets[i % 8], None))
            for s in slots[-depth:]: comp.wait(s)
        run(16)
        comp.mark(0); run(320); comp.mark(1)
        a = comp.marked_ms() * 1e3 / 320
        lap = comp.batch([sets[f % 8] for f in range(32)], [None] * 32)
        for _ in range(3): lap.launch()
        lap.wait()
        comp.mark(0)
        for _ in range(10): lap.launch()
        comp.mark(1); lap.wait()
        b = comp.marked_ms() * 1e3 / 320
        print("mode %d depth %d: enqueue/wait %.1f us per frame, batch lap (mode %s) %.1f us" % (mode, depth, a, lap.mode, b))
        del lap
