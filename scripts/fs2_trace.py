"""Pipeline trace of k_fs2 (SB_FS2_TRACE): where does a CTA's time go, and how do consecutive launches follow each other?
Run on the GPU box:   python scripts/fs2_trace.py [rig] [mode]     mode: single | stream4 | graph1 | graph4"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
path = os.environ.setdefault("SB_FS2_TRACE", "gpurun_out/fs2_trace.bin")      # needs a library built with -DSB_FS2_TRACE (STITCHB200_LIB=variants/...)
import torch                               # noqa: E402
import stitchingvideo_b200 as sv          # noqa: E402
from stitchingvideo_b200 import capi, rigs      # noqa: E402

rig = sys.argv[1] if len(sys.argv) > 1 else "c2"
mode = sys.argv[2] if len(sys.argv) > 2 else "single"
Ks, Rs, spec = rigs.cameras(rig)
n = spec["n_used"]
comp = sv.Compositor((spec["W"], spec["H"]), Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], gains=spec["gain_values"])
sets = [[capi.DeviceImage.from_torch(torch.from_numpy(rigs.frame(rig, s, i)).cuda()) for i in range(n)] for s in range(8)]
dump = capi.lib().sb_debug_fs2_trace_dump
if mode == "single":
    for k in range(3):
        comp.compose(sets[k])
    dump()
    comp.compose(sets[3])
elif mode == "stream4":
    comp.set_depth(4)
    slots = [comp.enqueue(sets[k % 8], None) for k in range(4)]
    for s in slots:
        comp.wait(s)
    dump()
    for rep in range(3):
        slots = [comp.enqueue(sets[k % 8], None) for k in range(4)]
        for s in slots:
            comp.wait(s)
else:
    comp.set_depth(1 if mode == "graph1" else 4)
    b = comp.batch([sets[k % 8] for k in range(12)], [None] * 12)     # (trace buffers are bound at capture time: 12 eager + 12 captured)
    b.launch(); b.wait()
    b.launch(); b.wait()
nl = dump()
t = np.fromfile(path, dtype=np.uint64).reshape(nl, -1, 64, 8).astype(np.int64)
if mode.startswith("graph"):
    t, nl = t[12:], nl - 12
t0 = t[t > 0].min()
print("mode %s: %d launches traced" % (mode, nl))
for k in range(nl):
    tk = t[k]
    v = tk[tk > 0]
    issued, wait, landed, done = [np.where(tk[:, :, j] > 0, tk[:, :, j] - t0, -1) for j in range(4)]
    fin = done.max(axis=1)
    print("launch %2d: first stamp %8.2f us  first landed (median CTA) %8.2f  CTA finish min %8.2f median %8.2f max %8.2f   span %.2f us" % (
        k, (v.min() - t0) / 1e3, np.median(landed[:, 0]) / 1e3, fin.min() / 1e3, np.median(fin) / 1e3, fin.max() / 1e3, (v.max() - v.min()) / 1e3))
tk = t[nl - 1]
issued, wait, landed, done = [np.where(tk[:, :, j] > 0, tk[:, :, j] - tk[tk > 0].min(), -1) for j in range(4)]
ok = done >= 0
w = np.where(ok, landed - wait, 0); c = np.where(ok, done - landed, 0); lat = np.where(ok, landed - issued, 0)
pstart, pend = [np.where(tk[:, :, j] > 0, tk[:, :, j] - tk[tk > 0].min(), -1) for j in (4, 5)]
print("producer, per tile (ns): blocked on ring space mean %.0f, descriptor + issue mean %.0f" % (np.where(ok, issued - pstart, 0)[ok].mean(), np.where(ok, pend - issued, 0)[ok].mean()))
print("last launch, per tile (ns): wait mean %.0f  compute mean %.0f  copy latency mean %.0f p90 %.0f;  share of a group's time spent waiting %.2f" % (
    w[ok].mean(), c[ok].mean(), lat[ok].mean(), np.percentile(lat[ok], 90), w.sum() / (w.sum() + c.sum())))
if "-v" in sys.argv:
    b = 0
    for s in range(0, int(ok[b].sum())):
        print("  CTA0 tile %2d pstart %7.2f issued %7.2f pend %7.2f | wait %7.2f landed %7.2f done %7.2f" % (s, pstart[b, s] / 1e3, issued[b, s] / 1e3, pend[b, s] / 1e3, wait[b, s] / 1e3, landed[b, s] / 1e3, done[b, s] / 1e3))
