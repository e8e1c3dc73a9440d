"""Target for ncu captures: build one compositor for a rig and run a few frames (device-resident inputs).
    ncu --set full --clock-control none --import-source on -k regex:'k_feather_fused|k_mb_' -s <skip> -c <n> \
        -o gpurun_out/prof python scripts/ncu_frame.py c2 3
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import stitchingvideo_b200 as sv  # noqa: E402
from stitchingvideo_b200 import capi, rigs  # noqa: E402

rig = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
Ks, Rs, spec = rigs.cameras(rig)
n, size = spec["n_used"], (spec["W"], spec["H"])
comp = sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], gains=spec["gain_values"])
sets = [[torch.from_numpy(rigs.frame(rig, s, i, smooth=0)).cuda() for i in range(n)] for s in range(2)]
dsets = [[capi.DeviceImage.from_torch(t) for t in s] for s in sets]
for it in range(frames):
    slot = comp.enqueue(dsets[it % 2], None, None)
    comp.wait(slot)
print("done", rig, frames)
