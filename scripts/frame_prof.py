"""One frame kernel sequence at a time, default launch structure (the command behind the ncu --set full captures in profiles/).
    python scripts/frame_prof.py c2|c3|app6 [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import capi, rigs
rig = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 12
Ks, Rs, spec = rigs.cameras(rig)
n, size = spec["n_used"], (spec["W"], spec["H"])
kw = {}
if spec.get("block_gains"):
    probe = sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"])
    kw["gain_maps"] = rigs.block_gain_maps(rig, [probe.camera_roi(i)[2:] for i in range(n)])
    del probe
if spec.get("crop"):
    kw["crop"], kw["crop_app_fill"] = spec["crop"], spec.get("crop_app_fill", False)
comp = sv.Compositor(size, Ks, Rs, warper=spec["warper"], scale=spec["scale"], blender=spec["blender"], num_bands=5,
                     weight_type=sv.CV_32F, sharpness=0.02, gains=spec["gain_values"], output_type=sv.CV_8UC3, **kw)
sets = [[capi.DeviceImage.from_torch(torch.from_numpy(rigs.frame(rig, s, i)).cuda()) for i in range(n)] for s in range(4)]
comp.set_depth(1)
for it in range(frames):
    comp.wait(comp.enqueue(sets[it % 4], None))
torch.cuda.synchronize()
print("done", rig, frames, comp.kernel_plan() if hasattr(comp, "kernel_plan") else "")
