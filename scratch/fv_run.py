"""time one build of the library (STITCHB200_LIB) on a rig and print a digest of its panorama"""
import sys, hashlib, numpy as np
sys.path.insert(0, '/root/repo')
import torch
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import rigs, capi
rig = sys.argv[1] if len(sys.argv) > 1 else 'c2'
Ks, Rs, spec = rigs.cameras(rig); n = spec['n_used']; size = (spec['W'], spec['H'])
comp = sv.Compositor(size, Ks, Rs, warper=spec['warper'], scale=spec['scale'], blender=spec['blender'], gains=spec['gain_values'])
frames = [rigs.frame(rig, 0, i) for i in range(n)]
pano, mask = comp.compose(frames)
dig = hashlib.md5(pano.tobytes() + mask.tobytes()).hexdigest()
sets = [[torch.from_numpy(rigs.frame(rig, s, i, smooth=0)).cuda() for i in range(n)] for s in range(3)]
dsets = [[capi.DeviceImage.from_torch(t) for t in s] for s in sets]
for it in range(5): comp.profile_frame(dsets[it % 3])
ms = [sum(r['ms'] for r in comp.profile_frame(dsets[it % 3])) for it in range(20)]
print('%s %s us/frame: min %.1f med %.1f  digest %s' % (sys.argv[2] if len(sys.argv) > 2 else '', rig, min(ms) * 1e3, sorted(ms)[10] * 1e3, dig))
