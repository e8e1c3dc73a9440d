import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import torch
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import rigs, capi
Ks,Rs,spec=rigs.cameras('c2'); n=spec['n_used']; size=(spec['W'],spec['H'])
comp=sv.Compositor(size,Ks,Rs,warper=spec['warper'],scale=spec['scale'],blender=spec['blender'],gains=spec['gain_values'])
sets=[[torch.from_numpy(rigs.frame('c2',s,i)).cuda() for i in range(n)] for s in range(3)]
dsets=[[capi.DeviceImage.from_torch(t) for t in s] for s in sets]
for var in (10,11):
    comp._h and capi.lib().sb_compositor_set_fused(comp._h, var)
    for it in range(3): comp.profile_frame(dsets[it%3])
    ms=[sum(r['ms'] for r in comp.profile_frame(dsets[it%3])) for it in range(10)]
    print('variant',var-10,'ms per frame: min %.4f med %.4f'%(min(ms), sorted(ms)[5]))
