import sys, numpy as np
sys.path.insert(0,'/root/repo')
import torch
import stitchingvideo_b200 as sv
from stitchingvideo_b200 import rigs, capi
rig=sys.argv[1] if len(sys.argv)>1 else 'c3'
Ks,Rs,spec=rigs.cameras(rig); n=spec['n_used']; size=(spec['W'],spec['H'])
comp=sv.Compositor(size,Ks,Rs,warper=spec['warper'],scale=spec['scale'],blender=(sys.argv[2] if len(sys.argv)>2 else spec['blender']),gains=spec['gain_values'])
sets=[[torch.from_numpy(rigs.frame(rig,s,i,smooth=0)).cuda() for i in range(n)] for s in range(3)]
dsets=[[capi.DeviceImage.from_torch(t) for t in s] for s in sets]
for it in range(3): comp.profile_frame(dsets[it%3])
agg={}
N=10
for it in range(N):
    for r in comp.profile_frame(dsets[it%3]):
        a=agg.setdefault(r['name'],[0.0,0.0,0]); a[0]+=r['ms']; a[1]+=r['bytes']; a[2]+=1
tot=0
for k,(ms,b,c) in sorted(agg.items(), key=lambda kv:-kv[1][0]):
    print('%-18s n=%4.1f  us/frame=%7.1f  MB=%7.1f  GB/s=%6.0f'%(k,c/N,ms/N*1e3,b/N/1e6,b/ms/1e6)); tot+=ms/N
print('total us/frame %.1f'%(tot*1e3))
print('--- one frame, per launch')
for r in comp.profile_frame(dsets[0]):
    print('  %-16s %7.1f us  %7.2f MB  %6.0f GB/s'%(r['name'], r['ms']*1e3, r['bytes']/1e6, r['bytes']/max(r['ms'],1e-9)/1e6))
