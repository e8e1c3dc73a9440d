import torch, time
n=31104000
h=[torch.empty(n,dtype=torch.uint8).pin_memory() for _ in range(4)]
d=[torch.empty(n,dtype=torch.uint8,device='cuda') for _ in range(4)]
ho=[torch.empty(21371040,dtype=torch.uint8).pin_memory() for _ in range(4)]
do=[torch.empty(21371040,dtype=torch.uint8,device='cuda') for _ in range(4)]
s1,s2=torch.cuda.Stream(),torch.cuda.Stream()
def run(h2d,d2h,iters=50):
    torch.cuda.synchronize(); t=time.perf_counter()
    for i in range(iters):
        if h2d:
            with torch.cuda.stream(s1): d[i%4].copy_(h[i%4],non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): ho[i%4].copy_(do[i%4],non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    return (n*iters/dt/1e9 if h2d else 0, 21371040*iters/dt/1e9 if d2h else 0)
for _ in range(2): run(True,True,10)
print('h2d only', run(True,False)); print('d2h only', run(False,True)); print('both', run(True,True))
# 5 chunks like 5 camera frames
hc=[[torch.empty(n//5,dtype=torch.uint8).pin_memory() for _ in range(5)] for _ in range(4)]
dc=[[torch.empty(n//5,dtype=torch.uint8,device='cuda') for _ in range(5)] for _ in range(4)]
torch.cuda.synchronize(); t=time.perf_counter()
for i in range(50):
    with torch.cuda.stream(s1):
        for k in range(5): dc[i%4][k].copy_(hc[i%4][k],non_blocking=True)
    with torch.cuda.stream(s2): ho[i%4].copy_(do[i%4],non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print('both, 5 chunks', n*50/dt/1e9, 21371040*50/dt/1e9, 'frames/s', 50/dt)
