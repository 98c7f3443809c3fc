import torch,time
n=1<<28
a=torch.empty(n,dtype=torch.uint8).pin_memory(); d=torch.empty(n,dtype=torch.uint8,device="cuda")
b=torch.empty(n,dtype=torch.uint8).pin_memory(); e=torch.empty(n,dtype=torch.uint8,device="cuda")
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def run(h2d,d2h,reps=8):
    torch.cuda.synchronize(); t=time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): d.copy_(a,non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): b.copy_(e,non_blocking=True)
    torch.cuda.synchronize(); dt=time.perf_counter()-t
    return reps*n/dt/1e9
print("H2D only", run(1,0)); print("D2H only", run(0,1)); 
x=run(1,1); print("both: each direction", x, "total", 2*x)
# many small copies from 8 streams: 8.3MB H2D + 3.1MB D2H each
ss=[torch.cuda.Stream() for _ in range(8)]
ha=[torch.empty(8294400,dtype=torch.uint8).pin_memory() for _ in range(8)]; da=[torch.empty(8294400,dtype=torch.uint8,device="cuda") for _ in range(8)]
hb=[torch.empty(3110400,dtype=torch.uint8).pin_memory() for _ in range(8)]; db=[torch.empty(3110400,dtype=torch.uint8,device="cuda") for _ in range(8)]
torch.cuda.synchronize(); t=time.perf_counter()
R=100
for r in range(R):
    for i in range(8):
        with torch.cuda.stream(ss[i]):
            da[i].copy_(ha[i],non_blocking=True); hb[i].copy_(db[i],non_blocking=True)
torch.cuda.synchronize(); dt=time.perf_counter()-t
print("8 streams x (8.3MB in + 3.1MB out): frames/s", 8*R/dt, "H2D GB/s", 8*R*8.2944e-3/dt, "D2H GB/s", 8*R*3.1104e-3/dt)
