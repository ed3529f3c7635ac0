import torch
dev=torch.device("cuda:0")
n=409_600_000//4
x=torch.empty(n,dtype=torch.float32,device=dev)
y=torch.empty(n,dtype=torch.float32,device=dev)
def t(f,reps=20):
    for _ in range(3): f()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(reps): f()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e)/reps*1e3
a=t(lambda: x.zero_()); print("memset 409.6MB: %.1f us  %.0f GB/s"%(a, 409.6e6/a/1e3))
a=t(lambda: x.fill_(1.5)); print("fill 409.6MB: %.1f us  %.0f GB/s"%(a, 409.6e6/a/1e3))
a=t(lambda: y.copy_(x)); print("copy 409.6MB: %.1f us  %.0f GB/s (r+w)"%(a, 2*409.6e6/a/1e3))
