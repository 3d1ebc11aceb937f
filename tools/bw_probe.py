import torch, time
dev='cuda'
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    s=torch.cuda.Event(enable_timing=True); e=torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e)/n*1e-3
N=1<<29  # 512M floats = 2 GiB
a=torch.empty(N, device=dev, dtype=torch.float32); b=torch.empty(N, device=dev, dtype=torch.float32)
t=timeit(lambda: a.fill_(1.0)); print("fill  write-only: %.0f GB/s"%(N*4/t/1e9))
t=timeit(lambda: b.copy_(a)); print("copy  r+w: %.0f GB/s"%(2*N*4/t/1e9))
t=timeit(lambda: a.sum()); print("sum   read-only: %.0f GB/s"%(N*4/t/1e9))
# 334 MB write like our obs (L2 effects)
M=1638400*51
c=torch.empty(M, device=dev, dtype=torch.float32)
t=timeit(lambda: c.fill_(1.0), 50); print("fill 334MB: %.0f GB/s  (%.1f us)"%(M*4/t/1e9, t*1e6))
d=torch.empty(1638400*22, device=dev, dtype=torch.float32)  # 88MB*... reads
def mix():
    c.fill_(1.0); d.sum()
t=timeit(mix, 50); print("fill 334MB + read 144MB sequential kernels: %.1f us"%(t*1e6))
