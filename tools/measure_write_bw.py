"""Write-only / read-only / copy HBM bandwidth on this GPU with plain torch kernels (context for the roofline denominators:
round 1 on B200: fill 7.5 TB/s, sum 6.9 TB/s, copy 6.66 TB/s)."""
import torch
x = torch.empty(1<<29, dtype=torch.float64, device='cuda')  # 4 GiB
y = torch.empty_like(x)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e-3
b = x.numel()*8
print("fill (write only) GB/s", b/t(lambda: x.fill_(1.5))/1e9)
print("zero_ (memset)    GB/s", b/t(lambda: x.zero_())/1e9)
print("copy (r+w)        GB/s", 2*b/t(lambda: y.copy_(x))/1e9)
print("sum (read only)   GB/s", b/t(lambda: x.sum())/1e9)
# strided-column writes like the linearize kernel: 216 rows of 2^20 doubles
A = torch.empty((216, 1<<20), dtype=torch.float64, device='cuda')
print("fill 216x2^20     GB/s", A.numel()*8/t(lambda: A.fill_(2.0))/1e9)
