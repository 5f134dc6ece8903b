"""Write-only and copy bandwidth of this GPU (context for the covariance build's HBM roofline)."""
import torch
x = torch.empty(1 << 28, dtype=torch.float64, device="cuda")   # 2 GiB
y = torch.empty_like(x)
def t(fn, reps=10):
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.fill_(1.5)); print(f"fill  2 GiB: {ms:.3f} ms  {x.numel()*8/ms/1e6:.0f} GB/s written")
ms = t(lambda: x.zero_()); print(f"zero  2 GiB: {ms:.3f} ms  {x.numel()*8/ms/1e6:.0f} GB/s written")
ms = t(lambda: y.copy_(x)); print(f"copy  2 GiB: {ms:.3f} ms  {2*x.numel()*8/ms/1e6:.0f} GB/s read+write")
ms = t(lambda: x.sum()); print(f"sum   2 GiB: {ms:.3f} ms  {x.numel()*8/ms/1e6:.0f} GB/s read")
