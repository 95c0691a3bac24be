"""Pure-write and copy HBM bandwidth probes (context for the obs-store roofline)."""
import torch
dev = torch.device('cuda:0')
x = torch.empty(1 << 30, dtype=torch.uint8, device=dev); y = torch.empty_like(x)
def t(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
ms = t(lambda: x.zero_()); print(f"memset 1 GiB: {ms:.3f} ms  {x.numel()/ms/1e6:.0f} GB/s write")
ms = t(lambda: x.fill_(7)); print(f"fill   1 GiB: {ms:.3f} ms  {x.numel()/ms/1e6:.0f} GB/s write")
ms = t(lambda: y.copy_(x)); print(f"copy   1 GiB: {ms:.3f} ms  {2*x.numel()/ms/1e6:.0f} GB/s read+write")
