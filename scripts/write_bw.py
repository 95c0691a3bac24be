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

# pure-write probes with this repo's own store patterns
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import twoarmy_b200 as pkg
L = pkg._capi.lib()
L.ta_debug_write_probe.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
for mode, name in ((0, "STG.128 grid-stride"), (1, "3 KB TMA bulk stores"), (2, "4224 B rows, 2 tiles deep"), (3, "4224 B rows, 8 tiles deep")):
    fn = lambda: L.ta_debug_write_probe(C.c_void_p(x.data_ptr()), x.numel(), mode, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    ms = t(fn); print(f"{name:24s} 1 GiB: {ms:.3f} ms  {x.numel()/ms/1e6:.0f} GB/s write")
