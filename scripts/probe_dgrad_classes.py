"""Probe: conv2's data gradient (3x3, stride 2, 128 -> 64 channels, 16x16 -> 33x33) as four stride-1 convolutions of
dz, one per output-pixel parity class, vs the GEMM + col2im path and cuDNN's own strided dgrad."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F, twoarmy_b200 as pkg
C1 = importlib.import_module(pkg.__name__ + ".conv1")
import ctypes as C
dev = "cuda:0"
B = int(os.environ.get("B", "4096"))
torch.manual_seed(0)
w = (torch.randn(128, 64, 3, 3, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)   # conv2.weight [co, ci, ky, kx]
dz = torch.randn(B, 128, 16, 16, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)


def class_weights(w):
    """class (pa, pb) -> conv2d weight [ci, co, kh, kw] for out[i, j] = sum_{u, v} dz[i - u, j - v] . W[:, :, ky(u), kx(v)]"""
    out = {}
    for pa in (0, 1):
        kys = [0, 2] if pa == 0 else [1]
        for pb in (0, 1):
            kxs = [0, 2] if pb == 0 else [1]
            k = w[:, :, kys][:, :, :, kxs]                      # [co, ci, u, v]
            k = k.flip(2, 3).permute(1, 0, 2, 3)                # conv2d correlates: flip; -> [ci, co, kh, kw]
            out[(pa, pb)] = k.contiguous(memory_format=torch.channels_last)
    return out


cw = class_weights(w)


def planes(dz):
    return {(pa, pb): F.conv2d(dz, cw[(pa, pb)], padding=(1 - pa, 1 - pb)) for pa in (0, 1) for pb in (0, 1)}


def gemm_col2im(dz):
    dz_rows = dz.permute(0, 2, 3, 1).reshape(-1, 128)
    dcols = dz_rows @ w.permute(0, 2, 3, 1).reshape(128, 9 * 64)
    gx = torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pkg._capi.check(pkg._capi.lib().ta_col2im_s2(C.c_void_p(dcols.data_ptr()), C.c_void_p(gx.data_ptr()), B, 33, 33, 64, 3, st))
    return gx


pl = planes(dz)
ref = gemm_col2im(dz).float()
for (pa, pb), t in pl.items():
    got = t.permute(0, 2, 3, 1).float()                     # [B, Hc, Wc, 64]
    want = ref[:, pa::2, pb::2]
    print((pa, pb), tuple(t.shape), "max abs diff", float((got - want).abs().max()), "of", float(want.abs().max()))


def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


torch.backends.cudnn.benchmark = True
print(f"4 class convs: {timeit(lambda: planes(dz)):.1f} us;  GEMM + col2im: {timeit(lambda: gemm_col2im(dz)):.1f} us")
for c in pl:
    print(c, f"{timeit(lambda: F.conv2d(dz, cw[c], padding=(1 - c[0], 1 - c[1]))):.1f} us")
x = torch.randn(B, 64, 33, 33, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
print(f"cuDNN strided dgrad: {timeit(lambda: torch.ops.aten.convolution_backward(dz, x, w, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False])):.1f} us")
