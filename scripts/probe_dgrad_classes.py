"""Probe: the data gradient of TINet's conv2 (3x3, stride 2, 64 -> 64 channels, dz 16x16 -> dx 33x33) and conv3 (4x4,
stride 2, 64 -> 128 channels, dz 7x7 -> dx 16x16) as four stride-1 convolutions of dz, one per parity class of the
input pixel, vs the GEMM + col2im path and cuDNN's own strided dgrad.   env: LAYER=2|3, B"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F, twoarmy_b200 as pkg
C1 = importlib.import_module(pkg.__name__ + ".conv1")
import ctypes as C
dev = "cuda:0"
B = int(os.environ.get("B", "4096"))
torch.manual_seed(0)
LAYER = int(os.environ.get("LAYER", "2"))
cout, cin, k, oh, hin = (64, 64, 3, 16, 33) if LAYER == 2 else (128, 64, 4, 7, 16)
w = (torch.randn(cout, cin, k, k, device=dev) * 0.05).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)   # [co, ci, ky, kx]
dz = torch.randn(B, cout, oh, oh, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
cwm = C1.parity_class_weights(w)                                   # merged: [4*cin, cout, 2, 2]
cw = []
for c in range(4):                                                  # the four separate class kernels (for comparison)
    blk = cwm[c * cin:(c + 1) * cin]
    kh, kw = (k - (c >> 1) + 1) // 2, (k - (c & 1) + 1) // 2
    cw.append(blk[:, :, 2 - kh:, 2 - kw:].contiguous(memory_format=torch.channels_last))


def planes(dz):
    return [F.conv2d(dz, wk, padding=(wk.shape[2] - 1, wk.shape[3] - 1)) for wk in cw]


def merged(dz):
    return F.conv2d(dz, cwm, padding=1)


def gemm_col2im(dz):
    dz_rows = dz.permute(0, 2, 3, 1).reshape(-1, cout)
    dcols = dz_rows @ w.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
    gx = torch.empty((B, hin, hin, cin), dtype=torch.bfloat16, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pkg._capi.check(pkg._capi.lib().ta_col2im_s2(C.c_void_p(dcols.data_ptr()), C.c_void_p(gx.data_ptr()), B, hin, hin, cin, k, st))
    return gx


pl = planes(dz)
ref = gemm_col2im(dz).float()
for idx, t in enumerate(pl):
    pa, pb = idx >> 1, idx & 1
    got = t.permute(0, 2, 3, 1).float()                     # [B, Hc, Wc, cin]
    want = ref[:, pa::2, pb::2][:, :got.shape[1], :got.shape[2]]
    print((pa, pb), tuple(t.shape), "max abs diff", float((got[:, :want.shape[1], :want.shape[2]] - want).abs().max()), "of", float(want.abs().max()))


def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


torch.backends.cudnn.benchmark = True
mg = merged(dz)
for idx, t in enumerate(pl):
    sub = mg[:, idx * cin:(idx + 1) * cin, :t.shape[2], :t.shape[3]]
    assert float((sub.float() - t.float()).abs().max()) <= 0.07, idx
print(f"layer {LAYER}, B = {B}: merged conv {timeit(lambda: merged(dz)):.1f} us; 4 class convs {timeit(lambda: planes(dz)):.1f} us;  GEMM + col2im {timeit(lambda: gemm_col2im(dz)):.1f} us")
x = torch.randn(B, cin, hin, hin, device=dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
print(f"cuDNN strided dgrad: {timeit(lambda: torch.ops.aten.convolution_backward(dz, x, w, None, [2, 2], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False])):.1f} us")
