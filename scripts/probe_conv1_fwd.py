"""Development probe: the two tcgen05 forward kernels of the first layer (mode 1: single-role CTAs, mode 2: warp-specialised with
tensor-map stores), with and without the ReLU bit mask, B = 4096 (609 MB / 571 MB written)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
L = pkg._capi.lib()
B = int(os.environ.get("B", "4096"))
g = torch.Generator(device="cuda").manual_seed(1)
w4 = torch.randn((256, 16), generator=g, device="cuda") * 0.3
b4 = torch.randn((256,), generator=g, device="cuda") * 0.1
x = torch.randint(0, 3, (B, 5, 289), generator=g, device="cuda", dtype=torch.uint8)
ys = [torch.empty((B, 33, 33, 64), dtype=torch.bfloat16, device="cuda") for _ in range(2)]     # alternate: 1.1 GB > L2
masks = [torch.empty((B * 289 * 8,), dtype=torch.int32, device="cuda") for _ in range(2)]
vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode in (2, 1):
    L.ta_debug_conv1_tc(mode)
    for mk in (True, False):
        def run(i):
            if mk:
                pkg._capi.check(L.ta_conv1_fwd_mask(vp(x), 1, x.stride(0), vp(w4), vp(b4), B, vp(ys[i & 1]), vp(masks[i & 1]), st))
            else:
                pkg._capi.check(L.ta_conv1_fwd(vp(x), 1, x.stride(0), vp(w4), vp(b4), B, vp(ys[i & 1]), st))
        for i in range(4):
            run(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            run(i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        mb = (B * 33 * 33 * 128 + (B * 289 * 32 if mk else 0)) / 1e6
        print(f"mode {mode} mask={mk}: {us:.1f} us, {mb / us * 1e-3 * 1e3:.0f} GB/s of output; fail {L.ta_debug_conv1_tc_failed()}")
L.ta_debug_conv1_tc(-1)
