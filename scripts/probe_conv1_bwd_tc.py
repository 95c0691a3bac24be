"""Development probe: conv1 weight / bias gradient through the tcgen05 kernel vs the FP32-FMA kernel."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo"); C1 = importlib.import_module(pkg.__name__ + ".conv1")
L = pkg._capi.lib()
torch.manual_seed(0)
net = P.TINet().cuda(); conv = net.cnn_base[0]
g = torch.Generator().manual_seed(2)


def grads(codes, gy):
    conv.weight.grad = None; conv.bias.grad = None
    y = C1.conv1_relu(codes[:, 1:5], conv)
    (y.float() * gy).sum().backward()
    torch.cuda.synchronize()
    return conv.weight.grad.clone(), conv.bias.grad.clone()


for B in (37, 700, 4096):
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (B, 5, 289), generator=g)].cuda()
    gy = torch.randn((B, 64, 33, 33), generator=torch.Generator().manual_seed(3)).cuda().to(torch.bfloat16).float()
    L.ta_debug_conv1_bwd_tc(0)
    gw0, gb0 = grads(codes, gy)
    if True:
        L.ta_debug_conv1_bwd_tc(1)
        gw, gb = grads(codes, gy)
        print(f"B {B}: fail flag {L.ta_debug_conv1_tc_failed()}  dW max abs diff {float((gw - gw0).abs().max()):.5f} of {float(gw0.abs().max()):.3f}"
              f"   db max abs diff {float((gb - gb0).abs().max()):.5f} of {float(gb0.abs().max()):.3f}", flush=True)
for tc in (0, 1):
    L.ta_debug_conv1_bwd_tc(tc)
    y = C1.conv1_relu(codes[:, 1:5], conv)
    gyb = gy.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    for _ in range(3): grads(codes, gy)
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10): grads(codes, gy)
    print(f"tc={tc}: fwd+bwd wall {(time.time() - t0) / 10 * 1e3:.3f} ms")
# kernel time: direct C calls between CUDA events
import ctypes as C
x = codes[:, 1:5]
y = C1.conv1_relu(x, conv).detach()
yb = y.permute(0, 2, 3, 1).contiguous() if y.stride(1) != 1 else y   # channels-last storage [B,33,33,64]
dyb = torch.randn((B, 33, 33, 64), device="cuda").to(torch.bfloat16)
dw4 = torch.empty((256, 16), device="cuda"); db4 = torch.empty((256,), device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr())
for tc in (0, 1):
    L.ta_debug_conv1_bwd_tc(tc)
    for _ in range(3): L.ta_conv1_bwd(vp(x), 1, x.stride(0), vp(y), vp(dyb), B, vp(dw4), vp(db4), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.ta_conv1_bwd(vp(x), 1, x.stride(0), vp(y), vp(dyb), B, vp(dw4), vp(db4), st)
    e1.record(); torch.cuda.synchronize()
    print(f"tc={tc}: ta_conv1_bwd B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call  (dy + y = {2 * dyb.numel() * 2 / 1e6:.0f} MB)")
w4, b4 = C1.fold(conv.weight.detach(), conv.bias.detach())
w4 = w4.float().contiguous(); b4 = b4.float().contiguous()
yo = torch.empty((B, 33, 33, 64), device="cuda", dtype=torch.bfloat16)
for tc in (0, 1):
    L.ta_debug_conv1_tc(tc)
    for _ in range(3): L.ta_conv1_fwd(vp(x), 1, x.stride(0), vp(w4), vp(b4), B, vp(yo), st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.ta_conv1_fwd(vp(x), 1, x.stride(0), vp(w4), vp(b4), B, vp(yo), st)
    e1.record(); torch.cuda.synchronize()
    print(f"tc={tc}: ta_conv1_fwd B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call  (y = {yo.numel() * 2 / 1e6:.0f} MB)")
planes = torch.randn((B, 17, 17, 4, 64), device="cuda").to(torch.bfloat16)
mask = torch.empty((B * 289 * 8,), dtype=torch.int32, device="cuda")
for _ in range(3): L.ta_conv1_fwd_mask(vp(x), 1, x.stride(0), vp(w4), vp(b4), B, vp(yo), vp(mask), st)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): L.ta_conv1_fwd_mask(vp(x), 1, x.stride(0), vp(w4), vp(b4), B, vp(yo), vp(mask), st)
e1.record(); torch.cuda.synchronize()
print(f"ta_conv1_fwd_mask B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call")
for name, yarg, marg in (("y", vp(y), None), ("bit mask", None, vp(mask))):
    for _ in range(3): L.ta_conv1_bwd_planes(vp(x), 1, x.stride(0), yarg, marg, vp(planes), 0, B, vp(dw4), vp(db4), st)
    res = dw4.clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): L.ta_conv1_bwd_planes(vp(x), 1, x.stride(0), yarg, marg, vp(planes), 0, B, vp(dw4), vp(db4), st)
    e1.record(); torch.cuda.synchronize()
    print(f"ta_conv1_bwd_planes ({name}) B={B}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per call; |dw4| = {float(res.abs().sum()):.1f}")
