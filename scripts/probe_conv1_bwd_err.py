"""Probe: conv1 weight / bias gradient, tcgen05 kernel vs the FP32-FMA kernel, by batch size and input dtype."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo"); C1 = importlib.import_module(pkg.__name__ + ".conv1")
L = pkg._capi.lib()
torch.manual_seed(0)
conv = P.TINet().cuda().cnn_base[0]
g = torch.Generator().manual_seed(2)


def grads(x, gy):
    conv.weight.grad = None; conv.bias.grad = None
    y = C1.conv1_relu(x, conv)
    y.backward(gy)
    torch.cuda.synchronize()
    return conv.weight.grad.clone(), conv.bias.grad.clone()


for dtype in ("u8", "f32"):
    for B in (37, 150, 300, 700, 2048, 4096):
        codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (B, 5, 289), generator=g)].cuda()
        x = codes[:, 1:5] if dtype == "u8" else P.decode_matrix(codes[:, 1:5]).contiguous()
        gy = torch.randn((B, 64, 33, 33), generator=torch.Generator().manual_seed(3)).cuda().to(torch.bfloat16)
        L.ta_debug_conv1_bwd_tc(0)
        gw0, gb0 = grads(x, gy)
        L.ta_debug_conv1_bwd_tc(1)
        gw, gb = grads(x, gy)
        gw2, gb2 = grads(x, gy)
        print(f"{dtype} B {B:5d}: tiles {(B * 289 + 127) // 128:6d}  dW rel err {float((gw - gw0).abs().max() / gw0.abs().max()):.2e}  "
              f"db rel err {float((gb - gb0).abs().max() / gb0.abs().max()):.2e}  run-to-run {float((gw - gw2).abs().max() / gw0.abs().max()):.1e}  "
              f"fail {L.ta_debug_conv1_tc_failed()}", flush=True)
