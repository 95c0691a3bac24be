"""Development probe: conv1 forward through the tcgen05 kernel (TA_CONV1_TC=1) vs the fp32 torch layer."""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo"); C1 = importlib.import_module(pkg.__name__ + ".conv1")
L = pkg._capi.lib()
torch.manual_seed(0)
net = P.TINet().cuda(); conv = net.cnn_base[0]
g = torch.Generator().manual_seed(2)
for B in (37, 4096):
    codes = torch.tensor([0, 1, 2, 4], dtype=torch.uint8)[torch.randint(0, 4, (B, 5, 289), generator=g)].cuda()
    xf = P.decode_matrix(codes[:, 1:5]).view(B, 4, 17, 17)
    with torch.no_grad():
        want = torch.relu(conv(net.upsamplingnearest(xf)))
        got = C1.conv1_relu(codes[:, 1:5], conv)
    torch.cuda.synchronize()
    print("B", B, "tc failed flag:", L.ta_debug_conv1_tc_failed(), "max abs diff", float((got.float() - want).abs().max()), "max", float(want.abs().max()),
          "mean abs diff", float((got.float() - want).abs().mean()))
for _ in range(3): C1.conv1_relu(codes[:, 1:5], conv)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(20): C1.conv1_relu(codes[:, 1:5], conv)
torch.cuda.synchronize(); print(f"conv1 fwd B=4096: {(time.time()-t0)/20*1e3:.3f} ms")
