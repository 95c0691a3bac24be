"""Per-layer fwd+bwd: cuDNN conv (bf16 channels_last) vs the GEMM it is equivalent to (cuBLAS bf16,
im2col cost excluded/included with torch ops) -- is an explicit im2col + GEMM path worth building?"""
import time, torch, torch.nn.functional as F
dev = torch.device("cuda:0"); B = 4096
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time() - t0) / n * 1e3
layers = [("conv1", 4, 64, 4, 68), ("conv2", 64, 64, 3, 33), ("conv3", 64, 128, 4, 16), ("conv4", 128, 256, 3, 7)]
for name, cin, cout, k, hin in layers:
    hout = (hin - k) // 2 + 1
    x = torch.randn(B, cin, hin, hin, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(name != "conv1")
    w = torch.randn(cout, cin, k, k, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    def conv():
        y = F.conv2d(x, w, stride=2)
        y.backward(torch.ones_like(y))
    M, K = B * hout * hout, cin * k * k
    A = torch.randn(M, K, device=dev, dtype=torch.bfloat16).requires_grad_(name != "conv1")
    Wm = torch.randn(K, cout, device=dev, dtype=torch.bfloat16).requires_grad_(True)
    def gemm():
        y = A @ Wm
        y.backward(torch.ones_like(y))
    def copyA():  # stands in for im2col (write A) + col2im (read dA): two streaming passes over A
        A.detach().clone()
    tc, tg, tcp = t(conv), t(gemm), t(copyA)
    print(f"{name}: cudnn {tc:6.2f} ms | gemm fwd+bwd {tg:6.2f} ms + 2 x A-copy {2*tcp:5.2f} ms  (M={M}, K={K}, N={cout}, A={M*K*2/1e6:.0f} MB)", flush=True)
