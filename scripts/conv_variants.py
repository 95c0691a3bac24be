"""fwd+bwd time of the TINet conv stack under layout / dtype variants (which cuDNN path is fastest?)."""
import time, torch, torch.nn as nn
dev = torch.device("cuda:0")
B = 4096
def stack():
    return nn.Sequential(nn.Conv2d(4, 64, 4, 2), nn.ReLU(), nn.Conv2d(64, 64, 3, 2), nn.ReLU(), nn.Conv2d(64, 128, 4, 2), nn.ReLU(),
                         nn.Conv2d(128, 256, 3, 2), nn.ReLU(), nn.Flatten(), nn.Linear(2304, 256)).to(dev)
def bench(tag, cl, dtype, amp):
    net = stack()
    if cl: net = net.to(memory_format=torch.channels_last)
    if not amp and dtype != torch.float32: net = net.to(dtype)
    x = torch.randn(B, 4, 68, 68, device=dev, dtype=torch.float32 if amp else dtype)
    if cl: x = x.contiguous(memory_format=torch.channels_last)
    def step():
        with torch.autocast("cuda", dtype=dtype, enabled=amp):
            y = net(x)
        y.float().sum().backward()
    for _ in range(3): step()
    torch.cuda.synchronize(); t0 = time.time()
    for _ in range(10): step()
    torch.cuda.synchronize()
    print(f"{tag:40s} {(time.time()-t0)/10*1e3:7.2f} ms fwd+bwd (B={B})", flush=True)
bench("bf16 autocast channels_last (current)", True, torch.bfloat16, True)
bench("bf16 autocast NCHW", False, torch.bfloat16, True)
bench("bf16 pure channels_last", True, torch.bfloat16, False)
bench("bf16 pure NCHW", False, torch.bfloat16, False)
bench("fp16 pure channels_last", True, torch.float16, False)
torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
bench("fp32 tf32 channels_last", True, torch.float32, False)
bench("fp32 tf32 NCHW", False, torch.float32, False)
