"""A few GAE launches at the BASELINE configs[3] size (128 x 16384) over rotating buffers, for ncu."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
A = importlib.import_module(pkg.__name__ + ".advantage")
dev = torch.device("cuda:0")
T, N, R = 128, 16384, 8
r = [torch.randn(T, N, device=dev) for _ in range(R)]
v = [torch.randn(T, N, device=dev) for _ in range(R)]
d = [(torch.rand(T, N, device=dev) < 0.02).to(torch.uint8) for _ in range(R)]
adv = [torch.empty(T, N, device=dev) for _ in range(R)]
ret = [torch.empty(T, N, device=dev) for _ in range(R)]
lv = torch.randn(N, device=dev)
for b in range(R):
    A.gae(r[b], v[b], d[b], 0.99, 0.95, True, last_value=lv, out=(adv[b], ret[b]))
torch.cuda.synchronize()
print("ok")
