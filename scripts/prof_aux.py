"""A few launches of every non-step kernel for ncu (stack push, state matrix, GAE, HER plan, conv1 fwd/bwd, col2im)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo"); H = importlib.import_module(pkg.__name__ + ".her")
A = importlib.import_module(pkg.__name__ + ".advantage"); C1 = importlib.import_module(pkg.__name__ + ".conv1")
dev = torch.device("cuda:0"); n = 65536
env = pkg.TwoarmyVecEnv(4, n, 17, device=dev, seed=1, autoreset=False); env.reset()
s0 = torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev); s1 = torch.zeros_like(s0)
p0 = torch.zeros((n, 5, 2), device=dev); p1 = torch.zeros_like(p0)
for _ in range(3):
    env.stack_push(None, s0, None, p0, init_all=True); env.stack_push(s0, s1, p0, p1); env.state_matrix(want_codes=True)
T, N = 128, 262144
r = torch.randn(T, N, device=dev); v = torch.randn(T, N, device=dev); d = (torch.rand(T, N, device=dev) < 0.02).to(torch.uint8); lv = torch.randn(N, device=dev)
for _ in range(3):
    A.gae(r, v, d, 0.99, 0.95, True, last_value=lv, normalize=True)
p = torch.zeros((128, 16384, 5, 2), device=dev); p[:, :, 4] = torch.randint(1, 16, (128, 16384, 2), device=dev).float()
dn = (torch.rand(128, 16384, device=dev) < 0.03)
for _ in range(3):
    H.plan(p, dn)
torch.manual_seed(0)
net = P.TINet().to(dev).to(memory_format=torch.channels_last)
x = torch.randint(0, 3, (4096, 5, 289), device=dev, dtype=torch.uint8)
for _ in range(3):
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = net(x[:, 0:4], torch.zeros(4096, 4, 2, device=dev), torch.zeros(4096, 2, device=dev))
    y.float().sum().backward()
torch.cuda.synchronize(); print("ok")
