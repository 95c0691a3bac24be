"""A few launches of the kernels changed in the last session of round 2, for ncu: the pipelined frame-stack push and
state matrix (65536 envs), the GAE launch at 128 x 16384 (plain and normalising), the predictor stacks (2048 envs)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
A = importlib.import_module(pkg.__name__ + ".advantage")
M = importlib.import_module(pkg.__name__ + ".predictor")
dev = torch.device("cuda:0"); n = 65536
env = pkg.TwoarmyVecEnv(4, n, 17, device=dev, seed=1, autoreset=False); env.reset()
s0 = torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev); s1 = torch.zeros_like(s0)
p0 = torch.zeros((n, 5, 2), device=dev); p1 = torch.zeros_like(p0)
env.stack_push(None, s0, None, p0, init_all=True)
for _ in range(3):
    env.stack_push(s0, s1, p0, p1); env.state_matrix(want_codes=True)
T, N = 128, 16384
r = torch.randn(T, N, device=dev); v = torch.randn(T, N, device=dev); d = (torch.rand(T, N, device=dev) < 0.02).to(torch.uint8)
lv = torch.randn(N, device=dev)
for _ in range(3):
    A.gae(r, v, d, 0.99, 0.95, True, last_value=lv)
    A.gae(r, v, d, 0.99, 0.95, True, last_value=lv, normalize=True)
torch.manual_seed(0)
agent = M.ppo_predictor(device=dev)
frames = torch.randint(0, 3, (2048, 4, 289), dtype=torch.uint8, device=dev)
for _ in range(3):
    agent.pred_states(frames)
torch.cuda.synchronize(); print("ok")
