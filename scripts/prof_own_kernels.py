"""A few launches of the own kernels that sit below their roofline, for `ncu --set full`: stack push / frame codes (65536 envs),
the tcgen05 first-layer forward (with mask) and weight gradient (planes + mask) at B = 4096, relu_bwd_bias, col2im, im2col."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0"); n = 65536
env = pkg.TwoarmyVecEnv(4, n, 17, device=dev, seed=1, autoreset=False); env.reset()
s0 = torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev); s1 = torch.zeros_like(s0)
p0 = torch.zeros((n, 5, 2), device=dev); p1 = torch.zeros_like(p0)
env.stack_push(None, s0, None, p0, init_all=True)
for _ in range(4):
    env.stack_push(s0, s1, p0, p1); env.state_matrix(want_codes=True)
torch.manual_seed(0)
agent = P.PPO(device=dev)
mb = 4096
g = torch.Generator(device=dev).manual_seed(1)
buf = {"s": torch.randint(0, 3, (mb, 5, 289), generator=g, device=dev, dtype=torch.uint8),
       "p": torch.randint(1, 16, (mb, 5, 2), generator=g, device=dev).float(),
       "a": torch.randint(0, 5, (mb, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(mb, 1),
       "r": torch.rand(mb, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(mb, 1, generator=g, device=dev) * 0.3 + 0.1)}
agent.two_streams = False
step, B, bs, _ = agent._make_step(buf, minibatch=mb)
idx = torch.arange(mb, device=dev)
for _ in range(3):
    step(idx)
torch.cuda.synchronize(); print("ok")
