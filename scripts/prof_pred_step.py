"""ONE optimiser step of the PPO + predictor agent (minibatch 4096, both 8-channel nets, fused_step.FusedNet8, eager
launches) for an ncu launch list: after three warm-up steps the fourth runs between cudaProfilerStart/Stop."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
M = importlib.import_module(pkg.__name__ + ".predictor")
dev = torch.device("cuda:0")
torch.manual_seed(0)
agent = M.ppo_predictor(device=dev)
agent.use_graph = False
mb = int(os.environ.get("MB", "4096")); B = mb
g = torch.Generator(device=dev).manual_seed(1)
buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8),
       "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
       "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
       "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
step, Bn, bs, _ = agent._make_step(buf, minibatch=mb)
assert type(agent._fused["actor"]).__name__ == "FusedNet8"
idx = torch.randperm(B, device=dev)[:mb].contiguous()
for _ in range(3):
    step(idx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step(idx)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
