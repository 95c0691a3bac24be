"""torch.profiler table of a few PPO.update minibatch steps (where does the update's time go?)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0")
torch.manual_seed(0)
agent = P.PPO(device=dev)
B = int(os.environ.get("B", "16384")); mb = int(os.environ.get("MB", "4096"))
g = torch.Generator(device=dev).manual_seed(1)
buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8),
       "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
       "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
       "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
agent.update(buf, minibatch=mb, epochs=1)
torch.cuda.synchronize()
t0 = time.time(); agent.update(buf, minibatch=mb, epochs=2); torch.cuda.synchronize(); dt = time.time() - t0
steps = 2 * (B // mb)
print(f"update: {dt/steps*1e3:.2f} ms per optimiser step (minibatch {mb}), {2*B/dt:.0f} sample-epochs/s")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    agent.update(buf, minibatch=mb, epochs=1)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
