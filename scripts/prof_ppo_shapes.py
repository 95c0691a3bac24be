"""torch.profiler of 4 optimiser steps with input shapes (which tensors do the elementwise passes touch?)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0"); torch.manual_seed(0)
agent = P.PPO(device=dev)
B, mb = 16384, 4096
g = torch.Generator(device=dev).manual_seed(1)
buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8), "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
       "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
       "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
agent.update(buf, minibatch=mb, epochs=1); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    agent.update(buf, minibatch=mb, epochs=1); torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=42, max_shapes_column_width=70))
