"""Development probe: wall time of EAGER optimiser steps and of the graph capture (the per-update() overhead beside the replays)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0")
torch.manual_seed(0)
agent = P.PPO(device=dev)
mb = 4096; B = mb * 4
g = torch.Generator(device=dev).manual_seed(1)
buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8),
       "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
       "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
       "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
for rnd in range(3):
    step, Bn, bs, _ = agent._make_step(buf, minibatch=mb)
    idx = torch.randperm(B, device=dev)[:mb].contiguous()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    ts = []
    with torch.cuda.stream(side):
        for _ in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            step(idx)
            t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
            ts.append(f"{(t1 - t0) * 1e3:.1f}+{(t2 - t1) * 1e3:.1f}")
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        step(idx)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    gr.replay(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"round {rnd}: eager steps (issue + drain ms) {ts}; capture {(t1 - t0) * 1e3:.1f} ms; first replay {(t2 - t1) * 1e3:.1f} ms", flush=True)
    del gr
