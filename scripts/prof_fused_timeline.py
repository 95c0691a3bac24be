"""Timeline of ONE graph-replayed optimiser step (both networks, two streams): kernel start offsets / durations / stream from
CUPTI (torch profiler).  Shows where the step's wall time goes beyond the sum of kernel times (gaps, serialisation)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0")
torch.manual_seed(0)
agent = P.PPO(device=dev)
agent.two_streams = os.environ.get("TWO", "1") == "1"
mb = 4096; B = mb * 4
g = torch.Generator(device=dev).manual_seed(1)
buf = {"s": torch.randint(0, 3, (B, 5, 289), generator=g, device=dev, dtype=torch.uint8),
       "p": torch.randint(1, 16, (B, 5, 2), generator=g, device=dev).float(),
       "a": torch.randint(0, 5, (B, 1), generator=g, device=dev), "g": torch.tensor([[2.0, 14.0]], device=dev).repeat(B, 1),
       "r": torch.rand(B, 1, generator=g, device=dev) - 0.5, "a_logp": torch.log(torch.rand(B, 1, generator=g, device=dev) * 0.3 + 0.1)}
step, Bn, bs, _ = agent._make_step(buf, minibatch=mb)
idx = torch.randperm(B, device=dev)[:mb].contiguous()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        step(idx)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    step(idx)
for _ in range(3):
    gr.replay()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    gr.replay()
e1.record(); torch.cuda.synchronize()
print(f"replay: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per step (two_streams={agent.two_streams})")
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    gr.replay()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if str(e.device_type).endswith("CUDA")]
ev.sort(key=lambda e: e.time_range.start)
t0 = ev[0].time_range.start
streams = {}
tot = 0.0
print(f"{'start us':>9s} {'dur us':>8s} {'gap':>7s} st  kernel")
last_end = {}
for e in ev:
    s = getattr(e, "stream", None)
    if s is None:
        s = e.device_index
    k = streams.setdefault(s, len(streams))
    start = e.time_range.start - t0
    gap = start - last_end.get(k, start)
    last_end[k] = start + e.device_time
    tot += e.device_time
    print(f"{start:9.1f} {e.device_time:8.1f} {gap:7.1f} {k:2d}  {e.name[:90]}")
end = max(e.time_range.start + e.device_time for e in ev) - t0
print(f"{len(ev)} device activities, sum {tot:.0f} us, span {end:.0f} us")
