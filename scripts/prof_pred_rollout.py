"""Where the time of one acting step of the PPO + predictor agent goes (BASELINE configs[4], 8192 envs): torch profiler
table of select_action_frames, and CUDA-event times of its stages (encoder, LSTM, decoder, actor)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
M = importlib.import_module(pkg.__name__ + ".predictor")
dev = torch.device("cuda:0")
torch.manual_seed(0)
agent = M.ppo_predictor(device=dev)
B = int(os.environ.get("B", "8192"))
frames = torch.randint(0, 3, (B, 4, 289), dtype=torch.uint8, device=dev)
pos = torch.zeros((B, 4, 2), device=dev); goal = torch.zeros((B, 2), device=dev)


def ev(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


ms, _ = ev(lambda: agent.select_action_frames(frames, pos, goal))
print(f"select_action_frames {ms:.2f} ms")
x = M.decode_matrix(frames).float()
with torch.no_grad(), agent._amp():
    agent.encoder.eval(); agent.decoder.eval(); agent.predictor.eval()
    ms, (z_c, _) = ev(lambda: agent.encoder(x.reshape(-1, 1, 289)))
    print(f"encoder {ms:.2f} ms")
    z_c = z_c.view(-1, 4, 64, 4, 4)
    ms, (z_pred, _) = ev(lambda: agent.predictor(z_c))
    print(f"lstm fast {ms:.2f} ms")
    os.environ["TA_LSTM_FAST"] = "0"
    ms, _ = ev(lambda: agent.predictor(z_c))
    print(f"lstm cudnn {ms:.2f} ms")
    os.environ["TA_LSTM_FAST"] = "1"
    ms, _ = ev(lambda: agent.decoder(z_pred[:, 3:7]))
    print(f"decoder {ms:.2f} ms")
    cat = agent._cat(frames)
    ms, _ = ev(lambda: agent.actor(cat, pos, goal))
    print(f"actor (8-channel TINet) {ms:.2f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    agent.select_action_frames(frames, pos, goal)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))
