"""A few pred_states calls of the PPO + predictor agent (2048 envs) for ncu: pred_encoder_kernel / pred_decoder_kernel."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
M = importlib.import_module(pkg.__name__ + ".predictor")
dev = torch.device("cuda:0")
torch.manual_seed(0)
agent = M.ppo_predictor(device=dev)
frames = torch.randint(0, 3, (2048, 4, 289), dtype=torch.uint8, device=dev)
for _ in range(3):
    agent.pred_states(frames)
torch.cuda.synchronize()
print("ok")
