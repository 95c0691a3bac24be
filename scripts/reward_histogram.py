import importlib, sys, torch
sys.path.insert(0, ".")
import twoarmy_b200 as pkg
P = importlib.import_module(pkg.__name__ + ".ppo")
dev = torch.device("cuda:0")
env = pkg.TwoarmyVecEnv(4, 4096, 17, device=dev, seed=9981, autoreset=True)
env.reset()
amap = torch.tensor([0, 1, 2, 3, 6], dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
hist = {}
for t in range(200):
    a = amap[torch.randint(0, 5, (4096,), generator=g, device=dev)]
    obs, rew, term, trunc, _ = env.step(a)
    for v, c in zip(*torch.unique(rew, return_counts=True)):
        hist[round(float(v), 3)] = hist.get(round(float(v), 3), 0) + int(c)
print("random policy, autoreset, 200 steps x 4096 envs: reward histogram", hist)
torch.manual_seed(0)
agent = P.PPO(device=dev)
env2 = pkg.TwoarmyVecEnv(4, 4096, 17, device=dev, seed=9981, autoreset=False)
roll = P.VecRollout(env2, agent, 64)
buf = roll.collect()
print("policy rollout: action histogram", torch.bincount(buf.a[:64].flatten(), minlength=5).tolist())
vals, cnt = torch.unique(buf.r[:64], return_counts=True)
print("policy rollout: reward histogram", {round(float(v), 3): int(c) for v, c in zip(vals, cnt)}, "ended", int(buf.ended[:64].sum()), "terminated", int(buf.d[:64].sum()))
