"""Few T=1 launches for ncu (rotating B batches; -s skips the warm-up launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
dev = torch.device('cuda:0')
n = int(os.environ.get('N', '65536')); V = int(os.environ.get('V', '17')); B = int(os.environ.get('B', '8'))
K = int(os.environ.get('K', '24')); T = int(os.environ.get('T', '1'))
envs = [pkg.TwoarmyVecEnv(4, n, V, device=dev, seed=1, env_id0=b * n) for b in range(B)]
for e in envs: e.reset()
amap = torch.tensor([0, 1, 2, 3, 6], dtype=torch.uint8, device=dev)
acts = amap[torch.randint(0, 5, (max(T, 16), n), device=dev)].contiguous()
if T == 1:
    outs = [dict(obs=torch.empty((n, V, V, 3), dtype=torch.uint8, device=dev), reward=torch.empty(n, device=dev),
                 terminated=torch.empty(n, dtype=torch.uint8, device=dev), truncated=torch.empty(n, dtype=torch.uint8, device=dev)) for _ in range(B)]
    for i in range(K):
        envs[i % B].step(acts[i % 16], out=outs[i % B])
else:
    for i in range(K):
        envs[i % B].rollout(acts[:T])
torch.cuda.synchronize()
print("ok")
