"""T=1 launch timing: python-loop launches vs CUDA-graph replay, rotating B env batches (L2 busting).
env: N, V, B, VER, STEPS"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
dev = torch.device('cuda:0')
n = int(os.environ.get('N', '65536')); V = int(os.environ.get('V', '17')); B = int(os.environ.get('B', '8'))
ver = int(os.environ.get('VER', '4')); steps = int(os.environ.get('STEPS', '2000'))
envs = [pkg.TwoarmyVecEnv(ver, n, V, device=dev, seed=1, env_id0=b * n) for b in range(B)]
for e in envs: e.reset()
amap = torch.tensor([0, 1, 2, 3, 6], dtype=torch.uint8, device=dev)
R = 16
acts = amap[torch.randint(0, 5, (R, n), device=dev)].contiguous()
outs = [dict(obs=torch.empty((n, V, V, 3), dtype=torch.uint8, device=dev), reward=torch.empty(n, device=dev),
             terminated=torch.empty(n, dtype=torch.uint8, device=dev), truncated=torch.empty(n, dtype=torch.uint8, device=dev)) for _ in range(B)]
def run(k0, k):
    for i in range(k0, k0 + k):
        envs[i % B].step(acts[i % R], out=outs[i % B])
def timeit(fn, reps):
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); fn(reps); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
run(0, 200)
ms = timeit(lambda k: run(0, k), steps)
print(f"loop  n={n} V={V} v{ver}: {ms/steps*1e3:.2f} us/launch  {n*steps/ms/1e6:.3f} Gsteps/s")
# graph of R*B... launches: lcm(B,R)=16 steps per replay
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    run(0, 16)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        run(0, 16)
def rg(k):
    for _ in range(k // 16): g.replay()
rg(160)
ms = timeit(rg, steps)
k = steps // 16 * 16
print(f"graph n={n} V={V} v{ver}: {ms/k*1e3:.2f} us/launch  {n*k/ms/1e6:.3f} Gsteps/s")
