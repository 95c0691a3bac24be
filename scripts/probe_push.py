"""stack_push_codes_tile_kernel at 65536 envs: time per launch (CUDA graph of 48 nodes over 8 rotating batches) for
one setting of the TA_PUSH_* knobs (read once per process, so run one process per setting)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import twoarmy_b200 as pkg
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from aux_kernels_bench import timeit

dev = torch.device("cuda:0")
n, B = 65536, 8
envs = [pkg.TwoarmyVecEnv(4, n, 17, device=dev, seed=1, env_id0=b * n) for b in range(B)]
for e in envs:
    e.reset()
sc = [torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev) for _ in range(B)]
so = [torch.zeros((n, 5, 289), dtype=torch.uint8, device=dev) for _ in range(B)]
po = torch.zeros((n, 5, 2), dtype=torch.float32, device=dev)
ps = [torch.zeros((n, 5, 2), dtype=torch.float32, device=dev) for _ in range(B)]
k = [0]


def push():
    b = k[0] % B; k[0] += 1
    envs[b].stack_push(sc[b], so[b], ps[b], po)


s = timeit(push)
byt = (80 + 16 + 4 * 289 + 5 * 289 + 2 * 40) * n
print(json.dumps({"knobs": {k_: v for k_, v in os.environ.items() if k_.startswith("TA_PUSH")}, "us": round(s * 1e6, 2),
                  "frac": round(byt / s / 1e9 / 6536.7, 3)}))
