import sys, time; sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, twoarmy_b200 as pkg
dev=torch.device('cuda:0')
import os
n=int(os.environ.get('N','65536')); V=int(os.environ.get('V','17')); T=int(os.environ.get('T','16'))
env=pkg.TwoarmyVecEnv(4,n,V,device=dev,seed=1); env.reset()
amap=torch.tensor([0,1,2,3,6],dtype=torch.uint8,device=dev)
acts=amap[torch.randint(0,5,(T,n),device=dev)].contiguous()
import ctypes as C
from twoarmy_b200 import _capi
L=_capi.lib()
obs=torch.empty((T,n,V,V,3),dtype=torch.uint8,device=dev); rew=torch.empty((T,n),device=dev); te=torch.empty((T,n),dtype=torch.uint8,device=dev); tr=torch.empty((T,n),dtype=torch.uint8,device=dev)
def go():
    _capi.check(L.ta_rollout(env._h, C.c_void_p(acts.data_ptr()), 1, T, C.c_void_p(obs.data_ptr()), C.c_void_p(rew.data_ptr()), C.c_void_p(te.data_ptr()), C.c_void_p(tr.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(5): go()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
R=40
e0.record()
for _ in range(R): go()
e1.record(); torch.cuda.synchronize()
ms=e0.elapsed_time(e1)/R
print(f"n={n} V={V} T={T}: {ms*1e3:.1f} us/launch, {ms*1e3/T:.2f} us/step, {n*T/ms/1e6:.3f} Gsteps/s, done_frac={float((te|tr).float().mean()):.4f}")
