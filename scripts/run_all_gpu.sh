# the whole GPU suite + the aux bench table
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tail -8
timeout 300 python bench.py --workload aux 2>/dev/null | tail -1 > gpurun_out/aux_new.json
python -c "
import json; d=json.load(open('gpurun_out/aux_new.json')); ks=d.get('kernels', d)
print({k:(round(v['us'],1), round(v['frac'],3)) for k,v in ks.items() if isinstance(v,dict) and 'us' in v})"
