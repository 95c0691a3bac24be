# featuriser parity tests + the aux bench table (us, fraction of the measured HBM peak)
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stack or push or roll or feat or matrix" 2>&1 | tail -2
timeout 200 python bench.py --workload aux 2>/dev/null | tail -1 > gpurun_out/aux_new.json
python -c "
import json; d=json.load(open('gpurun_out/aux_new.json')); ks=d.get('kernels', d)
print({k:(round(v['us'],1), round(v['frac'],3)) for k,v in ks.items() if isinstance(v,dict) and 'us' in v})"
