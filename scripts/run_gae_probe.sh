mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gae.py -q -m gpu -x 2>&1 | tail -3
timeout 120 python scripts/probe_aux_floor.py 2>/dev/null | tail -1 | tee gpurun_out/gae_probe.txt
timeout 300 python bench.py --workload aux 2>/dev/null | tail -1 > gpurun_out/aux_new.json
python -c "
import json; d=json.load(open('gpurun_out/aux_new.json')); ks=d.get('kernels', d)
print({k:(round(v['us'],1), round(v['frac'],3)) for k,v in ks.items() if isinstance(v,dict) and 'us' in v})"
