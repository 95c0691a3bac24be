# featuriser parity tests + the aux bench table (us, fraction of the measured HBM peak) + the harness floor probe
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stack or push or roll or feat or matrix or rollout" 2>&1 | tail -2
for mb in 5 6 8; do
TA_PUSH_MINB=$mb timeout 200 python bench.py --workload aux 2>/dev/null | tail -1 > gpurun_out/aux_new_$mb.json
python -c "
import json; d=json.load(open('gpurun_out/aux_new_$mb.json')); ks=d.get('kernels', d)
print($mb, {k:(round(v['us'],1), round(v['frac'],3)) for k,v in ks.items() if isinstance(v,dict) and 'us' in v})"
done
timeout 200 python scripts/probe_aux_floor.py 2>&1 | tail -1 | tee gpurun_out/aux_floor.json
