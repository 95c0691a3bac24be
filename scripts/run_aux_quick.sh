# featuriser parity tests + the aux bench table (us, fraction of the measured HBM peak)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "stack or push or roll or feat or matrix or rollout" 2>&1 | tail -3
for cfg in "TA_FEAT_CTAS=4" "TA_FEAT_CTAS=2" "TA_FEAT_CTAS=3" "TA_FEAT_CTAS=6" "TA_FEAT_CTAS=7"; do
env $cfg timeout 200 python bench.py --workload aux 2>/dev/null | tail -1 > gpurun_out/aux_new.json
python -c "
import json; d=json.load(open('gpurun_out/aux_new.json')); ks=d.get('kernels', d)
print('$cfg', {k:(round(v['us'],1), round(v['frac'],3)) for k,v in ks.items() if isinstance(v,dict) and 'us' in v and 'state' in k})"
done
