# development sweep: does leaving SM resources free let consecutive ta_step launches (PDL) overlap more?
for cfg in "0 0" "1 7" "1 8" "1 4" "2 4" "2 3" "3 2" "1 6"; do set -- $cfg
  TA_CTAS_PER_SM=$1 TA_WARPS_PER_CTA=$2 timeout 120 python bench.py --no-ppo --no-extra --no-cpu-baseline --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ctas_per_sm=$1 warps_per_cta=$2', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],3))"
done
