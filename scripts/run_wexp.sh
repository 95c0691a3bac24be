for w in 1 2 4 7 8; do echo "W=$w"; TA_WARPS_PER_CTA=$w python scripts/step_time.py 2>&1 | tail -1; done
for c in 5 7 10; do echo "W=1 CTAS_PER_SM=$c"; TA_CTAS_PER_SM=$c python scripts/step_time.py 2>&1 | tail -1; done
echo "W=2 V=7"; TA_WARPS_PER_CTA=2 V=7 python scripts/step_time.py 2>&1 | tail -1
echo "W=7 V=7"; TA_WARPS_PER_CTA=7 V=7 python scripts/step_time.py 2>&1 | tail -1
echo "W=7 nothing"; TA_WARPS_PER_CTA=7 TA_DEBUG_FLAGS=3 python scripts/step_time.py 2>&1 | tail -1
echo "W=7 rollout"; TA_WARPS_PER_CTA=7 python scripts/rollout_time.py 2>&1 | tail -1
