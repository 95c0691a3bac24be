for v in 1 0; do echo "TA_CONV2_DGRAD_TC=$v"; TA_CONV2_DGRAD_TC=$v timeout 300 python scripts/probe_eager_step.py 2>&1 | tail -4; done
