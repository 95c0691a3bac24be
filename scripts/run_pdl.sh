python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for pdl in 0 1; do for v in 17 7; do echo "PDL=$pdl V=$v"; TA_PDL=$pdl V=$v python scripts/step_time.py 2>&1 | tail -1; done; done
TA_PDL=1 TA_DEBUG_FLAGS=3 python scripts/step_time.py 2>&1 | tail -1
