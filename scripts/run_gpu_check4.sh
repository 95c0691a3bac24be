python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
python scripts/aux_kernels_bench.py > gpurun_out/aux.json 2> gpurun_out/aux.err; echo rc=$?; cat gpurun_out/aux.json; tail -3 gpurun_out/aux.err
TA_GAE_L=8 python scripts/aux_kernels_bench.py 2>&1 | tail -1
