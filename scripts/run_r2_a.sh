# round 2, GPU run A: host topology probe, full GPU tests, smoke, default bench line (with extra.ppo), reference arm
mkdir -p gpurun_out
{ lscpu | head -30; echo; nproc; cat /sys/devices/system/node/node*/cpulist 2>/dev/null; nvidia-smi topo -m; free -g | head -3; 
  cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c; } > gpurun_out/r2_host_probe.txt 2>&1
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2a_pytest.txt; tail -3 gpurun_out/r2a_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2a_bench.json'))
    print('value',d['value'],'us',d['roofline']['launch_us'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'])
    print('view7',d['extra']['view7'])
    p=d['extra']['ppo']; print('ppo',p['value'],p['ms_per_optimizer_step'],p.get('launches_per_optimizer_step'),p['phases_ms'])
    print('cpu',d['cpu_baseline']['value'],d['cpu_baseline']['port']['value'], p.get('cpu_baseline'))
except Exception as e: print('parse fail',e)
PY
