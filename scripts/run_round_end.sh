# round-end evidence refresh: full GPU tests, smoke, aux-kernel ncu list, PPO-step ncu launch list, one full capture of the
# tcgen05 weight-gradient kernel (planes + bit mask)
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash scripts/run_prof_aux.sh
bash scripts/run_prof_ppo_step.sh
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"conv1_bwd_tc" -s 30 -c 1 -o gpurun_out/conv1_bwd_tc_final -f python scripts/probe_conv1_bwd_tc.py > gpurun_out/ncu_bwd_final.log 2>&1; tail -1 gpurun_out/ncu_bwd_final.log
