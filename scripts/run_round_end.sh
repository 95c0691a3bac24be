# round-end evidence refresh: aux-kernel ncu list, PPO-step ncu launch list, smoke
bash scripts/run_prof_aux.sh
bash scripts/run_prof_ppo_step.sh
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
