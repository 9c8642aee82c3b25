# round 2, job 20 (1 GPU): launch list of the last build, smoke()
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v9_launches.csv python tools/profile_step.py --steps 2 --warmup 3 > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r02_v9_launches.csv
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
