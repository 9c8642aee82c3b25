# round 2, job 15 (1 GPU): the state that is handed in -- full GPU suite, the default bench line (trainer arms, CPU port),
# launch list and --set full capture of the six own kernels of a step
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu.log | head
cp gpurun_out/refsuite_b200.txt gpurun_out/r02_refsuite_b200.txt 2>/dev/null; cp gpurun_out/refsuite_b200_dropin_stats.json gpurun_out/r02_refsuite_dropin_stats.json 2>/dev/null; head -1 gpurun_out/r02_refsuite_b200.txt
python bench.py > gpurun_out/r02_v7_bench.json 2> gpurun_out/r02_v7_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v7_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "host", d.get("host_issue_ms_per_step"), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "stock", d["ref_cuda_stock"]["ms_per_step"], "big_s", d["big_s"]["ms_per_step"], d["big_s"]["fwd_only_ms"])
t=d.get("trainer") or {}
for k,v in t.items(): print(k, json.dumps(v)[:600])
print("cpu", d["cpu_baseline"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v7_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r02_v7_launches.csv
ncu --set full --clock-control none --import-source on -k regex:"raster_bwd2|raster_fwd|project_sh|pack_rows|isect_emit" -s 12 -c 6 -o gpurun_out/r02_v7_kernels -f python tools/profile_step.py --steps 1 --warmup 2 > gpurun_out/ncu_full.log 2>&1; ls -la gpurun_out/r02_v7_kernels.ncu-rep
