# round 2, job 19 (1 GPU): totals published by a kernel into pinned memory (no copy engine) -- full GPU suite + bench line
timeout 1500 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log; grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu.log | head
cp gpurun_out/refsuite_b200.txt gpurun_out/r02_refsuite_b200.txt 2>/dev/null; head -1 gpurun_out/r02_refsuite_b200.txt
python bench.py --no-trainer --no-cpu-baseline > gpurun_out/r02_v9_bench.json 2> gpurun_out/r02_v9_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v9_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "host", d.get("host_issue_ms_per_step"), "e2e", round(d["e2e"]["ms_per_step"],3), d["e2e"].get("host_issue_ms_per_step"), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "stock", d["ref_cuda_stock"]["ms_per_step"], "big_s", d["big_s"]["ms_per_step"], d["big_s"]["fwd_only_ms"])
PY
