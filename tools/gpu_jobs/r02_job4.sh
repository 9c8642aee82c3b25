# round 2, job 4: GPU suite after the test / harness fixes (full log kept), quick bench
timeout 1200 python -m pytest tests -m gpu -q --tb=short > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
grep -n "^FAILED\|^ERROR" gpurun_out/r02_pytest_gpu.log | head -20
grep -n "test_raster_vs_reference" -A25 gpurun_out/r02_pytest_gpu.log | grep -v "^--$" | head -60
cp gpurun_out/refsuite_b200.txt gpurun_out/r02_refsuite_b200.txt 2>/dev/null
cp gpurun_out/refsuite_b200_dropin_stats.json gpurun_out/r02_refsuite_dropin_stats.json 2>/dev/null
head -1 gpurun_out/r02_refsuite_b200.txt
python bench.py --steps 20 --warmup 5 --no-trainer > gpurun_out/r02_v2_bench.json 2> gpurun_out/r02_v2_bench.err; python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v2_bench.json").read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "cpu", d["cpu_baseline"])
PY
python bench.py --impl reference --steps 20 --warmup 5 | tail -c 700
# forward with the barrier-free ring (opt-in until validated): parity subset, then timing
GSB200_FWD_PIPE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "raster or rasterization" 2>&1 | tail -4
GSB200_FWD_PIPE=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_v2_bench_fwdpipe.json 2> gpurun_out/r02_v2_bench_fwdpipe.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_v2_bench_fwdpipe.json").read().strip().splitlines()[-1])
    print("FWD_PIPE ms/step", round(d["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "big_s", round(d["big_s"]["ms_per_step"],3), round(d["big_s"]["fwd_only_ms"],3))
except Exception as e: print("FWD_PIPE ERR", e, open("gpurun_out/r02_v2_bench_fwdpipe.err").read()[-600:])
PY
