python -m pytest tests -m gpu -x -q 2>&1 | tail -8
GSB200_BWD_ALGO=butterfly python bench.py --steps 20 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_bench_bwd1.json 2> gpurun_out/r02_bench_bwd1.err
python bench.py --steps 20 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_bench_bwd2.json 2> gpurun_out/r02_bench_bwd2.err
GSB200_EMIT=serial python bench.py --steps 20 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_bench_emitserial.json 2> gpurun_out/r02_bench_emitserial.err
python - <<PY
import json
for n in ("bwd1","bwd2","emitserial"):
    try:
        d=json.loads(open(f"gpurun_out/r02_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "ref_stock", (d.get("ref_cuda_stock") or {}).get("ms_per_step"), "ref_chain", (d.get("ref_cuda") or {}).get("ms_per_step"), "big_s", d.get("big_s"))
    except Exception as e: print(n, "ERR", e, open(f"gpurun_out/r02_bench_{n}.err").read()[-1500:])
PY
python tools/trainer_bench.py --backend b200 --strategy default --steps 160 --grad-stats > gpurun_out/r02_gradstats.json 2>gpurun_out/r02_gradstats.err; tail -c 700 gpurun_out/r02_gradstats.json; tail -3 gpurun_out/r02_gradstats.err
python tools/run_refsuite.py --backend b200 --timeout 1200 2>&1 | tail -3
