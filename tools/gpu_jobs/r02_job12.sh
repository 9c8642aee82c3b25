# round 2, job 12 (1 GPU): fused projection side outputs (tile counts + compositing row records), project_sh_bwd at
# 3 CTAs / SM for one camera -- parity, A/B, launch list
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_vs_reference_cuda.py -m gpu -q -x --tb=short 2>&1 | tail -5
run() { # name, env...
  name=$1; shift
  env "$@" python bench.py --steps 30 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_v5_bench_$name.json 2> gpurun_out/r02_v5_bench_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v5_bench_$name.json").read().strip().splitlines()[-1])
print("$name", "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "stock", round(d["ref_cuda_stock"]["ms_per_step"],3), "big_s", round(d["big_s"]["ms_per_step"],3), round(d["big_s"]["fwd_only_ms"],3))
PY
}
run rows GSB200_X=0
run norows GSB200_ROW_RECORDS=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v5_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r02_v5_launches.csv
