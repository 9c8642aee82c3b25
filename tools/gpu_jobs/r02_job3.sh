# round 2, job 3: full GPU suite (incl. reference suite + registry binding), A/B of the backward variants, ncu launch list +
# full captures, bench with the cfg5 trainer runs
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60
cp gpurun_out/refsuite_b200.txt gpurun_out/r02_refsuite_b200.txt 2>/dev/null
cp gpurun_out/refsuite_b200_dropin_stats.json gpurun_out/r02_refsuite_dropin_stats.json 2>/dev/null
for v in "GSB200_BWD_PIPE=0" "GSB200_BWD_PIPE=1" "GSB200_BWD_ALGO=butterfly"; do
  env $v python bench.py --steps 20 --warmup 5 --no-trainer --no-cpu-baseline > gpurun_out/r02_v1_ab_${v//=/_}.json 2> gpurun_out/r02_v1_ab_${v//=/_}.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_v1_ab_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("ab_")[1], "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "bwd", round(d["roofline"]["ms"],4), "fwd", round(d["roofline"]["raster_fwd"]["ms"],4), "ref_stock", (d.get("ref_cuda_stock") or {}).get("ms_per_step"), "big_s", {k:(round(v,3) if isinstance(v,float) else v) for k,v in (d.get("big_s") or {}).items() if k!="workload"})
    except Exception as e: print(f, "ERR", e, open(f.replace(".json",".err")).read()[-800:])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_v1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/r02_v1_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_bwd2 -s 4 -c 1 -o gpurun_out/r02_v1_bwd2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_fwd -s 4 -c 1 -o gpurun_out/r02_v1_fwd -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --trainer-steps 700 > gpurun_out/r02_v1_bench.json 2> gpurun_out/r02_v1_bench.err
tail -c 7000 gpurun_out/r02_v1_bench.json; tail -5 gpurun_out/r02_v1_bench.err
