# round 2, job 3: full GPU suite (incl. reference suite + registry binding), ncu launch list + full captures, bench, trainer
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60
cp gpurun_out/refsuite_b200.txt gpurun_out/r02_refsuite_b200.txt 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_v1_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > gpurun_out/r02_v1_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_bwd2 -s 4 -c 1 -o gpurun_out/r02_v1_bwd2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:raster_fwd -s 4 -c 1 -o gpurun_out/r02_v1_fwd -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-trainer > /dev/null 2>&1
python bench.py --steps 20 --warmup 5 --trainer-steps 700 > gpurun_out/r02_v1_bench.json 2> gpurun_out/r02_v1_bench.err
tail -c 6000 gpurun_out/r02_v1_bench.json; tail -5 gpurun_out/r02_v1_bench.err
