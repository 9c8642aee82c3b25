# round 2, job 9 (8 GPUs): DP bench with checks -- the SCALE configuration
# (the all-reduce kernels are checked against NCCL inside bench.py: dp.checks)
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_v7_bench_n8.json 2> gpurun_out/r02_v7_bench_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v7_bench_n8.json").read().strip().splitlines()[-1])
dp=d["dp"]; print("N=8 value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3)); print({k:dp[k] for k in dp if k!="gaussian_sharded"}); print(dp["gaussian_sharded"]["dense"], dp["gaussian_sharded"]["packed"])
PY
tail -3 gpurun_out/r02_v7_bench_n8.err | cut -c1-300
