# round 2, job 6 (2 GPUs): all-reduce kernels (dense + flattened row-sparse) vs NCCL with the full log, DP bench
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_nvls_check.py > gpurun_out/r02_nvls_check_n2.log 2>&1
grep -n "nvls check ok\|Error\|error\|assert\|Traceback" gpurun_out/r02_nvls_check_n2.log | head -30
grep -n "Traceback" -A25 gpurun_out/r02_nvls_check_n2.log | head -60
unset NCCL_DEBUG
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_v3_bench_n2.json 2> gpurun_out/r02_v3_bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v3_bench_n2.json").read().strip().splitlines()[-1])
print("N=2 value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3))
dp=d["dp"]; print({k:dp[k] for k in dp if k!="gaussian_sharded"}); print(dp["gaussian_sharded"]["dense"], dp["gaussian_sharded"]["packed"])
PY
tail -3 gpurun_out/r02_v3_bench_n2.err | cut -c1-300
