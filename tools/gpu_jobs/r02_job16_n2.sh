# round 2, job 16 (2 GPUs): final build -- parity incl. the 2-GPU tests, all-reduce check, DP bench line
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --tb=short > gpurun_out/r02_pytest_parity_n2.log 2>&1; tail -2 gpurun_out/r02_pytest_parity_n2.log
export NCCL_DEBUG=WARN GSB200_CHECK_FAST=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_nvls_check.py > gpurun_out/r02_nvls_check_n2.log 2>&1
grep -n "nvls check ok\|AssertionError" gpurun_out/r02_nvls_check_n2.log | cut -c1-200 | tail -6
unset NCCL_DEBUG GSB200_CHECK_FAST
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_v7_bench_n2.json 2> gpurun_out/r02_v7_bench_n2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v7_bench_n2.json").read().strip().splitlines()[-1])
dp=d["dp"]; print("N=2 ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), "value", round(d["value"],1)); print({k:dp[k] for k in dp if k!="gaussian_sharded"}); print(dp["gaussian_sharded"]["dense"], dp["gaussian_sharded"]["packed"])
PY
