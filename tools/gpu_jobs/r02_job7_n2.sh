# round 2, job 7 (2 GPUs): nvls check after the fix; is the arena (symmetric memory) what makes the DP compute slower?
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_nvls_check.py > gpurun_out/r02_nvls_check_n2.log 2>&1
grep -n "nvls check ok\|AssertionError" gpurun_out/r02_nvls_check_n2.log | cut -c1-200 | tail -6
unset NCCL_DEBUG
for v in "GSB200_ALLREDUCE=own" "GSB200_ALLREDUCE=nccl" "GSB200_ALLREDUCE_ALGO=nvls" "GSB200_ALLREDUCE_ROWS=0"; do
env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_v3_bench_n2_${v//=/_}.json 2> /dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_v3_bench_n2_${v//=/_}.json").read().strip().splitlines()[-1])
dp=d["dp"]; print("$v", "ms/step", round(d["ms_per_step"],3), "compute", dp["per_rank_compute_ms"], "allreduce", dp["per_rank_allreduce_ms"], dp["allreduce"], "exposed", dp["exposed_allreduce_ms"], "moved", dp["checks"].get("allreduce_payload_fraction_moved"))
PY
done
