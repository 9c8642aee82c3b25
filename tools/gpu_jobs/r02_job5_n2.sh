# round 2, job 5 (2 GPUs): own all-reduce kernels (dense + row-sparse) vs NCCL, gaussian-sharded check, DP bench with checks
export NCCL_DEBUG=WARN
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/dist_nvls_check.py 2>&1 | grep -v "^W09\|^\*\*\*\*\|^$" | tail -25
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/dist_sharded_check.py 2>&1 | grep -v "^W09\|^\*\*\*\*\|^$" | tail -8
unset NCCL_DEBUG
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_v2_bench_n2.json 2> gpurun_out/r02_v2_bench_n2.err
tail -c 3500 gpurun_out/r02_v2_bench_n2.json; tail -5 gpurun_out/r02_v2_bench_n2.err | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29614 bench.py --impl reference --gpus 2 --steps 5 --warmup 1 | tail -c 600
