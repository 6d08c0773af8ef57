#!/bin/bash
# last 8-GPU measurement of round 2: bench line with the persistent task-list GEMM
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
export NCCL_DEBUG=WARN
timeout 300 python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 2 --warmup 1 > $O/r02_bench_8gpu_final2.log 2> $O/r02_bench_8gpu_final2.err
grep "self-check" $O/r02_bench_8gpu_final2.err; grep '^{' $O/r02_bench_8gpu_final2.log | cut -c1-260
