#!/bin/bash
# final 8-GPU session of round 2: the bench line of the final code on 8 and on 4 GPUs (default NCCL exchange, fused panel solve)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export NCCL_DEBUG=WARN
timeout 400 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 3 --warmup 2 > $O/r02_bench_8gpu_final.log 2> $O/r02_bench_8gpu_final.err
timeout 400 $TR --nproc-per-node 4 --master-port 29706 bench.py --gpus 4 --steps 2 --warmup 1 > $O/r02_bench_4gpu_final.log 2> $O/r02_bench_4gpu_final.err
grep "self-check" $O/r02_bench_8gpu_final.err; grep '^{' $O/r02_bench_8gpu_final.log | cut -c1-260; grep '^{' $O/r02_bench_4gpu_final.log | cut -c1-260
