#!/bin/bash
# 8-GPU session of round 2: sharded N=40k solve on 8 / 4 GPUs, grid shapes, block sizes, then the bench line.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export NCCL_DEBUG=WARN
GPP_TRACE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29701 tools/dist_solve.py --N 40000 --NB 512 --nugget 1e-12 --reps 2 > $O/r02_dist8_N40k_NB512.log 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29702 tools/dist_solve.py --N 40000 --NB 512 --nugget 1e-12 --reps 2 --Q 2 > $O/r02_dist8_N40k_NB512_Q2.log 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29703 tools/dist_solve.py --N 40000 --NB 1024 --nugget 1e-12 --reps 2 > $O/r02_dist8_N40k_NB1024.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29704 tools/dist_solve.py --N 40000 --NB 512 --nugget 1e-12 --reps 2 > $O/r02_dist4_N40k_NB512.log 2>&1
NCCL_DEBUG=INFO timeout 400 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 2 --warmup 1 > $O/r02_bench_8gpu.log 2> $O/r02_bench_8gpu.err
grep -m 20 -i "nvls\|Connected all\|channels" $O/r02_bench_8gpu.err > $O/r02_nccl_info_8gpu.txt
for f in r02_dist8_N40k_NB512 r02_dist8_N40k_NB512_Q2 r02_dist8_N40k_NB1024 r02_dist4_N40k_NB512; do echo "== $f"; grep '^{' $O/$f.log | tail -1; grep "dist inverse (rank 0)" $O/$f.log | tail -1; done
grep "gn_step" $O/r02_dist8_N40k_NB512.log | tail -1
grep "self-check" $O/r02_bench_8gpu.err; grep '^{' $O/r02_bench_8gpu.log | cut -c1-1500
