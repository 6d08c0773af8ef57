#!/bin/bash
# second 8-GPU session of round 2: peer-store exchange vs NCCL exchange, fused panel solve, then the bench line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
export NCCL_DEBUG=WARN
GPP_TRACE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29701 tools/dist_solve.py --N 40000 --NB 512 --nugget 1e-12 --reps 2 > $O/r02b_dist8_p2p.log 2>&1
GPP_DIST_P2P=0 GPP_TRACE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29702 tools/dist_solve.py --N 40000 --NB 512 --nugget 1e-12 --reps 2 > $O/r02b_dist8_nccl.log 2>&1
timeout 300 $TR --nproc-per-node 4 --master-port 29704 tools/dist_solve.py --N 40000 --NB 512 --nugget 1e-12 --reps 2 > $O/r02b_dist4_p2p.log 2>&1
timeout 400 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 2 --warmup 1 > $O/r02b_bench_8gpu.log 2> $O/r02b_bench_8gpu.err
for f in r02b_dist8_p2p r02b_dist8_nccl r02b_dist4_p2p; do echo "== $f"; grep '^{' $O/$f.log | tail -1; grep "dist inverse (rank 0)" $O/$f.log | tail -1; grep "gn_step" $O/$f.log | tail -1; done
grep "panel chain" $O/r02b_dist8_p2p.log | tail -14
grep "self-check" $O/r02b_bench_8gpu.err; grep '^{' $O/r02b_bench_8gpu.log | cut -c1-900
