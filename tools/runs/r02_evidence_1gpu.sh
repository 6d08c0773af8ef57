#!/bin/bash
# single-GPU evidence session of round 2: launch list of one N=40k solve, ncu --set full of the round-2 kernels,
# LAPACK vs device at the bench size, the bench line, smoke
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
python __graft_entry__.py --smoke > $O/r02_smoke.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file $O/r02_launches_N40k_one_solve.csv python bench.py --steps 1 --warmup 0 --skip_cpu_baseline > $O/r02_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"trsv_persistent|potrf_tiled|trsm_panel|hess_kernel|hess_blocks|predict_kernel|grad_kernel" -c 14 -o $O/r02_prof_round2_kernels python tools/ncu_targets.py > $O/r02_ncu_targets.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_dmma" -s 200 -c 3 -o $O/r02_prof_task_gemm python tools/ncu_targets.py >> $O/r02_ncu_targets.log 2>&1
python tools/nugget_probe.py --N 20000 --nuggets 1e-13 --variants > $O/r02_nugget_variants_N20k.jsonl 2> $O/r02_nugget2.err
python tools/nugget_probe.py --N 40000 --nuggets 1e-13 --lapack_max 40000 > $O/r02_nugget_N40k.jsonl 2>> $O/r02_nugget2.err
python bench.py --steps 2 --warmup 1 > $O/r02_bench_1gpu.log 2> $O/r02_bench_1gpu.err
tail -2 $O/r02_smoke.log; tail -2 $O/r02_ncu_targets.log; cat $O/r02_nugget_variants_N20k.jsonl $O/r02_nugget_N40k.jsonl; cut -c1-600 $O/r02_bench_1gpu.log; ls -la $O | tail -12
