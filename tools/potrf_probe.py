"""Assemble + factorise once at a given size, left-looking (default) or right-looking task-list schedule (--rl): timing probe and
ncu target for the two big-GEMM shapes (long-K column update vs K = NB trailing update)."""
import argparse, json, math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import PDEs

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=20000)
ap.add_argument("--nugget", type=float, default=1e-10)
ap.add_argument("--rl", action="store_true")
ap.add_argument("--blocksum", type=int, default=1)
ap.add_argument("--reps", type=int, default=1)
a = ap.parse_args()
np.random.seed(0)
p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: 0 * x, rhs=lambda x, y: 0 * x)
p.sampled_pts(a.N, 4 * (math.ceil(math.sqrt(a.N)) + 1))
eng = p._engine()
eng.set_option("rl_potrf", 1 if a.rl else 0)
eng.set_option("blocksum", a.blocksum)
M = 2 * a.N + p.N_boundary
for rep in range(a.reps):
    p.Gram_matrix("Gaussian", 0.2, a.nugget, "adaptive")
    p._state = 'gram'
    eng.timer_start()
    info = eng.potrf(0)
    ms = eng.timer_stop()
    print(json.dumps(dict(N=a.N, M=M, schedule="right-looking" if a.rl else "left-looking", blocksum=a.blocksum, rep=rep, info=info,
                          potrf_ms=round(ms, 2), TFLOPs=round(M ** 3 / 3 / ms / 1e9, 2))), flush=True)
