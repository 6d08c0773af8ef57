"""Phase timing of the elliptic hot path at a given size (developer tool; bench.py is the contract)."""
import argparse, json, math, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import PDEs


def u(x1, x2):
    return np.sin(np.pi * x1) * np.sin(np.pi * x2) + 2 * np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)


def f(x1, x2):
    s1 = np.sin(np.pi * x1) * np.sin(np.pi * x2); s4 = np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)
    w = s1 + 2 * s4
    return 2 * np.pi ** 2 * s1 + 64 * np.pi ** 2 * s4 + w * w * w


ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=10000)
ap.add_argument("--nugget", type=float, default=1e-10)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--sigma", type=float, default=0.2)
a = ap.parse_args()
N = a.N
Nb = 4 * (math.ceil(math.sqrt(N)) + 1)
np.random.seed(0)
p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=u, rhs=f)
p.sampled_pts(N, Nb)
init = np.random.normal(0, 1, N)
M = 2 * N + p.N_boundary
for rep in range(a.reps):
    t0 = time.time()
    p.Gram_matrix("Gaussian", a.sigma, a.nugget, "adaptive")
    p.Gram_Cholesky()
    p.GN_method(a.steps, 1, init, print_hist=False)
    wall = time.time() - t0
    t = p.timings
    err = np.abs(u(p.X_domain[:, 0], p.X_domain[:, 1]) - p.sol_sampled_pts)
    print(json.dumps(dict(N=N, M=M, rep=rep, wall_s=round(wall, 3), info=p.chol_info,
                          asm_ms=round(t["assembly_ms"], 3), asm_GBs=round(8 * M * (M + 1) / 2 / t["assembly_ms"] / 1e6, 1),
                          potrf_ms=round(t["potrf_ms"], 2), potrf_TF=round(M ** 3 / 3 / t["potrf_ms"] / 1e9, 2),
                          inv_ms=round(t["inverse_ms"], 2), inv_TF=round(2 * M ** 3 / 3 / t["inverse_ms"] / 1e9, 2),
                          gn_ms=round(t["gn_ms"], 2), gn_ms_per_step=round(t["gn_ms"] / a.steps, 2),
                          loss=p.loss_hist, L2=float(np.sqrt(np.mean(err ** 2))), max=float(err.max()))))
