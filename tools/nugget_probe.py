"""Smallest nugget for which the FP64 Cholesky of Theta succeeds, GPU vs LAPACK (developer tool)."""
import argparse, sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import PDEs
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, nargs="+", default=[5000])
ap.add_argument("--nuggets", type=float, nargs="+", default=[1e-13, 1e-12, 1e-11, 1e-10])
ap.add_argument("--lapack_max", type=int, default=6000)
a = ap.parse_args()
for N in a.N:
    Nb = 4 * (math.ceil(math.sqrt(N)) + 1)
    np.random.seed(0)
    p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: 0 * x, rhs=lambda x, y: 0 * x)
    p.sampled_pts(N, Nb)
    for ng in a.nuggets:
        p.Gram_matrix("Gaussian", 0.2, ng, "adaptive")
        lap = None
        if N <= a.lapack_max:
            th = p.Theta
            try:
                np.linalg.cholesky(th); lap = 0
            except np.linalg.LinAlgError:
                lap = -1
        p.Gram_Cholesky()
        print(f"N={N} nugget={ng:g} gpu_info={p.chol_info} lapack={'ok' if lap == 0 else ('fail' if lap == -1 else 'n/a')}", flush=True)
