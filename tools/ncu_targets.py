"""Small driver that launches every round-2 kernel a handful of times, for `ncu --set full -k regex:...` captures:
  persistent vector solve, tiled Cholesky, Hessian / gradient / prediction kernels (elliptic N_domain = 10 000, C2-size Burgers),
  task-list GEMM, fused panel solve, block Hessian (sharded path, 4 virtual ranks)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import PDEs
from oracle import gp_oracle as o

np.random.seed(0)
N = 10000
p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f)
p.sampled_pts(N, 4 * (math.ceil(math.sqrt(N)) + 1))
p.Gram_matrix("Gaussian", 0.2, 1e-10, "adaptive"); p.Gram_Cholesky(); p.GN_method(1, 1, "rdm", print_hist=False)
p.extend_sol(np.random.uniform(0, 1, (2000, 2)))
b = PDEs.Burgers(alpha=1.0, nu=0.02, bdy=o.burgers_bdy, rhs=lambda t, x: 0)
b.sampled_pts(1000, 200); b.Gram_matrix("anisotropic_Gaussian", (0.3, 0.05), 1e-5); b.Gram_Cholesky(); b.GN_method(1, 1, "rdm", print_hist=False)
s = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f)
s.sampled_pts(4000, 260); s.shard(virtual_ranks=4)
s.Gram_matrix("Gaussian", 0.2, 1e-9, "adaptive"); s.Gram_Cholesky(); s.GN_method(1, 1, "rdm", print_hist=False)
print("ncu targets done", p.loss_hist[-1], b.loss_hist[-1], s.loss_hist[-1])
