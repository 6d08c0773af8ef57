"""Small driver that launches every round-2 kernel a handful of times, for `ncu --set full -k regex:...` captures:
  sharded path first (task-list GEMM, fused panel solve, block Hessian; 2 virtual ranks, N_domain = 1500), then the
  single-GPU elliptic solve at N_domain = 10 000 (persistent vector solves at M = 20 404, tiled Cholesky of the diagonal
  blocks, Hessian / gradient / prediction kernels)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import PDEs
from oracle import gp_oracle as o

np.random.seed(0)
s = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f)
s.sampled_pts(1500, 160); s.shard(virtual_ranks=2)
s.Gram_matrix("Gaussian", 0.2, 1e-9, "adaptive"); s.Gram_Cholesky(); s.GN_method(1, 1, "rdm", print_hist=False)
N = 10000
p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f)
p.sampled_pts(N, 4 * (math.ceil(math.sqrt(N)) + 1))
p.Gram_matrix("Gaussian", 0.2, 1e-10, "adaptive"); p.Gram_Cholesky(); p.GN_method(1, 1, "rdm", print_hist=False)
p.extend_sol(np.random.uniform(0, 1, (2000, 2)))
print("ncu targets done", p.loss_hist[-1], s.loss_hist[-1])
