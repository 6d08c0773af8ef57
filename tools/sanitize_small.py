"""Small end-to-end run of every PDE for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nonlinpdes_gpsolver_b200 import PDEs, InverseProblems
np.random.seed(0)
p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: np.sin(x + y), rhs=lambda x, y: x * y)
p.sampled_pts(300, 44); p.Gram_matrix("Gaussian", 0.2, 1e-6); p.Gram_Cholesky(); p.GN_method(2, 1, "rdm", print_hist=False)
p.extend_sol(np.random.uniform(0, 1, (50, 2)))
p.Gram_matrix("Gaussian", 0.2, 1e-6); p.Gram_Cholesky(); p.GN_relaxed_method(2, 1, "rdm", pen_lambda=1e-6, print_hist=False)
b = PDEs.Burgers(alpha=1.0, nu=0.02, bdy=lambda t, x: -np.sin(np.pi * x) * (t == 0), rhs=lambda t, x: 0)
b.sampled_pts(200, 60); b.Gram_matrix("anisotropic_Gaussian", (0.3, 0.05), 1e-5); b.Gram_Cholesky(); b.GN_method(2, 1, "rdm", print_hist=False)
e = PDEs.Eikonal(eps=0.1, bdy=lambda x, y: 0, rhs=lambda x, y: 1)
e.sampled_pts(700, 100); e.Gram_matrix("Gaussian", 0.2, 1e-5); e.Gram_Cholesky(); e.GN_method(2, 1, "zero", print_hist=False)   # M=2900: look-ahead path
d = InverseProblems.Darcy_flow2d(bdy=lambda x, y: 0, rhs=lambda x, y: 1)
d.sampled_pts(150, 40, 20); d.get_observation(np.zeros(20), 1e-2); d.Gram_matrix("Gaussian", 0.2, 1e-6); d.Gram_Cholesky(); d.GN_method(2, 1, "rdm", print_hist=False)
d.extend_sol(np.random.uniform(0, 1, (30, 2)))
# the sharded path (task-list GEMM, fused panel solve, block Hessian) with 4 virtual ranks on a 2 x 2 grid
sh = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: np.sin(x + y), rhs=lambda x, y: x * y)
sh._engine().set_option("NB", 128)
sh.sampled_pts(500, 60); sh.shard(virtual_ranks=4, Q=2)
sh.Gram_matrix("Gaussian", 0.2, 1e-6); sh.Gram_Cholesky(); sh.GN_method(2, 1, "rdm", print_hist=False)
print("sharded (virtual 2 x 2):", sh.loss_hist[-1])
print("sanitize run finished:", p.loss_hist[-1], b.loss_hist[-1], e.loss_hist[-1], d.loss_hist[-1])
