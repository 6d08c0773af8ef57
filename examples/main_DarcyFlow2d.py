"""Driver with the reference's command line (main_DarcyFlow2d.py): -div(a grad u) = f, infer a from noisy u."""
import argparse

import numpy as onp
from scipy.interpolate import griddata
from scipy.sparse import diags
from scipy.sparse.linalg import spsolve

from _common import make_parser, unit_grid
from nonlinpdes_gpsolver_b200.solver import solver_GP


def a(x1, x2):
    return onp.exp(onp.sin(2 * onp.pi * x1) + onp.sin(2 * onp.pi * x2)) + onp.exp(-onp.sin(2 * onp.pi * x1) - onp.sin(2 * onp.pi * x2))


def FD_Darcy_flow_2d(N, fun_a, f_val):
    """5-point FD with zero Dirichlet data, padded with the boundary (reference_solver/FD_for_Darcy_flow.py:8-32)."""
    hg = 1 / (N + 1)
    x_mid = (onp.arange(0, N + 1) + 0.5) * hg
    x_grid = onp.arange(1, N + 1) * hg
    mid, grid = onp.meshgrid(x_mid, x_grid)
    a1 = onp.reshape(fun_a(mid.flatten(), grid.flatten()), (N, N + 1))
    a2 = onp.transpose(onp.reshape(fun_a(grid.flatten(), mid.flatten()), (N, N + 1)))
    a_diag = onp.reshape(a1[:, :N] + a1[:, 1:] + a2[:N, :] + a2[1:, :], -1)
    a_super1 = onp.reshape(onp.append(a1[:, 1:N], onp.zeros((N, 1)), axis=1), -1)
    a_super2 = onp.reshape(a2[1:N, :], -1)
    A = diags([-a_super2, -a_super1[:-1], a_diag, -a_super1[:-1], -a_super2], [-N, -1, 0, 1, N], shape=(N ** 2, N ** 2), format='csc') / hg ** 2
    out = onp.zeros((N + 2, N + 2))
    out[1:N + 1, 1:N + 1] = onp.reshape(spsolve(A, onp.full(N * N, float(f_val))), (N, N))
    return out


cfg = make_parser('Darcy flow inverse problem GP solver',
                  [("--N_data", dict(type=int, default=60)), ("--noise_level", dict(type=float, default=1e-3))],
                  nugget=1e-8, N_domain=400, N_boundary=100, randomseed=9999).parse_args()
onp.random.seed(cfg.randomseed)
print(f"[Seeds] random seeds: {cfg.randomseed}")
solver = solver_GP(cfg, PDE_type="Darcy_flow2d")
solver.set_equation(bdy=lambda x1, x2: 0, rhs=lambda x1, x2: 1, domain=onp.array([[0, 1], [0, 1]]))
solver.auto_sample_IP(cfg.N_domain, cfg.N_boundary, cfg.N_data, sampled_type=cfg.sampled_type)
N_pts_per_dim = 80
XX, YY, X_test = unit_grid(N_pts_per_dim)
u_truth_grid = FD_Darcy_flow_2d(N_pts_per_dim - 2, a, 1.0)
data_u = griddata((XX.flatten(), YY.flatten()), u_truth_grid.reshape(-1), (solver.eqn.X_data[:, 0], solver.eqn.X_data[:, 1]), method='linear')
solver.get_observed_data(data_u, cfg.noise_level)
solver.solve()
solver.test(X_test)
test_u = onp.reshape(solver.eqn.extended_sol_u, (N_pts_per_dim, N_pts_per_dim))
test_a = onp.reshape(solver.eqn.extended_sol_a, (N_pts_per_dim, N_pts_per_dim))
err_u = onp.abs(test_u - u_truth_grid)
err_a = onp.abs(onp.exp(test_a) - a(XX, YY))
print(f'[Test error u] Max error {onp.max(err_u)}, L2 error {onp.sqrt(onp.mean(err_u ** 2))}')
print(f'[Test error a] Max error {onp.max(err_a)}, L2 error {onp.sqrt(onp.mean(err_a ** 2))}')
