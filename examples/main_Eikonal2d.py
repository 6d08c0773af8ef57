"""Driver with the reference's command line (main_Eikonal2d.py): |grad u|^2 = f^2 + eps Delta u."""
import argparse

import numpy as onp
import scipy.sparse
from scipy.sparse import diags, identity
from scipy.sparse.linalg import spsolve

from _common import make_parser, unit_grid
from nonlinpdes_gpsolver_b200.solver import solver_GP


def solve_Eikonal(N, epsilon):
    """Cole-Hopf transformed linear problem, 5-point FD (reference_solver/Cole_Hopf_for_Eikonal.py:7-36)."""
    hg = 1 / (N + 1)
    x_grid = onp.arange(1, N + 1) * hg
    a1, a2 = onp.ones((N, N + 1)), onp.ones((N + 1, N))
    a_diag = onp.reshape(a1[:, :N] + a1[:, 1:] + a2[:N, :] + a2[1:, :], -1)
    a_super1 = onp.reshape(onp.append(a1[:, 1:N], onp.zeros((N, 1)), axis=1), -1)
    a_super2 = onp.reshape(a2[1:N, :], -1)
    A = diags([-a_super2, -a_super1[:-1], a_diag, -a_super1[:-1], -a_super2], [-N, -1, 0, 1, N], shape=(N ** 2, N ** 2), format='csc')
    XX, YY = onp.meshgrid(x_grid, x_grid)
    f = onp.zeros((N, N))
    for sl in ((0, slice(None)), (N - 1, slice(None)), (slice(None), 0), (slice(None), N - 1)):
        f[sl] += epsilon ** 2 / hg ** 2
    sol_v = spsolve((identity(N ** 2) + (epsilon ** 2) * A / hg ** 2).tocsc(), f.flatten())
    return XX, YY, onp.reshape(-epsilon * onp.log(sol_v), (N, N))


cfg = make_parser('Eikonal equation GP solver', [("--eps", dict(type=float, default=1e-1))], initial_sol="zero").parse_args()
if cfg.randomseed is not None:
    onp.random.seed(cfg.randomseed)
solver = solver_GP(cfg, PDE_type="Eikonal")
solver.set_equation(bdy=lambda x1, x2: 0, rhs=lambda x1, x2: 1, domain=onp.array([[0, 1], [0, 1]]))
solver.auto_sample(cfg.N_domain, cfg.N_boundary, sampled_type=cfg.sampled_type)
solver.solve()
N_pts = 60
XX, YY, X_test = unit_grid(N_pts, trim=True)
solver.test(X_test)
XX, YY, test_truth = solve_Eikonal(N_pts - 2, cfg.eps)
solver.get_test_error(test_truth.flatten())
