"""Driver with the reference's command line (main_Eikonal2d.py): |grad u|^2 = f^2 + eps Delta u."""
import argparse

import numpy as onp
import scipy.sparse
from scipy.sparse import diags, identity
from scipy.sparse.linalg import spsolve

from _common import str2bool
from nonlinpdes_gpsolver_b200.solver import solver_GP


def get_parser():
    parser = argparse.ArgumentParser(description='Eikonal equation GP solver')
    parser.add_argument("--eps", type=float, default=1e-1)
    parser.add_argument("--kernel", type=str, default="Gaussian")
    parser.add_argument("--kernel_parameter", type=float, default=0.2)
    parser.add_argument("--nugget", type=float, default=1e-5)
    parser.add_argument("--nugget_type", type=str, default="adaptive", choices=["adaptive", "identity", 'none'])
    parser.add_argument("--sampled_type", type=str, default='random', choices=['random', 'grid'])
    parser.add_argument("--N_domain", type=int, default=1000)
    parser.add_argument("--N_boundary", type=int, default=200)
    parser.add_argument("--method", type=str, default='elimination')
    parser.add_argument("--initial_sol", type=str, default='zero')
    parser.add_argument("--GNsteps", type=int, default=8)
    parser.add_argument("--step_size", type=int, default=1)
    parser.add_argument("--print_hist", type=str2bool, default=True)
    parser.add_argument("--show_figure", type=str2bool, default=False)
    parser.add_argument("--randomseed", type=int, default=None)
    return parser.parse_args()


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


cfg = get_parser()
if cfg.randomseed is not None:
    onp.random.seed(cfg.randomseed)
solver = solver_GP(cfg, PDE_type="Eikonal")
solver.set_equation(bdy=lambda x1, x2: 0, rhs=lambda x1, x2: 1, domain=onp.array([[0, 1], [0, 1]]))
solver.auto_sample(cfg.N_domain, cfg.N_boundary, sampled_type=cfg.sampled_type)
solver.solve()
N_pts = 60
xx = onp.linspace(0, 1, N_pts)[1:-1]
XX, YY = onp.meshgrid(xx, xx)
X_test = onp.concatenate((XX.reshape(-1, 1), YY.reshape(-1, 1)), axis=1)
solver.test(X_test)
XX, YY, test_truth = solve_Eikonal(N_pts - 2, cfg.eps)
solver.get_test_error(test_truth.flatten())
