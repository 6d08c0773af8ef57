"""Driver with the reference's command line (main_Burgers1d.py): u_t + alpha u u_x - nu u_xx = 0."""
import argparse

import numpy as onp

from _common import str2bool
from nonlinpdes_gpsolver_b200.solver import solver_GP


def get_parser():
    parser = argparse.ArgumentParser(description='Burgers equation GP solver')
    parser.add_argument("--alpha", type=float, default=1.0)
    parser.add_argument("--nu", type=float, default=0.02)
    parser.add_argument("--kernel", type=str, default="anisotropic_Gaussian")
    parser.add_argument("--kernel_parameter", type=float, nargs=2, default=[0.3, 0.05])
    parser.add_argument("--nugget", type=float, default=1e-5)
    parser.add_argument("--nugget_type", type=str, default="adaptive", choices=["adaptive", "identity", 'none'])
    parser.add_argument("--sampled_type", type=str, default='random', choices=['random', 'grid'])
    parser.add_argument("--N_domain", type=int, default=1000)
    parser.add_argument("--N_boundary", type=int, default=200)
    parser.add_argument("--method", type=str, default='elimination')
    parser.add_argument("--initial_sol", type=str, default='rdm')
    parser.add_argument("--GNsteps", type=int, default=8)
    parser.add_argument("--step_size", type=int, default=1)
    parser.add_argument("--print_hist", type=str2bool, default=True)
    parser.add_argument("--show_figure", type=str2bool, default=False)
    parser.add_argument("--randomseed", type=int, default=0)
    return parser.parse_args()


cfg = get_parser()
onp.random.seed(cfg.randomseed)
print(f"[Seeds] random seeds: {cfg.randomseed}")
alpha, nu = cfg.alpha, cfg.nu
solver = solver_GP(cfg, PDE_type="Burgers")


def u(x1, x2):
    return -onp.sin(onp.pi * x2) * (x1 == 0) + 0 * (x2 == 0)


def f(x1, x2):
    return 0


solver.set_equation(bdy=u, rhs=f, domain=onp.array([[0, 1], [-1, 1]]))
solver.auto_sample(cfg.N_domain, cfg.N_boundary, sampled_type=cfg.sampled_type)
solver.solve()
Gauss_pts, weights = onp.polynomial.hermite.hermgauss(80)


def u_truth(x1, x2):   # Cole-Hopf with 80-point Gauss-Hermite quadrature (main_Burgers1d.py:87-92)
    temp = x2[:, None] - onp.sqrt(4 * nu * x1[:, None]) * Gauss_pts[None, :]
    e = onp.exp(-onp.cos(onp.pi * temp) / (2 * onp.pi * nu))
    return -onp.sum(weights * onp.sin(onp.pi * temp) * e, axis=1) / onp.sum(weights * e, axis=1)


N_pts = 60
XX, YY = onp.meshgrid(onp.linspace(0, 1, N_pts), onp.linspace(-1, 1, N_pts))
X_test = onp.concatenate((XX.reshape(-1, 1), YY.reshape(-1, 1)), axis=1)
solver.test(X_test)
solver.get_test_error(u_truth(X_test[:, 0], X_test[:, 1]))
