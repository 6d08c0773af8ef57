"""Driver with the reference's command line (main_Burgers1d.py): u_t + alpha u u_x - nu u_xx = 0."""
import argparse

import numpy as onp

from _common import make_parser, unit_grid
from nonlinpdes_gpsolver_b200.solver import solver_GP

cfg = make_parser('Burgers equation GP solver',
                  [("--alpha", dict(type=float, default=1.0)), ("--nu", dict(type=float, default=0.02))],
                  kernel="anisotropic_Gaussian", kernel_parameter=[0.3, 0.05], randomseed=0).parse_args()
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


XX, YY, X_test = unit_grid(60, -1.0, 1.0)
solver.test(X_test)
solver.get_test_error(u_truth(X_test[:, 0], X_test[:, 1]))
