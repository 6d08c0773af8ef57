"""Driver with the reference's command line (main_NonLinElliptic2d.py), numpy callables instead of jax.
    python examples/main_NonLinElliptic2d.py --kernel Gaussian --kernel_parameter 0.2 --nugget 1e-13 --N_domain 900 --N_boundary 124 --GNsteps 4
solves -Delta u + alpha*u^m = f in [0,1]^2 on the GPU."""
import argparse

import numpy as onp

from _common import make_parser, unit_grid
from nonlinpdes_gpsolver_b200.solver import solver_GP

cfg = make_parser('NonLinElliptic equation GP solver',
                  [("--alpha", dict(type=float, default=1.0)), ("--m", dict(type=float, default=3.0))],
                  nugget=1e-13, N_domain=900, N_boundary=124, GNsteps=4).parse_args()
if cfg.randomseed is not None:
    onp.random.seed(cfg.randomseed)
solver = solver_GP(cfg, PDE_type="Nonlinear_elliptic")
alpha, m = cfg.alpha, cfg.m


def u(x1, x2):
    return onp.sin(onp.pi * x1) * onp.sin(onp.pi * x2) + 2 * onp.sin(4 * onp.pi * x1) * onp.sin(4 * onp.pi * x2)


def f(x1, x2):   # -Laplace(u) + alpha*u^m, analytic (the reference differentiates u with jax.grad)
    s1 = onp.sin(onp.pi * x1) * onp.sin(onp.pi * x2)
    s4 = onp.sin(4 * onp.pi * x1) * onp.sin(4 * onp.pi * x2)
    return 2 * onp.pi ** 2 * s1 + 64 * onp.pi ** 2 * s4 + alpha * (u(x1, x2) ** m)


solver.set_equation(bdy=u, rhs=f, domain=onp.array([[0, 1], [0, 1]]), print_option=cfg.print_hist)
solver.auto_sample(cfg.N_domain, cfg.N_boundary, sampled_type=cfg.sampled_type, print_option=cfg.print_hist)
if cfg.show_figure:
    solver.show_sample()
solver.solve(method=cfg.method, pen_lambda=cfg.pen_lambda, print_option=cfg.print_hist)
pts_truth = u(solver.eqn.X_domain[:, 0], solver.eqn.X_domain[:, 1])
solver.collocation_pts_err(pts_truth)
XX, YY, X_test = unit_grid(60)
solver.test(X_test)
solver.get_test_error(u(X_test[:, 0], X_test[:, 1]))
if cfg.show_figure:
    solver.contour_of_test_err(XX, YY)
