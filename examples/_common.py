"""Shared pieces of the example drivers: the flag set of the reference's main_*.py scripts, built from one table."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def str2bool(v):
    """The reference declares its boolean flags with type=bool, so any non-empty string is True
    (main_NonLinElliptic2d.py:44-45).  Here "false"/"0"/"no" mean False."""
    return str(v).lower() not in ("false", "0", "no", "")


def make_parser(description, extra, **defaults):
    """`extra`: list of (flag, kwargs) specific to one driver; `defaults`: per-driver defaults of the shared flags."""
    d = dict(kernel="Gaussian", kernel_parameter=0.2, nugget=1e-5, N_domain=1000, N_boundary=200, GNsteps=8, initial_sol="rdm",
             randomseed=None)
    d.update(defaults)
    ap = argparse.ArgumentParser(description=description)
    for flag, kw in extra:
        ap.add_argument(flag, **kw)
    ap.add_argument("--kernel", type=str, default=d["kernel"])
    if isinstance(d["kernel_parameter"], (list, tuple)):
        ap.add_argument("--kernel_parameter", type=float, nargs=len(d["kernel_parameter"]), default=list(d["kernel_parameter"]))
    else:
        ap.add_argument("--kernel_parameter", type=float, default=d["kernel_parameter"])
    ap.add_argument("--nugget", type=float, default=d["nugget"])
    ap.add_argument("--nugget_type", type=str, default="adaptive", choices=["adaptive", "identity", "none"])
    ap.add_argument("--sampled_type", type=str, default="random", choices=["random", "grid"])
    for name in ("N_domain", "N_boundary", "GNsteps"):
        ap.add_argument("--" + name, type=int, default=d[name])
    ap.add_argument("--method", type=str, default="elimination", choices=["elimination", "relaxation"])
    ap.add_argument("--pen_lambda", type=float, default=1e-10)
    ap.add_argument("--initial_sol", type=str, default=d["initial_sol"])
    ap.add_argument("--step_size", type=int, default=1)
    ap.add_argument("--print_hist", type=str2bool, default=True)
    ap.add_argument("--show_figure", type=str2bool, default=False)
    ap.add_argument("--randomseed", type=int, default=d["randomseed"], help="numpy seed (None: unseeded, like the reference's elliptic/Eikonal drivers)")
    return ap


def unit_grid(n, lo2=0.0, hi2=1.0, trim=False):
    import numpy as onp
    xx, yy = onp.linspace(0, 1, n), onp.linspace(lo2, hi2, n)
    if trim:
        xx, yy = xx[1:-1], yy[1:-1]
    XX, YY = onp.meshgrid(xx, yy)
    return XX, YY, onp.concatenate((XX.reshape(-1, 1), YY.reshape(-1, 1)), axis=1)
