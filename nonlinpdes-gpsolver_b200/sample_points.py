"""Collocation-point sampling (reference: src/sample_points.py:5-102).

Stays on the host in numpy on purpose: points must be bit-identical to the reference's for a given
seed, and numpy's legacy global RNG stream is frozen by numpy's compatibility policy.  The draws below
are issued in exactly the reference's order."""
import numpy as onp
from numpy import random


def _faces(domain):
    return domain[0, 0], domain[0, 1], domain[1, 0], domain[1, 1]


def sampled_pts_rdm(N_domain, N_boundary, domain, time_dependent=False):
    x1l, x1r, x2l, x2r = _faces(domain)
    # interior: column 0 first, then column 1 (src/sample_points.py:13)
    c0 = random.uniform(x1l, x1r, (N_domain, 1))
    c1 = random.uniform(x2l, x2r, (N_domain, 1))
    X_domain = onp.concatenate((c0, c1), axis=1)
    if not time_dependent:
        n = int(N_boundary / 4)
        X_boundary = onp.zeros((n * 4, 2))
        X_boundary[0:n, 0] = random.uniform(x1l, x1r, n)            # bottom  (:18-19)
        X_boundary[0:n, 1] = x2l
        X_boundary[n:2 * n, 0] = x1r                                 # right   (:21-22)
        X_boundary[n:2 * n, 1] = random.uniform(x2l, x2r, n)
        X_boundary[2 * n:3 * n, 0] = random.uniform(x1l, x1r, n)    # top     (:24-25)
        X_boundary[2 * n:3 * n, 1] = x2r
        X_boundary[3 * n:4 * n, 1] = random.uniform(x2l, x2r, n)    # left    (:27-28)
        X_boundary[3 * n:4 * n, 0] = x1l
    else:
        n = int(N_boundary / 3)
        X_boundary = onp.zeros((n * 3, 2))
        X_boundary[0:n, 1] = random.uniform(x2l, x2r, n)            # t = t0  (:38-39)
        X_boundary[0:n, 0] = x1l
        X_boundary[n:2 * n, 0] = random.uniform(x1l, x1r, n)        # x = x_right (:41-42)
        X_boundary[n:2 * n, 1] = x2r
        X_boundary[2 * n:, 0] = random.uniform(x1l, x1r, n)         # x = x_left  (:44-45)
        X_boundary[2 * n:, 1] = x2l
    return X_domain, X_boundary


def sampled_pts_grid(N_domain, N_boundary, domain, time_dependent=False):
    x1l, x1r, x2l, x2r = _faces(domain)
    N_pts = int(onp.sqrt(N_domain + N_boundary)) - 2
    xx = onp.linspace(x1l, x1r, N_pts + 2)
    yy = onp.linspace(x2l, x2r, N_pts + 2)
    XX, YY = onp.meshgrid(xx, yy)
    e = N_pts + 1
    if not time_dependent:   # :56-77
        Xi, Yi = XX[1:e, 1:e], YY[1:e, 1:e]
        Xb = onp.concatenate((XX[0, 0:e], XX[e, 0:e], XX[0:e, 0], XX[0:e, e]))
        Yb = onp.concatenate((YY[0, 0:e], YY[e, 0:e], YY[0:e, 0], YY[0:e, e]))
    else:                    # :79-100 (no boundary at the final time)
        Xi, Yi = XX[1:e, 1:e + 1], YY[1:e, 1:e + 1]
        Xb = onp.concatenate((XX[0, 1:e + 1], XX[e, 1:e + 1], XX[0:e + 1, 0]))
        Yb = onp.concatenate((YY[0, 1:e + 1], YY[e, 1:e + 1], YY[0:e + 1, 0]))
    X_domain = onp.stack((Xi.flatten(), Yi.flatten()), axis=1)
    X_boundary = onp.stack((Xb, Yb), axis=1)
    return X_domain, X_boundary
