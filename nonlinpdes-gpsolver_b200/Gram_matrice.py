"""Gram builders with the reference's signatures (src/Gram_matrice.py:11-289).

These free functions return dense numpy matrices like the reference does (use them for inspection and
parity tests).  The solver classes keep Theta on the device instead."""
import numpy as onp

from . import _lib


def _slots(eqn):
    if eqn == "Darcy_flow2d":
        return [(0, "Darcy_flow2d"), (1, "Darcy_flow2d_a")]
    if eqn in ("Nonlinear_elliptic", "Burgers", "Eikonal"):
        return [(0, eqn)]
    raise ValueError(f"unknown eqn {eqn!r}")


def Gram_matrix_assembly(X_domain, X_boundary, eqn='Nonlinear_elliptic', kernel='Gaussian', kernel_parameter=0.2):
    """Dense symmetric Theta (Darcy: the tuple (Theta_u, Theta_a)), src/Gram_matrice.py:11-187."""
    eng = _lib.default_engine()
    eng.set_points(onp.asarray(X_domain)[:, :2], onp.asarray(X_boundary))
    out = []
    for slot, layout in _slots(eqn):
        eng.gram_assemble(slot, layout, kernel, kernel_parameter)
        out.append(eng.gram_download(slot, 0))
    return tuple(out) if len(out) > 1 else out[0]


def construct_Theta_test(X_test, X_domain, X_boundary, eqn='Nonlinear_elliptic', kernel='Gaussian', kernel_parameter=0.2):
    """N_test x M cross-Gram with the y-side functionals (Darcy: (Theta_u_test, Theta_a_test)),
    src/Gram_matrice.py:190-289."""
    eng = _lib.default_engine()
    eng.set_points(onp.asarray(X_domain)[:, :2], onp.asarray(X_boundary))
    out = []
    for slot, layout in _slots(eqn):
        eng.gram_assemble(slot, layout, kernel, kernel_parameter)
        out.append(eng.theta_test(slot, X_test))
    return tuple(out) if len(out) > 1 else out[0]
