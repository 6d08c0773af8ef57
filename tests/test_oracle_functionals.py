"""Second, independent oracle for the kernel functionals: torch.func autodiff in
float64, nesting grad exactly like src/kernels.py:8-179 nests jax.grad.  Checks the
closed forms of oracle/gp_oracle.py for both kernels and all 19 methods.  CPU only."""
import numpy as np
import pytest
import torch
from torch.func import grad, vmap

from oracle import gp_oracle as o


def _kappa_gauss(x1, x2, y1, y2, s):          # src/kernels.py:12-13
    return torch.exp(-(1 / (2 * s ** 2)) * ((x1 - y1) ** 2 + (x2 - y2) ** 2))


def _kappa_aniso(x1, x2, y1, y2, s):          # src/kernels.py:95-99
    r = ((x1 - y1) / s[0]) ** 2 + ((x2 - y2) / s[1]) ** 2
    return torch.exp(-r)


def _methods(k):
    """Same nesting as the reference: argnums 0,1 = x1,x2 ; 2,3 = y1,y2."""
    g = lambda f, i: grad(f, argnums=i)
    lap_x = lambda f: (lambda *a: g(g(f, 0), 0)(*a) + g(g(f, 1), 1)(*a))
    lap_y = lambda f: (lambda *a: g(g(f, 2), 2)(*a) + g(g(f, 3), 3)(*a))
    return {
        "kappa": k,
        "D_x1_kappa": g(k, 0), "D_x2_kappa": g(k, 1), "DD_x2_kappa": g(g(k, 1), 1),
        "D_y1_kappa": g(k, 2), "D_y2_kappa": g(k, 3), "DD_y2_kappa": g(g(k, 3), 3),
        "D_x1_D_y1_kappa": g(g(k, 0), 2), "D_x1_D_y2_kappa": g(g(k, 0), 3),
        "D_x1_DD_y2_kappa": g(g(g(k, 0), 3), 3), "D_x2_D_y2_kappa": g(g(k, 1), 3),
        "D_x2_D_y1_kappa": g(g(k, 1), 2), "D_x2_DD_y2_kappa": g(g(g(k, 1), 3), 3),
        "DD_x2_DD_y2_kappa": g(g(g(g(k, 1), 1), 3), 3),
        "Delta_x_kappa": lap_x(k), "Delta_y_kappa": lap_y(k),
        "Delta_x_Delta_y_kappa": lap_x(lap_y(k)),
        "Delta_x_D_y1_kappa": g(lap_x(k), 2), "Delta_x_D_y2_kappa": g(lap_x(k), 3),
    }


@pytest.mark.parametrize("kernel,param", [("Gaussian", 0.2), ("anisotropic_Gaussian", (0.3, 0.05)),
                                          ("anisotropic_Gaussian", (1 / 3, 1 / 20))])
def test_closed_forms_match_autodiff(kernel, param):
    rng = np.random.RandomState(1)
    n = 400
    pts = rng.uniform(0, 1, (n, 4))
    if kernel != "Gaussian":
        pts[:, 1] = pts[:, 1] * 2 - 1
        pts[:, 3] = pts[:, 1] + rng.uniform(-0.2, 0.2, n)   # keep the narrow x-scale in range
    pts[:5, 2:] = pts[:5, :2]                                # coincident pairs (diagonal of Theta)
    t = [torch.tensor(pts[:, i], dtype=torch.float64) for i in range(4)]
    if kernel == "Gaussian":
        k = lambda a, b, c, d: _kappa_gauss(a, b, c, d, param)
    else:
        k = lambda a, b, c, d: _kappa_aniso(a, b, c, d, param)
    for name, fn in _methods(k).items():
        ref = vmap(fn)(*t).numpy()
        opx, opy = o.REFERENCE_METHODS[name]
        got = o.functional(kernel, param, opx, opy, pts[:, 0], pts[:, 1], pts[:, 2], pts[:, 3])
        scale = np.max(np.abs(ref))
        # tolerance of BASELINE north_star (1e-12), relative to max(|ref|, 1e-4 * block max)
        tol = 1e-12 * np.maximum(np.abs(ref), 1e-4 * scale)
        assert np.all(np.abs(got - ref) <= tol), (name, np.max(np.abs(got - ref) / scale))


def test_block_layout_and_symmetry():
    rng = np.random.RandomState(0)
    Xd, Xb = rng.uniform(0, 1, (7, 2)), rng.uniform(0, 1, (4, 2))
    for eqn, M in [("Nonlinear_elliptic", 18), ("Burgers", 32), ("Eikonal", 32)]:
        Th = o.Gram_matrix_assembly(Xd, Xb, eqn, "Gaussian", 0.2)
        assert Th.shape == (M, M)
        assert np.array_equal(Th, Th.T)                     # exactly symmetric by construction
    Tu, Ta = o.Gram_matrix_assembly(Xd, Xb, "Darcy_flow2d", "Gaussian", 0.2)
    N = 7
    # Theta_a is the [d1; d2; id(interior)] sub-matrix of Theta_u (src/Gram_matrice.py:183-186)
    idx = np.r_[0:2 * N, 3 * N:4 * N]
    assert np.array_equal(Ta, Tu[np.ix_(idx, idx)])
    # analytic adaptive-nugget ratios
    _, r = o.add_nugget(o.Gram_matrix_assembly(Xd, Xb, "Nonlinear_elliptic", "Gaussian", 0.2),
                        "Nonlinear_elliptic", 7, 4, 1e-8, "adaptive")
    np.testing.assert_allclose(r[0], 7 * 8 / 0.2 ** 4 / 11, rtol=1e-14)


def test_sampler_matches_reference_call_order():
    """src/sample_points.py:5-48: RNG call order (bit-exact requirement)."""
    dom = np.array([[0.0, 1.0], [-1.0, 1.0]])
    np.random.seed(0)
    Xd, Xb = o.sampled_pts_rdm(10, 7, dom, time_dependent=True)
    np.random.seed(0)
    c0 = np.random.uniform(0, 1, (10, 1)); c1 = np.random.uniform(-1, 1, (10, 1))
    f0 = np.random.uniform(-1, 1, 2); f1 = np.random.uniform(0, 1, 2); f2 = np.random.uniform(0, 1, 2)
    assert np.array_equal(Xd, np.concatenate([c0, c1], 1)) and Xb.shape == (6, 2)
    assert np.array_equal(Xb[:2, 1], f0) and np.array_equal(Xb[2:4, 0], f1) and np.array_equal(Xb[4:, 0], f2)
    assert np.all(Xb[:2, 0] == 0) and np.all(Xb[2:4, 1] == 1) and np.all(Xb[4:, 1] == -1)
