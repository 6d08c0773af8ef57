"""Pin the CPU oracle against the reference's only recorded outputs: the stored
stdout of its notebooks (tests/golden/notebook_golden.json).  CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import gp_oracle as o

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_golden.json")))


def _grid(n, interior=False):
    if interior:
        xx = np.arange(1, n + 1) / (n + 1)
    else:
        xx = np.linspace(0, 1, n)
    XX, YY = np.meshgrid(xx, xx)
    return np.stack([XX.ravel(), YY.ravel()], 1)


@pytest.mark.parametrize("solve", ["lu", "tri"])
def test_elliptic_notebook(solve):
    g = G["elliptic"]
    np.random.seed(g["seed"])
    Xd, Xb = o.notebook_sample_points(g["N_domain"], g["N_boundary"])
    init = np.random.normal(0.0, 1.0, g["N_domain"])
    p = o.Nonlinear_elliptic2d(alpha=g["alpha"], m=g["m"])
    p.set_points(Xd, Xb, o.elliptic_f(Xd[:, 0], Xd[:, 1]), o.elliptic_u(Xb[:, 0], Xb[:, 1]))
    p.Gram_matrix("Gaussian", g["sigma"], g["nugget"], "adaptive")
    assert p.ratio == g["trace_ratio"]                      # bit-identical
    p.Gram_Cholesky(solve)
    p.GN_method(g["steps"], 1, init)
    np.testing.assert_allclose(p.loss_hist, g["loss_hist"], rtol=5e-11)
    err = np.abs(o.elliptic_u(Xd[:, 0], Xd[:, 1]) - p.sol_sampled_pts)
    np.testing.assert_allclose(np.sqrt(np.sum(err ** 2) / g["N_domain"]), g["pts_L2"], rtol=1e-10)
    np.testing.assert_allclose(err.max(), g["pts_max"], rtol=1e-10)
    Xt = _grid(100)
    p.extend_sol(Xt)
    e = np.abs(p.extended_sol - o.elliptic_u(Xt[:, 0], Xt[:, 1]))
    np.testing.assert_allclose(np.linalg.norm(e) / 100, g["test100_L2"], rtol=1e-10)
    np.testing.assert_allclose(e.max(), g["test100_max"], rtol=1e-10)

    # The notebook's second (tuned-sigma) run, :619-643, is not pinned: its RNG stream
    # position cannot be recovered from the stored cells (see golden json "elliptic_tuned").


def test_eikonal_notebook():
    g = G["eikonal"]
    np.random.seed(g["seed"])
    Xd, Xb = o.notebook_sample_points(g["N_domain"], g["N_boundary"])
    p = o.Eikonal(eps=g["eps"])
    p.set_points(Xd, Xb, np.ones(g["N_domain"]), np.zeros(g["N_boundary"]))
    p.Gram_matrix("Gaussian", g["sigma"], g["nugget"], "adaptive")
    assert list(p.ratio) == g["trace_ratio"]               # bit-identical
    p.Gram_Cholesky("lu")
    p.GN_method(g["steps"], 1.0, "zero")
    np.testing.assert_allclose(p.loss_hist, g["loss_hist"], rtol=5e-9)
    XX, YY, truth = o.solve_Eikonal(100, g["eps"])
    p.extend_sol(np.stack([XX.ravel(), YY.ravel()], 1))
    e = np.abs(p.extended_sol.reshape(100, 100) - truth)
    np.testing.assert_allclose(np.linalg.norm(e) / 100, g["test100_L2"], rtol=1e-8)
    np.testing.assert_allclose(e.max(), g["test100_max"], rtol=1e-8)


def test_darcy_notebook():
    from scipy.interpolate import griddata
    g = G["darcy"]
    np.random.seed(g["seed"])
    ut = o.FD_Darcy_flow_2d(100)
    xx = np.linspace(0, 1, 102)
    XX, YY = np.meshgrid(xx, xx)
    Xd, Xb = o.notebook_sample_points(g["N_domain"], g["N_boundary"])
    init = np.random.normal(0, 1.0, 6 * g["N_domain"])
    data_u = griddata((XX.flatten(), YY.flatten()), ut.reshape(-1), (Xd[:g["N_data"], 0], Xd[:g["N_data"], 1]),
                      method="linear")
    d = o.Darcy_flow2d()
    d.set_points(Xd, Xb, g["N_data"], np.ones(g["N_domain"]), np.zeros(g["N_boundary"]))
    d.get_observation(data_u, g["noise"])                   # draws the noise after the initial guess
    d.Gram_matrix("Gaussian", g["sigma"], g["nugget"], "adaptive")
    assert list(d.ratio_u) == g["trace_ratio_u"] and list(d.ratio_a) == g["trace_ratio_a"]
    d.Gram_Cholesky("lu")
    d.GN_method(g["steps"], 1, init)
    np.testing.assert_allclose(d.loss_hist, g["loss_hist"], rtol=2e-9)


def test_structured_hessian_and_percall_lu_equal_dense_oracle():
    """The two cost-model variants of the oracle (Hessian from the interior inverse block; LU of L redone per call like
    the reference's jnp.linalg.solve) give the same numbers as the default dense / factor-once form."""
    np.random.seed(7)
    N, Nb = 80, 24
    Xd, Xb = o.sampled_pts_rdm(N, Nb, np.array([[0.0, 1.0], [0.0, 1.0]]))
    for make, nz, init in ((lambda: o.Nonlinear_elliptic2d(1.0, 3), 1, None), (lambda: o.Eikonal(0.1), 3, "zero")):
        runs = []
        z0 = np.random.normal(0.0, 1.0, nz * N) if init is None else init
        for solve, structured in (("lu", False), ("tri", True), ("lu_percall", False)):
            p = make()
            p.set_points(Xd, Xb, o.elliptic_f(Xd[:, 0], Xd[:, 1]) if nz == 1 else np.ones(N),
                         o.elliptic_u(Xb[:, 0], Xb[:, 1]) if nz == 1 else np.zeros(Xb.shape[0]))
            p.Gram_matrix("Gaussian", 0.2, 1e-6, "adaptive")
            p.Gram_Cholesky(solve, structured=structured)
            p.GN_method(3, 1, z0)
            runs.append(p)
        z = runs[0].sol
        H0, H1 = runs[0].Hessian_GN(z), runs[1].Hessian_GN(z)
        assert np.max(np.abs(H0 - H1)) <= 1e-7 * np.max(np.abs(H0))
        np.testing.assert_allclose(runs[1].loss_hist, runs[0].loss_hist, rtol=1e-8)
        np.testing.assert_allclose(runs[2].loss_hist, runs[0].loss_hist, rtol=1e-12)
        assert runs[2]._s.n_factorizations >= 3 * 3 and runs[0]._s.n_factorizations == 1
