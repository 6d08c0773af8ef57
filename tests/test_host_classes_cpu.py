"""Host-side classes on the CPU: the problem classes and the solver_GP facade driven through a fake engine
(tests/fake_engine.py, backed by the oracle) and compared with the oracle's own classes run with the same seed.
Checks what the host layer owns: RNG call order (sampling -> [noise] -> initial guess), trace-ratio nugget,
attribute names, the GN loop bookkeeping, state errors, Darcy's two Gram systems."""
from types import SimpleNamespace

import numpy as np
import pytest

from fake_engine import FakeEngine
from oracle import gp_oracle as o

DOM = np.array([[0.0, 1.0], [0.0, 1.0]])
DOM_T = np.array([[0.0, 1.0], [-1.0, 1.0]])


@pytest.fixture()
def fake(monkeypatch):
    from nonlinpdes_gpsolver_b200 import _lib
    monkeypatch.setattr(_lib, "Engine", lambda *a, **k: FakeEngine())
    monkeypatch.setattr(_lib, "_default_engine", None)
    return _lib


def _oracle_run(ref, Xd, Xb, rhs_f, bdy_g, kernel, kp, nugget, steps, init):
    ref.set_points(Xd, Xb, rhs_f, bdy_g)
    ref.Gram_matrix(kernel, kp, nugget, "adaptive")
    ref.Gram_Cholesky("tri")
    ref.GN_method(steps, 1, init)
    return ref


@pytest.mark.parametrize("name", ["elliptic", "burgers", "eikonal"])
def test_problem_classes_match_oracle(fake, name):
    from nonlinpdes_gpsolver_b200 import PDEs
    N, Nb, steps, seed = 60, 20, 3, 5
    if name == "elliptic":
        p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f, domain=DOM)
        ref, kernel, kp, nug, init_kind, dom, td = o.Nonlinear_elliptic2d(1.0, 3), "Gaussian", 0.2, 1e-6, "rdm", DOM, False
    elif name == "burgers":
        p = PDEs.Burgers(alpha=1.0, nu=0.02, bdy=o.burgers_bdy, rhs=lambda a, b: 0, domain=DOM_T)
        ref, kernel, kp, nug, init_kind, dom, td = o.Burgers(1.0, 0.02), "anisotropic_Gaussian", (0.3, 0.05), 1e-5, "rdm", DOM_T, True
    else:
        p = PDEs.Eikonal(eps=0.1, bdy=lambda a, b: 0, rhs=lambda a, b: 1, domain=DOM)
        ref, kernel, kp, nug, init_kind, dom, td = o.Eikonal(0.1), "Gaussian", 0.2, 1e-5, "zero", DOM, False
    np.random.seed(seed)
    p.sampled_pts(N, Nb)
    p.Gram_matrix(kernel, kp, nug, "adaptive")
    theta = p.Theta
    with pytest.raises(RuntimeError):
        p.L                                              # not factorised yet
    p.Gram_Cholesky()
    assert np.array_equal(p.Theta, theta)                # like the reference, Theta stays readable (re-assembled on request)
    p.Gram_Cholesky()                                    # idempotent
    p.GN_method(steps, 1, init_kind, print_hist=False)
    # the oracle with the reference's RNG call order: points, then the initial guess
    np.random.seed(seed)
    Xd, Xb = o.sampled_pts_rdm(N, Nb, dom, time_dependent=td)
    nz = {"elliptic": 1}.get(name, 3)
    init = np.random.normal(0.0, 1.0, nz * N) if init_kind == "rdm" else "zero"
    _oracle_run(ref, Xd, Xb, p.rhs_f, p.bdy_g, kernel, kp, nug, steps, init)
    assert np.array_equal(p.X_domain, Xd) and np.array_equal(p.X_boundary, Xb)
    assert p.N_domain == N and p.N_boundary == Xb.shape[0]
    if init_kind == "rdm":
        assert np.array_equal(p.init_sol, init)
    assert np.array_equal(theta, ref.Theta)
    np.testing.assert_allclose(np.atleast_1d(p.ratio), np.atleast_1d(ref.ratio), rtol=0, atol=0)
    assert len(p.loss_hist) == steps + 1 and p.max_iter == steps and p.step_size == 1
    np.testing.assert_allclose(p.loss_hist, ref.loss_hist, rtol=1e-9)
    np.testing.assert_allclose(p.sol_vec, ref.sol_vec, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(p.sol_sampled_pts, ref.sol_sampled_pts, rtol=1e-7, atol=1e-9)
    Xt = np.random.uniform(dom[:, 0], dom[:, 1], (17, 2))
    p.extend_sol(Xt)
    ref.extend_sol(Xt)
    assert p.N_test == 17 and p.X_test.shape == (17, 2)
    np.testing.assert_allclose(p.extended_sol, ref.extended_sol, rtol=1e-6, atol=1e-8)
    g, H = p.grad_loss(p.sol), p.Hessian_GN(p.sol, p.sol)
    assert g.shape == (nz * N,) and H.shape == (nz * N, nz * N)
    if name != "burgers":
        # GN_loss(z, z_old) is the quadratic whose Hessian is Hessian_GN (src/PDEs.py:94-102): check by differences
        rng = np.random.RandomState(0)
        d = rng.standard_normal(nz * N) * 1e-3
        z0 = p.sol
        second = p.GN_loss(z0 + d, z0) - 2 * p.GN_loss(z0, z0) + p.GN_loss(z0 - d, z0)
        np.testing.assert_allclose(second, d @ H @ d, rtol=1e-5)
    with pytest.raises(ValueError):
        p.GN_method(1, 1, "bogus", print_hist=False)     # upstream: NameError on an undefined `sol`


def test_nugget_types(fake):
    from nonlinpdes_gpsolver_b200 import PDEs
    np.random.seed(1)
    p = PDEs.Nonlinear_elliptic2d(bdy=o.elliptic_u, rhs=o.elliptic_f)
    p.sampled_pts(30, 12)
    p.Gram_matrix("Gaussian", 0.2, 1e-3, "none")
    base = p.Theta
    p.Gram_matrix("Gaussian", 0.2, 1e-3, "identity")
    np.testing.assert_allclose(p.Theta - base, 1e-3 * np.eye(base.shape[0]), atol=1e-12)
    p.Gram_matrix("Gaussian", 0.2, 1e-3, "adaptive")
    d = np.diag(p.Theta - base)
    np.testing.assert_allclose(d[:30], 1e-3 * p.ratio, rtol=1e-9)
    np.testing.assert_allclose(d[30:], 1e-3, rtol=1e-9)
    with pytest.raises(ValueError):
        p.Gram_matrix("Gaussian", 0.2, 1e-3, "bogus")


def test_darcy_class_matches_oracle(fake):
    from nonlinpdes_gpsolver_b200 import InverseProblems
    N, Nb, nd, steps, seed = 40, 16, 7, 2, 9
    np.random.seed(seed)
    d = InverseProblems.Darcy_flow2d(bdy=lambda a, b: 0, rhs=lambda a, b: 1, domain=DOM)
    d.sampled_pts(N, Nb, nd)
    data = np.linspace(0.0, 0.1, nd)
    d.get_observation(data, 1e-2)
    d.Gram_matrix("Gaussian", 0.2, 1e-6, "adaptive")
    d.Gram_Cholesky()
    d.GN_method(steps, 1, "rdm", print_hist=False)
    # reference RNG order (main_DarcyFlow2d.py): points -> observation noise -> initial guess
    np.random.seed(seed)
    Xd, Xb = o.sampled_pts_rdm(N, Nb, DOM)
    ref = o.Darcy_flow2d()
    ref.set_points(Xd, Xb, nd, np.ones(N), np.zeros(Xb.shape[0]))
    ref.get_observation(data, 1e-2)
    init = np.random.normal(0.0, 1.0, 6 * N)
    ref.Gram_matrix("Gaussian", 0.2, 1e-6, "adaptive")
    ref.Gram_Cholesky("tri")
    ref.GN_method(steps, 1, init)
    assert np.array_equal(d.X_domain, Xd) and np.array_equal(d.X_data, Xd[:nd]) and d.N_data == nd
    assert np.array_equal(d.data_u, ref.data_u) and np.array_equal(d.init_sol, init)
    assert list(d.ratio_u) == list(ref.ratio_u) and list(d.ratio_a) == list(ref.ratio_a)
    np.testing.assert_allclose(d.loss_hist, ref.loss_hist, rtol=1e-9)
    np.testing.assert_allclose(d.sol_vec_a, ref.sol_vec_a, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(d.sol_vec_u, ref.sol_vec_u, rtol=1e-7, atol=1e-9)
    Xt = np.random.uniform(0, 1, (11, 2))
    d.extend_sol(Xt)
    ref.extend_sol(Xt)
    np.testing.assert_allclose(d.extended_sol_u, ref.extended_sol_u, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(d.extended_sol_a, ref.extended_sol_a, rtol=1e-6, atol=1e-8)
    assert [c for c in d._engine().calls if c.startswith(("potrf", "inverse"))] == ["potrf[0]", "potrf[1]", "inverse[0]", "inverse[1]"]


def test_facade_end_to_end_on_fake_engine(fake, capsys):
    from nonlinpdes_gpsolver_b200.solver import solver_GP
    cfg = SimpleNamespace(alpha=1.0, m=3.0, kernel="Gaussian", kernel_parameter=0.2, nugget=1e-6, nugget_type="adaptive",
                          GNsteps=3, step_size=1, initial_sol="rdm", print_hist=True)
    np.random.seed(3)
    s = solver_GP(cfg, PDE_type="Nonlinear_elliptic")
    s.set_equation(bdy=o.elliptic_u, rhs=lambda a, b: o.elliptic_f(a, b, 1.0, 3.0), domain=DOM)
    s.auto_sample(50, 16)
    s.solve()
    s.collocation_pts_err(o.elliptic_u(s.eqn.X_domain[:, 0], s.eqn.X_domain[:, 1]))
    Xt = np.random.uniform(0, 1, (9, 2))
    s.test(Xt)
    s.get_test_error(o.elliptic_u(Xt[:, 0], Xt[:, 1]))
    out = capsys.readouterr().out
    for needle in ("[Sample points] N_domain = 50, N_boundary = 16", "iter = 0 Loss =", "iter =  3 Gauss-Newton step size = 1  Loss = ",
                   "[Gauss Newton] Gauss Newton iteration finished", "[Collocation point error] L2 error", "[Test error] Max error"):
        assert needle in out, needle
    assert s.pts_L2_err == np.sqrt(np.sum(s.pts_err_all ** 2) / 50) and s.test_err_all.shape == (9,)
    # user-supplied points (the upstream get_sample bug is fixed)
    s.get_sample(s.eqn.X_domain.copy(), s.eqn.X_boundary.copy(), print_option=False)
    assert s.eqn.N_domain == 50


def test_sharded_class_flow_on_fake_engine(fake):
    """PDEs.shard(): the problem class routes assembly, nugget, factorisation, inverse and GN steps through the dist_* calls
    (and only those), keeps the reference's RNG order, and produces the oracle's numbers; other PDEs refuse to shard."""
    from nonlinpdes_gpsolver_b200 import PDEs
    N, Nb, steps, seed = 50, 16, 2, 11
    np.random.seed(seed)
    p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f, domain=DOM)
    p.sampled_pts(N, Nb)
    assert p.shard(virtual_ranks=4, Q=2) is p
    assert p._engine().dist_info() == dict(rank=0, world=4, P=2, Q=2)
    p.Gram_matrix("Gaussian", 0.2, 1e-6, "adaptive")
    p.Gram_Cholesky()
    p.GN_method(steps, 1, "rdm", print_hist=False)
    calls = p._engine().calls
    assert "dist_gram_assemble" in calls and "dist_potrf" in calls and "dist_inverse" in calls and calls.count("dist_gn_step") == steps
    assert not any(c.startswith("inverse[") for c in calls)
    np.random.seed(seed)
    Xd, Xb = o.sampled_pts_rdm(N, Nb, DOM)
    init = np.random.normal(0.0, 1.0, N)
    ref = _oracle_run(o.Nonlinear_elliptic2d(1.0, 3), Xd, Xb, p.rhs_f, p.bdy_g, "Gaussian", 0.2, 1e-6, steps, init)
    assert np.array_equal(p.init_sol, init) and p.ratio == ref.ratio
    np.testing.assert_allclose(p.loss_hist, ref.loss_hist, rtol=1e-9)
    assert np.array_equal(p.Theta, ref.Theta)                       # re-assembled on a scratch handle in sharded mode
    with pytest.raises(NotImplementedError):
        p.Hessian_GN(p.sol, p.sol)                                  # no dense read-outs of a sharded problem
    with pytest.raises(NotImplementedError):
        PDEs.Eikonal(eps=0.1, bdy=lambda a, b: 0, rhs=lambda a, b: 1).shard(virtual_ranks=2)
