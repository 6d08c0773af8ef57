"""CPU oracle for the GP-PDE Gauss-Newton hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the reference's algorithm
(yifanc96/NonLinPDEs-GPsolver).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product package never does (it fails loudly without its CUDA library).

Parity status: PINNED.  The oracle reproduces the stored stdout of the
reference's notebooks (the only recorded outputs upstream) -- see
``tests/test_oracle_golden.py`` and ``tests/golden/notebook_golden.json``:
  * elliptic  (notebooks/Nonlinear_Elliptic_Equation.ipynb:333-338, :426, :561)
  * Eikonal   (notebooks/Regularized_Eikonal equation_eps1e-2.ipynb:376-386, :278)
  * Darcy     (notebooks/Darcy_flow_IP_noisy.ipynb:491-499, :334-335)
The Burgers notebook is unseeded upstream, so Burgers is pinned only through the
independent autodiff oracle in ``tests/test_oracle_functionals.py``.

The arithmetic the reference delegates to JAX (un-vendored, version unpinned;
README.md:5) is restated as follows:
  * ``jax.grad`` chains over the kernel (src/kernels.py:8-179) -> closed forms
    (Hermite-type polynomials times the kernel value), see ``_h``.
  * ``jnp.linalg.cholesky`` -> ``numpy.linalg.cholesky`` (LAPACK dpotrf).
  * ``jnp.linalg.solve(L, .)`` -> LU with partial pivoting on the triangular
    factor (``solve='lu'``, what jaxlib's getrf/getrs does) or a triangular
    solve (``solve='tri'``, what the CUDA path does).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg as sla

# --------------------------------------------------------------------------
# kernel functionals: closed forms for src/kernels.py
# --------------------------------------------------------------------------

# A linear functional is a sum of monomials d1^m1 d2^m2 (partial derivatives along
# axis 1 and axis 2).  (m1, m2) tuples:
OPS = {
    "id": [(0, 0)],          # point evaluation            (kappa)
    "d1": [(1, 0)],          # d/dx1                       (D_x1_kappa / D_y1_kappa)
    "d2": [(0, 1)],          # d/dx2                       (D_x2_kappa / D_y2_kappa)
    "d22": [(0, 2)],         # d^2/dx2^2                   (DD_x2_kappa / DD_y2_kappa)
    "lap": [(2, 0), (0, 2)], # Laplacian                   (Delta_x_kappa / Delta_y_kappa)
}

# Row-operator list per equation (block layout of Theta, src/Gram_matrice.py):
#   name -> list of (op, uses_boundary_points)
LAYOUT = {
    "Nonlinear_elliptic": [("lap", False), ("id", True)],                          # :41-56
    "Burgers": [("d1", False), ("d2", False), ("d22", False), ("id", True)],       # :58-99
    "Eikonal": [("d1", False), ("d2", False), ("lap", False), ("id", True)],       # :100-135
    "Darcy_flow2d": [("d1", False), ("d2", False), ("lap", False), ("id", True)],  # :137-179 (Theta_u)
    "Darcy_flow2d_a": [("d1", False), ("d2", False), ("id", False)],               # :183-186 (Theta_a)
}


def kernel_scales(kernel, kernel_parameter):
    """(b1, b2) such that kappa = exp(-(b1 d1^2 + b2 d2^2)/2).

    Gaussian (src/kernels.py:12-13): b1 = b2 = 1/sigma^2.
    anisotropic_Gaussian (src/kernels.py:95-99): b_i = 2/scale_i^2.
    """
    if kernel == "Gaussian":
        s = float(kernel_parameter)
        a = 1.0 / (s * s)
        return a, a
    if kernel == "anisotropic_Gaussian":
        st, sx = float(kernel_parameter[0]), float(kernel_parameter[1])
        return 2.0 / (st * st), 2.0 / (sx * sx)
    raise ValueError(f"unknown kernel {kernel!r}")


def kappa(kernel, kernel_parameter, d1, d2):
    """Kernel value from coordinate differences, same expression tree as the
    reference (src/kernels.py:13 and :96-99)."""
    if kernel == "Gaussian":
        s = float(kernel_parameter)
        return np.exp(-(1 / (2 * s ** 2)) * (d1 ** 2 + d2 ** 2))
    st, sx = float(kernel_parameter[0]), float(kernel_parameter[1])
    return np.exp(-((d1 / st) ** 2 + (d2 / sx) ** 2))


def _h(k, u, b):
    """h_k with  d^k/dd^k exp(-b d^2/2) = (-1)^k h_k(u) exp(-b d^2/2),  u = b d."""
    if k == 0:
        return np.ones_like(u)
    if k == 1:
        return u
    if k == 2:
        return u * u - b
    if k == 3:
        return u * (u * u - 3.0 * b)
    if k == 4:
        u2 = u * u
        return u2 * u2 - 6.0 * b * u2 + 3.0 * b * b
    raise ValueError(k)


def functional(kernel, kernel_parameter, op_x, op_y, x1, x2, y1, y2):
    """L_x L_y kappa(x, y) for row operator ``op_x`` (acting on x) and ``op_y`` (on y).

    Each monomial  dx1^m1 dx2^m2 dy1^n1 dy2^n2 kappa
        = (-1)^(m1+m2) h_{m1+n1}(u1) h_{m2+n2}(u2) kappa.
    Covers every method of src/kernels.py:8-179, e.g.
    Delta_x_Delta_y_kappa = functional(.., 'lap', 'lap', ..).
    """
    b1, b2 = kernel_scales(kernel, kernel_parameter)
    d1 = x1 - y1
    d2 = x2 - y2
    u1 = b1 * d1
    u2 = b2 * d2
    poly = 0.0
    for (m1, m2) in OPS[op_x]:
        for (n1, n2) in OPS[op_y]:
            sgn = -1.0 if (m1 + m2) % 2 else 1.0
            poly = poly + sgn * _h(m1 + n1, u1, b1) * _h(m2 + n2, u2, b2)
    return poly * kappa(kernel, kernel_parameter, d1, d2)


# name of the reference method -> (op_x, op_y); used by the functional tests
REFERENCE_METHODS = {
    "kappa": ("id", "id"),
    "D_x1_kappa": ("d1", "id"), "D_x2_kappa": ("d2", "id"), "DD_x2_kappa": ("d22", "id"),
    "D_y1_kappa": ("id", "d1"), "D_y2_kappa": ("id", "d2"), "DD_y2_kappa": ("id", "d22"),
    "D_x1_D_y1_kappa": ("d1", "d1"), "D_x1_D_y2_kappa": ("d1", "d2"),
    "D_x1_DD_y2_kappa": ("d1", "d22"), "D_x2_D_y2_kappa": ("d2", "d2"),
    "D_x2_D_y1_kappa": ("d2", "d1"), "D_x2_DD_y2_kappa": ("d2", "d22"),
    "DD_x2_DD_y2_kappa": ("d22", "d22"),
    "Delta_x_kappa": ("lap", "id"), "Delta_y_kappa": ("id", "lap"),
    "Delta_x_Delta_y_kappa": ("lap", "lap"),
    "Delta_x_D_y1_kappa": ("lap", "d1"), "Delta_x_D_y2_kappa": ("lap", "d2"),
}


# --------------------------------------------------------------------------
# Gram builders: src/Gram_matrice.py
# --------------------------------------------------------------------------

def _block_points(Xd, Xb, with_bdy):
    return np.concatenate([Xd, Xb], axis=0) if with_bdy else Xd


def block_offsets(eqn, N, Nb):
    offs, o = [], 0
    for (_, wb) in LAYOUT[eqn]:
        offs.append(o)
        o += N + (Nb if wb else 0)
    return offs, o


def _assemble(Xd, Xb, eqn, kernel, kernel_parameter):
    Xd = np.asarray(Xd, dtype=np.float64)
    Xb = np.asarray(Xb, dtype=np.float64).reshape(-1, 2)
    N, Nb = Xd.shape[0], Xb.shape[0]
    lay = LAYOUT[eqn]
    offs, M = block_offsets(eqn, N, Nb)
    Theta = np.zeros((M, M))
    for p, (opx, wbx) in enumerate(lay):
        Xp = _block_points(Xd, Xb, wbx)
        for q, (opy, wby) in enumerate(lay):
            if q < p:
                continue
            Xq = _block_points(Xd, Xb, wby)
            blk = functional(kernel, kernel_parameter, opx, opy,
                             Xp[:, None, 0], Xp[:, None, 1], Xq[None, :, 0], Xq[None, :, 1])
            Theta[offs[p]:offs[p] + Xp.shape[0], offs[q]:offs[q] + Xq.shape[0]] = blk
            if q > p:  # the reference writes the transpose of the block it evaluated
                Theta[offs[q]:offs[q] + Xq.shape[0], offs[p]:offs[p] + Xp.shape[0]] = blk.T
    return Theta


def Gram_matrix_assembly(X_domain, X_boundary, eqn="Nonlinear_elliptic", kernel="Gaussian",
                         kernel_parameter=0.2):
    """src/Gram_matrice.py:11-187.  Darcy returns (Theta_u, Theta_a)."""
    if eqn == "Darcy_flow2d":
        return (_assemble(X_domain, X_boundary, "Darcy_flow2d", kernel, kernel_parameter),
                _assemble(X_domain, X_boundary, "Darcy_flow2d_a", kernel, kernel_parameter))
    return _assemble(X_domain, X_boundary, eqn, kernel, kernel_parameter)


def _theta_test(X_test, Xd, Xb, eqn, kernel, kernel_parameter):
    X_test = np.asarray(X_test, dtype=np.float64)
    N, Nb = Xd.shape[0], Xb.shape[0]
    offs, M = block_offsets(eqn, N, Nb)
    T = np.zeros((X_test.shape[0], M))
    for q, (opy, wby) in enumerate(LAYOUT[eqn]):
        Xq = _block_points(Xd, Xb, wby)
        T[:, offs[q]:offs[q] + Xq.shape[0]] = functional(
            kernel, kernel_parameter, "id", opy,
            X_test[:, None, 0], X_test[:, None, 1], Xq[None, :, 0], Xq[None, :, 1])
    return T


def construct_Theta_test(X_test, X_domain, X_boundary, eqn="Nonlinear_elliptic", kernel="Gaussian",
                         kernel_parameter=0.2):
    """src/Gram_matrice.py:190-289.  Darcy returns (Theta_u_test, Theta_a_test)."""
    Xd = np.asarray(X_domain, dtype=np.float64)
    Xb = np.asarray(X_boundary, dtype=np.float64).reshape(-1, 2)
    if eqn == "Darcy_flow2d":
        return (_theta_test(X_test, Xd, Xb, "Darcy_flow2d", kernel, kernel_parameter),
                _theta_test(X_test, Xd, Xb, "Darcy_flow2d_a", kernel, kernel_parameter))
    return _theta_test(X_test, Xd, Xb, eqn, kernel, kernel_parameter)


def add_nugget(Theta, eqn, N, Nb, nugget, nugget_type):
    """src/PDEs.py:56-73, :250-269, :391-409; src/InverseProblems.py:66-99.
    Returns (Theta + nugget*diag(r), ratios)."""
    offs, M = block_offsets(eqn, N, Nb)
    if nugget_type == "adaptive":
        tr = []
        for p, (_, wb) in enumerate(LAYOUT[eqn]):
            n = N + (Nb if wb else 0)
            tr.append(np.trace(Theta[offs[p]:offs[p] + n, offs[p]:offs[p] + n]))
        ratios = [t / tr[-1] for t in tr[:-1]]
        r = np.ones(M)
        for p, rt in enumerate(ratios):
            r[offs[p]:offs[p] + N] = rt
        return Theta + nugget * np.diag(r), ratios
    if nugget_type == "identity":
        return Theta + nugget * np.eye(M), None
    if nugget_type == "none":
        return Theta, None
    raise ValueError(nugget_type)


def cholesky_lower(Theta):
    """jnp.linalg.cholesky (src/PDEs.py:77): symmetrise, lower factor, NaN on failure."""
    A = 0.5 * (Theta + Theta.T)
    try:
        return np.linalg.cholesky(A)
    except np.linalg.LinAlgError:
        return np.full_like(A, np.nan)


class _Solver:
    """L^{-1} b the way the reference does it (LU of the triangular factor,
    solve='lu') or by forward substitution (solve='tri')."""

    def __init__(self, L, solve="lu"):
        self.L = L
        self.mode = solve
        self.n_factorizations = 0
        if solve in ("lu", "lu_percall") and not np.all(np.isfinite(L)):
            self.mode = "nan"
        elif solve == "lu":
            self.lu = sla.lu_factor(L)
            self.n_factorizations = 1

    def fwd(self, b):
        if self.mode == "nan":
            return np.full_like(np.asarray(b, dtype=float), np.nan)
        if self.mode == "lu_percall":
            # cost model of the reference: every jnp.linalg.solve(self.L, .) call inside loss / GN_loss
            # (src/PDEs.py:86,97) refactorises L by getrf; the backward solve of the same call reuses that LU
            self.lu = sla.lu_factor(self.L)
            self.n_factorizations += 1
            return sla.lu_solve(self.lu, b)
        if self.mode == "lu":
            return sla.lu_solve(self.lu, b)
        return sla.solve_triangular(self.L, b, lower=True)

    def bwd(self, b):  # L^{-T} b
        if self.mode == "nan":
            return np.full_like(np.asarray(b, dtype=float), np.nan)
        if self.mode in ("lu", "lu_percall"):
            return sla.lu_solve(self.lu, b, trans=1)
        return sla.solve_triangular(self.L, b, lower=True, trans="T")


def int_pow(z, m):
    """z**m as repeated multiplication for integer-valued m (the oracle's chosen
    expression tree for src/PDEs.py:84); pow otherwise."""
    if float(m) == int(m) and int(m) >= 1:
        r = z
        for _ in range(int(m) - 1):
            r = r * z
        return r
    if float(m) == 0.0:
        return np.ones_like(z)
    return z ** m


# --------------------------------------------------------------------------
# problem classes: residual F(z), coefficient form of J, GN loop
# --------------------------------------------------------------------------

class _Problem:
    """Shared GN machinery.  Sub-classes give ``eqn``, ``n_blocks`` (#z blocks),
    ``F(z)`` (stacked functional values, length M) and ``jac_coeffs(z)``: a dict
    {(row_block p, z_block q): length-N coefficient vector or scalar} so that
    J[p-block rows (first N), q-block cols] = diag(coeff)."""

    eqn = None
    time_dependent = False

    def set_points(self, X_domain, X_boundary, rhs_f, bdy_g):
        self.X_domain = np.asarray(X_domain, dtype=np.float64)
        self.X_boundary = np.asarray(X_boundary, dtype=np.float64).reshape(-1, 2)
        self.N_domain = self.X_domain.shape[0]
        self.N_boundary = self.X_boundary.shape[0]
        self.rhs_f = np.asarray(rhs_f, dtype=np.float64)
        self.bdy_g = np.asarray(bdy_g, dtype=np.float64)

    def Gram_matrix(self, kernel="Gaussian", kernel_parameter=0.2, nugget=1e-8, nugget_type="adaptive"):
        Theta = _assemble(self.X_domain, self.X_boundary, self.eqn, kernel, kernel_parameter)
        self.kernel, self.kernel_parameter = kernel, kernel_parameter
        self.nugget, self.nugget_type = nugget, nugget_type
        self.Theta, self.ratio = add_nugget(Theta, self.eqn, self.N_domain, self.N_boundary, nugget, nugget_type)
        if self.ratio is not None and len(self.ratio) == 1:
            self.ratio = self.ratio[0]

    def Gram_Cholesky(self, solve="lu", structured=False):
        """structured=True: Hessian_GN is formed from the interior block of Theta^{-1} (LAPACK dpotri) and the
        diagonal coefficient form of J instead of the dense M x n solves -- the same matrix, affordable at
        N_domain ~ 10^4 (tests/test_gpu_configs.py); checked against the dense form in tests/test_oracle_golden.py."""
        self.L = cholesky_lower(self.Theta)
        self._s = _Solver(self.L, solve)
        self._Ainv = None
        if structured:
            inv, info = sla.lapack.dpotri(self.L, lower=1)
            assert info == 0
            mint = len(LAYOUT[self.eqn]) * self.N_domain
            inv = inv[:mint, :mint]
            self._Ainv = np.tril(inv) + np.tril(inv, -1).T

    # -- dense Jacobian from the coefficient form
    def jacobian(self, z):
        N = self.N_domain
        offs, M = block_offsets(self.eqn, N, self.N_boundary)
        J = np.zeros((M, self.n_blocks * N))
        idx = np.arange(N)
        for (p, q), c in self.jac_coeffs(z).items():
            J[offs[p] + idx, q * N + idx] = c
        return J

    def loss(self, z):
        s = self._s.fwd(self.F(z))
        return float(np.dot(s, s))

    def grad_loss(self, z):
        # grad(loss): J^T L^{-T} (2 L^{-1} F)   (reverse-mode through linalg.solve)
        return self.jacobian(z).T @ self._s.bwd(2.0 * self._s.fwd(self.F(z)))

    def Hessian_GN(self, z):
        # hessian(GN_loss)(z, z) = J^T L^{-T} 2 L^{-1} J   (src/PDEs.py:101-102)
        if getattr(self, "_Ainv", None) is not None:
            return self._hessian_structured(z)
        J = self.jacobian(z)
        return J.T @ self._s.bwd(2.0 * self._s.fwd(J))

    def _hessian_structured(self, z):
        N, nz = self.N_domain, self.n_blocks
        cf = self.jac_coeffs(z)
        H = np.zeros((nz * N, nz * N))
        for (p, q), c in cf.items():
            c = np.broadcast_to(np.asarray(c, dtype=float), (N,))
            for (pp, qq), cc in cf.items():
                cc = np.broadcast_to(np.asarray(cc, dtype=float), (N,))
                H[q * N:(q + 1) * N, qq * N:(qq + 1) * N] += 2.0 * (c[:, None] * self._Ainv[p * N:(p + 1) * N, pp * N:(pp + 1) * N]) * cc[None, :]
        return H

    def init_guess(self, initial_sol):
        n = self.n_blocks * self.N_domain
        if isinstance(initial_sol, str) and initial_sol == "rdm":
            return np.random.normal(0.0, 1.0, n)       # src/PDEs.py:106
        if isinstance(initial_sol, str) and initial_sol == "zero":
            return np.zeros(n)                          # src/PDEs.py:461
        return np.asarray(initial_sol, dtype=np.float64).copy()

    def GN_method(self, max_iter=3, step_size=1, initial_sol="rdm", print_hist=False):
        """src/PDEs.py:104-135 (and :309-343, :457-498; src/InverseProblems.py:153-186)."""
        sol = self.init_guess(initial_sol)
        self.init_sol = sol
        hist = [self.loss(sol)]
        if print_hist:
            print("iter = 0", "Loss =", hist[-1])
        for it in range(1, max_iter + 1):
            temp = np.linalg.solve(self.Hessian_GN(sol), self.grad_loss(sol))
            sol = sol - step_size * temp
            hist.append(self.loss(sol))
            if print_hist:
                print("iter = ", it, "Gauss-Newton step size =", step_size, " Loss = ", hist[-1])
        self.max_iter, self.step_size, self.loss_hist = max_iter, step_size, hist
        self.sol = sol
        self._finish(sol)

    def extend_sol(self, X_test):
        """src/PDEs.py:203-208."""
        T = _theta_test(X_test, self.X_domain, self.X_boundary, self.eqn, self.kernel, self.kernel_parameter)
        temp = self._s.bwd(self._s.fwd(self.sol_vec))
        self.X_test, self.N_test = X_test, np.asarray(X_test).shape[0]
        self.extended_sol = T @ temp


class Nonlinear_elliptic2d(_Problem):
    """-Delta u + alpha u^m = f  (src/PDEs.py:18-208).  z = u(interior)."""
    eqn, n_blocks = "Nonlinear_elliptic", 1

    def __init__(self, alpha=1.0, m=3):
        self.alpha, self.m = alpha, m

    def F(self, z):  # src/PDEs.py:84-85
        return np.concatenate([self.alpha * int_pow(z, self.m) - self.rhs_f, z, self.bdy_g])

    def jac_coeffs(self, z):  # src/PDEs.py:95
        return {(0, 0): self.alpha * self.m * int_pow(z, self.m - 1), (1, 0): 1.0}

    def _finish(self, sol):
        self.sol_vec = self.F(sol)
        self.sol_sampled_pts = sol

    # relaxed variant (src/PDEs.py:137-201): z = [v; w]
    def GN_relaxed_method(self, max_iter=3, step_size=1, initial_sol="rdm", pen_lambda=1e-10):
        N = self.N_domain
        sol = np.random.normal(0.0, 1.0, 2 * N) if isinstance(initial_sol, str) else np.array(initial_sol, float)
        self.init_sol = sol
        E = np.zeros((2 * N + self.N_boundary, 2 * N))
        E[np.arange(2 * N), np.arange(2 * N)] = 1.0
        A2 = E.T @ self._s.bwd(2.0 * self._s.fwd(E))

        def parts(z):
            v, w = z[:N], z[N:]
            ss = self._s.fwd(np.concatenate([v, w, self.bdy_g]))
            ss2 = -v + self.alpha * int_pow(w, self.m) - self.rhs_f
            return ss, ss2

        def loss(z):
            ss, ss2 = parts(z)
            return float(ss @ ss + (ss2 @ ss2) / pen_lambda)

        hist = [loss(sol)]
        for _ in range(max_iter):
            v, w = sol[:N], sol[N:]
            ss, ss2 = parts(sol)
            c = self.alpha * self.m * int_pow(w, self.m - 1)
            B = np.concatenate([-np.eye(N), np.diag(c)], axis=1)
            g = E.T @ self._s.bwd(2.0 * ss) + (2.0 / pen_lambda) * (B.T @ ss2)
            H = A2 + (2.0 / pen_lambda) * (B.T @ B)
            sol = sol - step_size * np.linalg.solve(H, g)
            hist.append(loss(sol))
        self.max_iter, self.step_size, self.loss_hist = max_iter, step_size, hist
        self.sol_vec = np.concatenate([sol, self.bdy_g])
        self.sol_sampled_pts = sol[N:]


class Burgers(_Problem):
    """u_t + alpha u u_x - nu u_xx = 0 (src/PDEs.py:211-350).  z = [u; u_x; u_xx]."""
    eqn, n_blocks, time_dependent = "Burgers", 3, True

    def __init__(self, alpha=1.0, nu=0.2):
        self.alpha, self.nu = alpha, nu

    def F(self, z):  # src/PDEs.py:280-287
        N = self.N_domain
        v0, v2, v3 = z[:N], z[N:2 * N], z[2 * N:]
        return np.concatenate([self.nu * v3 + self.rhs_f - self.alpha * v0 * v2, v2, v3, v0, self.bdy_g])

    def jac_coeffs(self, z):  # src/PDEs.py:297-305
        N = self.N_domain
        v0, v2 = z[:N], z[N:2 * N]
        return {(0, 0): -self.alpha * v2, (0, 1): -self.alpha * v0, (0, 2): self.nu,
                (1, 1): 1.0, (2, 2): 1.0, (3, 0): 1.0}

    def Hessian_GN(self, z):  # explicit 2 ss^T ss, src/PDEs.py:306-307
        ss = self._s.fwd(self.jacobian(z))
        return 2.0 * (ss.T @ ss)

    def _finish(self, sol):
        self.sol_vec = self.F(sol)
        self.sol_sampled_pts = sol[:self.N_domain]


class Eikonal(_Problem):
    """|grad u|^2 = f^2 + eps Delta u (src/PDEs.py:352-505).  z = [u; u_x1; u_x2]."""
    eqn, n_blocks = "Eikonal", 3

    def __init__(self, eps=3):
        self.eps = eps

    def F(self, z):  # src/PDEs.py:420-428
        N = self.N_domain
        v0, v1, v2 = z[:N], z[N:2 * N], z[2 * N:]
        v3 = -(self.rhs_f ** 2 - v1 ** 2 - v2 ** 2) / self.eps
        return np.concatenate([v1, v2, v3, v0, self.bdy_g])

    def jac_coeffs(self, z):  # src/PDEs.py:444
        N = self.N_domain
        v1, v2 = z[N:2 * N], z[2 * N:]
        return {(0, 1): 1.0, (1, 2): 1.0, (2, 1): 2 * v1 / self.eps, (2, 2): 2 * v2 / self.eps, (3, 0): 1.0}

    def _finish(self, sol):
        self.sol_vec = self.F(sol)
        self.sol_sampled_pts = sol[:self.N_domain]


class Darcy_flow2d:
    """-div(a grad u) = f inverse problem (src/InverseProblems.py:16-196).
    z = [w0 = log a; w1; w2; v0 = u; v1; v2]."""

    def set_points(self, X_domain, X_boundary, N_data, rhs_f, bdy_g):
        self.X_domain = np.asarray(X_domain, dtype=np.float64)
        self.X_boundary = np.asarray(X_boundary, dtype=np.float64).reshape(-1, 2)
        self.N_domain, self.N_boundary = self.X_domain.shape[0], self.X_boundary.shape[0]
        self.N_data = N_data
        self.X_data = self.X_domain[:N_data]
        self.rhs_f = np.asarray(rhs_f, dtype=np.float64)
        self.bdy_g = np.asarray(bdy_g, dtype=np.float64)

    def get_observation(self, data_u, noise_level):  # src/InverseProblems.py:62-64
        self.data_u = data_u + noise_level * np.random.normal(0, 1.0, np.shape(data_u)[0])
        self.noise_level = noise_level

    def Gram_matrix(self, kernel="Gaussian", kernel_parameter=0.2, nugget=1e-10, nugget_type="adaptive"):
        Tu, Ta = Gram_matrix_assembly(self.X_domain, self.X_boundary, "Darcy_flow2d", kernel, kernel_parameter)
        self.kernel, self.kernel_parameter, self.nugget, self.nugget_type = kernel, kernel_parameter, nugget, nugget_type
        self.Theta_u, self.ratio_u = add_nugget(Tu, "Darcy_flow2d", self.N_domain, self.N_boundary, nugget, nugget_type)
        self.Theta_a, self.ratio_a = add_nugget(Ta, "Darcy_flow2d_a", self.N_domain, 0, nugget, nugget_type)

    def Gram_Cholesky(self, solve="lu"):
        self.L_u, self.L_a = cholesky_lower(self.Theta_u), cholesky_lower(self.Theta_a)
        self._su, self._sa = _Solver(self.L_u, solve), _Solver(self.L_a, solve)

    def _split(self, z):
        N = self.N_domain
        return [z[k * N:(k + 1) * N] for k in range(6)]

    def F(self, z):  # src/InverseProblems.py:107-117
        w0, w1, w2, v0, v1, v2 = self._split(z)
        v3 = -v1 * w1 - v2 * w2 + (-self.rhs_f) * np.exp(-w0)
        return np.concatenate([w1, w2, w0]), np.concatenate([v1, v2, v3, v0, self.bdy_g])

    def jacobians(self, z):  # src/InverseProblems.py:140
        N, Nb = self.N_domain, self.N_boundary
        w0, w1, w2, v0, v1, v2 = self._split(z)
        idx = np.arange(N)
        Ja = np.zeros((3 * N, 6 * N))
        Ja[idx, N + idx] = 1.0
        Ja[N + idx, 2 * N + idx] = 1.0
        Ja[2 * N + idx, idx] = 1.0
        Ju = np.zeros((4 * N + Nb, 6 * N))
        Ju[idx, 4 * N + idx] = 1.0
        Ju[N + idx, 5 * N + idx] = 1.0
        Ju[2 * N + idx, idx] = (-self.rhs_f) * (-np.exp(-w0))
        Ju[2 * N + idx, N + idx] = -v1
        Ju[2 * N + idx, 2 * N + idx] = -v2
        Ju[2 * N + idx, 4 * N + idx] = -w1
        Ju[2 * N + idx, 5 * N + idx] = -w2
        Ju[3 * N + idx, 3 * N + idx] = 1.0
        return Ja, Ju

    def loss(self, z):  # src/InverseProblems.py:106-120
        Fa, Fu = self.F(z)
        ta, tu = self._sa.fwd(Fa), self._su.fwd(Fu)
        v0 = z[3 * self.N_domain:4 * self.N_domain]
        return float(ta @ ta + tu @ tu + (1 / self.noise_level ** 2) * np.sum((v0[:self.N_data] - self.data_u) ** 2))

    def grad_loss(self, z):
        Fa, Fu = self.F(z)
        Ja, Ju = self.jacobians(z)
        g = Ja.T @ self._sa.bwd(2.0 * self._sa.fwd(Fa)) + Ju.T @ self._su.bwd(2.0 * self._su.fwd(Fu))
        N = self.N_domain
        g[3 * N:3 * N + self.N_data] += (2.0 / self.noise_level ** 2) * (z[3 * N:3 * N + self.N_data] - self.data_u)
        return g

    def Hessian_GN(self, z):
        Ja, Ju = self.jacobians(z)
        H = Ja.T @ self._sa.bwd(2.0 * self._sa.fwd(Ja)) + Ju.T @ self._su.bwd(2.0 * self._su.fwd(Ju))
        N = self.N_domain
        i = 3 * N + np.arange(self.N_data)
        H[i, i] += 2.0 / self.noise_level ** 2
        return H

    def GN_method(self, max_iter=3, step_size=1, initial_sol="rdm", print_hist=False):
        N = self.N_domain
        sol = np.random.normal(0.0, 1.0, 6 * N) if isinstance(initial_sol, str) else np.array(initial_sol, float)
        self.init_sol = sol
        hist = [self.loss(sol)]
        for it in range(1, max_iter + 1):
            sol = sol - step_size * np.linalg.solve(self.Hessian_GN(sol), self.grad_loss(sol))
            hist.append(self.loss(sol))
            if print_hist:
                print("iter = ", it, "Gauss-Newton step size =", step_size, " Loss = ", hist[-1])
        self.max_iter, self.step_size, self.loss_hist, self.sol = max_iter, step_size, hist, sol
        self.sol_vec_a = np.append(sol[N:3 * N], sol[:N])   # src/InverseProblems.py:176-177
        self.sol_vec_u = self.F(sol)[1]

    def extend_sol(self, X_test):  # src/InverseProblems.py:188-196
        Tu, Ta = construct_Theta_test(X_test, self.X_domain, self.X_boundary, "Darcy_flow2d", self.kernel, self.kernel_parameter)
        self.X_test, self.N_test = X_test, np.asarray(X_test).shape[0]
        self.extended_sol_a = Ta @ self._sa.bwd(self._sa.fwd(self.sol_vec_a))
        self.extended_sol_u = Tu @ self._su.bwd(self._su.fwd(self.sol_vec_u))


# --------------------------------------------------------------------------
# sampling: src/sample_points.py (pure numpy upstream; restated, same RNG call order)
# --------------------------------------------------------------------------

def sampled_pts_rdm(N_domain, N_boundary, domain, time_dependent=False):
    """src/sample_points.py:5-48.  Uses numpy's global legacy RNG in the
    reference's call order so that seeded runs are bit-identical."""
    from numpy import random
    x1l, x1r, x2l, x2r = domain[0, 0], domain[0, 1], domain[1, 0], domain[1, 1]
    X_domain = np.concatenate((random.uniform(x1l, x1r, (N_domain, 1)),
                               random.uniform(x2l, x2r, (N_domain, 1))), axis=1)
    if not time_dependent:
        n = int(N_boundary / 4)
        Xb = np.zeros((n * 4, 2))
        Xb[0:n, 0] = random.uniform(x1l, x1r, n); Xb[0:n, 1] = x2l            # bottom
        Xb[n:2 * n, 0] = x1r; Xb[n:2 * n, 1] = random.uniform(x2l, x2r, n)    # right
        Xb[2 * n:3 * n, 0] = random.uniform(x1l, x1r, n); Xb[2 * n:3 * n, 1] = x2r  # top
        Xb[3 * n:4 * n, 1] = random.uniform(x2l, x2r, n); Xb[3 * n:4 * n, 0] = x1l  # left
    else:
        n = int(N_boundary / 3)
        Xb = np.zeros((n * 3, 2))
        Xb[0:n, 1] = random.uniform(x2l, x2r, n); Xb[0:n, 0] = x1l            # t = 0
        Xb[n:2 * n, 0] = random.uniform(x1l, x1r, n); Xb[n:2 * n, 1] = x2r    # x = right
        Xb[2 * n:, 0] = random.uniform(x1l, x1r, n); Xb[2 * n:, 1] = x2l      # x = left
    return X_domain, Xb


def notebook_sample_points(N_domain, N_boundary):
    """The notebooks' own sampler (notebooks/Nonlinear_Elliptic_Equation.ipynb:153-172):
    one (N,2) uniform draw, then bottom-x / right-y / top-x / left-y."""
    from numpy import random
    X_domain = random.uniform(0.0, 1.0, (N_domain, 2))
    Xb = np.zeros((N_boundary, 2))
    n = int(N_boundary / 4)
    Xb[0:n, 0] = random.uniform(0.0, 1.0, n)
    Xb[n:2 * n, 0] += 1
    Xb[n:2 * n, 1] = random.uniform(0.0, 1.0, n)
    Xb[2 * n:3 * n, 0] = random.uniform(0.0, 1.0, n)
    Xb[2 * n:3 * n, 1] += 1
    Xb[3 * n:4 * n, 1] = random.uniform(0.0, 1.0, n)
    return X_domain, Xb


# --------------------------------------------------------------------------
# manufactured data of the drivers
# --------------------------------------------------------------------------

def elliptic_u(x1, x2):  # main_NonLinElliptic2d.py:60-61
    return np.sin(np.pi * x1) * np.sin(np.pi * x2) + 2 * np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)


def elliptic_f(x1, x2, alpha=1.0, m=3):  # main_NonLinElliptic2d.py:62-64, -Laplace(u) analytic
    s1 = np.sin(np.pi * x1) * np.sin(np.pi * x2)
    s4 = np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)
    return 2 * np.pi ** 2 * s1 + 64 * np.pi ** 2 * s4 + alpha * int_pow(s1 + 2 * s4, m)


def burgers_bdy(x1, x2):  # main_Burgers1d.py:66-67
    return -np.sin(np.pi * x2) * (x1 == 0) + 0 * (x2 == 0)


def burgers_truth(x1, x2, nu):  # main_Burgers1d.py:87-92 (Cole-Hopf, 80-pt Gauss-Hermite)
    pts, w = np.polynomial.hermite.hermgauss(80)
    temp = x2[:, None] - np.sqrt(4 * nu * x1[:, None]) * pts[None, :]
    e = np.exp(-np.cos(np.pi * temp) / (2 * np.pi * nu))
    return -np.sum(w * np.sin(np.pi * temp) * e, axis=1) / np.sum(w * e, axis=1)


def darcy_a(x1, x2):  # main_DarcyFlow2d.py:111-113
    return np.exp(np.sin(2 * np.pi * x1) + np.sin(2 * np.pi * x2)) + np.exp(-np.sin(2 * np.pi * x1) - np.sin(2 * np.pi * x2))


def _fd_operator(N, a1, a2):
    from scipy.sparse import diags
    a_diag = np.reshape(a1[:, :N] + a1[:, 1:] + a2[:N, :] + a2[1:, :], -1)
    a_super1 = np.reshape(np.append(a1[:, 1:N], np.zeros((N, 1)), axis=1), -1)
    a_super2 = np.reshape(a2[1:N, :], -1)
    return diags([-a_super2, -a_super1[:-1], a_diag, -a_super1[:-1], -a_super2], [-N, -1, 0, 1, N],
                 shape=(N ** 2, N ** 2), format="csc")


def FD_Darcy_flow_2d(N, fun_a=darcy_a, f_val=1.0):
    """reference_solver/FD_for_Darcy_flow.py:8-32 (5-point FD, zero Dirichlet, padded)."""
    from scipy.sparse.linalg import spsolve
    hg = 1 / (N + 1)
    x_mid = (np.arange(0, N + 1, 1) + 0.5) * hg
    x_grid = (np.arange(1, N + 1, 1)) * hg
    mid, grid = np.meshgrid(x_mid, x_grid)
    a1 = np.reshape(fun_a(mid.flatten(), grid.flatten()), (N, N + 1))
    a2 = np.transpose(np.reshape(fun_a(grid.flatten(), mid.flatten()), (N, N + 1)))
    A = _fd_operator(N, a1, a2) / (hg ** 2)
    sol = spsolve(A, np.full(N * N, float(f_val)))
    out = np.zeros((N + 2, N + 2))
    out[1:N + 1, 1:N + 1] = np.reshape(sol, (N, N))
    return out


def solve_Eikonal(N, epsilon):
    """reference_solver/Cole_Hopf_for_Eikonal.py:7-36."""
    from scipy.sparse import identity
    from scipy.sparse.linalg import spsolve
    hg = 1 / (N + 1)
    x_grid = (np.arange(1, N + 1, 1)) * hg
    A = _fd_operator(N, np.ones((N, N + 1)), np.ones((N + 1, N)))
    XX, YY = np.meshgrid(x_grid, x_grid)
    f = np.zeros((N, N))
    f[0, :] += epsilon ** 2 / hg ** 2
    f[N - 1, :] += epsilon ** 2 / hg ** 2
    f[:, 0] += epsilon ** 2 / hg ** 2
    f[:, N - 1] += epsilon ** 2 / hg ** 2
    mtx = (identity(N ** 2) + (epsilon ** 2) * A / (hg ** 2)).tocsc()
    sol_v = spsolve(mtx, f.flatten())
    return XX, YY, np.reshape(-epsilon * np.log(sol_v), (N, N))
