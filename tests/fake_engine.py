"""A CPU stand-in for ``_lib.Engine`` used ONLY by the CPU test-suite to exercise the host-side classes (state
machine, attribute names, RNG call order, nugget flow, Darcy's two Gram systems) without a GPU.  It implements the
Engine methods the classes call, backed by the oracle's numpy routines.  It is test infrastructure: the product never
imports it, and the GPU tests never use it."""
import numpy as np
import scipy.linalg as sla

from oracle import gp_oracle as o

_LAYOUT_OF = {"Nonlinear_elliptic": "Nonlinear_elliptic", "Burgers": "Burgers", "Eikonal": "Eikonal",
              "Darcy_flow2d": "Darcy_flow2d", "Darcy_flow2d_a": "Darcy_flow2d_a"}


class FakeEngine:
    def __init__(self):
        self.slots = {}
        self.calls = []

    # ---- plumbing
    def timer_start(self):
        pass

    def timer_stop(self):
        return 0.0

    def sync(self):
        pass

    def launch_count(self):
        return 0

    # ---- points / Gram
    def set_points(self, Xd, Xb):
        self.Xd, self.Xb = np.asarray(Xd, float), np.asarray(Xb, float).reshape(-1, 2)
        self.N, self.Nb = self.Xd.shape[0], self.Xb.shape[0]
        self.calls.append("set_points")

    def gram_assemble(self, slot, layout, kernel, kernel_parameter):
        Xb = self.Xb if layout != "Darcy_flow2d_a" else np.zeros((0, 2))
        self.slots[slot] = dict(layout=_LAYOUT_OF[layout], kernel=kernel, kp=kernel_parameter,
                                Theta=o._assemble(self.Xd, Xb, _LAYOUT_OF[layout], kernel, kernel_parameter), L=None)
        self.calls.append(f"gram_assemble[{slot}]")

    def gram_size(self, slot):
        M = self.slots[slot]["Theta"].shape[0]
        nblk = len(o.LAYOUT[self.slots[slot]["layout"]])
        return M, nblk * self.N

    def gram_get_diag(self, slot):
        return np.diag(self.slots[slot]["Theta"]).copy()

    def gram_add_diag(self, slot, add):
        self.slots[slot]["Theta"] = self.slots[slot]["Theta"] + np.diag(add)

    def gram_download(self, slot, what):
        s = self.slots[slot]
        return s["Theta"].copy() if what == 0 else s["L"].copy()

    def potrf(self, slot):
        s = self.slots[slot]
        s["L"] = o.cholesky_lower(s["Theta"])
        self.calls.append(f"potrf[{slot}]")
        return 0 if np.all(np.isfinite(s["L"])) else 1

    def inverse(self, slot):
        self.calls.append(f"inverse[{slot}]")

    def _fwd(self, slot, b):
        return sla.solve_triangular(self.slots[slot]["L"], b, lower=True)

    def _bwd(self, slot, b):
        return sla.solve_triangular(self.slots[slot]["L"], b, lower=True, trans="T")

    def solve_vec(self, slot, b):
        return self._bwd(slot, self._fwd(slot, b))

    # ---- Gauss-Newton through the oracle's problem classes (sharing this engine's Theta / L)
    def gn_setup(self, pde, params, rhs_f, bdy_g, data_u=None, noise=1.0):
        self.pde = pde
        if pde == "Darcy_flow2d":
            p = o.Darcy_flow2d()
            p.set_points(self.Xd, self.Xb, 0 if data_u is None else len(data_u), rhs_f, bdy_g)
            p.data_u, p.noise_level = np.asarray(data_u, float), noise
            p.L_u, p.L_a = self.slots[0]["L"], self.slots[1]["L"]
            p._su, p._sa = o._Solver(p.L_u, "tri"), o._Solver(p.L_a, "tri")
            self.nz = 6
        else:
            cls = {"Nonlinear_elliptic": o.Nonlinear_elliptic2d, "Nonlinear_elliptic_relaxed": o.Nonlinear_elliptic2d,
                   "Burgers": o.Burgers, "Eikonal": o.Eikonal}[pde]
            if pde.startswith("Nonlinear_elliptic"):
                p = cls(alpha=params[0], m=params[1])
            elif pde == "Burgers":
                p = cls(alpha=params[0], nu=params[1])
            else:
                p = cls(eps=params[0])
            p.set_points(self.Xd, self.Xb, rhs_f, bdy_g)
            p.L = self.slots[0]["L"]
            p._s = o._Solver(p.L, "tri")
            self.nz = {"Nonlinear_elliptic": 1, "Nonlinear_elliptic_relaxed": 2}.get(pde, 3)
        self.problem = p
        self.calls.append("gn_setup")

    def gn_set_z(self, z):
        self.z = np.array(z, float)

    def gn_get_z(self):
        return self.z.copy()

    def gn_loss(self):
        return self.problem.loss(self.z)

    def gn_step(self, step):
        p = self.problem
        self.z = self.z - step * np.linalg.solve(p.Hessian_GN(self.z), p.grad_loss(self.z))
        return p.loss(self.z)

    def gn_grad_hess(self, want_grad=True, want_hess=True):
        p = self.problem
        return (p.grad_loss(self.z) if want_grad else None), (p.Hessian_GN(self.z) if want_hess else None)

    def gn_residual(self, slot):
        F = self.problem.F(self.z)
        if self.pde == "Darcy_flow2d":
            return F[1] if slot == 0 else F[0]
        return F

    # ---- options / the sharded call surface (one fake "rank": same numbers as the single-device calls)
    def set_option(self, name, value):
        self.calls.append(f"set_option[{name}]")

    def dist_init_virtual(self, nranks):
        self.world, self.rank = int(nranks), 0
        self.calls.append(f"dist_init_virtual[{nranks}]")

    def dist_set_grid(self, P, Q):
        self.grid = (int(P), int(Q))

    def dist_info(self):
        return dict(rank=0, world=self.world, P=self.grid[0], Q=self.grid[1])

    def dist_gram_assemble(self, layout, kernel, kernel_parameter):
        self.gram_assemble(0, layout, kernel, kernel_parameter)
        self.calls.append("dist_gram_assemble")

    def dist_get_diag(self):
        return self.gram_get_diag(0)

    def dist_add_diag(self, add):
        self.gram_add_diag(0, add)

    def dist_potrf(self):
        self.calls.append("dist_potrf")
        return self.potrf(0)

    def dist_inverse(self):
        self.calls.append("dist_inverse")

    def dist_gn_step(self, step):
        self.calls.append("dist_gn_step")
        return self.gn_step(step)

    # ---- prediction
    def predict(self, slot, X_test, w):
        s = self.slots[slot]
        Xb = self.Xb if s["layout"] != "Darcy_flow2d_a" else np.zeros((0, 2))
        return o._theta_test(np.asarray(X_test, float), self.Xd, Xb, s["layout"], s["kernel"], s["kp"]) @ w
