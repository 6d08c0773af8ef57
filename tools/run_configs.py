"""BASELINE.json configs C1-C4 on the GPU next to the CPU oracle (reference-style LU solves), same seeds and
RNG call order as the reference drivers.  Prints one JSON line per config: timings, final loss, errors and
the relative difference of the solution errors (the north_star parity metric)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.interpolate import griddata
from nonlinpdes_gpsolver_b200 import PDEs, InverseProblems
from oracle import gp_oracle as o

ap = argparse.ArgumentParser()
ap.add_argument("--configs", nargs="+", default=["C1", "C2", "C3", "C4"])
ap.add_argument("--no_oracle", action="store_true")
a = ap.parse_args()
DOM, DOMT = np.array([[0.0, 1.0], [0.0, 1.0]]), np.array([[0.0, 1.0], [-1.0, 1.0]])


def grid(n, lo=0.0, hi=1.0, lo2=0.0, hi2=1.0, trim=False):
    xx, yy = np.linspace(lo, hi, n), np.linspace(lo2, hi2, n)
    if trim:
        xx, yy = xx[1:-1], yy[1:-1]
    XX, YY = np.meshgrid(xx, yy)
    return np.concatenate((XX.reshape(-1, 1), YY.reshape(-1, 1)), axis=1)


def errs(truth, got):
    e = np.abs(truth - got)
    return float(np.sqrt(np.sum(e ** 2) / e.size)), float(e.max())


def timed(p, kernel, kp, nugget, steps, init):
    t0 = time.perf_counter()
    p.Gram_matrix(kernel, kp, nugget, "adaptive")
    p.Gram_Cholesky()
    p.GN_method(steps, 1, init, print_hist=False)
    p._engine().sync()
    return time.perf_counter() - t0


def report(name, desc, p, ref, t_gpu, t_cpu, gpu_err, cpu_err, steps):
    out = dict(config=name, desc=desc, gpu_solve_s=round(t_gpu, 4), gpu_steps_per_s=round(steps / t_gpu, 2), timings_ms={k: round(v, 3) for k, v in p.timings.items()},
               gpu_final_loss=p.loss_hist[-1], gpu_errors=gpu_err)
    if ref is not None:
        out.update(cpu_oracle_solve_s=round(t_cpu, 2), cpu_steps_per_s=round(steps / t_cpu, 4), cpu_final_loss=ref.loss_hist[-1], cpu_errors=cpu_err,
                   rel_diff_final_loss=abs(p.loss_hist[-1] - ref.loss_hist[-1]) / abs(ref.loss_hist[-1]),
                   rel_diff_errors={k: abs(gpu_err[k] - cpu_err[k]) / cpu_err[k] for k in gpu_err})
    print(json.dumps(out), flush=True)


for name in a.configs:
    if name == "C1":      # main_NonLinElliptic2d.py defaults; harness seed 0
        N, Nb, steps, nug = 900, 124, 4, 1e-13
        np.random.seed(0)
        p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3.0, bdy=o.elliptic_u, rhs=lambda x, y: o.elliptic_f(x, y, 1.0, 3.0), domain=DOM)
        p.sampled_pts(N, Nb)
        init = np.random.normal(0.0, 1.0, N)
        timed(p, "Gaussian", 0.2, nug, steps, init)                                   # warm-up
        t_gpu = timed(p, "Gaussian", 0.2, nug, steps, init)
        Xt = grid(60); p.extend_sol(Xt)
        truth_pts, truth_t = o.elliptic_u(p.X_domain[:, 0], p.X_domain[:, 1]), o.elliptic_u(Xt[:, 0], Xt[:, 1])
        ge = dict(zip(("pts_L2", "pts_max", "test_L2", "test_max"), errs(truth_pts, p.sol_sampled_pts) + errs(truth_t, p.extended_sol)))
        ref = ce = None; t_cpu = 0
        if not a.no_oracle:
            ref = o.Nonlinear_elliptic2d(alpha=1.0, m=3.0)
            ref.set_points(p.X_domain, p.X_boundary, p.rhs_f, p.bdy_g)
            t0 = time.perf_counter(); ref.Gram_matrix("Gaussian", 0.2, nug, "adaptive"); ref.Gram_Cholesky("lu"); ref.GN_method(steps, 1, init); t_cpu = time.perf_counter() - t0
            ref.extend_sol(Xt)
            ce = dict(zip(("pts_L2", "pts_max", "test_L2", "test_max"), errs(truth_pts, ref.sol_sampled_pts) + errs(truth_t, ref.extended_sol)))
        report(name, "NonLinElliptic2d Gaussian 0.2 nugget 1e-13 N=900 Nb=124 GN=4", p, ref, t_gpu, t_cpu, ge, ce, steps)
    elif name == "C2":    # main_Burgers1d.py defaults; seed 0
        N, Nb, steps, nug, kp = 1000, 200, 8, 1e-5, (0.3, 0.05)
        np.random.seed(0)
        p = PDEs.Burgers(alpha=1.0, nu=0.02, bdy=o.burgers_bdy, rhs=lambda x, y: 0, domain=DOMT)
        p.sampled_pts(N, Nb)
        init = np.random.normal(0.0, 1.0, 3 * N)
        timed(p, "anisotropic_Gaussian", kp, nug, steps, init)
        t_gpu = timed(p, "anisotropic_Gaussian", kp, nug, steps, init)
        Xt = grid(60, 0, 1, -1, 1); p.extend_sol(Xt)
        truth = o.burgers_truth(Xt[:, 0], Xt[:, 1], 0.02)
        ge = dict(zip(("test_L2", "test_max"), errs(truth, p.extended_sol)))
        ref = ce = None; t_cpu = 0
        if not a.no_oracle:
            ref = o.Burgers(alpha=1.0, nu=0.02)
            ref.set_points(p.X_domain, p.X_boundary, p.rhs_f, p.bdy_g)
            t0 = time.perf_counter(); ref.Gram_matrix("anisotropic_Gaussian", kp, nug, "adaptive"); ref.Gram_Cholesky("lu"); ref.GN_method(steps, 1, init); t_cpu = time.perf_counter() - t0
            ref.extend_sol(Xt)
            ce = dict(zip(("test_L2", "test_max"), errs(truth, ref.extended_sol)))
        report(name, "Burgers1d anisotropic (0.3,0.05) nugget 1e-5 N=1000 Nb=198 GN=8", p, ref, t_gpu, t_cpu, ge, ce, steps)
    elif name == "C3":    # main_Eikonal2d.py with --eps 0.01; harness seed 0; zero initial guess
        N, Nb, steps, nug, eps = 1000, 200, 8, 1e-5, 1e-2
        np.random.seed(0)
        p = PDEs.Eikonal(eps=eps, bdy=lambda x, y: 0, rhs=lambda x, y: 1, domain=DOM)
        p.sampled_pts(N, Nb)
        timed(p, "Gaussian", 0.2, nug, steps, "zero")
        t_gpu = timed(p, "Gaussian", 0.2, nug, steps, "zero")
        Xt = grid(60, trim=True); p.extend_sol(Xt)
        _, _, truth = o.solve_Eikonal(58, eps)
        ge = dict(zip(("test_L2", "test_max"), errs(truth.flatten(), p.extended_sol)))
        ref = ce = None; t_cpu = 0
        if not a.no_oracle:
            ref = o.Eikonal(eps=eps)
            ref.set_points(p.X_domain, p.X_boundary, p.rhs_f, p.bdy_g)
            t0 = time.perf_counter(); ref.Gram_matrix("Gaussian", 0.2, nug, "adaptive"); ref.Gram_Cholesky("lu"); ref.GN_method(steps, 1, "zero"); t_cpu = time.perf_counter() - t0
            ref.extend_sol(Xt)
            ce = dict(zip(("test_L2", "test_max"), errs(truth.flatten(), ref.extended_sol)))
        report(name, "Eikonal2d eps 1e-2 Gaussian 0.2 nugget 1e-5 N=1000 Nb=200 GN=8 zero init", p, ref, t_gpu, t_cpu, ge, ce, steps)
    elif name == "C4":    # main_DarcyFlow2d.py defaults; seed 9999; RNG order: points -> observation noise -> initial guess
        N, Nb, nd, steps, nug, noise = 400, 100, 60, 8, 1e-8, 1e-3
        np.random.seed(9999)
        d = InverseProblems.Darcy_flow2d(bdy=lambda x, y: 0, rhs=lambda x, y: 1, domain=DOM)
        d.sampled_pts(N, Nb, nd)
        ut = o.FD_Darcy_flow_2d(78)
        xx = np.linspace(0, 1, 80); XX, YY = np.meshgrid(xx, xx)
        data_u = griddata((XX.flatten(), YY.flatten()), ut.reshape(-1), (d.X_data[:, 0], d.X_data[:, 1]), method="linear")
        d.get_observation(data_u, noise)
        init = np.random.normal(0.0, 1.0, 6 * N)
        timed(d, "Gaussian", 0.2, nug, steps, init)
        t_gpu = timed(d, "Gaussian", 0.2, nug, steps, init)
        Xt = np.concatenate((XX.reshape(-1, 1), YY.reshape(-1, 1)), axis=1); d.extend_sol(Xt)
        a_true = o.darcy_a(Xt[:, 0], Xt[:, 1])
        ge = dict(zip(("u_L2", "u_max", "a_L2", "a_max"), errs(ut.reshape(-1), d.extended_sol_u) + errs(a_true, np.exp(d.extended_sol_a))))
        ref = ce = None; t_cpu = 0
        if not a.no_oracle:
            ref = o.Darcy_flow2d()
            ref.set_points(d.X_domain, d.X_boundary, nd, d.rhs_f, d.bdy_g)
            ref.data_u, ref.noise_level = d.data_u, noise
            t0 = time.perf_counter(); ref.Gram_matrix("Gaussian", 0.2, nug, "adaptive"); ref.Gram_Cholesky("lu"); ref.GN_method(steps, 1, init); t_cpu = time.perf_counter() - t0
            ref.extend_sol(Xt)
            ce = dict(zip(("u_L2", "u_max", "a_L2", "a_max"), errs(ut.reshape(-1), ref.extended_sol_u) + errs(a_true, np.exp(ref.extended_sol_a))))
        report(name, "DarcyFlow2d Gaussian 0.2 nugget 1e-8 N=400 Nb=100 N_data=60 noise 1e-3 GN=8", d, ref, t_gpu, t_cpu, ge, ce, steps)
