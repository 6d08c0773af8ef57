"""Does Theta + nugget * diag(r) have an FP64 Cholesky factor?  GPU factorisation next to LAPACK dpotrf on the SAME matrix
(downloaded from the device), one JSON line per (N_domain, nugget).  Settles whether a failure at nugget 1e-13 is the
matrix (numerically indefinite in FP64) or our rounding.  Developer / evidence tool (profiles/r02_nugget_lapack_vs_gpu.jsonl)."""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.linalg as sla
from nonlinpdes_gpsolver_b200 import PDEs

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, nargs="+", default=[10000, 20000])
ap.add_argument("--nuggets", type=float, nargs="+", default=[1e-13, 1e-12])
ap.add_argument("--lapack_max", type=int, default=20000)
ap.add_argument("--variants", action="store_true", help="also try other schedules of the device factorisation (NB, right-looking)")
a = ap.parse_args()
try:
    from threadpoolctl import threadpool_limits
    threadpool_limits(limits=os.cpu_count())
except Exception:
    pass
for N in a.N:
    Nb = 4 * (math.ceil(math.sqrt(N)) + 1)
    np.random.seed(0)
    p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: 0 * x, rhs=lambda x, y: 0 * x)
    p.sampled_pts(N, Nb)
    M = 2 * N + p.N_boundary
    for ng in a.nuggets:
        p.Gram_matrix("Gaussian", 0.2, ng, "adaptive")
        rec = dict(N_domain=N, M=M, nugget=ng, sigma=0.2, trace_ratio=p.ratio)
        if N <= a.lapack_max:
            th = p.Theta                                   # the device's own matrix, symmetric, nugget included
            t0 = time.perf_counter()
            _, info = sla.lapack.dpotrf(th, lower=1, overwrite_a=1, clean=0)
            rec.update(lapack_dpotrf_info=int(info), lapack_seconds=round(time.perf_counter() - t0, 1), host_threads=os.cpu_count())
            del th
        p.Gram_Cholesky()
        rec.update(gpu_info=int(p.chol_info), gpu_potrf_ms=round(p.timings["potrf_ms"], 1))
        if a.variants:
            # the same matrix through other schedules: block width (the "panel width") and the right-looking sharded
            # schedule with one rank; near the edge the outcome depends on the rounding pattern, not on one of them being better
            var = {}
            for nb in (128, 256, 1024):
                q = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: 0 * x, rhs=lambda x, y: 0 * x)
                q._eng = p._engine()
                q._eng.set_option("NB", nb)
                q.get_sampled_points(p.X_domain, p.X_boundary)
                q.Gram_matrix("Gaussian", 0.2, ng, "adaptive")
                q.Gram_Cholesky()
                var[f"left_looking_NB{nb}"] = int(q.chol_info)
            p._engine().set_option("NB", 512)
            from nonlinpdes_gpsolver_b200 import _lib
            r = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=lambda x, y: 0 * x, rhs=lambda x, y: 0 * x)
            r._eng = _lib.Engine()
            r.get_sampled_points(p.X_domain, p.X_boundary)
            r.shard(virtual_ranks=1)
            # right-looking (sharded schedule, one rank): K = NB products per update; 'blocksum' sums them from zero and
            # subtracts once (what LAPACK's GEMM-based updates do) instead of entry-by-entry progressive subtraction
            for nb in (128, 256, 512):
                for bsum in (0, 1):
                    r._eng.set_option("NB", nb)
                    r._eng.set_option("blocksum", bsum)
                    r.Gram_matrix("Gaussian", 0.2, ng, "adaptive")
                    r.Gram_Cholesky()
                    var[f"right_looking_NB{nb}_{'blocksum' if bsum else 'progressive'}"] = int(r.chol_info)
            r._eng.close()
            rec["gpu_info_other_schedules"] = var
        rec["verdict"] = ("both fail: indefinite in FP64" if rec.get("lapack_dpotrf_info", 0) > 0 and rec["gpu_info"] > 0 else
                          "both succeed" if rec.get("lapack_dpotrf_info", 1) == 0 and rec["gpu_info"] == 0 else
                          "GPU succeeds, LAPACK fails" if rec["gpu_info"] == 0 else "GPU fails, LAPACK succeeds / not run")
        print(json.dumps(rec), flush=True)
