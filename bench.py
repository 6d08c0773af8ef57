#!/usr/bin/env python
"""Benchmark of the GP-PDE Gauss-Newton hot path (BASELINE.json metric: Gauss-Newton steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--N_domain 40000]

A "step" is one pass of the hot path over one batch of synthetic input: Gram assembly -> Cholesky ->
interior inverse -> GNsteps Gauss-Newton steps of the nonlinear elliptic problem (BASELINE configs[4],
N_domain collocation points, Gaussian sigma=0.2, nugget 1e-13, 4 GN steps; manufactured data of
main_NonLinElliptic2d.py:60-64; points from the reference's sampler with numpy seed 0).
value = GN steps completed per second, whole job, inputs resident in HBM; e2e = the same through the
public Python API with host buffers (host<->device copies inside the timed region).
N>1 (torchrun): the path is run as independent replicas, one problem per GPU (multi-GPU).
--impl reference: the CPU oracle (port of the reference; JAX is not installable here) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FP64_PEAK_FALLBACK_TFLOPS = 36.19   # cuBLAS DGEMM 8192^3 on this pool's B200 (profiles/r01_cublas_dgemm_cusolver_potrf.txt)
HBM_PEAK_FALLBACK_GBS = 6650.0      # B200_PROFILING.md fallback
# dram__bytes_read.sum + dram__bytes_write.sum of one large left-looking update launch of THIS workload (N_domain=40000,
# block column 100: 928 tiles, K=50688; algorithmic operand bytes: A rows 12.04e9 + B 0.21e9 + C read/write 0.24e9), from the
# ncu --set full capture profiles/r01_gemm_big_N40k.ncu-rep (48.2 ms, DMMA pipe 96 % of active cycles)
NCU_GEMM_TRAFFIC = {"bytes_per_launch": 13.385e9 + 0.138e9, "algorithmic_bytes_per_launch": 12.49e9,
                    "launch": "gemm_nt_dmma_kernel<128> 928 tiles K=50688 (N_domain=40000, block column 100)",
                    "source": "profiles/r01_ncu_summary.md"}


def u_true(x1, x2):
    return np.sin(np.pi * x1) * np.sin(np.pi * x2) + 2 * np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)


def f_rhs(x1, x2):
    s1 = np.sin(np.pi * x1) * np.sin(np.pi * x2)
    s4 = np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)
    w = s1 + 2 * s4
    return 2 * np.pi ** 2 * s1 + 64 * np.pi ** 2 * s4 + 1.0 * (w * w * w)


def n_boundary_for(N):
    return 4 * (math.ceil(math.sqrt(N)) + 1)      # notebooks' rule 4*(N_pts+1), SURVEY section 8


def flops_solve(M, n, gn_steps):
    """Algorithmic flops of one solve as executed by this implementation (FMA = 2):
    potrf M^3/3 + triangular inverse M^3/3 + U U^T M^3/3 + per GN step potrf(H) n^3/3."""
    return M ** 3 + gn_steps * n ** 3 / 3.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measure_fp64_peak(device):
    """cuBLAS DGEMM 8192^3 burst on this GPU: the FP64 roofline denominator (MEASURED_PEAKS.json has none)."""
    try:
        import torch
        with torch.cuda.device(device):
            n = 8192
            a = torch.randn(n, n, dtype=torch.float64, device="cuda")
            b = torch.randn(n, n, dtype=torch.float64, device="cuda")
            for _ in range(2):
                (a @ b.T)
            torch.cuda.synchronize()
            best = 1e30
            for _ in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); (a @ b.T); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            del a, b
            torch.cuda.empty_cache()
            return 2 * n ** 3 / best / 1e9, "measured live: cuBLAS DGEMM 8192^3 burst on this GPU"
    except Exception as e:  # pragma: no cover
        return FP64_PEAK_FALLBACK_TFLOPS, f"fallback {FP64_PEAK_FALLBACK_TFLOPS} (profiles/r01_cublas_dgemm_cusolver_potrf.txt): {e}"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json)"
    return HBM_PEAK_FALLBACK_GBS, "of fallback"


def cpu_port_sample(N_sample, gn_steps, nugget, workload_M, workload_n):
    """Times the oracle (numpy/LAPACK port of the reference's algorithm, incl. its LU solves on L) on a
    bounded sample and extrapolates to the workload with the reference's dense flop model."""
    from oracle import gp_oracle as o
    cores = os.cpu_count() or 1
    try:  # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host thread it can use
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    np.random.seed(0)
    Nb = n_boundary_for(N_sample)
    Xd, Xb = o.sampled_pts_rdm(N_sample, Nb, np.array([[0.0, 1.0], [0.0, 1.0]]))
    init = np.random.normal(0.0, 1.0, N_sample)
    p = o.Nonlinear_elliptic2d(alpha=1.0, m=3)
    p.set_points(Xd, Xb, f_rhs(Xd[:, 0], Xd[:, 1]), u_true(Xb[:, 0], Xb[:, 1]))
    t0 = time.perf_counter()
    p.Gram_matrix("Gaussian", 0.2, nugget, "adaptive")
    t1 = time.perf_counter()
    p.Gram_Cholesky("lu")
    t2 = time.perf_counter()
    p.GN_method(gn_steps, 1, init)
    t3 = time.perf_counter()
    Ms, ns = 2 * N_sample + Xb.shape[0], N_sample
    # reference flop model: potrf M^3/3, LU of L 2/3 M^3 (once here; the reference redoes it per
    # call), per step M^2 n (L^-1 J) + 2 M^2 n (L^-T) + M n^2 ... ; we scale the measured time by M^3.
    scale = (workload_M / Ms) ** 3
    t_sample = t3 - t0
    sps_sample = gn_steps / t_sample
    return {
        "value": sps_sample / scale, "unit": "GN steps/s", "cores": cores, "kind": "port",
        "sample": (f"oracle/gp_oracle.py (numpy+LAPACK port, reference-style LU solves) full solve at N_domain={N_sample} "
                   f"(M={Ms}) on {cores} host threads: {t_sample:.2f} s (assembly {t1 - t0:.2f}, potrf+LU {t2 - t1:.2f}, "
                   f"{gn_steps} GN steps {t3 - t2:.2f}) = {sps_sample:.4f} steps/s at the sample size; value = that "
                   f"extrapolated to M={workload_M} by the O(M^3) cost ratio {scale:.1f} (JAX unavailable offline)"),
        "measured_steps_per_s_at_sample": sps_sample, "sample_seconds": t_sample, "final_loss": p.loss_hist[-1],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N_domain", type=int, default=40000)
    ap.add_argument("--gn_steps", type=int, default=4)
    ap.add_argument("--nugget", type=float, default=1e-12)   # 1e-13 is numerically indefinite at N>=20k
    ap.add_argument("--cpu_sample_N", type=int, default=4000)    # ~15-20 s of CPU work on a 16-core host
    ap.add_argument("--skip_cpu_baseline", action="store_true")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = a.N_domain
    Nb = n_boundary_for(N)
    M, n = 2 * N + Nb, N
    workload = f"NonLinElliptic2d scaled N_domain={N} N_boundary={Nb} Gaussian sigma=0.2 nugget={a.nugget:g} GNsteps={a.gn_steps}"
    metric = "Gauss-Newton steps/sec"

    if a.impl == "reference":
        if rank != 0:
            return 0
        times = []
        res = None
        for it in range(a.warmup + a.steps):
            res = cpu_port_sample(a.cpu_sample_N, a.gn_steps, a.nugget, M, n)
            if it >= a.warmup:
                times.append(res["sample_seconds"])
        scale = (M / (2 * a.cpu_sample_N + n_boundary_for(a.cpu_sample_N))) ** 3
        t = float(np.mean(times)) if times else res["sample_seconds"]
        val = a.gn_steps / (t * scale)
        res["value"] = val
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": val, "unit": "GN steps/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t * scale, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "l2_policy": "n/a (CPU)", "note": "CPU port of the reference timed on a bounded sample, extrapolated by O(M^3)"},
            "cpu_baseline": res, "e2e": {"value": val, "unit": "GN steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }))
        return 0

    # ----------------------------- our arm -----------------------------
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from nonlinpdes_gpsolver_b200 import PDEs
    from types import SimpleNamespace
    from nonlinpdes_gpsolver_b200.solver import solver_GP

    fp64_peak, fp64_how = measure_fp64_peak(local_rank)
    hbm, hbm_how = hbm_peak()

    np.random.seed(0 + rank)                      # each replica its own problem instance
    dom = np.array([[0.0, 1.0], [0.0, 1.0]])
    prob = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=u_true, rhs=f_rhs, domain=dom)
    prob.sampled_pts(N, Nb)
    init = np.random.normal(0.0, 1.0, N)
    eng = prob._engine()

    def one_solve(p):
        p.Gram_matrix("Gaussian", 0.2, a.nugget, "adaptive")
        p.Gram_Cholesky()
        p.GN_method(a.gn_steps, 1, init, print_hist=False)

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()

    for _ in range(a.warmup):
        one_solve(prob)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    phase = {"assembly_ms": 0.0, "potrf_ms": 0.0, "inverse_ms": 0.0, "gn_ms": 0.0}
    eng.timer2_start()                             # CUDA events on the library's stream bracket the K timed solves
    for _ in range(a.steps):
        one_solve(prob)
        for k in phase:
            phase[k] += prob.timings[k]
    elapsed = eng.timer2_stop() / 1e3
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    if dist is not None:
        import torch
        tt = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed = float(tt.item())
    value = world * a.gn_steps * a.steps / elapsed
    final_loss = prob.loss_hist[-1]
    err = np.abs(u_true(prob.X_domain[:, 0], prob.X_domain[:, 1]) - prob.sol_sampled_pts)

    # e2e: through the public facade with host buffers every step (points, data vectors, initial guess in;
    # solution and sol_vec out), wall clock around the calls
    cfg = SimpleNamespace(alpha=1.0, m=3, kernel="Gaussian", kernel_parameter=0.2, nugget=a.nugget, nugget_type="adaptive",
                          GNsteps=a.gn_steps, step_size=1, initial_sol=init, print_hist=False)
    s = solver_GP(cfg, PDE_type="Nonlinear_elliptic")
    s.eqn = prob
    Xd_host, Xb_host = prob.X_domain.copy(), prob.X_boundary.copy()
    e2e_steps = min(a.steps, 2)                   # bounded: the e2e loop repeats whole solves
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        s.get_sample(Xd_host, Xb_host, print_option=False)          # H2D: points; evaluates rhs_f, bdy_g on host
        s.solve(print_option=False)                                  # H2D: data vectors, z0, nugget; D2H: diag, loss, z, sol_vec
        s.collocation_pts_err(u_true(Xd_host[:, 0], Xd_host[:, 1]), print_option=False)
    eng.sync()
    e2e_elapsed = time.perf_counter() - t0
    if dist is not None:
        tt = torch.tensor([e2e_elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_elapsed = float(tt.item())
    e2e_value = world * a.gn_steps * e2e_steps / e2e_elapsed
    h2d = 8 * (2 * (N + Nb) + N + Nb + n + M)                        # points, rhs_f, bdy_g, z0, nugget diag
    d2h = 8 * (M + (a.gn_steps + 1) + n + M)                         # diag, losses, z, sol_vec

    if dist is not None:
        dist.barrier()
    if rank == 0:
        T_potrf = phase["potrf_ms"] / a.steps
        T_inv = phase["inverse_ms"] / a.steps
        T_asm = phase["assembly_ms"] / a.steps
        T_gn = phase["gn_ms"] / a.steps
        potrf_tf = M ** 3 / 3.0 / T_potrf / 1e9
        gemm_tf = (M ** 3) / (T_potrf + T_inv) / 1e9               # potrf + inverse: M^3 flops, all through the DMMA GEMM
        asm_bytes = 8.0 * M * (M + 1) / 2 + 16.0 * (N + Nb)
        out = {
            "metric": metric, "value": value, "unit": "GN steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * elapsed / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "M": M, "n": n, "parallelism": f"replicas x{world}" if world > 1 else "single GPU",
                       "l2_policy": "inputs larger than L2 (Theta 52 GB at N=40k); no explicit flush",
                       "algorithm": "potrf + interior inverse once, then O(n^2) Hessian assembly + n x n potrf per GN step"},
            "phases_ms": {"assembly": T_asm, "potrf": T_potrf, "inverse": T_inv, "gn_total": T_gn, "gn_per_step": T_gn / a.gn_steps},
            "roofline": {"kernel": "gemm_nt_dmma_kernel (potrf + inverse phases, M^3 algorithmic flops)", "bound": "tensor",
                         "achieved": gemm_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": gemm_tf / fp64_peak,
                         "traffic": NCU_GEMM_TRAFFIC,
                         "peak_source": fp64_how, "potrf_tflops": potrf_tf, "potrf_frac": potrf_tf / fp64_peak,
                         "inverse_tflops": 2 * M ** 3 / 3.0 / T_inv / 1e9},
            "assembly_roofline": {"kernel": "gram_assemble_kernel", "bound": "hbm", "achieved": asm_bytes / T_asm / 1e6, "peak": hbm,
                                  "unit": "GB/s", "frac": asm_bytes / T_asm / 1e6 / hbm, "peak_source": hbm_how},
            "e2e": {"value": e2e_value, "unit": "GN steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_elapsed / e2e_steps, "steps": e2e_steps},
            "gpu_launches": int(launches), "clocks": clocks,
            "result": {"final_loss": final_loss, "pts_L2_err": float(np.sqrt(np.mean(err ** 2))), "pts_max_err": float(err.max()),
                       "chol_info": prob.chol_info},
        }
        if world == 1 and not a.skip_cpu_baseline:
            out["cpu_baseline"] = cpu_port_sample(a.cpu_sample_N, a.gn_steps, a.nugget, M, n)
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
