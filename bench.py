#!/usr/bin/env python
"""Benchmark of the GP-PDE Gauss-Newton hot path (BASELINE.json metric: Gauss-Newton steps/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--N_domain 40000]

A "step" is one pass of the hot path over one batch of synthetic input: Gram assembly -> Cholesky -> interior
inverse -> GNsteps Gauss-Newton steps of the nonlinear elliptic problem (BASELINE configs[4]: N_domain collocation
points, Gaussian sigma=0.2, 4 GN steps; manufactured data of main_NonLinElliptic2d.py:60-64; points from the
reference's sampler with numpy seed 0).  NOTE: the workload runs at nugget 1e-12, not the 1e-13 of configs[0]:
at N_domain >= 20 000 the device factorisation of Theta + 1e-13 * diag(r) breaks down (first non-positive pivot near
35 000 at N = 20 000 -- with every block width and with the right-looking schedule -- and 47 792 at N = 40 000), while
LAPACK dpotrf on the same N = 20 000 matrix still succeeds (profiles/r02_nugget_lapack_vs_gpu.jsonl; LAPACK cannot be run
at N = 40 000: M^2 exceeds its 32-bit indexing).  1e-12 is the smallest decade at which the device factor exists.

Every timed step goes through the public facade (solver_GP: get_sample -> solve -> collocation_pts_err) with HOST
buffers.  `e2e` is the wall clock around the K steps (host<->device copies included); `value` is the same K steps
timed on the device with CUDA events from the moment the step's inputs are resident in HBM.
N > 1 (torchrun): ONE problem sharded over the N GPUs (csrc/dist.cu), "scaling": "strong".
--impl reference: the CPU oracle (numpy/LAPACK port of the reference; JAX is not installable offline) on bounded samples.
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
from types import SimpleNamespace

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


# the dominant kernel of the sharded (right-looking) schedules is the same kernel in task-list mode, K = NB = 512:
# ncu --set full of the step-0 trailing update at N_domain=20000 (3003 block tasks, 48048 CTAs; profiles/r02_ncu_task_gemm_bulk_raw.csv):
# DRAM 11.83 GB read + 6.10 GB written; algorithmic: every C block read and written once (12.59 GB) + the 166 MB panel once
NCU_TASK_GEMM_TRAFFIC = {"bytes_per_launch": 11.835e9 + 6.096e9, "algorithmic_bytes_per_launch": 12.59e9 + 0.166e9,
                         "launch": "gemm_nt_dmma_kernel<128> task-list mode, 3003 blocks of 512x512, K=512 (N_domain=20000, step 0)",
                         "source": "profiles/r02_summary.md"}


def u_true(x1, x2):
    return np.sin(np.pi * x1) * np.sin(np.pi * x2) + 2 * np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)


def f_rhs(x1, x2):
    s1 = np.sin(np.pi * x1) * np.sin(np.pi * x2)
    s4 = np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)
    w = s1 + 2 * s4
    return 2 * np.pi ** 2 * s1 + 64 * np.pi ** 2 * s4 + 1.0 * (w * w * w)


def n_boundary_for(N):
    return 4 * (math.ceil(math.sqrt(N)) + 1)      # notebooks' rule 4*(N_pts+1), SURVEY section 8


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


def cpu_port_solve(N_sample, gn_steps, nugget, solve="lu"):
    """One full solve of the oracle (numpy/LAPACK port of the reference's algorithm, LU solves on L like
    jnp.linalg.solve) at N_sample collocation points on all host threads.  solve='lu_percall' redoes the LU of L in
    every loss / Hessian call, the reference's own cost model (src/PDEs.py:86,97)."""
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
    p.Gram_Cholesky(solve)
    t2 = time.perf_counter()
    p.GN_method(gn_steps, 1, init)
    t3 = time.perf_counter()
    Ms = 2 * N_sample + Xb.shape[0]
    return {"N_domain": N_sample, "M": Ms, "seconds": t3 - t0, "assembly_s": t1 - t0, "potrf_lu_s": t2 - t1, "gn_s": t3 - t2,
            "steps_per_s": gn_steps / (t3 - t0), "final_loss": p.loss_hist[-1], "cores": cores, "solve": solve}


def gpu_solve_at(N_sample, gn_steps, nugget, device, reps=3):
    """The single-GPU product path at a CPU-sample size (same points, same seed): GN steps/s, best of `reps`."""
    from nonlinpdes_gpsolver_b200 import PDEs, _lib
    np.random.seed(0)
    p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=u_true, rhs=f_rhs, domain=np.array([[0.0, 1.0], [0.0, 1.0]]))
    p._eng = _lib.Engine(device)
    p.sampled_pts(N_sample, n_boundary_for(N_sample))
    init = np.random.normal(0.0, 1.0, N_sample)
    best = 1e30
    for _ in range(reps + 1):
        p._eng.sync()
        t0 = time.perf_counter()
        p.Gram_matrix("Gaussian", 0.2, nugget, "adaptive")
        p.Gram_Cholesky()
        p.GN_method(gn_steps, 1, init, print_hist=False)
        p._eng.sync()
        best = min(best, time.perf_counter() - t0)
    loss = p.loss_hist[-1]
    p._eng.close()
    return {"N_domain": N_sample, "seconds": best, "steps_per_s": gn_steps / best, "final_loss": loss}


def cpu_baseline_block(sample, gn_steps, workload_M):
    scale = (workload_M / sample["M"]) ** 3
    return {
        "value": sample["steps_per_s"] / scale, "unit": "GN steps/s", "cores": sample["cores"], "kind": "port",
        "extrapolated": True,
        "sample": (f"oracle/gp_oracle.py (numpy+LAPACK port, reference-style LU solves) full solve at N_domain={sample['N_domain']} "
                   f"(M={sample['M']}) on {sample['cores']} host threads: {sample['seconds']:.2f} s (assembly {sample['assembly_s']:.2f}, "
                   f"potrf+LU {sample['potrf_lu_s']:.2f}, {gn_steps} GN steps {sample['gn_s']:.2f}) = {sample['steps_per_s']:.4f} steps/s "
                   f"MEASURED at the sample size; `value` is that EXTRAPOLATED to M={workload_M} by the O(M^3) cost ratio {scale:.1f} "
                   f"(the workload itself would take hours on the CPU; JAX unavailable offline)"),
        "measured_steps_per_s_at_sample": sample["steps_per_s"], "sample_seconds": sample["seconds"], "final_loss": sample["final_loss"],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--N_domain", type=int, default=40000)
    ap.add_argument("--gn_steps", type=int, default=4)
    ap.add_argument("--nugget", type=float, default=1e-12)   # see the module docstring: the device factor does not exist at 1e-13 at this size
    ap.add_argument("--cpu_sample_N", type=int, default=4000)    # ~12 s of CPU work on a 16-thread host
    ap.add_argument("--cpu_big_N", type=int, default=10000)      # reference arm only: one measured solve, ~3 min
    ap.add_argument("--cpu_lu_percall", action="store_true")     # reference arm: redo the LU of L per call like the reference
    ap.add_argument("--skip_cpu_baseline", action="store_true")
    ap.add_argument("--NB", type=int, default=0)
    ap.add_argument("--Q", type=int, default=0)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    N = a.N_domain
    Nb = n_boundary_for(N)
    M, n = 2 * N + Nb, N
    workload = (f"NonLinElliptic2d scaled N_domain={N} N_boundary={Nb} Gaussian sigma=0.2 nugget={a.nugget:g} GNsteps={a.gn_steps} "
                f"(BASELINE configs[4]; nugget 1e-12 instead of 1e-13: the device Cholesky breaks down at 1e-13 for N_domain >= 20000, "
                f"profiles/r02_nugget_lapack_vs_gpu.jsonl)")
    metric = "Gauss-Newton steps/sec"

    if a.impl == "reference":
        if rank != 0:
            return 0
        solve = "lu_percall" if a.cpu_lu_percall else "lu"
        for _ in range(min(a.warmup, 1)):
            cpu_port_solve(1000, a.gn_steps, a.nugget, solve)                # BLAS thread pool warm-up
        samples = [cpu_port_solve(a.cpu_sample_N, a.gn_steps, a.nugget, solve) for _ in range(max(1, a.steps))]
        t = float(np.mean([s["seconds"] for s in samples]))
        scale = (M / samples[0]["M"]) ** 3
        val = a.gn_steps / (t * scale)
        res = cpu_baseline_block(samples[-1], a.gn_steps, M)
        res["value"] = val
        big = None
        if a.cpu_big_N > a.cpu_sample_N:
            big = cpu_port_solve(a.cpu_big_N, a.gn_steps, a.nugget, solve)  # ONE measured solve at the largest size that fits the time limit
            res["measured_at_largest_size"] = {k: big[k] for k in ("N_domain", "M", "seconds", "steps_per_s", "final_loss", "solve")}
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": val, "unit": "GN steps/s", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * t * scale, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "same_config": False,
            "config": {"workload": workload, "l2_policy": "n/a (CPU)",
                       "note": (f"CPU port of the reference; each step = one full solve at N_domain={a.cpu_sample_N} (measured), value EXTRAPOLATED "
                                f"to the workload by (M/M_sample)^3 = {scale:.1f}; one additional measured solve at N_domain={a.cpu_big_N} is in "
                                "cpu_baseline.measured_at_largest_size (compare with the GPU line's same_size.N10000)")},
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
    from nonlinpdes_gpsolver_b200 import PDEs, _dist
    from nonlinpdes_gpsolver_b200.solver import solver_GP

    fp64_peak, fp64_how = measure_fp64_peak(local_rank)
    hbm, hbm_how = hbm_peak()

    def maxred(vals, op="max"):
        if dist is None:
            return list(vals)
        import torch
        t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return t.tolist()

    # N > 1: evidence that the NCCL path computes the right thing, before the timed runs (small sharded solve vs the
    # single-GPU path on this rank's own device)
    self_check = None
    if dist is not None:
        self_check = nccl_self_check(dist, local_rank, rank, maxred)

    np.random.seed(0)                             # every rank builds the same problem
    dom = np.array([[0.0, 1.0], [0.0, 1.0]])
    cfg = SimpleNamespace(alpha=1.0, m=3, kernel="Gaussian", kernel_parameter=0.2, nugget=a.nugget, nugget_type="adaptive",
                          GNsteps=a.gn_steps, step_size=1, initial_sol=None, print_hist=False)
    s = solver_GP(cfg, PDE_type="Nonlinear_elliptic")
    s.set_equation(bdy=u_true, rhs=f_rhs, domain=dom, print_option=False)
    s.auto_sample(N, Nb, print_option=False)
    cfg.initial_sol = np.random.normal(0.0, 1.0, N)
    prob = s.eqn
    eng = prob._engine()
    if a.NB:
        eng.set_option("NB", a.NB)
    grid = "single GPU"
    if dist is not None:
        prob.shard(dist, Q=a.Q or None)
        info = eng.dist_info()
        how = ("finished panels stored straight into the peers' replicated buffers over NVLink by the solve kernels (CUDA IPC) + flags"
               if eng.dist_exchange_mode() == "p2p" else "NCCL panel gathers")
        grid = f"sharded, one problem over {world} GPUs, {info['P']} x {info['Q']} block-cyclic owner-computes, {how}"
    Xd_host, Xb_host = prob.X_domain.copy(), prob.X_boundary.copy()
    truth = u_true(Xd_host[:, 0], Xd_host[:, 1])

    def barrier():
        eng.sync()
        if dist is not None:
            dist.barrier()

    def one_step():
        """One pass of the hot path through the facade with host buffers; returns the device-timed part (ms)."""
        s.get_sample(Xd_host, Xb_host, print_option=False)       # H2D: points; rhs_f / bdy_g evaluated on the host
        eng.timer2_start()                                       # inputs resident: device-timed region starts
        s.solve(print_option=False)                              # H2D: data vectors, z0, nugget; D2H: diagonal, losses, z, sol_vec
        ms = eng.timer2_stop()
        s.collocation_pts_err(truth, print_option=False)
        return ms

    for _ in range(a.warmup):
        one_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.launch_count()
    phase = {"assembly_ms": 0.0, "potrf_ms": 0.0, "inverse_ms": 0.0, "gn_ms": 0.0}
    dev_ms = 0.0
    t0 = time.perf_counter()
    for _ in range(a.steps):
        dev_ms += one_step()
        for k in phase:
            phase[k] += prob.timings[k]
    eng.sync()
    wall = time.perf_counter() - t0
    launches = eng.launch_count() - l0
    clocks = sampler.stop()
    red = maxred([dev_ms / 1e3, wall] + [phase[k] for k in ("assembly_ms", "potrf_ms", "inverse_ms", "gn_ms")])
    elapsed, e2e_elapsed = red[0], red[1]
    T_asm, T_potrf, T_inv, T_gn = [v / a.steps for v in red[2:]]
    launches_total = int(maxred([float(launches)], "sum")[0])
    value = a.gn_steps * a.steps / elapsed
    e2e_value = a.gn_steps * a.steps / e2e_elapsed
    final_loss = prob.loss_hist[-1]
    h2d = world * 8 * (2 * (N + Nb) + N + Nb + n + M)                # points, rhs_f, bdy_g, z0, nugget diag (every rank)
    d2h = world * 8 * (M + (a.gn_steps + 1) + n + M)                 # diag, losses, z, sol_vec

    if dist is not None:
        dist.barrier()
    if rank == 0:
        potrf_tf = M ** 3 / 3.0 / T_potrf / 1e9
        gemm_tf = (M ** 3) / (T_potrf + T_inv) / 1e9                # potrf + inverse: M^3 flops, all through the DMMA GEMM
        asm_bytes = 8.0 * M * (M + 1) / 2 + 16.0 * (N + Nb)
        out = {
            "metric": metric, "value": value, "unit": "GN steps/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * elapsed / a.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "M": M, "n": n, "parallelism": grid,
                       "l2_policy": "inputs larger than L2 (Theta 52 GB at N=40k); no explicit flush",
                       "timing": ("every step runs through solver_GP with host buffers; value = K steps / sum of the per-step CUDA-event intervals "
                                  "(from inputs resident to last GN step, max over ranks); e2e = K steps / wall clock around the same K steps"),
                       "algorithm": "potrf + interior inverse once, then O(n^2) Hessian assembly + n x n potrf per GN step"},
            "phases_ms": {"assembly": T_asm, "potrf": T_potrf, "inverse": T_inv, "gn_total": T_gn, "gn_per_step": T_gn / a.gn_steps},
            "roofline": {"kernel": "gemm_nt_dmma_kernel (potrf + inverse phases, M^3 algorithmic flops, aggregate over the GPUs)",
                         "bound": "tensor", "achieved": gemm_tf, "peak": fp64_peak * world, "unit": "TFLOP/s", "frac": gemm_tf / (fp64_peak * world),
                         "traffic": NCU_GEMM_TRAFFIC if world == 1 else NCU_TASK_GEMM_TRAFFIC, "peak_source": fp64_how + (f" x {world} GPUs" if world > 1 else ""),
                         "potrf_tflops": potrf_tf, "potrf_frac": potrf_tf / (fp64_peak * world),
                         "inverse_tflops": 2 * M ** 3 / 3.0 / T_inv / 1e9},
            "assembly_roofline": {"kernel": "gram_assemble_kernel" if world == 1 else "gram_rows_kernel (row-sharded)", "bound": "hbm",
                                  "achieved": asm_bytes / T_asm / 1e6, "peak": hbm * world, "unit": "GB/s",
                                  "frac": asm_bytes / T_asm / 1e6 / (hbm * world), "peak_source": hbm_how},
            "e2e": {"value": e2e_value, "unit": "GN steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": 1e3 * e2e_elapsed / a.steps, "steps": a.steps},
            "gpu_launches": launches_total, "clocks": clocks,
            "result": {"final_loss": final_loss, "pts_L2_err": float(s.pts_L2_err), "pts_max_err": float(s.pts_max_err),
                       "chol_info": prob.chol_info},
        }
        if self_check is not None:
            out["nccl_self_check"] = self_check
        if world == 1 and not a.skip_cpu_baseline:
            sample = cpu_port_solve(a.cpu_sample_N, a.gn_steps, a.nugget)
            cb = cpu_baseline_block(sample, a.gn_steps, M)
            same = gpu_solve_at(a.cpu_sample_N, a.gn_steps, a.nugget, local_rank)
            cb["same_size"] = {"N_domain": a.cpu_sample_N, "cpu_steps_per_s": sample["steps_per_s"], "gpu_steps_per_s": same["steps_per_s"],
                               "final_loss_cpu": sample["final_loss"], "final_loss_gpu": same["final_loss"]}
            cb["same_size_ratio"] = same["steps_per_s"] / sample["steps_per_s"]
            big = gpu_solve_at(a.cpu_big_N, a.gn_steps, a.nugget, local_rank, reps=2)
            cb["same_size"]["N10000"] = {"N_domain": a.cpu_big_N, "gpu_steps_per_s": big["steps_per_s"], "final_loss_gpu": big["final_loss"],
                                         "note": "CPU side: --impl reference line, cpu_baseline.measured_at_largest_size"}
            out["cpu_baseline"] = cb
        print(json.dumps(out))
    if dist is not None:
        eng.dist_finalize()
        dist.destroy_process_group()
    return 0


def nccl_self_check(dist, local_rank, rank, maxred):
    """Small sharded solve over the real NCCL path against the single-GPU path on the same inputs."""
    from nonlinpdes_gpsolver_b200 import PDEs, _lib
    res = {}
    try:
        Ns, nug, steps = 1500, 1e-8, 3
        dom = np.array([[0.0, 1.0], [0.0, 1.0]])

        def make():
            np.random.seed(5)
            p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=u_true, rhs=f_rhs, domain=dom)
            p._eng = _lib.Engine(local_rank)
            p._eng.set_option("NB", 256)
            p.sampled_pts(Ns, n_boundary_for(Ns))
            return p, np.random.normal(0.0, 1.0, Ns)

        sh, init = make()
        sh.shard(dist)
        ref, _ = make()
        for p in (sh, ref):
            p.Gram_matrix("Gaussian", 0.2, nug, "adaptive")
        theta = ref.Theta
        for p in (sh, ref):
            p.Gram_Cholesky()
            p.GN_method(steps, 1, init, print_hist=False)
        L = sh.L
        Mx = theta.shape[0]
        e_res = float(np.max(np.abs(L @ L.T - theta)) / (np.max(np.abs(theta)) * np.sqrt(Mx)))
        e_lap = float(np.max(np.abs(L - np.linalg.cholesky(theta))) / np.max(np.abs(L)))
        e_loss = float(np.max(np.abs(np.array(sh.loss_hist) - np.array(ref.loss_hist)) / np.abs(ref.loss_hist)))
        e_sol = float(np.max(np.abs(sh.sol_sampled_pts - ref.sol_sampled_pts)) / np.max(np.abs(ref.sol_sampled_pts)))
        vals = maxred([e_res, e_lap, e_loss, e_sol])
        res = {"N_domain": Ns, "world": dist.get_world_size(), "LLt_minus_Theta_rel": vals[0], "L_vs_lapack_dpotrf_rel": vals[1],
               "loss_hist_vs_single_gpu_rel": vals[2], "sol_vs_single_gpu_rel": vals[3],
               "ok": bool(vals[0] < 1e-13 and vals[2] < 5e-6 and vals[3] < 5e-6)}
        sh._eng.dist_finalize()
        sh._eng.close()
        ref._eng.close()
    except Exception as e:  # pragma: no cover
        res = {"ok": False, "error": repr(e)}
    if rank == 0:
        print("[bench] NCCL self-check: " + json.dumps(res), file=sys.stderr, flush=True)
    return res


if __name__ == "__main__":
    sys.exit(main())
