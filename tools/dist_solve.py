"""One elliptic solve sharded over N GPUs (torchrun --nproc-per-node N tools/dist_solve.py --N 40000), or over `--virtual n`
emulated ranks on one GPU.  Prints one JSON line per repetition with the per-phase times (max over ranks).
--check: backward error of the sharded factor, and the sharded solve against the single-GPU solve on the same inputs
(small sizes)."""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=10000)
ap.add_argument("--nugget", type=float, default=1e-10)
ap.add_argument("--gn_steps", type=int, default=4)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--NB", type=int, default=512)
ap.add_argument("--Q", type=int, default=1)
ap.add_argument("--virtual", type=int, default=0)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist = None
if not a.virtual:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from nonlinpdes_gpsolver_b200 import PDEs, _lib


def u_true(x1, x2):
    return np.sin(np.pi * x1) * np.sin(np.pi * x2) + 2 * np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)


def f_rhs(x1, x2):
    s1 = np.sin(np.pi * x1) * np.sin(np.pi * x2)
    s4 = np.sin(4 * np.pi * x1) * np.sin(4 * np.pi * x2)
    w = s1 + 2 * s4
    return 2 * np.pi ** 2 * s1 + 64 * np.pi ** 2 * s4 + 1.0 * (w * w * w)


N = a.N
Nb = 4 * (math.ceil(math.sqrt(N)) + 1)
np.random.seed(0)                                   # same points / initial guess on every rank
dom = np.array([[0.0, 1.0], [0.0, 1.0]])
prob = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=u_true, rhs=f_rhs, domain=dom)
prob._engine().set_option("NB", a.NB)
prob.sampled_pts(N, Nb)
init = np.random.normal(0.0, 1.0, N)
if a.virtual:
    prob.shard(virtual_ranks=a.virtual, Q=a.Q)
else:
    prob.shard(dist, Q=a.Q)
eng = prob._engine()
M = 2 * N + prob.N_boundary


def maxred(vals):
    if dist is None:
        return list(vals)
    import torch
    t = torch.tensor(list(vals), dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


for rep in range(a.reps):
    if dist is not None:
        dist.barrier()
    eng.sync()
    eng.timer2_start()
    prob.Gram_matrix("Gaussian", 0.2, a.nugget, "adaptive")
    prob.Gram_Cholesky()
    prob.GN_method(a.gn_steps, 1, init, print_hist=False)
    total = eng.timer2_stop()
    t = prob.timings
    asm, potrf, inv, gn, tot = maxred([t["assembly_ms"], t["potrf_ms"], t["inverse_ms"], t["gn_ms"], total])
    err = np.abs(u_true(prob.X_domain[:, 0], prob.X_domain[:, 1]) - prob.sol_sampled_pts)
    if rank == 0:
        print(json.dumps(dict(world=a.virtual or world, virtual=bool(a.virtual), grid=eng.dist_info(), exchange=eng.dist_exchange_mode(), N=N, M=M, NB=a.NB, rep=rep, info=prob.chol_info,
                              asm_ms=round(asm, 3), potrf_ms=round(potrf, 2), potrf_TF=round(M ** 3 / 3 / potrf / 1e9, 2),
                              inverse_ms=round(inv, 2), inverse_TF=round(2 * M ** 3 / 3 / inv / 1e9, 2), gn_ms=round(gn, 2),
                              gn_step_ms=round(gn / a.gn_steps, 2), solve_ms=round(tot, 2), final_loss=prob.loss_hist[-1],
                              pts_L2_err=float(np.sqrt(np.mean(err ** 2))))), flush=True)

if a.check:
    L = eng.gram_download(0, 1)                        # every rank holds the whole factor
    ref = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=u_true, rhs=f_rhs, domain=dom)
    ref._eng = _lib.Engine(local)
    ref._eng.set_option("NB", a.NB)
    ref.get_sampled_points(prob.X_domain, prob.X_boundary)
    ref.Gram_matrix("Gaussian", 0.2, a.nugget, "adaptive")
    theta = ref.Theta
    ref.Gram_Cholesky()
    ref.GN_method(a.gn_steps, 1, init, print_hist=False)
    Lsingle = ref.L
    e_res = float(np.max(np.abs(L @ L.T - theta)) / (np.max(np.abs(theta)) * np.sqrt(M)))     # backward error of the sharded factor
    e_lap = float(np.max(np.abs(L - Lsingle)) / np.max(np.abs(Lsingle)))                      # vs the single-GPU factor
    e_loss = float(np.max(np.abs(np.array(prob.loss_hist) - np.array(ref.loss_hist)) / np.abs(ref.loss_hist)))
    e_sol = float(np.max(np.abs(prob.sol_sampled_pts - ref.sol_sampled_pts)) / np.max(np.abs(ref.sol_sampled_pts)))
    Xt = np.random.RandomState(1).uniform(0, 1, (64, 2))
    prob.extend_sol(Xt); ref.extend_sol(Xt)
    e_pred = float(np.max(np.abs(prob.extended_sol - ref.extended_sol)))
    ok = e_res < 1e-13 and e_lap < 1e-5 and e_loss < 1e-6 and e_sol < 1e-6 and e_pred < 1e-5
    vals = maxred([e_res, e_lap, e_loss, e_sol, e_pred, 0.0 if ok else 1.0])
    if rank == 0:
        print(json.dumps(dict(check="sharded factor: backward error and vs single-GPU factor; sharded solve vs single-GPU solve",
                              LLt_minus_Theta_rel=vals[0], L_vs_single_gpu=vals[1], loss_hist_rel=vals[2], sol_rel=vals[3], pred_abs=vals[4],
                              ok=bool(vals[5] == 0.0))), flush=True)
    assert ok, (e_res, e_lap, e_loss, e_sol, e_pred)
if dist is not None:
    eng.dist_finalize()
    dist.destroy_process_group()
