"""Distributed assembly + Cholesky over N GPUs (torchrun --nproc-per-node N tools/dist_potrf.py --N 40000).
--check gathers L on rank 0 and compares it with the single-GPU factor (small sizes only)."""
import argparse, json, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from nonlinpdes_gpsolver_b200 import _lib, _dist
from nonlinpdes_gpsolver_b200.sample_points import sampled_pts_rdm

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=10000)
ap.add_argument("--nugget", type=float, default=1e-10)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--NB", type=int, default=512)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = _lib.Engine(local)
eng.set_option("NB", a.NB)
_dist.init_engine_distributed(eng, dist)

N = a.N
Nb = 4 * (math.ceil(math.sqrt(N)) + 1)
np.random.seed(0)                                   # same points on every rank
Xd, Xb = sampled_pts_rdm(N, Nb, np.array([[0.0, 1.0], [0.0, 1.0]]))
eng.set_points(Xd, Xb)
M = 2 * N + Xb.shape[0]


def nugget_vec(diag):
    tr1, tr2 = np.sum(diag[:N]), np.sum(diag[N:])
    r = np.ones(M); r[:N] = tr1 / tr2
    return a.nugget * r


for rep in range(a.reps):
    dist.barrier(); torch.cuda.synchronize()
    eng.timer_start()
    eng.dist_gram_assemble("Nonlinear_elliptic", "Gaussian", 0.2)
    t_asm = eng.timer_stop()
    eng.dist_add_diag(nugget_vec(eng.dist_get_diag()))
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    eng.timer_start()
    info = eng.dist_potrf()
    t_potrf = eng.timer_stop()
    tt = torch.tensor([t_asm, t_potrf], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    if rank == 0:
        t_asm_m, t_potrf_m = tt.tolist()
        print(json.dumps(dict(world=world, N=N, M=M, NB=a.NB, rep=rep, info=info, asm_ms=round(t_asm_m, 3),
                              asm_GBs=round(8 * M * (M + 1) / 2 / t_asm_m / 1e6, 1), potrf_ms=round(t_potrf_m, 2),
                              potrf_TF=round(M ** 3 / 3 / t_potrf_m / 1e9, 2), potrf_TF_per_gpu=round(M ** 3 / 3 / t_potrf_m / 1e9 / world, 2))), flush=True)

if a.check:
    piece = eng.dist_download_local()
    pieces = [None] * world
    dist.all_gather_object(pieces, piece)
    if rank == 0:
        L = _dist.assemble_from_locals(pieces, M, a.NB)
        ref = _lib.Engine(local)
        ref.set_option("NB", a.NB)
        ref.set_points(Xd, Xb)
        ref.gram_assemble(0, "Nonlinear_elliptic", "Gaussian", 0.2)
        ref.gram_add_diag(0, nugget_vec(ref.gram_get_diag(0)))
        assert ref.potrf(0) == 0
        Lref = ref.gram_download(0, 1)
        err = np.max(np.abs(L - Lref)) / np.max(np.abs(Lref))
        print(json.dumps(dict(check="dist L vs single-GPU L", max_rel_diff=float(err), ok=bool(err < 1e-9))), flush=True)
        assert err < 1e-9
eng.dist_finalize()
dist.destroy_process_group()
