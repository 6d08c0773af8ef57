"""Multi-GPU path (one solve sharded over the GPUs of a box).
CPU part: ownership maps and a world_size-2 gloo run of the host-side plumbing.
GPU part, ONE GPU: the task plans of n "virtual" ranks run one after the other on the device (shared storage, no NCCL) --
ownership, task lists and kernels of the sharded Cholesky / inverse / Hessian path against the single-GPU path and LAPACK.
GPU part, >= 2 GPUs: the real NCCL path (tools/dist_solve.py --check)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import gp_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DOM = np.array([[0.0, 1.0], [0.0, 1.0]])


def test_block_ownership_maps():
    from nonlinpdes_gpsolver_b200 import _dist
    nblk = 11
    for world, Q in ((1, 1), (2, 1), (2, 2), (3, 1), (4, 2), (8, 1), (8, 4), (8, 2)):
        P = world // Q
        seen = {}
        for r in range(world):
            for blk in _dist.owned_blocks(nblk, r, P, Q):
                assert blk not in seen and _dist.block_owner(blk[0], blk[1], P, Q) == r
                seen[blk] = r
        assert len(seen) == nblk * (nblk + 1) // 2              # every lower block has exactly one owner
    M, NB = 1300, 512
    for world, Q in ((1, 1), (2, 1), (3, 1), (8, 1), (4, 2)):
        P = world // Q
        rows = np.concatenate([_dist.held_rows(M, NB, p * Q, P, Q) for p in range(P)])
        assert np.array_equal(np.sort(rows), np.arange(M))
    assert np.array_equal(_dist.held_rows(M, NB, 1, 2), np.arange(512, 1024))
    assert _dist.held_rows(600, 512, 5, 8).size == 0            # more ranks than block rows
    assert _dist.grid_shape(8, 2) == (4, 2)
    with pytest.raises(ValueError):
        _dist.grid_shape(8, 3)


def test_task_plans_are_consistent_on_the_host():
    """The real plan builder of csrc/dist.cu, run on the CPU through the C ABI (no GPU): for every step of the sharded
    Cholesky and of U = L^-T each block is solved / updated by exactly one rank -- its owner -- with the right K range and
    extents; every Hessian block has exactly one owner with its four Theta^-1 sub-block tasks."""
    from nonlinpdes_gpsolver_b200 import _lib
    for (n, NB) in ((1300, 128), (2048, 256), (5000, 512), (700, 128)):
        for (P, Q) in ((1, 1), (2, 1), (3, 1), (2, 2), (8, 1), (2, 4), (4, 2), (1, 8)):
            for phase in (0, 1):
                assert _lib.dist_plan_check(n, NB, P, Q, phase) == 0, (n, NB, P, Q, phase)
            assert _lib.dist_plan_check(n, NB, P, Q, 2, nb_extra=96) == 0, (n, NB, P, Q)
    assert _lib.dist_plan_check(1000, 100, 2, 1, 0) == -1            # block size must be a multiple of 64


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GPP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from nonlinpdes_gpsolver_b200 import _dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
# owner-computes + gather on replicated storage, the structure of csrc/dist.cu, with numpy blocks and gloo:
# right-looking Cholesky of a small SPD matrix, block (bi, bc) updated only by its owner, panels gathered
n, NB, P, Q = 96, 16, world, 1
nblk = n // NB
rng = np.random.RandomState(0)
B = rng.standard_normal((n, n)); S = B @ B.T + n * np.eye(n)
A = S.copy()
blk = lambda M, i, j: M[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB]
for j in range(nblk):
    root = _dist.block_owner(j, j, P, Q)
    d = torch.from_numpy(np.linalg.cholesky(blk(A, j, j)) if rank == root else np.zeros((NB, NB)))
    dist.broadcast(d, src=root)
    blk(A, j, j)[:] = d.numpy()
    for bi in range(j + 1, nblk):
        own = _dist.block_owner(bi, j, P, Q)
        t = torch.from_numpy(np.linalg.solve(d.numpy(), blk(A, bi, j).T).T.copy() if rank == own else np.zeros((NB, NB)))
        dist.broadcast(t, src=own)                       # panel gather
        blk(A, bi, j)[:] = t.numpy()
    for (bi, bc) in _dist.owned_blocks(nblk, rank, P, Q):
        if bc > j:
            blk(A, bi, bc)[:] -= blk(A, bi, j) @ blk(A, bc, j).T
L = np.tril(A)
assert np.max(np.abs(L - np.linalg.cholesky(S))) < 1e-10
# the unique-id exchange pattern of init_engine_distributed
box = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
assert box[0] == bytes(range(128))
# max-over-ranks timing reduction used by bench.py / tools/dist_solve.py
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == float(world)
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_gloo_world2_host_logic(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_GLOO_WORKER)
    env = dict(os.environ, GPP_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def _problem(PDEs, N, Nb, NB, seed=3):
    np.random.seed(seed)
    p = PDEs.Nonlinear_elliptic2d(alpha=1.0, m=3, bdy=o.elliptic_u, rhs=o.elliptic_f, domain=DOM)
    p._engine().set_option("NB", NB)
    p.sampled_pts(N, Nb)
    return p, np.random.normal(0.0, 1.0, N)


@pytest.mark.gpu
@pytest.mark.parametrize("nranks,Q,N,Nb,NB", [
    (1, 1, 500, 60, 128),       # world = 1 goes through the same right-looking code
    (2, 1, 700, 100, 128),
    (3, 1, 700, 100, 128),      # ragged: 12 block rows over 3 ranks, last block partial
    (4, 2, 700, 100, 128),      # 2 x 2 grid
    (8, 1, 900, 124, 128),
    (8, 4, 900, 124, 128),      # 2 x 4 grid
    (2, 1, 1500, 160, 512),     # NB = 512: recursion depth of the panel solve, 128-wide GEMM tiles
])
def test_virtual_sharded_solve_matches_single_gpu(nranks, Q, N, Nb, NB):
    """The sharded path (plans of `nranks` ranks emulated on one GPU) against the single-GPU path on the same inputs:
    factor, loss history, solution, prediction; and the factor against Theta (backward error)."""
    from nonlinpdes_gpsolver_b200 import PDEs
    nug, steps = 1e-8, 3
    ref, init = _problem(PDEs, N, Nb, NB)
    ref.Gram_matrix("Gaussian", 0.2, nug, "adaptive")
    theta = ref.Theta
    ref.Gram_Cholesky()
    ref.GN_method(steps, 1, init, print_hist=False)
    sh, init2 = _problem(PDEs, N, Nb, NB)
    assert np.array_equal(init, init2)
    sh.shard(virtual_ranks=nranks, Q=Q)
    assert sh._engine().dist_info() == dict(rank=0, world=nranks, P=nranks // Q, Q=Q)
    sh.Gram_matrix("Gaussian", 0.2, nug, "adaptive")
    assert sh.ratio == ref.ratio
    sh.Gram_Cholesky()
    assert sh.chol_info == 0
    L, M = sh.L, theta.shape[0]
    assert np.max(np.abs(L @ L.T - theta)) <= 1e-13 * np.max(np.abs(theta)) * np.sqrt(M)
    # the two paths round differently (tiled / left-looking single-GPU kernels vs right-looking sharded schedule); at
    # nugget 1e-8 that is amplified to ~1e-7 in the loss (band: 10 eps / nugget = 2e-7 on the solution error)
    assert np.max(np.abs(L - ref.L)) <= 1e-5 * np.max(np.abs(L))
    sh.GN_method(steps, 1, init, print_hist=False)
    np.testing.assert_allclose(sh.loss_hist, ref.loss_hist, rtol=1e-6)
    np.testing.assert_allclose(sh.sol_sampled_pts, ref.sol_sampled_pts, atol=1e-6 * np.max(np.abs(ref.sol_sampled_pts)))
    Xt = np.random.RandomState(1).uniform(0, 1, (40, 2))
    sh.extend_sol(Xt)
    ref.extend_sol(Xt)
    np.testing.assert_allclose(sh.extended_sol, ref.extended_sol, atol=1e-6)
    # a second solve on the same handle reuses the cached plans and buffers
    sh.Gram_matrix("Gaussian", 0.2, nug, "adaptive")
    sh.Gram_Cholesky()
    sh.GN_method(steps, 1, init, print_hist=False)
    np.testing.assert_allclose(sh.loss_hist, ref.loss_hist, rtol=1e-6)


@pytest.mark.gpu
def test_virtual_sharded_factor_failure_is_reported():
    from nonlinpdes_gpsolver_b200 import PDEs
    p, _ = _problem(PDEs, 600, 80, 128)
    p.shard(virtual_ranks=4, Q=2)
    p.Gram_matrix("Gaussian", 0.2, 0.0, "none")          # no nugget: numerically indefinite
    p.Gram_Cholesky()
    assert p.chol_info > 0


@pytest.mark.gpu
@pytest.mark.parametrize("Q,p2p", [(1, "0"), (2, "0"), (1, "1"), (2, "1")])
def test_sharded_solve_two_gpus(Q, p2p):
    """The real exchange paths on 2 GPUs: NCCL all-gather / broadcast (default) and peer stores over NVLink (GPP_DIST_P2P=1)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29612", os.path.join(ROOT, "tools", "dist_solve.py"), "--N", "1500", "--NB", "256", "--reps", "1",
                          "--nugget", "1e-8", "--Q", str(Q), "--check"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, GPP_DIST_P2P=p2p))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"ok": true' in out.stdout
    assert ('"exchange": "p2p"' if p2p == "1" else '"exchange": "nccl"') in out.stdout


@pytest.mark.gpu
def test_sharded_solve_single_rank_nccl():
    """world = 1 through torchrun: communicator creation and the whole sharded call sequence on the driver's 1-GPU box."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", os.path.join(ROOT, "tools", "dist_solve.py"), "--N", "1200", "--NB", "256", "--reps", "1",
                          "--nugget", "1e-8", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"ok": true' in out.stdout
