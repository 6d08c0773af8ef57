"""Multi-GPU path.  CPU part: world_size-2 gloo run of the host-side logic (ownership map, gathering the
row-sharded factor, max-over-ranks timing reduction).  GPU part (needs >= 2 GPUs): distributed assembly +
Cholesky over NCCL against the single-GPU factor."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_ownership_map():
    from nonlinpdes_gpsolver_b200 import _dist
    M, NB = 1300, 512
    for world in (1, 2, 3, 8):
        seen = np.concatenate([_dist.local_row_map(M, NB, r, world) for r in range(world)])
        assert np.array_equal(np.sort(seen), np.arange(M))
    assert np.array_equal(_dist.local_row_map(M, NB, 1, 2), np.arange(512, 1024))
    assert np.array_equal(_dist.local_row_map(M, NB, 0, 2), np.r_[np.arange(0, 512), np.arange(1024, 1300)])
    assert _dist.local_row_map(600, 512, 5, 8).size == 0            # more ranks than block rows


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GPP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from nonlinpdes_gpsolver_b200 import _dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
M, NB = 700, 128
rng = np.random.RandomState(0)
A = np.tril(rng.standard_normal((M, M)))
piece = A[_dist.local_row_map(M, NB, rank, world)]            # what gpp_dist_download_local returns
pieces = [None] * world
dist.all_gather_object(pieces, piece)
B = _dist.assemble_from_locals(pieces, M, NB)
assert np.array_equal(A, B)
# the unique-id exchange pattern of init_engine_distributed
box = [bytes(range(128)) if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
assert box[0] == bytes(range(128))
# max-over-ranks timing reduction used by bench.py / tools/dist_potrf.py
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


@pytest.mark.gpu
def test_dist_potrf_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29612", os.path.join(ROOT, "tools", "dist_potrf.py"), "--N", "1500", "--NB", "256", "--reps", "1",
                          "--nugget", "1e-8", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"ok": true' in out.stdout


@pytest.mark.gpu
def test_dist_potrf_single_rank_matches():
    """world = 1 exercises the row-panel assembly and the distributed code path without NCCL traffic."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=1", "--master-addr", "127.0.0.1",
                          "--master-port", "29613", os.path.join(ROOT, "tools", "dist_potrf.py"), "--N", "1200", "--NB", "256", "--reps", "1",
                          "--nugget", "1e-8", "--check"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert '"ok": true' in out.stdout
