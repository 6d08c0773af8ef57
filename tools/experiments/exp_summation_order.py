"""CPU experiment (round 2): does the summation order of the Cholesky updates, or a compensated (extended-precision) diagonal, move the
smallest nugget at which Theta + nugget * diag(r) factors in FP64?  Result (N_domain = 1000 / 1500): LAPACK, a blocked right-looking
factorisation with block 64 / 128 / 512, the entry-by-entry (progressive) order and the compensated-diagonal variant all break down within a
factor 2 of each other (progressive and small blocks last longest, block 512 first); recomputing the diagonal in extended precision changes
nothing: the error that makes a pivot non-positive sits in the panel entries, not in the diagonal accumulation.
    python tools/experiments/exp_summation_order.py 1000"""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, scipy.linalg as sla
from oracle import gp_oracle as o
np.random.seed(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
import math
Nb = 4 * (math.ceil(math.sqrt(N)) + 1)
Xd, Xb = o.sampled_pts_rdm(N, Nb, np.array([[0., 1.], [0., 1.]]))
Th0 = o._assemble(Xd, Xb, "Nonlinear_elliptic", "Gaussian", 0.2)
M = Th0.shape[0]

def blocked_chol(A, NB=128, comp=False):
    """right-looking blocked Cholesky (float64); comp: before factoring a diagonal block, recompute its diagonal entries
    d_rr = theta_rr - sum_{k<j0} l_rk^2 in extended precision from the final L entries"""
    A = A.copy(); n = A.shape[0]
    diag0 = np.diag(A).astype(np.longdouble).copy()
    for j0 in range(0, n, NB):
        j1 = min(n, j0 + NB)
        if comp and j0 > 0:
            Lrow = A[j0:j1, :j0].astype(np.longdouble)
            d = diag0[j0:j1] - np.sum(Lrow * Lrow, axis=1)
            A[np.arange(j0, j1), np.arange(j0, j1)] = d.astype(np.float64)
        try:
            Ljj = np.linalg.cholesky(A[j0:j1, j0:j1])
        except np.linalg.LinAlgError:
            return j0
        A[j0:j1, j0:j1] = Ljj
        if j1 < n:
            A[j1:, j0:j1] = sla.solve_triangular(Ljj, A[j1:, j0:j1].T, lower=True).T
            A[j1:, j1:] -= A[j1:, j0:j1] @ A[j1:, j0:j1].T
    return -1

for nug in [1e-13, 5e-14, 3e-14, 2e-14, 1e-14, 5e-15, 3e-15, 2e-15, 1e-15]:
    Th, _ = o.add_nugget(Th0.copy(), "Nonlinear_elliptic", N, Xb.shape[0], nug, "adaptive")
    try:
        np.linalg.cholesky(Th); lap = "ok"
    except np.linalg.LinAlgError:
        lap = "FAIL"
    r0 = blocked_chol(Th, 128, False); r1 = blocked_chol(Th, 128, True)
    print(f"N={N} M={M} nugget={nug:g} lapack={lap} blocked={'ok' if r0<0 else 'FAIL@%d'%r0} blocked+compdiag={'ok' if r1<0 else 'FAIL@%d'%r1}", flush=True)

def rank1_chol(A):
    """unblocked right-looking: every entry receives its updates one by one (the 'progressive' order)"""
    A = A.copy(); n = A.shape[0]
    for k in range(n):
        d = A[k, k]
        if not d > 0: return k
        A[k, k] = np.sqrt(d)
        if k + 1 < n:
            A[k+1:, k] /= A[k, k]
            v = A[k+1:, k]
            A[k+1:, k+1:] -= np.outer(v, v)
    return -1
print("--- progressive (rank-1) vs blocked summation")
for nug in [3e-14, 2e-14, 1.5e-14, 1e-14, 7e-15, 5e-15]:
    Th, _ = o.add_nugget(Th0.copy(), "Nonlinear_elliptic", N, Xb.shape[0], nug, "adaptive")
    t0=time.time(); r = rank1_chol(Th); t1=time.time()
    res = {nb: blocked_chol(Th, nb, False) for nb in (64, 128, 512)}
    print(f"nugget={nug:g} rank1={'ok' if r<0 else 'FAIL@%d'%r} ({t1-t0:.0f}s) " + " ".join(f"blocked{nb}={'ok' if v<0 else 'FAIL@%d'%v}" for nb,v in res.items()), flush=True)
