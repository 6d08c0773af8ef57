// Blocked FP64 Cholesky, triangular solves and the interior block of the inverse.
//
// Replaces jnp.linalg.cholesky (src/PDEs.py:75-80, :271-276, :411-416;
// src/InverseProblems.py:101-103) and the jnp.linalg.solve(L, .) calls of the GN
// loop (src/PDEs.py:86,97,118; ...).  All O(n^3) work is routed through the
// TMA + DMMA kernel of gemm_dmma.cu; the kernels in this file are the small
// O(n^2 nb) pieces: 64x64 diagonal factorisations, 64-wide triangular solves by
// substitution (no explicit inverses: nuggets down to 1e-13 leave no slack) and
// the vector triangular solves.
#include "gpp_internal.cuh"

#include <cstdlib>

namespace {

constexpr int BASE = 64;
constexpr int LDS_PAD = BASE + 1;

// ---------------------------------------------------------------------------
// potrf of one diagonal block (n <= 64): one thread per row, the row lives in registers.
// Right-looking over columns j = 0..63 (fully unrolled): thread j publishes sqrt(a_jj), every thread i > j
// publishes l_ij = a_ij / l_jj, then a_ik -= l_ij l_kj for k > j.  Entry (i, k) therefore receives its updates
// one by one in column order, each rounded relative to the remainder (same rounding model as the GEMM
// updates -- and the same operation sequence as a left-looking dot product started at a_ik).
// A non-positive / NaN pivot is recorded once in *info (1-based global index);
// sqrt then produces NaN which propagates, like jnp.linalg.cholesky's NaN output.
// Blocks with n < 64 are padded with the identity.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(BASE)
potrf_base_kernel(double* __restrict__ A, long ld, int n, int gidx0, int* info) {
  __shared__ double stage[BASE * LDS_PAD];
  __shared__ double col[BASE];
  __shared__ double sdiag;
  const int i = threadIdx.x;
  for (int r = 0; r < n; ++r)         // coalesced row loads
    if (i <= r) stage[r * LDS_PAD + i] = A[(long)r * ld + i];
  __syncthreads();
  double a[BASE];
#pragma unroll
  for (int k = 0; k < BASE; ++k) a[k] = (i < n && k <= i) ? stage[i * LDS_PAD + k] : (k == i ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < BASE; ++j) {
    if (i == j) {
      const double d = a[j];
      if (!(d > 0.0) && j < n) atomicCAS(info, 0, gidx0 + j + 1);
      const double sq = sqrt(d);
      a[j] = sq;
      sdiag = sq;
    }
    __syncthreads();
    double l = 0.0;
    if (i > j) { l = a[j] / sdiag; a[j] = l; }
    col[i] = l;
    __syncthreads();
#pragma unroll
    for (int k = j + 1; k < BASE; ++k) a[k] = fma(-l, col[k], a[k]);   // entries k > i are never read
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < BASE; ++k) stage[i * LDS_PAD + k] = a[k];
  __syncthreads();
  for (int r = 0; r < n; ++r)
    if (i <= r) A[(long)r * ld + i] = stage[r * LDS_PAD + i];
}

// ---------------------------------------------------------------------------
// X * L^T = P  for an nb x nb (nb <= 64) lower-triangular L, P is rows x nb, in place.
// One thread per row, the row lives in registers; substitution along the row, right-looking and fully unrolled:
//   x_j = p_j / L_jj, then p_k -= x_j L_kj for k > j   (the same operation sequence per entry as
//   x_j = (p_j - sum_{k<j} x_k L_jk) / L_jj evaluated term by term).
// L is held transposed in shared memory (a column of L is contiguous: broadcast 128-bit loads); it is padded
// with the identity for nb < 64.  blockIdx.x -> 128 rows: rows are either contiguous (row_stride_blk == 0) or
// block-cyclic: logical row block b of row_nb rows sits at physical rows (row_first_blk + b * row_stride_blk) * row_nb.
// ---------------------------------------------------------------------------
constexpr int TRSM_ROWS = 128;
__device__ __forceinline__ long trsm_phys_row(const TrsmRows& m, int r) {
  if (m.stride_blk == 0) return r;
  return (long)(m.first_blk + (r / m.nb) * m.stride_blk) * m.nb + r % m.nb;
}
__global__ void __launch_bounds__(TRSM_ROWS)
trsm_base_kernel(double* __restrict__ P, long ldp, const TrsmRows rm, const double* __restrict__ L, long ldl, int nb) {
  extern __shared__ double sm[];
  double* sLt = sm;                      // [64][64] transposed: sLt[j * 64 + k] = L[k][j]
  double* sP = sm + BASE * BASE;         // TRSM_ROWS x LDS_PAD
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TRSM_ROWS;
  for (int e = tid; e < BASE * BASE; e += blockDim.x) {
    const int k = e / BASE, j = e % BASE;          // L[k][j], coalesced over j
    double v = (k == j) ? 1.0 : 0.0;
    if (k < nb && j < nb) v = (j <= k) ? L[(long)k * ldl + j] : 0.0;
    sLt[j * BASE + k] = v;
  }
  const int nrow = min(TRSM_ROWS, rm.rows - r0);
  for (int e = tid; e < nrow * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    sP[i * LDS_PAD + j] = P[trsm_phys_row(rm, r0 + i) * ldp + j];
  }
  __syncthreads();
  if (tid < nrow) {
    double x[BASE];
#pragma unroll
    for (int k = 0; k < BASE; ++k) x[k] = (k < nb) ? sP[tid * LDS_PAD + k] : 0.0;
#pragma unroll
    for (int j = 0; j < BASE; ++j) {
      const double* lc = sLt + j * BASE;
      const double xj = x[j] / lc[j];
      x[j] = xj;
#pragma unroll
      for (int k = j + 1; k < BASE; ++k) x[k] = fma(-xj, lc[k], x[k]);
    }
#pragma unroll
    for (int k = 0; k < BASE; ++k) if (k < nb) sP[tid * LDS_PAD + k] = x[k];
  }
  __syncthreads();
  for (int e = tid; e < nrow * nb; e += blockDim.x) {
    const int i = e / nb, j = e % nb;
    P[trsm_phys_row(rm, r0 + i) * ldp + j] = sP[i * LDS_PAD + j];
  }
}

// ---------------------------------------------------------------------------
// Tiled right-looking Cholesky of a whole (small) matrix in ONE persistent kernel: n <= a few thousand, 64 x 64 tiles.
// The blocked path above needs ~50 dependent launches per 512-wide block column; for the latency-bound configurations
// (BASELINE C1-C4: M ~ 2000-4200, n ~ 900-3000) and for the diagonal blocks of the large factorisations that launch
// chain IS the run time.  Here all CTAs are resident and meet at a global barrier (atomic counter, release / acquire):
//   step k:  B(k): tiles (i, k), i > k, solved against the factored diagonal tile  X L_kk^T = A_ik     | barrier
//            C(k): tiles (i, j), k < j <= i:  A_ij -= L_ik L_jk^T (entry by entry in k order: progressive
//                  accumulation); the CTA that updates tile (k+1, k+1) factors it right away  | barrier
// Tile data goes through L2 (ld.cg) because other CTAs rewrite it between phases.
// ---------------------------------------------------------------------------
constexpr int TT = 64;             // tile edge
constexpr int TPAD = 66;           // k-major operand tiles: row stride (doubles), 16-byte aligned rows
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire_u32(bar) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// Factor the tile held row-wise in shared memory (st[r * LDS_PAD + c], lower part valid, nb x nb; rows >= nb are treated
// as identity rows).  Threads 0..63 own one row each; called by the whole CTA (block barriers inside).  The 64 columns are
// processed in two passes of 32 so that a thread keeps 32 doubles in registers; per entry the operation sequence is that
// of potrf_base_kernel (updates applied one by one in column order).
__device__ __forceinline__ void tile_potrf_smem(double* st, double* col, double* sdiag, int nb, int gidx0, int* info) {
  const int i = threadIdx.x;
  if (i < TT) {
    // only the two warps that own the 64 rows synchronise inside (named barrier 1); the rest of the CTA waits below
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int c0 = 32 * pass;
      double a[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) a[k] = (i < nb && c0 + k <= i) ? st[i * LDS_PAD + c0 + k] : (c0 + k == i ? 1.0 : 0.0);
      if (pass == 1) {
        // entries (i, 32..63) receive the updates of columns 0..31 first
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const double lij = (i < nb) ? st[i * LDS_PAD + j] : 0.0;
#pragma unroll
          for (int k = 0; k < 32; ++k) a[k] = fma(-lij, st[(32 + k) * LDS_PAD + j], a[k]);
        }
      }
#pragma unroll
      for (int jj = 0; jj < 32; ++jj) {
        const int j = c0 + jj;
        if (i == j) {
          const double d = a[jj];
          if (!(d > 0.0) && j < nb) atomicCAS(info, 0, gidx0 + j + 1);
          const double sq = sqrt(d);
          a[jj] = sq;
          *sdiag = sq;
        }
        asm volatile("bar.sync 1, 64;" ::: "memory");
        double l = 0.0;
        if (i > j) { l = a[jj] / *sdiag; a[jj] = l; }
        col[i] = l;
        asm volatile("bar.sync 1, 64;" ::: "memory");
#pragma unroll
        for (int k = jj + 1; k < 32; ++k) a[k] = fma(-l, col[c0 + k], a[k]);
      }
#pragma unroll
      for (int k = 0; k < 32; ++k) st[i * LDS_PAD + c0 + k] = a[k];
      asm volatile("bar.sync 1, 64;" ::: "memory");
    }
  }
  __syncthreads();
}

// rows of st (row-major, threads 0..63 one row each) solved against the transposed factor in sLt (sLt[j * TPAD + k] =
// L[k][j], identity-padded): X L^T = P in place, two passes of 32 columns; operation sequence of trsm_base_kernel
__device__ __forceinline__ void tile_trsm_smem(double* st, const double* sLt) {
  const int tid = threadIdx.x;
  if (tid >= TT) return;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const int c0 = 32 * pass;
    double x[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) x[k] = st[tid * LDS_PAD + c0 + k];
    if (pass == 1) {
#pragma unroll 4
      for (int j = 0; j < 32; ++j) {
        const double xj = st[tid * LDS_PAD + j];
        const double* lc = sLt + j * TPAD + 32;
#pragma unroll
        for (int k = 0; k < 32; ++k) x[k] = fma(-xj, lc[k], x[k]);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
      const double* lc = sLt + (c0 + jj) * TPAD + c0;
      const double xj = x[jj] / lc[jj];
      x[jj] = xj;
#pragma unroll
      for (int k = jj + 1; k < 32; ++k) x[k] = fma(-xj, lc[k], x[k]);
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) st[tid * LDS_PAD + c0 + k] = x[k];
  }
}

__global__ void __launch_bounds__(256, 1)
potrf_tiled_kernel(double* __restrict__ A, long ld, int n, int gidx0, int* info, unsigned* bar) {
  extern __shared__ double tsm[];
  double* sAt = tsm;                       // [TT][TPAD]  operand tile, k-major: sAt[kk * TPAD + r]
  double* sBt = tsm + TT * TPAD;           // [TT][TPAD]
  double* st = tsm + 2 * TT * TPAD;        // [TT][LDS_PAD] row-major tile (diagonal factor / solve staging)
  double* col = st + TT * LDS_PAD;         // [TT]
  double* sdiag = col + TT;
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const int nt = (n + TT - 1) / TT;
  unsigned phase = 0;
  auto rows_of_tile = [&](int t) { return min(TT, n - t * TT); };

  // tile (0, 0) is factored by CTA 0 before the first barrier
  if (blockIdx.x == 0) {
    const int nb = rows_of_tile(0);
    for (int e = tid; e < TT * TT; e += 256) {
      const int r = e / TT, c = e % TT;
      st[r * LDS_PAD + c] = (r < nb && c <= r) ? __ldcg(A + (long)r * ld + c) : 0.0;
    }
    __syncthreads();
    tile_potrf_smem(st, col, sdiag, nb, gidx0, info);
    for (int e = tid; e < TT * TT; e += 256) {
      const int r = e / TT, c = e % TT;
      if (r < nb && c <= r) A[(long)r * ld + c] = st[r * LDS_PAD + c];
    }
  }
  grid_barrier(bar, (++phase) * G);

  for (int k = 0; k < nt; ++k) {
    const int k0 = k * TT, nbk = rows_of_tile(k);
    // ---- B(k): solve the tiles below the diagonal tile
    if (k + 1 < nt) {
      // L_kk transposed into sAt (sAt[j * TPAD + kk] = L[kk][j]), padded with the identity
      bool have_l = false;
      for (int i = k + 1 + blockIdx.x; i < nt; i += G) {
        if (!have_l) {
          for (int e = tid; e < TT * TT; e += 256) {
            const int r = e / TT, c = e % TT;      // L[r][c]
            double v = (r == c) ? 1.0 : 0.0;
            if (r < nbk && c < nbk) v = (c <= r) ? __ldcg(A + (long)(k0 + r) * ld + k0 + c) : 0.0;
            sAt[c * TPAD + r] = v;
          }
          have_l = true;
        }
        const int i0 = i * TT, nbi = rows_of_tile(i);
        for (int e = tid; e < TT * TT; e += 256) {
          const int r = e / TT, c = e % TT;
          st[r * LDS_PAD + c] = (r < nbi && c < nbk) ? __ldcg(A + (long)(i0 + r) * ld + k0 + c) : 0.0;
        }
        __syncthreads();
        tile_trsm_smem(st, sAt);
        __syncthreads();
        for (int e = tid; e < TT * TT; e += 256) {
          const int r = e / TT, c = e % TT;
          if (r < nbi && c < nbk) A[(long)(i0 + r) * ld + k0 + c] = st[r * LDS_PAD + c];
        }
        __syncthreads();
      }
    }
    grid_barrier(bar, (++phase) * G);
    if (k + 1 >= nt) break;
    // ---- C(k): trailing update; linear index t over the lower-triangular tile set starting at (k+1, k+1)
    const int m = nt - k - 1;
    const long ntile = (long)m * (m + 1) / 2;
    for (long t = blockIdx.x; t < ntile; t += G) {
      int ii = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
      while ((long)(ii + 1) * (ii + 2) / 2 <= t) ++ii;
      while ((long)ii * (ii + 1) / 2 > t) --ii;
      const int jj = (int)(t - (long)ii * (ii + 1) / 2);
      const int ti = k + 1 + ii, tj = k + 1 + jj;
      const int i0 = ti * TT, j0 = tj * TT, nbi = rows_of_tile(ti), nbj = rows_of_tile(tj);
      // operand tiles, k-major
      for (int e = tid; e < TT * TT; e += 256) {
        const int r = e / TT, c = e % TT;
        sAt[c * TPAD + r] = (r < nbi && c < nbk) ? __ldcg(A + (long)(i0 + r) * ld + k0 + c) : 0.0;
        sBt[c * TPAD + r] = (r < nbj && c < nbk) ? __ldcg(A + (long)(j0 + r) * ld + k0 + c) : 0.0;
      }
      const int ty = tid >> 4, tx = tid & 15;
      double acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int gr = ty * 4 + r, gc = tx * 4 + c;
          acc[r][c] = (gr < nbi && gc < nbj) ? __ldcg(A + (long)(i0 + gr) * ld + j0 + gc) : 0.0;
        }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < TT; ++kk) {
        const double2 a01 = *reinterpret_cast<const double2*>(sAt + kk * TPAD + ty * 4);
        const double2 a23 = *reinterpret_cast<const double2*>(sAt + kk * TPAD + ty * 4 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(sBt + kk * TPAD + tx * 4);
        const double2 b23 = *reinterpret_cast<const double2*>(sBt + kk * TPAD + tx * 4 + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fma(-a[r], b[c], acc[r][c]);
      }
      if (ti == tj && ii == 0) {
        // the next diagonal tile: factor it now (its column is solved right after the barrier)
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) st[(ty * 4 + r) * LDS_PAD + tx * 4 + c] = acc[r][c];
        __syncthreads();
        tile_potrf_smem(st, col, sdiag, nbi, gidx0 + i0, info);
        for (int e = tid; e < TT * TT; e += 256) {
          const int r = e / TT, c = e % TT;
          if (r < nbi && c <= r) A[(long)(i0 + r) * ld + j0 + c] = st[r * LDS_PAD + c];
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int gr = ty * 4 + r, gc = tx * 4 + c;
            if (gr < nbi && gc < nbj && (ti != tj || gc <= gr)) A[(long)(i0 + gr) * ld + j0 + gc] = acc[r][c];
          }
      }
      __syncthreads();
    }
    grid_barrier(bar, (++phase) * G);
  }
}
constexpr int TILED_SMEM = (2 * TT * TPAD + TT * LDS_PAD + TT + 8) * 8;

// ---------------------------------------------------------------------------
// Fused panel solve  X L^T = P  (in place), L = nbw x nbw lower (nbw <= NB), P = rows x nbw, rows contiguous or
// block-cyclic (TrsmRows).  ONE launch: a CTA owns 64 rows and sweeps the 64-wide column blocks cb = 0, 1, ...:
//   T = P[:, cb] - sum_{c < cb} X[:, c] L[cb, c]^T  (64^3 tile products, FP64 FMA from shared memory), then the
//   64 x 64 substitution against L[cb, cb].  CTAs never wait for each other.
// Used on the critical path of the sharded factorisations (csrc/dist.cu), where the recursive solve costs ~15
// dependent launches per column, each waiting for an SM of the concurrently running trailing update.
// Per entry the updates are applied one by one in column order (the rounding model of the recursive path).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1)
trsm_panel_kernel(double* __restrict__ P, long ldp, const TrsmRows rm, const double* __restrict__ L, long ldl, int nbw,
                  const __grid_constant__ PeerPush pp) {
  extern __shared__ double tsm[];
  double* sAt = tsm;                       // [TT][TPAD] X tile (k-major) / transposed diagonal factor
  double* sBt = tsm + TT * TPAD;           // [TT][TPAD] L[cb, c] tile (k-major)
  double* st = tsm + 2 * TT * TPAD;        // [TT][LDS_PAD] row-major staging of the tile being solved
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TT;          // logical first row
  const int nrow = min(TT, rm.rows - r0);
  const int ncb = (nbw + TT - 1) / TT;
  const int ty = tid >> 4, tx = tid & 15;
  // physical rows of this CTA's 64 logical rows (a 64-row tile never straddles a cyclic block: nb % 64 == 0)
  const long prow0 = trsm_phys_row(rm, r0);
  double* Pt = P + prow0 * ldp;
  for (int cb = 0; cb < ncb; ++cb) {
    const int c0 = cb * TT, wcb = min(TT, nbw - c0);
    double acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int gr = ty * 4 + r, gc = tx * 4 + c;
        acc[r][c] = (gr < nrow && gc < wcb) ? Pt[(long)gr * ldp + c0 + gc] : 0.0;
      }
    for (int cc = 0; cc < cb; ++cc) {
      const int k0 = cc * TT;              // all earlier column blocks are full (64 wide)
      for (int e = tid; e < TT * TT; e += 256) {
        const int r = e / TT, c = e % TT;
        sAt[c * TPAD + r] = (r < nrow) ? Pt[(long)r * ldp + k0 + c] : 0.0;                 // X[:, cc] (already solved)
        sBt[c * TPAD + r] = (r < wcb) ? L[(long)(c0 + r) * ldl + k0 + c] : 0.0;            // L[cb, cc]
      }
      __syncthreads();
#pragma unroll 8
      for (int kk = 0; kk < TT; ++kk) {
        const double2 a01 = *reinterpret_cast<const double2*>(sAt + kk * TPAD + ty * 4);
        const double2 a23 = *reinterpret_cast<const double2*>(sAt + kk * TPAD + ty * 4 + 2);
        const double2 b01 = *reinterpret_cast<const double2*>(sBt + kk * TPAD + tx * 4);
        const double2 b23 = *reinterpret_cast<const double2*>(sBt + kk * TPAD + tx * 4 + 2);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y};
        const double b[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fma(-a[r], b[c], acc[r][c]);
      }
      __syncthreads();
    }
    // diagonal factor L[cb, cb] transposed (identity padded) and the tile to solve
    for (int e = tid; e < TT * TT; e += 256) {
      const int r = e / TT, c = e % TT;    // L[r][c]
      double v = (r == c) ? 1.0 : 0.0;
      if (r < wcb && c < wcb) v = (c <= r) ? L[(long)(c0 + r) * ldl + c0 + c] : 0.0;
      sAt[c * TPAD + r] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) st[(ty * 4 + r) * LDS_PAD + tx * 4 + c] = acc[r][c];
    __syncthreads();
    tile_trsm_smem(st, sAt);
    __syncthreads();
    for (int e = tid; e < TT * TT; e += 256) {
      const int r = e / TT, c = e % TT;
      if (r < nrow && c < wcb) {
        const double v = st[r * LDS_PAD + c];
        const long off = (long)r * ldp + c0 + c;
        Pt[off] = v;
        for (int q = 0; q < pp.npeers; ++q) pp.base[q][prow0 * ldp + off] = v;     // the gather, fused: NVLink stores
      }
    }
    __syncthreads();      // the solved tile is read back (as X[:, cb]) by other threads of this CTA in the next sweep
  }
  if (pp.npeers > 0) {
    // all peer stores of this CTA before its arrival; the last CTA publishes the flags
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(pp.counter, 1u);
      if (prev == gridDim.x - 1) {
        atomicExch(pp.counter, 0u);
        __threadfence_system();
        for (int q = 0; q < pp.npeers; ++q)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pp.flag[q]), "l"(pp.seq) : "memory");
      }
    }
  }
}
constexpr int PANEL_SMEM = (2 * TT * TPAD + TT * LDS_PAD) * 8;

__global__ void fill_identity_kernel(double* __restrict__ A, long ld, int rows, int cols) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < (long)rows * cols) {
    int i = (int)(e / cols), j = (int)(e % cols);
    A[(long)i * ld + j] = (i == j) ? 1.0 : 0.0;
  }
}

// mirror the lower triangle of an n x n matrix into its upper triangle (32x32 smem transpose)
__global__ void symmetrize_kernel(double* __restrict__ A, long ld, int n) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;   // block row / col, bj <= bi processed
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    int r = bi * 32 + k, c = bj * 32 + tx;
    tile[k][tx] = (r < n && c < n) ? A[(long)r * ld + c] : 0.0;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    int r = bj * 32 + k, c = bi * 32 + tx;   // transposed position
    if (r < n && c < n && c > r) A[(long)r * ld + c] = tile[tx][k];
  }
}

// ---------------------------------------------------------------------------
// Vector triangular solves  L x = b  /  L^T x = b  (in place in x): ONE persistent kernel.
//
// 64-row blocks; CTA c owns blocks c, c + G, c + 2G, ... (G = resident CTAs, so every CTA is on the device and the
// spin-waits below cannot deadlock).  For its block rb a CTA streams the 64 x 64 blocks L[rb, jb], jb < rb (forward; for
// the transposed solve L[jb, rb], jb > rb) against the finished parts of x, each as soon as the global counter `done`
// says x[jb] is final; then it solves the diagonal block (one warp, 64 shuffle steps, reciprocal diagonal prepared off
// the critical path), publishes x[rb] and advances the counter (release / acquire at GPU scope).
// The L blocks are issued before the wait, so the only serial part per block is: counter round trip, one 64 x 64
// product, the cross-warp reduction and the diagonal solve.  Replaces two launches per 64 rows (jnp.linalg.solve(L, vec)
// in the reference: src/PDEs.py:86, :205).
// ---------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool TRANS>
__global__ void __launch_bounds__(256, 2)
trsv_persistent_kernel(const double* __restrict__ L, long ld, int n, double* __restrict__ x, int* __restrict__ done) {
  __shared__ double sD[BASE * LDS_PAD];     // diagonal block (lower part)
  __shared__ double sR[8][BASE];            // per-warp partial sums
  const int nblk = (n + BASE - 1) / BASE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int ready = 0;                            // blocks known to be final (in solve order)
  for (int it = blockIdx.x; it < nblk; it += gridDim.x) {
    const int rb = TRANS ? nblk - 1 - it : it;        // block solved at position `it` of the order
    const int r0 = rb * BASE;
    const int nb = min(BASE, n - r0);
    // diagonal block and reciprocals: independent of x, loaded first
    {
      const int c = threadIdx.x & 63;
      for (int r = threadIdx.x >> 6; r < BASE; r += 4) {
        double v = (r == c) ? 1.0 : 0.0;
        if (r < nb && c <= r) v = L[(long)(r0 + r) * ld + r0 + c];
        sD[r * LDS_PAD + c] = v;
      }
    }
    double acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0;
    if (!TRANS) {
      // rows 8 warp .. 8 warp + 7 of the block; lane -> columns 2 lane, 2 lane + 1 of every earlier block
      const double* rowp[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int r = min(r0 + 8 * warp + k, n - 1);
        rowp[k] = L + (long)r * ld + 2 * lane;
      }
      for (int jb = 0; jb < rb; ++jb) {
        double2 l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) l[k] = __ldcs(reinterpret_cast<const double2*>(rowp[k] + jb * BASE));
        if (jb >= ready) {
          do { ready = ld_acquire(done); } while (ready <= jb);
        }
        const double2 xv = __ldcg(reinterpret_cast<const double2*>(x + jb * BASE + 2 * lane));
#pragma unroll
        for (int k = 0; k < 8; ++k) { acc[k] = fma(l[k].x, xv.x, acc[k]); acc[k] = fma(l[k].y, xv.y, acc[k]); }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sR[0][8 * warp + k] = v;
      }
    } else {
      // transposed: rows 8 warp .. 8 warp + 7 of every later block jb; lane -> columns 2 lane, 2 lane + 1 of block rb
      for (int s = 0; s < it; ++s) {
        const int jb = nblk - 1 - s;
        const int nrow = min(BASE, n - jb * BASE);
        double2 l[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 8 * warp + k;
          l[k] = (r < nrow) ? __ldcs(reinterpret_cast<const double2*>(L + (long)(jb * BASE + r) * ld + r0 + 2 * lane)) : make_double2(0.0, 0.0);
        }
        if (s >= ready) {
          do { ready = ld_acquire(done); } while (ready <= s);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int r = 8 * warp + k;
          const double xr = (r < nrow) ? __ldcg(x + jb * BASE + r) : 0.0;
          acc[0] = fma(l[k].x, xr, acc[0]);
          acc[1] = fma(l[k].y, xr, acc[1]);
        }
      }
      sR[warp][2 * lane] = acc[0];
      sR[warp][2 * lane + 1] = acc[1];
    }
    __syncthreads();
    if (warp == 0) {
      // v = b - (sum of the products); two entries per lane: i = lane, lane + 32
      double v0, v1;
      if (!TRANS) {
        v0 = ((lane < nb) ? x[r0 + lane] : 0.0) - sR[0][lane];
        v1 = ((lane + 32 < nb) ? x[r0 + lane + 32] : 0.0) - sR[0][lane + 32];
      } else {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { s0 += sR[w][lane]; s1 += sR[w][lane + 32]; }
        v0 = ((lane < nb) ? x[r0 + lane] : 0.0) - s0;
        v1 = ((lane + 32 < nb) ? x[r0 + lane + 32] : 0.0) - s1;
      }
      const double inv0 = 1.0 / sD[lane * LDS_PAD + lane], inv1 = 1.0 / sD[(lane + 32) * LDS_PAD + lane + 32];
      if (!TRANS) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const double xj = __shfl_sync(0xffffffffu, v0 * inv0, j);
          if (lane == j) v0 = xj;
          if (lane > j) v0 = fma(-sD[lane * LDS_PAD + j], xj, v0);
          v1 = fma(-sD[(lane + 32) * LDS_PAD + j], xj, v1);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const double xj = __shfl_sync(0xffffffffu, v1 * inv1, j);
          if (lane == j) v1 = xj;
          if (lane > j) v1 = fma(-sD[(lane + 32) * LDS_PAD + 32 + j], xj, v1);
        }
      } else {
        // L^T x = v: backward over j; v_i -= L[j][i] x_j for i < j
#pragma unroll
        for (int j = 31; j >= 0; --j) {
          const double xj = __shfl_sync(0xffffffffu, v1 * inv1, j);
          if (lane == j) v1 = xj;
          if (lane < j) v1 = fma(-sD[(32 + j) * LDS_PAD + 32 + lane], xj, v1);
          v0 = fma(-sD[(32 + j) * LDS_PAD + lane], xj, v0);
        }
#pragma unroll
        for (int j = 31; j >= 0; --j) {
          const double xj = __shfl_sync(0xffffffffu, v0 * inv0, j);
          if (lane == j) v0 = xj;
          if (lane < j) v0 = fma(-sD[j * LDS_PAD + lane], xj, v0);
        }
      }
      if (lane < nb) x[r0 + lane] = v0;
      if (lane + 32 < nb) x[r0 + lane + 32] = v1;
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release(done, it + 1);
    }
    __syncthreads();    // sD / sR are reused by the next block of this CTA
  }
}

}  // namespace

int trsm_base_launch(gpp_handle* h, double* P, long ldp, const TrsmRows& rm, const double* L, long ldl, int nbl) {
  if (rm.rows <= 0 || nbl <= 0) return GPP_OK;
  static bool attr[64] = {false};        // function attributes are per device
  const int smem = (BASE * BASE + TRSM_ROWS * LDS_PAD) * 8;
  if (h->device >= 64 || !attr[h->device]) {
    CUDA_TRY(h, cudaFuncSetAttribute(trsm_base_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (h->device < 64) attr[h->device] = true;
  }
  trsm_base_kernel<<<(rm.rows + TRSM_ROWS - 1) / TRSM_ROWS, TRSM_ROWS, smem, h->cur>>>(P, ldp, rm, L, ldl, nbl);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

int trsm_panel_launch(gpp_handle* h, double* P, long ldp, const TrsmRows& rm, const double* L, long ldl, int nbw, const PeerPush* push) {
  if (rm.rows <= 0 || nbw <= 0) return GPP_OK;
  PeerPush pp{};
  if (push) pp = *push;
  if (rm.stride_blk != 0 && rm.nb % TT) { h->err = "trsm_panel: cyclic block size must be a multiple of 64"; return -1; }
  static bool attr[64] = {false};
  if (h->device >= 64 || !attr[h->device]) {
    CUDA_TRY(h, cudaFuncSetAttribute(trsm_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM));
    if (h->device < 64) attr[h->device] = true;
  }
  trsm_panel_kernel<<<(rm.rows + TT - 1) / TT, 256, PANEL_SMEM, h->cur>>>(P, ldp, rm, L, ldl, nbw, pp);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

int fill_identity_launch(gpp_handle* h, double* A, long ld, int rows, int cols) {
  const long tot = (long)rows * cols;
  if (tot <= 0) return GPP_OK;
  fill_identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->cur>>>(A, ld, rows, cols);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

int trsm_right_lt(gpp_handle* h, const Mat& P, int pr0, int pc0, int rows, const Mat& L, int lr0, int lc0, int nb) {
  if (rows <= 0 || nb <= 0) return GPP_OK;
  if (nb <= BASE) {
    TrsmRows rm{rows, 0, 0, 0};
    return trsm_base_launch(h, P.base + (long)pr0 * P.ld + pc0, P.ld, rm, L.base + (long)lr0 * L.ld + lc0, L.ld, nb);
  }
  const int hh = (int)round_up((nb + 1) / 2, BASE);
  int rc = trsm_right_lt(h, P, pr0, pc0, rows, L, lr0, lc0, hh);
  if (rc) return rc;
  GemmDesc d{};
  d.mapA = P.map; d.mapB = L.map; d.mapAdiag = nullptr; d.mapBdiag = nullptr;
  d.a_row0 = pr0; d.b_row0 = lr0 + hh;
  d.C = P.base + (long)pr0 * P.ld + pc0 + hh; d.ldc = P.ld;
  d.Cin = d.C; d.ldcin = P.ld;
  d.m = rows; d.n = nb - hh;
  d.k0 = pc0; d.k1 = pc0 + hh; d.kb_off = lc0 - pc0;
  d.ktri = 0; d.diag_nb = 0; d.alpha = -1.0; d.lower_only = 0;
  rc = gemm_nt_launch(h, d);
  if (rc) return rc;
  return trsm_right_lt(h, P, pr0, pc0 + hh, rows, L, lr0 + hh, lc0 + hh, nb - hh);
}

// whole-matrix tiled factorisation (one persistent launch) of the nb x nb block at (r0, c0)
static int potrf_tiled(gpp_handle* h, const Mat& A, int r0, int c0, int nb, int gidx0) {
  static int grid_cap[64] = {0};
  int cap = h->device < 64 ? grid_cap[h->device] : 0;
  if (!cap) {
    CUDA_TRY(h, cudaFuncSetAttribute(potrf_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILED_SMEM));
    int per_sm = 0, sms = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, potrf_tiled_kernel, 256, TILED_SMEM));
    CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    cap = (per_sm > 0 ? per_sm : 1) * sms;
    if (h->device < 64) grid_cap[h->device] = cap;
  }
  if (!h->d_bar) CUDA_TRY(h, cudaMalloc(&h->d_bar, 64 * sizeof(unsigned)));
  unsigned* bar = h->d_bar + (h->bar_calls++ & 63);
  CUDA_TRY(h, cudaMemsetAsync(bar, 0, sizeof(unsigned), h->cur));
  const int nt = (nb + TT - 1) / TT;
  const long tiles = (long)nt * (nt - 1) / 2;          // largest trailing update
  int grid = (int)(tiles < 1 ? 1 : (tiles < cap ? tiles : cap));
  if (h->tiled_grid_limit > 0 && grid > h->tiled_grid_limit) grid = h->tiled_grid_limit;
  potrf_tiled_kernel<<<grid, 256, TILED_SMEM, h->cur>>>(A.base + (long)r0 * A.ld + c0, A.ld, nb, gidx0, h->d_info, bar);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

int potrf_diag(gpp_handle* h, const Mat& A, int r0, int c0, int nb, int gidx0) {
  if (nb > BASE && h->tiled_potrf) return potrf_tiled(h, A, r0, c0, nb, gidx0);
  if (nb <= BASE) {
    potrf_base_kernel<<<1, BASE, 0, h->cur>>>(A.base + (long)r0 * A.ld + c0, A.ld, nb, gidx0, h->d_info);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return GPP_OK;
  }
  const int hh = (int)round_up((nb + 1) / 2, BASE);
  int rc = potrf_diag(h, A, r0, c0, hh, gidx0);
  if (rc) return rc;
  rc = trsm_right_lt(h, A, r0 + hh, c0, nb - hh, A, r0, c0, hh);
  if (rc) return rc;
  GemmDesc d{};
  d.mapA = A.map; d.mapB = A.map;
  d.a_row0 = r0 + hh; d.b_row0 = r0 + hh;
  d.C = A.base + (long)(r0 + hh) * A.ld + c0 + hh; d.ldc = A.ld; d.Cin = d.C; d.ldcin = A.ld;
  d.m = nb - hh; d.n = nb - hh; d.k0 = c0; d.k1 = c0 + hh; d.kb_off = 0;
  d.alpha = -1.0; d.lower_only = 1;
  rc = gemm_nt_launch(h, d);
  if (rc) return rc;
  return potrf_diag(h, A, r0 + hh, c0 + hh, nb - hh, gidx0 + hh);
}

// Left-looking blocked Cholesky: block column j first receives all earlier updates in long-K GEMMs
// (accumulators stay in registers), then its diagonal block is factorised and the rows below are solved
// against it.
//
// Look-ahead schedule (h->lookahead): the update of column j is split at its last block of K,
//   G1(j): K = [0, (j-1) NB)   needs columns <= j-2   -> streams sG[j & 1]
//   G2(j): K = [(j-1) NB, j NB) needs column j-1       -> high-priority stream sP, followed by Panel(j)
// so that G1(j+1) runs while G2(j) and the latency-bound panel chain of column j are in flight, and fills the
// SMs left idle by the last wave of G1(j).  K is still consumed in ascending order (progressive accumulation).
static cudaEvent_t next_event(gpp_handle* h, size_t& used) {
  if (used == h->evpool.size()) {
    cudaEvent_t e;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    h->evpool.push_back(e);
  }
  return h->evpool[used++];
}

int potrf_lower(gpp_handle* h, double* A, long ld, int n, const TMap2* map) {
  Mat M{A, ld, map};
  const int NB = h->NB;
  const int nblk = (n + NB - 1) / NB;
  CUDA_TRY(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), h->stream));
  if (h->tiled_potrf && n <= h->tiled_max_n) return potrf_tiled(h, M, 0, 0, n, 0);   // latency-bound sizes: one persistent kernel
  auto update = [&](int j0, int nbj, int k0, int k1) -> int {
    if (k1 <= k0) return GPP_OK;
    GemmDesc d{};
    d.mapA = map; d.mapB = map;
    d.a_row0 = j0; d.b_row0 = j0;
    d.C = A + (long)j0 * ld + j0; d.ldc = ld; d.Cin = d.C; d.ldcin = ld;
    d.m = n - j0; d.n = nbj; d.k0 = k0; d.k1 = k1; d.kb_off = 0;
    d.alpha = -1.0; d.lower_only = 0;
    return gemm_nt_launch(h, d);
  };
  auto panel = [&](int j0, int nbj) -> int {
    int rc = potrf_diag(h, M, j0, j0, nbj, j0);
    if (rc) return rc;
    return trsm_right_lt(h, M, j0 + nbj, j0, n - j0 - nbj, M, j0, j0, nbj);
  };
  const bool la = h->lookahead && nblk >= 4 && h->sP != nullptr;
  if (!la) {
    for (int j0 = 0; j0 < n; j0 += NB) {
      const int nbj = (n - j0 < NB) ? (n - j0) : NB;
      int rc = update(j0, nbj, 0, j0);
      if (rc) return rc;
      rc = panel(j0, nbj);
      if (rc) return rc;
    }
    return GPP_OK;
  }
  size_t used = 0;
  cudaEvent_t ev_start = next_event(h, used);
  CUDA_TRY(h, cudaEventRecord(ev_start, h->stream));
  for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaStreamWaitEvent(h->sG[k], ev_start, 0));
  CUDA_TRY(h, cudaStreamWaitEvent(h->sP, ev_start, 0));
  std::vector<cudaEvent_t> evP(nblk);
  int rc = GPP_OK;
  for (int j = 0; j < nblk && !rc; ++j) {
    const int j0 = j * NB;
    const int nbj = (n - j0 < NB) ? (n - j0) : NB;
    if (j >= 2) {
      cudaStream_t sg = h->sG[j & 1];
      CUDA_TRY(h, cudaStreamWaitEvent(sg, evP[j - 2], 0));
      h->cur = sg;
      rc = update(j0, nbj, 0, j0 - NB);
      if (rc) break;
      cudaEvent_t eg = next_event(h, used);
      CUDA_TRY(h, cudaEventRecord(eg, sg));
      CUDA_TRY(h, cudaStreamWaitEvent(h->sP, eg, 0));
    }
    h->cur = h->sP;
    if (j >= 1) rc = update(j0, nbj, j0 - NB, j0);
    if (!rc) rc = panel(j0, nbj);
    evP[j] = next_event(h, used);
    CUDA_TRY(h, cudaEventRecord(evP[j], h->sP));
  }
  h->cur = h->stream;
  if (rc) return rc;
  CUDA_TRY(h, cudaStreamWaitEvent(h->stream, evP[nblk - 1], 0));
  for (int k = 0; k < 2; ++k) {   // the G streams end before the last panel, but keep the main stream ordered after them
    cudaEvent_t e = next_event(h, used);
    CUDA_TRY(h, cudaEventRecord(e, h->sG[k]));
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, e, 0));
  }
  return GPP_OK;
}

// U = L^{-T} (block column by block column, into the strict upper triangle of the buffer,
// clean diagonal blocks in udiag), then Ainv = (U U^T)[0:Mint, 0:Mint] = (Theta^{-1}) interior block.
int inverse_interior(gpp_handle* h, GramSlot& s) {
  const int NB = h->NB, M = s.M;
  static const bool trace = getenv("GPP_TRACE") != nullptr;
  if (trace) cudaEventRecord(h->ev[2], h->stream);
  Mat T{s.T, s.ld, &s.mapT};
  Mat UD{s.udiag, (long)NB, &s.mapUdiag};
  const int nblk = (M + NB - 1) / NB;
  // block column i of U:  T_i = -U[0:o, 0:o] L[i, 0:o]^T  (o = i NB; rows are upper triangular: ktri), then T_i L_ii^{-T}
  auto col_gemm = [&](int i, int rows, int k0, int k1, bool accumulate) -> int {
    const int o = i * NB;
    const int nbi = (M - o < NB) ? (M - o) : NB;
    if (rows <= 0 || k1 <= k0) return GPP_OK;
    GemmDesc d{};
    d.mapA = &s.mapT; d.mapAdiag = &s.mapUdiag; d.mapB = &s.mapT; d.mapBdiag = nullptr;
    d.a_row0 = 0; d.b_row0 = o;
    d.C = s.T + o; d.ldc = s.ld;
    d.Cin = accumulate ? d.C : nullptr; d.ldcin = s.ld;
    d.m = rows; d.n = nbi; d.k0 = k0; d.k1 = k1; d.kb_off = 0;
    d.ktri = 1; d.diag_nb = NB; d.alpha = -1.0; d.lower_only = 0;
    return gemm_nt_launch(h, d);
  };
  auto diag_inverse = [&](int i) -> int {            // udiag block i = L_ii^{-T}  (X L_ii^T = I)
    const int o = i * NB;
    const int nbi = (M - o < NB) ? (M - o) : NB;
    const long tot = (long)NB * NB;
    fill_identity_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, h->cur>>>(s.udiag + (long)o * NB, NB, NB, NB);
    h->launches++;
    return trsm_right_lt(h, UD, o, 0, nbi, T, o, o, nbi);
  };
  const bool la = h->lookahead && nblk >= 4 && h->sP != nullptr;
  if (!la) {
    for (int i = 0; i < nblk; ++i) {
      const int o = i * NB;
      const int nbi = (M - o < NB) ? (M - o) : NB;
      int rc = diag_inverse(i);
      if (rc) return rc;
      rc = col_gemm(i, o, 0, o, false);
      if (rc) return rc;
      rc = trsm_right_lt(h, T, 0, o, o, T, o, o, nbi);
      if (rc) return rc;
    }
  } else {
    // look-ahead, same schedule as potrf_lower: G1(i) = K < (i-1) NB on alternating streams (needs columns <= i-2),
    // G2(i) = last K block + the solve against L_ii on the high-priority stream
    size_t used = 0;
    cudaEvent_t ev_start = next_event(h, used);
    CUDA_TRY(h, cudaEventRecord(ev_start, h->stream));
    for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaStreamWaitEvent(h->sG[k], ev_start, 0));
    CUDA_TRY(h, cudaStreamWaitEvent(h->sP, ev_start, 0));
    std::vector<cudaEvent_t> evP(nblk);
    int rc = GPP_OK;
    for (int i = 0; i < nblk && !rc; ++i) {
      const int o = i * NB;
      const int nbi = (M - o < NB) ? (M - o) : NB;
      bool have_g1 = false;
      if (i >= 2) {
        cudaStream_t sg = h->sG[i & 1];
        CUDA_TRY(h, cudaStreamWaitEvent(sg, evP[i - 2], 0));
        h->cur = sg;
        rc = col_gemm(i, o - NB, 0, o - NB, false);          // rows of blocks 0..i-2, K up to (i-1) NB
        if (rc) break;
        cudaEvent_t eg = next_event(h, used);
        CUDA_TRY(h, cudaEventRecord(eg, sg));
        CUDA_TRY(h, cudaStreamWaitEvent(h->sP, eg, 0));
        have_g1 = true;
      }
      h->cur = h->sP;
      rc = diag_inverse(i);
      if (rc) break;
      if (i >= 1) {
        // rows of block i-1 start here (no G1 part): clear them so that one accumulating GEMM covers all rows
        CUDA_TRY(h, cudaMemset2DAsync(s.T + (long)(o - NB) * s.ld + o, s.ld * 8, 0, (size_t)nbi * 8, NB, h->sP));
        if (have_g1) {
          rc = col_gemm(i, o, o - NB, o, true);
        } else {
          rc = col_gemm(i, o, 0, o, false);                  // i == 1: single GEMM
        }
        if (rc) break;
        rc = trsm_right_lt(h, T, 0, o, o, T, o, o, nbi);
        if (rc) break;
      }
      evP[i] = next_event(h, used);
      CUDA_TRY(h, cudaEventRecord(evP[i], h->sP));
    }
    h->cur = h->stream;
    if (rc) return rc;
    CUDA_TRY(h, cudaStreamWaitEvent(h->stream, evP[nblk - 1], 0));
    for (int k = 0; k < 2; ++k) {
      cudaEvent_t e = next_event(h, used);
      CUDA_TRY(h, cudaEventRecord(e, h->sG[k]));
      CUDA_TRY(h, cudaStreamWaitEvent(h->stream, e, 0));
    }
  }
  if (trace) cudaEventRecord(h->ev[3], h->stream);
  {
    GemmDesc d{};
    d.mapA = &s.mapT; d.mapAdiag = &s.mapUdiag; d.mapB = &s.mapT; d.mapBdiag = &s.mapUdiag;
    d.a_row0 = 0; d.b_row0 = 0;
    d.C = s.Ainv; d.ldc = s.ldA; d.Cin = nullptr;
    d.m = s.Mint; d.n = s.Mint; d.k0 = 0; d.k1 = (int)round_up(M, 16); d.kb_off = 0;
    d.ktri = 1; d.diag_nb = NB; d.alpha = 1.0; d.lower_only = 1;
    int rc = gemm_nt_launch(h, d);
    if (rc) return rc;
    if (trace) cudaEventRecord(h->ev[4], h->stream);
    dim3 grid((s.Mint + 31) / 32, (s.Mint + 31) / 32), blk(32, 8);
    symmetrize_kernel<<<grid, blk, 0, h->cur>>>(s.Ainv, s.ldA, s.Mint);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
  }
  if (trace) {
    cudaEventRecord(h->ev[5], h->stream);
    cudaEventSynchronize(h->ev[5]);
    float t[3];
    for (int k = 0; k < 3; ++k) cudaEventElapsedTime(&t[k], h->ev[2 + k], h->ev[3 + k]);
    fprintf(stderr, "[gpp trace] inverse: U = L^-T %.2f ms | U U^T %.2f | symmetrize %.2f\n", t[0], t[1], t[2]);
  }
  return GPP_OK;
}

int trsv_lower(gpp_handle* h, const double* L, long ld, int n, double* x, bool transposed) {
  if (n <= 0) return GPP_OK;
  if ((ld & 1) || (reinterpret_cast<uintptr_t>(L) & 15) || (reinterpret_cast<uintptr_t>(x) & 15)) { h->err = "trsv: operands must be 16-byte aligned"; return -1; }
  static int grid_cap[64] = {0};
  int cap = h->device < 64 ? grid_cap[h->device] : 0;
  if (!cap) {
    int per_sm = 0, sms = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trsv_persistent_kernel<false>, 256, 0));
    CUDA_TRY(h, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
    cap = (per_sm > 0 ? per_sm : 1) * sms;
    if (h->device < 64) grid_cap[h->device] = cap;
  }
  if (!h->d_trsv_flag) CUDA_TRY(h, cudaMalloc(&h->d_trsv_flag, 16 * sizeof(int)));
  int* flag = h->d_trsv_flag + (h->trsv_calls++ & 15);
  CUDA_TRY(h, cudaMemsetAsync(flag, 0, sizeof(int), h->cur));
  const int nblk = (n + BASE - 1) / BASE;
  const int grid = nblk < cap ? nblk : cap;
  if (transposed) trsv_persistent_kernel<true><<<grid, 256, 0, h->cur>>>(L, ld, n, x, flag);
  else trsv_persistent_kernel<false><<<grid, 256, 0, h->cur>>>(L, ld, n, x, flag);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}
