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
  __shared__ double sInv[BASE];             // reciprocal diagonal
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

int potrf_diag(gpp_handle* h, const Mat& A, int r0, int c0, int nb, int gidx0) {
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
