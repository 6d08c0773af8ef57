// FP64 "NT" GEMM / SYRK on sm_100a:  C = Cin + alpha * A * B^T  with A (m x K) and
// B (n x K) both row-major (K contiguous).  This one kernel carries every O(M^3) part
// of the hot path (reference: jnp.linalg.cholesky src/PDEs.py:77, the L^{-1}J solves
// and J^T Theta^{-1} J products of src/PDEs.py:97-102): left-looking Cholesky updates,
// the triangular-inverse updates and the U U^T product.
//
// Design (B200): tcgen05 has no FP64 kind, so FP64 tensor work is mma.sync m8n8k4
// (SASS DMMA.8x8x4, measured pipe peak 37.0 TFLOP/s, profiles/r01_fp64_pipe_microbench.txt)
// with register accumulators.  Operand tiles (TILE rows x 16 k) are brought in by TMA
// (cp.async.bulk.tensor.2d, 128B swizzle) into a 6-stage mbarrier ring by one
// producer warp; the consumer warps each own a 32x32 block of the CTA tile
// (128x128: 16 warps, 1 CTA/SM; 64x64: 4 warps, 2 CTAs/SM).
// The k-slot -> column mapping inside a 16-wide chunk is permuted (same for A and B)
// so that every fragment load is a conflict-free LDS.64 under the 128B swizzle.
// ncu (profiles/r01_ncu_summary.md): DMMA pipe busy 96 % of the active cycles on the
// bench-size update; DRAM traffic within 8 % of the algorithmic operand bytes.
#include "gpp_internal.cuh"

namespace {

constexpr int BK = 16, STAGES = 6;
// TMA box = one operand tile: TILE rows x 16 doubles (each buffer carries a descriptor per tile size, TMap2)
// Two tile sizes share one code path: 128 x 128 (16 consumer warps, 1 CTA / SM) for the bulk, 64 x 64 (4 consumer warps,
// 2 CTAs / SM) for launches with too few 128-tiles to fill the GPU (panel chain, small matrices).  Per output element
// both consume K in the same order, so results do not depend on the tile size.
template <int TILE> struct Cfg {
  static constexpr int WG = TILE / 32;                       // warp grid is WG x WG, each warp 32 x 32
  static constexpr int CONSUMER_WARPS = WG * WG;
  static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
  static constexpr int TILE_BYTES = TILE * BK * 8;           // one operand tile
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;         // A + B
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int MIN_CTAS = TILE == 128 ? 1 : 2;
};

struct GemmParams {
  CUtensorMap mapA, mapB, mapAdiag, mapBdiag;
  int has_adiag, has_bdiag;
  int a_row0, b_row0;
  double* C; long ldc;
  const double* Cin; long ldcin;
  int m, n, k0, k1, kb_off, ktri, diag_nb;
  double alpha;
  int lower_only;
  int tiles_m, tiles_n;
  // block-diagonal batch mode (multi-GPU Cholesky): tile t -> local block row bd_lblk0 + t / ntri, lower tile t % ntri of
  // the NB x NB diagonal block of that block row; global column of the block = ((lblk * bd_world) + bd_rank) * bd_nb
  int bd_mode, bd_world, bd_rank, bd_lblk0, bd_nb, bd_M;
  // task-list mode (multi-GPU path): blockIdx.x / (task_tps^2) indexes tasks[]; see GemmTask in gpp_internal.cuh
  const GemmTask* tasks;
  int task_tps;
  int blocksum;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}

template <int TILE>
__global__ void __launch_bounds__(Cfg<TILE>::THREADS, Cfg<TILE>::MIN_CTAS)
gemm_nt_dmma_kernel(const __grid_constant__ GemmParams p) {
  constexpr int BM = TILE, BN = TILE;
  constexpr int CONSUMER_WARPS = Cfg<TILE>::CONSUMER_WARPS, WG = Cfg<TILE>::WG;
  constexpr int TILE_BYTES = Cfg<TILE>::TILE_BYTES, STAGE_BYTES = Cfg<TILE>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  // 1024B alignment for the 128B swizzle atom
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  // tile index -> (tm, tn); lower_only enumerates the lower-triangular tile set row by row
  int tm, tn;
  if (p.lower_only) {
    // tiles (i, j) with j <= i + shift where shift accounts for a_row0/b_row0 offsets being equal
    long t = blockIdx.x;
    int i = static_cast<int>((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
    while ((long)(i + 1) * (i + 2) / 2 <= t) ++i;
    while ((long)i * (i + 1) / 2 > t) --i;
    tm = i;
    tn = static_cast<int>(t - (long)i * (i + 1) / 2);
  } else {
    tm = blockIdx.x / p.tiles_n;
    tn = blockIdx.x % p.tiles_n;
  }
  int row_base = tm * BM;    // within the output block
  int col_base = tn * BN;
  int a_row = p.a_row0 + row_base;
  int b_row = p.b_row0 + col_base;
  double* Cp = p.C;
  const double* Cinp = p.Cin;
  int m_lim = p.m, n_lim = p.n;
  if (p.bd_mode) {
    const int nt = p.bd_nb / BM, ntri = nt * (nt + 1) / 2;
    const int lblk = p.bd_lblk0 + blockIdx.x / ntri;
    int tri = blockIdx.x % ntri, ti = 0;
    while ((ti + 1) * (ti + 2) / 2 <= tri) ++ti;
    const int tj = tri - ti * (ti + 1) / 2;
    const int gcol0 = (lblk * p.bd_world + p.bd_rank) * p.bd_nb;      // global index of the block's first row / column
    row_base = ti * BM; col_base = tj * BN;
    a_row = lblk * p.bd_nb + row_base;                                // local rows of both operands
    b_row = lblk * p.bd_nb + col_base;
    Cp = p.C + (long)(lblk * p.bd_nb) * p.ldc + gcol0;
    Cinp = Cp;
    m_lim = n_lim = (p.bd_M - gcol0 < p.bd_nb) ? (p.bd_M - gcol0) : p.bd_nb;
    if (row_base >= m_lim || col_base >= n_lim) return;
  }

  int k_lo = p.k0, k_hi = p.k1, kb_off = p.kb_off;
  if (p.tasks) {
    const int per = p.task_tps * p.task_tps;
    const int4* tp = reinterpret_cast<const int4*>(p.tasks + blockIdx.x / per);
    const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);   // {a_row, b_row, c_row, c_col} {m, n, k0, k1} {kb_off, tri, -, -}
    const int tile = blockIdx.x % per;
    const int ti = tile / p.task_tps, tj = tile % p.task_tps;
    row_base = ti * BM; col_base = tj * BN;
    if (row_base >= t1.x || col_base >= t1.y || (t2.y && tj > ti)) return;     // uniform over the CTA, before any barrier
    a_row = t0.x + row_base; b_row = t0.y + col_base;
    Cp = p.C + (long)t0.z * p.ldc + t0.w;
    Cinp = p.Cin ? p.Cin + (long)t0.z * p.ldcin + t0.w : nullptr;
    m_lim = t1.x; n_lim = t1.y;
    k_lo = t1.z; k_hi = t1.w; kb_off = t2.x;
  }
  if (p.ktri) {
    int kb = (a_row / p.diag_nb) * p.diag_nb;
    if (kb > k_lo) k_lo = kb;
  }
  const int nchunks = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == CONSUMER_WARPS) {
    // ===== TMA producer =====
    if (lane == 0) {
      const int a_dblk = a_row / p.diag_nb, b_dblk = b_row / p.diag_nb;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % STAGES;
        const uint32_t ph = (c / STAGES) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem + s * STAGE_BYTES;
        uint8_t* sb = sa + TILE_BYTES;
        mbar_expect_tx(&full[s], STAGE_BYTES);
        const int k = k_lo + c * BK;
        const int kblk = k / p.diag_nb;
        if (p.has_adiag && kblk == a_dblk) tma_load_2d(sa, &p.mapAdiag, &full[s], k - kblk * p.diag_nb, a_row);
        else tma_load_2d(sa, &p.mapA, &full[s], k, a_row);
        if (p.has_bdiag && kblk == b_dblk) tma_load_2d(sb, &p.mapBdiag, &full[s], k - kblk * p.diag_nb, b_row);
        else tma_load_2d(sb, &p.mapB, &full[s], k + kb_off, b_row);
      }
    }
    return;
  }

  // ===== consumers: WG x WG warps, each 32 x 32 =====
  const int wm = warp / WG, wn = warp % WG;
  const int g = lane >> 2, t = lane & 3;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // Progressive update (alpha = +-1 with an addend): start the accumulators at alpha*Cin so that the running
  // value is the remainder Cin - sum_k a b (same order as a right-looking update).  For the nearly singular
  // Gram matrices of this solver the remainder shrinks quickly with k, and rounding each partial result
  // relative to the remainder -- not to the partial sum -- is what keeps pivots of size ~nugget positive.
  const bool progressive = (Cinp != nullptr) && (p.alpha == 1.0 || p.alpha == -1.0) && !p.blocksum;
  if (progressive) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row_base + wm * 32 + i * 8 + g;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int cc = col_base + wn * 32 + j * 8 + 2 * t;
        if (r < m_lim && cc < n_lim) {
          const double* src = Cinp + (long)r * p.ldcin + cc;
          acc[i][j][0] = p.alpha * src[0];
          if (cc + 1 < n_lim) acc[i][j][1] = p.alpha * src[1];
        }
      }
    }
  }

  // per-thread swizzled offsets: step s reads logical 16B chunk (s + 4*(t>>1)), element t&1
  uint32_t koff[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) koff[s] = ((((s + 4 * (t >> 1)) ^ g) << 4) + ((t & 1) << 3));
  const uint32_t a_thr = (wm * 32 + g) * 128;
  const uint32_t b_thr = (wn * 32 + g) * 128;
  const uint32_t smem_base = smem_u32(smem);

  for (int c = 0; c < nchunks; ++c) {
    const int s = c % STAGES;
    const uint32_t ph = (c / STAGES) & 1;
    mbar_wait(&full[s], ph);
    const uint32_t sa = smem_base + s * STAGE_BYTES + a_thr;
    const uint32_t sb = smem_base + s * STAGE_BYTES + TILE_BYTES + b_thr;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = lds64(sa + i * 1024 + koff[ks]);
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = lds64(sb + j * 1024 + koff[ks]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ===== epilogue: C = Cin + alpha * acc =====
  if (Cinp && !progressive) {
    // block summation: the addend joins here.  Full, 16-byte aligned tiles: the four 128-bit addend loads of a row are
    // issued together, then the four stores (ncu source page of the element-by-element form: 13 % of the samples sat
    // on 16 serialised load -> add -> store chains).  A thread reads an element before it overwrites that same element,
    // so the two pointers may be declared non-aliasing.
    const int rb = row_base + wm * 32 + g, cb = col_base + wn * 32 + 2 * t;
    double* __restrict__ Cw = Cp;
    const double* __restrict__ Cr = Cinp;
    if ((rb + 24 < m_lim) && (cb + 25 < n_lim) && ((reinterpret_cast<uintptr_t>(Cw) & 15) == 0) && !(p.ldc & 1) &&
        ((reinterpret_cast<uintptr_t>(Cr) & 15) == 0) && !(p.ldcin & 1)) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long ro = (long)(rb + i * 8);
        double2 cin[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) cin[j] = *reinterpret_cast<const double2*>(Cr + ro * p.ldcin + cb + j * 8);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<double2*>(Cw + ro * p.ldc + cb + j * 8) =
              make_double2(p.alpha * acc[i][j][0] + cin[j].x, p.alpha * acc[i][j][1] + cin[j].y);
      }
      return;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = row_base + wm * 32 + i * 8 + g;
    if (r >= m_lim) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cc = col_base + wn * 32 + j * 8 + 2 * t;
      if (cc >= n_lim) continue;
      double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
      double* dst = Cp + (long)r * p.ldc + cc;
      const bool two = (cc + 1 < n_lim);
      if (Cinp && !progressive) {
        const double* src = Cinp + (long)r * p.ldcin + cc;
        v0 += src[0];
        if (two) v1 += src[1];
      }
      if (two && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
      } else {
        dst[0] = v0;
        if (two) dst[1] = v1;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent task-list variant (right-looking schedules: K = NB per task, thousands of tiles per launch).  ncu on the
// one-CTA-per-tile launch (profiles/r02_ncu_task_gemm_bulk_raw.csv): DMMA pipe active 83 % of the SM-active cycles against
// 95 % for the long-K launches -- ~13 us of launch, task fetch, barrier set-up, pipeline fill and epilogue per 65 us tile.
// Here a CTA stays on its SM and walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...; the mbarrier ring keeps running
// across tiles, so the producer warp is already fetching the next tile's operands while the consumer warps store the
// current one.  Same k order per output element as the one-shot kernel.
// ---------------------------------------------------------------------------------------------------------------------
template <int TILE>
__global__ void __launch_bounds__(Cfg<TILE>::THREADS, Cfg<TILE>::MIN_CTAS)
gemm_tasks_persistent_kernel(const __grid_constant__ GemmParams p, int total_tiles) {
  constexpr int BM = TILE, BN = TILE;
  constexpr int CONSUMER_WARPS = Cfg<TILE>::CONSUMER_WARPS, WG = Cfg<TILE>::WG;
  constexpr int TILE_BYTES = Cfg<TILE>::TILE_BYTES, STAGE_BYTES = Cfg<TILE>::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONSUMER_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int per = p.task_tps * p.task_tps;
  uint32_t cbase = 0;                  // chunks consumed by this CTA before the current tile (same count in every warp)

  const int wm = warp / WG, wn = warp % WG;
  const int g = lane >> 2, t = lane & 3;
  uint32_t koff[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) koff[s] = ((((s + 4 * (t >> 1)) ^ g) << 4) + ((t & 1) << 3));
  const uint32_t a_thr = (wm * 32 + g) * 128;
  const uint32_t b_thr = (wn * 32 + g) * 128;
  const uint32_t smem_base = smem_u32(smem);

  for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
    const int4* tp = reinterpret_cast<const int4*>(p.tasks + tile_id / per);
    const int4 t0 = __ldg(tp), t1 = __ldg(tp + 1), t2 = __ldg(tp + 2);   // {a_row, b_row, c_row, c_col} {m, n, k0, k1} {kb_off, tri, -, -}
    const int tile = tile_id % per;
    const int ti = tile / p.task_tps, tj = tile % p.task_tps;
    const int row_base = ti * BM, col_base = tj * BN;
    if (row_base >= t1.x || col_base >= t1.y || (t2.y && tj > ti)) continue;   // same decision in every warp: no chunks
    const int a_row = t0.x + row_base, b_row = t0.y + col_base;
    const int k_lo = t1.z, k_hi = t1.w, kb_off = t2.x;
    const int nchunks = (k_hi > k_lo) ? (k_hi - k_lo + BK - 1) / BK : 0;

    if (warp == CONSUMER_WARPS) {
      // ===== TMA producer: runs ahead of the consumers across tile boundaries =====
      if (lane == 0) {
        for (int c = 0; c < nchunks; ++c) {
          const uint32_t gc = cbase + c;
          const int s = gc % STAGES;
          const uint32_t ph = (gc / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + TILE_BYTES;
          mbar_expect_tx(&full[s], STAGE_BYTES);
          const int k = k_lo + c * BK;
          tma_load_2d(sa, &p.mapA, &full[s], k, a_row);
          tma_load_2d(sb, &p.mapB, &full[s], k + kb_off, b_row);
        }
      }
      cbase += nchunks;
      continue;
    }

    // ===== consumers (block summation: the products of the tile are summed from zero, the addend joins in the epilogue) =====
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const uint32_t gc = cbase + c;
      const int s = gc % STAGES;
      const uint32_t ph = (gc / STAGES) & 1;
      mbar_wait(&full[s], ph);
      const uint32_t sa = smem_base + s * STAGE_BYTES + a_thr;
      const uint32_t sb = smem_base + s * STAGE_BYTES + TILE_BYTES + b_thr;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = lds64(sa + i * 1024 + koff[ks]);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = lds64(sb + j * 1024 + koff[ks]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
    }
    cbase += nchunks;
    // ===== epilogue: C = Cin + alpha * acc.  The task is decoded again here (cache hit) so that nothing but the
    // accumulators and the ring state lives across the main loop (96 registers per thread).
    {
      const int4 e0 = __ldg(tp), e1 = __ldg(tp + 1);
      // the addend is read before the same thread overwrites the same element, so for the compiler's purposes the two
      // pointers do not alias: with __restrict__ it may issue the loads of a row ahead of the stores of the previous one
      // (ncu source page before this change: 13 % of the samples sat on the 16 serialised load -> add -> store chains)
      double* __restrict__ Cp = p.C + (long)e0.z * p.ldc + e0.w;
      const double* __restrict__ Cinp = p.Cin ? p.Cin + (long)e0.z * p.ldcin + e0.w : nullptr;
      const int rb = (tile / p.task_tps) * BM + wm * 32 + g, cb = (tile % p.task_tps) * BN + wn * 32 + 2 * t;
      const bool interior = (rb + 24 < e1.x) && (cb + 25 < e1.y) && ((reinterpret_cast<uintptr_t>(Cp) & 15) == 0) && !(p.ldc & 1) &&
                            Cinp && ((reinterpret_cast<uintptr_t>(Cinp) & 15) == 0) && !(p.ldcin & 1);
      if (interior) {
        // full tile, 16-byte aligned: four 128-bit addend loads per row in flight, then four 128-bit stores
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const long ro = (long)(rb + i * 8);
          double2 cin[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) cin[j] = *reinterpret_cast<const double2*>(Cinp + ro * p.ldcin + cb + j * 8);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2*>(Cp + ro * p.ldc + cb + j * 8) =
                make_double2(p.alpha * acc[i][j][0] + cin[j].x, p.alpha * acc[i][j][1] + cin[j].y);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = rb + i * 8;
          if (r >= e1.x) continue;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int cc = cb + j * 8;
            if (cc >= e1.y) continue;
            double v0 = p.alpha * acc[i][j][0], v1 = p.alpha * acc[i][j][1];
            double* dst = Cp + (long)r * p.ldc + cc;
            const bool two = (cc + 1 < e1.y);
            if (Cinp) {
              const double* src = Cinp + (long)r * p.ldcin + cc;
              v0 += src[0];
              if (two) v1 += src[1];
            }
            if (two && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
              *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
            } else {
              dst[0] = v0;
              if (two) dst[1] = v1;
            }
          }
        }
      }
    }
  }
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled g_encode = nullptr;

}  // namespace

int make_tensor_map(gpp_handle* h, TMap2* map, const double* base, long rows, long cols, long ld) {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { h->err = "cuTensorMapEncodeTiled not available"; return GPP_CUDA_ERR; }
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 8) & 15)) { h->err = "tensor map: base/stride not 16B aligned"; return -1; }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 8};
  cuuint32_t estr[2] = {1, 1};
  for (int v = 0; v < 2; ++v) {
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)(v == 0 ? 128 : 64)};
    CUresult r = g_encode(v == 0 ? &map->m128 : &map->m64, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), gdim, gstr,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { h->err = "cuTensorMapEncodeTiled failed: " + std::to_string((int)r); return GPP_CUDA_ERR; }
  }
  return GPP_OK;
}

namespace {
template <int TILE>
int launch_tile(gpp_handle* h, const GemmDesc& d, GemmParams& p) {
  auto pick = [](const TMap2* m) -> const CUtensorMap& { return TILE == 128 ? m->m128 : m->m64; };
  p.mapA = pick(d.mapA);
  p.mapB = pick(d.mapB);
  p.mapAdiag = pick(d.mapAdiag ? d.mapAdiag : d.mapA);
  p.mapBdiag = pick(d.mapBdiag ? d.mapBdiag : d.mapB);
  static bool attr_set[64] = {false};      // function attributes are per device
  if (h->device >= 64 || !attr_set[h->device]) {
    CUDA_TRY(h, cudaFuncSetAttribute(gemm_nt_dmma_kernel<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<TILE>::SMEM_BYTES));
    if (h->device < 64) attr_set[h->device] = true;
  }
  p.tiles_m = (d.m + TILE - 1) / TILE;
  p.tiles_n = (d.n + TILE - 1) / TILE;
  long ntiles;
  if (p.bd_mode) {
    const int nt = d.bd_nb / TILE;
    ntiles = (long)d.bd_count * (nt * (nt + 1) / 2);
  } else if (d.lower_only) {
    ntiles = (long)p.tiles_m * (p.tiles_m + 1) / 2;     // square lower-triangular tile set (a_row0 == b_row0, m == n)
  } else {
    ntiles = (long)p.tiles_m * p.tiles_n;
  }
  gemm_nt_dmma_kernel<TILE><<<(unsigned)ntiles, Cfg<TILE>::THREADS, Cfg<TILE>::SMEM_BYTES, h->cur>>>(p);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}
}  // namespace

int gemm_nt_launch(gpp_handle* h, const GemmDesc& d) {
  if (d.bd_count <= 0 && (d.m <= 0 || d.n <= 0)) return GPP_OK;
  GemmParams p;
  p.has_adiag = d.mapAdiag != nullptr;
  p.has_bdiag = d.mapBdiag != nullptr;
  p.a_row0 = d.a_row0; p.b_row0 = d.b_row0;
  p.C = d.C; p.ldc = d.ldc; p.Cin = d.Cin; p.ldcin = d.ldcin;
  p.m = d.m; p.n = d.n; p.k0 = d.k0; p.k1 = d.k1; p.kb_off = d.kb_off; p.ktri = d.ktri;
  p.diag_nb = d.diag_nb > 0 ? d.diag_nb : (1 << 30);
  p.alpha = d.alpha; p.lower_only = d.lower_only;
  p.bd_mode = d.bd_count > 0; p.bd_world = d.bd_world; p.bd_rank = d.bd_rank; p.bd_lblk0 = d.bd_lblk0; p.bd_nb = d.bd_nb; p.bd_M = d.bd_M;
  p.tasks = nullptr; p.task_tps = 0; p.blocksum = 0;
  // tile choice: 64 x 64 when the 128-tiling would occupy less than half of the SMs
  long t128;
  if (p.bd_mode) { const int nt = d.bd_nb / 128; t128 = (long)d.bd_count * (nt * (nt + 1) / 2); }
  else { const long tm = (d.m + 127) / 128, tn = (d.n + 127) / 128; t128 = d.lower_only ? tm * (tm + 1) / 2 : tm * tn; }
  const int force = h->force_tile;
  const bool small = force ? (force == 64) : (t128 < 74);
  return small ? launch_tile<64>(h, d, p) : launch_tile<128>(h, d, p);
}

namespace {
template <int TILE>
int launch_tasks(gpp_handle* h, const GemmTaskDesc& d, GemmParams& p) {
  p.mapA = TILE == 128 ? d.mapA->m128 : d.mapA->m64;
  p.mapB = TILE == 128 ? d.mapB->m128 : d.mapB->m64;
  p.mapAdiag = p.mapA; p.mapBdiag = p.mapB;
  static bool attr_set[64] = {false};
  if (h->device >= 64 || !attr_set[h->device]) {
    CUDA_TRY(h, cudaFuncSetAttribute(gemm_nt_dmma_kernel<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<TILE>::SMEM_BYTES));
    if (h->device < 64) attr_set[h->device] = true;
  }
  p.task_tps = d.bs / TILE;
  const long ntiles = (long)d.ntasks * p.task_tps * p.task_tps;
  if (h->persistent_gemm && (p.blocksum || p.Cin == nullptr)) {
    static bool attr_p[64] = {false};
    static int sms[64] = {0};
    if (h->device >= 64 || !attr_p[h->device]) {
      CUDA_TRY(h, cudaFuncSetAttribute(gemm_tasks_persistent_kernel<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<TILE>::SMEM_BYTES));
      int n = 0;
      CUDA_TRY(h, cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, h->device));
      if (h->device < 64) { attr_p[h->device] = true; sms[h->device] = n; }
    }
    int nsm = h->device < 64 ? sms[h->device] : 148;
    if (nsm <= 0) nsm = 148;
    const long cap = (long)nsm * Cfg<TILE>::MIN_CTAS;
    const unsigned grid = (unsigned)(ntiles < cap ? ntiles : cap);
    gemm_tasks_persistent_kernel<TILE><<<grid, Cfg<TILE>::THREADS, Cfg<TILE>::SMEM_BYTES, h->cur>>>(p, (int)ntiles);
  } else {
    gemm_nt_dmma_kernel<TILE><<<(unsigned)ntiles, Cfg<TILE>::THREADS, Cfg<TILE>::SMEM_BYTES, h->cur>>>(p);
  }
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}
}  // namespace

int gemm_tasks_launch(gpp_handle* h, const GemmTaskDesc& d) {
  if (d.ntasks <= 0) return GPP_OK;
  if (d.bs <= 0 || d.bs % 128) { h->err = "task block size must be a multiple of 128"; return -1; }
  GemmParams p{};
  p.has_adiag = p.has_bdiag = 0;
  p.C = d.C; p.ldc = d.ldc; p.Cin = d.Cin; p.ldcin = d.ldcin;
  p.diag_nb = 1 << 30;
  p.alpha = d.alpha;
  p.tasks = d.tasks;
  p.blocksum = d.blocksum;
  p.tiles_m = p.tiles_n = 1;     // unused in task mode (kept non-zero: the generic tile decode still runs)
  const long t128 = (long)d.ntasks * (d.bs / 128) * (d.bs / 128);
  const int force = h->force_tile;
  const bool small = force ? (force == 64) : (t128 < 74);
  return small ? launch_tasks<64>(h, d, p) : launch_tasks<128>(h, d, p);
}
