// Gram assembly (lower triangle of Theta) and the prediction path.
//
// Replaces Gram_matrix_assembly / construct_Theta_test (src/Gram_matrice.py:11-187,
// :190-289) and the 19 autodiff functionals of src/kernels.py:8-179.  Every entry
//   Theta[(p,i),(q,j)] = L_p^x L_q^y kappa(x_i, x_j)
// is (polynomial in u1 = b1 d1, u2 = b2 d2) * kappa, kappa = exp(-(b1 d1^2 + b2 d2^2)/2):
//   d^m/dx^m d^n/dy^n along one axis = (-1)^m h_{m+n}(u) kappa,
//   h1 = u, h2 = u^2 - b, h3 = u (u^2 - 3b), h4 = u^4 - 6 b u^2 + 3 b^2.
// One exp per point pair is shared by all blocks the pair feeds; only the lower triangle
// of Theta is written (off-diagonal blocks p > q in full, diagonal blocks for i >= j).
//
// This file is compiled with -fmad=false: the polynomial prefactors use the same
// multiply/add sequence as the CPU oracle (oracle/gp_oracle.py::_h, functional), so
// Theta differs from it only through exp() (<= 1 ulp on both sides).
#include "gpp_internal.cuh"

namespace {

struct KParams {
  double b1, b2;        // polynomial scales
  double c3b1, c6b1, c3bb1, c3b2, c6b2, c3bb2;  // 3b, 6b, (3b)b per axis (host-evaluated like the oracle)
  double e1, e2;        // exponent: Gaussian: e1 = -(1/(2 sigma^2)); anisotropic: e1 = s_t, e2 = s_x
};

struct AsmParams {
  KParams k;
  const double* X;      // (N+Nb) x 2
  int N, Nb;
  double* T; long ld;
  int off[GPP_MAX_BLOCKS];
  int size[GPP_MAX_BLOCKS];
};

template <int KERNEL>
__device__ __forceinline__ double kappa_of(const KParams& k, double d1, double d2) {
  if (KERNEL == 0) {
    // exp(-(1/(2 sigma^2)) * (d1^2 + d2^2))            src/kernels.py:13
    return exp(k.e1 * (d1 * d1 + d2 * d2));
  } else {
    // exp(-((d1/s_t)^2 + (d2/s_x)^2))                   src/kernels.py:96-99
    double q1 = d1 / k.e1, q2 = d2 / k.e2;
    return exp(-(q1 * q1 + q2 * q2));
  }
}

// h[0..4] for one axis
__device__ __forceinline__ void hermite(double u, double b, double c3b, double c6b, double c3bb, double* h) {
  const double u2 = u * u;
  h[0] = 1.0;
  h[1] = u;
  h[2] = u2 - b;
  h[3] = u * (u2 - c3b);
  h[4] = u2 * u2 - c6b * u2 + c3bb;
}

// monomials d1^m1 d2^m2 of each row operator (OP_LAP has two)
__host__ __device__ constexpr int mono_n(int op) { return op == OP_LAP ? 2 : 1; }
__host__ __device__ constexpr int mono_m1(int op, int a) { return op == OP_D1 ? 1 : (op == OP_LAP ? (a == 0 ? 2 : 0) : 0); }
__host__ __device__ constexpr int mono_m2(int op, int a) {
  return op == OP_D2 ? 1 : (op == OP_D22 ? 2 : (op == OP_LAP ? (a == 0 ? 0 : 2) : 0));
}

// polynomial prefactor of L_x(OPX) L_y(OPY) kappa, summed in the oracle's order
template <int OPX, int OPY>
__device__ __forceinline__ double prefactor(const double* h1, const double* h2) {
  double poly = 0.0;
#pragma unroll
  for (int a = 0; a < mono_n(OPX); ++a) {
#pragma unroll
    for (int b = 0; b < mono_n(OPY); ++b) {
      const int m1 = mono_m1(OPX, a), m2 = mono_m2(OPX, a);
      const int n1 = mono_m1(OPY, b), n2 = mono_m2(OPY, b);
      const double sg = ((m1 + m2) & 1) ? -1.0 : 1.0;
      poly = poly + (sg * h1[m1 + n1]) * h2[m2 + n2];
    }
  }
  return poly;
}

__host__ __device__ constexpr int lay_nblk(int lay) { return lay == LAY_ELLIPTIC ? 2 : (lay == LAY_DARCY_A ? 3 : 4); }
__host__ __device__ constexpr int lay_op(int lay, int p) {
  return lay == LAY_ELLIPTIC ? (p == 0 ? OP_LAP : OP_ID)
       : lay == LAY_BURGERS  ? (p == 0 ? OP_D1 : p == 1 ? OP_D2 : p == 2 ? OP_D22 : OP_ID)
       : lay == LAY_EIKONAL  ? (p == 0 ? OP_D1 : p == 1 ? OP_D2 : p == 2 ? OP_LAP : OP_ID)
       :                       (p == 0 ? OP_D1 : p == 1 ? OP_D2 : OP_ID);
}

// ---------------------------------------------------------------------------------------------------
// Assembly kernel.  One CTA per lower tile pair (I >= J) of 64 x 64 collocation points, 8 warps.
//   phase A (row-wise): lane -> columns j = j0 + 2*lane, +1; warp w -> rows i = i0 + w + 8*rr.  One exp per
//     point pair; every block (p >= q) the pair feeds is written from registers with 128-bit stores
//     (a warp writes 512 contiguous bytes of one row of Theta).  u1, u2, kappa are stashed in shared memory.
//   phase B (only I > J): the mirrored entries Theta[(p, j), (q, i)], p > q, of the same unordered pairs are
//     produced from the stash with d -> -d (no second exp) and written transposed, again 128-bit coalesced.
// Diagonal tiles (I == J) visit all ordered pairs in phase A and need no phase B.
// ---------------------------------------------------------------------------------------------------
constexpr int TP = 64;            // points per tile side
constexpr int SPAD = TP + 1;      // stash row stride (doubles)
constexpr int ASM_SMEM = (3 * TP * SPAD + 4 * TP) * 8;

// Theta is written once and is far larger than L2: streaming (evict-first) stores
__device__ __forceinline__ void store_pair(double* dst, double e0, double e1, bool ok0, bool ok1, bool vec) {
  if (ok0 && ok1 && vec) {
    __stcs(reinterpret_cast<double2*>(dst), make_double2(e0, e1));
  } else {
    if (ok0) __stcs(dst, e0);
    if (ok1) __stcs(dst + 1, e1);
  }
}

// rows: point r (global index) in block P; cols: points c, c+1 in block Q
template <int LAYOUT, int P, int Q, bool FULL, bool DIAGMASK>
__device__ __forceinline__ void emit_block(const AsmParams& a, int r, int c, const double* h1, const double* h2,
                                           const double* h1b, const double* h2b, double kap0, double kap1) {
  bool ok0 = true, ok1 = true;
  if (!FULL) {
    if (r >= a.size[P]) return;
    ok0 = c < a.size[Q];
    ok1 = c + 1 < a.size[Q];
  }
  if (DIAGMASK && P == Q) { ok0 = ok0 && (c <= r); ok1 = ok1 && (c + 1 <= r); }
  if (!ok0 && !ok1) return;
  constexpr int OPX = lay_op(LAYOUT, P), OPY = lay_op(LAYOUT, Q);
  const double e0 = prefactor<OPX, OPY>(h1, h2) * kap0;
  const double e1 = prefactor<OPX, OPY>(h1b, h2b) * kap1;
  double* dst = a.T + (long)(a.off[P] + r) * a.ld + a.off[Q] + c;
  store_pair(dst, e0, e1, ok0, ok1, ((a.off[Q] | (int)(a.ld & 1)) & 1) == 0);
}

// MIRROR = false: all blocks p >= q (phase A).  MIRROR = true: only p > q (phase B).
template <int LAYOUT, int P, int Q, bool FULL, bool DIAGMASK, bool MIRROR>
struct EmitAll {
  __device__ static __forceinline__ void run(const AsmParams& a, int r, int c, const double* h1, const double* h2,
                                             const double* h1b, const double* h2b, double k0, double k1) {
    if ((!MIRROR && Q <= P) || (MIRROR && Q < P)) emit_block<LAYOUT, P, Q, FULL, DIAGMASK>(a, r, c, h1, h2, h1b, h2b, k0, k1);
    if constexpr (Q + 1 < lay_nblk(LAYOUT)) EmitAll<LAYOUT, P, Q + 1, FULL, DIAGMASK, MIRROR>::run(a, r, c, h1, h2, h1b, h2b, k0, k1);
    else if constexpr (P + 1 < lay_nblk(LAYOUT)) EmitAll<LAYOUT, P + 1, 0, FULL, DIAGMASK, MIRROR>::run(a, r, c, h1, h2, h1b, h2b, k0, k1);
  }
};

template <int LAYOUT, int KERNEL, bool FULL, bool DIAG>
__device__ __forceinline__ void gram_tile(const AsmParams& a, int i0, int j0, double* sm) {
  double* su1 = sm;                       // [TP][SPAD]
  double* su2 = sm + TP * SPAD;
  double* skp = sm + 2 * TP * SPAD;
  double* sx = sm + 3 * TP * SPAD;        // row points [TP][2], col points [TP][2]
  const int ntot = a.N + a.Nb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int r = threadIdx.x >> 1, c = threadIdx.x & 1;   // 256 threads: 64 rows x 2 + 64 cols x 2
    if (r < TP) sx[r * 2 + c] = (i0 + r < ntot) ? a.X[(long)(i0 + r) * 2 + c] : 0.0;
    else sx[2 * TP + (r - TP) * 2 + c] = (j0 + (r - TP) < ntot) ? a.X[(long)(j0 + r - TP) * 2 + c] : 0.0;
  }
  __syncthreads();
  // ---- phase A
  {
    const int jl = 2 * lane;
    const double y1a = sx[2 * TP + jl * 2], y2a = sx[2 * TP + jl * 2 + 1];
    const double y1b = sx[2 * TP + jl * 2 + 2], y2b = sx[2 * TP + jl * 2 + 3];
#pragma unroll 2
    for (int rr = 0; rr < TP / 8; ++rr) {
      const int il = warp + 8 * rr;
      const int i = i0 + il, j = j0 + jl;
      if (!FULL && (i >= ntot || j >= ntot)) continue;
      const double x1 = sx[il * 2], x2 = sx[il * 2 + 1];
      double h1[5], h2[5], h1b[5], h2b[5];
      const double d1 = x1 - y1a, d2 = x2 - y2a;
      const double u1 = a.k.b1 * d1, u2 = a.k.b2 * d2;
      hermite(u1, a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1);
      hermite(u2, a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2);
      const double k0 = kappa_of<KERNEL>(a.k, d1, d2);
      const double d1b = x1 - y1b, d2b = x2 - y2b;
      const double u1b = a.k.b1 * d1b, u2b = a.k.b2 * d2b;
      hermite(u1b, a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1b);
      hermite(u2b, a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2b);
      const double k1 = kappa_of<KERNEL>(a.k, d1b, d2b);
      EmitAll<LAYOUT, 0, 0, FULL, DIAG, false>::run(a, i, j, h1, h2, h1b, h2b, k0, k1);
      if (!DIAG) {
        su1[il * SPAD + jl] = u1; su1[il * SPAD + jl + 1] = u1b;
        su2[il * SPAD + jl] = u2; su2[il * SPAD + jl + 1] = u2b;
        skp[il * SPAD + jl] = k0; skp[il * SPAD + jl + 1] = k1;
      }
    }
  }
  if (DIAG) return;
  __syncthreads();
  // ---- phase B: rows = column points j, cols = row points i (two per lane), d -> -d
  {
    const int il = 2 * lane;
#pragma unroll 2
    for (int rr = 0; rr < TP / 8; ++rr) {
      const int jl = warp + 8 * rr;
      const int i = i0 + il, j = j0 + jl;
      if (!FULL && (i >= ntot || j >= ntot)) continue;
      double h1[5], h2[5], h1b[5], h2b[5];
      hermite(-su1[il * SPAD + jl], a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1);
      hermite(-su2[il * SPAD + jl], a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2);
      hermite(-su1[(il + 1) * SPAD + jl], a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1b);
      hermite(-su2[(il + 1) * SPAD + jl], a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2b);
      EmitAll<LAYOUT, 0, 0, FULL, false, true>::run(a, j, i, h1, h2, h1b, h2b, skp[il * SPAD + jl], skp[(il + 1) * SPAD + jl]);
    }
  }
}

template <int LAYOUT, int KERNEL>
__global__ void __launch_bounds__(256, 2)
gram_assemble_kernel(const __grid_constant__ AsmParams a) {
  extern __shared__ double asm_smem[];
  // lower-triangular tile enumeration: blockIdx.x -> (I, J), J <= I
  const long t = blockIdx.x;
  int I = static_cast<int>((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
  while ((long)(I + 1) * (I + 2) / 2 <= t) ++I;
  while ((long)I * (I + 1) / 2 > t) --I;
  const int J = static_cast<int>(t - (long)I * (I + 1) / 2);
  const int i0 = I * TP, j0 = J * TP;
  const bool full = (i0 + TP <= a.N) && (j0 + TP <= a.N);   // all 64 x 64 points interior: every block is valid
  if (I == J) {
    if (full) gram_tile<LAYOUT, KERNEL, true, true>(a, i0, j0, asm_smem);
    else gram_tile<LAYOUT, KERNEL, false, true>(a, i0, j0, asm_smem);
  } else {
    if (full) gram_tile<LAYOUT, KERNEL, true, false>(a, i0, j0, asm_smem);
    else gram_tile<LAYOUT, KERNEL, false, false>(a, i0, j0, asm_smem);
  }
}

// ---- prediction: out[t] = sum_q sum_j [id_x L_q^y kappa](x_t, x_j) * w[off_q + j] --------------
struct PredParams {
  KParams k;
  const double* X; int N, Nb;
  int nblk; int op[GPP_MAX_BLOCKS]; int off[GPP_MAX_BLOCKS]; int size[GPP_MAX_BLOCKS];
  const double* xt; int ntest;
  const double* w; double* out;     // predict
  double* Tt; long ldo;             // theta_test materialisation
};

__device__ __forceinline__ double prefactor_y(int op, const double* h1, const double* h2) {
  switch (op) {
    case OP_ID:  return prefactor<OP_ID, OP_ID>(h1, h2);
    case OP_D1:  return prefactor<OP_ID, OP_D1>(h1, h2);
    case OP_D2:  return prefactor<OP_ID, OP_D2>(h1, h2);
    case OP_D22: return prefactor<OP_ID, OP_D22>(h1, h2);
    default:     return prefactor<OP_ID, OP_LAP>(h1, h2);
  }
}

// one warp per test point, lanes stride over collocation points, fixed-order warp reduction
template <int KERNEL>
__global__ void __launch_bounds__(256)
predict_kernel(const __grid_constant__ PredParams p) {
  const int t = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (t >= p.ntest) return;
  const double x1 = p.xt[(long)t * 2], x2 = p.xt[(long)t * 2 + 1];
  const int ntot = p.N + p.Nb;
  double acc = 0.0;
  for (int j = lane; j < ntot; j += 32) {
    const double d1 = x1 - p.X[(long)j * 2], d2 = x2 - p.X[(long)j * 2 + 1];
    double h1[5], h2[5];
    hermite(p.k.b1 * d1, p.k.b1, p.k.c3b1, p.k.c6b1, p.k.c3bb1, h1);
    hermite(p.k.b2 * d2, p.k.b2, p.k.c3b2, p.k.c6b2, p.k.c3bb2, h2);
    const double kap = kappa_of<KERNEL>(p.k, d1, d2);
    for (int q = 0; q < p.nblk; ++q)
      if (j < p.size[q]) acc += (prefactor_y(p.op[q], h1, h2) * kap) * p.w[p.off[q] + j];
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) p.out[t] = acc;
}

template <int KERNEL>
__global__ void __launch_bounds__(256)
theta_test_kernel(const __grid_constant__ PredParams p) {
  const int ntot = p.N + p.Nb;
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (long)p.ntest * ntot) return;
  const int t = (int)(e / ntot), j = (int)(e % ntot);
  const double d1 = p.xt[(long)t * 2] - p.X[(long)j * 2], d2 = p.xt[(long)t * 2 + 1] - p.X[(long)j * 2 + 1];
  double h1[5], h2[5];
  hermite(p.k.b1 * d1, p.k.b1, p.k.c3b1, p.k.c6b1, p.k.c3bb1, h1);
  hermite(p.k.b2 * d2, p.k.b2, p.k.c3b2, p.k.c6b2, p.k.c3bb2, h2);
  const double kap = kappa_of<KERNEL>(p.k, d1, d2);
  for (int q = 0; q < p.nblk; ++q)
    if (j < p.size[q]) p.Tt[(long)t * p.ldo + p.off[q] + j] = prefactor_y(p.op[q], h1, h2) * kap;
}

KParams make_kparams(const GramSlot& s) {
  KParams k;
  k.b1 = s.kp_b1; k.b2 = s.kp_b2;
  k.c3b1 = 3.0 * k.b1; k.c6b1 = 6.0 * k.b1; k.c3bb1 = 3.0 * k.b1 * k.b1;
  k.c3b2 = 3.0 * k.b2; k.c6b2 = 6.0 * k.b2; k.c3bb2 = 3.0 * k.b2 * k.b2;
  k.e1 = s.kp_e1; k.e2 = s.kp_e2;
  return k;
}

PredParams make_pred(gpp_handle* h, GramSlot& s) {
  PredParams p{};
  p.k = make_kparams(s);
  p.X = h->Xall; p.N = s.N; p.Nb = s.Nb;
  p.nblk = s.lay.nblk;
  for (int q = 0; q < s.lay.nblk; ++q) {
    p.op[q] = s.lay.op[q]; p.off[q] = s.off[q];
    p.size[q] = s.N + (s.lay.with_bdy[q] ? s.Nb : 0);
  }
  return p;
}

template <int LAYOUT>
int launch_asm(gpp_handle* h, GramSlot& s, const AsmParams& a) {
  const int ntot = s.N + s.Nb;
  const long nt = (ntot + TP - 1) / TP;
  const unsigned grid = (unsigned)(nt * (nt + 1) / 2);
  if (s.kernel_id == 0) {
    CUDA_TRY(h, cudaFuncSetAttribute(gram_assemble_kernel<LAYOUT, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, ASM_SMEM));
    gram_assemble_kernel<LAYOUT, 0><<<grid, 256, ASM_SMEM, h->stream>>>(a);
  } else {
    CUDA_TRY(h, cudaFuncSetAttribute(gram_assemble_kernel<LAYOUT, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, ASM_SMEM));
    gram_assemble_kernel<LAYOUT, 1><<<grid, 256, ASM_SMEM, h->stream>>>(a);
  }
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

}  // namespace

int gram_assemble(gpp_handle* h, GramSlot& s) {
  AsmParams a{};
  a.k = make_kparams(s);
  a.X = h->Xall; a.N = s.N; a.Nb = s.Nb; a.T = s.T; a.ld = s.ld;
  for (int q = 0; q < s.lay.nblk; ++q) {
    a.off[q] = s.off[q];
    a.size[q] = s.N + (s.lay.with_bdy[q] ? s.Nb : 0);
  }
  switch (s.layout_id) {
    case LAY_ELLIPTIC: return launch_asm<LAY_ELLIPTIC>(h, s, a);
    case LAY_BURGERS:  return launch_asm<LAY_BURGERS>(h, s, a);
    case LAY_EIKONAL:  return launch_asm<LAY_EIKONAL>(h, s, a);
    case LAY_DARCY_A:  return launch_asm<LAY_DARCY_A>(h, s, a);
  }
  h->err = "bad layout";
  return -1;
}

int gram_predict(gpp_handle* h, GramSlot& s, const double* d_xtest, int ntest, const double* d_w, double* d_out) {
  PredParams p = make_pred(h, s);
  p.xt = d_xtest; p.ntest = ntest; p.w = d_w; p.out = d_out;
  if (s.kernel_id == 0) predict_kernel<0><<<(ntest + 7) / 8, 256, 0, h->stream>>>(p);
  else predict_kernel<1><<<(ntest + 7) / 8, 256, 0, h->stream>>>(p);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

int gram_theta_test(gpp_handle* h, GramSlot& s, const double* d_xtest, int ntest, double* d_out, long ldo) {
  PredParams p = make_pred(h, s);
  p.xt = d_xtest; p.ntest = ntest; p.Tt = d_out; p.ldo = ldo;
  const long tot = (long)ntest * (s.N + s.Nb);
  if (s.kernel_id == 0) theta_test_kernel<0><<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(p);
  else theta_test_kernel<1><<<(unsigned)((tot + 255) / 256), 256, 0, h->stream>>>(p);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

// ---- scalar functionals for the kernel classes (src/kernels.py methods), vectorised over n pairs ----
namespace {
struct EvalParams {
  KParams k; int opx, opy; const double* x1; const double* x2; const double* y1; const double* y2; long n; double* out;
};
template <int KERNEL>
__global__ void kernel_eval_kernel(const EvalParams p) {
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= p.n) return;
  const double d1 = p.x1[e] - p.y1[e], d2 = p.x2[e] - p.y2[e];
  double h1[5], h2[5];
  hermite(p.k.b1 * d1, p.k.b1, p.k.c3b1, p.k.c6b1, p.k.c3bb1, h1);
  hermite(p.k.b2 * d2, p.k.b2, p.k.c3b2, p.k.c6b2, p.k.c3bb2, h2);
  double poly = 0.0;
  for (int a = 0; a < mono_n(p.opx); ++a)
    for (int b = 0; b < mono_n(p.opy); ++b) {
      const int m1 = mono_m1(p.opx, a), m2 = mono_m2(p.opx, a), n1 = mono_m1(p.opy, b), n2 = mono_m2(p.opy, b);
      const double sg = ((m1 + m2) & 1) ? -1.0 : 1.0;
      poly = poly + (sg * h1[m1 + n1]) * h2[m2 + n2];
    }
  p.out[e] = poly * kappa_of<KERNEL>(p.k, d1, d2);
}
}  // namespace

int gram_kernel_eval(gpp_handle* h, int kernel, const double* kparams, int opx, int opy, const double* d_in, long n,
                     double* d_out) {
  GramSlot tmp;
  tmp.kp_b1 = kparams[0]; tmp.kp_b2 = kparams[1]; tmp.kp_e1 = kparams[2]; tmp.kp_e2 = kparams[3];
  EvalParams p{};
  p.k = make_kparams(tmp); p.opx = opx; p.opy = opy;
  p.x1 = d_in; p.x2 = d_in + n; p.y1 = d_in + 2 * n; p.y2 = d_in + 3 * n; p.n = n; p.out = d_out;
  if (kernel == 0) kernel_eval_kernel<0><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(p);
  else kernel_eval_kernel<1><<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(p);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return GPP_OK;
}

// ---- row-panel assembly for the row-sharded (multi-GPU) layout ------------------------------------------
// Rows [i_begin, i_begin + nrows) of row-operator block PROW, every column block q <= PROW (q == PROW: j <= i),
// written to dst (row 0 of dst = point i_begin).  Ordered pairs: no mirrored writes, so a rank touches only
// the rows it owns and the assembly needs no exchange.
namespace {
struct RowParams {
  AsmParams a;
  int i_begin, nrows;
  double* dst; long ldd;
};

template <int LAYOUT, int P, int Q>
struct EmitRow {
  __device__ static __forceinline__ void run(const RowParams& r, double* drow, int i, int j, const double* h1, const double* h2,
                                             const double* h1b, const double* h2b, double k0, double k1) {
    if (Q <= P) {
      bool ok0 = j < r.a.size[Q], ok1 = j + 1 < r.a.size[Q];
      if (P == Q) { ok0 = ok0 && (j <= i); ok1 = ok1 && (j + 1 <= i); }
      if (ok0 || ok1) {
        constexpr int OPX = lay_op(LAYOUT, P), OPY = lay_op(LAYOUT, Q);
        const double e0 = prefactor<OPX, OPY>(h1, h2) * k0;
        const double e1 = prefactor<OPX, OPY>(h1b, h2b) * k1;
        store_pair(drow + r.a.off[Q] + j, e0, e1, ok0, ok1, ((r.a.off[Q] | (int)(r.ldd & 1)) & 1) == 0);
      }
    }
    if constexpr (Q + 1 <= P) EmitRow<LAYOUT, P, Q + 1>::run(r, drow, i, j, h1, h2, h1b, h2b, k0, k1);
  }
};

template <int LAYOUT, int KERNEL, int P>
__global__ void __launch_bounds__(256)
gram_rows_kernel(const __grid_constant__ RowParams r) {
  const AsmParams& a = r.a;
  const int ntot = a.N + a.Nb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 64 + 2 * lane;
  if (j >= ntot) return;
  const bool v1 = j + 1 < ntot;
  const double y1a = a.X[(long)j * 2], y2a = a.X[(long)j * 2 + 1];
  const double y1b = v1 ? a.X[(long)(j + 1) * 2] : 0.0, y2b = v1 ? a.X[(long)(j + 1) * 2 + 1] : 0.0;
  for (int rr = warp; rr < 64; rr += 8) {
    const int il = blockIdx.y * 64 + rr;
    if (il >= r.nrows) break;
    const int i = r.i_begin + il;
    const double x1 = a.X[(long)i * 2], x2 = a.X[(long)i * 2 + 1];
    double h1[5], h2[5], h1b[5], h2b[5];
    const double d1 = x1 - y1a, d2 = x2 - y2a;
    hermite(a.k.b1 * d1, a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1);
    hermite(a.k.b2 * d2, a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2);
    const double k0 = kappa_of<KERNEL>(a.k, d1, d2);
    const double d1b = x1 - y1b, d2b = x2 - y2b;
    hermite(a.k.b1 * d1b, a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1b);
    hermite(a.k.b2 * d2b, a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2b);
    const double k1 = v1 ? kappa_of<KERNEL>(a.k, d1b, d2b) : 0.0;
    EmitRow<LAYOUT, P, 0>::run(r, r.dst + (long)il * r.ldd, i, j, h1, h2, h1b, h2b, k0, k1);
  }
}

template <int LAYOUT, int P>
int launch_rows_p(gpp_handle* h, GramSlot& s, const RowParams& r) {
  if constexpr (P < lay_nblk(LAYOUT)) {
    // columns needed: blocks q <= P; the widest is q = P itself (j <= i) or any earlier block (all N points)
    const int ncol = (P == 0) ? (r.i_begin + r.nrows) : r.a.size[0] > r.a.size[P] ? r.a.size[0] : r.a.size[P];
    dim3 grid((ncol + 63) / 64, (r.nrows + 63) / 64);
    if (s.kernel_id == 0) gram_rows_kernel<LAYOUT, 0, P><<<grid, 256, 0, h->cur>>>(r);
    else gram_rows_kernel<LAYOUT, 1, P><<<grid, 256, 0, h->cur>>>(r);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    return GPP_OK;
  } else {
    h->err = "bad row block";
    return -1;
  }
}

template <int LAYOUT>
int launch_rows(gpp_handle* h, GramSlot& s, int prow, const RowParams& r) {
  switch (prow) {
    case 0: return launch_rows_p<LAYOUT, 0>(h, s, r);
    case 1: return launch_rows_p<LAYOUT, 1>(h, s, r);
    case 2: return launch_rows_p<LAYOUT, 2>(h, s, r);
    case 3: return launch_rows_p<LAYOUT, 3>(h, s, r);
  }
  h->err = "bad row block";
  return -1;
}
}  // namespace

int gram_assemble_rows(gpp_handle* h, GramSlot& s, int prow, int i_begin, int nrows, double* dst, long ld) {
  if (nrows <= 0) return GPP_OK;
  RowParams r{};
  r.a.k = make_kparams(s);
  r.a.X = h->Xall; r.a.N = s.N; r.a.Nb = s.Nb; r.a.T = nullptr; r.a.ld = ld;
  for (int q = 0; q < s.lay.nblk; ++q) {
    r.a.off[q] = s.off[q];
    r.a.size[q] = s.N + (s.lay.with_bdy[q] ? s.Nb : 0);
  }
  r.i_begin = i_begin; r.nrows = nrows; r.dst = dst; r.ldd = ld;
  switch (s.layout_id) {
    case LAY_ELLIPTIC: return launch_rows<LAY_ELLIPTIC>(h, s, prow, r);
    case LAY_BURGERS:  return launch_rows<LAY_BURGERS>(h, s, prow, r);
    case LAY_EIKONAL:  return launch_rows<LAY_EIKONAL>(h, s, prow, r);
    case LAY_DARCY_A:  return launch_rows<LAY_DARCY_A>(h, s, prow, r);
  }
  h->err = "bad layout";
  return -1;
}
