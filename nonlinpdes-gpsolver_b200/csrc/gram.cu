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

constexpr int TI = 64, TJ = 64;   // pair tile: 64 rows x 64 cols; 8 warps, each warp one row at a time, lane -> 2 cols

template <int LAYOUT, int P, int Q>
__device__ __forceinline__ void emit_block(const AsmParams& a, int i, int j, const double* h1, const double* h2,
                                           const double* h1b, const double* h2b, double kap0, double kap1, bool v1) {
  // entry (row block P, point i ; col block Q, points j and j+1)
  if (i >= a.size[P] || j >= a.size[Q]) return;
  if (P == Q && j > i) return;
  constexpr int OPX = lay_op(LAYOUT, P), OPY = lay_op(LAYOUT, Q);
  const double e0 = prefactor<OPX, OPY>(h1, h2) * kap0;
  double* dst = a.T + (long)(a.off[P] + i) * a.ld + a.off[Q] + j;
  bool two = v1 && (j + 1 < a.size[Q]) && !(P == Q && j + 1 > i);
  if (two) {
    const double e1 = prefactor<OPX, OPY>(h1b, h2b) * kap1;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
      *reinterpret_cast<double2*>(dst) = make_double2(e0, e1);
    } else {
      dst[0] = e0; dst[1] = e1;
    }
  } else {
    dst[0] = e0;
  }
}

template <int LAYOUT, int P, int Q>
struct EmitAll {
  __device__ static __forceinline__ void run(const AsmParams& a, int i, int j, const double* h1, const double* h2,
                                             const double* h1b, const double* h2b, double k0, double k1, bool v1) {
    if (Q <= P) emit_block<LAYOUT, P, Q>(a, i, j, h1, h2, h1b, h2b, k0, k1, v1);
    if constexpr (Q + 1 < lay_nblk(LAYOUT)) EmitAll<LAYOUT, P, Q + 1>::run(a, i, j, h1, h2, h1b, h2b, k0, k1, v1);
    else if constexpr (P + 1 < lay_nblk(LAYOUT)) EmitAll<LAYOUT, P + 1, 0>::run(a, i, j, h1, h2, h1b, h2b, k0, k1, v1);
  }
};

template <int LAYOUT, int KERNEL>
__global__ void __launch_bounds__(256)
gram_assemble_kernel(const __grid_constant__ AsmParams a) {
  __shared__ double sxi[TI][2];
  const int ntot = a.N + a.Nb;
  const int i0 = blockIdx.y * TI, j0 = blockIdx.x * TJ;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < TI * 2) {
    int r = threadIdx.x >> 1, c = threadIdx.x & 1;
    sxi[r][c] = (i0 + r < ntot) ? a.X[(long)(i0 + r) * 2 + c] : 0.0;
  }
  __syncthreads();
  const int j = j0 + 2 * lane;
  if (j >= ntot) return;
  const bool v1 = (j + 1 < ntot);
  const double y1a = a.X[(long)j * 2], y2a = a.X[(long)j * 2 + 1];
  const double y1b = v1 ? a.X[(long)(j + 1) * 2] : 0.0, y2b = v1 ? a.X[(long)(j + 1) * 2 + 1] : 0.0;
  for (int rr = warp; rr < TI; rr += 8) {
    const int i = i0 + rr;
    if (i >= ntot) break;
    const double x1 = sxi[rr][0], x2 = sxi[rr][1];
    double h1[5], h2[5], h1b[5], h2b[5];
    const double d1 = x1 - y1a, d2 = x2 - y2a;
    hermite(a.k.b1 * d1, a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1);
    hermite(a.k.b2 * d2, a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2);
    const double k0 = kappa_of<KERNEL>(a.k, d1, d2);
    const double d1b = x1 - y1b, d2b = x2 - y2b;
    hermite(a.k.b1 * d1b, a.k.b1, a.k.c3b1, a.k.c6b1, a.k.c3bb1, h1b);
    hermite(a.k.b2 * d2b, a.k.b2, a.k.c3b2, a.k.c6b2, a.k.c3bb2, h2b);
    const double k1 = kappa_of<KERNEL>(a.k, d1b, d2b);
    EmitAll<LAYOUT, 0, 0>::run(a, i, j, h1, h2, h1b, h2b, k0, k1, v1);
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
  dim3 grid((ntot + TJ - 1) / TJ, (ntot + TI - 1) / TI);
  if (s.kernel_id == 0) gram_assemble_kernel<LAYOUT, 0><<<grid, 256, 0, h->stream>>>(a);
  else gram_assemble_kernel<LAYOUT, 1><<<grid, 256, 0, h->stream>>>(a);
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
