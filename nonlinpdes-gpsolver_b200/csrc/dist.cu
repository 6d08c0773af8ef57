// Multi-GPU path: one process per GPU, block rows of Theta dealt cyclically to the ranks
// (a P x 1 process grid of the 2-D block-cyclic family), NCCL over NVLink for the one exchange per
// block column.  The reference has no multi-device code at all: this is new.
//
//   assembly : every Gram entry depends on two points only -> each rank fills the block rows it owns
//              (gram_assemble_rows), no exchange.
//   Cholesky : left-looking by block column j.  The owner of block row j brings its diagonal block up to
//              date, factorises it, and broadcasts the finished block row L[j, 0:(j+1)NB] (ncclBroadcast);
//              every rank then updates and solves its own block rows below j with the same DMMA GEMM and
//              substitution kernels as the single-GPU path.  Work per step is ~ (#block rows below j) / P.
#include "../../include/gpp.h"
#include "gpp_internal.cuh"

#include <nccl.h>

#include <cstring>
#include <vector>

namespace {

struct DistState {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // row-sharded Theta / L of slot 0
  double* Tloc = nullptr;
  long ld = 0;
  int M = 0, nloc = 0;          // global size, local row count
  TMap2 mapLoc;
  double* rowbuf = nullptr;     // NB x M staging of the broadcast block row (two buffers alternate)
  double* rowbuf2 = nullptr;
  double* dvec = nullptr;       // M doubles scratch
  bool factored = false;
};

DistState* ds(gpp_handle* h) { return static_cast<DistState*>(h->dist); }

#define NCCL_TRY(h, expr)                                                      \
  do {                                                                         \
    ncclResult_t _r = (expr);                                                  \
    if (_r != ncclSuccess) {                                                   \
      (h)->err = std::string(#expr) + ": " + ncclGetErrorString(_r);           \
      return GPP_CUDA_ERR + 2;                                                 \
    }                                                                          \
  } while (0)

// block row b (global) -> owner, local block index
inline int owner_of(int b, int P) { return b % P; }
inline int local_blk(int b, int P) { return b / P; }
// number of block rows <= j owned by rank
inline int owned_upto(int j, int rank, int P) { return j >= rank ? (j - rank) / P + 1 : 0; }

__global__ void dist_get_diag_kernel(const double* __restrict__ T, long ld, int lrow0, int g0, int nrows, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) out[g0 + i] = T[(long)(lrow0 + i) * ld + g0 + i];
}
__global__ void dist_add_diag_kernel(double* __restrict__ T, long ld, int lrow0, int g0, int nrows, const double* __restrict__ add) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nrows) T[(long)(lrow0 + i) * ld + g0 + i] += add[g0 + i];
}

}  // namespace

extern "C" {

int gpp_dist_unique_id(unsigned char* id128) {
  if (!id128) return -1;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return GPP_CUDA_ERR + 2;
  memcpy(id128, &id, 128);
  return GPP_OK;
}

int gpp_dist_init(gpp_handle* h, int rank, int world, const unsigned char* id128) {
  if (h) cudaSetDevice(h->device);
  if (!h || !id128) return -1;
  if (world < 1 || rank < 0 || rank >= world) { h->err = "bad rank/world"; return -2; }
  if (h->dist) { h->err = "already initialised"; return -3; }
  CUDA_TRY(h, cudaSetDevice(h->device));
  DistState* d = new DistState();
  d->rank = rank; d->world = world;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  NCCL_TRY(h, ncclCommInitRank(&d->comm, world, id, rank));
  h->dist = d;
  return GPP_OK;
}

int gpp_dist_finalize(gpp_handle* h) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist) return -1;
  DistState* d = ds(h);
  cudaStreamSynchronize(h->stream);
  if (d->comm) ncclCommDestroy(d->comm);
  if (d->Tloc) cudaFree(d->Tloc);
  if (d->rowbuf) cudaFree(d->rowbuf);
  if (d->rowbuf2) cudaFree(d->rowbuf2);
  if (d->dvec) cudaFree(d->dvec);
  delete d;
  h->dist = nullptr;
  return GPP_OK;
}

// Row-sharded Gram_matrix_assembly (src/Gram_matrice.py:11-187): this rank's block rows of Theta.
int gpp_dist_gram_assemble(gpp_handle* h, int layout, int kernel, const double* kparams) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist) return -1;
  if (!h->Xall) { h->err = "points not set"; return -1; }
  if (layout < 0 || layout > 3 || kernel < 0 || kernel > 1 || !kparams) return -2;
  CUDA_TRY(h, cudaSetDevice(h->device));
  DistState* d = ds(h);
  GramSlot& s = h->slot[0];
  s.layout_id = layout; s.lay = make_layout(layout);
  s.N = h->N; s.Nb = (layout == LAY_DARCY_A) ? 0 : h->Nb;
  int o = 0;
  for (int p = 0; p < s.lay.nblk; ++p) { s.off[p] = o; o += s.N + (s.lay.with_bdy[p] ? s.Nb : 0); }
  s.off[s.lay.nblk] = o;
  s.M = o; s.Mint = s.lay.nblk * s.N;
  s.kernel_id = kernel;
  s.kp_b1 = kparams[0]; s.kp_b2 = kparams[1]; s.kp_e1 = kparams[2]; s.kp_e2 = kparams[3];
  const int NB = h->NB, P = d->world, M = s.M;
  const int nblk = (M + NB - 1) / NB;
  int nloc = 0;
  for (int b = d->rank; b < nblk; b += P) nloc += (M - b * NB < NB) ? (M - b * NB) : NB;
  const long ld = round_up(M, 16);
  if (!d->Tloc || d->M != M || d->nloc != nloc) {
    if (d->Tloc) cudaFree(d->Tloc);
    if (d->rowbuf) cudaFree(d->rowbuf);
    if (d->rowbuf2) cudaFree(d->rowbuf2);
    if (d->dvec) cudaFree(d->dvec);
    d->M = M; d->nloc = nloc; d->ld = ld;
    CUDA_TRY(h, cudaMalloc(&d->Tloc, sizeof(double) * (size_t)(nloc > 0 ? nloc : 1) * ld));
    CUDA_TRY(h, cudaMalloc(&d->rowbuf, sizeof(double) * (size_t)NB * ld));
    CUDA_TRY(h, cudaMalloc(&d->rowbuf2, sizeof(double) * (size_t)NB * ld));
    CUDA_TRY(h, cudaMalloc(&d->dvec, sizeof(double) * (size_t)M));
    if (nloc > 0) {
      int rc = make_tensor_map(h, &d->mapLoc, d->Tloc, nloc, M, ld);
      if (rc) return rc;
    }
  }
  d->factored = false;
  // owned block row b covers global rows [b NB, b NB + nb); split at the row-operator block boundaries
  for (int b = d->rank; b < nblk; b += P) {
    const int g0 = b * NB, g1 = (g0 + NB < M) ? g0 + NB : M;
    const int l0 = local_blk(b, P) * NB;
    for (int p = 0; p < s.lay.nblk; ++p) {
      const int lo = g0 > s.off[p] ? g0 : s.off[p];
      const int hi = g1 < s.off[p + 1] ? g1 : s.off[p + 1];
      if (hi <= lo) continue;
      int rc = gram_assemble_rows(h, s, p, lo - s.off[p], hi - lo, d->Tloc + (long)(l0 + lo - g0) * ld, ld);
      if (rc) return rc;
    }
  }
  return GPP_OK;
}

int gpp_dist_local_rows(gpp_handle* h, int* nloc, int* M) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist) return -1;
  if (nloc) *nloc = ds(h)->nloc;
  if (M) *M = ds(h)->M;
  return GPP_OK;
}

// diag_out[M]: the full diagonal on every rank (sum all-reduce of the owned parts)
int gpp_dist_get_diag(gpp_handle* h, double* diag_out) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist || !diag_out) return -1;
  DistState* d = ds(h);
  const int NB = h->NB, P = d->world, M = d->M;
  CUDA_TRY(h, cudaMemsetAsync(d->dvec, 0, sizeof(double) * M, h->stream));
  for (int b = d->rank; b * NB < M; b += P) {
    const int g0 = b * NB, nr = (M - g0 < NB) ? (M - g0) : NB;
    dist_get_diag_kernel<<<(nr + 255) / 256, 256, 0, h->stream>>>(d->Tloc, d->ld, local_blk(b, P) * NB, g0, nr, d->dvec);
    h->launches++;
  }
  NCCL_TRY(h, ncclAllReduce(d->dvec, d->dvec, M, ncclDouble, ncclSum, d->comm, h->stream));
  CUDA_TRY(h, cudaMemcpyAsync(diag_out, d->dvec, sizeof(double) * M, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GPP_OK;
}

int gpp_dist_add_diag(gpp_handle* h, const double* add) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist || !add) return -1;
  DistState* d = ds(h);
  const int NB = h->NB, P = d->world, M = d->M;
  CUDA_TRY(h, cudaMemcpyAsync(d->dvec, add, sizeof(double) * M, cudaMemcpyHostToDevice, h->stream));
  for (int b = d->rank; b * NB < M; b += P) {
    const int g0 = b * NB, nr = (M - g0 < NB) ? (M - g0) : NB;
    dist_add_diag_kernel<<<(nr + 255) / 256, 256, 0, h->stream>>>(d->Tloc, d->ld, local_blk(b, P) * NB, g0, nr, d->dvec);
    h->launches++;
  }
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  return GPP_OK;
}

// Update + solve + own-diagonal update of local rows [lrow, lrow + nrows) in block column j against the received block
// row (rowbuf holds L[j, 0:(j+1)NB] packed with leading dimension ldrow), on h->cur.
static int dist_step_rows(gpp_handle* h, DistState* d, int j, int lrow, int nrows, double* rowbuf) {
  if (nrows <= 0) return GPP_OK;
  const int NB = h->NB, M = d->M;
  const int j0 = j * NB;
  const int nbj = (M - j0 < NB) ? (M - j0) : NB;
  const long ldrow = j0 + NB;
  Mat Mloc{d->Tloc, d->ld, &d->mapLoc};
  TMap2 mapRow;
  int rc = make_tensor_map(h, &mapRow, rowbuf, nbj, j0 + nbj, ldrow);
  if (rc) return rc;
  if (j > 0) {
    GemmDesc g{};
    g.mapA = &d->mapLoc; g.mapB = &mapRow;
    g.a_row0 = lrow; g.b_row0 = 0;
    g.C = d->Tloc + (long)lrow * d->ld + j0; g.ldc = d->ld; g.Cin = g.C; g.ldcin = d->ld;
    g.m = nrows; g.n = nbj; g.k0 = 0; g.k1 = j0; g.kb_off = 0; g.alpha = -1.0; g.lower_only = 0;
    rc = gemm_nt_launch(h, g);
    if (rc) return rc;
  }
  Mat Row{rowbuf, ldrow, &mapRow};
  rc = trsm_right_lt(h, Mloc, lrow, j0, nrows, Row, 0, j0, nbj);
  if (rc) return rc;
  // right-looking update of this rank's own diagonal blocks with the freshly solved column j: keeps every diagonal
  // block up to date (in column order), so its owner can factorise it the moment its turn comes
  GemmDesc g{};
  g.mapA = &d->mapLoc; g.mapB = &d->mapLoc;
  g.C = d->Tloc; g.ldc = d->ld; g.Cin = d->Tloc; g.ldcin = d->ld;
  g.k0 = j0; g.k1 = j0 + nbj; g.kb_off = 0; g.alpha = -1.0;
  g.bd_count = (nrows + NB - 1) / NB; g.bd_world = d->world; g.bd_rank = d->rank; g.bd_lblk0 = lrow / NB; g.bd_nb = NB; g.bd_M = M;
  return gemm_nt_launch(h, g);
}

// factorise the (up-to-date) diagonal block of block row j and pack L[j, 0:(j+1)NB] into rowbuf, on h->cur
static int dist_factor_and_pack(gpp_handle* h, DistState* d, int j, double* rowbuf) {
  const int NB = h->NB, M = d->M, P = d->world;
  const int j0 = j * NB;
  const int nbj = (M - j0 < NB) ? (M - j0) : NB;
  const long ldrow = j0 + NB;
  const int lr = local_blk(j, P) * NB;
  Mat Mloc{d->Tloc, d->ld, &d->mapLoc};
  int rc = potrf_diag(h, Mloc, lr, j0, nbj, j0);
  if (rc) return rc;
  CUDA_TRY(h, cudaMemcpy2DAsync(rowbuf, ldrow * 8, d->Tloc + (long)lr * d->ld, d->ld * 8, (size_t)(j0 + nbj) * 8, nbj,
                                cudaMemcpyDeviceToDevice, h->cur));
  return GPP_OK;
}

// Distributed X.Gram_Cholesky (src/PDEs.py:75-80).  *info as in gpp_potrf, identical on every rank.
//
// Look-ahead: the owner of block row j+1 treats that block row first (high-priority stream sP): update + solve in
// column j, factorise its diagonal block, pack, broadcast -- while every rank is still busy with the bulk of
// column j on the main stream.  Two row buffers alternate; events order bulk(j) after bcast(j) and bcast(j+2) after
// bulk(j).  NCCL calls are issued in the same order (j = 0, 1, ...) on the side stream of every rank.
int gpp_dist_potrf(gpp_handle* h, int* info) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist) return -1;
  DistState* d = ds(h);
  if (!d->Tloc) { h->err = "assemble first"; return -2; }
  if (d->factored) { h->err = "already factored"; return -3; }
  CUDA_TRY(h, cudaSetDevice(h->device));
  const int NB = h->NB, P = d->world, M = d->M, rank = d->rank;
  const int nblk = (M + NB - 1) / NB;
  cudaStream_t S = h->stream, Sp = h->sP;
  double* rb[2] = {d->rowbuf, d->rowbuf2};
  CUDA_TRY(h, cudaMemsetAsync(h->d_info, 0, sizeof(int), S));
  size_t used = 0;
  auto new_event = [&]() {
    if (used == h->evpool.size()) {
      cudaEvent_t e;
      cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      h->evpool.push_back(e);
    }
    return h->evpool[used++];
  };
  constexpr int NREG = 3;
  auto clampi = [&](long v) { return (int)(v > d->nloc ? d->nloc : v); };
  const int bound[NREG + 1] = {0, clampi(round_up(d->nloc / 2, NB)), clampi(round_up((3L * d->nloc) / 4, NB)), d->nloc};
  std::vector<cudaEvent_t> ev_bc(nblk), ev_bulk(NREG * nblk);
  cudaEvent_t ev0 = new_event();
  CUDA_TRY(h, cudaEventRecord(ev0, S));
  CUDA_TRY(h, cudaStreamWaitEvent(Sp, ev0, 0));
  for (int k = 0; k < 2; ++k) CUDA_TRY(h, cudaStreamWaitEvent(h->sG[k], ev0, 0));
  auto bcast = [&](int j) -> int {     // on Sp
    const int j0 = j * NB;
    const int nbj = (M - j0 < NB) ? (M - j0) : NB;
    if (P > 1) NCCL_TRY(h, ncclBroadcast(rb[j & 1], rb[j & 1], (size_t)nbj * (j0 + NB), ncclDouble, owner_of(j, P), d->comm, Sp));
    ev_bc[j] = new_event();
    CUDA_TRY(h, cudaEventRecord(ev_bc[j], Sp));
    return GPP_OK;
  };
  int rc = GPP_OK;
  // prologue: block row 0
  h->cur = Sp;
  if (rank == owner_of(0, P)) rc = dist_factor_and_pack(h, d, 0, rb[0]);
  if (!rc) rc = bcast(0);
  for (int j = 0; j < nblk && !rc; ++j) {
    const int lstart = owned_upto(j, rank, P) * NB;        // first local row below block row j
    int bulk_start = lstart;
    if (j + 1 < nblk) {
      // rb[(j+1)&1] was last read by bulk(j-1); block row j+1's earlier columns were written by bulk(<= j-1)
      if (j >= 1) {
        for (int k = 0; k < NREG; ++k) CUDA_TRY(h, cudaStreamWaitEvent(Sp, ev_bulk[NREG * (j - 1) + k], 0));
      }
      if (rank == owner_of(j + 1, P)) {
        h->cur = Sp;
        const int nb1 = (M - (j + 1) * NB < NB) ? (M - (j + 1) * NB) : NB;
        rc = dist_step_rows(h, d, j, lstart, nb1, rb[j & 1]);   // block row j+1 is this rank's first row block below j
        if (!rc) rc = dist_factor_and_pack(h, d, j + 1, rb[(j + 1) & 1]);
        if (rc) break;
        bulk_start = lstart + NB;
      }
      rc = bcast(j + 1);
      if (rc) break;
    }
    // bulk of column j: the remaining rows, cut at fixed local-row boundaries into regions that stay on the same
    // stream for the whole factorisation.  A row only depends on its own history and on the received block row, so
    // region A may start column j+1 while region B still finishes column j: the second stream fills the SMs the last
    // wave of the other one leaves idle.
    for (int k = 0; k < NREG && !rc; ++k) {
      const int lo = bulk_start > bound[k] ? bulk_start : bound[k];
      const int hi = bound[k + 1];
      cudaStream_t sg = h->sG[k & 1];
      CUDA_TRY(h, cudaStreamWaitEvent(sg, ev_bc[j], 0));
      h->cur = sg;
      if (hi > lo) rc = dist_step_rows(h, d, j, lo, hi - lo, rb[j & 1]);
      ev_bulk[NREG * j + k] = new_event();
      CUDA_TRY(h, cudaEventRecord(ev_bulk[NREG * j + k], sg));
    }
  }
  h->cur = S;
  if (rc) return rc;
  {
    cudaEvent_t e = new_event();
    CUDA_TRY(h, cudaEventRecord(e, Sp));
    CUDA_TRY(h, cudaStreamWaitEvent(S, e, 0));
    for (int k = 0; k < 2; ++k) {
      cudaEvent_t eg = new_event();
      CUDA_TRY(h, cudaEventRecord(eg, h->sG[k]));
      CUDA_TRY(h, cudaStreamWaitEvent(S, eg, 0));
    }
  }
  if (P > 1) NCCL_TRY(h, ncclAllReduce(h->d_info, h->d_info, 1, ncclInt, ncclMax, d->comm, S));
  int hinfo = 0;
  CUDA_TRY(h, cudaMemcpyAsync(&hinfo, h->d_info, sizeof(int), cudaMemcpyDeviceToHost, S));
  CUDA_TRY(h, cudaStreamSynchronize(S));
  if (info) *info = hinfo;
  d->factored = true;
  return GPP_OK;
}

// out: nloc x M dense rows of this rank (lower triangle of Theta / L; entries above the diagonal are zeroed)
int gpp_dist_download_local(gpp_handle* h, double* out) {
  if (h) cudaSetDevice(h->device);
  if (!h || !h->dist || !out) return -1;
  DistState* d = ds(h);
  const int NB = h->NB, P = d->world, M = d->M;
  CUDA_TRY(h, cudaMemcpy2DAsync(out, (size_t)M * 8, d->Tloc, d->ld * 8, (size_t)M * 8, d->nloc, cudaMemcpyDeviceToHost, h->stream));
  CUDA_TRY(h, cudaStreamSynchronize(h->stream));
  long lr = 0;
  for (int b = d->rank; b * NB < M; b += P) {
    const int g0 = b * NB, nr = (M - g0 < NB) ? (M - g0) : NB;
    for (int i = 0; i < nr; ++i, ++lr)
      for (int c = g0 + i + 1; c < M; ++c) out[lr * M + c] = 0.0;
  }
  return GPP_OK;
}

}  // extern "C"
