// Is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) on sm_100a a chain of IEEE FMAs?  One warp computes D = A B + C for many random
// 8x4 / 4x8 / 8x8 operands with a wide dynamic range and compares every element of D bit by bit with
//   fma(a3,b3, fma(a2,b2, fma(a1,b1, fma(a0,b0, c))))  in all 24 orders of k,
// and measures the error of DMMA and of the ascending FMA chain against a double-double reference.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o dmma_exactness dmma_exactness.cu && ./dmma_exactness
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__global__ void dmma_kernel(const double* A, const double* B, const double* C, double* D, int ntile) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  for (int tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const double* a = A + (long)tile * 32;   // 8 x 4 row-major
    const double* b = B + (long)tile * 32;   // 4 x 8 row-major (k x n)
    const double* c = C + (long)tile * 64;   // 8 x 8
    double* d = D + (long)tile * 64;
    double av = a[g * 4 + t];                // A fragment: row g, col t
    double bv = b[t * 8 + g];                // B fragment: row t (k), col g
    double c0 = c[g * 8 + 2 * t], c1 = c[g * 8 + 2 * t + 1];
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv));
    d[g * 8 + 2 * t] = c0; d[g * 8 + 2 * t + 1] = c1;
  }
}

static void two_prod(double a, double b, double& p, double& e) { p = a * b; e = fma(a, b, -p); }
static void two_sum(double a, double b, double& s, double& e) { s = a + b; double bb = s - a; e = (a - (s - bb)) + (b - bb); }

int main() {
  const int ntile = 20000;
  double *hA = (double*)malloc(sizeof(double) * 32 * ntile), *hB = (double*)malloc(sizeof(double) * 32 * ntile);
  double *hC = (double*)malloc(sizeof(double) * 64 * ntile), *hD = (double*)malloc(sizeof(double) * 64 * ntile);
  srand(1);
  auto rnd = [](int spread) { double m = (double)rand() / RAND_MAX * 2 - 1; int e = rand() % (2 * spread + 1) - spread; return ldexp(m, e); };
  for (int i = 0; i < 32 * ntile; ++i) { hA[i] = rnd(8); hB[i] = rnd(8); }
  for (int i = 0; i < 64 * ntile; ++i) hC[i] = rnd(12);
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, sizeof(double) * 32 * ntile); cudaMalloc(&dB, sizeof(double) * 32 * ntile);
  cudaMalloc(&dC, sizeof(double) * 64 * ntile); cudaMalloc(&dD, sizeof(double) * 64 * ntile);
  cudaMemcpy(dA, hA, sizeof(double) * 32 * ntile, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(double) * 32 * ntile, cudaMemcpyHostToDevice);
  cudaMemcpy(dC, hC, sizeof(double) * 64 * ntile, cudaMemcpyHostToDevice);
  dmma_kernel<<<64, 32>>>(dA, dB, dC, dD, ntile);
  if (cudaMemcpy(hD, dD, sizeof(double) * 64 * ntile, cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
  int perm[24][4], np = 0;
  for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) for (int c = 0; c < 4; ++c) for (int d = 0; d < 4; ++d)
    if (a != b && a != c && a != d && b != c && b != d && c != d) { perm[np][0] = a; perm[np][1] = b; perm[np][2] = c; perm[np][3] = d; ++np; }
  long match[24] = {0}, total = 0, any = 0;
  double err_dmma = 0, err_fma = 0, bias_dmma = 0, bias_fma = 0;
  for (int tile = 0; tile < ntile; ++tile)
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) {
      const double* a = hA + (long)tile * 32 + i * 4; const double* b = hB + (long)tile * 32; const double c = hC[(long)tile * 64 + i * 8 + j];
      const double got = hD[(long)tile * 64 + i * 8 + j];
      bool hit = false;
      for (int p = 0; p < 24; ++p) {
        double acc = c;
        for (int s = 0; s < 4; ++s) { int k = perm[p][s]; acc = fma(a[k], b[k * 8 + j], acc); }
        if (acc == got) { match[p]++; hit = true; }
      }
      any += hit; ++total;
      // double-double reference of c + sum a_k b_k
      double hi = c, lo = 0;
      for (int k = 0; k < 4; ++k) { double p, pe, s, se; two_prod(a[k], b[k * 8 + j], p, pe); two_sum(hi, p, s, se); hi = s; lo += se + pe; }
      const double ref = hi + lo, ulp = fabs(ref) > 0 ? ldexp(1.0, ilogb(ref) - 52) : 1e-300;
      double accf = c; for (int k = 0; k < 4; ++k) accf = fma(a[k], b[k * 8 + j], accf);
      const double ed = ((got - hi) - lo) / ulp, ef = ((accf - hi) - lo) / ulp;
      err_dmma += fabs(ed); err_fma += fabs(ef); bias_dmma += ed; bias_fma += ef;
    }
  printf("elements %ld; DMMA result equals SOME order of an FMA chain: %.4f %%\n", total, 100.0 * any / total);
  for (int p = 0; p < 24; ++p) printf("  k order %d%d%d%d: %.4f %%\n", perm[p][0], perm[p][1], perm[p][2], perm[p][3], 100.0 * match[p] / total);
  printf("mean |error| in ulp of the exact result: DMMA %.4f, FMA chain (k ascending) %.4f;  mean signed error: DMMA %+.5f, FMA chain %+.5f\n",
         err_dmma / total, err_fma / total, bias_dmma / total, bias_fma / total);
  return 0;
}
