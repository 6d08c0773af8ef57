/* CPU restatement of the device exp used by the Darcy residual kernel (TEST INFRASTRUCTURE ONLY).
 *
 * The reference evaluates jnp.exp(-w0) (src/InverseProblems.py:114) with XLA's exp, which is not
 * available here and is not bit-identical to either glibc's or CUDA's exp.  To make the residual
 * kernels bit-exact between host and device (north_star), both sides use this one algorithm:
 * Cody-Waite reduction x = k ln2 + r, degree-13 Taylor polynomial in Horner form with explicit fma,
 * scaling by ldexp.  Same constants, same operation order as gpp_exp in
 * nonlinpdes-gpsolver_b200/csrc/gn.cu.  Build: gcc -O2 -ffp-contract=off -shared -fPIC (see Makefile). */
#include <math.h>

double gpp_exp_ref(double x) {
  if (x != x) return x;
  if (x > 709.782712893384) return INFINITY;
  if (x < -745.1332191019412) return 0.0;
  const double L2E = 1.4426950408889634074;
  const double LN2HI = 6.93147180369123816490e-01, LN2LO = 1.90821492927058770002e-10;
  const double kf = rint(x * L2E);
  double r = fma(-kf, LN2HI, x);
  r = fma(-kf, LN2LO, r);
  double p = 1.0 / 6227020800.0;
  p = fma(p, r, 1.0 / 479001600.0);
  p = fma(p, r, 1.0 / 39916800.0);
  p = fma(p, r, 1.0 / 3628800.0);
  p = fma(p, r, 1.0 / 362880.0);
  p = fma(p, r, 1.0 / 40320.0);
  p = fma(p, r, 1.0 / 5040.0);
  p = fma(p, r, 1.0 / 720.0);
  p = fma(p, r, 1.0 / 120.0);
  p = fma(p, r, 1.0 / 24.0);
  p = fma(p, r, 1.0 / 6.0);
  p = fma(p, r, 0.5);
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  return ldexp(p, (int)kf);
}

void gpp_exp_ref_array(const double* x, double* out, long n) {
  for (long i = 0; i < n; ++i) out[i] = gpp_exp_ref(x[i]);
}
