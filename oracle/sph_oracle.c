/* ----------------------------------------------------------------------------
 * sph_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 * See sph_oracle.h for the scope statement and the pinning status.
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC)
 * -ffp-contract=off keeps the reference's generic-x86-64 (no FMA) rounding.
 * --------------------------------------------------------------------------*/
#include "sph_oracle.h"
#include <math.h>
#include <float.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- Utilities/SpheralFunctions.hh:30-52,86-92 ; Utilities/safeInv.hh:13-27 ; FastMath.hh:203-223 ----*/
static inline double orc_sgn(double x) { return x < 0.0 ? -1.0 : 1.0; }
static inline double orc_sq(double x) { return x*x; }
static inline double orc_pow4(double x) { return x*x*x*x; }
static inline double orc_safeInv(double x, double fuzz) { return x/(x*x + fuzz); }
static inline double orc_safeInvVar(double x) { return orc_sgn(x)/fmax(1.0e-30, fabs(x)); }
static inline int    orc_fuzzyEqual(double a, double b, double fuzz) {
  return fabs(a - b) <= fuzz*fmax(1.0, fabs(a) + fabs(b));
}

/* ---- analytic kernels ------------------------------------------------------------
 * BSpline:    Kernel/BSplineKernelInline.hh:8-90
 * WendlandC4: Kernel/WendlandC4KernelInline.hh:8-95
 * WendlandC2: Kernel/WendlandC2KernelInline.hh:8-90                                  */
/* ---- NBSplineKernel(order): Kernel/NBSplineKernel.cc:17-122, NBSplineKernelInline.hh:55-97 ----------------------------------
 * W(eta) = A/(k-1)! * sum_{i=0..k} (-1)^i C(k,i) (eta - i + k/2)_+^(k-1), k = order + 1; the derivatives lower the exponent and the
 * factorial; kernel extent (order + 1)/2 in INTEGER arithmetic; A from simpsonsVolumeIntegral with 10000 bins
 * (Kernel/VolumeIntegrationFunctions.cc:22-70, Utilities/simpsonsIntegration.hh:20-53). */
static int nbs_factorial(int n) { if (n < 0) return 2147483647; int r = 1; for (int i = 1; i < n + 1; ++i) r *= i; return r; }
static double nbs_sum(int order, double eta, int lower) {      /* lower = 1, 2, 3: value, gradient, second derivative (unnormalised) */
  const int k = order + 1;
  const int e = (k - lower > -lower + 1) ? k - lower : -lower + 1;           /* max(0,k-1), max(-1,k-2), max(-2,k-3) */
  const double halfk = 0.5*k;
  double result = 0.0;
  for (int i = 0; i <= k; ++i) {
    const double x = eta - i + halfk;
    const int binom = nbs_factorial(k)/(nbs_factorial(k - i)*nbs_factorial(i));
    result += pow(-1.0, i)*binom*(x >= 0.0 ? pow(x, e) : 0.0);
  }
  return result/nbs_factorial(e);
}
static double nbs_volume_normalization(int order, int ndim) {
  static double cache[16][4];
  if (order < 16 && cache[order][ndim] != 0.0) return cache[order][ndim];
  const double kext = (double)((order + 1)/2);
  const unsigned numBins = 10000u;
  const double dx = (kext - 0.0)/numBins;
  double result = 0.0;
  for (unsigned i = 0; i < numBins + 1u; ++i) {
    const double r = 0.0 + i*dx;
    const double ve = (ndim == 1 ? 2.0 : ndim == 2 ? 2.0*M_PI*r : 4.0*M_PI*r*r);
    const double integrand = ve*((r >= kext) ? 0.0 : nbs_sum(order, r, 1)*1.0*1.0);
    if (i == 0 || i == numBins) result += integrand;
    else if (i % 2 == 0) result += 2.0*integrand;
    else result += 4.0*integrand;
  }
  result *= dx/3.0;
  const double A = 1.0/result;
  if (order < 16) cache[order][ndim] = A;
  return A;
}

double orc_kernel_extent(int kind, int ndim) {
  (void)ndim;
  if (kind >= ORC_KERNEL_NBSPLINE) return (double)((kind - ORC_KERNEL_NBSPLINE + 1)/2);
  switch (kind) {
    case ORC_KERNEL_BSPLINE: return 2.0;
    case ORC_KERNEL_WENDLANDC4: return 1.0;
    case ORC_KERNEL_WENDLANDC2: return 1.0;
    default: return 0.0;
  }
}

void orc_kernel_analytic(int kind, int ndim, double eta, double* W, double* gradW, double* grad2W) {
  double w = 0.0, g = 0.0, g2 = 0.0;
  if (kind == ORC_KERNEL_BSPLINE) {
    const double A = (ndim == 1 ? 2.0/3.0 : ndim == 2 ? 10.0/(7.0*M_PI) : 1.0/M_PI);
    if (eta < 1.0) {
      const double eta2 = eta*eta;
      w  = A*1.0*(1.0 - 1.5*eta2 + 0.75*eta2*eta);
      g  = -A*1.0*(3.0 - 2.25*eta)*eta;
      g2 = -A*1.0*(3 - 4.5*eta);
    } else if (eta < 2.0) {
      const double t = 2.0 - eta;
      w  = A*1.0*0.25*(t*t*t);
      g  = -A*1.0*0.75*(t*t);
      g2 = A*1.0*1.5*(2 - eta);
    }
  } else if (kind == ORC_KERNEL_WENDLANDC4) {
    const double A = (ndim == 1 ? 3.0/2.0 : ndim == 2 ? 9.0/M_PI : 495.0/(32.0*M_PI));
    const double in = (eta < 1.0) ? 1.0 : 0.0;
    const double eta2 = eta*eta;
    if (ndim == 1) {
      w  = A*1.0*(pow(1.0 - eta, 5)*(1.0 + 5.0*eta + 8.0*eta2))*in;
      g  = A*1.0*(-14.0*pow(1.0 - eta, 4)*eta*(1.0 + 4.0*eta))*in;
      g2 = A*1.0*(-14.0*pow(eta - 1.0, 3)*(24.0*eta2 - 3.0*eta - 1.0))*in;
    } else {
      w  = A*1.0*(pow(1.0 - eta, 6)*(1.0 + 6.0*eta + (35.0/3.0)*eta2))*in;
      g  = A*1.0*((56.0/3.0)*pow(eta - 1.0, 5)*eta*(5.0*eta + 1.0))*in;
      g2 = A*1.0*((56.0/3.0)*pow(eta - 1.0, 4)*(35.0*eta2 - 4.0*eta - 1.0))*in;
    }
  } else if (kind == ORC_KERNEL_WENDLANDC2) {
    const double A = (ndim == 1 ? 5.0/4.0 : ndim == 2 ? 7.0/M_PI : 21.0/(2.0*M_PI));
    const double in = (eta < 1.0) ? 1.0 : 0.0;
    const double eta2 = eta*eta;
    if (ndim == 1) {
      w  = A*1.0*(pow(1.0 - eta, 3)*(1.0 + 3.0*eta))*in;
      g  = A*1.0*(-12.0*pow(1.0 - eta, 2)*eta)*in;
      g2 = A*1.0*(-12.0*(3.0*eta2 - 4.0*eta + 1.0))*in;
    } else {
      w  = A*1.0*(pow(1.0 - eta, 4)*(1.0 + 4.0*eta))*in;
      g  = A*1.0*(20.0*pow(eta - 1.0, 3)*eta)*in;
      g2 = A*1.0*(20.0*pow(eta - 1.0, 2)*(4.0*eta - 1.0))*in;
    }
  }
  else if (kind >= ORC_KERNEL_NBSPLINE) {
    const int order = kind - ORC_KERNEL_NBSPLINE;
    if (eta < orc_kernel_extent(kind, ndim)) {
      const double A = nbs_volume_normalization(order, ndim);
      w = nbs_sum(order, eta, 1)*(A*1.0);
      g = nbs_sum(order, eta, 2)*(A*1.0);
      g2 = nbs_sum(order, eta, 3)*(A*1.0);
    }
  }
  if (W) *W = w;
  if (gradW) *gradW = g;
  if (grad2W) *grad2W = g2;
}

/* ---- QuadraticInterpolator::initialize -------------------------------------------
 * Utilities/QuadraticInterpolator.cc:59-100.  The 3x3 solve is Eigen 5.0.0's fixed-size
 * inverse (Eigen/src/LU/InverseImpl.h, compute_inverse<Matrix,3>: cofactors / determinant
 * expanded along column 0) followed by a dense 3x3 * 3x1 product.  Eigen is NOT vendored in
 * /root/reference (spack dependency eigen@5.0.0); the published cofactor algorithm is restated. */
static void eigen_inverse3_times(const double A[3][3], const double B[3], double X[3]) {
  double c[3][3];   /* c[i][j] = cofactor_3x3<i,j> */
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    c[i][j] = A[i1][j1]*A[i2][j2] - A[i1][j2]*A[i2][j1];
  }
  const double det = c[0][0]*A[0][0] + c[1][0]*A[1][0] + c[2][0]*A[2][0];
  const double invdet = 1.0/det;
  double inv[3][3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv[j][i] = c[i][j]*invdet;
  for (int r = 0; r < 3; ++r) X[r] = inv[r][0]*B[0] + inv[r][1]*B[1] + inv[r][2]*B[2];
}

void orc_quadratic_fit(double xmin, double xmax, size_t n, const double* yvals, double* coeffs, size_t* n1out, double* xstepout) {
  const size_t N1 = (n - 1u)/2u - 1u;
  const double xstep = (xmax - xmin)/(double)(N1 + 1u);
  for (size_t i0 = 0; i0 <= N1; ++i0) {
    const double x0 = xmin + (double)i0*xstep;
    const double x1 = x0 + 0.5*xstep;
    const double x2 = x0 + xstep;
    const double A[3][3] = {{1.0, x0, x0*x0}, {1.0, x1, x1*x1}, {1.0, x2, x2*x2}};
    const double B[3] = {yvals[2u*i0], yvals[2u*i0 + 1u], yvals[2u*i0 + 2u]};
    double X[3];
    eigen_inverse3_times(A, B, X);
    coeffs[3*i0] = X[0]; coeffs[3*i0 + 1] = X[1]; coeffs[3*i0 + 2] = X[2];
  }
  if (n1out) *n1out = N1;
  if (xstepout) *xstepout = xstep;
}

/* number of samples after the "make odd" rule (QuadraticInterpolatorInline.hh:31) */
static size_t odd_samples(size_t n) { return (n % 2 == 0) ? n + 1 : n; }

size_t orc_table_ncoef(size_t numPoints) {
  const size_t n = odd_samples(numPoints);
  const size_t N1 = (n - 1u)/2u - 1u;
  return 3u*(N1 + 1u);
}

/* TableKernel ctor (Kernel/TableKernel.cc:169-209): three QuadraticInterpolators on [0,kext] */
int orc_table_build(int kind, int ndim, size_t numPoints, double* Wc, double* gWc, double* g2Wc,
                    double* kextout, size_t* n1, double* xstep) {
  const double kext = orc_kernel_extent(kind, ndim);
  if (kext <= 0.0 || numPoints < 2) return 1;
  const size_t n = odd_samples(numPoints);
  const double step = (kext - 0.0)/(double)(n - 1u);
  double* y0 = (double*)malloc(n*8); double* y1 = (double*)malloc(n*8); double* y2 = (double*)malloc(n*8);
  for (size_t i = 0; i < n; ++i) orc_kernel_analytic(kind, ndim, 0.0 + (double)i*step, &y0[i], &y1[i], &y2[i]);
  orc_quadratic_fit(0.0, kext, n, y0, Wc, n1, xstep);
  orc_quadratic_fit(0.0, kext, n, y1, gWc, n1, xstep);
  orc_quadratic_fit(0.0, kext, n, y2, g2Wc, n1, xstep);
  free(y0); free(y1); free(y2);
  if (kextout) *kextout = kext;
  return 0;
}

/* QuadraticInterpolatorView::lowerBound / operator() (Utilities/QuadraticInterpolatorViewInline.hh:8-21,71-77) */
static inline double quad_eval(const double* c, double xmin, double xstep, size_t n1, double x) {
  const double t = fmax(0.0, x - xmin)/xstep;
  size_t k = (size_t)t; if (k > n1) k = n1;
  const size_t i0 = 3u*k;
  return c[i0] + (c[i0 + 1] + c[i0 + 2]*x)*x;
}

/* TableKernelView::kernelAndGradValue (Kernel/TableKernelViewInline.hh:84-99) */
void orc_table_eval(const orc_table* t, double eta, double Hdet, double* W, double* gW) {
  if (eta < t->kext) {
    const double q = fmax(0.0, eta - t->xmin)/t->xstep;
    size_t k = (size_t)q; if (k > t->n1) k = t->n1;
    const size_t i0 = 3u*k;
    *W  = Hdet*(t->Wcoef[i0] + (t->Wcoef[i0 + 1] + t->Wcoef[i0 + 2]*eta)*eta);
    *gW = Hdet*(t->gradWcoef[i0] + (t->gradWcoef[i0 + 1] + t->gradWcoef[i0 + 2]*eta)*eta);
  } else {
    *W = 0.0; *gW = 0.0;
  }
}

/* ---- CubicHermiteInterpolator ---------------------------------------------------------
 * eval: Utilities/CubicHermiteInterpolatorViewInline.hh:8-33,103-108
 * gradient knots: CubicHermiteInterpolator.cc:160-195 (Han & Guo 2018 tridiagonal system; the reference
 *   solves it with Eigen::SparseLU, restated here with the Thomas algorithm -- same linear system)
 * makeMonotonic: CubicHermiteInterpolator.cc:104-150 (Fritsch-Carlson)                               */
double orc_cubic_hermite_eval(size_t n, double xmin, double xmax, double xstep, const double* v, double x) {
  if (x < xmin) return v[0] + v[n]*(x - xmin);
  if (x > xmax) return v[n - 1u] + v[2u*n - 1u]*(x - xmin);      /* sic: (x - mXmin) in the reference */
  size_t i0 = (size_t)(fmax(0.0, x - xmin)/xstep); if (i0 > n - 2u) i0 = n - 2u;
  const double t = fmax(0.0, fmin(1.0, (x - xmin - (double)i0*xstep)/xstep));
  const double t2 = t*t, t3 = t*t2;
  return ((2.0*t3 - 3.0*t2 + 1.0)*v[i0] + (-2.0*t3 + 3.0*t2)*v[i0 + 1u] +
          xstep*((t3 - 2.0*t2 + t)*v[n + i0] + (t3 - t2)*v[n + i0 + 1u]));
}

static void hermite_gradient_knots(size_t n, double xstep, double* v) {
  double* a = (double*)malloc(n*8); double* b = (double*)malloc(n*8);
  double* c = (double*)malloc(n*8); double* r = (double*)malloc(n*8);
  for (size_t k = 0; k < n; ++k) { a[k] = -0.5; b[k] = 4.0; c[k] = -0.5; }
  c[0] = -1.0; a[n - 1] = -1.0;
  r[0] = 3.0*(v[1] - v[0])/xstep;
  r[n - 1] = 3.0*(v[n - 1] - v[n - 2])/xstep;
  for (size_t k = 1; k < n - 1; ++k) r[k] = 1.5*(v[k + 1] - v[k - 1])/xstep;
  for (size_t k = 1; k < n; ++k) { const double m = a[k]/b[k - 1]; b[k] -= m*c[k - 1]; r[k] -= m*r[k - 1]; }
  v[n + n - 1] = r[n - 1]/b[n - 1];
  for (size_t k = n - 1; k-- > 0;) v[n + k] = (r[k] - c[k]*v[n + k + 1])/b[k];
  free(a); free(b); free(c); free(r);
}

static void hermite_make_monotonic(size_t n, double xstep, double* v) {
  double* cg = (double*)malloc((n - 1)*8);
  const double dxInv = 1.0/xstep;
  for (size_t k = 0; k < n - 1; ++k) cg[k] = (v[k + 1] - v[k])*dxInv;
  for (size_t k = 1; k < n - 1; ++k) {
    if (cg[k - 1]*cg[k] <= 0.0) v[n + k] = 0.0;
    if (cg[k] == 0.0) { v[n + k] = 0.0; v[n + k + 1] = 0.0; }
  }
  int done = 0;
  while (!done) {
    done = 1;
    for (size_t k = 0; k < n - 1; ++k) {
      double alpha = v[n + k]/cg[k], beta = v[n + k + 1]/cg[k];
      if (alpha < 0.0) { v[n + k] = 0.0; alpha = 0.0; done = 0; }
      if (beta < 0.0) { v[n + k + 1] = 0.0; beta = 0.0; done = 0; }
      const double tau = 3.0/sqrt(alpha*alpha + beta*beta);
      if (tau < 1.0) { v[n + k] = 0.99*tau*alpha*cg[k]; v[n + k + 1] = 0.99*tau*beta*cg[k]; done = 0; }
    }
  }
  free(cg);
}

/* sumKernelValues (Kernel/TableKernel.cc:24-66) with kernelValueSPH (TableKernelViewInline.hh:118-127) */
static double sum_kernel_values(const orc_table* t, int ndim, double nPerh) {
  const double deta = 1.0/nPerh;
  double result = 0.0, etar = deta;
  while (etar < t->kext) {
    const double k = fabs(quad_eval(t->gradWcoef, t->xmin, t->xstep, t->n1, etar));
    if (ndim == 1) result += 2.0*k;
    else if (ndim == 2) result += 2.0*M_PI*etar/deta*k;
    else result += 4.0*M_PI*orc_sq(etar/deta)*k;
    etar += deta;
  }
  return ndim == 1 ? result : ndim == 2 ? sqrt(result) : pow(result, 1.0/3.0);
}

typedef struct { size_t n; double xmin, xmax, xstep; const double* v; double target; } bisect_ctx;
static double bisect_f(const bisect_ctx* c, double x) { return orc_cubic_hermite_eval(c->n, c->xmin, c->xmax, c->xstep, c->v, x) - c->target; }
/* Utilities/bisectRoot.hh:18-78 */
static double bisect_root(const bisect_ctx* c, double xmin, double xmax) {
  const double xacc = 1.0e-15, yacc = 1.0e-10;
  const double fxmin = bisect_f(c, xmin), fxmax = bisect_f(c, xmax);
  if (orc_fuzzyEqual(fxmin, 0.0, yacc)) return xmin;
  if (orc_fuzzyEqual(fxmax, 0.0, yacc)) return xmax;
  double x0, x1;
  if (fxmin < 0.0) { x0 = xmin; x1 = xmax; } else { x0 = xmax; x1 = xmin; }
  double rootSafe = 0.0;
  for (unsigned iter = 0; iter < 100u; ++iter) {
    const double dx = 0.5*(x1 - x0);
    rootSafe = x0 + dx;
    if (fabs(dx) <= xacc) return rootSafe;
    const double f = bisect_f(c, rootSafe);
    if (orc_fuzzyEqual(f, 0.0, yacc)) return rootSafe;
    if (f < 0.0) x0 = rootSafe; else x1 = rootSafe;
  }
  return rootSafe;
}

/* TableKernel.cc:181-208 */
int orc_table_build_nperh(const orc_table* t, int ndim, size_t numPoints, double minNperh, double maxNperh,
                          double* wsumVals, double* wsumRange, double* nperhVals, double* nperhRange) {
  double mn = fmax(minNperh, 1.1/t->kext), mx = maxNperh;
  if (mx <= mn) mx = 4.0*mn;
  const size_t n = numPoints;
  const double step = (mx - mn)/(double)(n - 1u);
  for (size_t i = 0; i < n; ++i) wsumVals[i] = sum_kernel_values(t, ndim, mn + (double)i*step);
  hermite_gradient_knots(n, step, wsumVals);
  const double w0 = orc_cubic_hermite_eval(n, mn, mx, step, wsumVals, mn);
  const double w1 = orc_cubic_hermite_eval(n, mn, mx, step, wsumVals, mx);
  const double wstep = (w1 - w0)/(double)(n - 1u);
  bisect_ctx c = { n, mn, mx, step, wsumVals, 0.0 };
  for (size_t i = 0; i < n; ++i) { c.target = w0 + (double)i*wstep; nperhVals[i] = bisect_root(&c, mn, mx); }
  hermite_gradient_knots(n, wstep, nperhVals);
  hermite_make_monotonic(n, step, wsumVals);
  hermite_make_monotonic(n, wstep, nperhVals);
  wsumRange[0] = mn; wsumRange[1] = mx; nperhRange[0] = w0; nperhRange[1] = w1;
  return 0;
}

/* ---- dimension-specific bodies ------------------------------------------------------*/
#define D 3
#include "sph_oracle_dim.inc"
#undef D
#define D 2
#include "sph_oracle_dim.inc"
#undef D
#define D 1
#include "sph_oracle_dim.inc"
#undef D

size_t orc_pairs_bruteforce(int ndim, size_t nInt, size_t nGhost, const double* pos, const double* H,
                            double kext, uint32_t* pi, uint32_t* pj, size_t cap, uint32_t* counts) {
  return ndim == 3 ? pairs_bruteforce_3d(nInt, nGhost, pos, H, kext, pi, pj, cap, counts)
                   : (ndim == 2 ? pairs_bruteforce_2d(nInt, nGhost, pos, H, kext, pi, pj, cap, counts)
                               : pairs_bruteforce_1d(nInt, nGhost, pos, H, kext, pi, pj, cap, counts));
}
size_t orc_pairs_cells(int ndim, size_t nInt, size_t nGhost, const double* pos, const double* H,
                       double kext, uint32_t* pi, uint32_t* pj, size_t cap, uint32_t* counts) {
  return ndim == 3 ? pairs_cells_3d(nInt, nGhost, pos, H, kext, pi, pj, cap, counts)
                   : (ndim == 2 ? pairs_cells_2d(nInt, nGhost, pos, H, kext, pi, pj, cap, counts)
                               : pairs_cells_1d(nInt, nGhost, pos, H, kext, pi, pj, cap, counts));
}
int orc_evaluate_derivatives(const orc_options* o, const orc_table* W, const orc_table* WQ,
                             size_t nInt, size_t nGhost, const orc_state* s,
                             size_t npairs, const uint32_t* pi, const uint32_t* pj,
                             const uint32_t* numNeighbors, orc_derivs* d, int nthreads) {
  if (o->compatibleEnergy && o->evolveTotalEnergy) return 2;      /* VERIFY2 at SPH.cc:97-98 */
  return o->ndim == 3 ? evaluate_derivatives_3d(o, W, WQ, nInt, nGhost, s, npairs, pi, pj, numNeighbors, d, nthreads)
                      : (o->ndim == 2 ? evaluate_derivatives_2d(o, W, WQ, nInt, nGhost, s, npairs, pi, pj, numNeighbors, d, nthreads)
                               : evaluate_derivatives_1d(o, W, WQ, nInt, nGhost, s, npairs, pi, pj, numNeighbors, d, nthreads));
}
int orc_update_energy_compatible(int ndim, size_t nInt, size_t nGhost, const double* mass, const double* vel,
                                 const double* DvDt, const double* DepsDt0, size_t npairs,
                                 const uint32_t* pi, const uint32_t* pj, const double* pacc,
                                 double multiplier, double* eps) {
  return ndim == 3 ? update_energy_3d(nInt, nGhost, mass, vel, DvDt, DepsDt0, npairs, pi, pj, pacc, multiplier, eps)
                   : (ndim == 2 ? update_energy_2d(nInt, nGhost, mass, vel, DvDt, DepsDt0, npairs, pi, pj, pacc, multiplier, eps)
                               : update_energy_1d(nInt, nGhost, mass, vel, DvDt, DepsDt0, npairs, pi, pj, pacc, multiplier, eps));
}

/* ---- CRKSPH (crk_oracle_dim.inc) ----------------------------------------------------------------*/
int orc_crk_sum_volume(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* H,
                       size_t npairs, const uint32_t* pi, const uint32_t* pj, double* vol) {
  return ndim == 3 ? crk_sum_volume_3d(W, nInt, nGhost, pos, H, npairs, pi, pj, vol)
                   : (ndim == 2 ? crk_sum_volume_2d(W, nInt, nGhost, pos, H, npairs, pi, pj, vol)
                               : crk_sum_volume_1d(W, nInt, nGhost, pos, H, npairs, pi, pj, vol));
}
int orc_crk_corrections(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* H,
                        const double* vol, size_t npairs, const uint32_t* pi, const uint32_t* pj, double* corr) {
  return ndim == 3 ? crk_corrections_3d(W, nInt, nGhost, pos, H, vol, npairs, pi, pj, corr)
                   : (ndim == 2 ? crk_corrections_2d(W, nInt, nGhost, pos, H, vol, npairs, pi, pj, corr)
                               : crk_corrections_1d(W, nInt, nGhost, pos, H, vol, npairs, pi, pj, corr));
}
int orc_crk_sum_density(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* mass,
                        const double* vol, const double* H, size_t npairs, const uint32_t* pi, const uint32_t* pj,
                        double rhoMin, double rhoMax, double* rho) {
  return ndim == 3 ? crk_sum_density_3d(W, nInt, nGhost, pos, mass, vol, H, npairs, pi, pj, rhoMin, rhoMax, rho)
                   : (ndim == 2 ? crk_sum_density_2d(W, nInt, nGhost, pos, mass, vol, H, npairs, pi, pj, rhoMin, rhoMax, rho)
                               : crk_sum_density_1d(W, nInt, nGhost, pos, mass, vol, H, npairs, pi, pj, rhoMin, rhoMax, rho));
}
int orc_crk_evaluate_derivatives(const orc_options* o, const orc_table* W, size_t nInt, size_t nGhost,
                                 const orc_state* s, const double* vol, const double* corr,
                                 size_t npairs, const uint32_t* pi, const uint32_t* pj, orc_derivs* d) {
  if (o->compatibleEnergy && o->evolveTotalEnergy) return 2;
  return o->ndim == 3 ? crk_evaluate_derivatives_3d(o, W, nInt, nGhost, s, vol, corr, npairs, pi, pj, d)
                      : (o->ndim == 2 ? crk_evaluate_derivatives_2d(o, W, nInt, nGhost, s, vol, corr, npairs, pi, pj, d)
                               : crk_evaluate_derivatives_1d(o, W, nInt, nGhost, s, vol, corr, npairs, pi, pj, d));
}
/* RKUtilities::evaluateKernelAndGradient for one point pair (test hook for the RK interpolation pinning test) */
void orc_rk_kernel_grad(int ndim, const orc_table* W, const double* x, const double* H, const double* corr,
                        double* WR, double* gradWR) {
  if (ndim == 3) rk_kernel_grad_3d(W, x, H, corr, WR, gradWR); else if (ndim == 2) rk_kernel_grad_2d(W, x, H, corr, WR, gradWR);
  else rk_kernel_grad_1d(W, x, H, corr, WR, gradWR);
}

/* ---- per-step callers (step_oracle_dim.inc) ------------------------------------------------------------------*/
/* helpers of sph_oracle_dim.inc that the step file needs again (those macros are undefined at this point) */
#define D 3
#include "step_oracle_dim.inc"
#undef D
#define D 2
#include "step_oracle_dim.inc"
#undef D
#define D 1
#include "step_oracle_dim.inc"
#undef D

int orc_sum_mass_density(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* mass,
                         const double* H, size_t npairs, const uint32_t* pi, const uint32_t* pj, double* rho) {
  return ndim == 3 ? sum_mass_density_3d(W, nInt, nGhost, pos, mass, H, npairs, pi, pj, rho)
                   : (ndim == 2 ? sum_mass_density_2d(W, nInt, nGhost, pos, mass, H, npairs, pi, pj, rho)
                               : sum_mass_density_1d(W, nInt, nGhost, pos, mass, H, npairs, pi, pj, rho));
}
int orc_omega_gradh(int ndim, const orc_table* W, size_t nInt, size_t nGhost, const double* pos, const double* H,
                    size_t npairs, const uint32_t* pi, const uint32_t* pj, const uint32_t* numNeighbors, double* omega) {
  return ndim == 3 ? omega_gradh_3d(W, nInt, nGhost, pos, H, npairs, pi, pj, numNeighbors, omega)
                   : (ndim == 2 ? omega_gradh_2d(W, nInt, nGhost, pos, H, npairs, pi, pj, numNeighbors, omega)
                               : omega_gradh_1d(W, nInt, nGhost, pos, H, npairs, pi, pj, numNeighbors, omega));
}
void orc_eos_gamma_law(const orc_step_options* so, size_t n, const double* rho, const double* eps, double* P, double* cs) {
  eos_gamma_law_3d(so, n, rho, eps, P, cs);
}
int orc_state_update(const orc_options* o, const orc_step_options* so, size_t nInt, size_t nGhost, double multiplier,
                     int timeAdvanceOnly, int epsDone, const orc_derivs* d,
                     double* pos, double* vel, double* H, double* rho, double* eps, double* P, double* cs) {
  return o->ndim == 3 ? state_update_3d(o, so, nInt, nGhost, multiplier, timeAdvanceOnly, epsDone, d, pos, vel, H, rho, eps, P, cs)
                      : (o->ndim == 2 ? state_update_2d(o, so, nInt, nGhost, multiplier, timeAdvanceOnly, epsDone, d, pos, vel, H, rho, eps, P, cs)
                               : state_update_1d(o, so, nInt, nGhost, multiplier, timeAdvanceOnly, epsDone, d, pos, vel, H, rho, eps, P, cs));
}
double orc_hydro_dt(const orc_options* o, const orc_step_options* so, size_t nInt, const double* vel, const double* H,
                    const double* rho, const double* cs, const orc_derivs* d, size_t npairs, const uint32_t* pi,
                    const uint32_t* pj, int* reason, uint32_t* node) {
  return o->ndim == 3 ? hydro_dt_3d(o, so, nInt, vel, H, rho, cs, d, npairs, pi, pj, reason, node)
                      : (o->ndim == 2 ? hydro_dt_2d(o, so, nInt, vel, H, rho, cs, d, npairs, pi, pj, reason, node)
                               : hydro_dt_1d(o, so, nInt, vel, H, rho, cs, d, npairs, pi, pj, reason, node));
}
void orc_sym_bound(int ndim, double* H, double minv, double maxv) {
  if (ndim == 3) sym_bound_3d(H, minv, maxv); else if (ndim == 2) sym_bound_2d(H, minv, maxv); else sym_bound_1d(H, minv, maxv);
}
