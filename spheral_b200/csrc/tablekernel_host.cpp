// tablekernel_host.cpp -- host-side construction of a TableKernel payload (product code, no GPU needed).
//
// When the package is driven by a real Spheral, the table comes from Spheral's own TableKernel through
// sphb200_set_kernel_table.  Stand-alone (tests, bench, the Python mirror) there is no Spheral, so this file builds
// the same payload: TableKernel<Dim>::TableKernel(kernel, numPoints, minNperh, maxNperh) (Kernel/TableKernel.cc:169-209)
//   = QuadraticInterpolator fits of W, gradW, grad2W on [0,kext]      (Utilities/QuadraticInterpolator.cc:59-100)
//   + CubicHermite lookups Wsum(nperh) and nperh(Wsum)                (Utilities/CubicHermiteInterpolator.cc:63-195)
#include "sphb200.h"
#include <cmath>
#include <functional>
#include <vector>

namespace {

// NBSplineKernel(order) (Kernel/NBSplineKernel.cc:17-122): Schoenberg's B-spline of order k = order + 1,
//   W(eta) = A/(k-1)! sum_{i=0..k} (-1)^i C(k,i) (eta - i + k/2)_+^(k-1),
// its derivatives by lowering the exponent and the factorial; kernel extent (order + 1)/2 in integer arithmetic (:116); the volume
// normalisation A by composite Simpson integration of W over the kernel volume with 10000 bins (:119-121,
// Kernel/VolumeIntegrationFunctions.cc:22-70, Utilities/simpsonsIntegration.hh:20-53) -- not the closed form, so that the table
// carries the reference's own normalisation error (~1e-9).
struct NBSpline {
  int order, ndim; double A = 1.0;
  static double fact(int n) { double r = 1.0; for (int i = 2; i <= n; ++i) r *= i; return r; }
  double extent() const { return double((order + 1)/2); }
  double sum(double eta, int lower) const {           // lower = 1, 2, 3: value, first, second derivative (unnormalised)
    const int k = order + 1, e = std::max(1 - lower, k - lower);
    double r = 0.0;
    for (int i = 0; i <= k; ++i) {
      const double x = eta - i + 0.5*k;
      const double binom = fact(k)/(fact(k - i)*fact(i));
      if (x >= 0.0) r += ((i & 1) ? -1.0 : 1.0)*binom*std::pow(x, e);
    }
    return r/fact(std::max(e, 0));
  }
  void normalise() {
    const unsigned bins = 10000u;
    const double kext = extent(), dx = kext/bins;
    double acc = 0.0;
    for (unsigned i = 0; i <= bins; ++i) {
      const double r = i*dx;
      const double shell = ndim == 1 ? 2.0 : (ndim == 2 ? 2.0*M_PI*r : 4.0*M_PI*r*r);
      const double f = shell*(r >= kext ? 0.0 : sum(r, 1));
      acc += (i == 0 || i == bins) ? f : ((i % 2 == 0) ? 2.0*f : 4.0*f);
    }
    A = 1.0/(acc*dx/3.0);
  }
};

struct Analytic {
  int kind, ndim;
  NBSpline nbs{0, 0};
  Analytic(int k, int d) : kind(k), ndim(d) {
    if (kind >= SPHB200_KERNEL_NBSPLINE) { nbs = NBSpline{kind - SPHB200_KERNEL_NBSPLINE, ndim}; nbs.normalise(); }
  }
  double extent() const { return kind >= SPHB200_KERNEL_NBSPLINE ? nbs.extent() : (kind == SPHB200_KERNEL_BSPLINE ? 2.0 : 1.0); }
  // value / first / second derivative at eta with Hdet = 1 (Kernel/BSplineKernelInline.hh:38-90,
  // WendlandC4KernelInline.hh:38-95, WendlandC2KernelInline.hh:36-90)
  void eval(double eta, double& w, double& g, double& g2) const {
    w = g = g2 = 0.0;
    if (kind >= SPHB200_KERNEL_NBSPLINE) {
      if (eta < nbs.extent()) { w = nbs.sum(eta, 1)*nbs.A; g = nbs.sum(eta, 2)*nbs.A; g2 = nbs.sum(eta, 3)*nbs.A; }
    } else if (kind == SPHB200_KERNEL_BSPLINE) {
      const double A = ndim == 1 ? 2.0/3.0 : (ndim == 2 ? 10.0/(7.0*M_PI) : 1.0/M_PI);
      if (eta < 1.0) {
        const double e2 = eta*eta;
        w = A*(1.0 - 1.5*e2 + 0.75*e2*eta); g = -A*(3.0 - 2.25*eta)*eta; g2 = -A*(3 - 4.5*eta);
      } else if (eta < 2.0) {
        const double t = 2.0 - eta;
        w = A*0.25*(t*t*t); g = -A*0.75*(t*t); g2 = A*1.5*(2 - eta);
      }
    } else if (kind == SPHB200_KERNEL_WENDLANDC4) {
      const double A = ndim == 1 ? 3.0/2.0 : (ndim == 2 ? 9.0/M_PI : 495.0/(32.0*M_PI));
      const double in = eta < 1.0 ? 1.0 : 0.0, e2 = eta*eta;
      if (ndim == 1) {
        w = A*(std::pow(1.0 - eta, 5)*(1.0 + 5.0*eta + 8.0*e2))*in;
        g = A*(-14.0*std::pow(1.0 - eta, 4)*eta*(1.0 + 4.0*eta))*in;
        g2 = A*(-14.0*std::pow(eta - 1.0, 3)*(24.0*e2 - 3.0*eta - 1.0))*in;
      } else {
        w = A*(std::pow(1.0 - eta, 6)*(1.0 + 6.0*eta + (35.0/3.0)*e2))*in;
        g = A*((56.0/3.0)*std::pow(eta - 1.0, 5)*eta*(5.0*eta + 1.0))*in;
        g2 = A*((56.0/3.0)*std::pow(eta - 1.0, 4)*(35.0*e2 - 4.0*eta - 1.0))*in;
      }
    } else if (kind == SPHB200_KERNEL_WENDLANDC2) {
      const double A = ndim == 1 ? 5.0/4.0 : (ndim == 2 ? 7.0/M_PI : 21.0/(2.0*M_PI));
      const double in = eta < 1.0 ? 1.0 : 0.0, e2 = eta*eta;
      if (ndim == 1) {
        w = A*(std::pow(1.0 - eta, 3)*(1.0 + 3.0*eta))*in;
        g = A*(-12.0*std::pow(1.0 - eta, 2)*eta)*in;
        g2 = A*(-12.0*(3.0*e2 - 4.0*eta + 1.0))*in;
      } else {
        w = A*(std::pow(1.0 - eta, 4)*(1.0 + 4.0*eta))*in;
        g = A*(20.0*std::pow(eta - 1.0, 3)*eta)*in;
        g2 = A*(20.0*std::pow(eta - 1.0, 2)*(4.0*eta - 1.0))*in;
      }
    }
  }
};

// exact parabola through three equally spaced samples via the inverse of the monomial Vandermonde matrix
// (cofactor expansion / determinant, the fixed-size 3x3 inverse Eigen uses for QuadraticInterpolator.cc:84-91)
struct Quadratic {
  double xmin = 0, xstep = 0; size_t n1 = 0; std::vector<double> c;
  void fit(double x0, double x1, const std::vector<double>& y) {
    const size_t n = y.size();
    n1 = (n - 1)/2 - 1;
    xmin = x0; xstep = (x1 - x0)/double(n1 + 1);
    c.assign(3*(n1 + 1), 0.0);
    for (size_t k = 0; k <= n1; ++k) {
      const double a = xmin + double(k)*xstep, b = a + 0.5*xstep, d = a + xstep;
      const double A[3][3] = {{1.0, a, a*a}, {1.0, b, b*b}, {1.0, d, d*d}};
      double cof[3][3];
      for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        cof[i][j] = A[i1][j1]*A[i2][j2] - A[i1][j2]*A[i2][j1];
      }
      const double invdet = 1.0/(cof[0][0]*A[0][0] + cof[1][0]*A[1][0] + cof[2][0]*A[2][0]);
      const double B[3] = {y[2*k], y[2*k + 1], y[2*k + 2]};
      for (int r = 0; r < 3; ++r)
        c[3*k + r] = (cof[0][r]*invdet)*B[0] + (cof[1][r]*invdet)*B[1] + (cof[2][r]*invdet)*B[2];
    }
  }
  double operator()(double x) const {
    size_t k = size_t(std::max(0.0, x - xmin)/xstep);
    if (k > n1) k = n1;
    return c[3*k] + (c[3*k + 1] + c[3*k + 2]*x)*x;
  }
};

struct Hermite {
  size_t n = 0; double xmin = 0, xmax = 0, xstep = 0; std::vector<double> v;   // n values then n gradients
  double operator()(double x) const {
    if (x < xmin) return v[0] + v[n]*(x - xmin);
    if (x > xmax) return v[n - 1] + v[2*n - 1]*(x - xmin);
    size_t i0 = size_t(std::max(0.0, x - xmin)/xstep);
    if (i0 > n - 2) i0 = n - 2;
    const double t = std::max(0.0, std::min(1.0, (x - xmin - double(i0)*xstep)/xstep)), t2 = t*t, t3 = t*t2;
    return (2.0*t3 - 3.0*t2 + 1.0)*v[i0] + (-2.0*t3 + 3.0*t2)*v[i0 + 1] + xstep*((t3 - 2.0*t2 + t)*v[n + i0] + (t3 - t2)*v[n + i0 + 1]);
  }
  void init(double x0, double x1, size_t np, const std::function<double(double)>& F) {
    n = np; xmin = x0; xmax = x1; xstep = (x1 - x0)/double(n - 1);
    v.assign(2*n, 0.0);
    for (size_t i = 0; i < n; ++i) v[i] = F(xmin + double(i)*xstep);
    // minimal-derivative-oscillation knots (Han & Guo 2018): tridiagonal system, Thomas algorithm
    std::vector<double> a(n, -0.5), b(n, 4.0), cc(n, -0.5), r(n);
    cc[0] = -1.0; a[n - 1] = -1.0;
    r[0] = 3.0*(v[1] - v[0])/xstep; r[n - 1] = 3.0*(v[n - 1] - v[n - 2])/xstep;
    for (size_t k = 1; k + 1 < n; ++k) r[k] = 1.5*(v[k + 1] - v[k - 1])/xstep;
    for (size_t k = 1; k < n; ++k) { const double m = a[k]/b[k - 1]; b[k] -= m*cc[k - 1]; r[k] -= m*r[k - 1]; }
    v[2*n - 1] = r[n - 1]/b[n - 1];
    for (size_t k = n - 1; k-- > 0;) v[n + k] = (r[k] - cc[k]*v[n + k + 1])/b[k];
  }
  void makeMonotonic() {            // Fritsch-Carlson
    std::vector<double> cg(n - 1);
    for (size_t k = 0; k + 1 < n; ++k) cg[k] = (v[k + 1] - v[k])*(1.0/xstep);
    for (size_t k = 1; k + 1 < n; ++k) {
      if (cg[k - 1]*cg[k] <= 0.0) v[n + k] = 0.0;
      if (cg[k] == 0.0) { v[n + k] = 0.0; v[n + k + 1] = 0.0; }
    }
    bool done = false;
    while (!done) {
      done = true;
      for (size_t k = 0; k + 1 < n; ++k) {
        double al = v[n + k]/cg[k], be = v[n + k + 1]/cg[k];
        if (al < 0.0) { v[n + k] = 0.0; al = 0.0; done = false; }
        if (be < 0.0) { v[n + k + 1] = 0.0; be = 0.0; done = false; }
        const double tau = 3.0/std::sqrt(al*al + be*be);
        if (tau < 1.0) { v[n + k] = 0.99*tau*al*cg[k]; v[n + k + 1] = 0.99*tau*be*cg[k]; done = false; }
      }
    }
  }
};

inline bool fuzzyEq(double a, double b, double f) { return std::fabs(a - b) <= f*std::max(1.0, std::fabs(a) + std::fabs(b)); }

double bisect(const std::function<double(double)>& f, double xmin, double xmax) {    // Utilities/bisectRoot.hh
  const double fa = f(xmin), fb = f(xmax);
  if (fuzzyEq(fa, 0.0, 1e-10)) return xmin;
  if (fuzzyEq(fb, 0.0, 1e-10)) return xmax;
  double x0 = fa < 0.0 ? xmin : xmax, x1 = fa < 0.0 ? xmax : xmin, root = x0;
  for (unsigned it = 0; it < 100; ++it) {
    const double dx = 0.5*(x1 - x0);
    root = x0 + dx;
    if (std::fabs(dx) <= 1e-15) return root;
    const double fr = f(root);
    if (fuzzyEq(fr, 0.0, 1e-10)) return root;
    if (fr < 0.0) x0 = root; else x1 = root;
  }
  return root;
}

}  // namespace

extern "C" size_t sphb200_table_ncoef(size_t numPoints) {
  const size_t n = (numPoints % 2 == 0) ? numPoints + 1 : numPoints;
  return 3*((n - 1)/2);
}

extern "C" int sphb200_table_kernel_build(int kind, int ndim, size_t numPoints, double minNperh, double maxNperh,
                                          double* kextOut, double* xstepOut, size_t* n1Out,
                                          double* Wcoef, double* gradWcoef, double* grad2Wcoef,
                                          double* nperhVals, double* nperhRange, double* wsumVals, double* wsumRange) {
  const bool nbspline = kind >= SPHB200_KERNEL_NBSPLINE + 1 && kind <= SPHB200_KERNEL_NBSPLINE + 11;      // orders 1 .. 11 (kernel extent >= 1)
  if (((kind < 0 || kind > SPHB200_KERNEL_WENDLANDC2) && !nbspline) || ndim < 1 || ndim > 3 || numPoints < 3) return 1;
  const Analytic K(kind, ndim);
  const double kext = K.extent();
  const size_t n = (numPoints % 2 == 0) ? numPoints + 1 : numPoints;
  const double step = kext/double(n - 1);
  std::vector<double> y0(n), y1(n), y2(n);
  for (size_t i = 0; i < n; ++i) K.eval(double(i)*step, y0[i], y1[i], y2[i]);
  Quadratic qW, qG, qG2;
  qW.fit(0.0, kext, y0); qG.fit(0.0, kext, y1); qG2.fit(0.0, kext, y2);
  for (size_t k = 0; k < qW.c.size(); ++k) { Wcoef[k] = qW.c[k]; gradWcoef[k] = qG.c[k]; if (grad2Wcoef) grad2Wcoef[k] = qG2.c[k]; }
  if (kextOut) *kextOut = kext;
  if (xstepOut) *xstepOut = qW.xstep;
  if (n1Out) *n1Out = qW.n1;
  if (nperhVals && wsumVals) {
    double mn = std::max(minNperh, 1.1/kext), mx = maxNperh;
    if (mx <= mn) mx = 4.0*mn;
    auto sumKernel = [&](double nPerh) {                       // TableKernel.cc:24-66, kernelValueSPH = |gradW table|
      const double deta = 1.0/nPerh;
      double res = 0.0, etar = deta;
      while (etar < kext) {
        const double kv = std::fabs(qG(etar));
        res += ndim == 1 ? 2.0*kv : (ndim == 2 ? 2.0*M_PI*etar/deta*kv : 4.0*M_PI*(etar/deta)*(etar/deta)*kv);
        etar += deta;
      }
      return ndim == 1 ? res : (ndim == 2 ? std::sqrt(res) : std::pow(res, 1.0/3.0));
    };
    Hermite wsum, nperh;
    wsum.init(mn, mx, numPoints, sumKernel);
    const double w0 = wsum(mn), w1 = wsum(mx);
    nperh.init(w0, w1, numPoints, [&](double Wsum) { return bisect([&](double x) { return wsum(x) - Wsum; }, mn, mx); });
    wsum.makeMonotonic(); nperh.makeMonotonic();
    for (size_t k = 0; k < 2*numPoints; ++k) { wsumVals[k] = wsum.v[k]; nperhVals[k] = nperh.v[k]; }
    wsumRange[0] = mn; wsumRange[1] = mx; nperhRange[0] = w0; nperhRange[1] = w1;
  }
  return 0;
}
