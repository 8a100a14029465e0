// sym_eigen.cuh -- eigenvalues / eigenvectors of the symmetric H tensor on the device (shared by steps.cu and boundary.cu).
#pragma once
#include "sphb200_internal.cuh"
#include <cfloat>

namespace {

// ---- symmetric eigen-decomposition (the role of GeomSymmetricTensor::eigenVectors, GeomSymmetricTensorInline.hh:2279-2364) -------
// 2-D: the reference's closed form.  3-D: the reference calls Eigen::SelfAdjointEigenSolver; every use here rebuilds
// R diag(f(lambda)) R^T, which is independent of eigenvector sign / order / degenerate-subspace basis, so cyclic Jacobi is used.
template <int DIM> __device__ void sym_eigen(const double* H, double* lam, double* V) {
  if (DIM == 2) {
    const double fscale = fmax(10.0*DBL_EPSILON, fmax(fabs(H[0]), fmax(fabs(H[1]), fabs(H[2]))));
    const double fi = 1.0/fscale;
    const double axx = H[0]*fi, axy = H[1]*fi, ayy = H[2]*fi;
    if (fabs(axy) < 1.0e-50) { lam[0] = H[0]; lam[1] = H[2]; V[0] = 1; V[1] = 0; V[2] = 0; V[3] = 1; return; }
    const double theta = 0.5*atan2(2.0*axy, ayy - axx);
    const double xh = cos(theta), yh = sin(theta);
    lam[0] = (xh*(axx*xh - axy*yh) - yh*(axy*xh - ayy*yh))*fscale;
    lam[1] = (yh*(axx*yh + axy*xh) + xh*(axy*yh + ayy*xh))*fscale;
    V[0] = xh; V[1] = yh; V[2] = -yh; V[3] = xh;
  } else {
    double A[3][3] = {{H[0], H[1], H[2]}, {H[1], H[3], H[4]}, {H[2], H[4], H[5]}};
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
      const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
      const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
      if (off <= 1.0e-300 || off <= 1.0e-18*diag) break;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = p + 1; q < 3; ++q) {
          if (A[p][q] == 0.0) continue;
          const double th = (A[q][q] - A[p][p])/(2.0*A[p][q]);
          const double tt = d_sgn(th)/(fabs(th) + sqrt(th*th + 1.0));
          const double cc = 1.0/sqrt(tt*tt + 1.0), ss = tt*cc;
#pragma unroll
          for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = cc*akp - ss*akq; A[k][q] = ss*akp + cc*akq; }
#pragma unroll
          for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = cc*apk - ss*aqk; A[q][k] = ss*apk + cc*aqk; }
#pragma unroll
          for (int k = 0; k < 3; ++k) { const double rkp = R[k][p], rkq = R[k][q]; R[k][p] = cc*rkp - ss*rkq; R[k][q] = ss*rkp + cc*rkq; }
        }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) { lam[k] = A[k][k];
#pragma unroll
      for (int l = 0; l < 3; ++l) V[3*k + l] = R[k][l]; }
  }
}
template <int DIM> __device__ void sym_rebuild(const double* lam, const double* V, double* H) {
  double F[DIM][DIM];
#pragma unroll
  for (int r = 0; r < DIM; ++r)
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) s += V[DIM*r + k]*lam[k]*V[DIM*c + k];
      F[r][c] = s;
    }
  if (DIM == 3) { H[0] = F[0][0]; H[1] = 0.5*(F[0][1] + F[1][0]); H[2] = 0.5*(F[0][2] + F[2][0]); H[3] = F[1][1]; H[4] = 0.5*(F[1][2] + F[2][1]); H[5] = F[2][2]; }
  else { H[0] = F[0][0]; H[1] = 0.5*(F[0][1] + F[1][0]); H[2] = F[1][1]; }
}
// min(maxv, max(minv, H)): enforceMinEigenValue then enforceMaxEigenValue (GeomSymmetricTensorInline.hh:2554-2640): H is returned
// untouched unless an eigenvalue is out of bounds
template <int DIM> __device__ void sym_bound(double* H, double minv, double maxv) {
  double lam[DIM], V[DIM*DIM];
  sym_eigen<DIM>(H, lam, V);
  double lo = lam[0], hi = lam[0];
#pragma unroll
  for (int k = 1; k < DIM; ++k) { lo = fmin(lo, lam[k]); hi = fmax(hi, lam[k]); }
  if (lo < minv) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) lam[k] = fmax(lam[k], minv);
    sym_rebuild<DIM>(lam, V, H);
    sym_eigen<DIM>(H, lam, V);
    hi = lam[0];
#pragma unroll
    for (int k = 1; k < DIM; ++k) hi = fmax(hi, lam[k]);
  }
  if (hi > maxv) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) lam[k] = fmin(lam[k], maxv);
    sym_rebuild<DIM>(lam, V, H);
  }
}
// GeomSymmetricTensor::eigenValues() (GeomSymmetricTensorInline.hh:2210-2256): the closed form GenericHydro::dt and
// findNodesTouchingThroughPlanes use; lo / hi = minElement() / maxElement()
template <int DIM> __device__ void sym_eigenvalue_range(const double* H, double& lo, double& hi) {
  if (DIM == 2) {
    if (fabs(H[1]) < 1.0e-50) { lo = fmin(H[0], H[2]); hi = fmax(H[0], H[2]); return; }
    const double b = H[0] + H[2], c = H[0]*H[2] - H[1]*H[1];
    const double q = 0.5*(b + d_sgn(b)*sqrt(fmax(0.0, b*b - 4.0*c)));
    lo = fmin(q, c/q); hi = fmax(q, c/q);
    return;
  }
  const double fscale = fmax(10.0*DBL_EPSILON, fmax(fmax(fabs(H[0]), fabs(H[1])), fmax(fmax(fabs(H[2]), fabs(H[3])), fmax(fabs(H[4]), fabs(H[5])))));
  const double fi = 1.0/fscale;
  const double a00 = H[0]*fi, a01 = H[1]*fi, a02 = H[2]*fi, a11 = H[3]*fi, a12 = H[4]*fi, a22 = H[5]*fi;
  const double c0 = a00*a11*a22 + 2.0*a01*a02*a12 - a00*a12*a12 - a11*a02*a02 - a22*a01*a01;
  const double c1 = a00*a11 - a01*a01 + a00*a22 - a02*a02 + a11*a22 - a12*a12;
  const double c2 = a00 + a11 + a22;
  const double third = 1.0/3.0;
  const double c2Div3 = c2*third;
  const double aDiv3 = fmin(0.0, third*(c1 - c2*c2Div3));
  const double mbDiv2 = 0.5*(c0 + c2Div3*(2.0*c2Div3*c2Div3 - c1));
  const double q = fmin(0.0, mbDiv2*mbDiv2 + aDiv3*aDiv3*aDiv3);
  const double mag = sqrt(-aDiv3);
  const double angle = atan2(sqrt(-q), mbDiv2)*third;
  const double cs = cos(angle), sn = sin(angle), s3 = sqrt(3.0);
  const double e0 = fscale*(c2Div3 + 2.0*mag*cs), e1 = fscale*(c2Div3 - mag*(cs + s3*sn)), e2 = fscale*(c2Div3 - mag*(cs - s3*sn));
  lo = fmin(e0, fmin(e1, e2)); hi = fmax(e0, fmax(e1, e2));
}
template <int DIM> __device__ double sym_max_eigenvalue(const double* H) { double lo, hi; sym_eigenvalue_range<DIM>(H, lo, hi); return hi; }
template <int DIM> __device__ double sym_min_eigenvalue(const double* H) { double lo, hi; sym_eigenvalue_range<DIM>(H, lo, hi); return lo; }


}  // namespace
