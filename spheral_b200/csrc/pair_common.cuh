// pair_common.cuh -- device helpers shared by the pair-loop kernels (derivs.cu: SPH, crk.cu: CRKSPH):
// fast reciprocal / rsqrt, the shared-memory TableKernel lookup, small tensor algebra, the Balsara switch,
// the ASPH smoothing-scale derivative and the CubicHermite lookup.
#pragma once
#include "sphb200_internal.cuh"

// Table record n1+1 is all zeros (staged by the kernels).  With SPHB200_TABLE_ZERO=1 the lookup clamps eta >= kext onto
// that record instead of selecting 0 afterwards.  Default 0: the variant every committed measurement was taken with.
#ifndef SPHB200_TABLE_ZERO
#define SPHB200_TABLE_ZERO 0
#endif

namespace {

// Fast FP64 reciprocal / reciprocal square root for the pair loop: the hardware seed (MUFU.RCP64H / MUFU.RSQ64H) refined to
// full double precision (relative error ~1e-16), without the IEEE corner-case slow path of `/` and sqrt() -- arguments in
// the pair loop are finite, positive and far from the denormal range.  The 1e-10 parity bar leaves 6 digits of head room.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));       // relative error e0 <= 2^-23
  double e = fma(-x, r, 1.0);
  e = fma(e, e, e);                                            // r*(1 + e + e^2): error e0^3 ~ 2^-69
  return fma(r, e, r);
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));     // relative error e0 <= 2^-22
  const double t = y*y;
  const double e = fma(-t, x, 1.0);
  const double p = fma(e, 0.375, 0.5);
  const double q = e*y;
  return fma(p, q, y);                                         // y*(1 + e/2 + 3e^2/8): error O(e0^3) ~ 2^-64
}

__device__ __forceinline__ double2 lds128(unsigned addr) {      // 32-bit shared-window address: no generic->shared conversion per use
  double2 v;
  asm("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

// Reads of shared memory that cp.async fills (the neighbour-row ring of k_sph_derivs): volatile, so that the compiler keeps them
// behind the cp.async.wait_group / __syncwarp that publish the data.  The plain lds128 above is a pure function of its address to
// the compiler and may be hoisted out of a loop -- fine for the read-only table, wrong for a ring slot (seen when a look-ahead
// read of the next slot was tried: it was moved above the copies that fill it).
__device__ __forceinline__ double2 lds128v(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}

// TableKernelView::kernelAndGradValue (Kernel/TableKernelViewInline.hh:84-99) with
// QuadraticInterpolatorView::lowerBound (Utilities/QuadraticInterpolatorViewInline.hh:71-77), WITHOUT the Hdet factor
// (the caller multiplies; the raw gradient value is also kernelValueSPH of TableKernelViewInline.hh:118-127).
// The interval index is size_t(max(0,x-xmin)/xstep); a reciprocal multiply is used unless the quotient is within 1e-9
// of an integer, where the true division decides (an off-by-one interval would change W at the 1e-6 level).
__device__ __forceinline__ void table_eval_raw(unsigned tab, double kext, double xmin, double xstep, double rxstep,
                                               uint32_t n1, double eta, double& W, double& gW) {
  const double x = eta - xmin;                  // max(0, .) of the reference is applied to the integer index below
  const double q = x*rxstep;
  int k = __double2int_rz(q);
  const double fr = q - (double)k;
#if SPHB200_TABLE_ZERO
  k = max(min(k, (int)n1 + 1), 0);              // record n1+1 is all zeros: eta >= kext gives W = gradW = 0 without a select
  if (fr < 1.0e-9 || fr > 1.0 - 1.0e-9)         // within 1e-9 of an interval edge (or of kext): the reference's own expressions decide
    k = (eta < kext) ? min((int)(fmax(x, 0.0)/xstep), (int)n1) : (int)n1 + 1;
  const unsigned c = tab + 48u*(unsigned)k;
  const double2 c01 = lds128(c), c23 = lds128(c + 16u), c45 = lds128(c + 32u);
  W  = fma(fma(c23.x, eta, c01.y), eta, c01.x);
  gW = fma(fma(c45.y, eta, c45.x), eta, c23.y);
#else
  if (fr < 1.0e-9 || fr > 1.0 - 1.0e-9) k = (int)(fmax(x, 0.0)/xstep);
  k = max(min(k, (int)n1), 0);
  const unsigned c = tab + 48u*(unsigned)k;
  const double2 c01 = lds128(c), c23 = lds128(c + 16u), c45 = lds128(c + 32u);
  const bool in = eta < kext;
  W  = in ? fma(fma(c23.x, eta, c01.y), eta, c01.x) : 0.0;
  gW = in ? fma(fma(c45.y, eta, c45.x), eta, c23.y) : 0.0;
#endif
}

// The gradient half of table_eval_raw (same interval selection, same coefficients): for the side of a pair whose W is not needed.
__device__ __forceinline__ void table_eval_grad(unsigned tab, double kext, double xmin, double xstep, double rxstep,
                                                uint32_t n1, double eta, double& gW) {
  const double x = eta - xmin;
  const double q = x*rxstep;
  int k = __double2int_rz(q);
  const double fr = q - (double)k;
  if (fr < 1.0e-9 || fr > 1.0 - 1.0e-9) k = (int)(fmax(x, 0.0)/xstep);
  k = max(min(k, (int)n1), 0);
  const unsigned c = tab + 48u*(unsigned)k;
  const double2 c23 = lds128(c + 16u), c45 = lds128(c + 32u);
  gW = (eta < kext) ? fma(fma(c45.y, eta, c45.x), eta, c23.y) : 0.0;
}

template <int DIM> __device__ __forceinline__ double rootnu(double x) {
  if (DIM == 3) return d_sgn(x)*pow(fabs(x), 0.3333333333333333);   // Dimension.hh:94, FastMath.hh:152-163
  return sqrt(x);
}
template <int DIM> __device__ __forceinline__ double ten_det(const double* T) {
  if (DIM == 3) return (T[0]*T[4]*T[8] + T[1]*T[5]*T[6] + T[2]*T[3]*T[7] - T[0]*T[5]*T[7] - T[1]*T[3]*T[8] - T[2]*T[4]*T[6]);
  return T[0]*T[3] - T[1]*T[2];
}
template <int DIM> __device__ __forceinline__ void ten_inverse(const double* T, double* o) {
  const double di = 1.0/ten_det<DIM>(T);
  if (DIM == 3) {
    const double xx = T[0], xy = T[1], xz = T[2], yx = T[3], yy = T[4], yz = T[5], zx = T[6], zy = T[7], zz = T[8];
    o[0] = (yy*zz - yz*zy)*di; o[1] = (xz*zy - xy*zz)*di; o[2] = (xy*yz - xz*yy)*di;
    o[3] = (yz*zx - yx*zz)*di; o[4] = (xx*zz - xz*zx)*di; o[5] = (xz*yx - xx*yz)*di;
    o[6] = (yx*zy - yy*zx)*di; o[7] = (xy*zx - xx*zy)*di; o[8] = (xx*yy - xy*yx)*di;
  } else {
    const double xx = T[0], xy = T[1], yx = T[2], yy = T[3];
    o[0] = yy*di; o[1] = -xy*di; o[2] = -yx*di; o[3] = xx*di;
  }
}
template <int DIM> __device__ __forceinline__ void ten_mul(const double* A, const double* B, double* o) {
  double t[DIM*DIM];
#pragma unroll
  for (int r = 0; r < DIM; ++r)
#pragma unroll
    for (int c = 0; c < DIM; ++c) {
      double s = A[r*DIM]*B[c];
#pragma unroll
      for (int k = 1; k < DIM; ++k) s += A[r*DIM + k]*B[k*DIM + c];
      t[r*DIM + c] = s;
    }
#pragma unroll
  for (int k = 0; k < DIM*DIM; ++k) o[k] = t[k];
}
template <int DIM> __device__ __forceinline__ double ten_trace(const double* T) { return DIM == 3 ? T[0] + T[4] + T[8] : T[0] + T[3]; }
template <int DIM> __device__ __forceinline__ void ten_dot(const double* T, const double* v, double* o) {
#pragma unroll
  for (int r = 0; r < DIM; ++r) {
    double s = T[r*DIM]*v[0];
#pragma unroll
    for (int k = 1; k < DIM; ++k) s += T[r*DIM + k]*v[k];
    o[r] = s;
  }
}

// ArtificialViscosityHandle::calcBalsaraShearCorrection (ArtificialViscosityHandleInline.hh:47-63)
template <int DIM> __device__ __forceinline__ double balsara(const sphb200_options& o, const double* DvDx, double Hdet, double cs) {
  const double div = fabs(ten_trace<DIM>(DvDx));
  double curl;
  if (DIM == 3) { const double a = DvDx[7] - DvDx[5], b = DvDx[2] - DvDx[6], c = DvDx[3] - DvDx[1]; curl = sqrt(a*a + b*b + c*c); }
  else curl = fabs(DvDx[2] - DvDx[1]);
  const double hmaxinverse = rootnu<DIM>(Hdet);
  const double x = div + curl + o.eps2*fmax(o.negligibleSoundSpeed, cs)*hmaxinverse;
  return div*(d_sgn(x)/fmax(1.0e-30, fabs(x)));
}

// smoothingScaleDerivative (SmoothingScale/SmoothingScaleUtilities.hh:46-85)
template <int DIM> __device__ __forceinline__ void asph_DHDt(const double* H, const double* T, double* o) {
  if (DIM == 3) {
    const double Hxx = H[0], Hxy = H[1], Hxz = H[2], Hyy = H[3], Hyz = H[4], Hzz = H[5];
    const double Txx = T[0], Txy = T[1], Txz = T[2], Tyx = T[3], Tyy = T[4], Tyz = T[5], Tzx = T[6], Tzy = T[7], Tzz = T[8];
    const double AA = Hxx*Txy - Hxy*(Txx - Tyy) + Hxz*Tzy - Hyy*Tyx - Hyz*Tzx;
    const double BB = Hxx*Txz + Hxy*Tyz - Hxz*(Txx - Tzz) - Hyz*Tyx - Hzz*Tzx;
    const double CC = Hxy*Txz + Hyy*Tyz - Hyz*(Tyy - Tzz) - Hxz*Txy - Hzz*Tzy;
    const double thpt = Hyy + Hzz;
    const double Ga = (Hxx + Hyy)*thpt - Hxz*Hxz;
    const double Gb = (Hyy + Hzz)*Hyz + Hxy*Hxz;
    const double Gc = (Hxx + Hzz)*thpt - Hxy*Hxy;
    const double Gd = thpt*AA + Hxz*CC;
    const double Ge = thpt*BB - Hxy*CC;
    const double ack = 1.0/(Ga*Gc - Gb*Gb);
    const double Gdot = (Gc*Gd - Gb*Ge)*ack;
    const double Tdot = (Gb*Gd - Ga*Ge)*ack;
    const double Phidot = (Hxz*Gdot + Hxy*Tdot + CC)/thpt;
    o[0] = -Hxx*Txx + Hxy*(Gdot - Tyx) - Hxz*(Tdot + Tzx);
    o[1] = Hyy*Gdot - Hyz*Tdot - Hxx*Txy - Hxy*Tyy - Hxz*Tzy;
    o[2] = Hyz*Gdot - Hzz*Tdot - Hxx*Txz - Hxy*Tyz - Hxz*Tzz;
    o[3] = Hyz*(Phidot - Tzy) - Hxy*(Gdot + Txy) - Hyy*Tyy;
    o[4] = Hxy*Tdot - Hyy*Phidot - Hxz*Txy - Hyz*Tyy - Hzz*Tzy;
    o[5] = Hxz*(Tdot - Txz) - Hyz*(Phidot + Tyz) - Hzz*Tzz;
  } else {
    const double Hxx = H[0], Hyx = H[1], Hyy = H[2];
    const double Txx = T[0], Txy = T[1], Tyx = T[2], Tyy = T[3];
    const double thetaDot = (Hxx*Txy - Hyy*Tyx - Hyx*(Txx - Tyy))/(Hxx + Hyy);
    o[0] = Hyx*(thetaDot - Tyx) - Hxx*Txx;
    o[1] = -(Hxx*thetaDot + Hyx*Txx + Hyy*Tyx);
    o[2] = -Hyx*(thetaDot + Txy) - Hyy*Tyy;
  }
}

// CubicHermiteInterpolatorView::operator() (Utilities/CubicHermiteInterpolatorViewInline.hh:8-33,103-108)
__device__ __forceinline__ double hermite_eval(const double* __restrict__ v, uint32_t n, double xmin, double xmax, double xstep, double x) {
  if (x < xmin) return v[0] + v[n]*(x - xmin);
  if (x > xmax) return v[n - 1u] + v[2u*n - 1u]*(x - xmin);
  uint32_t i0 = (uint32_t)(fmax(0.0, x - xmin)/xstep);
  i0 = min(i0, n - 2u);
  const double t = fmax(0.0, fmin(1.0, (x - xmin - (double)i0*xstep)/xstep));
  const double t2 = t*t, t3 = t*t2;
  return ((2.0*t3 - 3.0*t2 + 1.0)*v[i0] + (-2.0*t3 + 3.0*t2)*v[i0 + 1u] +
          xstep*((t3 - 2.0*t2 + t)*v[n + i0] + (t3 - t2)*v[n + i0 + 1u]));
}

}  // namespace
