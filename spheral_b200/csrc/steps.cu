// steps.cu -- the per-step callers either side of the derivative path, device-resident:
//
//   k_sph_sum_density   computeSPHSumMassDensity            SPH/computeSPHSumMassDensity.cc:15-89   (SPHBase::preStepInitialize)
//   k_sph_omega         computeSPHOmegaGradhCorrection      SPH/computeSPHOmegaGradhCorrection.cc:20-113 (SPHBase::postStateUpdate)
//   k_eos_gamma         PressurePolicy / SoundSpeedPolicy with GammaLawGas (Material/GammaLawGas.cc:185-189, 233-238)
//   k_state_update      State::update with the policies the hydro and smoothing-scale packages register
//                       (IncrementState, IncrementBoundedState, ReplaceBoundedState, IncrementASPHHtensor; DataBase/State.cc:221-300)
//   k_dt_nodes/_pairs   GenericHydro::dt                    Physics/GenericHydro.cc:112-381
//
// so that no field leaves the GPU between the stages of CheapSynchronousRK2 (Integrator/CheapSynchronousRK2.cc:40-132): the
// host integrator (spheral_b200/integrator.py) issues these calls and reads back one number per step, the time step.
// The pair loops use the same i-centric gather and the same warp-cooperative record ring as the derivative kernels.
#include "sphb200_internal.cuh"
#include "pair_common.cuh"
#include "nbr_ring.cuh"
#include "sym_eigen.cuh"
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace {

constexpr int RB = 256;
// Warps per CTA x ring depth, measured at 8 M (ms: dt vote / grad-h correction / sum density): 8 x 4: 5.07 / 8.14 / 11.63,
// 16 x 2: 3.86 / 7.37 / 9.77, 16 x 3: 4.03 / 7.45 / 10.44, 24 x 2: 3.77 / 7.32 / 10.72 -- these loops have little arithmetic per
// edge, so resident warps hide the gather better than a deeper ring, and a larger CTA shares one staged kernel table.
#ifndef SPHB200_STEP_WARPS
#define SPHB200_STEP_WARPS 16
#endif
#ifndef SPHB200_STEP_STAGES
#define SPHB200_STEP_STAGES 2
#endif
constexpr int SW = SPHB200_STEP_WARPS, SS = SPHB200_STEP_STAGES;     // warps per CTA and ring depth of the light pair loops

template <int DIM> struct StepPrefix { static constexpr int MASS = ((Dm<DIM>::R_M + 1)*8 + 15)/16*16; };    // position .. mass
template <int DIM> using MassRing = NbrRing<DIM, 0, 0, SS, StepPrefix<DIM>::MASS, false>;
template <int DIM> using PosRing = NbrRing<DIM, 0, 0, SS, (DIM == 3 ? 32 : 16), false>;
// dt pair loop: its own compact 32-byte record {v (DIM), nodeScale} per node, built by k_node_scale; no node rows at all
struct DtRing {
  static constexpr int REC = 32, RB = ring_pad(REC), STAGEB = 32*RB, WARPB = SS*STAGEB;
  unsigned base; const unsigned char* recs;
  mutable unsigned rot = 0u;                         // see NbrRing::rot
  __device__ __forceinline__ unsigned stage(uint32_t p) const { return base + ((p + rot) % SS)*(unsigned)STAGEB; }
  __device__ __forceinline__ void issue(uint32_t p, uint32_t jrow, int lane) const {
    ring_copy_records<REC, RB, REC>(stage(p), recs, jrow, lane);
    ring_commit();
  }
  __device__ __forceinline__ void read(uint32_t k, int lane, double* o) const {
    const double2 a = ring_lds128(stage(k) + (unsigned)lane*RB), b = ring_lds128(stage(k) + (unsigned)lane*RB + 16u);
    o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
  }
};

struct LoopArgs {
  const double* rows; const double* aux2; const uint32_t* perm;
  const uint32_t* nbrCount; const uint32_t* tileRows; const unsigned long long* tileOff; const uint32_t* nbr;
  const double* tabW; double kext, xmin, xstep; uint32_t n1; double W0;
  size_t n; uint32_t nInt;
  double* out;                                       // api-order destination (rho / omega)
  // classic ASPH ideal H: sorted component-major derivative arrays (component c of slot i at [c*cap + i]) and the nperh look-up
  double *dM0, *dM1, *dHideal; size_t cap;
  const double* nperhVals; uint32_t nperhN; double nperhXmin, nperhXmax, nperhXstep;
  double nPerh, hminInv, hmaxInv, hminratio;
  // dt
  unsigned long long* best;                          // {dt bits, tag} pairs per CTA
};

struct TileLane { size_t i; bool inRange, active; uint32_t o, cnt, rowsT; unsigned long long base; };
__device__ __forceinline__ TileLane tile_lane(const LoopArgs& a, size_t tile, int lane) {
  TileLane t;
  t.i = tile*SPHB200_TILE + lane;
  t.inRange = t.i < a.n;
  t.o = t.inRange ? a.perm[t.i] : 0xffffffffu;
  t.active = t.inRange && t.o < a.nInt;
  t.cnt = t.active ? a.nbrCount[t.i] : 0u;
  t.rowsT = a.tileRows[tile];
  t.base = a.tileOff[tile] + lane;
  return t;
}
__device__ __forceinline__ unsigned stage_table(double* smem, const double* __restrict__ tab, uint32_t n1) {
  const uint32_t nW = 6u*(n1 + 2u);
  for (uint32_t k = threadIdx.x; k < nW; k += blockDim.x) smem[k] = (k < nW - 6u) ? tab[k] : 0.0;
  __syncthreads();
  return (unsigned)__cvta_generic_to_shared(smem);
}
template <typename Ring> __device__ __forceinline__ Ring make_ring(const LoopArgs& a, unsigned ringBase, int warp) {
  Ring r;
  r.base = ringBase + (unsigned)warp*(unsigned)Ring::WARPB;
  r.rows = reinterpret_cast<const unsigned char*>(a.rows);
  r.x1 = nullptr; r.x2 = nullptr;
  r.aux2 = reinterpret_cast<const unsigned char*>(a.aux2);
  return r;
}

// ---- computeSPHSumMassDensity: rho_i = m_i W(0) Hdet_i + sum_j m_j Hdet_j W(|Hj (ri - rj)|)  (scatter form: the neighbour's H) ----
template <int DIM>
__global__ void __launch_bounds__(32*SW, 1) k_sph_sum_density(LoopArgs a) {
  using D = Dm<DIM>;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const MassRing<DIM> ring = make_ring<MassRing<DIM>>(a, tW + 48u*(a.n1 + 2u), warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*SW + warp; tile < nTiles; tile += (size_t)gridDim.x*SW) {
    const TileLane t = tile_lane(a, tile, lane);
    double ri[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) ri[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_POS + k] : 0.0;
    double sum = 0.0;
    ring_walk<MassRing<DIM>, SS>(ring, lane, t.rowsT, t.cnt,
      [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t) {
        double rw[StepPrefix<DIM>::MASS/8];
        ring.read_row(k, lane, rw);
        const double Hdetj = sym_det<DIM>(rw + D::R_H);        // recomputed: cheaper than a per-lane 16-byte copy (nbr_ring.cuh)
        double rij[DIM], eta[DIM];
#pragma unroll
        for (int q = 0; q < DIM; ++q) rij[q] = ri[q] - rw[D::R_POS + q];
        sym_dot<DIM>(rw + D::R_H, rij, eta);
        const double e2 = vdot<DIM>(eta, eta);
        double W, gW;
        table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, e2*fast_rsqrt(e2 + 1.0e-300), W, gW);
        sum = fma(rw[D::R_M], W*Hdetj, sum);                 // :78  massDensity_i += mj*Wj
      });
    if (t.active) a.out[t.o] = sum + a.rows[t.i*D::ROW + D::R_M]*(a.W0*a.aux2[2*t.i]);      // :37-46 self contribution
  }
}

// ---- computeSPHOmegaGradhCorrection: omega_i = max(1e-30, -sum_j eta_i gW_i / (nDim (sum_j W_i + Hdet_i W0))) -------------------
template <int DIM>
__global__ void __launch_bounds__(32*SW, 1) k_sph_omega(LoopArgs a) {
  using D = Dm<DIM>;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const PosRing<DIM> ring = make_ring<PosRing<DIM>>(a, tW + 48u*(a.n1 + 2u), warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*SW + warp; tile < nTiles; tile += (size_t)gridDim.x*SW) {
    const TileLane t = tile_lane(a, tile, lane);
    double ri[DIM], Hi[D::NS];
#pragma unroll
    for (int k = 0; k < DIM; ++k) ri[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_POS + k] : 0.0;
#pragma unroll
    for (int k = 0; k < D::NS; ++k) Hi[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_H + k] : 0.0;
    double wsum = 0.0, gsum = 0.0;
    ring_walk<PosRing<DIM>, SS>(ring, lane, t.rowsT, t.cnt,
      [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t) {
        double rj[DIM == 3 ? 4 : 2];
        ring.read_row(k, lane, rj);
        double rij[DIM], eta[DIM];
#pragma unroll
        for (int q = 0; q < DIM; ++q) rij[q] = ri[q] - rj[q];
        sym_dot<DIM>(Hi, rij, eta);
        const double e2 = vdot<DIM>(eta, eta);
        const double etaMag = e2*fast_rsqrt(e2 + 1.0e-300);
        double W, gW;
        table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, etaMag, W, gW);
        wsum += W; gsum = fma(etaMag, gW, gsum);             // :85-89, Hdet_i factored out of both sums
      });
    if (t.active) {
      double om = 1.0;                                       // :99-100 isolated point
      if (t.cnt != 0u) {
        const double Hdeti = sym_det<DIM>(Hi);
        const double o1 = wsum*Hdeti + Hdeti*a.W0;           // :103-106
        om = fmax(1.0e-30, -(gsum*Hdeti)/((double)DIM*o1));
      }
      a.out[t.o] = om;
    }
  }
}

// ---- ASPHClassicSmoothingScale::evaluateDerivatives (SmoothingScale/ASPHClassicSmoothingScale.cc:130-378), general case -------------------
// The pair part needs the node's own H only (W_SPH,i = |gradW(|H_i x_ij|)|), so the neighbours' positions are all that is streamed:
//   m0_i = sum W_SPH,i ; m1_i = -sum W_SPH,i eta_i ; psi_i = sum W_SPH,i^2 x_ij (x) x_ij / |x_ij|^5            (:206-214)
// and the per-node part turns the second moment into the unit-determinant shape of the ideal H (eigenvalues 1/sqrt(lambda), bounded
// below by hminratio times the largest, :289-312), scales it to the target neighbour count like the SPH ideal H (:339-367) and bounds
// the eigenvalues (:371-375).  Every tensor function is of the form R f(lambda) R^T, so one eigen-decomposition of psi carries the
// chain the reference evaluates with three (eigenVectors, sqrt, eigenVectors): the rotations coincide.  DHDt is the ASPH tensor
// derivative k_sph_derivs has already stored.  Not exported: the second moment itself (the reference enrolls it as a derivative field
// only for its own use and restart).
template <int DIM>
__global__ void __launch_bounds__(32*SW, 1) k_asph_classic(LoopArgs a) {
  using D = Dm<DIM>;
  constexpr int NS = D::NS;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const PosRing<DIM> ring = make_ring<PosRing<DIM>>(a, tW + 48u*(a.n1 + 2u), warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  const double tiny = 1.0e-50;
  for (size_t tile = (size_t)blockIdx.x*SW + warp; tile < nTiles; tile += (size_t)gridDim.x*SW) {
    const TileLane t = tile_lane(a, tile, lane);
    double ri[DIM], Hi[NS];
#pragma unroll
    for (int k = 0; k < DIM; ++k) ri[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_POS + k] : 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) Hi[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_H + k] : 0.0;
    double m0 = 0.0, m1[DIM], psi[NS];
#pragma unroll
    for (int k = 0; k < DIM; ++k) m1[k] = 0.0;
#pragma unroll
    for (int k = 0; k < NS; ++k) psi[k] = 0.0;
    ring_walk<PosRing<DIM>, SS>(ring, lane, t.rowsT, t.cnt,
      [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t) {
        double rj[DIM == 3 ? 4 : 2];
        ring.read_row(k, lane, rj);
        double rij[DIM], eta[DIM];
#pragma unroll
        for (int q = 0; q < DIM; ++q) rij[q] = ri[q] - rj[q];
        sym_dot<DIM>(Hi, rij, eta);
        const double e2 = vdot<DIM>(eta, eta);
        double gW;
        table_eval_grad(tW, a.kext, a.xmin, a.xstep, rx, a.n1, e2*fast_rsqrt(e2 + 1.0e-300), gW);
        const double W = fabs(gW);                             // kernelValueSPH (TableKernelViewInline.hh:118-127)
        m0 += W;
#pragma unroll
        for (int q = 0; q < DIM; ++q) m1[q] = fma(-W, eta[q], m1[q]);
        const double r2 = vdot<DIM>(rij, rij);
        // safeInvVar(|x_ij|^5): coincident nodes give 1e30 times an exactly zero dyad
        const double rinv = (r2 > 0.0) ? fast_rsqrt(r2) : 0.0;
        const double w5 = (W*W)*((rinv*rinv)*(rinv*rinv)*rinv);
        int c = 0;
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int q = r; q < DIM; ++q) { psi[c] = fma(w5, rij[r]*rij[q], psi[c]); ++c; }
      });
    if (!t.active) continue;
    const size_t cap = a.cap, i = t.i;
    const double z0 = rootnu<DIM>(fmax(0.0, m0));                                             // :269
    a.dM0[i] = z0;
#pragma unroll
    for (int q = 0; q < DIM; ++q) a.dM1[(size_t)q*cap + i] = m1[q];
    const bool isolated = fabs(z0) <= 1.0e-15*fmax(1.0, fabs(z0));                             // fuzzyEqual(z0, 0)
    const double cur = isolated ? 0.5*a.nPerh : fmax(0.0, hermite_eval(a.nperhVals, a.nperhN, a.nperhXmin, a.nperhXmax, a.nperhXstep, z0));
    const double sv = fmin(4.0, fmax(0.25, a.nPerh/cur));                                      // :279
    const double psiweight = fmax(0.0, fmin(1.0, 2.0/sv - 1.0));                               // :285
    double lam[DIM], V[DIM*DIM];
    bool shaped = false;
    if (psiweight > 0.0 && sym_det<DIM>(psi) > 0.0) {
      double mx = 0.0;
#pragma unroll
      for (int k = 0; k < NS; ++k) mx = fmax(mx, fabs(psi[k]));
      const double mi = 1.0/mx;
#pragma unroll
      for (int k = 0; k < NS; ++k) psi[k] *= mi;                                                // :291 (a positive scale: same eigenvectors, same sign tests)
      sym_eigen<DIM>(psi, lam, V);
      double lmin = lam[0];
#pragma unroll
      for (int k = 1; k < DIM; ++k) lmin = fmin(lmin, lam[k]);
      shaped = lmin > 0.0;                                                                      // :289
    }
    if (shaped) {
      const double dpsi = sym_det<DIM>(psi);
      if (dpsi > 1.0e-10) {                                                                     // :292-293
        const double f = 1.0/rootnu<DIM>(fabs(dpsi) + tiny);
#pragma unroll
        for (int k = 0; k < DIM; ++k) lam[k] *= f;
      } else {                                                                                  // :295 psi = one
#pragma unroll
        for (int k = 0; k < DIM; ++k) lam[k] = 1.0;
      }
      double vmax = 0.0, prod = 1.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) { lam[k] = 1.0/sqrt(lam[k]); vmax = fmax(vmax, lam[k]); }  // :301
      const double psimin = vmax*a.hminratio;
#pragma unroll
      for (int k = 0; k < DIM; ++k) { lam[k] = fmax(psimin, lam[k]); prod *= lam[k]; }          // :303 constructSymTensorWithMaxDiagonal
      const double g = 1.0/rootnu<DIM>(prod + tiny);                                            // :307
#pragma unroll
      for (int k = 0; k < DIM; ++k) lam[k] = 1.0/sqrt(fmax(0.0, lam[k]*g));                     // :312 psi.sqrt().Inverse()
    } else {
#pragma unroll
      for (int k = 0; k < DIM; ++k) lam[k] = 1.0;                                               // :334 SymTensor::one()
#pragma unroll
      for (int k = 0; k < DIM*DIM; ++k) V[k] = 0.0;
#pragma unroll
      for (int k = 0; k < DIM; ++k) V[k*DIM + k] = 1.0;
    }
    const double aa = (sv < 1.0 ? 0.4*(1.0 + sv*sv) : 0.4*(1.0 + 1.0/(sv*sv*sv + tiny)));       // :339-344
    const double f = rootnu<DIM>(sym_det<DIM>(Hi))/(1.0 - aa + aa*sv);                          // :367 general case
    double lmin = 1.0e300;
#pragma unroll
    for (int k = 0; k < DIM; ++k) { lam[k] *= f; lmin = fmin(lmin, lam[k]); }
    const double hminEffInv = fmin(a.hminInv, fmax(a.hmaxInv, lmin)/a.hminratio);               // :372
#pragma unroll
    for (int k = 0; k < DIM; ++k) lam[k] = fmax(a.hmaxInv, fmin(hminEffInv, lam[k]));           // :373 constructSymTensorWithBoundedDiagonal
    double Hid[NS];
    sym_rebuild<DIM>(lam, V, Hid);
#pragma unroll
    for (int k = 0; k < NS; ++k) a.dHideal[(size_t)k*cap + i] = Hid[k];
  }
}

// ---- GammaLawGas behind PressurePolicy / SoundSpeedPolicy: every node of the Field (ghosts included) ------------------------------
__global__ void __launch_bounds__(RB) k_eos_gamma(const double* __restrict__ rho, const double* __restrict__ eps, size_t n,
                                                  sphb200_gamma_law e, double* __restrict__ P, double* __restrict__ cs) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  const double g1 = e.gamma - 1.0;
  double p = g1*rho[i]*eps[i] - e.externalPressure;                                         // GammaLawGas.cc:188, EquationOfStateInline.hh:98-106
  p = (p < e.minimumPressure ? (e.minPressureType == 0 ? e.minimumPressure : 0.0) : (p > e.maximumPressure ? e.maximumPressure : p));
  P[i] = p;
  cs[i] = sqrt(fmax(0.0, e.gamma*g1*eps[i]));                                               // GammaLawGas.cc:237
}

// ---- State::update over the internal nodes ------------------------------------------------------------------------------------------
struct UpdateArgs {
  const uint32_t* permEval; size_t nEval, capEval; uint32_t nInt;
  const double* deriv[DV_COUNT];
  double *pos, *vel, *H, *rho, *eps;
  double* vol; const double* mass;                     // CRKSPH: the volume evolves with ContinuityVolumePolicy (null otherwise)
  double multiplier; int timeAdvanceOnly, epsDone;
  int hEvolution /*SPHB200_H_**/, HEvolution /*0 ideal 1 integrate 2 fixed*/;
  double rhoMin, rhoMax, hminInv, hmaxInv, hminratio;
};
template <int DIM>
__global__ void __launch_bounds__(RB) k_state_update(UpdateArgs a) {
  constexpr int NS = Dm<DIM>::NS;
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;        // sorted slot of the evaluation the derivatives came from
  if (s >= a.nEval) return;
  const size_t o = a.permEval[s];
  if (o >= a.nInt) return;
  const size_t cap = a.capEval;
  const double mult = a.multiplier;
  a.rho[o] = fmin(a.rhoMax, fmax(a.rhoMin, a.rho[o] + mult*a.deriv[DV_DRHODT][s]));        // IncrementBoundedStateInline.hh:69
  if (!a.epsDone) a.eps[o] += mult*a.deriv[DV_DEPSDT][s];                                  // IncrementStateInline.hh:69
#pragma unroll
  for (int q = 0; q < DIM; ++q) {
    a.pos[o*DIM + q] += mult*a.deriv[DV_DXDT][(size_t)q*cap + s];
    a.vel[o*DIM + q] += mult*a.deriv[DV_DVDT][(size_t)q*cap + s];
  }
  double Hi[NS];
#pragma unroll
  for (int q = 0; q < NS; ++q) Hi[q] = a.H[o*NS + q];
  const bool fixedH = (a.HEvolution == 2 || a.hEvolution == SPHB200_H_NONE);                // FixedH
  if (fixedH) {
  } else if (a.hEvolution == SPHB200_H_ASPH) {                                                    // IncrementASPHHtensor.cc:82-88
#pragma unroll
    for (int q = 0; q < NS; ++q) Hi[q] += mult*a.deriv[DV_DHDT][(size_t)q*cap + s];
    double lam[DIM], V[DIM*DIM];
    sym_eigen<DIM>(Hi, lam, V);
    double lo = lam[0];
#pragma unroll
    for (int k = 1; k < DIM; ++k) lo = fmin(lo, lam[k]);
    const double hminEffInv = fmin(a.hminInv, fmax(a.hmaxInv, lo)/a.hminratio);
#pragma unroll
    for (int k = 0; k < DIM; ++k) lam[k] = fmax(a.hmaxInv, fmin(hminEffInv, lam[k]));
    sym_rebuild<DIM>(lam, V, Hi);
  } else if (a.HEvolution == 1 || a.timeAdvanceOnly) {                                      // IntegrateH, or IdealH degraded to an increment
#pragma unroll
    for (int q = 0; q < NS; ++q) Hi[q] += mult*a.deriv[DV_DHDT][(size_t)q*cap + s];
    sym_bound<DIM>(Hi, a.hmaxInv, a.hminInv);
  } else {                                                                                  // IdealH: ReplaceBoundedState("new H")
#pragma unroll
    for (int q = 0; q < NS; ++q) Hi[q] = a.deriv[DV_HIDEAL][(size_t)q*cap + s];
    sym_bound<DIM>(Hi, a.hmaxInv, a.hminInv);
  }
  if (!fixedH) {
#pragma unroll
    for (int q = 0; q < NS; ++q) a.H[o*NS + q] = Hi[q];
  }
  if (a.vol) {
    // ContinuityVolumePolicy::update (RK/ContinuityVolumePolicy.cc:33-66; CRKSPHBase.cc:155 enrolls the volume with it).  It depends
    // on the mass and the mass density, so it fires after the density update above; in timeAdvanceOnly mode the same expression
    // runs (UpdatePolicyBase.hh:54-61).  safeInvVar: sgn(x)/max(1e-30, |x|) (Utilities/safeInv.hh:24-27).
    const double m = a.mass[o], rhoN = a.rho[o];
    const double rho2 = rhoN*rhoN;
    const double volMin = 0.5*m*(d_sgn(rhoN)/fmax(1.0e-30, fabs(rhoN)));
    const double volMax = ((DIM == 3) ? 4.0*M_PI/(3.0*sym_det<DIM>(Hi)) : M_PI/sym_det<DIM>(Hi));
    const double dVdt = -m*(d_sgn(rho2)/fmax(1.0e-30, fabs(rho2)))*a.deriv[DV_DRHODT][s];
    a.vol[o] = fmax(volMin, fmin(volMax, a.vol[o] + mult*dVdt));
  }
}

// ---- iterateIdealH (Utilities/iterateIdealH.cc:120-190): one sweep H <- "new H" over the nodes not yet converged -----------------------
// deltaH_i = max |phi - 1| over the eigenvalues phi of H1^(1/2) H^-1 H1^(1/2).  The SPH smoothing scale's ideal H is a multiple of the
// identity (SPHSmoothingScale.cc:262-268), H1 = h1 I, for which phi = h1/lambda(H): only the eigenvalue range of the current H is needed.
template <int DIM>
__global__ void __launch_bounds__(RB) k_iterate_h(const uint32_t* __restrict__ perm, size_t n, size_t cap, uint32_t nInt,
                                                  const double* __restrict__ Hideal, double* __restrict__ H, uint32_t* __restrict__ done,
                                                  double tolerance, unsigned long long* __restrict__ maxDeltaBits) {
  constexpr int NS = Dm<DIM>::NS;
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  double delta = 0.0;
  if (s < n) {
    const size_t o = perm[s];
    if (o < nInt && done[o] == 0u) {
      double Hi[NS], H1[NS];
#pragma unroll
      for (int q = 0; q < NS; ++q) { Hi[q] = H[o*NS + q]; H1[q] = Hideal[(size_t)q*cap + s]; }
      double lo, hi;
      const bool iso1 = (DIM == 3) ? (H1[1] == 0.0 && H1[2] == 0.0 && H1[4] == 0.0 && H1[3] == H1[0] && H1[5] == H1[0]) : (H1[1] == 0.0 && H1[2] == H1[0]);
      if (iso1) {
        sym_eigenvalue_range<DIM>(Hi, lo, hi);
        const double h1 = H1[0];
        delta = fmax(fabs(h1/hi - 1.0), fabs(h1/lo - 1.0));
      } else {
        // a tensor ideal H (the classic ASPH package): phi = eigenvalues of (H1^(1/2) H^-1 H1^(1/2)).Symmetric(), iterateIdealH.cc:188-192
        double lam[DIM], V[DIM*DIM], S[NS], Sf[DIM][DIM], Hf[DIM][DIM], Hv[DIM][DIM], T[DIM][DIM], P[DIM][DIM], Ps[NS];
        sym_eigen<DIM>(H1, lam, V);
#pragma unroll
        for (int k = 0; k < DIM; ++k) lam[k] = sqrt(fmax(0.0, lam[k]));
        sym_rebuild<DIM>(lam, V, S);
        if (DIM == 3) {
          Sf[0][0] = S[0]; Sf[0][1] = Sf[1][0] = S[1]; Sf[0][DIM - 1] = Sf[DIM - 1][0] = S[2]; Sf[1][1] = S[3]; Sf[1][DIM - 1] = Sf[DIM - 1][1] = S[4]; Sf[DIM - 1][DIM - 1] = S[5];
          Hf[0][0] = Hi[0]; Hf[0][1] = Hf[1][0] = Hi[1]; Hf[0][DIM - 1] = Hf[DIM - 1][0] = Hi[2]; Hf[1][1] = Hi[3]; Hf[1][DIM - 1] = Hf[DIM - 1][1] = Hi[4]; Hf[DIM - 1][DIM - 1] = Hi[5];
        } else {
          Sf[0][0] = S[0]; Sf[0][1] = Sf[1][0] = S[1]; Sf[1][1] = S[2];
          Hf[0][0] = Hi[0]; Hf[0][1] = Hf[1][0] = Hi[1]; Hf[1][1] = Hi[2];
        }
        { double Tin[DIM*DIM], Tout[DIM*DIM];
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int q = 0; q < DIM; ++q) Tin[r*DIM + q] = Hf[r][q];
          ten_inverse<DIM>(Tin, Tout);
#pragma unroll
          for (int r = 0; r < DIM; ++r)
#pragma unroll
            for (int q = 0; q < DIM; ++q) Hv[r][q] = Tout[r*DIM + q]; }
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int q = 0; q < DIM; ++q) { double t = 0.0;
#pragma unroll
            for (int m = 0; m < DIM; ++m) t += Sf[r][m]*Hv[m][q];
            T[r][q] = t; }
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int q = 0; q < DIM; ++q) { double t = 0.0;
#pragma unroll
            for (int m = 0; m < DIM; ++m) t += T[r][m]*Sf[m][q];
            P[r][q] = t; }
        if (DIM == 3) { Ps[0] = P[0][0]; Ps[1] = 0.5*(P[0][1] + P[1][0]); Ps[2] = 0.5*(P[0][DIM - 1] + P[DIM - 1][0]); Ps[3] = P[1][1]; Ps[4] = 0.5*(P[1][DIM - 1] + P[DIM - 1][1]); Ps[5] = P[DIM - 1][DIM - 1]; }
        else { Ps[0] = P[0][0]; Ps[1] = 0.5*(P[0][1] + P[1][0]); Ps[2] = P[1][1]; }
        sym_eigenvalue_range<DIM>(Ps, lo, hi);
        delta = fmax(fabs(lo - 1.0), fabs(hi - 1.0));
      }
      if (delta <= tolerance) done[o] = 1u;
#pragma unroll
      for (int q = 0; q < NS; ++q) H[o*NS + q] = H1[q];
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) delta = fmax(delta, __shfl_down_sync(0xffffffffu, delta, off));
  if ((threadIdx.x & 31) == 0 && delta > 0.0) atomicMax(maxDeltaBits, (unsigned long long)__double_as_longlong(delta));
}

// ---- GenericHydro::dt ------------------------------------------------------------------------------------------------------------------
// A candidate is {dt, tag}; tag = phase<<40 | node<<3 | reason orders equal dt values the way the reference meets them (nodes in
// index order with the checks in source order, then the pair loop), so the reported reason / node match the reference's.
__device__ __forceinline__ void dt_take(double& best, unsigned long long& tag, double v, unsigned long long t) {
  if (v < best || (v == best && t < tag)) { best = v; tag = t; }
}
__device__ __forceinline__ void dt_block_reduce(double best, unsigned long long tag, unsigned long long* out) {
  __shared__ double sb[32]; __shared__ unsigned long long st[32];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ob = __shfl_down_sync(0xffffffffu, best, off);
    const unsigned long long ot = __shfl_down_sync(0xffffffffu, tag, off);
    dt_take(best, tag, ob, ot);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { sb[warp] = best; st[warp] = tag; }
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    best = lane < nw ? sb[lane] : DBL_MAX; tag = lane < nw ? st[lane] : ~0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const double ob = __shfl_down_sync(0xffffffffu, best, off);
      const unsigned long long ot = __shfl_down_sync(0xffffffffu, tag, off);
      dt_take(best, tag, ob, ot);
    }
    if (lane == 0) { out[2*blockIdx.x] = (unsigned long long)__double_as_longlong(best); out[2*blockIdx.x + 1] = tag; }
  }
}
struct DtNodeArgs {
  const uint32_t* permEval; size_t nEval, capEval; uint32_t nInt;
  const double *vel, *H, *rho, *cs;                  // api order
  const double *maxQ, *DvDx, *DvDt;                  // sorted (evaluation order)
  double nPerh; int useVelMag;
  unsigned long long* best;
};
template <int DIM>
__global__ void __launch_bounds__(RB) k_dt_nodes(DtNodeArgs a) {
  constexpr int NS = Dm<DIM>::NS, NT = Dm<DIM>::NT;
  const double tiny = DBL_EPSILON;
  double best = DBL_MAX; unsigned long long tag = ~0ull;
  for (size_t s = (size_t)blockIdx.x*RB + threadIdx.x; s < a.nEval; s += (size_t)gridDim.x*RB) {
    const size_t o = a.permEval[s];
    if (o >= a.nInt) continue;
    double Hi[NS], v[DIM], acc[DIM];
#pragma unroll
    for (int q = 0; q < NS; ++q) Hi[q] = a.H[o*NS + q];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { v[q] = a.vel[o*DIM + q]; acc[q] = a.DvDt[(size_t)q*a.capEval + s]; }
    const double nodeScale = 1.0/sym_max_eigenvalue<DIM>(Hi)/a.nPerh;                               // :190
    const unsigned long long base = (unsigned long long)o << 3;
    dt_take(best, tag, nodeScale/(a.cs[o] + tiny), base | 0ull);                                       // :197 sound speed
    dt_take(best, tag, nodeScale/(sqrt(a.maxQ[s]/a.rho[o]) + tiny), base | 1ull);                      // :253 artificial viscosity
    double div = 0.0;
#pragma unroll
    for (int q = 0; q < DIM; ++q) div += a.DvDx[(size_t)(q*DIM + q)*a.capEval + s];
    dt_take(best, tag, 1.0/(fabs(div) + tiny), base | 2ull);                                           // :272 velocity divergence
    const double vmag = sqrt(vdot<DIM>(v, v)), amag = sqrt(vdot<DIM>(acc, acc));
    dt_take(best, tag, 0.1*fmax(nodeScale/(vmag + tiny), vmag/(amag + tiny)), base | 3ull);            // :288 total acceleration
    if (a.useVelMag) dt_take(best, tag, nodeScale/(vmag + 1.0e-10), base | 4ull);                      // :305 velocity magnitude
    (void)NT;
  }
  dt_block_reduce(best, tag, a.best);
}
// per sorted node: the dt record {v (DIM), [0,] nodeScale} (one closed-form eigenvalue per node instead of one per edge)
template <int DIM>
__global__ void __launch_bounds__(RB) k_node_scale(const double* __restrict__ rows, size_t n, double nPerh, double* __restrict__ out) {
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  double Hi[Dm<DIM>::NS];
#pragma unroll
  for (int q = 0; q < Dm<DIM>::NS; ++q) Hi[q] = rows[s*Dm<DIM>::ROW + Dm<DIM>::R_H + q];
  out[4*s] = rows[s*Dm<DIM>::ROW + Dm<DIM>::R_VEL]; out[4*s + 1] = rows[s*Dm<DIM>::ROW + Dm<DIM>::R_VEL + 1];
  out[4*s + 2] = (DIM == 3) ? rows[s*Dm<DIM>::ROW + Dm<DIM>::R_VEL + 2] : 0.0;
  out[4*s + 3] = 1.0/sym_max_eigenvalue<DIM>(Hi)/nPerh;
}
// :320-346 pairwise velocity difference limit: min over pairs of min(scale_i, scale_j)/|v_i - v_j|
template <int DIM>
__global__ void __launch_bounds__(32*SW, 1) k_dt_pairs(LoopArgs a) {
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DtRing ring;
  ring.base = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)warp*(unsigned)DtRing::WARPB;
  ring.recs = reinterpret_cast<const unsigned char*>(a.aux2);                                  // {v, nodeScale} records
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  const double tiny = DBL_EPSILON;
  double best = DBL_MAX; unsigned long long tag = ~0ull;
  for (size_t tile = (size_t)blockIdx.x*SW + warp; tile < nTiles; tile += (size_t)gridDim.x*SW) {
    const TileLane t = tile_lane(a, tile, lane);
    double ri[4] = {0.0, 0.0, 0.0, 0.0};
    if (t.inRange) {
#pragma unroll
      for (int q = 0; q < 4; ++q) ri[q] = a.aux2[4*t.i + q];
    }
    ring_walk<DtRing, SS>(ring, lane, t.rowsT, t.cnt,
      [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t j) {
        double rj[4];
        ring.read(k, lane, rj);
        double vij[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) vij[q] = ri[q] - rj[q];
        const double vm = sqrt(vij[0]*vij[0] + vij[1]*vij[1] + vij[2]*vij[2]);
        const double dtv = fmin(ri[3], rj[3])*(1.0/fmax(tiny, vm));                           // safeInvVar(|vij|, tiny)
        if (dtv <= best) {                                                                   // rare: resolve the pair's i_node only then
          const uint32_t oj = a.perm[j];
          const unsigned long long node = t.o < oj ? t.o : oj;
          dt_take(best, tag, dtv, (1ull << 40) | (node << 3) | 5ull);
        }
      });
  }
  dt_block_reduce(best, tag, a.best);
}
__global__ void __launch_bounds__(RB) k_dt_final(const unsigned long long* __restrict__ cand, int n, unsigned long long* __restrict__ out) {
  double best = DBL_MAX; unsigned long long tag = ~0ull;
  for (int k = threadIdx.x; k < n; k += RB) dt_take(best, tag, __longlong_as_double((long long)cand[2*k]), cand[2*k + 1]);
  dt_block_reduce(best, tag, out);
}

void fill_loop_args(sphb200_ctx* c, LoopArgs& a) {
  a = LoopArgs{};
  a.rows = c->rows; a.aux2 = c->aux2; a.perm = c->perm; a.nbrCount = c->nbrCount; a.tileRows = c->tileRows; a.tileOff = c->tileOff; a.nbr = c->nbr;
  a.tabW = c->W.coef; a.kext = c->W.kext; a.xmin = c->W.xmin; a.xstep = c->W.xstep; a.n1 = c->W.n1;
  a.n = c->n; a.nInt = (uint32_t)c->nInt;
}
double table_W0(const TableDev& t) { return t.hostW.empty() ? 0.0 : t.hostW[0]; }     // W(0) = a0 of the first interval

template <typename K> int launch_loop(sphb200_ctx* c, K kern, const LoopArgs& a, const char* name, size_t ringWarpBytes, bool withTable,
                                      unsigned* grid = nullptr) {
  const size_t shm = (withTable ? (size_t)6*(c->W.n1 + 2)*sizeof(double) : 0) + (size_t)SW*ringWarpBytes;
  if (shm > 227*1024) return sphb200_fail(c, "kernel table too large for shared memory");
  CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  int nsm = 148, perSM = 1;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, 32*SW, shm) != cudaSuccess || perSM < 1) perSM = 1;
  const unsigned nb = (unsigned)std::min<size_t>((c->nTiles + SW - 1)/SW, (size_t)nsm*perSM);
  kern<<<nb, 32*SW, shm, c->stream>>>(a);
  KERNEL_CHECK(c, name);
  if (grid) *grid = nb;
  return 0;
}
int loop_ready(sphb200_ctx* c, const char* who) {
  if (!c->pairsValid) return sphb200_fail(c, std::string(who) + ": connectivity is stale or missing (call build_pairs first)");
  if (!c->W.set) return sphb200_fail(c, std::string(who) + ": kernel table not set");
  if (!c->rowsValid && sphb200_pack_rows(c)) return 1;
  return 0;
}

}  // namespace

extern "C" {

}  // extern "C"
// Called by sphb200_evaluate_derivatives after the pair loop when opt.hEvolution == SPHB200_H_ASPH_CLASSIC
int sphb200_launch_asph_classic(sphb200_ctx* c) {
  if (c->n == 0) return 0;
  if (!c->W.nperhVals || c->W.nperhN < 2)
    return sphb200_fail(c, "evaluateDerivatives: ASPHClassicSmoothingScale needs the TableKernel nperh lookup (nperhVals) but none was set");
  LoopArgs a; fill_loop_args(c, a);
  a.dM0 = c->deriv[DV_M0]; a.dM1 = c->deriv[DV_M1]; a.dHideal = c->deriv[DV_HIDEAL]; a.cap = c->cap;
  a.nperhVals = c->W.nperhVals; a.nperhN = c->W.nperhN; a.nperhXmin = c->W.nperhXmin; a.nperhXmax = c->W.nperhXmax; a.nperhXstep = c->W.nperhXstep;
  a.nPerh = c->opt.nPerh; a.hminInv = 1.0/c->opt.hmin; a.hmaxInv = 1.0/c->opt.hmax; a.hminratio = c->opt.hminratio;
  if (c->ndim == 3) { if (launch_loop(c, k_asph_classic<3>, a, "k_asph_classic", PosRing<3>::WARPB, true)) return 1; }
  else              { if (launch_loop(c, k_asph_classic<2>, a, "k_asph_classic", PosRing<2>::WARPB, true)) return 1; }
  return 0;
}
extern "C" {

int sphb200_sum_mass_density(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->have[S_MASS]) return sphb200_fail(c, "sum_mass_density: the mass is not on the device");
  if (c->n == 0) return 0;
  if (loop_ready(c, "sum_mass_density")) return 1;
  LoopArgs a; fill_loop_args(c, a);
  a.W0 = table_W0(c->W); a.out = c->api[S_RHO];
  if (c->ndim == 3) { if (launch_loop(c, k_sph_sum_density<3>, a, "k_sph_sum_density", MassRing<3>::WARPB, true)) return 1; }
  else              { if (launch_loop(c, k_sph_sum_density<2>, a, "k_sph_sum_density", MassRing<2>::WARPB, true)) return 1; }
  c->have[S_RHO] = true;
  c->rowsValid = false;                              // the rows carry rho and P/rho^2
  return 0;
}

int sphb200_compute_omega_gradh(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->n == 0) return 0;
  if (loop_ready(c, "compute_omega_gradh")) return 1;
  LoopArgs a; fill_loop_args(c, a);
  a.W0 = table_W0(c->W); a.out = c->api[S_OMEGA];
  if (!c->have[S_OMEGA]) {                           // ghost entries default to 1 (SPHBase.cc:210 resizes the field with 1.0)
    std::vector<double> ones(c->n, 1.0);
    CU_CHECK(c, cudaMemcpyAsync(c->api[S_OMEGA], ones.data(), c->n*sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
  }
  if (c->ndim == 3) { if (launch_loop(c, k_sph_omega<3>, a, "k_sph_omega", PosRing<3>::WARPB, true)) return 1; }
  else              { if (launch_loop(c, k_sph_omega<2>, a, "k_sph_omega", PosRing<2>::WARPB, true)) return 1; }
  c->have[S_OMEGA] = true;
  c->rowsValid = false;                              // P/rho^2 in the rows carries 1/omega
  return 0;
}

int sphb200_update_eos_gamma_law(sphb200_ctx* c, const sphb200_gamma_law* e) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!e) return sphb200_fail(c, "update_eos_gamma_law: null equation of state");
  if (!(e->gamma > 1.0)) return sphb200_fail(c, "update_eos_gamma_law: gamma must exceed 1");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->have[S_RHO] || !c->have[S_EPS]) return sphb200_fail(c, "update_eos_gamma_law: mass density and specific thermal energy must be on the device");
  if (c->n == 0) return 0;
  k_eos_gamma<<<(unsigned)((c->n + RB - 1)/RB), RB, 0, c->stream>>>(c->api[S_RHO], c->api[S_EPS], c->n, *e, c->api[S_P], c->api[S_CS]);
  KERNEL_CHECK(c, "k_eos_gamma");
  c->have[S_P] = c->have[S_CS] = true;
  c->rowsValid = false;
  return 0;
}

int sphb200_state_update(sphb200_ctx* c, const sphb200_step_options* so, double multiplier, int timeAdvanceOnly) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!so) return sphb200_fail(c, "state_update: null options");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->derivNodeValid) return sphb200_fail(c, "state_update: no derivatives on the device (call evaluate_derivatives first)");
  if (c->nIntEval != c->nInt) return sphb200_fail(c, "state_update: the internal node count changed since the derivatives were evaluated");
  for (int s : {S_POS, S_VEL, S_H, S_RHO, S_EPS})
    if (!c->have[s]) return sphb200_fail(c, "state_update: position, velocity, H, mass density and specific thermal energy must be on the device");
  if (c->nInt == 0) return 0;
  int epsDone = 0;
  if (c->opt.compatibleEnergy && !timeAdvanceOnly) {
    // SpecificThermalEnergyPolicy fires before velocity and position (they declare it as a dependency, SPHBase.cc:237-245)
    if (sphb200_update_energy_compatible(c, multiplier)) return 1;
    epsDone = 1;
  }
  UpdateArgs a{};
  a.permEval = c->permEval; a.nEval = c->nEval; a.capEval = c->capEval; a.nInt = (uint32_t)c->nInt;
  for (int s = 0; s < DV_COUNT; ++s) a.deriv[s] = c->deriv[s];
  a.pos = c->api[S_POS]; a.vel = c->api[S_VEL]; a.H = c->api[S_H]; a.rho = c->api[S_RHO]; a.eps = c->api[S_EPS];
  a.multiplier = multiplier; a.timeAdvanceOnly = timeAdvanceOnly; a.epsDone = epsDone;
  if (c->opt.hydro == SPHB200_HYDRO_CRKSPH && c->have[S_VOLUME] && c->have[S_MASS]) { a.vol = c->api[S_VOLUME]; a.mass = c->api[S_MASS]; }
  a.hEvolution = c->opt.hEvolution; a.HEvolution = so->HEvolution;
  a.rhoMin = so->rhoMin; a.rhoMax = so->rhoMax; a.hminInv = 1.0/c->opt.hmin; a.hmaxInv = 1.0/c->opt.hmax; a.hminratio = so->hminratio;
  const unsigned nb = (unsigned)((c->nEval + RB - 1)/RB);
  if (c->ndim == 3) k_state_update<3><<<nb, RB, 0, c->stream>>>(a); else k_state_update<2><<<nb, RB, 0, c->stream>>>(a);
  KERNEL_CHECK(c, "k_state_update");
  // positions and H moved: the rows are stale.  The connectivity is deliberately kept -- the reference evaluates the mid-step
  // derivatives on the connectivity of the step start (CheapSynchronousRK2.cc:76-90) -- and the next build_pairs re-sorts anyway.
  c->rowsValid = false;
  return sphb200_update_eos_gamma_law(c, &so->eos);
}

int sphb200_state_copy(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  for (int s = 0; s < S_COUNT; ++s) {
    if (!c->have[s] || !c->api[s]) { c->have0[s] = false; continue; }
    const size_t bytes = c->cap*(size_t)sphb200_state_width(c->ndim, s)*sizeof(double);
    if (c->cap0[s] < c->cap) {
      if (c->api0[s]) cudaFree(c->api0[s]);
      c->api0[s] = nullptr; c->cap0[s] = 0;
      CU_CHECK(c, cudaMalloc((void**)&c->api0[s], bytes));
      c->cap0[s] = c->cap;
    }
    CU_CHECK(c, cudaMemcpyAsync(c->api0[s], c->api[s], c->n*(size_t)sphb200_state_width(c->ndim, s)*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    c->have0[s] = true;
  }
  c->n0 = c->n;
  return 0;
}

int sphb200_state_assign(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->n0 != c->n) return sphb200_fail(c, "state_assign: the node count changed since state_copy");
  for (int s = 0; s < S_COUNT; ++s) {
    if (!c->have0[s]) continue;
    CU_CHECK(c, cudaMemcpyAsync(c->api[s], c->api0[s], c->n*(size_t)sphb200_state_width(c->ndim, s)*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  }
  c->rowsValid = false;
  return 0;
}

int sphb200_iterate_ideal_h(sphb200_ctx* c, int firstSweep, double tolerance, double* maxDeltaH) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->opt.hEvolution != SPHB200_H_SPH && c->opt.hEvolution != SPHB200_H_ASPH_CLASSIC)
    return sphb200_fail(c, "iterate_ideal_h: the SPH and the classic ASPH smoothing scale compute an ideal H on the device (the Voronoi-cell ideal H of ASPHSmoothingScale is out of scope)");
  if (!c->derivsValid || !c->pairsValid) return sphb200_fail(c, "iterate_ideal_h: needs the derivatives ('new H') of the current connectivity");
  if (maxDeltaH) *maxDeltaH = 0.0;
  if (c->nInt == 0) return 0;
  // one flag per INTERNAL node (the only entries used); the number of internal nodes cannot change between the sweeps of one
  // relaxation, so the flags of earlier sweeps survive whatever the ghost set does to the node capacity
  if (!firstSweep && (!c->hDone || c->hDoneCap < c->nInt))
    return sphb200_fail(c, "iterate_ideal_h: the number of internal nodes grew since the first sweep (start again with firstSweep = 1)");
  if (firstSweep) {
    if (sphb200_ensure(c, c->hDone, c->hDoneCap, c->nInt)) return 1;
    CU_CHECK(c, cudaMemsetAsync(c->hDone, 0, c->nInt*sizeof(uint32_t), c->stream));
  }
  CU_CHECK(c, cudaMemsetAsync(c->counters + 9, 0, sizeof(unsigned long long), c->stream));
  const unsigned nb = (unsigned)((c->n + RB - 1)/RB);
  if (c->ndim == 3) k_iterate_h<3><<<nb, RB, 0, c->stream>>>(c->perm, c->n, c->cap, (uint32_t)c->nInt, c->deriv[DV_HIDEAL], c->api[S_H], c->hDone, tolerance, c->counters + 9);
  else              k_iterate_h<2><<<nb, RB, 0, c->stream>>>(c->perm, c->n, c->cap, (uint32_t)c->nInt, c->deriv[DV_HIDEAL], c->api[S_H], c->hDone, tolerance, c->counters + 9);
  KERNEL_CHECK(c, "k_iterate_h");
  unsigned long long bits = 0;
  CU_CHECK(c, cudaMemcpyAsync(&bits, c->counters + 9, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  double d; memcpy(&d, &bits, 8);
  if (maxDeltaH) *maxDeltaH = d;
  // H changed: the connectivity and everything derived from it are stale
  c->sortValid = c->rowsValid = c->pairsValid = c->derivsValid = false;
  return 0;
}

int sphb200_compute_dt(sphb200_ctx* c, double cfl, int useVelocityMagnitudeForDt, double* dt, int* reason, uint32_t* node) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->derivNodeValid)
    return sphb200_fail(c, "compute_dt: no derivatives on the device (call evaluate_derivatives first; a growth of the node arrays discards them: n = " +
                        std::to_string(c->n) + ", capacity " + std::to_string(c->cap) + ", at the evaluation " + std::to_string(c->nEval) + " / " + std::to_string(c->capEval) + ")");
  if (c->nIntEval != c->nInt) return sphb200_fail(c, "compute_dt: the internal node count changed since the derivatives were evaluated");
  for (int s : {S_VEL, S_H, S_RHO, S_CS})
    if (!c->have[s]) return sphb200_fail(c, "compute_dt: velocity, H, mass density and sound speed must be on the device");
  if (c->nInt == 0) { if (dt) *dt = DBL_MAX; if (reason) *reason = -1; if (node) *node = 0; return 0; }
  if (loop_ready(c, "compute_dt")) return 1;
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  const int nbNodes = (int)std::min<size_t>((c->nEval + RB - 1)/RB, (size_t)nsm*4);
  const size_t maxCand = (size_t)nbNodes + (size_t)nsm*8;
  if (sphb200_ensure(c, c->dtCand, c->dtCandCap, 2*maxCand + 2)) return 1;
  if (sphb200_ensure(c, c->dtAux, c->dtAuxCap, 4*c->cap)) return 1;
  DtNodeArgs na{};
  na.permEval = c->permEval; na.nEval = c->nEval; na.capEval = c->capEval; na.nInt = (uint32_t)c->nInt;
  na.vel = c->api[S_VEL]; na.H = c->api[S_H]; na.rho = c->api[S_RHO]; na.cs = c->api[S_CS];
  na.maxQ = c->deriv[DV_MAXQ]; na.DvDx = c->deriv[DV_DVDX]; na.DvDt = c->deriv[DV_DVDT];
  na.nPerh = c->opt.nPerh; na.useVelMag = useVelocityMagnitudeForDt; na.best = c->dtCand;
  if (c->ndim == 3) k_dt_nodes<3><<<nbNodes, RB, 0, c->stream>>>(na); else k_dt_nodes<2><<<nbNodes, RB, 0, c->stream>>>(na);
  KERNEL_CHECK(c, "k_dt_nodes");
  { const unsigned nb = (unsigned)((c->n + RB - 1)/RB);
    if (c->ndim == 3) k_node_scale<3><<<nb, RB, 0, c->stream>>>(c->rows, c->n, c->opt.nPerh, c->dtAux);
    else              k_node_scale<2><<<nb, RB, 0, c->stream>>>(c->rows, c->n, c->opt.nPerh, c->dtAux);
    KERNEL_CHECK(c, "k_node_scale"); }
  LoopArgs a; fill_loop_args(c, a);
  a.aux2 = c->dtAux; a.best = c->dtCand + 2*(size_t)nbNodes;
  unsigned nbPairs = 0;
  if (c->ndim == 3) { if (launch_loop(c, k_dt_pairs<3>, a, "k_dt_pairs", DtRing::WARPB, false, &nbPairs)) return 1; }
  else              { if (launch_loop(c, k_dt_pairs<2>, a, "k_dt_pairs", DtRing::WARPB, false, &nbPairs)) return 1; }
  if ((size_t)nbNodes + nbPairs > maxCand) return sphb200_fail(c, "compute_dt: internal candidate buffer too small");
  k_dt_final<<<1, RB, 0, c->stream>>>(c->dtCand, nbNodes + (int)nbPairs, c->dtCand + 2*maxCand);
  KERNEL_CHECK(c, "k_dt_final");
  unsigned long long res[2];
  CU_CHECK(c, cudaMemcpyAsync(res, c->dtCand + 2*maxCand, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  double best; memcpy(&best, &res[0], 8);
  if (dt) *dt = best*cfl;                                                                   // :378 scale by the cfl safety factor
  if (reason) *reason = (res[1] == ~0ull) ? -1 : (int)(res[1] & 7ull);
  if (node) *node = (res[1] == ~0ull) ? 0u : (uint32_t)((res[1] >> 3) & 0xffffffffull);
  return 0;
}

}  // extern "C"
