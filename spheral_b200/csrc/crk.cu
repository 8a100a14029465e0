// crk.cu -- the CRKSPH path (BASELINE config 4): RKOrder::LinearOrder reproducing kernels with
// RKVolumeType::RKSumVolume, the settings of tests/functional/Hydro/Sedov/Sedov-spherical-3d.py:60-61.
//
//   k_crk_volume        computeRKSumVolume                       RK/computeRKSumVolume.cc:33-116
//   k_crk_corrections   RKUtilities::computeCorrections          RK/RKUtilities.cc:252-491   (needHessian = false)
//   k_crk_sum_density   computeCRKSPHSumMassDensity              CRKSPH/computeCRKSPHSumMassDensity.cc:21-131
//   k_crk_derivs        CRKSPH<Dim>::evaluateDerivativesImpl     CRKSPH/CRKSPH.cc:176-440
//                       + SPHSmoothingScale / ASPHSmoothingScale::evaluateDerivatives (CRKSPHHydros.py:93-103)
//
// Same formulation as derivs.cu: one warp per tile of 32 Morton-consecutive nodes, lane <-> node i, every node gathers over
// its own complete neighbour list (the reference needs per-node lists for the corrections anyway, RKUtilities.cc:370-380),
// so nothing is scattered and a node's summation order is fixed by its list.  Each kernel writes the internal entries of
// its field; ghost entries come from the caller (boundary conditions / halo), as in the reference.
#include "sphb200_internal.cuh"
#include "pair_common.cuh"
#include "nbr_ring.cuh"
#include <algorithm>
#include <cmath>

#ifndef SPHB200_CRK_LIGHT_WARPS
#define SPHB200_CRK_LIGHT_WARPS 8
#endif
#ifndef SPHB200_CRK_LIGHT_STAGES
#define SPHB200_CRK_LIGHT_STAGES 4
#endif
#ifndef SPHB200_CRK_WARPS
#define SPHB200_CRK_WARPS 8
#endif
#ifndef SPHB200_CRK_STAGES
#define SPHB200_CRK_STAGES 2
#endif

namespace {

constexpr int RB = 256;
// Persistent CTAs, one per SM.  Light loops (volume, corrections, sum density): LW warps with a ring LS deep of node rows;
// the derivative loop: CW warps with a ring CS deep of {node row, RK corrections, {volume, Q velocity gradient}} (352 B per
// neighbour in 3-D, so the ring depth is what shared memory allows).
constexpr int LW = SPHB200_CRK_LIGHT_WARPS, LS = SPHB200_CRK_LIGHT_STAGES;
constexpr int CW = SPHB200_CRK_WARPS, CS = SPHB200_CRK_STAGES;

template <int DIM> struct Ck {
  static constexpr int PS = DIM + 1;            // polynomialSize of LinearOrder (RKUtilitiesInline.hh:44-55)
  static constexpr int NC = PS*(1 + DIM);       // correctionsSize(false): C, then dC_d per direction
  static constexpr int CST = (DIM == 3) ? 16 : 10;   // stride of the sorted copy (16-byte aligned records)
  static constexpr int QST = (DIM == 3) ? 10 : 6;    // sorted record {volume, Q velocity gradient (ndim^2), pad}
};
template <int DIM> struct LightPrefix { static constexpr int BYTES = ((Dm<DIM>::R_M + 1)*8 + 15)/16*16; };   // position .. mass: 112 B (3-D) / 64 B (2-D)
template <int DIM> using LightRing = NbrRing<DIM, 0, 0, LS, LightPrefix<DIM>::BYTES>;
template <int DIM> using PosRing = NbrRing<DIM, 0, 0, LS, (DIM == 3 ? 32 : 16), false>;     // the position only, no aux record
template <int DIM> using DerivRing = NbrRing<DIM, Ck<DIM>::CST*8, Ck<DIM>::QST*8, CS, Dm<DIM>::ROW*8, false>;   // det Hj recomputed, Vj in the Q record

// api (host order, AoS, width doubles per node) -> sorted copy with stride `stride`
__global__ void __launch_bounds__(RB) k_gather_sorted(const double* __restrict__ api, int width, int stride,
                                                      const uint32_t* __restrict__ perm, size_t n, double* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x*RB + threadIdx.x;
  if (t >= n*(size_t)stride) return;
  const size_t s = t/stride; const int q = (int)(t - s*stride);
  out[t] = (q < width) ? api[(size_t)perm[s]*width + q] : 0.0;
}

// {volume (host order), Q velocity gradient (already sorted, or absent)} -> sorted records of stride qst
__global__ void __launch_bounds__(RB) k_crk_qrec(const double* __restrict__ volApi, const double* __restrict__ dvdxqS, int nt, int qst,
                                                 const uint32_t* __restrict__ perm, size_t n, double* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x*RB + threadIdx.x;
  if (t >= n*(size_t)qst) return;
  const size_t s = t/qst; const int q = (int)(t - s*qst);
  out[t] = (q == 0) ? volApi[perm[s]] : ((q <= nt && dvdxqS) ? dvdxqS[s*nt + (q - 1)] : 0.0);
}

// sorted volume + the CRK aux record {det H, volume} the light loops stream through the ring in place of {det H, 1/rho}
__global__ void __launch_bounds__(RB) k_crk_volaux(const double* __restrict__ volApi, const double* __restrict__ aux2,
                                                   const uint32_t* __restrict__ perm, size_t n, double* __restrict__ volS, double* __restrict__ vaux) {
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  const double v = volApi[perm[s]];
  volS[s] = v;
  vaux[2*s] = aux2[2*s]; vaux[2*s + 1] = v;
}

struct CrkArgs {
  const double* rows; const double* aux2; const uint32_t* perm;
  const uint32_t* nbrCount; const uint32_t* tileRows; const unsigned long long* tileOff; const uint32_t* nbr;
  const double* volS; const double* corrS; const double* qrecS;
  const double *auxDvDxQ, *auxfCl, *auxfCq;
  const double* tabW; double kext, xmin, xstep; uint32_t n1;
  const double* nperhVals; uint32_t nperhN; double nperhXmin, nperhXmax, nperhXstep;
  double W0, gW0;                 // table values at eta = 0 (Hdet = 1)
  double etaVolMax, rhoMin, rhoMax;
  size_t n, cap; uint32_t nInt;
  sphb200_options o;
  double* volApi; double* corrApi; double* rhoApi; const double* massApi;
  double* deriv[DV_COUNT];
  double* pacc;
};

// stage the interleaved W/gradW table (+ one all-zero record) in shared memory; returns its 32-bit shared address
__device__ __forceinline__ unsigned stage_table(double* smem, const double* __restrict__ tab, uint32_t n1) {
  const uint32_t nW = 6u*(n1 + 2u);
  for (uint32_t k = threadIdx.x; k < nW; k += blockDim.x) smem[k] = (k < nW - 6u) ? tab[k] : 0.0;
  __syncthreads();
  return (unsigned)__cvta_generic_to_shared(smem);
}

template <int DIM> __device__ __forceinline__ void load_row(const double* __restrict__ rows, size_t j, double* rw) {
  constexpr int ROW = Dm<DIM>::ROW;
  const double2* p = reinterpret_cast<const double2*>(rows + j*ROW);
#pragma unroll
  for (int q = 0; q < ROW/2; ++q) { const double2 v = __ldg(p + q); rw[2*q] = v.x; rw[2*q + 1] = v.y; }
}

// per-lane view of one tile of 32 Morton-consecutive nodes
struct TileLane { size_t i; bool inRange, active; uint32_t o, cnt, rowsT; unsigned long long base; };
__device__ __forceinline__ TileLane tile_lane(const CrkArgs& a, size_t tile, int lane) {
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
template <typename Ring> __device__ __forceinline__ Ring make_ring(const CrkArgs& a, unsigned tW, int warp) {
  Ring r;
  r.base = tW + 48u*(a.n1 + 2u) + (unsigned)warp*(unsigned)Ring::WARPB;
  r.rows = reinterpret_cast<const unsigned char*>(a.rows);
  r.x1 = reinterpret_cast<const unsigned char*>(a.corrS);
  r.x2 = reinterpret_cast<const unsigned char*>(a.qrecS);
  r.aux2 = reinterpret_cast<const unsigned char*>(a.aux2);
  return r;
}

// ---- computeRKSumVolume ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32*LW, 1) k_crk_volume(CrkArgs a) {
  using D = Dm<DIM>;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const PosRing<DIM> ring = make_ring<PosRing<DIM>>(a, tW, warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*LW + warp; tile < nTiles; tile += (size_t)gridDim.x*LW) {
    const TileLane t = tile_lane(a, tile, lane);
    double ri[DIM], Hi[D::NS];
#pragma unroll
    for (int k = 0; k < DIM; ++k) ri[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_POS + k] : 0.0;
#pragma unroll
    for (int k = 0; k < D::NS; ++k) Hi[k] = t.inRange ? a.rows[t.i*D::ROW + D::R_H + k] : 0.0;
    double sum = 0.0;
    ring_walk<PosRing<DIM>, LS>(ring, lane, t.rowsT, t.cnt,
      [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t) {
        double rj[DIM == 3 ? 4 : 2];
        ring.read_row(k, lane, rj);                            // the position leads the row
        double rij[DIM], eta[DIM];
#pragma unroll
        for (int q = 0; q < DIM; ++q) rij[q] = ri[q] - rj[q];
        sym_dot<DIM>(Hi, rij, eta);
        const double e2 = vdot<DIM>(eta, eta);
        const double etaMag = e2*fast_rsqrt(e2 + 1.0e-300);
        double W, gW;
        table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, etaMag, W, gW);
        sum += W;                                            // Wi = Hdeti * W(eta_i), Hdeti applied once below
      });
    if (t.active) {
      const double Hdeti = sym_det<DIM>(Hi);
      const double s = sum*Hdeti + Hdeti*a.W0;               // :104-113 self contribution and the eta-space cap
      a.volApi[t.o] = fmin(a.etaVolMax/Hdeti, 1.0/s);
    }
  }
}

// ---- Eigen::ColPivHouseholderQR (the solver of RKUtilities.cc:404-425), N x N in registers ---------------------------------
// Published algorithm: at step k the remaining column of largest norm is swapped into place, a Householder reflector
// annihilates it below the diagonal and is applied to the trailing columns; solve = Q^T b, back substitution on the
// leading rank x rank triangle (rank: |R_kk| > eps*N*max|R_kk|), permutation undone.  All loops have compile-time
// bounds so the matrix never leaves registers.
template <int N> struct QrDev { double A[N][N]; double tau[N]; int perm[N]; int rank; };

template <int N> __device__ __forceinline__ void qr_factor(QrDev<N>& q) {
#pragma unroll
  for (int k = 0; k < N; ++k) q.perm[k] = k;
  double maxPivot = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    int best = k; double bestN = -1.0;
#pragma unroll
    for (int c = k; c < N; ++c) {
      double s = 0.0;
#pragma unroll
      for (int r = k; r < N; ++r) s = fma(q.A[r][c], q.A[r][c], s);
      if (s > bestN) { bestN = s; best = c; }
    }
#pragma unroll
    for (int c = k + 1; c < N; ++c)
      if (c == best) {
#pragma unroll
        for (int r = 0; r < N; ++r) { const double t = q.A[r][k]; q.A[r][k] = q.A[r][c]; q.A[r][c] = t; }
        const int t = q.perm[k]; q.perm[k] = q.perm[c]; q.perm[c] = t;
      }
    const double c0 = q.A[k][k];
    double tail = 0.0;
#pragma unroll
    for (int r = k + 1; r < N; ++r) tail = fma(q.A[r][k], q.A[r][k], tail);
    double beta, tau;
    if (tail <= 2.2250738585072014e-308) {
      tau = 0.0; beta = c0;
#pragma unroll
      for (int r = k + 1; r < N; ++r) q.A[r][k] = 0.0;
    } else {
      beta = sqrt(fma(c0, c0, tail)); if (c0 >= 0.0) beta = -beta;
      const double inv = 1.0/(c0 - beta);
#pragma unroll
      for (int r = k + 1; r < N; ++r) q.A[r][k] *= inv;
      tau = (beta - c0)/beta;
    }
    q.A[k][k] = beta; q.tau[k] = tau;
    maxPivot = fmax(maxPivot, fabs(beta));
#pragma unroll
    for (int c = k + 1; c < N; ++c) {
      double s = q.A[k][c];
#pragma unroll
      for (int r = k + 1; r < N; ++r) s = fma(q.A[r][k], q.A[r][c], s);
      s *= tau;
      q.A[k][c] -= s;
#pragma unroll
      for (int r = k + 1; r < N; ++r) q.A[r][c] = fma(-s, q.A[r][k], q.A[r][c]);
    }
  }
  const double thr = maxPivot*(2.220446049250313e-16*(double)N);
  int rank = 0;
#pragma unroll
  for (int k = 0; k < N; ++k) rank += (fabs(q.A[k][k]) > thr) ? 1 : 0;
  q.rank = rank;
}

template <int N> __device__ __forceinline__ void qr_solve(const QrDev<N>& q, const double* rhs, double* x) {
  double c[N];
#pragma unroll
  for (int k = 0; k < N; ++k) c[k] = rhs[k];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    double s = c[k];
#pragma unroll
    for (int r = k + 1; r < N; ++r) s = fma(q.A[r][k], c[r], s);
    s *= q.tau[k];
    c[k] -= s;
#pragma unroll
    for (int r = k + 1; r < N; ++r) c[r] = fma(-s, q.A[r][k], c[r]);
  }
  const int nz = q.rank;
#pragma unroll
  for (int k = N - 1; k >= 0; --k) {
    double s = c[k];
#pragma unroll
    for (int r = k + 1; r < N; ++r) s -= (r < nz) ? q.A[k][r]*c[r] : 0.0;
    c[k] = (k < nz) ? s/q.A[k][k] : 0.0;
  }
#pragma unroll
  for (int t = 0; t < N; ++t) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) v = (q.perm[k] == t) ? c[k] : v;
    x[t] = v;
  }
}

// ---- RKUtilities<Dim, LinearOrder>::computeCorrections ----------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32*LW, 1) k_crk_corrections(CrkArgs a) {
  using D = Dm<DIM>;
  constexpr int PS = Ck<DIM>::PS, NC = Ck<DIM>::NC;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LightRing<DIM> ring = make_ring<LightRing<DIM>>(a, tW, warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*LW + warp; tile < nTiles; tile += (size_t)gridDim.x*LW) {
  const TileLane t = tile_lane(a, tile, lane);
  const size_t i = t.i; const uint32_t o = t.o;
  double ri[DIM];
#pragma unroll
  for (int k = 0; k < DIM; ++k) ri[k] = t.inRange ? a.rows[i*D::ROW + D::R_POS + k] : 0.0;

  double M[PS][PS], dM[DIM][PS][PS];
#pragma unroll
  for (int k = 0; k < PS; ++k)
#pragma unroll
    for (int l = 0; l < PS; ++l) { M[k][l] = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) dM[d][k][l] = 0.0; }

  // addToMatrix (RKUtilities.cc:296-352): w, dw from the base kernel of node j evaluated at x_ij
  auto add = [&](const double* xij, double vj, double w, const double* dw) {
    double p[PS];
    p[0] = 1.0;
#pragma unroll
    for (int q = 0; q < DIM; ++q) p[1 + q] = xij[q];
    const double vw = vj*w;
#pragma unroll
    for (int k = 0; k < PS; ++k)
#pragma unroll
      for (int l = k; l < PS; ++l) {
        const double pp = p[k]*p[l];
        M[k][l] = fma(vw, pp, M[k][l]);
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          // dp_d = e_{1+d}: (dp[k] p[l] + p[k] dp[l]) w + p[k] p[l] dw_d
          double t = pp*dw[d];
          if (k == 1 + d) t = fma(p[l], w, t);
          if (l == 1 + d) t = fma(p[k], w, t);
          dM[d][k][l] = fma(vj, t, dM[d][k][l]);
        }
      }
  };

  ring_walk<LightRing<DIM>, LS>(ring, lane, t.rowsT, t.cnt,
    [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
    [&](uint32_t k, uint32_t) {
    double rw[LightPrefix<DIM>::BYTES/8];
    ring.read_row(k, lane, rw);
    const double2 ax = ring.read_aux(k, lane);              // {det Hj, Vj}
    const double Hdetj = ax.x, vj = ax.y;
    double xij[DIM], eta[DIM], Heta[DIM], dw[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) xij[q] = ri[q] - rw[D::R_POS + q];
    sym_dot<DIM>(rw + D::R_H, xij, eta);
    const double e2 = vdot<DIM>(eta, eta);
    const double inv = fast_rsqrt(e2 + 1.0e-300);
    double W, gW;
    table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, e2*inv, W, gW);
    sym_dot<DIM>(rw + D::R_H, eta, Heta);                  // H.etaUnit*dk == (dk/|eta|) * (H.eta)  (RKUtilities.cc:70-77)
    const double sc = gW*Hdetj*inv;
#pragma unroll
    for (int q = 0; q < DIM; ++q) dw[q] = sc*Heta[q];
    add(xij, vj, W*Hdetj, dw);
  });
  if (!t.active) continue;
  {
    // self contribution (RKUtilities.cc:383): x = 0, unitVector() of the zero vector is (1,0,0) (GeomVectorInline.hh:998-1001)
    double Hi[D::NS], xij[DIM], dw[DIM];
#pragma unroll
    for (int k = 0; k < D::NS; ++k) Hi[k] = a.rows[i*D::ROW + D::R_H + k];
    const double Hdeti = a.aux2[2*i];
    const double dk = a.gW0*Hdeti;
#pragma unroll
    for (int q = 0; q < DIM; ++q) { xij[q] = 0.0; dw[q] = Hi[q]*dk; }      // first column of the symmetric H: xx, xy[, xz]
    add(xij, a.volS[i], a.W0*Hdeti, dw);
  }
  QrDev<PS> q;
#pragma unroll
  for (int k = 0; k < PS; ++k)
#pragma unroll
    for (int l = 0; l < PS; ++l) q.A[k][l] = (l >= k) ? M[k][l] : M[l][k];
  qr_factor<PS>(q);
  double rhs[PS], C[PS], dC[PS];
  rhs[0] = 1.0;
#pragma unroll
  for (int k = 1; k < PS; ++k) rhs[k] = 0.0;
  qr_solve<PS>(q, rhs, C);
  double* out = a.corrApi + (size_t)o*NC;
#pragma unroll
  for (int k = 0; k < PS; ++k) out[k] = C[k];
#pragma unroll
  for (int d = 0; d < DIM; ++d) {
#pragma unroll
    for (int k = 0; k < PS; ++k) {
      double s = 0.0;
#pragma unroll
      for (int l = 0; l < PS; ++l) s = fma((l >= k) ? dM[d][k][l] : dM[d][l][k], C[l], s);
      rhs[k] = -s;
    }
    qr_solve<PS>(q, rhs, dC);
#pragma unroll
    for (int k = 0; k < PS; ++k) out[PS*(1 + d) + k] = dC[k];
  }
  }   // tile loop
}

// ---- computeCRKSPHSumMassDensity ----------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(32*LW, 1) k_crk_sum_density(CrkArgs a) {
  using D = Dm<DIM>;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const LightRing<DIM> ring = make_ring<LightRing<DIM>>(a, tW, warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*LW + warp; tile < nTiles; tile += (size_t)gridDim.x*LW) {
    const TileLane t = tile_lane(a, tile, lane);
    const size_t i = t.i;
    double ri[DIM];
#pragma unroll
    for (int k = 0; k < DIM; ++k) ri[k] = t.inRange ? a.rows[i*D::ROW + D::R_POS + k] : 0.0;
    double wsum = 0.0, md = 0.0, vol1 = 0.0;
    ring_walk<LightRing<DIM>, LS>(ring, lane, t.rowsT, t.cnt,
      [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t) {
        double rw[LightPrefix<DIM>::BYTES/8];
        ring.read_row(k, lane, rw);
        const double2 ax = ring.read_aux(k, lane);            // {det Hj, Vj}
        const double Hdetj = ax.x, Vj = ax.y;
        double rij[DIM], eta[DIM];
#pragma unroll
        for (int q = 0; q < DIM; ++q) rij[q] = ri[q] - rw[D::R_POS + q];
        sym_dot<DIM>(rw + D::R_H, rij, eta);
        const double e2 = vdot<DIM>(eta, eta);
        double W, gW;
        table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, e2*fast_rsqrt(e2 + 1.0e-300), W, gW);
        const double VW = Vj*(W*Hdetj);                        // :87-92, i side
        wsum += VW; md = fma(rw[D::R_M], VW, md); vol1 = fma(Vj, VW, vol1);
      });
    if (t.active) {
      const double mi = a.rows[i*D::ROW + D::R_M], Vi = a.volS[i], Hdeti = a.aux2[2*i];
      const double ws = wsum + Vi*Hdeti*a.W0;                  // :115-122
      const double v1 = (vol1 + Vi*Vi*Hdeti*a.W0)/ws;
      a.rhoApi[t.o] = fmax(fmax(a.rhoMin, 0.1*mi*Hdeti), fmin(a.rhoMax, (md + mi*Vi*Hdeti*a.W0)/(ws*v1)));
    }
  }
}

// ---- CRKSPH<Dim>::evaluateDerivativesImpl + smoothing-scale sub-package ---------------------------------------------------------
//   GEN  : every option of the reference at run time (Balsara, Cl/Cq multipliers, linear / quadraticInExpansion, any
//          XSPH / smoothing-scale / compatible-energy choice)
//   !GEN : the factory configuration of CRKSPHHydros.py (XSPH, SPH smoothing scale, compatible energy; LIMITED picks
//          LimitedMonaghanGingold or plain MonaghanGingold) fixed at compile time, so the unused paths cost no registers
template <int DIM, bool GEN, bool LIMITED>
__global__ void __launch_bounds__(32*CW, 1) k_crk_derivs(CrkArgs a) {
  using D = Dm<DIM>;
  constexpr int NS = D::NS, NT = D::NT, ROW = D::ROW, PS = Ck<DIM>::PS, CST = Ck<DIM>::CST, QST = Ck<DIM>::QST;
  extern __shared__ __align__(16) double smem[];
  const unsigned tW = stage_table(smem, a.tabW, a.n1);
  const double rx = 1.0/a.xstep;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const DerivRing<DIM> ring = make_ring<DerivRing<DIM>>(a, tW, warp);
  const size_t nTiles = (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*CW + warp; tile < nTiles; tile += (size_t)gridDim.x*CW) {
  const TileLane t = tile_lane(a, tile, lane);
  const size_t i = t.i;
  const bool inRange = t.inRange, active = t.active;
  const sphb200_options& op = a.o;
  const bool xsph = GEN ? op.XSPH != 0 : true, hsph = GEN ? op.hEvolution == SPHB200_H_SPH : true;
  const bool compat = GEN ? op.compatibleEnergy != 0 : true;
  const bool limited = GEN ? op.Qkind == SPHB200_Q_LIMITED_MG : LIMITED;
  const bool bals = GEN && op.balsara;
  const bool needQ = limited || bals;
  const bool mult = GEN && a.auxfCl != nullptr;
  const bool linExp = GEN && op.linearInExpansion, quadExp = GEN && op.quadraticInExpansion;
  const double etaCrit = op.etaCritFrac/op.nPerh, rEtaFold = op.nPerh/op.etaFoldFrac;

  double rwi[ROW], ci_[CST];
  if (inRange) {
    load_row<DIM>(a.rows, i, rwi);
#pragma unroll
    for (int q = 0; q < CST; ++q) ci_[q] = a.corrS[i*CST + q];
  } else {
#pragma unroll
    for (int q = 0; q < ROW; ++q) rwi[q] = 0.0;
#pragma unroll
    for (int q = 0; q < CST; ++q) ci_[q] = 0.0;
    rwi[D::R_M] = 1.0; rwi[D::R_RHO] = 1.0;
  }
  const double* ri = rwi + D::R_POS; const double* vi = rwi + D::R_VEL; const double* Hi = rwi + D::R_H;
  const double mi = rwi[D::R_M], rhoi = rwi[D::R_RHO], Pi = rwi[D::R_PRHO], csi = rwi[D::R_CS];
  const double Hdeti = inRange ? a.aux2[2*i] : 0.0;
  const double voli = inRange ? a.volS[i] : 0.0;
  const double miInv = 1.0/mi;
  double DvDxQi[NT];
#pragma unroll
  for (int q = 0; q < NT; ++q) DvDxQi[q] = (needQ && inRange) ? a.auxDvDxQ[i*NT + q] : 0.0;
  const double fCli = (mult && inRange) ? a.auxfCl[i] : 1.0, fCqi = (mult && inRange) ? a.auxfCq[i] : 1.0;
  const double balsi = (bals && inRange) ? balsara<DIM>(op, DvDxQi, Hdeti, csi) : 1.0;

  double DepsDt = 0, maxQ = 0, effQ = 0, m0 = 0;
  double DvDt[DIM], XdV[DIM], m1[DIM], DvDx[NT];
#pragma unroll
  for (int q = 0; q < DIM; ++q) { DvDt[q] = 0; XdV[q] = 0; m1[q] = 0; }
#pragma unroll
  for (int q = 0; q < NT; ++q) DvDx[q] = 0;

  double* const paccTile = compat ? a.pacc + (size_t)DIM*a.tileOff[tile] + lane : nullptr;

  ring_walk<DerivRing<DIM>, CS>(ring, lane, t.rowsT, t.cnt,
    [&](uint32_t p) -> uint32_t { return (p < t.cnt) ? a.nbr[t.base + (unsigned long long)p*SPHB200_TILE] : 0u; },
    [&](uint32_t k, uint32_t j) {
    // records are read from the ring where they are first needed (42 doubles of neighbour state would otherwise be live at once)
    double rw[ROW], cj_[CST], qj_[QST];
    ring.read_row(k, lane, rw);
    const double Hdetj = sym_det<DIM>(rw + D::R_H);         // recomputed: a per-lane 16-byte aux copy costs as much as the 32 rows (nbr_ring.cuh)
    const double volj = ring_lds128(ring.stage(k) + DerivRing<DIM>::X2OFF + (unsigned)lane*DerivRing<DIM>::X2B).x;   // {Vj, Q velocity gradient}
    const double* rj = rw + D::R_POS; const double* vj = rw + D::R_VEL; const double* Hj = rw + D::R_H;
    const double mj = rw[D::R_M], rhoj = rw[D::R_RHO], Pj = rw[D::R_PRHO], csj = rw[D::R_CS];

    // CRKSPH.cc:333-336
    double rij[DIM], vij[DIM], etai[DIM], etaj[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { rij[q] = ri[q] - rj[q]; vij[q] = vi[q] - vj[q]; }
    sym_dot<DIM>(Hi, rij, etai);
    sym_dot<DIM>(Hj, rij, etaj);
    const double e2i = vdot<DIM>(etai, etai), e2j = vdot<DIM>(etaj, etaj);
    const double invi = fast_rsqrt(e2i + 1.0e-300), invj = fast_rsqrt(e2j + 1.0e-300);
    const double etaMagi = e2i*invi, etaMagj = e2j*invj;

    // base kernels (RKUtilities.cc:64-78): w_j = W(|Hj rij|) Hdetj at x = +rij; w_i = W(|Hi rij|) Hdeti at x = -rij
    double Wbi, gWbi, Wbj, gWbj;
    table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, etaMagi, Wbi, gWbi);
    table_eval_raw(tW, a.kext, a.xmin, a.xstep, rx, a.n1, etaMagj, Wbj, gWbj);
    const double gWiRaw = gWbi;
    Wbi *= Hdeti; gWbi *= Hdeti; Wbj *= Hdetj; gWbj *= Hdetj;
    double Hei[DIM], Hej[DIM];
    sym_dot<DIM>(Hi, etai, Hei);
    sym_dot<DIM>(Hj, etaj, Hej);
    const double sj = gWbj*invj, si = -(gWbi*invi);          // x = -rij flips the unit vector of the i-side base gradient

    ring.template read_x1<CST>(k, lane, cj_);
    // :339-341 evaluateKernelAndGradient (RKUtilities.cc:180-209), P = {1, x}, dP_d = e_{1+d}
    //   (Wj, gradWj) = WR( rij, Hj, corrections_i)     (Wi, gradWi) = WR(-rij, Hi, corrections_j)
    double CPj = ci_[0], CPi = cj_[0];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { CPj = fma(ci_[1 + q], rij[q], CPj); CPi = fma(-cj_[1 + q], rij[q], CPi); }
    double gradWj[DIM], gradWi[DIM], dg[DIM];
#pragma unroll
    for (int d = 0; d < DIM; ++d) {
      double tj = ci_[1 + d] + ci_[PS*(1 + d)], ti = cj_[1 + d] + cj_[PS*(1 + d)];
#pragma unroll
      for (int q = 0; q < DIM; ++q) { tj = fma(ci_[PS*(1 + d) + 1 + q], rij[q], tj); ti = fma(-cj_[PS*(1 + d) + 1 + q], rij[q], ti); }
      gradWj[d] = fma(tj, Wbj, CPj*(sj*Hej[d]));
      gradWi[d] = fma(ti, Wbi, CPi*(si*Hei[d]));
      dg[d] = gradWj[d] - gradWi[d];
    }
    const double Wj = CPj*Wbj;

    // :344-359 artificial viscosity (MonaghanGingoldViscosity.cc:69-100 / LimitedMonaghanGingoldViscosity.cc:140-218)
    double vijQ[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) vijQ[q] = vij[q];
    double Clij = op.Cl, Cqij = op.Cq;
    {
      double fshear = 1.0, fClj = 1.0, fCqj = 1.0;
      if (mult) { fClj = a.auxfCl[j]; fCqj = a.auxfCq[j]; }
      if (needQ) {
        ring.template read_x2<QST>(k, lane, qj_);            // {Vj, Q velocity gradient of j}
        const double* DvDxQj = qj_ + 1;
        if (bals) fshear = 0.5*(balsi + balsara<DIM>(op, DvDxQj, Hdetj, csj));
        if (limited) {
          double xij[DIM], t1[DIM], t2[DIM];
#pragma unroll
          for (int q = 0; q < DIM; ++q) xij[q] = 0.5*rij[q];
          ten_dot<DIM>(DvDxQi, xij, t1); const double gradi = vdot<DIM>(t1, xij);
          ten_dot<DIM>(DvDxQj, xij, t2); const double gradj = vdot<DIM>(t2, xij);
          // safeInvVar / van Leer with refined hardware reciprocals (|argument| >= 1e-30, so never denormal)
          const double agj = fabs(gradj), agi = fabs(gradi);
          const double rri = gradi*(d_sgn(gradj)*fast_rcp(agj > 1.0e-30 ? agj : 1.0e-30));
          const double rrj = gradj*(d_sgn(gradi)*fast_rcp(agi > 1.0e-30 ? agi : 1.0e-30));
          const double x = rri < rrj ? rri : rrj;
          double phi = 0.0;
          if (x > 0.0) { const double r1 = fast_rcp(1.0 + x); phi = 2.0*r1*2.0*x*r1; }
          const double etaij = etaMagi < etaMagj ? etaMagi : etaMagj;
          if (etaij < etaCrit) { const double z = (etaij - etaCrit)*rEtaFold; phi *= exp(-z*z); }
#pragma unroll
          for (int q = 0; q < DIM; ++q) vijQ[q] = (vi[q] - phi*t1[q]) - (vj[q] + phi*t2[q]);
        }
      }
      if (GEN) {
        Clij = 0.5*(fCli + fClj)*fshear*op.Cl;
        Cqij = 0.5*(fCqi + fCqj)*fshear*op.Cq;
      }
    }
    const double mui = vdot<DIM>(vijQ, etai)*fast_rcp(e2i + op.eps2);
    const double muj = vdot<DIM>(vijQ, etaj)*fast_rcp(e2j + op.eps2);
    const double mui0 = mui < 0.0 ? mui : 0.0, muj0 = muj < 0.0 ? muj : 0.0;
    const double ei = -Clij*csi*(linExp ? mui : mui0) + Cqij*(quadExp ? -d_sgn(mui)*mui*mui : mui0*mui0);
    const double ej = -Clij*csj*(linExp ? muj : muj0) + Cqij*(quadExp ? -d_sgn(muj)*muj*muj : muj0*muj0);
    const double Qi = rhoi*ei, Qj = rhoj*ej;                 // rho_i^2 QPiij = rho_i e_i = Qi (QPiij = e_i/rho_i)
    const double vdg = vdot<DIM>(vij, dg);
    { const double q4 = 4.0*Qi; maxQ = q4 > maxQ ? q4 : maxQ; }   // :354
    effQ = fma(volj*Qi, Wj, effQ);                           // :356

    // :362-368 velocity gradient
    { double g[DIM];
#pragma unroll
      for (int q = 0; q < DIM; ++q) g[q] = volj*gradWj[q];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int c2 = 0; c2 < DIM; ++c2) DvDx[r*DIM + c2] = fma(-vij[r], g[c2], DvDx[r*DIM + c2]);
    }

    // :376-384 Type III interpoint force, :387-389 energy
    const double hvv = 0.5*voli*volj;
    const double fsc = hvv*((Pi + Pj) + (Qi + Qj));           // force = fsc * deltagrad
    const double accsc = fsc*miInv;
#pragma unroll
    for (int q = 0; q < DIM; ++q) DvDt[q] = fma(-accsc, dg[q], DvDt[q]);
    if (compat) {
      // stored per directed edge as force/(mi mj): -mj*stored = -force/mi is the pair acceleration of the pair's i-node (:384),
      // mi*stored the (antisymmetric) one seen from its j-node -- the convention k_energy / k_emit_pacc use for SPH
      const double ps = accsc*fast_rcp(mj);
      double* const paccRow = paccTile + (size_t)k*(DIM*32);
#pragma unroll
      for (int q = 0; q < DIM; ++q) paccRow[32*q] = ps*dg[q];
    }
    DepsDt = fma(hvv*miInv, (Pj + Qj)*vdg, DepsDt);           // workQi = rho_j^2 QPiji vij.deltagrad = Qj vdg

    // :395-398 XSPH
    if (xsph) {
      const double w = volj*Wj;
#pragma unroll
      for (int q = 0; q < DIM; ++q) XdV[q] = fma(-w, vij[q], XdV[q]);
    }
    // SPHSmoothingScale.cc:186-222
    if (hsph) {
      const double WSPHi = fabs(gWiRaw);
      m0 += WSPHi;
#pragma unroll
      for (int q = 0; q < DIM; ++q) m1[q] = fma(-WSPHi, etai[q], m1[q]);
    }
  });

  if (!inRange) continue;
  const size_t cap = a.cap;
  auto put = [&](int slot, int comp, double v) { a.deriv[slot][(size_t)comp*cap + i] = v; };
  if (!active) {
    for (int s = 0; s < DV_COUNT; ++s) { const int w = sphb200_deriv_width(DIM, s); for (int q = 0; q < w; ++q) put(s, q, 0.0); }
    continue;
  }
  // :409-437
#pragma unroll
  for (int q = 0; q < DIM; ++q) { put(DV_DXDT, q, xsph ? vi[q] + XdV[q] : vi[q]); put(DV_DVDT, q, DvDt[q]); put(DV_XSPHDV, q, XdV[q]); put(DV_GRADRHO, q, 0.0); }
  put(DV_DRHODT, 0, -rhoi*ten_trace<DIM>(DvDx));
  if (GEN && op.evolveTotalEnergy) DepsDt = mi*(vdot<DIM>(vi, DvDt) + DepsDt);
  put(DV_DEPSDT, 0, DepsDt);
  put(DV_RHOSUM, 0, 0.0); put(DV_NORM, 0, 0.0); put(DV_MAXQ, 0, maxQ); put(DV_EFFQ, 0, effQ); put(DV_XSPHW, 0, 0.0);
#pragma unroll
  for (int q = 0; q < NT; ++q) { put(DV_DVDX, q, DvDx[q]); put(DV_LOCALDVDX, q, DvDx[q]); put(DV_M, q, 0.0); put(DV_LOCALM, q, 0.0); }
  // smoothing scale (as in k_sph_derivs)
  if (hsph) {
    const double z0 = rootnu<DIM>(fmax(0.0, m0));
    put(DV_M0, 0, z0);
#pragma unroll
    for (int q = 0; q < DIM; ++q) put(DV_M1, q, m1[q]);
    const double tr = ten_trace<DIM>(DvDx);
    const double dinv = 1.0/(double)DIM;
#pragma unroll
    for (int q = 0; q < NS; ++q) put(DV_DHDT, q, ((-Hi[q])*dinv)*tr);
    const bool isolated = fabs(z0 - 0.0) <= 1.0e-15*fmax(1.0, fabs(z0));
    const double cur = isolated ? 0.5*op.nPerh : fmax(0.0, hermite_eval(a.nperhVals, a.nperhN, a.nperhXmin, a.nperhXmax, a.nperhXstep, z0));
    const double sv = fmin(4.0, fmax(0.25, op.nPerh/(cur + 1.0e-30)));
    const double aa = (sv < 1.0 ? 0.4*(1.0 + sv*sv) : 0.4*(1.0 + 1.0/(sv*sv*sv)));
    const double hi0 = 1.0/Hi[0];
    const double hi1 = fmin(op.hmax, fmax(op.hmin, hi0*(1.0 - aa + aa*sv)));
    const double hinv = 1.0/hi1;
    if (DIM == 3) { put(DV_HIDEAL, 0, hinv); put(DV_HIDEAL, 1, 0.0); put(DV_HIDEAL, 2, 0.0); put(DV_HIDEAL, 3, hinv); put(DV_HIDEAL, 4, 0.0); put(DV_HIDEAL, 5, hinv); }
    else { put(DV_HIDEAL, 0, hinv); put(DV_HIDEAL, 1, 0.0); put(DV_HIDEAL, 2, hinv); }
  } else {
    put(DV_M0, 0, 0.0);
#pragma unroll
    for (int q = 0; q < DIM; ++q) put(DV_M1, q, 0.0);
    double dh[NS];
    if (GEN && sphb200_is_asph(op.hEvolution)) asph_DHDt<DIM>(Hi, DvDx, dh);      // classic ASPH: the ideal H follows in k_asph_classic
    else {
#pragma unroll
      for (int q = 0; q < NS; ++q) dh[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) { put(DV_DHDT, q, dh[q]); put(DV_HIDEAL, q, 0.0); }
  }
  }   // tile loop
}

double host_table_value(const TableDev& t, double eta, bool grad) {
  if (!(eta < t.kext)) return 0.0;
  const double q = std::max(0.0, eta - t.xmin)/t.xstep;
  const size_t k = std::min<size_t>((size_t)q, t.n1);
  const std::vector<double>& c = grad ? t.hostG : t.hostW;
  return c[3*k] + (c[3*k + 1] + c[3*k + 2]*eta)*eta;
}

int crk_common(sphb200_ctx* c, CrkArgs& a, const char* who) {
  if (c->opt.hydro != SPHB200_HYDRO_CRKSPH) return sphb200_fail(c, std::string(who) + ": the context was not created for CRKSPH (options.hydro)");
  if (!c->pairsValid) return sphb200_fail(c, std::string(who) + ": connectivity is stale or missing (call build_pairs first)");
  if (!c->W.set) return sphb200_fail(c, std::string(who) + ": kernel table not set");
  // positions / H may have moved on a kept connectivity (the mid-step evaluation of CheapSynchronousRK2): the sorted rows follow
  if (c->n && !c->rowsValid && sphb200_pack_rows(c)) return 1;
  a = CrkArgs{};
  a.rows = c->rows; a.aux2 = c->aux2; a.perm = c->perm; a.nbrCount = c->nbrCount; a.tileRows = c->tileRows; a.tileOff = c->tileOff; a.nbr = c->nbr;
  a.volS = c->crkVolS; a.corrS = c->crkCorrS; a.qrecS = c->crkQS;
  a.tabW = c->W.coef; a.kext = c->W.kext; a.xmin = c->W.xmin; a.xstep = c->W.xstep; a.n1 = c->W.n1;
  a.nperhVals = c->W.nperhVals; a.nperhN = c->W.nperhN; a.nperhXmin = c->W.nperhXmin; a.nperhXmax = c->W.nperhXmax; a.nperhXstep = c->W.nperhXstep;
  a.W0 = host_table_value(c->W, 0.0, false); a.gW0 = host_table_value(c->W, 0.0, true);
  const double half = 0.5*c->W.kext;
  a.etaVolMax = (c->ndim == 3) ? half*half*half*(4.0/3.0*M_PI) : half*half*M_PI;      // computeRKSumVolume.cc:46
  a.n = c->n; a.cap = c->cap; a.nInt = (uint32_t)c->nInt; a.o = c->opt;
  a.volApi = c->api[S_VOLUME]; a.corrApi = c->api[S_RKCORR]; a.rhoApi = c->api[S_RHO]; a.massApi = c->api[S_MASS];
  for (int s = 0; s < DV_COUNT; ++s) a.deriv[s] = c->deriv[s];
  return 0;
}

int gather(sphb200_ctx* c, int slot, int stride, double* out) {
  const int w = sphb200_state_width(c->ndim, slot);
  const size_t total = c->n*(size_t)stride;
  k_gather_sorted<<<(unsigned)((total + RB - 1)/RB), RB, 0, c->stream>>>(c->api[slot], w, stride, c->perm, c->n, out);
  KERNEL_CHECK(c, "k_gather_sorted");
  return 0;
}
// volume -> sorted copy + {det H, volume} records
int gather_volume(sphb200_ctx* c) {
  k_crk_volaux<<<(unsigned)((c->n + RB - 1)/RB), RB, 0, c->stream>>>(c->api[S_VOLUME], c->aux2, c->perm, c->n, c->crkVolS, c->crkAux);
  KERNEL_CHECK(c, "k_crk_volaux");
  return 0;
}

// persistent launch: one CTA per SM, `warps` tiles in flight per CTA, ring of `ringWarpBytes` per warp behind the table
template <typename K> int launch_tiles(sphb200_ctx* c, K kern, const CrkArgs& a, const char* name, int warps, size_t ringWarpBytes) {
  const size_t shm = (size_t)6*(c->W.n1 + 2)*sizeof(double) + (size_t)warps*ringWarpBytes;
  if (shm > 227*1024) return sphb200_fail(c, "kernel table too large for shared memory");
  CU_CHECK(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  int perSM = 1;                                     // resident CTAs per SM (registers + shared memory): the persistent grid fills them
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, 32*warps, shm) != cudaSuccess || perSM < 1) perSM = 1;
  const unsigned nb = (unsigned)std::min<size_t>((c->nTiles + warps - 1)/warps, (size_t)nsm*perSM);
  kern<<<nb, 32*warps, shm, c->stream>>>(a);
  KERNEL_CHECK(c, name);
  return 0;
}
template <int DIM, typename K> int launch_light(sphb200_ctx* c, K kern, const CrkArgs& a, const char* name) {
  return launch_tiles(c, kern, a, name, LW, (size_t)LightRing<DIM>::WARPB);
}

}  // namespace

int sphb200_launch_crk_derivs(sphb200_ctx* c) {
  CrkArgs a;
  if (crk_common(c, a, "evaluateDerivatives")) return 1;
  const bool needQ = (c->opt.Qkind == SPHB200_Q_LIMITED_MG) || c->opt.balsara;
  const bool mult = c->have[S_FCL] && c->have[S_FCQ];
  a.auxDvDxQ = needQ ? c->auxDvDxQ : nullptr; a.auxfCl = mult ? c->auxfCl : nullptr; a.auxfCq = mult ? c->auxfCq : nullptr;
  if (gather_volume(c)) return 1;
  if (gather(c, S_RKCORR, c->ndim == 3 ? Ck<3>::CST : Ck<2>::CST, c->crkCorrS)) return 1;
  {
    const int nt = c->ndim*c->ndim, qst = c->ndim == 3 ? Ck<3>::QST : Ck<2>::QST;
    const size_t total = c->n*(size_t)qst;
    k_crk_qrec<<<(unsigned)((total + RB - 1)/RB), RB, 0, c->stream>>>(c->api[S_VOLUME], a.auxDvDxQ, nt, qst, c->perm, c->n, c->crkQS);
    KERNEL_CHECK(c, "k_crk_qrec");
  }
  a.aux2 = c->crkAux;
  c->paccMode = PACC_FULL; c->paccWidth = c->ndim;      // the RK-corrected pair force has no two-scalar form
  if (c->opt.compatibleEnergy && sphb200_ensure(c, c->pacc, c->paccCap, c->nSlots*(size_t)c->ndim)) return 1;
  a.pacc = c->pacc;
  if (c->opt.hEvolution == SPHB200_H_SPH && (!a.nperhVals || a.nperhN < 2))
    return sphb200_fail(c, "evaluateDerivatives: SPHSmoothingScale needs the TableKernel nperh lookup (nperhVals) but none was set");
  const sphb200_options& o = c->opt;
  const bool fast = o.XSPH && o.hEvolution == SPHB200_H_SPH && o.compatibleEnergy && !o.evolveTotalEnergy && !o.balsara && !mult &&
                    !o.linearInExpansion && !o.quadraticInExpansion;
  const bool lim = o.Qkind == SPHB200_Q_LIMITED_MG;
  int rc;
  if (c->ndim == 3) rc = !fast ? launch_tiles(c, k_crk_derivs<3, true, true>, a, "k_crk_derivs", CW, (size_t)DerivRing<3>::WARPB)
                        : lim ? launch_tiles(c, k_crk_derivs<3, false, true>, a, "k_crk_derivs", CW, (size_t)DerivRing<3>::WARPB)
                              : launch_tiles(c, k_crk_derivs<3, false, false>, a, "k_crk_derivs", CW, (size_t)DerivRing<3>::WARPB);
  else              rc = !fast ? launch_tiles(c, k_crk_derivs<2, true, true>, a, "k_crk_derivs", CW, (size_t)DerivRing<2>::WARPB)
                        : lim ? launch_tiles(c, k_crk_derivs<2, false, true>, a, "k_crk_derivs", CW, (size_t)DerivRing<2>::WARPB)
                              : launch_tiles(c, k_crk_derivs<2, false, false>, a, "k_crk_derivs", CW, (size_t)DerivRing<2>::WARPB);
  if (rc) return 1;
  c->derivsValid = true;
  return 0;
}

extern "C" {

int sphb200_crk_compute_volume(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  CrkArgs a;
  if (crk_common(c, a, "crk_compute_volume")) return 1;
  if (c->n == 0) return 0;
  if (c->ndim == 3) { if (launch_tiles(c, k_crk_volume<3>, a, "k_crk_volume", LW, (size_t)PosRing<3>::WARPB)) return 1; }
  else              { if (launch_tiles(c, k_crk_volume<2>, a, "k_crk_volume", LW, (size_t)PosRing<2>::WARPB)) return 1; }
  c->have[S_VOLUME] = true;
  return 0;
}

int sphb200_crk_compute_corrections(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  CrkArgs a;
  if (crk_common(c, a, "crk_compute_corrections")) return 1;
  if (!c->have[S_VOLUME]) return sphb200_fail(c, "crk_compute_corrections: the volume is not on the device (call crk_compute_volume or upload it)");
  if (c->n == 0) return 0;
  if (gather_volume(c)) return 1;
  a.aux2 = c->crkAux;
  if (c->ndim == 3) { if (launch_light<3>(c, k_crk_corrections<3>, a, "k_crk_corrections")) return 1; }
  else              { if (launch_light<2>(c, k_crk_corrections<2>, a, "k_crk_corrections")) return 1; }
  c->have[S_RKCORR] = true;
  return 0;
}

int sphb200_crk_sum_mass_density(sphb200_ctx* c, double rhoMin, double rhoMax) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  CrkArgs a;
  if (crk_common(c, a, "crk_sum_mass_density")) return 1;
  if (!c->have[S_VOLUME] || !c->have[S_MASS]) return sphb200_fail(c, "crk_sum_mass_density: volume and mass must be on the device");
  if (c->n == 0) return 0;
  if (!c->rowsValid && sphb200_pack_rows(c)) return 1;        // the rows carry the masses
  if (gather_volume(c)) return 1;
  a.aux2 = c->crkAux;
  a.rhoMin = rhoMin; a.rhoMax = rhoMax;
  if (c->ndim == 3) { if (launch_light<3>(c, k_crk_sum_density<3>, a, "k_crk_sum_density")) return 1; }
  else              { if (launch_light<2>(c, k_crk_sum_density<2>, a, "k_crk_sum_density")) return 1; }
  c->have[S_RHO] = true;
  c->rowsValid = false;                                       // the rows carry rho as well
  return 0;
}

}  // extern "C"
