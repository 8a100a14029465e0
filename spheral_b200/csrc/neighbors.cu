// neighbors.cu -- K1 (Morton cell sort) and K2 (neighbour / pair build) of the SPH hot path.
//
// Replaces, with identical results:
//   Neighbor::updateNodes            TreeNeighbor.cc:370-455 (tree build)      -> cell keys + counting sort
//   ConnectivityMap::computeConnectivity  ConnectivityMap.cc:747-1152          -> k_neighbors (count, fill)
// The pair set is decided ONLY by the predicate of ConnectivityMap.cc:916-925
//        (Hi.rij)^2 <= kext^2  ||  (Hj.rij)^2 <= kext^2 ,  i != j
// evaluated with the reference's operation order and no FMA contraction (eta2_exact); the cell grid merely supplies
// a superset of candidates, exactly like the reference's tree walk.
#include "sphb200_internal.cuh"
#include <cmath>
#include <cstring>

namespace {

constexpr int RB = 256;   // threads for simple per-node kernels

template <int DIM> struct Fr;                     // FP32 pre-filter row: relpos, H, lo2, hi2
template <> struct Fr<3> { static constexpr int ROW = 12, R_H = 3, R_LO = 9, R_HI = 10; };
template <> struct Fr<2> { static constexpr int ROW = 8,  R_H = 2, R_LO = 5, R_HI = 6; };

// ---- K1a: bounding box + maximum per-axis kernel extent ---------------------------------------------------------
// extent_a(i) = kext*sqrt((H^-2)_aa): half width along axis a of node i's ellipsoid |H r| <= kext
// (the role of Neighbor::HExtent, NeighborInline.hh:52-64).
template <int DIM> __device__ __forceinline__ void h_extent(const double* H, double kext, double* ext) {
  if (DIM == 3) {
    const double det = sym_det<3>(H);
    const double ixx = (H[3]*H[5] - H[4]*H[4])/det, ixy = (H[2]*H[4] - H[1]*H[5])/det, ixz = (H[1]*H[4] - H[2]*H[3])/det;
    const double iyy = (H[0]*H[5] - H[2]*H[2])/det, iyz = (H[1]*H[2] - H[0]*H[4])/det, izz = (H[0]*H[3] - H[1]*H[1])/det;
    ext[0] = kext*sqrt(ixx*ixx + ixy*ixy + ixz*ixz);
    ext[1] = kext*sqrt(ixy*ixy + iyy*iyy + iyz*iyz);
    ext[2] = kext*sqrt(ixz*ixz + iyz*iyz + izz*izz);
  } else {
    const double det = sym_det<2>(H);
    const double ixx = H[2]/det, ixy = -H[1]/det, iyy = H[0]/det;
    ext[0] = kext*sqrt(ixx*ixx + ixy*ixy);
    ext[1] = kext*sqrt(ixy*ixy + iyy*iyy);
  }
}

__device__ __forceinline__ double warp_min(double v) { for (int d = 16; d; d >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, d)); return v; }
__device__ __forceinline__ double warp_max(double v) { for (int d = 16; d; d >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, d)); return v; }

// partial[blk*9 + (0..2 lo, 3..5 hi, 6..8 ext)]
template <int DIM>
__global__ void __launch_bounds__(RB) k_bbox(const double* __restrict__ pos, const double* __restrict__ H, size_t n, double kext,
                                             double* __restrict__ partial) {
  constexpr int NS = Dm<DIM>::NS;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, ex[3] = {0, 0, 0};
  for (size_t i = (size_t)blockIdx.x*RB + threadIdx.x; i < n; i += (size_t)gridDim.x*RB) {
    double h[NS], e[DIM];
#pragma unroll
    for (int k = 0; k < NS; ++k) h[k] = H[i*NS + k];
    h_extent<DIM>(h, kext, e);
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      const double x = pos[i*DIM + a];
      lo[a] = fmin(lo[a], x); hi[a] = fmax(hi[a], x); ex[a] = fmax(ex[a], e[a]);
    }
  }
  __shared__ double sm[RB/32][9];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int a = 0; a < 3; ++a) {
    const double l = warp_min(lo[a]), h2 = warp_max(hi[a]), e2 = warp_max(ex[a]);
    if (lane == 0) { sm[w][a] = l; sm[w][3 + a] = h2; sm[w][6 + a] = e2; }
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double v = sm[0][threadIdx.x];
    for (int k = 1; k < RB/32; ++k) v = (threadIdx.x < 3) ? fmin(v, sm[k][threadIdx.x]) : fmax(v, sm[k][threadIdx.x]);
    partial[blockIdx.x*9 + threadIdx.x] = v;
  }
}
__global__ void k_bbox_final(const double* __restrict__ partial, int nb, double* __restrict__ out) {
  if (threadIdx.x < 9) {
    double v = partial[threadIdx.x];
    for (int k = 1; k < nb; ++k) v = (threadIdx.x < 3) ? fmin(v, partial[k*9 + threadIdx.x]) : fmax(v, partial[k*9 + threadIdx.x]);
    out[threadIdx.x] = v;
  }
}

// ---- K1b: cell key + histogram --------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(RB) k_cell_count(const double* __restrict__ pos, size_t n, GridDev g,
                                                   uint32_t* __restrict__ keyOut, uint32_t* __restrict__ cellCount) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  uint32_t key = 0;
#pragma unroll
  for (int a = 0; a < DIM; ++a) key |= dilate(g, a, cell_coord(pos[i*DIM + a], g.lo[a], g.cs[a], g.nc[a]));
  keyOut[i] = key;
  atomicAdd(&cellCount[key], 1u);
}

// ---- K1c: scatter into cells, then order each cell by original index (deterministic layout) ------------------------------
__global__ void __launch_bounds__(RB) k_cell_scatter(const uint32_t* __restrict__ key, size_t n, uint32_t* __restrict__ cursor,
                                                     uint32_t* __restrict__ perm) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = atomicAdd(&cursor[key[i]], 1u);
  perm[s] = (uint32_t)i;
}
__global__ void __launch_bounds__(RB) k_cell_order(const uint32_t* __restrict__ cellStart, uint32_t tableSize, uint32_t* __restrict__ perm) {
  const uint32_t c = blockIdx.x*RB + threadIdx.x;
  if (c >= tableSize) return;
  const uint32_t b = cellStart[c], e = cellStart[c + 1];
  for (uint32_t a = b + 1; a < e; ++a) {          // insertion sort, cells hold O(10) nodes
    const uint32_t v = perm[a];
    uint32_t k = a;
    while (k > b && perm[k - 1] > v) { perm[k] = perm[k - 1]; --k; }
    perm[k] = v;
  }
}

// ---- K1d: gather the host-ordered fields into Morton-sorted 128-byte node rows --------------------------------------------
struct PackArgs {
  const double *pos, *vel, *H, *mass, *rho, *P, *omega, *cs, *DvDxQ, *fCl, *fCq;
  double *rows, *auxPneg, *auxSomr2, *auxDvDxQ, *auxfCl, *auxfCq;
  const uint32_t *perm, *keyApi;
  uint32_t* skey;
  size_t n;
  float* frows; GridDev g; double kext; double csmax;
};
template <int DIM>
__global__ void __launch_bounds__(RB) k_pack(PackArgs a) {
  using D = Dm<DIM>;
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= a.n) return;
  const size_t o = a.perm[s];
  double* r = a.rows + s*D::ROW;
#pragma unroll
  for (int k = 0; k < DIM; ++k) { r[D::R_POS + k] = a.pos[o*DIM + k]; r[D::R_VEL + k] = a.vel ? a.vel[o*DIM + k] : 0.0; }
#pragma unroll
  for (int k = 0; k < D::NS; ++k) r[D::R_H + k] = a.H[o*D::NS + k];
  const double m = a.mass ? a.mass[o] : 0.0, rho = a.rho ? a.rho[o] : 1.0, P = a.P ? a.P[o] : 0.0;
  const double om = a.omega ? a.omega[o] : 1.0, cs = a.cs ? a.cs[o] : 0.0;
  const double safeOmega = om/(om*om + 1.0e-30);                 // safeInv, Utilities/safeInv.hh:13-19 (SPH.cc:310)
  r[D::R_M] = m; r[D::R_RHO] = rho; r[D::R_CS] = cs;
  r[D::R_PRHO] = safeOmega*P/(rho*rho);                          // SPH.cc:425 with Peff == P
  if (DIM == 2) r[11] = 0.0;
  if (a.auxPneg) { a.auxPneg[s] = (P < 0.0 ? -P : 0.0); a.auxSomr2[s] = safeOmega/(rho*rho); }
  if (a.auxDvDxQ) {
#pragma unroll
    for (int k = 0; k < D::NT; ++k) a.auxDvDxQ[s*D::NT + k] = a.DvDxQ[o*D::NT + k];
  }
  if (a.auxfCl) { a.auxfCl[s] = a.fCl[o]; a.auxfCq[s] = a.fCq[o]; }
  if (a.skey) a.skey[s] = a.keyApi[o];
  if (a.frows) {
    // FP32 pre-filter row: position relative to the node's cell corner, H, and the squared thresholds of the error band.
    // |eta_f32 - eta| <= B = ||H||_F * csmax * 2^-17 for candidates in the 3^DIM stencil (DESIGN.md "K2 error band"), so
    //   eta_f32^2 <= (kext-B)^2  => certainly inside ;  eta_f32^2 > (kext+B)^2  => certainly outside.
    using F = Fr<DIM>;
    float* f = a.frows + s*F::ROW;
    double hf = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      const double x = a.pos[o*DIM + k];
      const int c = cell_coord(x, a.g.lo[k], a.g.cs[k], a.g.nc[k]);
      f[k] = (float)(x - (a.g.lo[k] + (double)c*a.g.cs[k]));
    }
#pragma unroll
    for (int k = 0; k < D::NS; ++k) { const double h = a.H[o*D::NS + k]; f[F::R_H + k] = (float)h; hf += h*h; }
    if (DIM == 3) hf += a.H[o*6 + 1]*a.H[o*6 + 1] + a.H[o*6 + 2]*a.H[o*6 + 2] + a.H[o*6 + 4]*a.H[o*6 + 4];
    else hf += a.H[o*3 + 1]*a.H[o*3 + 1];
    const double B = sqrt(hf)*a.csmax*7.62939453125e-06;       // 2^-17
    const double lo = fmax(a.kext - B, 0.0), hi = a.kext + B;
    f[F::R_LO] = __double2float_rd(lo*lo*(1.0 - 1.0e-6));
    f[F::R_HI] = __double2float_ru(hi*hi*(1.0 + 1.0e-6));
    f[F::ROW - 1] = 0.f;
  }
}

// ---- K2: neighbour build ----------------------------------------------------------------------------------------------------
// One warp per tile of 32 consecutive Morton-sorted nodes; lane <-> node i.  The warp walks the union of the 3^DIM cell
// stencils of the distinct cells its nodes live in ("candidates", visited in a fixed order); every candidate j is read once
// (uniform address -> broadcast) and tested by all lanes.
//   k_nbr_count_candidates : integer walk, number of candidates per tile (sizes the hit-mask buffer)
//   k_nbr_test             : the predicate, once per (i, candidate): FP32 pre-filter with a rigorous error band, exact FP64
//                            evaluation (reference operation order, no FMA) for the rare in-band cases; emits one hit bit
//                            per (lane, candidate) plus the per-node counts
//   k_nbr_fill             : integer walk that expands the hit masks into the sliced-ELL lists (warp-shuffle lookups)
struct NbrArgs {
  const double* rows; const float* frows; const uint32_t* perm; const uint32_t* skey; const uint32_t* cellStart;
  size_t n; uint32_t nInt; double kext2; GridDev g;
  uint32_t* nbrCount; uint32_t* tileRows; const unsigned long long* tileOff; uint32_t* nbr;
  uint32_t* tileWords; const unsigned long long* maskOff; uint32_t* mask;
  unsigned long long* counters;
};


// Per-warp candidate walk shared by the three kernels: calls f(jb, je, sx, sy, sz) for every stencil cell, warp-uniformly.
template <int DIM, typename F>
__device__ __forceinline__ void walk_cells(const GridDev& g, const uint32_t* __restrict__ cellStart, unsigned leaders,
                                           const int* ci, F&& f) {
  for (unsigned lm = leaders; lm; lm &= lm - 1) {
    const int L = __ffs(lm) - 1;
    int lc[3];
    lc[0] = __shfl_sync(0xffffffffu, ci[0], L); lc[1] = __shfl_sync(0xffffffffu, ci[1], L); lc[2] = __shfl_sync(0xffffffffu, ci[2], L);
    const int zlo = (DIM == 3) ? -1 : 0, zhi = (DIM == 3) ? 1 : 0;
    for (int dz = zlo; dz <= zhi; ++dz) {
      const int sz = lc[2] + dz;
      if (DIM == 3 && (sz < 0 || sz >= g.nc[2])) continue;
      for (int dy = -1; dy <= 1; ++dy) {
        const int sy = lc[1] + dy;
        if (sy < 0 || sy >= g.nc[1]) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int sx = lc[0] + dx;
          if (sx < 0 || sx >= g.nc[0]) continue;
          // skip cells already visited through an earlier leader's stencil
          bool seen = false;
          for (unsigned pm = leaders & ((1u << L) - 1u); pm && !seen; pm &= pm - 1) {
            const int P = __ffs(pm) - 1;
            const int px = __shfl_sync(0xffffffffu, ci[0], P), py = __shfl_sync(0xffffffffu, ci[1], P), pz = __shfl_sync(0xffffffffu, ci[2], P);
            seen = (abs(px - sx) <= 1) && (abs(py - sy) <= 1) && (DIM == 2 || abs(pz - sz) <= 1);
          }
          if (seen) continue;
          uint32_t key = dilate(g, 0, sx) | dilate(g, 1, sy);
          if (DIM == 3) key |= dilate(g, 2, sz);
          const uint32_t jb = cellStart[key], je = cellStart[key + 1];
          if (je > jb) f(jb, je, sx, sy, sz);
        }
      }
    }
  }
}

// common per-lane prologue: identity, cell coordinates, leaders
template <int DIM>
__device__ __forceinline__ bool tile_prologue(const NbrArgs& a, size_t& tile, int& lane, size_t& i, bool& inRange, bool& active,
                                              uint32_t& origi, int* ci, unsigned& leaders) {
  using D = Dm<DIM>;
  lane = threadIdx.x & 31;
  tile = (size_t)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile*SPHB200_TILE >= a.n) return false;
  i = tile*SPHB200_TILE + lane;
  inRange = i < a.n;
  origi = inRange ? a.perm[i] : 0xffffffffu;
  active = inRange && origi < a.nInt;
  ci[0] = ci[1] = ci[2] = 0;
  uint32_t keyi = 0xffffffffu;
  if (inRange) {
    const double* r = a.rows + i*D::ROW;
#pragma unroll
    for (int k = 0; k < DIM; ++k) ci[k] = cell_coord(r[D::R_POS + k], a.g.lo[k], a.g.cs[k], a.g.nc[k]);
    keyi = a.skey[i];
  }
  const unsigned actMask = __ballot_sync(0xffffffffu, active);
  const unsigned same = __match_any_sync(0xffffffffu, active ? keyi : (0x80000000u | lane));
  const bool leader = active && ((__ffs(same & actMask) - 1) == lane);
  leaders = __ballot_sync(0xffffffffu, leader);
  return true;
}

template <int DIM>
__global__ void __launch_bounds__(128) k_nbr_count_candidates(NbrArgs a) {
  size_t tile, i; int lane, ci[3]; bool inRange, active; uint32_t origi; unsigned leaders;
  if (!tile_prologue<DIM>(a, tile, lane, i, inRange, active, origi, ci, leaders)) return;
  uint32_t T = 0;
  walk_cells<DIM>(a.g, a.cellStart, leaders, ci, [&](uint32_t jb, uint32_t je, int, int, int) { T += je - jb; });
  if (lane == 0) a.tileWords[tile] = (T + 31u)/32u;
}

// exact predicate (ConnectivityMap.cc:912-925) from the FP64 rows
template <int DIM>
__device__ __noinline__ bool exact_pair(const double* __restrict__ rows, size_t i, size_t j, double kext2) {
  using D = Dm<DIM>;
  const double* pi = rows + i*D::ROW; const double* pj = rows + j*D::ROW;
  double rij[DIM], Hi[D::NS], Hj[D::NS];
#pragma unroll
  for (int k = 0; k < DIM; ++k) rij[k] = __dadd_rn(pi[D::R_POS + k], -pj[D::R_POS + k]);
#pragma unroll
  for (int k = 0; k < D::NS; ++k) { Hi[k] = pi[D::R_H + k]; Hj[k] = pj[D::R_H + k]; }
  return eta2_exact<DIM>(Hi, rij) <= kext2 || eta2_exact<DIM>(Hj, rij) <= kext2;
}

template <int DIM> __device__ __forceinline__ float eta2_f32(const float* H, const float* r) {
  if (DIM == 3) {
    const float ex = fmaf(H[2], r[2], fmaf(H[1], r[1], H[0]*r[0]));
    const float ey = fmaf(H[4], r[2], fmaf(H[3], r[1], H[1]*r[0]));
    const float ez = fmaf(H[5], r[2], fmaf(H[4], r[1], H[2]*r[0]));
    return fmaf(ez, ez, fmaf(ey, ey, ex*ex));
  } else {
    const float ex = fmaf(H[1], r[1], H[0]*r[0]);
    const float ey = fmaf(H[2], r[1], H[1]*r[0]);
    return fmaf(ey, ey, ex*ex);
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) k_nbr_test(NbrArgs a) {
  using F = Fr<DIM>;
  size_t tile, i; int lane, ci[3]; bool inRange, active; uint32_t origi; unsigned leaders;
  if (!tile_prologue<DIM>(a, tile, lane, i, inRange, active, origi, ci, leaders)) return;
  float reli[DIM], Hi[Dm<DIM>::NS], lo2i = 0.f, hi2i = 0.f;
  if (inRange) {
    const float* fr = a.frows + i*F::ROW;
#pragma unroll
    for (int k = 0; k < DIM; ++k) reli[k] = fr[k];
#pragma unroll
    for (int k = 0; k < Dm<DIM>::NS; ++k) Hi[k] = fr[F::R_H + k];
    lo2i = fr[F::R_LO]; hi2i = fr[F::R_HI];
  } else {
#pragma unroll
    for (int k = 0; k < DIM; ++k) reli[k] = 0.f;
#pragma unroll
    for (int k = 0; k < Dm<DIM>::NS; ++k) Hi[k] = 0.f;
  }
  const float csf[3] = {(float)a.g.cs[0], (float)a.g.cs[1], (float)a.g.cs[2]};
  uint32_t cnt = 0, word = 0, t = 0;
  uint32_t* mrow = a.mask + a.maskOff[tile] + lane;

  walk_cells<DIM>(a.g, a.cellStart, leaders, ci, [&](uint32_t jb, uint32_t je, int sx, int sy, int sz) {
    // the error band of the FP32 filter is derived for candidates in the lane's own 3^DIM stencil; anything farther is a
    // certain miss because every cell is at least one kernel extent wide
    const int kx = ci[0] - sx, ky = ci[1] - sy, kz = (DIM == 3) ? ci[2] - sz : 0;
    const bool near = active && abs(kx) <= 1 && abs(ky) <= 1 && abs(kz) <= 1;
    float bse[DIM];
    bse[0] = fmaf((float)kx, csf[0], reli[0]);
    bse[1] = fmaf((float)ky, csf[1], reli[1]);
    if (DIM == 3) bse[2] = fmaf((float)kz, csf[2], reli[2]);
    for (uint32_t j = jb; j < je; ++j) {
      const float4* fj = reinterpret_cast<const float4*>(a.frows + (size_t)j*F::ROW);   // uniform address: broadcast
      float rw[F::ROW];
#pragma unroll
      for (int q = 0; q < F::ROW/4; ++q) { const float4 v = __ldg(fj + q); rw[4*q] = v.x; rw[4*q + 1] = v.y; rw[4*q + 2] = v.z; rw[4*q + 3] = v.w; }
      float r[DIM];
#pragma unroll
      for (int k = 0; k < DIM; ++k) r[k] = bse[k] - rw[k];
      const float e2i = eta2_f32<DIM>(Hi, r);
      const float e2j = eta2_f32<DIM>(rw + F::R_H, r);
      const bool cand = near && (j != (uint32_t)i);
      bool hit = cand && (e2i <= lo2i || e2j <= rw[F::R_LO]);
      const bool amb = cand && !hit && !(e2i > hi2i && e2j > rw[F::R_HI]);
      if (amb) hit = exact_pair<DIM>(a.rows, i, j, a.kext2);
      word |= (hit ? 1u : 0u) << (t & 31u);
      cnt += hit ? 1u : 0u;
      ++t;
      if ((t & 31u) == 0u) { mrow[(size_t)((t >> 5) - 1u)*SPHB200_TILE] = word; word = 0; }
    }
  });
  if (t & 31u) mrow[(size_t)(t >> 5)*SPHB200_TILE] = word;

  if (inRange) a.nbrCount[i] = cnt;
  uint32_t mx = cnt;
  unsigned long long sAll = cnt;
  for (int d = 16; d; d >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    sAll += __shfl_xor_sync(0xffffffffu, sAll, d);
  }
  if (lane == 0) {
    a.tileRows[tile] = mx;
    if (sAll) atomicAdd(&a.counters[1], sAll);
  }
}

template <int DIM>
__global__ void __launch_bounds__(128) k_nbr_fill(NbrArgs a) {
  size_t tile, i; int lane, ci[3]; bool inRange, active; uint32_t origi; unsigned leaders;
  if (!tile_prologue<DIM>(a, tile, lane, i, inRange, active, origi, ci, leaders)) return;
  const uint32_t* mrow = a.mask + a.maskOff[tile] + lane;
  uint32_t* out = a.nbr + a.tileOff[tile] + lane;
  uint32_t cnt = 0, hi = 0, t = 0;
  uint32_t candJ = 0, candO = 0;           // lane L holds candidate (32*w + L) of the current chunk
  auto flush = [&](uint32_t w) {
    uint32_t m = mrow[(size_t)w*SPHB200_TILE];
    while (__any_sync(0xffffffffu, m != 0u)) {
      const int b = m ? (__ffs(m) - 1) : 0;
      const uint32_t j = __shfl_sync(0xffffffffu, candJ, b);
      const uint32_t oj = __shfl_sync(0xffffffffu, candO, b);
      if (m) {
        const uint32_t up = (oj > origi) ? 1u : 0u;
        out[(size_t)cnt*SPHB200_TILE] = j | (up << 31);
        ++cnt; hi += up;
        m &= m - 1u;
      }
    }
  };
  walk_cells<DIM>(a.g, a.cellStart, leaders, ci, [&](uint32_t jb, uint32_t je, int, int, int) {
    uint32_t j = jb;
    while (j < je) {
      const uint32_t pos = t & 31u;                         // first free lane of the chunk
      const uint32_t take = min(32u - pos, je - j);
      if ((uint32_t)lane >= pos && (uint32_t)lane < pos + take) { candJ = j + ((uint32_t)lane - pos); candO = a.perm[candJ]; }
      j += take; t += take;
      if ((t & 31u) == 0u) flush((t >> 5) - 1u);
    }
  });
  if (t & 31u) flush(t >> 5);
  unsigned long long sHi = hi;
  for (int d = 16; d; d >>= 1) sHi += __shfl_xor_sync(0xffffffffu, sHi, d);
  if (lane == 0 && sHi) atomicAdd(&a.counters[0], sHi);
}

}  // namespace

// ---- host drivers --------------------------------------------------------------------------------------------------------------

template <typename T> int sphb200_ensure(sphb200_ctx* c, T*& p, size_t& cap, size_t need) {
  if (need <= cap && p) return 0;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  const size_t newCap = need + need/16 + 64;
  CU_CHECK(c, cudaMalloc((void**)&p, newCap*sizeof(T)));
  cap = newCap;
  return 0;
}
template int sphb200_ensure<uint32_t>(sphb200_ctx*, uint32_t*&, size_t&, size_t);
template int sphb200_ensure<double>(sphb200_ctx*, double*&, size_t&, size_t);
template int sphb200_ensure<float>(sphb200_ctx*, float*&, size_t&, size_t);

static int build_grid(sphb200_ctx* c, const double* bb /*lo3 hi3 ext3*/) {
  GridDev& g = c->grid;
  const int nd = c->ndim;
  const int maxBitsTotal = 27;               // 128 Mi table entries
  int totalBits;
  double cs[3] = {1, 1, 1};
  for (int a = 0; a < 3; ++a) { g.lo[a] = 0; g.cs[a] = 1; g.nc[a] = 1; g.bits[a] = 0; g.mask[a] = 0; }
  for (int a = 0; a < nd; ++a) {
    double e = bb[6 + a];
    if (!(e > 0.0) || !std::isfinite(e)) return sphb200_fail(c, "build_pairs: non-positive or non-finite kernel extent (bad H?)");
    cs[a] = e*(1.0 + 1.0e-9);
  }
  // grow cells (never shrink) until the Morton table fits
  for (int iter = 0; iter < 64; ++iter) {
    totalBits = 0;
    for (int a = 0; a < nd; ++a) {
      const double span = bb[3 + a] - bb[a];
      double m = std::floor(span/cs[a]) + 1.0;
      if (m < 1.0) m = 1.0;
      if (m > 32768.0) m = 32768.0;
      int nc = (int)m;
      if ((double)nc*cs[a] < span) cs[a] = span/nc*(1.0 + 1.0e-9);   // clamped axis: stretch the cells to cover the span
      int b = 0; while ((1 << b) < nc) ++b;
      g.nc[a] = nc; g.bits[a] = b; totalBits += b;
    }
    // keep the table within a small multiple of the node count as well
    int nodeBits = 0; while (((size_t)1 << nodeBits) < c->n) ++nodeBits;
    const int limit = std::min(maxBitsTotal, std::max(10, nodeBits + 1));
    if (totalBits <= limit) break;
    int amax = 0; for (int a = 1; a < nd; ++a) if (g.bits[a] > g.bits[amax]) amax = a;
    cs[amax] *= 2.0;
  }
  for (int a = 0; a < nd; ++a) { g.lo[a] = bb[a]; g.cs[a] = cs[a]; }
  // interleave: level by level, axes that still have bits
  int pos = 0;
  for (int l = 0; l < 16; ++l)
    for (int a = 0; a < nd; ++a)
      if (l < g.bits[a]) { g.bitpos[a][l] = (uint8_t)pos; g.mask[a] |= (1u << pos); ++pos; }
  g.tableSize = 1u << pos;
  return 0;
}

int sphb200_pack_rows(sphb200_ctx* c) {
  if (!c->sortValid) return sphb200_fail(c, "internal: pack_rows before sort");
  const bool tens = c->opt.epsTensile != 0.0;
  const bool needQ = (c->opt.Qkind == SPHB200_Q_LIMITED_MG) || c->opt.balsara;
  const bool mult = c->have[S_FCL] && c->have[S_FCQ];
  if (needQ && !c->have[S_DVDXQ]) return sphb200_fail(c, "evaluateDerivatives: the artificial viscosity needs the 'velocity gradient for artificial viscosity' field (DvDxQ) but none was uploaded");
  size_t cap;
  if (tens) { cap = c->auxPneg ? c->cap : 0; if (sphb200_ensure(c, c->auxPneg, cap, c->cap)) return 1; cap = c->auxSomr2 ? c->cap : 0; if (sphb200_ensure(c, c->auxSomr2, cap, c->cap)) return 1; }
  if (needQ) { cap = c->auxDvDxQ ? c->cap*9 : 0; if (sphb200_ensure(c, c->auxDvDxQ, cap, c->cap*9)) return 1; }
  if (mult) { cap = c->auxfCl ? c->cap : 0; if (sphb200_ensure(c, c->auxfCl, cap, c->cap)) return 1; cap = c->auxfCq ? c->cap : 0; if (sphb200_ensure(c, c->auxfCq, cap, c->cap)) return 1; }
  PackArgs a{};
  a.pos = c->api[S_POS]; a.vel = c->have[S_VEL] ? c->api[S_VEL] : nullptr; a.H = c->api[S_H];
  a.mass = c->have[S_MASS] ? c->api[S_MASS] : nullptr; a.rho = c->have[S_RHO] ? c->api[S_RHO] : nullptr;
  a.P = c->have[S_P] ? c->api[S_P] : nullptr; a.omega = c->have[S_OMEGA] ? c->api[S_OMEGA] : nullptr;
  a.cs = c->have[S_CS] ? c->api[S_CS] : nullptr;
  a.DvDxQ = needQ ? c->api[S_DVDXQ] : nullptr; a.fCl = mult ? c->api[S_FCL] : nullptr; a.fCq = mult ? c->api[S_FCQ] : nullptr;
  a.rows = c->rows; a.auxPneg = tens ? c->auxPneg : nullptr; a.auxSomr2 = tens ? c->auxSomr2 : nullptr;
  a.auxDvDxQ = needQ ? c->auxDvDxQ : nullptr; a.auxfCl = mult ? c->auxfCl : nullptr; a.auxfCq = mult ? c->auxfCq : nullptr;
  a.perm = c->perm; a.keyApi = c->cellKeyApi; a.skey = c->skey; a.n = c->n;
  { size_t fcap = c->frows ? c->frowsCap : 0;
    if (sphb200_ensure(c, c->frows, fcap, c->cap*(size_t)(c->ndim == 3 ? 12 : 8))) return 1;
    c->frowsCap = fcap; }
  a.frows = c->frows; a.g = c->grid;
  a.kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  a.csmax = std::max(c->grid.cs[0], std::max(c->grid.cs[1], c->ndim == 3 ? c->grid.cs[2] : 0.0));
  const unsigned nb = (unsigned)((c->n + RB - 1)/RB);
  if (c->ndim == 3) k_pack<3><<<nb, RB, 0, c->stream>>>(a); else k_pack<2><<<nb, RB, 0, c->stream>>>(a);
  KERNEL_CHECK(c, "k_pack");
  c->rowsValid = true;
  return 0;
}

int sphb200_sort_and_pack(sphb200_ctx* c) {
  const size_t n = c->n;
  if (!c->have[S_POS] || !c->have[S_H]) return sphb200_fail(c, "build_pairs: position and H must be uploaded first");
  if (!c->W.set) return sphb200_fail(c, "build_pairs: kernel table not set (need the kernel extent)");
  const double kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  // bbox + extents
  const int nbb = (int)std::min<size_t>(296, (n + RB - 1)/RB);
  if (c->ndim == 3) k_bbox<3><<<nbb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], n, kext, c->reduceBuf);
  else              k_bbox<2><<<nbb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], n, kext, c->reduceBuf);
  KERNEL_CHECK(c, "k_bbox");
  k_bbox_final<<<1, 32, 0, c->stream>>>(c->reduceBuf, nbb, c->reduceBuf + 296*9);
  KERNEL_CHECK(c, "k_bbox_final");
  CU_CHECK(c, cudaMemcpyAsync(c->reduceHost, c->reduceBuf + 296*9, 9*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int a = 0; a < c->ndim; ++a)
    if (!std::isfinite(c->reduceHost[a]) || !std::isfinite(c->reduceHost[3 + a]) || !std::isfinite(c->reduceHost[6 + a]))
      return sphb200_fail(c, "build_pairs: non-finite position or H");
  if (build_grid(c, c->reduceHost)) return 1;

  const size_t tbl = (size_t)c->grid.tableSize + 1;
  if (sphb200_ensure(c, c->cellStart, c->cellCap, tbl)) return 1;
  if (sphb200_ensure(c, c->cellCursor, c->cellCursorCap, tbl)) return 1;
  CU_CHECK(c, cudaMemsetAsync(c->cellStart, 0, tbl*sizeof(uint32_t), c->stream));
  const unsigned nb = (unsigned)((n + RB - 1)/RB);
  if (c->ndim == 3) k_cell_count<3><<<nb, RB, 0, c->stream>>>(c->api[S_POS], n, c->grid, c->cellKeyApi, c->cellStart);
  else              k_cell_count<2><<<nb, RB, 0, c->stream>>>(c->api[S_POS], n, c->grid, c->cellKeyApi, c->cellStart);
  KERNEL_CHECK(c, "k_cell_count");
  if (sphb200_scan_u32(c, c->cellStart, c->cellStart, c->grid.tableSize)) return 1;
  CU_CHECK(c, cudaMemcpyAsync(c->cellCursor, c->cellStart, (size_t)c->grid.tableSize*sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  k_cell_scatter<<<nb, RB, 0, c->stream>>>(c->cellKeyApi, n, c->cellCursor, c->perm);
  KERNEL_CHECK(c, "k_cell_scatter");
  k_cell_order<<<(c->grid.tableSize + RB - 1)/RB, RB, 0, c->stream>>>(c->cellStart, c->grid.tableSize, c->perm);
  KERNEL_CHECK(c, "k_cell_order");
  c->sortValid = true;
  return sphb200_pack_rows(c);
}

int sphb200_neighbors(sphb200_ctx* c) {
  const size_t n = c->n;
  c->nTiles = (n + SPHB200_TILE - 1)/SPHB200_TILE;
  NbrArgs a{};
  a.rows = c->rows; a.frows = c->frows; a.perm = c->perm; a.skey = c->skey; a.cellStart = c->cellStart;
  a.n = n; a.nInt = (uint32_t)c->nInt;
  const double kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  a.kext2 = kext*kext; a.g = c->grid;
  a.nbrCount = c->nbrCount; a.tileRows = c->tileRows; a.tileOff = c->tileOff; a.nbr = nullptr; a.counters = c->counters;
  a.tileWords = c->tileWords; a.maskOff = c->maskOff; a.mask = nullptr;
  CU_CHECK(c, cudaMemsetAsync(c->counters, 0, 2*sizeof(unsigned long long), c->stream));
  const int wpb = 4;
  const unsigned nb = (unsigned)((c->nTiles + wpb - 1)/wpb);
  // 1. candidates per tile -> hit-mask offsets
  if (c->ndim == 3) k_nbr_count_candidates<3><<<nb, wpb*32, 0, c->stream>>>(a); else k_nbr_count_candidates<2><<<nb, wpb*32, 0, c->stream>>>(a);
  KERNEL_CHECK(c, "k_nbr_count_candidates");
  if (sphb200_scan_tiles(c, c->tileWords, c->maskOff, c->nTiles)) return 1;
  CU_CHECK(c, cudaMemcpyAsync(c->countersHost + 3, c->maskOff + c->nTiles, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  if (sphb200_ensure(c, c->mask, c->maskCap, (size_t)c->countersHost[3] + 32)) return 1;
  a.mask = c->mask;
  // 2. the predicate, once per (node, candidate)
  if (c->ndim == 3) k_nbr_test<3><<<nb, wpb*32, 0, c->stream>>>(a); else k_nbr_test<2><<<nb, wpb*32, 0, c->stream>>>(a);
  KERNEL_CHECK(c, "k_nbr_test");
  if (sphb200_scan_tiles(c, c->tileRows, c->tileOff, c->nTiles)) return 1;
  CU_CHECK(c, cudaMemcpyAsync(c->countersHost + 2, c->tileOff + c->nTiles, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  c->nSlots = (size_t)c->countersHost[2];
  if (sphb200_ensure(c, c->nbr, c->nbrCap, c->nSlots + 32)) return 1;
  a.nbr = c->nbr;
  // 3. expand the masks into the sliced-ELL lists
  if (c->ndim == 3) k_nbr_fill<3><<<nb, wpb*32, 0, c->stream>>>(a); else k_nbr_fill<2><<<nb, wpb*32, 0, c->stream>>>(a);
  KERNEL_CHECK(c, "k_nbr_fill");
  CU_CHECK(c, cudaMemcpyAsync(c->countersHost, c->counters, 2*sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  c->npairs = (size_t)c->countersHost[0];
  c->nEdges = (size_t)c->countersHost[1];
  c->pairsValid = true;
  c->stats.directed_edges = c->nEdges;
  return 0;
}
