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
#include <cstdlib>

namespace {

constexpr int RB = 256;   // threads for simple per-node kernels

// FP32 pre-filter row of K2, 16 floats = four float4 (both dimensions):
//   q0 = { x-rel, y-rel, z-rel, r2lo }   position relative to the node's cell corner; |r|^2 <= r2lo  => certainly a neighbour
//   q1 = { r2hi, e2lo, e2hi, - }         |r|^2 > r2hi => certainly not (this node's side);  eta^2 thresholds of the error band
//   q2 = { Hxx, Hxy, Hxz, Hyy }  q3 = { Hyz, Hzz, -, - }      (2-D: Hxx, Hxy, Hyy in q2.xyz)
struct Fr { static constexpr int ROW = 16, R_R2LO = 3, R_R2HI = 4, R_E2LO = 5, R_E2HI = 6, R_H = 8; };

// Bounds on the extreme eigenvalues of a symmetric positive H: cyclic Jacobi rotations, then Gershgorin discs on the
// (similar) rotated matrix, so lmin/lmax are true bounds up to the round-off of the rotations (the caller adds a 1e-9
// relative margin).  Exact for diagonal H.  Needed only to turn the ellipsoid |H r| <= kext into an inner and an outer sphere.
__device__ __forceinline__ void jacobi_rot(double& app, double& aqq, double& apq, double& arp, double& arq) {
  if (apq == 0.0) return;
  const double theta = (aqq - app)/(2.0*apq);
  const double t = (theta < 0.0 ? -1.0 : 1.0)/(fabs(theta) + sqrt(theta*theta + 1.0));
  const double c = 1.0/sqrt(t*t + 1.0), sn = t*c;
  app -= t*apq; aqq += t*apq; apq = 0.0;
  const double rp = c*arp - sn*arq, rq = sn*arp + c*arq;
  arp = rp; arq = rq;
}
template <int DIM> __device__ __forceinline__ void sym_eig_bounds(const double* H, double& lmin, double& lmax) {
  if (DIM == 2) {
    const double m = 0.5*(H[0] + H[2]), d = sqrt(0.25*(H[0] - H[2])*(H[0] - H[2]) + H[1]*H[1]);
    lmin = m - d; lmax = m + d;
    return;
  }
  double a00 = H[0], a01 = H[1], a02 = H[2], a11 = H[3], a12 = H[4], a22 = H[5];
  for (int sweep = 0; sweep < 10; ++sweep) {
    if (fabs(a01) + fabs(a02) + fabs(a12) <= 1.0e-13*(fabs(a00) + fabs(a11) + fabs(a22))) break;
    jacobi_rot(a00, a11, a01, a02, a12);
    jacobi_rot(a00, a22, a02, a01, a12);
    jacobi_rot(a11, a22, a12, a01, a02);
  }
  const double off = fabs(a01) + fabs(a02) + fabs(a12);
  lmin = fmin(a00, fmin(a11, a22)) - off;
  lmax = fmax(a00, fmax(a11, a22)) + off;
}

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

// partial[blk*12 + (0..2 lo, 3..5 hi, 6..8 largest extent, 9..11 sum of extents)]
template <int DIM>
__global__ void __launch_bounds__(RB) k_bbox(const double* __restrict__ pos, const double* __restrict__ H, size_t n, double kext,
                                             double* __restrict__ partial) {
  constexpr int NS = Dm<DIM>::NS;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, ex[3] = {0, 0, 0}, es[3] = {0, 0, 0};
  for (size_t i = (size_t)blockIdx.x*RB + threadIdx.x; i < n; i += (size_t)gridDim.x*RB) {
    double h[NS], e[DIM];
#pragma unroll
    for (int k = 0; k < NS; ++k) h[k] = H[i*NS + k];
    h_extent<DIM>(h, kext, e);
#pragma unroll
    for (int a = 0; a < DIM; ++a) {
      const double x = pos[i*DIM + a];
      lo[a] = fmin(lo[a], x); hi[a] = fmax(hi[a], x); ex[a] = fmax(ex[a], e[a]); es[a] += e[a];
    }
  }
  __shared__ double sm[RB/32][12];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int a = 0; a < 3; ++a) {
    const double l = warp_min(lo[a]), h2 = warp_max(hi[a]), e2 = warp_max(ex[a]);
    double sum = es[a];
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, d);
    if (lane == 0) { sm[w][a] = l; sm[w][3 + a] = h2; sm[w][6 + a] = e2; sm[w][9 + a] = sum; }
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = sm[0][threadIdx.x];
    for (int k = 1; k < RB/32; ++k) v = (threadIdx.x < 3) ? fmin(v, sm[k][threadIdx.x]) : (threadIdx.x < 9 ? fmax(v, sm[k][threadIdx.x]) : v + sm[k][threadIdx.x]);
    partial[blockIdx.x*12 + threadIdx.x] = v;
  }
}
__global__ void k_bbox_final(const double* __restrict__ partial, int nb, double* __restrict__ out) {   // 12 warps, one per output
  const int q = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double v = (q < 3) ? 1e300 : (q < 9 ? -1e300 : 0.0);
  for (int k = lane; k < nb; k += 32) {
    const double p = partial[k*12 + q];
    v = (q < 3) ? fmin(v, p) : (q < 9 ? fmax(v, p) : v + p);
  }
  if (q < 3) v = warp_min(v); else if (q < 9) v = warp_max(v);
  else { for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d); }
  if (lane == 0) out[q] = v;
}

// ---- K1b: cell key + histogram --------------------------------------------------------------------------------------
// Ghost split (ghostBase = table size, 0 = off): ghost nodes (i >= nInt) are counted in a second copy of the cell table that follows
// the first, so the sorted order is [internal nodes in Morton order | ghost nodes in Morton order].  Ghosts build no lists; mixed
// among the internal nodes they leave lanes of the tiles next to a boundary idle in every pair loop (a tile costs what its longest
// list costs).  The key stored per node stays the plain cell key; a cell is then two ranges of the sorted order.
template <int DIM>
__global__ void __launch_bounds__(RB) k_cell_count(const double* __restrict__ pos, size_t n, GridDev g, const uint32_t* __restrict__ dilTab,
                                                   uint32_t* __restrict__ keyOut, uint32_t* __restrict__ cellCount, size_t nInt, uint32_t ghostBase) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  uint32_t key = 0;
#pragma unroll
  for (int a = 0; a < DIM; ++a) key |= dilTab[a*SPHB200_DIL + cell_coord(pos[i*DIM + a], g.lo[a], g.cs[a], g.nc[a])];
  keyOut[i] = key;
  atomicAdd(&cellCount[key + (i >= nInt ? ghostBase : 0u)], 1u);
}

// ---- K1c: scatter into cells, then order each cell by original index (deterministic layout) ------------------------------
__global__ void __launch_bounds__(RB) k_cell_scatter(const uint32_t* __restrict__ key, size_t n, uint32_t* __restrict__ cursor,
                                                     uint32_t* __restrict__ perm, size_t nInt, uint32_t ghostBase) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  const uint32_t s = atomicAdd(&cursor[key[i] + (i >= nInt ? ghostBase : 0u)], 1u);
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

// ---- K1c': stencil radius per cell (only when the grid is finer than the largest extent) ---------------------------------------
// A pair (i, j) is a neighbour pair if EITHER ellipsoid contains the other node, so a tile must walk as far as the extent of its
// own nodes (gather) and as far as the extent of any node that may contain them (scatter).  Every node with an extent of more
// than one cell writes its radius (in cells) over the block of cells it can reach; a tile then walks the largest radius
// recorded on the cells of its nodes.
template <int DIM>
__global__ void __launch_bounds__(RB) k_cell_reach(const double* __restrict__ pos, const double* __restrict__ H, size_t n, double kext, GridDev g,
                                                   const uint32_t* __restrict__ dilTab, int rmax, uint32_t* __restrict__ reach) {
  constexpr int NS = Dm<DIM>::NS;
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= n) return;
  double h[NS], e[DIM];
#pragma unroll
  for (int k = 0; k < NS; ++k) h[k] = H[i*NS + k];
  h_extent<DIM>(h, kext, e);
  int r = 1, ci[3] = {0, 0, 0};
#pragma unroll
  for (int a = 0; a < DIM; ++a) {
    r = max(r, (int)ceil(e[a]/g.cs[a]));
    ci[a] = cell_coord(pos[i*DIM + a], g.lo[a], g.cs[a], g.nc[a]);
  }
  if (r <= 1) return;
  r = min(r, rmax);
  const int zlo = (DIM == 3) ? max(ci[2] - r, 0) : 0, zhi = (DIM == 3) ? min(ci[2] + r, g.nc[2] - 1) : 0;
  for (int z = zlo; z <= zhi; ++z)
    for (int y = max(ci[1] - r, 0); y <= min(ci[1] + r, g.nc[1] - 1); ++y) {
      const uint32_t kyz = dilTab[SPHB200_DIL + y] | ((DIM == 3) ? dilTab[2*SPHB200_DIL + z] : 0u);
      for (int x = max(ci[0] - r, 0); x <= min(ci[0] + r, g.nc[0] - 1); ++x) {
        uint32_t* p = reach + (dilTab[x] | kyz);
        if (*p < (uint32_t)r) atomicMax(p, (uint32_t)r);
      }
    }
}

// candidate-volume proxy of the fine grid: sum over nodes of (2 r + 1)^DIM with r the radius recorded on the node's cell
template <int DIM>
__global__ void __launch_bounds__(RB) k_reach_cost(const uint32_t* __restrict__ keyApi, size_t n, const uint32_t* __restrict__ reach,
                                                   unsigned long long* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  unsigned long long v = 0;
  if (i < n) { const unsigned long long w = 2ull*max(1u, reach[keyApi[i]]) + 1ull; v = (DIM == 3) ? w*w*w : w*w; }
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0 && v) atomicAdd(out, v);
}

// ---- K1d: gather the host-ordered fields into Morton-sorted 128-byte node rows --------------------------------------------
struct PackArgs {
  const double *pos, *vel, *H, *mass, *rho, *P, *omega, *cs, *DvDxQ, *fCl, *fCq;
  double *rows, *aux2, *auxPneg, *auxSomr2, *auxDvDxQ, *auxfCl, *auxfCq;
  const uint32_t *perm, *keyApi;
  uint32_t* skey;
  size_t n;
  float* frows; GridDev g; double kext; double csmax;   // csmax: largest cell width x (3 rad + 4)/7 (FP32 error of k*cs + rel, |k| <= rad)
  int rawP;                       // CRKSPH: the row carries P itself (CRKSPH.cc:376 uses Pi + Pj), not safeInv(omega)*P/rho^2
  unsigned long long* aniso;      // set to 1 if any H is not a multiple of the identity (selects the isotropic pair loop)
  const uint32_t* invPerm; size_t first, count;   // RANGE mode: re-pack the rows of the nodes [first, first + count) only
};
// RANGE: only the nodes [first, first + count) of the host order (ghosts whose non-geometric fields have just landed, with positions
// and H -- hence the sort, the FP32 rows and the isotropy flag -- unchanged) are packed again, through the inverse permutation.
template <int DIM, bool RANGE>
__global__ void __launch_bounds__(RB) k_pack(PackArgs a) {
  using D = Dm<DIM>;
  const size_t t = (size_t)blockIdx.x*RB + threadIdx.x;
  if (t >= (RANGE ? a.count : a.n)) return;
  const size_t s = RANGE ? (size_t)a.invPerm[a.first + t] : t;
  const size_t o = RANGE ? a.first + t : (size_t)a.perm[s];
  double* r = a.rows + s*D::ROW;
#pragma unroll
  for (int k = 0; k < DIM; ++k) { r[D::R_POS + k] = a.pos[o*DIM + k]; r[D::R_VEL + k] = a.vel ? a.vel[o*DIM + k] : 0.0; }
  double Hn[D::NS];                                              // this node's H, read once
#pragma unroll
  for (int k = 0; k < D::NS; ++k) { Hn[k] = a.H[o*D::NS + k]; r[D::R_H + k] = Hn[k]; }
  if (!RANGE) { const bool iso = (DIM == 3) ? (Hn[1] == 0.0 && Hn[2] == 0.0 && Hn[4] == 0.0 && Hn[3] == Hn[0] && Hn[5] == Hn[0])
                                             : (Hn[1] == 0.0 && Hn[2] == Hn[0]);
    if (!iso) *a.aniso = 1ull; }
  const double m = a.mass ? a.mass[o] : 0.0, rho = a.rho ? a.rho[o] : 1.0, P = a.P ? a.P[o] : 0.0;
  const double om = a.omega ? a.omega[o] : 1.0, cs = a.cs ? a.cs[o] : 0.0;
  const double safeOmega = om/(om*om + 1.0e-30);                 // safeInv, Utilities/safeInv.hh:13-19 (SPH.cc:310)
  r[D::R_M] = m; r[D::R_RHO] = rho; r[D::R_CS] = cs;
  r[D::R_PRHO] = a.rawP ? P : safeOmega*P/(rho*rho);             // SPH.cc:425 with Peff == P
  if (DIM == 2) r[11] = 0.0;
  a.aux2[2*s] = sym_det<DIM>(Hn); a.aux2[2*s + 1] = 1.0/rho;
  if (a.auxPneg) { a.auxPneg[s] = (P < 0.0 ? -P : 0.0); a.auxSomr2[s] = safeOmega/(rho*rho); }
  if (a.auxDvDxQ) {
#pragma unroll
    for (int k = 0; k < D::NT; ++k) a.auxDvDxQ[s*D::NT + k] = a.DvDxQ[o*D::NT + k];
  }
  if (a.auxfCl) { a.auxfCl[s] = a.fCl[o]; a.auxfCq[s] = a.fCq[o]; }
  if (!RANGE && a.skey) a.skey[s] = a.keyApi[o];
  if (!RANGE && a.frows) {
    // FP32 pre-filter row (layout: struct Fr).  Error model (DESIGN.md "K2 error bands"): every relative-position component
    // is reconstructed in FP32 as k*cs + rel_i - rel_j with |error| <= 7*2^-24*cs, so for a pair at true distance <= R
    //   | |r|^2_f32 - |r|^2 | <= 2e-6*R*cs + 1e-6*R^2 + 1e-11*cs^2 =: margin(R)            (spheres)
    //   | eta_f32 - eta |     <= ||H||_F * cs * 2^-17 =: B                                     (ellipsoids, 3^DIM stencil)
    float* f = a.frows + s*Fr::ROW;
    double h[D::NS], hf = 0.0;
#pragma unroll
    for (int k = 0; k < D::NS; ++k) { h[k] = Hn[k]; hf += h[k]*h[k]; }
    if (DIM == 3) hf += h[1]*h[1] + h[2]*h[2] + h[4]*h[4]; else hf += h[1]*h[1];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float rel = 0.f;
      if (k < DIM) {
        const double x = a.pos[o*DIM + k];
        const int c = cell_coord(x, a.g.lo[k], a.g.cs[k], a.g.nc[k]);
        rel = (float)(x - (a.g.lo[k] + (double)c*a.g.cs[k]));
      }
      f[k] = rel;
    }
    const double cs = a.csmax;
    double lmin, lmax;
    sym_eig_bounds<DIM>(h, lmin, lmax);
    lmin -= 1.0e-9*lmax; lmax *= 1.0 + 1.0e-9;
    float r2lo = -1.f, r2hi = 3.0e38f;
    if (lmin > 0.0) {
      const double Rmax = a.kext/lmin, Rmin = a.kext/lmax;
      const double mhi = 2.0e-6*Rmax*cs + 1.0e-6*Rmax*Rmax + 1.0e-11*cs*cs;
      const double mlo = 2.0e-6*Rmin*cs + 1.0e-6*Rmin*Rmin + 1.0e-11*cs*cs;
      const double hi = (Rmax*Rmax + mhi)*(1.0 + 1.0e-7);
      if (hi < 3.0e38) r2hi = __double2float_ru(hi);
      if (Rmin > 1.0e-3*cs) r2lo = __double2float_rd((Rmin*Rmin - mlo)*(1.0 - 1.0e-7));
    }
    f[Fr::R_R2LO] = r2lo; f[Fr::R_R2HI] = r2hi;
    const double B = sqrt(hf)*cs*7.62939453125e-06;       // 2^-17
    const double lo = fmax(a.kext - B, 0.0), hi = a.kext + B;
    f[Fr::R_E2LO] = __double2float_rd(lo*lo*(1.0 - 1.0e-6));
    f[Fr::R_E2HI] = __double2float_ru(hi*hi*(1.0 + 1.0e-6));
    f[7] = __uint_as_float((uint32_t)o);                 // original index (bit pattern), for the pair orientation
#pragma unroll
    for (int k = 0; k < 6; ++k) f[Fr::R_H + k] = (k < D::NS) ? (float)h[k] : 0.f;
    f[14] = 0.f; f[15] = 0.f;
  }
}

// ---- K1d': the same rows, written the other way round (round 2) -----------------------------------------------------------------
// k_pack walks the SORTED slots: every field is gathered through perm (32 different lines per load instruction) and every thread
// writes its own 128-byte row eight bytes at a time (32 different lines per store instruction); the ncu profile of the 8 M step shows
// it LSU-bound at 2 TB/s of DRAM traffic (1.7 ms).  k_pack_scatter walks the nodes in HOST order instead: a warp reads the fields of
// 32 consecutive nodes (coalesced), builds their rows in shared memory and then writes each row to its sorted slot with one full
// 128-byte line per 8 lanes (64 bytes per 4 lanes for the FP32 rows).  Same values bit for bit (the arithmetic is shared).
// FROWS: also build the FP32 pre-filter row of the neighbour build (cell-relative position, Jacobi eigenvalue bounds of H, band
// thresholds).  A re-pack on a valid connectivity (values changed: sum density, EOS, a state update on the step-start pair lists,
// fields that landed after the build) leaves it out: only K2 reads those rows, and the next sort packs them again.
template <int DIM, bool FROWS = true>
__device__ __forceinline__ void pack_one(const PackArgs& a, size_t o, double* r, float* f, double* aux, bool& aniso) {
  using D = Dm<DIM>;
#pragma unroll
  for (int k = 0; k < DIM; ++k) { r[D::R_POS + k] = a.pos[o*DIM + k]; r[D::R_VEL + k] = a.vel ? a.vel[o*DIM + k] : 0.0; }
  double Hn[D::NS];
#pragma unroll
  for (int k = 0; k < D::NS; ++k) { Hn[k] = a.H[o*D::NS + k]; r[D::R_H + k] = Hn[k]; }
  aniso = !((DIM == 3) ? (Hn[1] == 0.0 && Hn[2] == 0.0 && Hn[4] == 0.0 && Hn[3] == Hn[0] && Hn[5] == Hn[0]) : (Hn[1] == 0.0 && Hn[2] == Hn[0]));
  const double m = a.mass ? a.mass[o] : 0.0, rho = a.rho ? a.rho[o] : 1.0, P = a.P ? a.P[o] : 0.0;
  const double om = a.omega ? a.omega[o] : 1.0, cs = a.cs ? a.cs[o] : 0.0;
  const double safeOmega = om/(om*om + 1.0e-30);                 // safeInv, Utilities/safeInv.hh:13-19 (SPH.cc:310)
  r[D::R_M] = m; r[D::R_RHO] = rho; r[D::R_CS] = cs;
  r[D::R_PRHO] = a.rawP ? P : safeOmega*P/(rho*rho);             // SPH.cc:425 with Peff == P
  if (DIM == 2) r[11] = 0.0;
  aux[0] = sym_det<DIM>(Hn); aux[1] = 1.0/rho; aux[2] = (P < 0.0 ? -P : 0.0); aux[3] = safeOmega/(rho*rho);
  if constexpr (!FROWS) return;
  // FP32 pre-filter row: see k_pack for the error model
  double hf = 0.0;
#pragma unroll
  for (int k = 0; k < D::NS; ++k) hf += Hn[k]*Hn[k];
  if (DIM == 3) hf += Hn[1]*Hn[1] + Hn[2]*Hn[2] + Hn[4]*Hn[4]; else hf += Hn[1]*Hn[1];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float rel = 0.f;
    if (k < DIM) {
      const double x = a.pos[o*DIM + k];
      const int c = cell_coord(x, a.g.lo[k], a.g.cs[k], a.g.nc[k]);
      rel = (float)(x - (a.g.lo[k] + (double)c*a.g.cs[k]));
    }
    f[k] = rel;
  }
  const double csw = a.csmax;
  double lmin, lmax;
  sym_eig_bounds<DIM>(Hn, lmin, lmax);
  lmin -= 1.0e-9*lmax; lmax *= 1.0 + 1.0e-9;
  float r2lo = -1.f, r2hi = 3.0e38f;
  if (lmin > 0.0) {
    const double Rmax = a.kext/lmin, Rmin = a.kext/lmax;
    const double mhi = 2.0e-6*Rmax*csw + 1.0e-6*Rmax*Rmax + 1.0e-11*csw*csw;
    const double mlo = 2.0e-6*Rmin*csw + 1.0e-6*Rmin*Rmin + 1.0e-11*csw*csw;
    const double hi = (Rmax*Rmax + mhi)*(1.0 + 1.0e-7);
    if (hi < 3.0e38) r2hi = __double2float_ru(hi);
    if (Rmin > 1.0e-3*csw) r2lo = __double2float_rd((Rmin*Rmin - mlo)*(1.0 - 1.0e-7));
  }
  f[Fr::R_R2LO] = r2lo; f[Fr::R_R2HI] = r2hi;
  const double B = sqrt(hf)*csw*7.62939453125e-06;       // 2^-17
  const double lo = fmax(a.kext - B, 0.0), hi2 = a.kext + B;
  f[Fr::R_E2LO] = __double2float_rd(lo*lo*(1.0 - 1.0e-6));
  f[Fr::R_E2HI] = __double2float_ru(hi2*hi2*(1.0 + 1.0e-6));
  f[7] = __uint_as_float((uint32_t)o);
#pragma unroll
  for (int k = 0; k < 6; ++k) f[Fr::R_H + k] = (k < D::NS) ? (float)Hn[k] : 0.f;
  f[14] = 0.f; f[15] = 0.f;
}

constexpr int PS_WARPS = 4;
template <int DIM, bool FROWS>
__global__ void __launch_bounds__(32*PS_WARPS) k_pack_scatter(PackArgs a) {
  using D = Dm<DIM>;
  constexpr int ROW = D::ROW, RS = ROW + 2;                      // +2 doubles: conflict-free 128-bit reads of a row's chunks
  __shared__ __align__(16) double srow[PS_WARPS][32*RS];
  __shared__ __align__(16) float sfr[PS_WARPS][32*(Fr::ROW + 4)];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t o0 = ((size_t)blockIdx.x*PS_WARPS + w)*32;
  if (o0 >= a.n) return;
  const size_t o = o0 + lane;
  const bool in = o < a.n;
  uint32_t s = 0;
  if (in) {
    double r[ROW], aux[4]; float f[Fr::ROW]; bool aniso;
    pack_one<DIM, FROWS>(a, o, r, f, aux, aniso);
    s = a.invPerm[o];
    if (aniso) *a.aniso = 1ull;
#pragma unroll
    for (int k = 0; k < ROW; ++k) srow[w][lane*RS + k] = r[k];
    if constexpr (FROWS) {
#pragma unroll
      for (int k = 0; k < Fr::ROW; ++k) sfr[w][lane*(Fr::ROW + 4) + k] = f[k];
    }
    a.aux2[2*(size_t)s] = aux[0]; a.aux2[2*(size_t)s + 1] = aux[1];
    if (a.auxPneg) { a.auxPneg[s] = aux[2]; a.auxSomr2[s] = aux[3]; }
    if (a.auxDvDxQ) {
#pragma unroll
      for (int k = 0; k < D::NT; ++k) a.auxDvDxQ[(size_t)s*D::NT + k] = a.DvDxQ[o*D::NT + k];
    }
    if (a.auxfCl) { a.auxfCl[s] = a.fCl[o]; a.auxfCq[s] = a.fCq[o]; }
    if (a.skey) a.skey[s] = a.keyApi[o];
  }
  __syncwarp();
  const int nhere = (int)min((size_t)32, a.n - o0);
  // node rows: CH 16-byte chunks each; 32/CH... lanes are dealt (row, chunk) pairs in order, so consecutive lanes fill one line
  constexpr int CH = ROW/2;
  for (int t = lane; t < 32*CH; t += 32) {
    const int rr = t/CH, ch = t - rr*CH;
    const uint32_t sr = __shfl_sync(0xffffffffu, s, rr);
    if (rr < nhere) {
      const double2 v = *reinterpret_cast<const double2*>(&srow[w][rr*RS + 2*ch]);
      *reinterpret_cast<double2*>(a.rows + (size_t)sr*ROW + 2*ch) = v;
    }
  }
  if (FROWS && a.frows) {
    for (int t = lane; t < 32*4; t += 32) {
      const int rr = t >> 2, ch = t & 3;
      const uint32_t sr = __shfl_sync(0xffffffffu, s, rr);
      if (rr < nhere) {
        const float4 v = *reinterpret_cast<const float4*>(&sfr[w][rr*(Fr::ROW + 4) + 4*ch]);
        *reinterpret_cast<float4*>(a.frows + (size_t)sr*Fr::ROW + 4*ch) = v;
      }
    }
  }
}

// ---- K2: neighbour build ----------------------------------------------------------------------------------------------------
// One warp per tile of 32 consecutive Morton-sorted nodes; lane <-> node i.  The candidates of a tile are the nodes of the
// union of the 3^DIM cell stencils of the distinct cells its nodes live in, visited in a fixed order, so every candidate j
// is fetched once per tile and tested by all lanes.
//   k_tile_runs : integer walk over the stencils; emits the tile's candidate runs (<= 32 consecutive sorted slots of one
//                 cell: first slot, length, cell coordinates)
//   k_nbr_build : the predicate, once per (i, candidate), and the lists.  The 64-byte FP32 rows of a run are loaded
//                 coalesced (one lane per candidate, one run ahead) and broadcast through shared memory.  Stage 1 is |r|^2
//                 against the nodes' inner and outer spheres (decisive for isotropic H); stage 2 the FP32 ellipsoid test
//                 with a rigorous error band; the rare in-band cases are decided by the exact FP64 evaluation (reference
//                 operation order, no FMA).  Hits are appended to the lane's column of a shared-memory list, the tile's
//                 block of the sliced-ELL array is allocated with one atomic and written as whole 128-byte rows.
// Both are launched back to back without a host round trip: buffer capacities come from the previous build and a kernel
// that would overflow one skips its tile; the host checks the totals once at the end and redoes the build with larger
// buffers if needed.
struct NbrArgs {
  const double* rows; const float* frows; const uint32_t* perm; const uint32_t* skey; const uint32_t* cellStart;
  const uint32_t* dilTab;
  uint32_t ghostBase;                // ghost split: the ghost nodes of cell k are cellStart[ghostBase + k] .. cellStart[ghostBase + k + 1] (0: off)
  size_t n; uint32_t nInt; double kext2; GridDev g;
  uint32_t* nbrCount; uint32_t* tileRows; unsigned long long* tileOff; uint32_t* nbr; unsigned long long nbrCap;
  uint4* runs; unsigned long long runsCap; uint32_t* tileRunStart; uint32_t* tileRunCount;
  int fine; GridDev gf;              // two-level walk: skey / cellStart / dilTab refer to the fine grid gf (cells of half the width)
  const uint32_t* cellReach;         // per cell key: stencil radius its tiles must walk (null: radius 1 everywhere)
  uint32_t* tileRadius;              // per tile: stencil radius used by k_tile_runs, read back by k_nbr_build
  unsigned long long* counters;      // [0] hits on ghost candidates  [1] directed edges  [2] run cursor  [3] list cursor  [4] longest list  [5] tiles with too many runs for the 16-bit list codes
};

// Per-warp candidate walk: calls f(jb, je, sx, sy, sz) for every non-empty stencil cell, warp-uniformly.
template <int DIM, typename F>
__device__ __forceinline__ void walk_cells(const GridDev& g, const uint32_t* __restrict__ dilTab,
                                           const uint32_t* __restrict__ cellStart, uint32_t ghostBase, unsigned leaders, const int* ci, int rad, F&& f) {
  for (unsigned lm = leaders; lm; lm &= lm - 1) {
    const int L = __ffs(lm) - 1;
    int lc[3];
    lc[0] = __shfl_sync(0xffffffffu, ci[0], L); lc[1] = __shfl_sync(0xffffffffu, ci[1], L); lc[2] = __shfl_sync(0xffffffffu, ci[2], L);
    const int zlo = (DIM == 3) ? -rad : 0, zhi = (DIM == 3) ? rad : 0;
    for (int dz = zlo; dz <= zhi; ++dz) {
      const int sz = lc[2] + dz;
      if (DIM == 3 && (sz < 0 || sz >= g.nc[2])) continue;
      for (int dy = -rad; dy <= rad; ++dy) {
        const int sy = lc[1] + dy;
        if (sy < 0 || sy >= g.nc[1]) continue;
        const uint32_t kyz = dilTab[SPHB200_DIL + sy] | ((DIM == 3) ? dilTab[2*SPHB200_DIL + sz] : 0u);
        for (int dx = -rad; dx <= rad; ++dx) {
          const int sx = lc[0] + dx;
          if (sx < 0 || sx >= g.nc[0]) continue;
          // skip cells already visited through an earlier leader's stencil
          bool seen = false;
          for (unsigned pm = leaders & ((1u << L) - 1u); pm && !seen; pm &= pm - 1) {
            const int P = __ffs(pm) - 1;
            const int px = __shfl_sync(0xffffffffu, ci[0], P), py = __shfl_sync(0xffffffffu, ci[1], P), pz = __shfl_sync(0xffffffffu, ci[2], P);
            seen = (abs(px - sx) <= rad) && (abs(py - sy) <= rad) && (DIM == 2 || abs(pz - sz) <= rad);
          }
          if (seen) continue;
          const uint32_t key = dilTab[sx] | kyz;
          const uint32_t jb = cellStart[key], je = cellStart[key + 1];
          if (je > jb) f(jb, je, sx, sy, sz);
          if (ghostBase) {
            const uint32_t gb = cellStart[ghostBase + key], ge = cellStart[ghostBase + key + 1];
            if (ge > gb) f(gb, ge, sx, sy, sz);
          }
        }
      }
    }
  }
}

// The same walk handing out the stencil cells' coordinates only (two-level walk: the caller looks the children up itself).
template <int DIM, typename F>
__device__ __forceinline__ void walk_cells_coarse(const GridDev& g, unsigned leaders, const int* ci, int rad, F&& f) {
  for (unsigned lm = leaders; lm; lm &= lm - 1) {
    const int L = __ffs(lm) - 1;
    int lc[3];
    lc[0] = __shfl_sync(0xffffffffu, ci[0], L); lc[1] = __shfl_sync(0xffffffffu, ci[1], L); lc[2] = __shfl_sync(0xffffffffu, ci[2], L);
    const int zlo = (DIM == 3) ? -rad : 0, zhi = (DIM == 3) ? rad : 0;
    for (int dz = zlo; dz <= zhi; ++dz) {
      const int sz = lc[2] + dz;
      if (DIM == 3 && (sz < 0 || sz >= g.nc[2])) continue;
      for (int dy = -rad; dy <= rad; ++dy) {
        const int sy = lc[1] + dy;
        if (sy < 0 || sy >= g.nc[1]) continue;
        for (int dx = -rad; dx <= rad; ++dx) {
          const int sx = lc[0] + dx;
          if (sx < 0 || sx >= g.nc[0]) continue;
          bool seen = false;
          for (unsigned pm = leaders & ((1u << L) - 1u); pm && !seen; pm &= pm - 1) {
            const int P = __ffs(pm) - 1;
            const int px = __shfl_sync(0xffffffffu, ci[0], P), py = __shfl_sync(0xffffffffu, ci[1], P), pz = __shfl_sync(0xffffffffu, ci[2], P);
            seen = (abs(px - sx) <= rad) && (abs(py - sy) <= rad) && (DIM == 2 || abs(pz - sz) <= rad);
          }
          if (!seen) f(sx, sy, sz);
        }
      }
    }
  }
}

// common per-lane prologue: identity, cell coordinates
template <int DIM>
__device__ __forceinline__ bool tile_prologue(const NbrArgs& a, size_t& tile, int& lane, size_t& i, bool& inRange, bool& active,
                                              uint32_t& origi, int* ci) {
  using D = Dm<DIM>;
  lane = threadIdx.x & 31;
  tile = (size_t)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tile*SPHB200_TILE >= a.n) return false;
  i = tile*SPHB200_TILE + lane;
  inRange = i < a.n;
  origi = inRange ? a.perm[i] : 0xffffffffu;
  active = inRange && origi < a.nInt;
  ci[0] = ci[1] = ci[2] = 0;
  if (inRange) {
    const double* r = a.rows + i*D::ROW;
#pragma unroll
    for (int k = 0; k < DIM; ++k) ci[k] = cell_coord(r[D::R_POS + k], a.g.lo[k], a.g.cs[k], a.g.nc[k]);
  }
  return true;
}

constexpr int RUN_CAP = 128;         // candidate runs per tile buffered in shared memory (typical: 40-70)

template <int DIM, bool FINE>
__global__ void __launch_bounds__(128) k_tile_runs(NbrArgs a) {
  __shared__ uint4 sruns[4][RUN_CAP];
  size_t tile, i; int lane, ci[3]; bool inRange, active; uint32_t origi;
  if (!tile_prologue<DIM>(a, tile, lane, i, inRange, active, origi, ci)) return;
  const int w = threadIdx.x >> 5;
  // leaders: first internal lane of every distinct cell of the tile
  // two-level walk: the sort key is the fine one; the coarse cell (leaders, stencil, FP32 geometry) is its parent
  const uint32_t keyf = inRange ? a.skey[i] : 0xffffffffu;
  const uint32_t keyi = (FINE && inRange) ? (keyf >> DIM) : keyf;
  const unsigned actMask = __ballot_sync(0xffffffffu, active);
  const unsigned same = __match_any_sync(0xffffffffu, active ? keyi : (0x80000000u | lane));
  const bool leader = active && ((__ffs(same & actMask) - 1) == lane);
  const unsigned leaders = __ballot_sync(0xffffffffu, leader);
  // stencil radius of this tile: the largest reach recorded for the cells of its nodes (k_cell_reach: a node's own extent in
  // cells AND the extents of the large nodes around it, so that gather and scatter neighbours are both inside the walk)
  int rad = 1;
  if (a.cellReach) {
    uint32_t r = (active && keyi != 0xffffffffu) ? max(1u, a.cellReach[keyi]) : 1u;
    r = __reduce_max_sync(0xffffffffu, r);
    rad = (int)r;
  }
  if (lane == 0 && a.tileRadius) a.tileRadius[tile] = (uint32_t)rad;
  uint32_t R = 0;
  auto emit = [&](uint32_t jb, uint32_t je, int sx, int sy, int sz, uint32_t lo, bool toShared) {
    // a cell with more than 32 nodes becomes several runs
    for (uint32_t b = jb; b < je; b += 32u) {
      const uint4 rec = make_uint4(b, min(32u, je - b), (uint32_t)sx | ((uint32_t)sy << 16), (uint32_t)sz);
      if (lane == 0) {
        if (toShared) { if (R < RUN_CAP) sruns[w][R] = rec; }
        else if (R >= RUN_CAP) a.runs[lo + R] = rec;
      }
      ++R;
    }
  };
  // Two-level walk (experimental): the nodes are sorted by a key with one more level, so the 2^DIM children of a stencil cell are
  // 2^DIM consecutive entries of cellStart and the cell is still one contiguous range.  Of every stencil cell only the children
  // inside the hull of the tile's own fine cells, grown by two fine cells, can hold a neighbour (|dx| <= extent <= 2 fine widths on
  // every axis, for the gather and the scatter side alike); consecutive needed children form one run.
  int flo[3] = {0, 0, 0}, fhi[3] = {0, 0, 0};
  if constexpr (FINE) {
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      const int f = inRange ? cell_coord(a.rows[i*Dm<DIM>::ROW + Dm<DIM>::R_POS + k], a.gf.lo[k], a.gf.cs[k], a.gf.nc[k]) : 0;
      flo[k] = __reduce_min_sync(0xffffffffu, active ? f : 0x7fffffff) - 2;
      fhi[k] = __reduce_max_sync(0xffffffffu, active ? f : -0x7fffffff) + 2;
    }
  }
  auto emit_fine = [&](int sx, int sy, int sz, uint32_t lo, bool toShared) {
    const uint32_t k0 = a.dilTab[2*sx] | a.dilTab[SPHB200_DIL + 2*sy] | ((DIM == 3) ? a.dilTab[2*SPHB200_DIL + 2*sz] : 0u);   // child (0,0,0)
    const int s3[3] = {sx, sy, sz};
    int first = -1;
    for (int ch = 0; ch <= (1 << DIM); ++ch) {
      bool need = ch < (1 << DIM);
      if (need) {
#pragma unroll
        for (int k = 0; k < DIM; ++k) {
          const int f = 2*s3[k] + ((ch >> a.gf.bitpos[k][0]) & 1);
          need = need && f >= flo[k] && f <= fhi[k];
        }
      }
      if (need && first < 0) first = ch;
      if (!need && first >= 0) {
        const uint32_t jb = a.cellStart[k0 + (uint32_t)first], je = a.cellStart[k0 + (uint32_t)ch];
        if (je > jb) emit(jb, je, sx, sy, sz, lo, toShared);
        first = -1;
      }
    }
  };
  // Box walk (the common case: stencil radius 1, the tile's cells within a small box): the lanes share out the cells of the box
  // [cmin - 1, cmax + 1], keep those inside the 3^DIM stencil of at least one leader and look their ranges up in parallel; the
  // runs come out in box order (z, y, x), a fixed function of the input.  ~150 instructions per tile against ~3.8 k for the
  // serial walk with its leader-by-leader duplicate test (k_tile_runs was 1.1 ms of the 8 M step: profiles/r02_notes.md).
  bool boxed = false;
  if constexpr (!FINE) {
    int bl[3] = {0, 0, 0}, be[3] = {1, 1, 1};
#pragma unroll
    for (int k = 0; k < DIM; ++k) {
      const int lo = __reduce_min_sync(0xffffffffu, leader ? ci[k] : 0x7fffffff), hi = __reduce_max_sync(0xffffffffu, leader ? ci[k] : -0x7fffffff);
      bl[k] = max(lo - 1, 0);
      be[k] = min(hi + 1, a.g.nc[k] - 1) - bl[k] + 1;
    }
    const int ncell = be[0]*be[1]*be[2];
    if (rad == 1 && leaders != 0u && ncell <= 128) {
      boxed = true;
      for (int base = 0; base < ncell; base += 32) {
        const int idx = base + lane;
        int sx = 0, sy = 0, sz = 0;
        bool want = idx < ncell;
        if (want) {
          sx = bl[0] + idx % be[0]; const int t = idx/be[0];
          sy = bl[1] + t % be[1]; sz = bl[2] + t/be[1];
        }
        bool inSt = false;                                   // within one cell of some leader's cell?
        for (unsigned lm = leaders; lm; lm &= lm - 1) {
          const int L = __ffs(lm) - 1;
          const int lx = __shfl_sync(0xffffffffu, ci[0], L), ly = __shfl_sync(0xffffffffu, ci[1], L), lz = __shfl_sync(0xffffffffu, ci[2], L);
          inSt = inSt || (abs(lx - sx) <= 1 && abs(ly - sy) <= 1 && (DIM == 2 || abs(lz - sz) <= 1));
        }
        uint32_t jb = 0, je = 0, gb = 0, ge = 0;             // the cell's internal nodes, then (ghost split) its ghost nodes
        if (want && inSt) {
          const uint32_t key = a.dilTab[sx] | a.dilTab[SPHB200_DIL + sy] | ((DIM == 3) ? a.dilTab[2*SPHB200_DIL + sz] : 0u);
          jb = a.cellStart[key]; je = a.cellStart[key + 1];
          if (a.ghostBase) { gb = a.cellStart[a.ghostBase + key]; ge = a.cellStart[a.ghostBase + key + 1]; }
        }
        const uint32_t nr = ((je - jb + 31u) >> 5) + ((ge - gb + 31u) >> 5);     // a range of more than 32 nodes becomes several runs
        uint32_t pre = nr;                                   // inclusive prefix sum over the lanes
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, pre, d); if (lane >= d) pre += t; }
        const uint32_t tot = __shfl_sync(0xffffffffu, pre, 31);
        uint32_t at = R + pre - nr;
        for (uint32_t b = jb; b < je; b += 32u, ++at)
          if (at < (uint32_t)RUN_CAP) sruns[w][at] = make_uint4(b, min(32u, je - b), (uint32_t)sx | ((uint32_t)sy << 16), (uint32_t)sz);
        for (uint32_t b = gb; b < ge; b += 32u, ++at)
          if (at < (uint32_t)RUN_CAP) sruns[w][at] = make_uint4(b, min(32u, ge - b), (uint32_t)sx | ((uint32_t)sy << 16), (uint32_t)sz);
        R += tot;
      }
      if (R > (uint32_t)RUN_CAP) { boxed = false; R = 0; }   // more runs than the table holds: the serial walk (it spills to global memory)
    }
  }
  if (!boxed) {
    if constexpr (FINE) walk_cells_coarse<DIM>(a.g, leaders, ci, rad, [&](int sx, int sy, int sz) { emit_fine(sx, sy, sz, 0u, true); });
    else walk_cells<DIM>(a.g, a.dilTab, a.cellStart, a.ghostBase, leaders, ci, rad, [&](uint32_t jb, uint32_t je, int sx, int sy, int sz) { emit(jb, je, sx, sy, sz, 0u, true); });
  }
  __syncwarp();
  const uint32_t Rtot = R;
  if (Rtot >= 2048u && lane == 0) atomicAdd(&a.counters[5], 1ull);      // list codes are run << 5 | candidate in 16 bits
  unsigned long long start = 0;
  if (lane == 0) start = atomicAdd(&a.counters[2], (unsigned long long)Rtot);
  start = __shfl_sync(0xffffffffu, start, 0);
  if (lane == 0) { a.tileRunStart[tile] = (uint32_t)start; a.tileRunCount[tile] = Rtot; }
  if (start + Rtot > a.runsCap) return;              // the host sees the cursor past the capacity and redoes the build
  for (uint32_t k = lane; k < min(Rtot, (uint32_t)RUN_CAP); k += 32) a.runs[start + k] = sruns[w][k];
  if (Rtot > RUN_CAP) {                              // rare: very ragged tile, walk again for the tail
    R = 0;
    if constexpr (FINE) walk_cells_coarse<DIM>(a.g, leaders, ci, rad, [&](int sx, int sy, int sz) { emit_fine(sx, sy, sz, (uint32_t)start, false); });
    else walk_cells<DIM>(a.g, a.dilTab, a.cellStart, a.ghostBase, leaders, ci, rad, [&](uint32_t jb, uint32_t je, int sx, int sy, int sz) { emit(jb, je, sx, sy, sz, (uint32_t)start, false); });
  }
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

constexpr int CROW = 20;             // floats per candidate row in shared memory (16 + 4 pad: conflict-free 128-bit stores)
constexpr int S1F = 5*32;            // floats of the stage-1 SoA copy of a run: x[32] y[32] z[32] r2lo[32] r2hi[32]

// Packed FP32 pairs (sm_100: FADD2 / FMUL2 / FFMA2 work on two floats held in a 64-bit register).  Stage 1 of the neighbour
// test is issue-bound (profiles/r01_notes.md); with the candidates of a run stored as SoA, one instruction advances two
// candidates.  Same operations and rounding as the scalar expressions (rn, no ftz): results are bit-identical.
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) { unsigned long long r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) { unsigned long long r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) { unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
#ifndef SPHB200_NBR_PACKED
#define SPHB200_NBR_PACKED 1
#endif
#ifndef SPHB200_NBR_PAD
#define SPHB200_NBR_PAD 0            // diagnostic: extra shared memory per warp (bytes), to measure how k_nbr_build reacts to occupancy
#endif
constexpr int NB_WARPS = 4;          // warps (tiles) per CTA of k_nbr_build
constexpr int JB_CAP = 128;          // runs per tile whose first slot is cached in shared memory for the flush

// Shared memory of k_nbr_build per warp: candidate rows of the current run (32*CROW floats), first slots of the runs
// (JB_CAP words) and the list under construction, `listRows` rows x 32 lanes of 16-bit codes (run << 5 | candidate).
// The lists are assembled [row][lane] (conflict-free for any per-lane row) and written out as whole 128-byte rows; a
// lane-private 4-byte global store per hit would cost a 32-byte L2 sector transaction each (profiles/r01_notes.md).
template <int DIM>
__global__ void __launch_bounds__(32*NB_WARPS) k_nbr_build(NbrArgs a, int listRows) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int w = threadIdx.x >> 5;
  const size_t perWarp = (size_t)32*CROW*4 + (size_t)S1F*4 + (size_t)JB_CAP*4 + (size_t)listRows*64 + SPHB200_NBR_PAD;
  float* const sc = reinterpret_cast<float*>(smemRaw + w*perWarp);
  float* const s1 = sc + 32*CROW;                   // stage-1 SoA copy of the current run
  uint32_t* const sjb = reinterpret_cast<uint32_t*>(s1 + S1F);
  unsigned short* const slist = reinterpret_cast<unsigned short*>(sjb + JB_CAP);

  size_t tile, i; int lane, ci[3]; bool inRange, active; uint32_t origi;
  if (!tile_prologue<DIM>(a, tile, lane, i, inRange, active, origi, ci)) return;
  const uint32_t R = a.tileRunCount[tile];
  const unsigned long long rs = a.tileRunStart[tile];
  const int rad = a.tileRadius ? (int)a.tileRadius[tile] : 1;
  if (rs + R > a.runsCap || R >= 2048u) {                        // capacity miss: leave a consistent, empty tile; the host redoes the build
    if (inRange) a.nbrCount[i] = 0;
    if (lane == 0) { a.tileRows[tile] = 0; a.tileOff[tile] = 0; }
    return;
  }
  float reli[3] = {0.f, 0.f, 0.f}, Hi[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, r2loi = -1.f, r2hii = 0.f, e2loi = 0.f, e2hii = 0.f;
  const float4* __restrict__ frows4 = reinterpret_cast<const float4*>(a.frows);
  if (inRange) {
    const float4 q0 = frows4[i*4], q1 = frows4[i*4 + 1], q2 = frows4[i*4 + 2], q3 = frows4[i*4 + 3];
    reli[0] = q0.x; reli[1] = q0.y; reli[2] = q0.z; r2loi = q0.w;
    r2hii = q1.x; e2loi = q1.y; e2hii = q1.z;
    if (DIM == 3) { Hi[0] = q2.x; Hi[1] = q2.y; Hi[2] = q2.z; Hi[3] = q2.w; Hi[4] = q3.x; Hi[5] = q3.y; }
    else { Hi[0] = q2.x; Hi[1] = q2.y; Hi[2] = q2.z; }
  }
  const float csf[3] = {(float)a.g.cs[0], (float)a.g.cs[1], (float)a.g.cs[2]};
  uint32_t cnt = 0, ghostHits = 0;
  unsigned short* lp = slist + lane;               // next free entry of this lane's column

  // software pipeline: the rows of run r+1 travel to registers while run r is tested out of shared memory
  const float4 FAR0 = make_float4(-1.0e30f, 0.f, 0.f, -1.f), FAR1 = make_float4(-1.f, 0.f, 0.f, 0.f);   // padding candidate: at infinity
  uint4 rec = (R > 0u) ? __ldg(a.runs + rs) : make_uint4(0u, 0u, 0u, 0u);
  float4 p0 = FAR0, p1 = FAR1, p2 = FAR1, p3 = FAR1;
  if ((uint32_t)lane < rec.y) { const size_t c = (size_t)(rec.x + lane)*4; p0 = __ldg(frows4 + c); p1 = __ldg(frows4 + c + 1); p2 = __ldg(frows4 + c + 2); p3 = __ldg(frows4 + c + 3); }
  for (uint32_t r = 0; r < R; ++r) {
    const uint32_t jb = rec.x, len = rec.y;
    const int kx = ci[0] - (int)(rec.z & 0xffffu), ky = ci[1] - (int)(rec.z >> 16), kz = (DIM == 3) ? ci[2] - (int)rec.w : 0;
    __syncwarp();                                                   // every lane is done with the previous run's rows
    { float4* d = reinterpret_cast<float4*>(sc + lane*CROW); d[0] = p0; d[1] = p1; d[2] = p2; d[3] = p3; }
#if SPHB200_NBR_PACKED
    s1[lane] = p0.x; s1[32 + lane] = p0.y; s1[64 + lane] = p0.z; s1[96 + lane] = p0.w; s1[128 + lane] = p1.x;
#endif
    if (lane == 0 && r < (uint32_t)JB_CAP) sjb[r] = jb;
    // ghost candidates of the run (original index >= nInt): a hit on one of them is a pair counted once, not twice
    const unsigned ghostWord = __ballot_sync(0xffffffffu, (uint32_t)lane < len && __float_as_uint(p1.w) >= a.nInt);
    __syncwarp();
    p0 = FAR0; p1 = FAR1; p2 = FAR1; p3 = FAR1;
    if (r + 1u < R) {
      rec = __ldg(a.runs + rs + r + 1u);
      if ((uint32_t)lane < rec.y) { const size_t c = (size_t)(rec.x + lane)*4; p0 = __ldg(frows4 + c); p1 = __ldg(frows4 + c + 1); p2 = __ldg(frows4 + c + 2); p3 = __ldg(frows4 + c + 3); }
    }
    // A candidate outside the lane's (2 rad + 1)^DIM stencil is a certain miss (rad cells cover the extent of this node and
    // of every node that can reach it, k_cell_reach) and the error bands are derived for candidates inside it: such lanes
    // (and ghost / padding lanes) see the run at infinity.
    const bool near = active && abs(kx) <= rad && abs(ky) <= rad && abs(kz) <= rad;
    float bse[3];
    bse[0] = near ? fmaf((float)kx, csf[0], reli[0]) : 1.0e30f;
    bse[1] = fmaf((float)ky, csf[1], reli[1]);
    bse[2] = (DIM == 3) ? fmaf((float)kz, csf[2], reli[2]) : 0.f;

    // stage 1, branch-free: |r|^2 against the inner spheres (certain hit) and the outer spheres (possible hit).  Rows past
    // the end of the run are padding candidates at infinity, so the loop runs in groups of four.
    uint32_t hitWord = 0, inWord = 0;
#if SPHB200_NBR_PACKED
    const unsigned long long bx2 = f2_pack(bse[0], bse[0]), by2 = f2_pack(bse[1], bse[1]), bz2 = f2_pack(bse[2], bse[2]);
    for (uint32_t c4 = 0; c4 < len; c4 += 4u) {
      const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(s1 + c4), Y = *reinterpret_cast<const ulonglong2*>(s1 + 32 + c4);   // broadcast
      const float4 LO = *reinterpret_cast<const float4*>(s1 + 96 + c4), HI = *reinterpret_cast<const float4*>(s1 + 128 + c4);
      const unsigned long long rxa = f2_sub(bx2, X.x), rxb = f2_sub(bx2, X.y), rya = f2_sub(by2, Y.x), ryb = f2_sub(by2, Y.y);
      unsigned long long qa = f2_fma(rya, rya, f2_mul(rxa, rxa)), qb = f2_fma(ryb, ryb, f2_mul(rxb, rxb));
      if (DIM == 3) {
        const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(s1 + 64 + c4);
        const unsigned long long rza = f2_sub(bz2, Z.x), rzb = f2_sub(bz2, Z.y);
        qa = f2_fma(rza, rza, qa); qb = f2_fma(rzb, rzb, qb);
      }
      float r2v[4];
      f2_unpack(qa, r2v[0], r2v[1]); f2_unpack(qb, r2v[2], r2v[3]);
      const float lo[4] = {LO.x, LO.y, LO.z, LO.w}, hi[4] = {HI.x, HI.y, HI.z, HI.w};
      uint32_t hn = 0, in = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r2v[u] <= fmaxf(r2loi, lo[u])) hn |= 1u << u;
        if (r2v[u] <= fmaxf(r2hii, hi[u])) in |= 1u << u;
      }
      hitWord |= hn << c4; inWord |= in << c4;
    }
#else
    for (uint32_t c4 = 0; c4 < len; c4 += 4u) {
      uint32_t hn = 0, in = 0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float4 q0 = *reinterpret_cast<const float4*>(sc + (c4 + u)*CROW);          // broadcast
        const float4 q1 = *reinterpret_cast<const float4*>(sc + (c4 + u)*CROW + 4);
        const float rx = bse[0] - q0.x, ry = bse[1] - q0.y, rz = bse[2] - q0.z;
        const float r2 = (DIM == 3) ? fmaf(rz, rz, fmaf(ry, ry, rx*rx)) : fmaf(ry, ry, rx*rx);
        if (r2 <= fmaxf(r2loi, q0.w)) hn |= 1u << u;
        if (r2 <= fmaxf(r2hii, q1.x)) in |= 1u << u;
      }
      hitWord |= hn << c4; inWord |= in << c4;
    }
#endif
    // stage 2: candidates inside an outer sphere but no inner one ask the ellipsoids (FP32 with error band, else exact)
    uint32_t amb = inWord & ~hitWord;
    for (unsigned au = __reduce_or_sync(0xffffffffu, amb); au; au &= au - 1u) {
      const uint32_t c = (uint32_t)(__ffs(au) - 1);
      if ((amb >> c) & 1u) {
        const float4 q0 = *reinterpret_cast<const float4*>(sc + c*CROW), q1 = *reinterpret_cast<const float4*>(sc + c*CROW + 4);
        const float4 q2 = *reinterpret_cast<const float4*>(sc + c*CROW + 8), q3 = *reinterpret_cast<const float4*>(sc + c*CROW + 12);
        float rv[3], Hj[6];
        rv[0] = bse[0] - q0.x; rv[1] = bse[1] - q0.y; rv[2] = bse[2] - q0.z;
        if (DIM == 3) { Hj[0] = q2.x; Hj[1] = q2.y; Hj[2] = q2.z; Hj[3] = q2.w; Hj[4] = q3.x; Hj[5] = q3.y; }
        else { Hj[0] = q2.x; Hj[1] = q2.y; Hj[2] = q2.z; }
        const float e2i = eta2_f32<DIM>(Hi, rv), e2j = eta2_f32<DIM>(Hj, rv);
        bool hit = (e2i <= e2loi) || (e2j <= q1.y);
        if (!hit && !(e2i > e2hii && e2j > q1.z)) hit = exact_pair<DIM>(a.rows, i, (size_t)jb + c, a.kext2);
        if (hit) hitWord |= 1u << c;
      }
    }
    const uint32_t self = (uint32_t)i - jb;                        // this node's own position in the run, if any
    if (self < len) hitWord &= ~(1u << self);
    cnt += __popc(hitWord);
    ghostHits += __popc(hitWord & ghostWord);
    // append the hits to this lane's column of the list; the bound is checked once per run (a list that does not fit
    // is reported through counters[4] and the host redoes the build with a larger staging area)
    if (cnt <= (uint32_t)listRows) {
      const uint32_t rbase = r << 5;
      for (uint32_t m = hitWord; m; m &= m - 1u) { *lp = (unsigned short)(rbase | (uint32_t)(__ffs(m) - 1)); lp += 32; }
    }
  }

  if (inRange) a.nbrCount[i] = cnt;
  uint32_t mx = cnt;
  unsigned long long sAll = cnt, sGhost = ghostHits;
  for (int d = 16; d; d >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    sAll += __shfl_xor_sync(0xffffffffu, sAll, d);
    sGhost += __shfl_xor_sync(0xffffffffu, sGhost, d);
  }
  // allocate the tile's block of the sliced-ELL array (placement order is irrelevant; the content is deterministic)
  unsigned long long off = 0;
  if (lane == 0) {
    off = atomicAdd(&a.counters[3], (unsigned long long)mx*SPHB200_TILE);
    a.tileRows[tile] = mx; a.tileOff[tile] = off;
    if (sAll) atomicAdd(&a.counters[1], sAll);
    if (sGhost) atomicAdd(&a.counters[0], sGhost);
    atomicMax(&a.counters[4], (unsigned long long)mx);          // longest list: sizes the staging (too small -> host redoes)
  }
  off = __shfl_sync(0xffffffffu, off, 0);
  if (off + (unsigned long long)mx*SPHB200_TILE > a.nbrCap || mx > (uint32_t)listRows) return;
  __syncwarp();
  uint32_t* out = a.nbr + off + lane;
  for (uint32_t k = 0; k < mx; ++k) {
    uint32_t slot = 0u;                                                       // padding entries point at slot 0 (never used)
    if (k < cnt) {                                                            // beyond a lane's own list the staging holds stale codes
      const uint32_t code = slist[k*32u + lane];
      const uint32_t rr = code >> 5;
      const uint32_t jb = (rr < (uint32_t)JB_CAP) ? sjb[rr] : __ldg(&a.runs[rs + rr].x);
      slot = jb + (code & 31u);
    }
    out[(size_t)k*SPHB200_TILE] = slot;
  }
}

// ---- K2, second version: flattened candidate stream (round 2) ------------------------------------------------------------------
// Same inputs (the tile's candidate runs of k_tile_runs), same three-stage predicate, same lists in the same order as k_nbr_build,
// restructured after the round-2 profile of the 8 M configuration (profiles/r02_notes.md): k_nbr_build spent its instructions on
// candidates that no node of the tile can reach (924 tested per node for 85 hits) and on per-run overhead (a warp-uniform stage-2
// loop over the union of the lanes' ambiguous candidates, a divergent list append per run).  Here
//   * positions live in ONE frame per tile (relative to the cell of the tile's first node), so a candidate is placed once, when its
//     run is loaded, and stage 1 needs no per-run set-up;
//   * the lane that loads a candidate tests it against the bounding box of the tile's nodes grown by the reach of either side
//     (|r| <= max(outer radius) is necessary for a pair) and only survivors enter a shared-memory ring; they are tested in full
//     chunks of 32, whatever run they came from;
//   * stage 2 is a per-lane loop over the lane's own ambiguous candidates.
// A tile whose nodes straddle a jump of the Morton curve (cells more than one apart) is handled in passes, one per cluster of
// lanes, so that every coordinate of a pass stays within four cell widths of its origin (the FP32 error model of k_pack).
// Only used with a stencil radius of 1 (cells at least as wide as every kernel extent).
constexpr int Q2_RING = 64;            // surviving candidates buffered per warp: two chunks of 32
constexpr int Q2_FIXED = 5*Q2_RING*4 + Q2_RING*32 + Q2_RING + JB_CAP*4;   // bytes per warp in front of the per-chunk records
constexpr int Q2_CHUNKB = 32*4 + 32*2; // per chunk of 32 survivors: one hit mask per lane + the survivors' list codes

// Hits are not appended one by one: a chunk leaves one 32-bit hit mask per lane (bit c <-> survivor c of the chunk) and the list
// codes of its survivors (run << 5 | position in the run); the lists are expanded from the masks when the tile's block of the
// sliced-ELL array is written, whole 128-byte rows at a time.  (The per-hit append of k_nbr_build is a divergent loop -- 17
// iterations per chunk for 4.5 hits per lane -- with a shared-memory load in its dependency chain: 26 % of the samples of the
// first version of this kernel.)
// STAGEH: the H tensors of the survivors are staged in the ring for stage 2.  When the previous build saw only isotropic H, stage 2
// is a rarity (a 1e-5-wide shell) and fetches its two rows from global memory instead: 16 of ~31 LSU wavefronts per run less.
template <int DIM, bool STAGEH>
#ifndef SPHB200_NB2_CTAS
#define SPHB200_NB2_CTAS 4
#endif
__global__ void __launch_bounds__(32*NB_WARPS, SPHB200_NB2_CTAS) k_nbr_build2(NbrArgs a, int maxChunks) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  const int w = threadIdx.x >> 5;
  const size_t perWarp = (size_t)Q2_FIXED + (size_t)maxChunks*Q2_CHUNKB;
  unsigned char* const wb = smemRaw + w*perWarp;
  float* const qx = reinterpret_cast<float*>(wb);                    // SoA of the ring: tile-frame position, inner / outer radius^2
  float* const qy = qx + Q2_RING; float* const qz = qy + Q2_RING; float* const qlo = qz + Q2_RING; float* const qhi = qlo + Q2_RING;
  float* const qrow = qhi + Q2_RING;                                  // stage-2 record per entry: H (6), e2lo, e2hi
  unsigned char* const qself = reinterpret_cast<unsigned char*>(qrow + 8*Q2_RING);    // bits 0-4: lane of this tile the candidate IS (0x1f with bit 5: none); bit 7: ghost
  uint32_t* const sjb = reinterpret_cast<uint32_t*>(qself + Q2_RING);
  uint32_t* const smask = sjb + JB_CAP;                               // [chunk][lane]
  unsigned short* const scode = reinterpret_cast<unsigned short*>(smask + 32*(size_t)maxChunks);   // [chunk][survivor]

  size_t tile, i; int lane, ci[3]; bool inRange, active; uint32_t origi;
  if (!tile_prologue<DIM>(a, tile, lane, i, inRange, active, origi, ci)) return;
  const uint32_t R = a.tileRunCount[tile];
  const unsigned long long rs = a.tileRunStart[tile];
  if (rs + R > a.runsCap || R >= 2048u) {                        // capacity miss: leave a consistent, empty tile; the host redoes the build
    if (inRange) a.nbrCount[i] = 0;
    if (lane == 0) { a.tileRows[tile] = 0; a.tileOff[tile] = 0; }
    return;
  }
  float reli[3] = {0.f, 0.f, 0.f}, Hi[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, r2loi = -1.f, r2hii = 0.f, e2loi = 0.f, e2hii = 0.f;
  const float4* __restrict__ frows4 = reinterpret_cast<const float4*>(a.frows);
  if (inRange) {
    const float4 q0 = frows4[i*4], q1 = frows4[i*4 + 1], q2 = frows4[i*4 + 2], q3 = frows4[i*4 + 3];
    reli[0] = q0.x; reli[1] = q0.y; reli[2] = q0.z; r2loi = q0.w;
    r2hii = q1.x; e2loi = q1.y; e2hii = q1.z;
    if (DIM == 3) { Hi[0] = q2.x; Hi[1] = q2.y; Hi[2] = q2.z; Hi[3] = q2.w; Hi[4] = q3.x; Hi[5] = q3.y; }
    else { Hi[0] = q2.x; Hi[1] = q2.y; Hi[2] = q2.z; }
  }
  const float csf[3] = {(float)a.g.cs[0], (float)a.g.cs[1], (float)a.g.cs[2]};
  uint32_t cnt = 0, ghostHits = 0;
  for (uint32_t r = lane; r < min(R, (uint32_t)JB_CAP); r += 32u) sjb[r] = __ldg(&a.runs[rs + r].x);
  const uint32_t tile0 = (uint32_t)(tile*SPHB200_TILE);
  const unsigned ltMask = (1u << lane) - 1u;
  uint32_t qtail = 0;                                  // survivors so far (chunk = index >> 5, ring entry = index & 63)

  for (unsigned unhandled = __ballot_sync(0xffffffffu, active); unhandled; ) {
    // ---- one pass: the lanes whose cell is within one cell of the first unhandled lane's
    const int L0 = __ffs(unhandled) - 1;
    const int c0x = __shfl_sync(0xffffffffu, ci[0], L0), c0y = __shfl_sync(0xffffffffu, ci[1], L0), c0z = __shfl_sync(0xffffffffu, ci[2], L0);
    const int kix = ci[0] - c0x, kiy = ci[1] - c0y, kiz = (DIM == 3) ? ci[2] - c0z : 0;
    const bool inPass = ((unhandled >> lane) & 1u) && abs(kix) <= 1 && abs(kiy) <= 1 && abs(kiz) <= 1;
    unhandled &= ~__ballot_sync(0xffffffffu, inPass);
    // my node in the pass frame; lanes outside the pass sit at infinity and hit nothing
    float pi[3];
    pi[0] = inPass ? fmaf((float)kix, csf[0], reli[0]) : 1.0e30f;
    pi[1] = fmaf((float)kiy, csf[1], reli[1]);
    pi[2] = (DIM == 3) ? fmaf((float)kiz, csf[2], reli[2]) : 0.f;
    // bounding box of the pass nodes (coordinates shifted to be positive: integer min / max of the bit patterns), grown by 1e-5
    // cell widths (two orders above the FP32 error of the coordinates), and their largest outer radius
    float blo[3] = {0.f, 0.f, 0.f}, bhi[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < DIM; ++q) {
      const float sh = 2.0f*csf[q];
      const unsigned v = __float_as_uint(pi[q] + sh);
      blo[q] = __uint_as_float(__reduce_min_sync(0xffffffffu, inPass ? v : 0x7f000000u)) - sh - 1.0e-5f*csf[q];
      bhi[q] = __uint_as_float(__reduce_max_sync(0xffffffffu, inPass ? v : 0u)) - sh + 1.0e-5f*csf[q];
    }
    const float r2reach = __uint_as_float(__reduce_max_sync(0xffffffffu, inPass ? __float_as_uint(fmaxf(r2hii, 0.f)) : 0u));
    const unsigned long long bx2 = f2_pack(pi[0], pi[0]), by2 = f2_pack(pi[1], pi[1]), bz2 = f2_pack(pi[2], pi[2]);

    uint32_t qhead = qtail;                            // a multiple of 32 here
    // ---- chunk of up to 32 buffered candidates: stage 1 for every lane, stage 2 for the ambiguous ones, one hit mask per lane
    auto process = [&](uint32_t first, uint32_t n) {
      const uint32_t base = first & (Q2_RING - 1), chunk = first >> 5;
      uint32_t hitWord = 0, inWord = 0;
      for (uint32_t c4 = 0; c4 < n; c4 += 4u) {
        const ulonglong2 X = *reinterpret_cast<const ulonglong2*>(qx + base + c4), Y = *reinterpret_cast<const ulonglong2*>(qy + base + c4);   // broadcast
        const float4 LO = *reinterpret_cast<const float4*>(qlo + base + c4), HI = *reinterpret_cast<const float4*>(qhi + base + c4);
        const unsigned long long rxa = f2_sub(bx2, X.x), rxb = f2_sub(bx2, X.y), rya = f2_sub(by2, Y.x), ryb = f2_sub(by2, Y.y);
        unsigned long long qa = f2_fma(rya, rya, f2_mul(rxa, rxa)), qb = f2_fma(ryb, ryb, f2_mul(rxb, rxb));
        if (DIM == 3) {
          const ulonglong2 Z = *reinterpret_cast<const ulonglong2*>(qz + base + c4);
          const unsigned long long rza = f2_sub(bz2, Z.x), rzb = f2_sub(bz2, Z.y);
          qa = f2_fma(rza, rza, qa); qb = f2_fma(rzb, rzb, qb);
        }
        float r2v[4];
        f2_unpack(qa, r2v[0], r2v[1]); f2_unpack(qb, r2v[2], r2v[3]);
        const float lo[4] = {LO.x, LO.y, LO.z, LO.w}, hi[4] = {HI.x, HI.y, HI.z, HI.w};
        uint32_t hn = 0, in = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (r2v[u] <= fmaxf(r2loi, lo[u])) hn |= 1u << u;
          if (r2v[u] <= fmaxf(r2hii, hi[u])) in |= 1u << u;
        }
        hitWord |= hn << c4; inWord |= in << c4;
      }
      if (n < 32u) { const uint32_t m = (1u << n) - 1u; hitWord &= m; inWord &= m; }
      const uint32_t t = ((uint32_t)lane < n) ? (uint32_t)qself[base + lane] : 0x3fu;
      // a node is not its own neighbour: the (at most 32 per tile) buffered entries that ARE nodes of this tile clear their lane's bit
      for (unsigned sm = __ballot_sync(0xffffffffu, (t & 0x20u) == 0u); sm; sm &= sm - 1u) {
        const int c = __ffs(sm) - 1;
        const uint32_t L = __shfl_sync(0xffffffffu, t, c) & 0x1fu;
        if ((uint32_t)lane == L) { hitWord &= ~(1u << c); inWord &= ~(1u << c); }
      }
      // stage 2: inside an outer sphere but no inner one -> the ellipsoids (FP32 with error band, else exact), lane by lane
      uint32_t amb = inWord & ~hitWord;
      while (__any_sync(0xffffffffu, amb != 0u)) {
        if (amb) {
          const uint32_t c = (uint32_t)(__ffs(amb) - 1);
          amb &= amb - 1u;
          const uint32_t e = base + c;
          float4 h0, h1;
          size_t slotj = 0;
          if (STAGEH) { h0 = *reinterpret_cast<const float4*>(qrow + 8*e); h1 = *reinterpret_cast<const float4*>(qrow + 8*e + 4); }
          else {
            const uint32_t code = (chunk < (uint32_t)maxChunks) ? scode[32u*chunk + c] : 0u;
            const uint32_t rr = code >> 5;
            slotj = (size_t)((rr < (uint32_t)JB_CAP) ? sjb[rr] : __ldg(&a.runs[rs + rr].x)) + (code & 31u);
            const float4 g1 = __ldg(frows4 + slotj*4 + 1), g3 = __ldg(frows4 + slotj*4 + 3);
            h0 = __ldg(frows4 + slotj*4 + 2); h1 = make_float4(g3.x, g3.y, g1.y, g1.z);
          }
          float rv[3], Hj[6];
          rv[0] = pi[0] - qx[e]; rv[1] = pi[1] - qy[e]; rv[2] = pi[2] - qz[e];
          if (DIM == 3) { Hj[0] = h0.x; Hj[1] = h0.y; Hj[2] = h0.z; Hj[3] = h0.w; Hj[4] = h1.x; Hj[5] = h1.y; }
          else { Hj[0] = h0.x; Hj[1] = h0.y; Hj[2] = h0.z; }
          const float e2i = eta2_f32<DIM>(Hi, rv), e2j = eta2_f32<DIM>(Hj, rv);
          bool hit = (e2i <= e2loi) || (e2j <= h1.z);
          if (!hit && !(e2i > e2hii && e2j > h1.w)) {
            if (STAGEH) {
              const uint32_t code = (chunk < (uint32_t)maxChunks) ? scode[32u*chunk + c] : 0u;
              const uint32_t rr = code >> 5;
              slotj = (size_t)((rr < (uint32_t)JB_CAP) ? sjb[rr] : __ldg(&a.runs[rs + rr].x)) + (code & 31u);
            }
            hit = (chunk < (uint32_t)maxChunks) && exact_pair<DIM>(a.rows, i, slotj, a.kext2);
          }
          if (hit) hitWord |= 1u << c;
        }
      }
      const unsigned ghostWord = __ballot_sync(0xffffffffu, (t & 0x80u) != 0u);
      cnt += __popc(hitWord);
      ghostHits += __popc(hitWord & ghostWord);
      if (chunk < (uint32_t)maxChunks) smask[32u*chunk + lane] = hitWord;
      __syncwarp();
    };

    // ---- the runs: one lane per candidate, records two runs and rows one run ahead of their use (two register sets, A and B,
    //      alternate: the rows of run r + 1 travel while run r is culled and queued)
    auto load_rows = [&](const uint4& rc, float4& p0, float4& p1, float4& p2, float4& p3) {
      const size_t c = (size_t)(rc.x + ((uint32_t)lane < rc.y ? lane : 0))*4;
      p0 = __ldg(frows4 + c); p1 = __ldg(frows4 + c + 1);
      if (STAGEH) { p2 = __ldg(frows4 + c + 2); p3 = __ldg(frows4 + c + 3); }
    };
    auto do_run = [&](uint32_t r, const uint4& rc, const float4& c0, const float4& c1, const float4& c2, const float4& c3) {
      const uint32_t jb = rc.x, len = rc.y;
      const int kx = (int)(rc.z & 0xffffu) - c0x, ky = (int)(rc.z >> 16) - c0y, kz = (DIM == 3) ? (int)rc.w - c0z : 0;
      // a pass node sits within one cell of the origin cell and reaches one cell further: runs beyond two cells are out of reach
      if (abs(kx) > 2 || abs(ky) > 2 || abs(kz) > 2) return;
      float pj[3];
      pj[0] = fmaf((float)kx, csf[0], c0.x); pj[1] = fmaf((float)ky, csf[1], c0.y); pj[2] = (DIM == 3) ? fmaf((float)kz, csf[2], c0.z) : 0.f;
      float d2 = 0.f;
#pragma unroll
      for (int q = 0; q < DIM; ++q) { const float d = fmaxf(fmaxf(blo[q] - pj[q], pj[q] - bhi[q]), 0.f); d2 = fmaf(d, d, d2); }
      const bool keep = (uint32_t)lane < len && d2 <= fmaxf(r2reach, c1.x)*1.00001f;
      const unsigned keepMask = __ballot_sync(0xffffffffu, keep);
      if (keepMask == 0u) return;
      if (keep) {
        const uint32_t sidx = qtail + (uint32_t)__popc(keepMask & ltMask);
        const uint32_t e = sidx & (Q2_RING - 1);
        const uint32_t slot = jb + (uint32_t)lane;
        qx[e] = pj[0]; qy[e] = pj[1]; qz[e] = pj[2]; qlo[e] = c0.w; qhi[e] = c1.x;
        if (STAGEH) {
          *reinterpret_cast<float4*>(qrow + 8*e) = c2;
          *reinterpret_cast<float4*>(qrow + 8*e + 4) = make_float4(c3.x, c3.y, c1.y, c1.z);
        }
        qself[e] = (unsigned char)(((slot - tile0 < 32u) ? (slot - tile0) : 0x3fu) | ((__float_as_uint(c1.w) >= a.nInt) ? 0x80u : 0u));
        if ((sidx >> 5) < (uint32_t)maxChunks) scode[sidx] = (unsigned short)((r << 5) | (uint32_t)lane);
      }
      qtail += (uint32_t)__popc(keepMask);
      if (qtail - qhead >= 32u) {
        __syncwarp();
        process(qhead, 32u);
        qhead += 32u;
      }
    };
    uint4 recA = (R > 0u) ? __ldg(a.runs + rs) : make_uint4(0u, 0u, 0u, 0u);
    uint4 recB = (R > 1u) ? __ldg(a.runs + rs + 1u) : make_uint4(0u, 0u, 0u, 0u);
    float4 a0, a1, a2, a3, b0, b1, b2, b3;
    a0 = a1 = a2 = a3 = b0 = b1 = b2 = b3 = make_float4(0.f, 0.f, 0.f, 0.f);
    load_rows(recA, a0, a1, a2, a3);
    for (uint32_t r = 0; r < R; r += 2u) {
      if (r + 1u < R) load_rows(recB, b0, b1, b2, b3);
      const uint4 rcA = recA;
      if (r + 2u < R) recA = __ldg(a.runs + rs + r + 2u);
      do_run(r, rcA, a0, a1, a2, a3);
      if (r + 1u >= R) break;
      if (r + 2u < R) load_rows(recA, a0, a1, a2, a3);
      const uint4 rcB = recB;
      if (r + 3u < R) recB = __ldg(a.runs + rs + r + 3u);
      do_run(r + 1u, rcB, b0, b1, b2, b3);
    }
    const uint32_t nLeft = qtail - qhead;
    if (nLeft) {
      // pad the last group of four with candidates at infinity
      if ((uint32_t)lane >= nLeft && (uint32_t)lane < ((nLeft + 3u) & ~3u)) { const uint32_t e = (qhead & (Q2_RING - 1)) + lane; qx[e] = -1.0e30f; qy[e] = 0.f; qz[e] = 0.f; qlo[e] = -1.f; qhi[e] = -1.f; }
      __syncwarp();
      process(qhead, nLeft);
      qtail = qhead + 32u;                             // the next pass starts a new chunk
    }
  }

  const uint32_t nChunks = (qtail + 31u) >> 5;
  if (inRange) a.nbrCount[i] = cnt;
  uint32_t mx = cnt;
  unsigned long long sAll = cnt, sGhost = ghostHits;
  for (int d = 16; d; d >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    sAll += __shfl_xor_sync(0xffffffffu, sAll, d);
    sGhost += __shfl_xor_sync(0xffffffffu, sGhost, d);
  }
  unsigned long long off = 0;
  if (lane == 0) {
    off = atomicAdd(&a.counters[3], (unsigned long long)mx*SPHB200_TILE);
    a.tileRows[tile] = mx; a.tileOff[tile] = off;
    if (sAll) atomicAdd(&a.counters[1], sAll);
    if (sGhost) atomicAdd(&a.counters[0], sGhost);
    atomicMax(&a.counters[4], (unsigned long long)mx);
    atomicMax(&a.counters[6], (unsigned long long)nChunks);      // sizes the per-chunk records of the next build (too few -> host redoes)
  }
  off = __shfl_sync(0xffffffffu, off, 0);
  if (off + (unsigned long long)mx*SPHB200_TILE > a.nbrCap || nChunks > (uint32_t)maxChunks) return;
  __syncwarp();
  // ---- expand the masks into the tile's block of the sliced-ELL array, row by row (row k: the k-th neighbour of every lane),
  //      four rows per trip so that the shared-memory look-ups of different rows overlap
  uint32_t* out = a.nbr + off + lane;
  uint32_t ch = 0, m = (nChunks > 0u) ? smask[lane] : 0u;
  for (uint32_t k = 0; k < mx; k += 4u) {
    uint32_t idx[4], slot[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      idx[u] = 0xffffffffu;
      if (k + u < cnt) {
        while (m == 0u) { ++ch; m = smask[32u*ch + lane]; }
        idx[u] = 32u*ch + (uint32_t)(__ffs(m) - 1);
        m &= m - 1u;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) slot[u] = (idx[u] != 0xffffffffu) ? (uint32_t)scode[idx[u]] : 0xffffffffu;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t code = slot[u];
      uint32_t sl = 0u;                                                         // padding entries point at slot 0 (never used)
      if (code != 0xffffffffu) {
        const uint32_t rr = code >> 5;
        sl = ((rr < (uint32_t)JB_CAP) ? sjb[rr] : __ldg(&a.runs[rs + rr].x)) + (code & 31u);
      }
      if (k + u < mx) out[(size_t)(k + u)*SPHB200_TILE] = sl;
    }
  }
}

// dilated (Morton) coordinate tables: dilTab[axis*SPHB200_DIL + c] = bits of c spread to the axis' key positions
__global__ void __launch_bounds__(RB) k_dilate_table(GridDev g, uint32_t* __restrict__ tab) {
  const int c = blockIdx.x*RB + threadIdx.x;
  const int axis = blockIdx.y;
  if (c < g.nc[axis]) tab[axis*SPHB200_DIL + c] = dilate(g, axis, c);
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
template int sphb200_ensure<uint4>(sphb200_ctx*, uint4*&, size_t&, size_t);
template int sphb200_ensure<unsigned long long>(sphb200_ctx*, unsigned long long*&, size_t&, size_t);

static int build_grid(sphb200_ctx* c, const double* bb /*lo3 hi3 ext3 sumext3*/) {
  GridDev& g = c->grid;
  const int nd = c->ndim;
  const int maxBitsTotal = 27;               // 128 Mi table entries
  int totalBits;
  double cs[3] = {1, 1, 1};
  for (int a = 0; a < 3; ++a) { g.lo[a] = 0; g.cs[a] = 1; g.nc[a] = 1; g.bits[a] = 0; g.mask[a] = 0; }
  for (int a = 0; a < nd; ++a) {
    double e = bb[6 + a];
    if (!(e > 0.0) || !std::isfinite(e)) return sphb200_fail(c, "build_pairs: non-positive or non-finite kernel extent (bad H?)");
    // Cell width: the largest extent when the extents are similar (then the 3^DIM stencil of a node's cell holds every
    // possible neighbour); with a heavy tail of large nodes, 1.25 x the mean extent, and the tiles near large nodes walk a
    // wider stencil (k_cell_reach / tileRadius) -- otherwise a handful of large nodes would coarsen the grid for everyone.
    const double emean = bb[9 + a]/(double)c->n;
    double w = e;
    if (!c->forceR1 && emean > 0.0 && 1.25*emean < e) w = std::max(1.25*emean, e/(double)SPHB200_MAX_STENCIL);
    cs[a] = w*(1.0 + 1.0e-9);
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
  // stencil radius (in cells) that covers the largest extent
  c->stencilR = 1;
  for (int a = 0; a < nd; ++a) c->stencilR = std::max(c->stencilR, (int)std::ceil(bb[6 + a]/cs[a]));
  if (c->stencilR > SPHB200_MAX_STENCIL) return sphb200_fail(c, "internal: stencil radius exceeds the supported maximum");
  // interleave: level by level, axes that still have bits
  int pos = 0;
  for (int l = 0; l < 16; ++l)
    for (int a = 0; a < nd; ++a)
      if (l < g.bits[a]) { g.bitpos[a][l] = (uint8_t)pos; g.mask[a] |= (1u << pos); ++pos; }
  g.tableSize = 1u << pos;
  // Two-level walk (experimental, off unless SPHB200_FINE_WALK=1 is set in the environment): sort by cells of half the width.  With one
  // more bit per axis, interleaved level by level like the coarse key, the lowest ndim bits of the fine key are the child index and
  // fine key >> ndim is the coarse key, so a coarse cell stays one contiguous range of the sorted order.
  c->fineWalk = false;
  static const bool fineWanted = [] { const char* e = std::getenv("SPHB200_FINE_WALK"); return e && e[0] == '1'; }();
  if (fineWanted && c->stencilR == 1 && pos + nd <= maxBitsTotal) {
    bool fits = true;
    for (int a = 0; a < nd; ++a) fits = fits && 2*g.nc[a] <= SPHB200_DIL && g.bits[a] + 1 <= 16;
    if (fits) {
      GridDev& f = c->gridFine;
      f = g;
      for (int a = 0; a < 3; ++a) { f.mask[a] = 0; f.bits[a] = 0; }
      for (int a = 0; a < nd; ++a) { f.nc[a] = 2*g.nc[a]; f.cs[a] = 0.5*g.cs[a]; f.bits[a] = g.bits[a] + 1; }
      int fp = 0;
      for (int l = 0; l < 16; ++l)
        for (int a = 0; a < nd; ++a)
          if (l < f.bits[a]) { f.bitpos[a][l] = (uint8_t)fp; f.mask[a] |= (1u << fp); ++fp; }
      f.tableSize = 1u << fp;
      c->fineWalk = true;
    }
  }
  return 0;
}

static bool sphb200_ghost_split_wanted() {      // SPHB200_GHOST_SPLIT=0 in the environment: ghosts sorted among the internal nodes (A/B, tests)
  const char* e = std::getenv("SPHB200_GHOST_SPLIT");
  return !(e && e[0] == '0');
}
static bool sphb200_nbr_v2_wanted() {          // read at every build: the tests switch between the two builders inside one process
  const char* e = std::getenv("SPHB200_NBR_V2");
  return !(e && e[0] == '0');
}

static int pack_rows_impl(sphb200_ctx* c, bool range, size_t first, size_t count, bool withFrows = true) {
  if (!c->sortValid) return sphb200_fail(c, "internal: pack_rows before sort");
  c->rowsAtEval = false;                       // the rows are about to change: compressed pair forces can no longer be expanded
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
  a.rows = c->rows; a.aux2 = c->aux2; a.auxPneg = tens ? c->auxPneg : nullptr; a.auxSomr2 = tens ? c->auxSomr2 : nullptr;
  a.auxDvDxQ = needQ ? c->auxDvDxQ : nullptr; a.auxfCl = mult ? c->auxfCl : nullptr; a.auxfCq = mult ? c->auxfCq : nullptr;
  a.perm = c->perm; a.keyApi = c->cellKeyApi; a.skey = c->skey; a.n = c->n;
  a.rawP = (c->opt.hydro == SPHB200_HYDRO_CRKSPH) ? 1 : 0;
  a.aniso = c->counters + 8;
  if (range) {
    if (first + count > c->n) return sphb200_fail(c, "internal: pack_rows range exceeds the node count");
    if (count == 0) return 0;
    if (sphb200_inverse_perm(c)) return 1;
    a.invPerm = c->invPerm; a.first = first; a.count = count;
    const unsigned nbr = (unsigned)((count + RB - 1)/RB);
    if (c->ndim == 3) k_pack<3, true><<<nbr, RB, 0, c->stream>>>(a); else k_pack<2, true><<<nbr, RB, 0, c->stream>>>(a);
    KERNEL_CHECK(c, "k_pack(range)");
    return 0;
  }
  CU_CHECK(c, cudaMemsetAsync(c->counters + 8, 0, sizeof(unsigned long long), c->stream));
  { size_t fcap = c->frows ? c->frowsCap : 0;
    if (sphb200_ensure(c, c->frows, fcap, c->cap*(size_t)Fr::ROW)) return 1;
    c->frowsCap = fcap; }
  a.frows = c->frows; a.g = c->grid;
  a.kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  // FP32 error of a reconstructed relative position, in units of 7 * 2^-24 cell widths: k*cs + rel_i - rel_j with |k| <= rad in
  // k_nbr_build; k_nbr_build2 places both nodes in the frame of the tile (|k_i| <= 1, |k_j| <= 2): 12 * 2^-24, budgeted as 14
  if (withFrows) c->nbrV2 = sphb200_nbr_v2_wanted() && c->stencilR == 1 && !c->fineWalk;      // a values-only re-pack leaves the FP32 rows, hence their error model, alone
  a.csmax = std::max(c->grid.cs[0], std::max(c->grid.cs[1], c->ndim == 3 ? c->grid.cs[2] : 0.0))*(c->nbrV2 ? 2.0 : (3.0*c->stencilR + 4.0)/7.0);
  static const bool gatherPack = [] { const char* e = std::getenv("SPHB200_PACK_GATHER"); return e && e[0] == '1'; }();     // A/B switch: the round-1 kernel
  if (gatherPack) {
    const unsigned nb = (unsigned)((c->n + RB - 1)/RB);
    if (c->ndim == 3) k_pack<3, false><<<nb, RB, 0, c->stream>>>(a); else k_pack<2, false><<<nb, RB, 0, c->stream>>>(a);
    KERNEL_CHECK(c, "k_pack");
  } else {
    if (sphb200_inverse_perm(c)) return 1;
    a.invPerm = c->invPerm;
    const unsigned nb = (unsigned)((c->n + 32*PS_WARPS - 1)/(32*PS_WARPS));
    if (withFrows) { if (c->ndim == 3) k_pack_scatter<3, true><<<nb, 32*PS_WARPS, 0, c->stream>>>(a); else k_pack_scatter<2, true><<<nb, 32*PS_WARPS, 0, c->stream>>>(a); }
    else           { if (c->ndim == 3) k_pack_scatter<3, false><<<nb, 32*PS_WARPS, 0, c->stream>>>(a); else k_pack_scatter<2, false><<<nb, 32*PS_WARPS, 0, c->stream>>>(a); }
    KERNEL_CHECK(c, "k_pack_scatter");
  }
  c->rowsValid = true;
  return 0;
}

// Every caller but the sort re-packs on a valid connectivity (it checked pairsValid): the FP32 rows of the neighbour build are left alone.
// SPHB200_PACK_VALUES_ONLY=0 packs them all the same (A/B).
int sphb200_pack_rows(sphb200_ctx* c) {
  static const bool valuesOnly = [] { const char* e = std::getenv("SPHB200_PACK_VALUES_ONLY"); return !(e && e[0] == '0'); }();
  return pack_rows_impl(c, false, 0, 0, !(valuesOnly && c->pairsValid));
}
int sphb200_pack_rows_all(sphb200_ctx* c) { return pack_rows_impl(c, false, 0, 0, true); }
// Re-pack the rows of the nodes [first, first + count) of the host order after their non-geometric fields changed (late halo fields)
int sphb200_pack_rows_range(sphb200_ctx* c, size_t first, size_t count) { return pack_rows_impl(c, true, first, count); }

int sphb200_bounds_reduce(sphb200_ctx* c, size_t count) {
  const double kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  const int nbb = (int)std::min<size_t>(296, (count + RB - 1)/RB);
  if (c->ndim == 3) k_bbox<3><<<nbb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], count, kext, c->reduceBuf);
  else              k_bbox<2><<<nbb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], count, kext, c->reduceBuf);
  KERNEL_CHECK(c, "k_bbox");
  k_bbox_final<<<1, 32*12, 0, c->stream>>>(c->reduceBuf, nbb, c->reduceBuf + 296*12);
  KERNEL_CHECK(c, "k_bbox_final");
  CU_CHECK(c, cudaMemcpyAsync(c->reduceHost, c->reduceBuf + 296*12, 12*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  return 0;
}

int sphb200_sort_and_pack(sphb200_ctx* c) {
  const size_t n = c->n;
  if (c->forceR1 && c->coarseHold > 0 && --c->coarseHold == 0) c->forceR1 = false;    // re-examine the fine grid now and then
  if (!c->have[S_POS] || !c->have[S_H]) return sphb200_fail(c, "build_pairs: position and H must be uploaded first");
  if (!c->W.set) return sphb200_fail(c, "build_pairs: kernel table not set (need the kernel extent)");
  const double kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  // bbox + extents
  if (sphb200_bounds_reduce(c, n)) return 1;
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int a = 0; a < c->ndim; ++a)
    if (!std::isfinite(c->reduceHost[a]) || !std::isfinite(c->reduceHost[3 + a]) || !std::isfinite(c->reduceHost[6 + a]))
      return sphb200_fail(c, "build_pairs: non-finite position or H");
  if (build_grid(c, c->reduceHost)) return 1;

  const GridDev& gs = c->fineWalk ? c->gridFine : c->grid;          // the grid the nodes are sorted by
  // ghost split: a second copy of the cell table for the ghost nodes (see k_cell_count); not with the experimental two-level walk
  c->ghostBase = (c->nGhost > 0 && !c->fineWalk && sphb200_ghost_split_wanted() && 2ull*gs.tableSize < 0x7fffffffull) ? gs.tableSize : 0u;
  const uint32_t cells = gs.tableSize + c->ghostBase;
  const size_t tbl = (size_t)cells + 1;
  if (sphb200_ensure(c, c->cellStart, c->cellCap, tbl)) return 1;
  if (sphb200_ensure(c, c->cellCursor, c->cellCursorCap, tbl)) return 1;
  CU_CHECK(c, cudaMemsetAsync(c->cellStart, 0, tbl*sizeof(uint32_t), c->stream));
  const unsigned nb = (unsigned)((n + RB - 1)/RB);
  {
    int ncmax = std::max(gs.nc[0], std::max(gs.nc[1], gs.nc[2]));
    k_dilate_table<<<dim3((unsigned)((ncmax + RB - 1)/RB), (unsigned)c->ndim), RB, 0, c->stream>>>(gs, c->dilTab);
    KERNEL_CHECK(c, "k_dilate_table");
  }
  if (c->ndim == 3) k_cell_count<3><<<nb, RB, 0, c->stream>>>(c->api[S_POS], n, gs, c->dilTab, c->cellKeyApi, c->cellStart, c->nInt, c->ghostBase);
  else              k_cell_count<2><<<nb, RB, 0, c->stream>>>(c->api[S_POS], n, gs, c->dilTab, c->cellKeyApi, c->cellStart, c->nInt, c->ghostBase);
  KERNEL_CHECK(c, "k_cell_count");
  if (sphb200_scan_u32(c, c->cellStart, c->cellStart, cells)) return 1;
  CU_CHECK(c, cudaMemcpyAsync(c->cellCursor, c->cellStart, (size_t)cells*sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  k_cell_scatter<<<nb, RB, 0, c->stream>>>(c->cellKeyApi, n, c->cellCursor, c->perm, c->nInt, c->ghostBase);
  KERNEL_CHECK(c, "k_cell_scatter");
  k_cell_order<<<(cells + RB - 1)/RB, RB, 0, c->stream>>>(c->cellStart, cells, c->perm);
  KERNEL_CHECK(c, "k_cell_order");
  c->sortValid = true;
  c->chunkListsValid = false;
  if (c->stencilR > 1) {
    if (sphb200_ensure(c, c->cellReach, c->cellReachCap, tbl)) return 1;
    CU_CHECK(c, cudaMemsetAsync(c->cellReach, 0, tbl*sizeof(uint32_t), c->stream));
    if (c->ndim == 3) k_cell_reach<3><<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], n, kext, c->grid, c->dilTab, c->stencilR, c->cellReach);
    else              k_cell_reach<2><<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->api[S_H], n, kext, c->grid, c->dilTab, c->stencilR, c->cellReach);
    KERNEL_CHECK(c, "k_cell_reach");
    // Is the fine grid worth it?  Candidate volume per node ~ (2 r + 1)^DIM cells of the fine width against 3^DIM cells of the
    // largest extent.  Large nodes that are rare but spread evenly put every tile within reach of one of them, and the fine
    // grid then costs more than the coarse one: measured on the reach map itself (one extra host round trip, only on builds
    // with a heavy tail of extents), and remembered for the next builds.
    CU_CHECK(c, cudaMemsetAsync(c->counters + 10, 0, sizeof(unsigned long long), c->stream));
    if (c->ndim == 3) k_reach_cost<3><<<nb, RB, 0, c->stream>>>(c->cellKeyApi, n, c->cellReach, c->counters + 10);
    else              k_reach_cost<2><<<nb, RB, 0, c->stream>>>(c->cellKeyApi, n, c->cellReach, c->counters + 10);
    KERNEL_CHECK(c, "k_reach_cost");
    unsigned long long sumCells = 0;
    CU_CHECK(c, cudaMemcpyAsync(&sumCells, c->counters + 10, sizeof(sumCells), cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    double volFine = 1.0, volCoarse = 1.0;
    for (int a = 0; a < c->ndim; ++a) { volFine *= c->grid.cs[a]; volCoarse *= 3.0*c->reduceHost[6 + a]; }
    if ((double)sumCells*volFine > 0.8*(double)n*volCoarse) {
      c->forceR1 = true; c->coarseHold = 32;             // wide cells for the next 32 builds, then look again
      return sphb200_sort_and_pack(c);
    }
  }
  return sphb200_pack_rows_all(c);
}

int sphb200_neighbors(sphb200_ctx* c) {
  const size_t n = c->n;
  c->nTiles = (n + SPHB200_TILE - 1)/SPHB200_TILE;
  const double kext = std::max(c->W.kext, c->WQ.set ? c->WQ.kext : 0.0);
  const unsigned nb = (unsigned)((c->nTiles + 3)/4);
  // first guesses for the variable-size buffers; afterwards the capacities of the previous build are reused
  if (!c->runs && sphb200_ensure(c, c->runs, c->runsCap, c->nTiles*72 + 1024)) return 1;
  if (!c->nbr && sphb200_ensure(c, c->nbr, c->nbrCap, n*(size_t)(c->ndim == 3 ? 150 : 60) + 1024)) return 1;
  if (c->listRows <= 0) c->listRows = (c->ndim == 3) ? 192 : 96;
  for (int attempt = 0; attempt < 5; ++attempt) {
    NbrArgs a{};
    a.rows = c->rows; a.frows = c->frows; a.perm = c->perm; a.skey = c->skey; a.cellStart = c->cellStart; a.dilTab = c->dilTab;
    a.ghostBase = c->ghostBase;
    a.n = n; a.nInt = (uint32_t)c->nInt; a.kext2 = kext*kext; a.g = c->grid;
    a.nbrCount = c->nbrCount; a.tileRows = c->tileRows; a.tileOff = c->tileOff; a.nbr = c->nbr; a.nbrCap = c->nbrCap;
    a.runs = c->runs; a.runsCap = c->runsCap; a.tileRunStart = c->tileRunStart; a.tileRunCount = c->tileRunCount;
    a.counters = c->counters;
    a.fine = c->fineWalk ? 1 : 0; a.gf = c->gridFine;
    a.cellReach = (c->stencilR > 1) ? c->cellReach : nullptr;
    if (c->stencilR > 1) { if (sphb200_ensure(c, c->tileRadius, c->tileRadiusCap, c->nTiles + 1)) return 1; a.tileRadius = c->tileRadius; }
    CU_CHECK(c, cudaMemsetAsync(c->counters, 0, 8*sizeof(unsigned long long), c->stream));
    // 1. candidate runs per tile
    if (c->fineWalk) { if (c->ndim == 3) k_tile_runs<3, true><<<nb, 128, 0, c->stream>>>(a); else k_tile_runs<2, true><<<nb, 128, 0, c->stream>>>(a); }
    else             { if (c->ndim == 3) k_tile_runs<3, false><<<nb, 128, 0, c->stream>>>(a); else k_tile_runs<2, false><<<nb, 128, 0, c->stream>>>(a); }
    KERNEL_CHECK(c, "k_tile_runs");
    // 2. the predicate, once per (node, candidate), and the sliced-ELL lists
    {
      if (c->nbrChunks <= 0) c->nbrChunks = (c->ndim == 3) ? 48 : 24;
      const size_t perWarp = c->nbrV2 ? (size_t)Q2_FIXED + (size_t)c->nbrChunks*Q2_CHUNKB
                                      : (size_t)32*CROW*4 + (size_t)S1F*4 + (size_t)JB_CAP*4 + (size_t)c->listRows*64 + SPHB200_NBR_PAD;
      int warps = NB_WARPS;
      while (warps > 1 && warps*perWarp > 220*1024) warps >>= 1;     // very long lists: fewer tiles per CTA
      if (perWarp > 220*1024)
        return sphb200_fail(c, "build_pairs: a node has more neighbours than the list staging can hold (H far too large for the node spacing?)");
      const size_t shm = warps*perWarp;
      const unsigned nbb = (unsigned)((c->nTiles + warps - 1)/warps);
      if (c->nbrV2) {
        // the H tensors of the candidates are staged for stage 2 unless the previous build saw isotropic H only (a hint, not a promise:
        // both variants decide every pair correctly, the unstaged one merely pays a global fetch per ambiguous candidate)
        const bool stageH = !c->isoHint;
#define SPHB200_LAUNCH_NB2(D, S) do { CU_CHECK(c, cudaFuncSetAttribute(k_nbr_build2<D, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); \
                                       k_nbr_build2<D, S><<<nbb, 32*warps, shm, c->stream>>>(a, c->nbrChunks); } while (0)
        if (c->ndim == 3) { if (stageH) SPHB200_LAUNCH_NB2(3, true); else SPHB200_LAUNCH_NB2(3, false); }
        else              { if (stageH) SPHB200_LAUNCH_NB2(2, true); else SPHB200_LAUNCH_NB2(2, false); }
#undef SPHB200_LAUNCH_NB2
      } else {
        if (c->ndim == 3) { CU_CHECK(c, cudaFuncSetAttribute(k_nbr_build<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); k_nbr_build<3><<<nbb, 32*warps, shm, c->stream>>>(a, c->listRows); }
        else              { CU_CHECK(c, cudaFuncSetAttribute(k_nbr_build<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm)); k_nbr_build<2><<<nbb, 32*warps, shm, c->stream>>>(a, c->listRows); }
      }
      KERNEL_CHECK(c, "k_nbr_build");
    }
    // one host round trip: totals and capacity check
    CU_CHECK(c, cudaMemcpyAsync(c->countersHost, c->counters, 9*sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    const size_t needRuns = (size_t)c->countersHost[2], needNbr = (size_t)c->countersHost[3], needRows = (size_t)c->countersHost[4];
    if (c->countersHost[5] != 0ull) return 2;       // a tile has too many candidate runs for the list codes: the caller rebuilds with wide cells
    const bool okRuns = needRuns <= c->runsCap;
    const size_t needChunks = (size_t)c->countersHost[6];
    const bool okRows = c->nbrV2 ? needChunks <= (size_t)c->nbrChunks : needRows <= (size_t)c->listRows;
    const bool okNbr = okRuns && needNbr <= c->nbrCap;               // the list size is only known once the test ran everywhere
    if (okRuns && okRows && okNbr) {
      c->nEdges = (size_t)c->countersHost[1];
      // every internal-internal pair appears as two directed edges, every internal-ghost pair as one
      c->npairs = (size_t)((c->countersHost[1] + c->countersHost[0])/2);
      c->nSlots = needNbr;
      // staging for the next build: longest list + 1/8 (in an evolving problem the longest list grows from step to step, and a
      // build that overflows its staging is redone at full cost)
      c->listRows = (int)((needRows + needRows/8 + 8 + 7)/8*8);
      if (c->nbrV2) c->nbrChunks = (int)(needChunks + needChunks/4 + 2);
      c->pairsValid = true;
      c->allIsotropic = (c->countersHost[8] == 0ull);       // k_pack of this build saw only H = h^-1 I
      c->isoHint = c->allIsotropic;
      c->stats.directed_edges = c->nEdges;
      return 0;
    }
    if (!okRuns && sphb200_ensure(c, c->runs, c->runsCap, needRuns + needRuns/8)) return 1;
    if (okRuns && !okRows) {
      if (c->nbrV2) c->nbrChunks = (int)(needChunks + needChunks/4 + 4);
      else c->listRows = (int)(needRows + needRows/8 + 16);
    }
    if (okRuns && needNbr > c->nbrCap && sphb200_ensure(c, c->nbr, c->nbrCap, needNbr + needNbr/8)) return 1;
  }
  return sphb200_fail(c, "build_pairs: neighbour buffers failed to converge");
}
