// energy.cu -- K6 compatible-energy update and the NodePairList / PairwiseField exporters.
//
// Replaces SpecificThermalEnergyPolicy::update (Hydro/SpecificThermalEnergyPolicy.cc:47-174).  The reference walks the
// pair list once and scatters the discrete pair work into both nodes; here node i walks its own directed edges
// (which hold deltaDvDt of SPH.cc:427 in i's orientation) and gathers only its own share, so no atomics are needed.
#include "sphb200_internal.cuh"
#include "nbr_ring.cuh"
#include <algorithm>
#include <cfloat>

namespace {

constexpr int RB = 256;

// erow[s] = { v + DvDt*hdt (DIM), DepsDt0, m, original index | position (DIM), H (NS) }, one record per sorted node with the stride of a node
// row (so that the record ring of nbr_ring.cuh streams it like one); USED doubles of it are filled.
// The geometry tail exists only for the compressed pair-force modes (PACC_ISO: position and 1/h; PACC_TENSOR: position and H), which are
// expanded here from the node rows the evaluation read: delta = sd r_ij, or a H_i.(H_i.r_ij) + b H_j.(H_j.r_ij).
template <int DIM, int MODE> struct ERow {
  static constexpr int GEOM = (MODE == PACC_FULL) ? 0 : (MODE == PACC_ISO ? DIM + 1 : DIM + Dm<DIM>::NS);
  static constexpr int USED = DIM + 3 + GEOM;
  static constexpr int ES = Dm<DIM>::ROW;
  static constexpr int COPY = (USED*8 + 15)/16*16;          // bytes of a record the ring copies
  static_assert(USED <= ES, "the energy record must fit a node row");
};
// Ring depth x resident CTAs per SM, measured at 8 M (tensor mode) / 1 M (isotropic mode), per-lane gather before: 14.85 / 1.71 ms:
//   4 x 1: 18.85 / 1.66    3 x 2: 11.75 / 1.67    2 x 3: 10.95 / 1.27 ms -- this loop has little arithmetic per edge to overlap with
// and an IEEE division in its chain, so warps (24 per SM at 80 registers) hide more than ring depth does.
#ifndef SPHB200_ENERGY_STAGES
#define SPHB200_ENERGY_STAGES 2
#endif
#ifndef SPHB200_ENERGY_CTAS
#define SPHB200_ENERGY_CTAS 3
#endif
constexpr int EW = 8, ESTAGES = SPHB200_ENERGY_STAGES;        // warps per CTA and ring depth of k_energy
template <int DIM, int MODE> using ERing = NbrRing<DIM, 0, 0, ESTAGES, ERow<DIM, MODE>::COPY, false>;

template <int DIM, int MODE>
__global__ void __launch_bounds__(RB) k_energy_prep(const double* __restrict__ velApi, const double* __restrict__ massApi,
                                                    const uint32_t* __restrict__ perm, const double* __restrict__ DvDt,
                                                    const double* __restrict__ DepsDt, const double* __restrict__ rows, size_t n, size_t cap, double hdt,
                                                    double* __restrict__ erow) {
  using D = Dm<DIM>;
  constexpr int ES = ERow<DIM, MODE>::ES;
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  const size_t o = perm[s];
#pragma unroll
  for (int q = 0; q < DIM; ++q) erow[s*ES + q] = velApi[o*DIM + q] + DvDt[(size_t)q*cap + s]*hdt;
  erow[s*ES + DIM] = DepsDt[s];
  erow[s*ES + DIM + 1] = massApi[o];
  erow[s*ES + DIM + 2] = (double)o;            // original index: orients the pair (i_node < j_node, NodePairIdxType.hh:34-58)
  if (MODE != PACC_FULL) {
#pragma unroll
    for (int q = 0; q < DIM; ++q) erow[s*ES + DIM + 3 + q] = rows[s*D::ROW + D::R_POS + q];
    if (MODE == PACC_ISO) erow[s*ES + 2*DIM + 3] = rows[s*D::ROW + D::R_H];
    else {
#pragma unroll
      for (int q = 0; q < D::NS; ++q) erow[s*ES + 2*DIM + 3 + q] = rows[s*D::ROW + D::R_H + q];
    }
  }
}

// the pair force of a directed edge in i's orientation, expanded from its stored words
template <int DIM, int MODE>
__device__ __forceinline__ void pacc_expand(const double* __restrict__ pacc, unsigned long long slot, const double* gi, const double* gj, double* d) {
  constexpr int W = (MODE == PACC_ISO) ? 1 : (MODE == PACC_TENSOR ? 2 : DIM);
  if (MODE == PACC_FULL) {
#pragma unroll
    for (int q = 0; q < DIM; ++q) d[q] = pacc[pacc_at(W, slot, q)];
    return;
  }
  double rij[DIM];
#pragma unroll
  for (int q = 0; q < DIM; ++q) rij[q] = gi[q] - gj[q];
  if (MODE == PACC_ISO) {
    const double sd = pacc[pacc_at(W, slot, 0)];
#pragma unroll
    for (int q = 0; q < DIM; ++q) d[q] = sd*rij[q];
  } else {
    const double a = pacc[pacc_at(W, slot, 0)], b = pacc[pacc_at(W, slot, 1)];
    double ei[DIM], ej[DIM], hi[DIM], hj[DIM];
    sym_dot<DIM>(gi + DIM, rij, ei); sym_dot<DIM>(gi + DIM, ei, hi);
    sym_dot<DIM>(gj + DIM, rij, ej); sym_dot<DIM>(gj + DIM, ej, hj);
#pragma unroll
    for (int q = 0; q < DIM; ++q) d[q] = fma(a, hi[q], b*hj[q]);
  }
}

// One warp per tile, lane <-> node i, persistent CTAs; the neighbours' records arrive through the warp-cooperative ring (a per-lane
// gather of the 120-byte record was L1-wavefront bound: 14.1 ms of the 83 ms RK2 step at 8 M, profiles/r02_launches_rk2_step_8m_summary.txt).
template <int DIM, int MODE>
__global__ void __launch_bounds__(32*EW, SPHB200_ENERGY_CTAS) k_energy(const double* __restrict__ erow, const uint32_t* __restrict__ perm,
                                                     const uint32_t* __restrict__ nbrCount, const uint32_t* __restrict__ tileRows,
                                                     const unsigned long long* __restrict__ tileOff, const uint32_t* __restrict__ nbr,
                                                     const double* __restrict__ pacc, size_t nSlots, size_t n, uint32_t nInt,
                                                     double multiplier, double* __restrict__ epsApi) {
  constexpr int ES = ERow<DIM, MODE>::ES, GEOM = ERow<DIM, MODE>::GEOM, NREC = ERow<DIM, MODE>::COPY/8;
  using Ring = ERing<DIM, MODE>;
  extern __shared__ __align__(16) double smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Ring ring;
  ring.base = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)warp*(unsigned)Ring::WARPB;
  ring.rows = reinterpret_cast<const unsigned char*>(erow);
  ring.x1 = nullptr; ring.x2 = nullptr; ring.aux2 = nullptr;
  const size_t nTiles = (n + SPHB200_TILE - 1)/SPHB200_TILE;
  for (size_t tile = (size_t)blockIdx.x*EW + warp; tile < nTiles; tile += (size_t)gridDim.x*EW) {
    const size_t i = tile*SPHB200_TILE + lane;
    const bool inRange = i < n;
    const uint32_t o = inRange ? perm[i] : 0xffffffffu;
    const bool active = inRange && o < nInt;
    double vi[DIM], Di = 0, mi = 1, gi[GEOM > 0 ? GEOM : 1];
    if (inRange) {
#pragma unroll
      for (int q = 0; q < DIM; ++q) vi[q] = erow[i*ES + q];
      Di = erow[i*ES + DIM]; mi = erow[i*ES + DIM + 1];
#pragma unroll
      for (int q = 0; q < GEOM; ++q) gi[q] = erow[i*ES + DIM + 3 + q];
    } else {
#pragma unroll
      for (int q = 0; q < DIM; ++q) vi[q] = 0;
#pragma unroll
      for (int q = 0; q < (GEOM > 0 ? GEOM : 1); ++q) gi[q] = 0;
    }
    const uint32_t cnt = active ? nbrCount[i] : 0u;
    const uint32_t rows = tileRows[tile];
    const unsigned long long base = tileOff[tile] + lane;
    double acc = 0.0;
    ring_walk<Ring, ESTAGES>(ring, lane, rows, cnt,
      [&](uint32_t p) -> uint32_t { return (p < cnt) ? nbr[base + (unsigned long long)p*SPHB200_TILE] : 0u; },
      [&](uint32_t k, uint32_t) {
        const unsigned long long slot = base + (unsigned long long)k*SPHB200_TILE;
        double rj[NREC];
        ring.read_row(k, lane, rj);
        const double* vj = rj;
        const double* gj = rj + DIM + 3;
        double d[DIM];
        pacc_expand<DIM, MODE>(pacc, slot, gi, gj, d);
        const double Dj = rj[DIM], mj = rj[DIM + 1];
        const bool up = rj[DIM + 2] > (double)o;                          // original index of j > original index of i
        if (up) {
          // i is the pair's i-node: paccij = -mj*deltaDvDt (SPH.cc:430); duij = (vj12 - vi12).paccij
          double du = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; ++q) du += (vj[q] - vi[q])*(-mj*d[q]);
          const double sg = du < 0.0 ? -1.0 : 1.0;
          const double wti = fmax(DBL_EPSILON, Di*sg), wtj = fmax(DBL_EPSILON, Dj*sg);
          const double wi = wti/(wti + wtj);
          acc += wi*du;
        } else {
          // i is the pair's j-node; the stored delta is in i's orientation, so pacc(j<-i) = mi*delta
          double du = 0.0;
#pragma unroll
          for (int q = 0; q < DIM; ++q) du += (vi[q] - vj[q])*(mi*d[q]);
          const double sg = du < 0.0 ? -1.0 : 1.0;
          const double wtj = fmax(DBL_EPSILON, Dj*sg), wti = fmax(DBL_EPSILON, Di*sg);
          const double wj = wtj/(wtj + wti);
          acc += (1.0 - wj)*du*mj/mi;
        }
      });
    if (active) epsApi[o] += acc*multiplier;
  }
}

// ---- exporters: NodePairList (sorted (i,j), i<j in ORIGINAL numbering) and PairwiseField --------------------------------
__global__ void __launch_bounds__(128) k_hi_count(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ nbrCount,
                                                  const unsigned long long* __restrict__ tileOff, const uint32_t* __restrict__ nbr,
                                                  size_t n, uint32_t nInt, uint32_t* __restrict__ hiByOrig) {
  const size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = perm[i];
  if (o >= nInt) return;
  const unsigned long long base = tileOff[i/SPHB200_TILE] + (i % SPHB200_TILE);
  const uint32_t cnt = nbrCount[i];
  uint32_t h = 0;
  for (uint32_t k = 0; k < cnt; ++k) h += (perm[nbr[base + (unsigned long long)k*SPHB200_TILE]] > o) ? 1u : 0u;
  hiByOrig[o] = h;
}

__global__ void __launch_bounds__(128) k_emit_pairs(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ nbrCount,
                                                    const unsigned long long* __restrict__ tileOff, const uint32_t* __restrict__ nbr,
                                                    size_t n, uint32_t nInt, const unsigned long long* __restrict__ pairOff,
                                                    uint32_t* __restrict__ outI, uint32_t* __restrict__ outJ,
                                                    unsigned long long* __restrict__ outSlot) {
  const size_t i = (size_t)blockIdx.x*blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t o = perm[i];
  if (o >= nInt) return;
  const unsigned long long base = tileOff[i/SPHB200_TILE] + (i % SPHB200_TILE);
  const uint32_t cnt = nbrCount[i];
  const unsigned long long p0 = pairOff[o];
  unsigned long long t = p0;
  for (uint32_t k = 0; k < cnt; ++k) {
    const unsigned long long slot = base + (unsigned long long)k*SPHB200_TILE;
    const uint32_t jo = perm[nbr[slot]];
    if (jo > o) {
      unsigned long long q = t;                    // insertion sort by original j
      while (q > p0 && outJ[q - 1] > jo) { outJ[q] = outJ[q - 1]; outSlot[q] = outSlot[q - 1]; --q; }
      outJ[q] = jo; outSlot[q] = slot; outI[t] = o;
      ++t;
    }
  }
}

template <int DIM, int MODE>
__global__ void __launch_bounds__(RB) k_emit_pacc(const uint32_t* __restrict__ outI, const uint32_t* __restrict__ outJ,
                                                  const unsigned long long* __restrict__ outSlot, const uint32_t* __restrict__ invPerm,
                                                  const double* __restrict__ rows, const double* __restrict__ massApi,
                                                  const double* __restrict__ pacc, size_t nSlots, size_t npairs, double* __restrict__ out) {
  using D = Dm<DIM>;
  const size_t k = (size_t)blockIdx.x*RB + threadIdx.x;
  if (k >= npairs) return;
  const double mj = massApi[outJ[k]];
  const unsigned long long slot = outSlot[k];
  double d[DIM], gi[DIM + D::NS], gj[DIM + D::NS];
  if (MODE != PACC_FULL) {                       // geometry of both ends from the node rows the evaluation read
    const double* ri = rows + (size_t)invPerm[outI[k]]*D::ROW; const double* rj = rows + (size_t)invPerm[outJ[k]]*D::ROW;
#pragma unroll
    for (int q = 0; q < DIM; ++q) { gi[q] = ri[D::R_POS + q]; gj[q] = rj[D::R_POS + q]; }
#pragma unroll
    for (int q = 0; q < D::NS; ++q) { gi[DIM + q] = ri[D::R_H + q]; gj[DIM + q] = rj[D::R_H + q]; }
  }
  pacc_expand<DIM, MODE>(pacc, slot, gi, gj, d);
#pragma unroll
  for (int q = 0; q < DIM; ++q) out[k*DIM + q] = -mj*d[q];    // SPH.cc:430
}

// u32 counts -> u64 exclusive offsets (single block serial-by-chunks; export path only, not on the hot path)
__global__ void k_scan64_small(const uint32_t* __restrict__ in, unsigned long long* __restrict__ out, size_t n) {
  __shared__ unsigned long long carry;
  __shared__ unsigned long long ws[32];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (size_t b = 0; b <= n; b += blockDim.x) {
    const size_t idx = b + threadIdx.x;
    unsigned long long v = (idx < n) ? in[idx] : 0ull, inc = v;
    for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d); if ((threadIdx.x & 31) >= d) inc += t; }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) {
      unsigned long long s = (threadIdx.x < (blockDim.x >> 5)) ? ws[threadIdx.x] : 0ull, si = s;
      for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(0xffffffffu, si, d); if (threadIdx.x >= d) si += t; }
      ws[threadIdx.x] = si - s;
    }
    __syncthreads();
    const unsigned long long ex = carry + ws[threadIdx.x >> 5] + inc - v;
    if (idx <= n) out[idx] = ex;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = ex + v;
    __syncthreads();
  }
}

}  // namespace

template <int DIM, int MODE>
static int launch_energy_mode(sphb200_ctx* c, double multiplier) {
  const size_t n = c->n;
  constexpr int ES = ERow<DIM, MODE>::ES;
  size_t need = n*ES*sizeof(double);
  if (need > c->stageBytes) {
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stageBytes = 0;
    CU_CHECK(c, cudaMalloc((void**)&c->stage, need + need/8));
    c->stageBytes = need + need/8;
  }
  const double hdt = 0.5*multiplier;
  const unsigned nb = (unsigned)((n + RB - 1)/RB);
  k_energy_prep<DIM, MODE><<<nb, RB, 0, c->stream>>>(c->api[S_VEL], c->api[S_MASS], c->perm, c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], c->rows, n, c->cap, hdt, c->stage);
  KERNEL_CHECK(c, "k_energy_prep");
  const size_t shm = (size_t)EW*ERing<DIM, MODE>::WARPB;
  CU_CHECK(c, cudaFuncSetAttribute(k_energy<DIM, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  int nsm = 148, perSM = 1;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, k_energy<DIM, MODE>, 32*EW, shm) != cudaSuccess || perSM < 1) perSM = 1;
  const unsigned nbt = (unsigned)std::min<size_t>((c->nTiles + EW - 1)/EW, (size_t)nsm*perSM);
  k_energy<DIM, MODE><<<nbt, 32*EW, shm, c->stream>>>(c->stage, c->perm, c->nbrCount, c->tileRows, c->tileOff, c->nbr, c->pacc, c->nSlots, n, (uint32_t)c->nInt, multiplier, c->api[S_EPS]);
  KERNEL_CHECK(c, "k_energy");
  return 0;
}

int sphb200_launch_energy(sphb200_ctx* c, double multiplier) {
  if (c->paccMode != PACC_FULL && !c->rowsAtEval)
    return sphb200_fail(c, "update_energy_compatible: the node rows of the evaluation were re-packed since (a derivative-consuming call on new state came "
                           "first); evaluate the derivatives again before the compatible energy update");
  if (c->ndim == 3) {
    if (c->paccMode == PACC_ISO) return launch_energy_mode<3, PACC_ISO>(c, multiplier);
    if (c->paccMode == PACC_TENSOR) return launch_energy_mode<3, PACC_TENSOR>(c, multiplier);
    return launch_energy_mode<3, PACC_FULL>(c, multiplier);
  }
  if (c->paccMode == PACC_ISO) return launch_energy_mode<2, PACC_ISO>(c, multiplier);
  if (c->paccMode == PACC_TENSOR) return launch_energy_mode<2, PACC_TENSOR>(c, multiplier);
  return launch_energy_mode<2, PACC_FULL>(c, multiplier);
}

int sphb200_pairs_to_host(sphb200_ctx* c, uint32_t* pi, uint32_t* pj, size_t cap, double* paccOut, size_t paccCap) {
  const size_t n = c->n, np = c->npairs;
  if (cap < np && (pi || pj)) return sphb200_fail(c, "download_pairs: buffer too small");
  if (paccOut && paccCap < np*(size_t)c->ndim) return sphb200_fail(c, "download_pair_accelerations: buffer too small");
  uint32_t *hi = nullptr, *dI = nullptr, *dJ = nullptr; unsigned long long *off = nullptr, *dS = nullptr; double* dP = nullptr;
  int rc = 0;
  auto cleanup = [&]() { cudaFree(hi); cudaFree(dI); cudaFree(dJ); cudaFree(off); cudaFree(dS); cudaFree(dP); };
#define PCHK(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cleanup(); return sphb200_fail(c, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)
  PCHK(cudaMalloc((void**)&hi, (c->nInt + 1)*sizeof(uint32_t)));
  PCHK(cudaMalloc((void**)&off, (c->nInt + 1)*sizeof(unsigned long long)));
  PCHK(cudaMalloc((void**)&dI, (np + 1)*sizeof(uint32_t)));
  PCHK(cudaMalloc((void**)&dJ, (np + 1)*sizeof(uint32_t)));
  PCHK(cudaMalloc((void**)&dS, (np + 1)*sizeof(unsigned long long)));
  PCHK(cudaMemsetAsync(hi, 0, (c->nInt + 1)*sizeof(uint32_t), c->stream));
  const unsigned nb = (unsigned)((n + 127)/128);
  if (n) {
    k_hi_count<<<nb, 128, 0, c->stream>>>(c->perm, c->nbrCount, c->tileOff, c->nbr, n, (uint32_t)c->nInt, hi);
    c->stats.launches++;
    k_scan64_small<<<1, 1024, 0, c->stream>>>(hi, off, c->nInt);
    c->stats.launches++;
    k_emit_pairs<<<nb, 128, 0, c->stream>>>(c->perm, c->nbrCount, c->tileOff, c->nbr, n, (uint32_t)c->nInt, off, dI, dJ, dS);
    c->stats.launches++;
  }
  if (pi) PCHK(cudaMemcpyAsync(pi, dI, np*sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (pj) PCHK(cudaMemcpyAsync(pj, dJ, np*sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  if (paccOut && np) {
    PCHK(cudaMalloc((void**)&dP, np*(size_t)c->ndim*sizeof(double)));
    const unsigned nbp = (unsigned)((np + RB - 1)/RB);
    if (c->paccMode != PACC_FULL) {
      if (!c->rowsAtEval) { cleanup(); return sphb200_fail(c, "download_pair_accelerations: the node rows of the evaluation were re-packed since; evaluate the derivatives again"); }
      if (sphb200_inverse_perm(c)) { cleanup(); return 1; }
    }
#define SPHB200_EMIT(D, M) k_emit_pacc<D, M><<<nbp, RB, 0, c->stream>>>(dI, dJ, dS, c->invPerm, c->rows, c->api[S_MASS], c->pacc, c->nSlots, np, dP)
    if (c->ndim == 3) { if (c->paccMode == PACC_ISO) SPHB200_EMIT(3, PACC_ISO); else if (c->paccMode == PACC_TENSOR) SPHB200_EMIT(3, PACC_TENSOR); else SPHB200_EMIT(3, PACC_FULL); }
    else              { if (c->paccMode == PACC_ISO) SPHB200_EMIT(2, PACC_ISO); else if (c->paccMode == PACC_TENSOR) SPHB200_EMIT(2, PACC_TENSOR); else SPHB200_EMIT(2, PACC_FULL); }
#undef SPHB200_EMIT
    c->stats.launches++;
    PCHK(cudaMemcpyAsync(paccOut, dP, np*(size_t)c->ndim*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  PCHK(cudaStreamSynchronize(c->stream));
  PCHK(cudaGetLastError());
#undef PCHK
  cleanup();
  return rc;
}
