// energy.cu -- K6 compatible-energy update and the NodePairList / PairwiseField exporters.
//
// Replaces SpecificThermalEnergyPolicy::update (Hydro/SpecificThermalEnergyPolicy.cc:47-174).  The reference walks the
// pair list once and scatters the discrete pair work into both nodes; here node i walks its own directed edges
// (which hold deltaDvDt of SPH.cc:427 in i's orientation) and gathers only its own share, so no atomics are needed.
#include "sphb200_internal.cuh"
#include <cfloat>

namespace {

constexpr int RB = 256;

// erow[s] = { v + DvDt*hdt (DIM), DepsDt0, m, original index }   stride DIM+3
template <int DIM>
__global__ void __launch_bounds__(RB) k_energy_prep(const double* __restrict__ velApi, const double* __restrict__ massApi,
                                                    const uint32_t* __restrict__ perm, const double* __restrict__ DvDt,
                                                    const double* __restrict__ DepsDt, size_t n, size_t cap, double hdt,
                                                    double* __restrict__ erow) {
  constexpr int ES = DIM + 3;
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  const size_t o = perm[s];
#pragma unroll
  for (int q = 0; q < DIM; ++q) erow[s*ES + q] = velApi[o*DIM + q] + DvDt[(size_t)q*cap + s]*hdt;
  erow[s*ES + DIM] = DepsDt[s];
  erow[s*ES + DIM + 1] = massApi[o];
  erow[s*ES + DIM + 2] = (double)o;            // original index: orients the pair (i_node < j_node, NodePairIdxType.hh:34-58)
}

template <int DIM>
__global__ void __launch_bounds__(128) k_energy(const double* __restrict__ erow, const uint32_t* __restrict__ perm,
                                                const uint32_t* __restrict__ nbrCount, const uint32_t* __restrict__ tileRows,
                                                const unsigned long long* __restrict__ tileOff, const uint32_t* __restrict__ nbr,
                                                const double* __restrict__ pacc, size_t nSlots, size_t n, uint32_t nInt,
                                                double multiplier, double* __restrict__ epsApi) {
  constexpr int ES = DIM + 3;
  const int lane = threadIdx.x & 31;
  const size_t tile = (size_t)blockIdx.x*(blockDim.x >> 5) + (threadIdx.x >> 5);
  const size_t i = tile*SPHB200_TILE + lane;
  if (tile*SPHB200_TILE >= n) return;
  const bool inRange = i < n;
  const uint32_t o = inRange ? perm[i] : 0xffffffffu;
  const bool active = inRange && o < nInt;
  double vi[DIM], Di = 0, mi = 1;
  if (inRange) {
#pragma unroll
    for (int q = 0; q < DIM; ++q) vi[q] = erow[i*ES + q];
    Di = erow[i*ES + DIM]; mi = erow[i*ES + DIM + 1];
  } else {
#pragma unroll
    for (int q = 0; q < DIM; ++q) vi[q] = 0;
  }
  const uint32_t cnt = active ? nbrCount[i] : 0u;
  const uint32_t rows = tileRows[tile];
  const unsigned long long base = tileOff[tile];
  double acc = 0.0;
  for (uint32_t k = 0; k < rows; ++k) {
    if (k >= cnt) continue;
    const unsigned long long slot = base + (unsigned long long)k*SPHB200_TILE + lane;
    const uint32_t j = nbr[slot];
    double vj[DIM], d[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { vj[q] = erow[(size_t)j*ES + q]; d[q] = pacc[pacc_index<DIM>(slot, q)]; }
    const double Dj = erow[(size_t)j*ES + DIM], mj = erow[(size_t)j*ES + DIM + 1];
    const bool up = erow[(size_t)j*ES + DIM + 2] > (double)o;         // original index of j > original index of i
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
  }
  if (active) epsApi[o] += acc*multiplier;
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

template <int DIM>
__global__ void __launch_bounds__(RB) k_emit_pacc(const uint32_t* __restrict__ outJ, const unsigned long long* __restrict__ outSlot,
                                                  const double* __restrict__ massApi, const double* __restrict__ pacc, size_t nSlots,
                                                  size_t npairs, double* __restrict__ out) {
  const size_t k = (size_t)blockIdx.x*RB + threadIdx.x;
  if (k >= npairs) return;
  const double mj = massApi[outJ[k]];
  const unsigned long long slot = outSlot[k];
#pragma unroll
  for (int q = 0; q < DIM; ++q) out[k*DIM + q] = -mj*pacc[pacc_index<DIM>(slot, q)];    // SPH.cc:430
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

int sphb200_launch_energy(sphb200_ctx* c, double multiplier) {
  const size_t n = c->n;
  const int ES = c->ndim + 3;
  size_t need = n*ES*sizeof(double);
  if (need > c->stageBytes) {
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stageBytes = 0;
    CU_CHECK(c, cudaMalloc((void**)&c->stage, need + need/8));
    c->stageBytes = need + need/8;
  }
  const double hdt = 0.5*multiplier;
  const unsigned nb = (unsigned)((n + RB - 1)/RB);
  const unsigned nbt = (unsigned)((c->nTiles + 3)/4);
  if (c->ndim == 3) {
    k_energy_prep<3><<<nb, RB, 0, c->stream>>>(c->api[S_VEL], c->api[S_MASS], c->perm, c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], n, c->cap, hdt, c->stage);
    KERNEL_CHECK(c, "k_energy_prep");
    k_energy<3><<<nbt, 128, 0, c->stream>>>(c->stage, c->perm, c->nbrCount, c->tileRows, c->tileOff, c->nbr, c->pacc, c->nSlots, n, (uint32_t)c->nInt, multiplier, c->api[S_EPS]);
  } else {
    k_energy_prep<2><<<nb, RB, 0, c->stream>>>(c->api[S_VEL], c->api[S_MASS], c->perm, c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], n, c->cap, hdt, c->stage);
    KERNEL_CHECK(c, "k_energy_prep");
    k_energy<2><<<nbt, 128, 0, c->stream>>>(c->stage, c->perm, c->nbrCount, c->tileRows, c->tileOff, c->nbr, c->pacc, c->nSlots, n, (uint32_t)c->nInt, multiplier, c->api[S_EPS]);
  }
  KERNEL_CHECK(c, "k_energy");
  return 0;
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
    if (c->ndim == 3) k_emit_pacc<3><<<nbp, RB, 0, c->stream>>>(dJ, dS, c->api[S_MASS], c->pacc, c->nSlots, np, dP);
    else              k_emit_pacc<2><<<nbp, RB, 0, c->stream>>>(dJ, dS, c->api[S_MASS], c->pacc, c->nSlots, np, dP);
    c->stats.launches++;
    PCHK(cudaMemcpyAsync(paccOut, dP, np*(size_t)c->ndim*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  PCHK(cudaStreamSynchronize(c->stream));
  PCHK(cudaGetLastError());
#undef PCHK
  cleanup();
  return rc;
}
