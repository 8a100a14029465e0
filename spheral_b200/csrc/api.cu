// api.cu -- the C ABI of libsphb200 (include/sphb200.h): context life cycle, field movement, and the drivers that
// string the kernels of neighbors.cu / derivs.cu / energy.cu together.  No CPU fallback exists: every compute entry
// point runs CUDA kernels on the context's device or fails with an error.
#include "sphb200_internal.cuh"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

static std::string g_lastError;
static std::mutex g_errMutex;

int sphb200_fail(sphb200_ctx* c, const std::string& msg) {
  if (c) c->err = msg;
  std::lock_guard<std::mutex> lk(g_errMutex);
  g_lastError = msg;
  return 1;
}

namespace {

constexpr int RB = 256;

// sorted SoA (component-major) -> host AoS order, for downloads
__global__ void __launch_bounds__(RB) k_unpermute(const double* __restrict__ src, size_t cap, const uint32_t* __restrict__ perm,
                                                  size_t n, int width, double* __restrict__ dst) {
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  const size_t o = perm[s];
  for (int q = 0; q < width; ++q) dst[o*width + q] = src[(size_t)q*cap + s];
}
// host AoS (original order) -> component-major derivative array stored in ORIGINAL order (restart: permEval = identity)
__global__ void __launch_bounds__(RB) k_derivs_in(const double* __restrict__ src, size_t cap, size_t n, int width, double* __restrict__ dst) {
  const size_t o = (size_t)blockIdx.x*RB + threadIdx.x;
  if (o >= n) return;
  for (int q = 0; q < width; ++q) dst[(size_t)q*cap + o] = src[o*width + q];
}
__global__ void __launch_bounds__(RB) k_iota(uint32_t* __restrict__ p, size_t n) {
  const size_t o = (size_t)blockIdx.x*RB + threadIdx.x;
  if (o < n) p[o] = (uint32_t)o;
}
// DvDx (sorted SoA) -> api DvDxQ (AoS, original order)
__global__ void __launch_bounds__(RB) k_copy_dvdx(const double* __restrict__ src, size_t cap, const uint32_t* __restrict__ perm,
                                                  size_t n, size_t nLimit, int width, double* __restrict__ dst) {
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  const size_t o = perm[s];
  if (o >= nLimit) return;
  for (int q = 0; q < width; ++q) dst[o*width + q] = src[(size_t)q*cap + s];
}
__global__ void __launch_bounds__(RB) k_counts_by_orig(const uint32_t* __restrict__ nbrCount, const uint32_t* __restrict__ perm,
                                                       size_t n, uint32_t nInt, uint32_t* __restrict__ out) {
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s >= n) return;
  const uint32_t o = perm[s];
  if (o < nInt) out[o] = nbrCount[s];
}
// halo: gather listed nodes of one field into a staging block / scatter a block into a contiguous ghost range
// all masked fields of one side in ONE launch (a launch per field left the GPU idle between ~36 tiny operations per refresh)
struct HaloFields { const double* src[S_COUNT]; double* dst[S_COUNT]; int width[S_COUNT]; unsigned long long off[S_COUNT]; int nf; unsigned long long total; };
__global__ void __launch_bounds__(RB) k_halo_pack_all(HaloFields f, const uint32_t* __restrict__ nodes, size_t count, double* __restrict__ out) {
  const size_t t = (size_t)blockIdx.x*RB + threadIdx.x;
  if (t >= f.total) return;
  int a = 0;
  while (a + 1 < f.nf && t >= f.off[a + 1]) ++a;
  const size_t u = t - f.off[a];
  const int w = f.width[a];
  const size_t k = u/w; const int q = (int)(u - k*w);
  out[t] = f.src[a][(size_t)nodes[k]*w + q];
}
__global__ void __launch_bounds__(RB) k_halo_unpack_all(HaloFields f, size_t firstGhost, const double* __restrict__ in) {
  const size_t t = (size_t)blockIdx.x*RB + threadIdx.x;
  if (t >= f.total) return;
  int a = 0;
  while (a + 1 < f.nf && t >= f.off[a + 1]) ++a;
  const size_t u = t - f.off[a];
  f.dst[a][firstGhost*(size_t)f.width[a] + u] = in[t];
}
// slab halo selection: flags, then (after exclusive scans) ordered scatter of the node indices
// original index -> sorted slot
__global__ void __launch_bounds__(RB) k_inverse_perm(const uint32_t* __restrict__ perm, size_t n, uint32_t* __restrict__ inv) {
  const size_t s = (size_t)blockIdx.x*RB + threadIdx.x;
  if (s < n) inv[perm[s]] = (uint32_t)s;
}
// finalizeDerivatives over a domain boundary: {DvDt (ndim), DepsDt} of listed nodes, sorted component-major arrays <-> staging
// (DvDt block of count*ndim doubles, then DepsDt block of count doubles)
__global__ void __launch_bounds__(RB) k_halo_pack_derivs(const double* __restrict__ DvDt, const double* __restrict__ DepsDt, size_t cap, int ndim,
                                                         const uint32_t* __restrict__ inv, const uint32_t* __restrict__ nodes, size_t count,
                                                         double* __restrict__ out) {
  const size_t k = (size_t)blockIdx.x*RB + threadIdx.x;
  if (k >= count) return;
  const size_t s = inv[nodes[k]];
  for (int q = 0; q < ndim; ++q) out[k*ndim + q] = DvDt[(size_t)q*cap + s];
  out[count*ndim + k] = DepsDt[s];
}
__global__ void __launch_bounds__(RB) k_halo_unpack_derivs(double* __restrict__ DvDt, double* __restrict__ DepsDt, size_t cap, int ndim,
                                                           const uint32_t* __restrict__ inv, size_t first, size_t count,
                                                           const double* __restrict__ in) {
  const size_t k = (size_t)blockIdx.x*RB + threadIdx.x;
  if (k >= count) return;
  const size_t s = inv[first + k];
  for (int q = 0; q < ndim; ++q) DvDt[(size_t)q*cap + s] = in[k*ndim + q];
  DepsDt[s] = in[count*ndim + k];
}
__global__ void __launch_bounds__(RB) k_halo_flags(const double* __restrict__ pos, int ndim, int axis, size_t count, double lowCut, double highCut,
                                                   uint32_t* __restrict__ fLow, uint32_t* __restrict__ fHigh) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= count) return;
  const double x = pos[i*ndim + axis];
  fLow[i] = (x < lowCut) ? 1u : 0u;
  fHigh[i] = (x >= highCut) ? 1u : 0u;
}
// same, with the cut positions derived on the device: width = widthDev[axis] * (1 + 1e-9) (the all-reduced largest kernel
// extent; the factor matches the host variant's safety margin)
__global__ void __launch_bounds__(RB) k_halo_flags_dev(const double* __restrict__ pos, int ndim, int axis, size_t count, double lo, double hi,
                                                       const double* __restrict__ widthDev, uint32_t* __restrict__ fLow, uint32_t* __restrict__ fHigh) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= count) return;
  const double w = widthDev[axis]*(1.0 + 1.0e-9);
  const double x = pos[i*ndim + axis];
  fLow[i] = (x < lo + w) ? 1u : 0u;
  fHigh[i] = (x >= hi - w) ? 1u : 0u;
}
__global__ void k_halo_counts(const uint32_t* __restrict__ offLow, const uint32_t* __restrict__ offHigh, size_t count, size_t cap, long long* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { out[0] = (long long)offLow[count]; out[1] = (long long)offHigh[count]; out[2] = (long long)cap; }
}
__global__ void __launch_bounds__(RB) k_halo_scatter(const uint32_t* __restrict__ offLow, const uint32_t* __restrict__ offHigh, size_t count,
                                                     size_t cap, uint32_t* __restrict__ outLow, uint32_t* __restrict__ outHigh) {
  const size_t i = (size_t)blockIdx.x*RB + threadIdx.x;
  if (i >= count) return;
  if (offLow[i + 1] != offLow[i] && offLow[i] < cap) outLow[offLow[i]] = (uint32_t)i;
  if (offHigh[i + 1] != offHigh[i] && offHigh[i] < cap) outHigh[offHigh[i]] = (uint32_t)i;
}
// dependent-chain FP64 FMA throughput probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x*1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int k = 0; k < iters; ++k) {
    x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
  }
  const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
  if (s == 12345.678) out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

// (Re)size the per-node arrays.  The host-ordered state fields keep their contents when the capacity grows (the ghost
// count changes from step to step while the internal state lives on the device); everything else is scratch.
int alloc_nodes(sphb200_ctx* c, size_t n) {
  if (n <= c->cap) return 0;
  const size_t cap = n + n/32 + 32;
  const size_t keep = c->n;                     // nodes whose state must survive
  const bool crk = c->opt.hydro == SPHB200_HYDRO_CRKSPH;
  for (int s = 0; s < S_COUNT; ++s) {
    if (!crk && (s == S_VOLUME || s == S_RKCORR)) continue;       // CRKSPH-only fields
    const size_t w = (size_t)sphb200_state_width(c->ndim, s);
    double* p = nullptr;
    CU_CHECK(c, cudaMalloc((void**)&p, cap*w*sizeof(double)));
    if (c->api[s] && c->have[s] && keep)
      CU_CHECK(c, cudaMemcpyAsync(p, c->api[s], keep*w*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    else c->have[s] = false;
    if (c->api[s]) { CU_CHECK(c, cudaStreamSynchronize(c->stream)); cudaFree(c->api[s]); }
    c->api[s] = p;
  }
  auto reall = [&](double*& p, size_t cnt) -> int {
    if (p) cudaFree(p);
    p = nullptr;
    CU_CHECK(c, cudaMalloc((void**)&p, cnt*sizeof(double)));
    return 0;
  };
  if (reall(c->rows, cap*(size_t)(c->ndim == 3 ? 16 : 12))) return 1;
  if (reall(c->aux2, cap*2)) return 1;
  if (crk && (reall(c->crkVolS, cap) || reall(c->crkCorrS, cap*(size_t)(c->ndim == 3 ? 16 : 10)) ||
              reall(c->crkQS, cap*(size_t)(c->ndim == 3 ? 10 : 6)) || reall(c->crkAux, cap*2))) return 1;
  // node-wise derivatives survive a growth of the node set (the ghost count changes from step to step while the integrator
  // still needs the previous evaluation): component c of a field moves from stride capEval to stride cap
  const bool keepDerivs = c->derivNodeValid && c->nEval > 0;
  for (int s = 0; s < DV_COUNT; ++s) {
    const int w = sphb200_deriv_width(c->ndim, s);
    double* p = nullptr;
    CU_CHECK(c, cudaMalloc((void**)&p, cap*(size_t)w*sizeof(double)));
    if (keepDerivs && c->deriv[s])
      for (int q = 0; q < w; ++q)
        CU_CHECK(c, cudaMemcpyAsync(p + (size_t)q*cap, c->deriv[s] + (size_t)q*c->capEval, c->nEval*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    if (c->deriv[s]) { CU_CHECK(c, cudaStreamSynchronize(c->stream)); cudaFree(c->deriv[s]); }
    c->deriv[s] = p;
  }
  if (keepDerivs) c->capEval = cap;
  auto reall32 = [&](uint32_t*& p, size_t cnt) -> int {
    if (p) cudaFree(p);
    p = nullptr;
    CU_CHECK(c, cudaMalloc((void**)&p, cnt*sizeof(uint32_t)));
    return 0;
  };
  if (reall32(c->cellKeyApi, cap) || reall32(c->perm, cap) || reall32(c->skey, cap) || reall32(c->nbrCount, cap)) return 1;
  const size_t nt = (cap + SPHB200_TILE - 1)/SPHB200_TILE + 1;
  if (reall32(c->tileRows, nt) || reall32(c->tileRunStart, nt) || reall32(c->tileRunCount, nt)) return 1;
  if (c->tileOff) cudaFree(c->tileOff);
  c->tileOff = nullptr;
  CU_CHECK(c, cudaMalloc((void**)&c->tileOff, (nt + 1)*sizeof(unsigned long long)));
  // aux arrays are sized lazily from cap
  for (double** p : {&c->auxPneg, &c->auxSomr2, &c->auxDvDxQ, &c->auxfCl, &c->auxfCq}) { if (*p) cudaFree(*p); *p = nullptr; }
  if (c->frows) { cudaFree(c->frows); c->frows = nullptr; c->frowsCap = 0; }
  c->cap = cap;
  c->sortValid = c->rowsValid = c->pairsValid = c->derivsValid = false;
  if (!keepDerivs) c->derivNodeValid = false;
  return 0;
}

}  // namespace

int sphb200_join_uploads(sphb200_ctx* c, bool all) {
  if (c->pendGeomUp) { CU_CHECK(c, cudaStreamWaitEvent(c->stream, c->evGeomUp, 0)); c->pendGeomUp = false; }
  if (all && c->pendRestUp) { CU_CHECK(c, cudaStreamWaitEvent(c->stream, c->evRestUp, 0)); c->pendRestUp = false; }
  if (all && c->ghostRefillPending) {
    // plane ghosts were generated while the non-geometric fields of their control nodes were still on their way from the host
    // (sphb200_reflect_set_ghost_nodes waits for positions and H only, so that the neighbour build overlaps the rest of the upload):
    // now that everything has landed, give those ghosts the values of their controls
    c->ghostRefillPending = false;
    if (sphb200_reflect_apply_ghosts(c, ~((1u << S_POS) | (1u << S_H)))) return 1;
  }
  return 0;
}

namespace {

int ensure_stage(sphb200_ctx* c, size_t bytes) {
  if (bytes <= c->stageBytes) return 0;
  if (c->stage) cudaFree(c->stage);
  c->stage = nullptr; c->stageBytes = 0;
  CU_CHECK(c, cudaMalloc((void**)&c->stage, bytes + bytes/8));
  c->stageBytes = bytes + bytes/8;
  return 0;
}

}  // namespace

extern "C" {

int sphb200_abi_version(void) { return SPHB200_ABI_VERSION; }

const char* sphb200_last_error(const sphb200_ctx* c) {
  if (c) return c->err.c_str();
  return g_lastError.c_str();
}

static int check_options(sphb200_ctx* c, const sphb200_options* o) {
  if (!o) return sphb200_fail(c, "null options");
  if (o->ndim != 2 && o->ndim != 3) return sphb200_fail(c, "ndim must be 2 or 3");
  // SPH.cc:97-98
  if (o->compatibleEnergy && o->evolveTotalEnergy)
    return sphb200_fail(c, "SPH error : you cannot simultaneously use both compatibleEnergyEvolution and evolveTotalEnergy");
  if (!(o->nPerh > 0.0)) return sphb200_fail(c, "nPerh must be positive");
  if (o->Qkind != SPHB200_Q_MG && o->Qkind != SPHB200_Q_LIMITED_MG) return sphb200_fail(c, "unknown Qkind");
  if (o->hEvolution < SPHB200_H_SPH || o->hEvolution > SPHB200_H_ASPH_CLASSIC) return sphb200_fail(c, "unknown hEvolution");
  if (o->hEvolution == SPHB200_H_ASPH_CLASSIC) {
    if (!(o->hminratio > 0.0) || !(o->hmin > 0.0) || !(o->hmax > 0.0)) return sphb200_fail(c, "SPHB200_H_ASPH_CLASSIC needs positive hmin, hmax and hminratio");
  }
  if (o->hydro != SPHB200_HYDRO_SPH && o->hydro != SPHB200_HYDRO_CRKSPH) return sphb200_fail(c, "unknown hydro");
  return 0;
}

int sphb200_create(sphb200_ctx** out, int device, const sphb200_options* opts) {
  if (!out) return sphb200_fail(nullptr, "null ctx pointer");
  *out = nullptr;
  if (check_options(nullptr, opts)) return 1;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return sphb200_fail(nullptr, std::string("no CUDA device available (libsphb200 has no CPU fallback): ") + cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return sphb200_fail(nullptr, "device index out of range");
  sphb200_ctx* c = new sphb200_ctx();
  c->device = device; c->ndim = opts->ndim; c->opt = *opts;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return sphb200_fail(nullptr, "cudaSetDevice/cudaStreamCreate failed");
  }
  for (auto& ev : c->ev) cudaEventCreate(&ev);
  cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&c->evGeomUp, cudaEventDisableTiming); cudaEventCreateWithFlags(&c->evRestUp, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&c->evMainMark, cudaEventDisableTiming);
  cudaMalloc((void**)&c->reduceBuf, (296*12 + 16)*sizeof(double));
  cudaMallocHost((void**)&c->reduceHost, 16*sizeof(double));
  cudaMalloc((void**)&c->counters, 16*sizeof(unsigned long long));      // [0,8) neighbour build, [8] anisotropy flag of k_pack
  cudaMemset(c->counters, 0, 16*sizeof(unsigned long long));
  cudaMalloc((void**)&c->dilTab, 3*SPHB200_DIL*sizeof(uint32_t));
  cudaMallocHost((void**)&c->countersHost, 16*sizeof(unsigned long long));
  cudaMalloc((void**)&c->chunkCount, SPHB200_MAX_CHUNKS*sizeof(uint32_t));
  cudaMallocHost((void**)&c->chunkCountHost, SPHB200_MAX_CHUNKS*sizeof(uint32_t));
  for (auto& ev : c->evChunk) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
  if (cudaGetLastError() != cudaSuccess) { sphb200_destroy(c); return sphb200_fail(nullptr, "context allocation failed"); }
  *out = c;
  return 0;
}

void sphb200_destroy(sphb200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaStreamDestroy(c->copyStream); }
  if (c->stream) cudaStreamSynchronize(c->stream);
  for (cudaEvent_t ev : {c->evGeomUp, c->evRestUp, c->evMainMark}) if (ev) cudaEventDestroy(ev);
  for (int s = 0; s < S_COUNT; ++s) cudaFree(c->api[s]);
  for (int s = 0; s < DV_COUNT; ++s) cudaFree(c->deriv[s]);
  for (void* p : {(void*)c->W.coef, (void*)c->W.nperhVals, (void*)c->WQ.coef, (void*)c->WQ.nperhVals, (void*)c->cellKeyApi, (void*)c->cellStart,
                  (void*)c->cellCursor, (void*)c->perm, (void*)c->skey, (void*)c->reduceBuf, (void*)c->rows, (void*)c->aux2, (void*)c->auxPneg, (void*)c->auxSomr2,
                  (void*)c->auxDvDxQ, (void*)c->auxfCl, (void*)c->auxfCq, (void*)c->nbrCount, (void*)c->tileRows, (void*)c->tileOff, (void*)c->nbr,
                  (void*)c->counters, (void*)c->frows, (void*)c->scanTmp, (void*)c->pacc, (void*)c->stage,
                  (void*)c->runs, (void*)c->tileRunStart, (void*)c->tileRunCount, (void*)c->dilTab, (void*)c->crkVolS, (void*)c->crkCorrS, (void*)c->crkQS, (void*)c->crkAux, (void*)c->permEval, (void*)c->dtCand, (void*)c->dtAux}) cudaFree(p);
  for (int s = 0; s < S_COUNT; ++s) if (c->api0[s]) cudaFree(c->api0[s]);
  for (int p = 0; p < SPHB200_MAX_PLANES; ++p) if (c->planeCtl[p]) cudaFree(c->planeCtl[p]);
  if (c->invPerm) cudaFree(c->invPerm);
  if (c->hDone) cudaFree(c->hDone);
  if (c->cellReach) cudaFree(c->cellReach);
  if (c->tileRadius) cudaFree(c->tileRadius);
  cudaFreeHost(c->reduceHost); cudaFreeHost(c->countersHost);
  if (c->chunkList) cudaFree(c->chunkList);
  if (c->chunkCount) cudaFree(c->chunkCount);
  if (c->chunkCountHost) cudaFreeHost(c->chunkCountHost);
  for (auto& ev : c->evChunk) if (ev) cudaEventDestroy(ev);
  for (auto& ev : c->ev) if (ev) cudaEventDestroy(ev);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int sphb200_set_options(sphb200_ctx* c, const sphb200_options* o) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (check_options(c, o)) return 1;
  if (o->ndim != c->ndim) return sphb200_fail(c, "ndim cannot change after creation");
  if (o->hydro != c->opt.hydro) return sphb200_fail(c, "hydro (SPH | CRKSPH) cannot change after creation");
  const bool repack = (o->epsTensile != c->opt.epsTensile) || (o->Qkind != c->opt.Qkind) || (o->balsara != c->opt.balsara);
  c->opt = *o;
  if (repack) c->rowsValid = false;
  return 0;
}

int sphb200_sync(sphb200_ctx* c) {
  if (c && c->copyStream) cudaStreamSynchronize(c->copyStream);
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  CU_CHECK(c, cudaGetLastError());
  return 0;
}

void* sphb200_stream(sphb200_ctx* c) { return c ? (void*)c->stream : nullptr; }

int sphb200_set_kernel_table(sphb200_ctx* c, int which, double kext, double xmin, double xstep, size_t n1,
                             const double* Wc, const double* gWc, const double* /*g2Wc*/,
                             size_t nperhN, double nperhXmin, double nperhXmax, const double* nperhVals,
                             size_t /*wsumN*/, double /*wsumXmin*/, double /*wsumXmax*/, const double* /*wsumVals*/) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (which != SPHB200_TABLE_W && which != SPHB200_TABLE_WPI) return sphb200_fail(c, "set_kernel_table: bad table id");
  if (!Wc || !gWc || !(kext > 0.0) || !(xstep > 0.0)) return sphb200_fail(c, "set_kernel_table: bad arguments");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  TableDev& t = (which == SPHB200_TABLE_W) ? c->W : c->WQ;
  const size_t nint = n1 + 1;
  std::vector<double> inter(6*nint);
  for (size_t k = 0; k < nint; ++k) for (int q = 0; q < 3; ++q) { inter[6*k + q] = Wc[3*k + q]; inter[6*k + 3 + q] = gWc[3*k + q]; }
  if (t.coef) cudaFree(t.coef);
  t.coef = nullptr;
  CU_CHECK(c, cudaMalloc((void**)&t.coef, inter.size()*sizeof(double)));
  CU_CHECK(c, cudaMemcpyAsync(t.coef, inter.data(), inter.size()*sizeof(double), cudaMemcpyHostToDevice, c->stream));
  t.kext = kext; t.xmin = xmin; t.xstep = xstep; t.n1 = (uint32_t)n1;
  t.hostW.assign(Wc, Wc + 3*nint); t.hostG.assign(gWc, gWc + 3*nint);
  if (t.nperhVals) { cudaFree(t.nperhVals); t.nperhVals = nullptr; }
  t.nperhN = 0;
  if (nperhVals && nperhN >= 2) {
    CU_CHECK(c, cudaMalloc((void**)&t.nperhVals, 2*nperhN*sizeof(double)));
    CU_CHECK(c, cudaMemcpyAsync(t.nperhVals, nperhVals, 2*nperhN*sizeof(double), cudaMemcpyHostToDevice, c->stream));
    t.nperhN = (uint32_t)nperhN; t.nperhXmin = nperhXmin; t.nperhXmax = nperhXmax; t.nperhXstep = (nperhXmax - nperhXmin)/double(nperhN - 1);
  }
  CU_CHECK(c, cudaStreamSynchronize(c->stream));     // inter is a local buffer
  t.set = true;
  // oneKernel = (W == WQ), SPH.cc:185 / TableKernelViewInline.hh:104-112
  c->oneKernel = !c->WQ.set || (c->WQ.kext == c->W.kext && c->WQ.n1 == c->W.n1 && c->WQ.xstep == c->W.xstep &&
                                c->WQ.hostW == c->W.hostW && c->WQ.hostG == c->W.hostG);
  c->pairsValid = false;
  return 0;
}

int sphb200_set_nodes(sphb200_ctx* c, size_t nInternal, size_t nGhost) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device));
  const size_t n = nInternal + nGhost;
  if (n >= 0x7fffffffull) return sphb200_fail(c, "set_nodes: more than 2^31-1 nodes per GPU is not supported");
  if (n > c->cap) {      // reallocation frees arrays that in-flight kernels AND in-flight uploads may use: drain both streams first
    const bool refill = c->ghostRefillPending; c->ghostRefillPending = false;     // the ghosts are being redefined anyway
    if (sphb200_join_uploads(c, true)) return 1;
    c->ghostRefillPending = refill;
    CU_CHECK(c, cudaStreamSynchronize(c->stream));
    CU_CHECK(c, cudaStreamSynchronize(c->copyStream));
  }   // otherwise only the node counts change: uploads still in flight on the copy stream keep landing in the same arrays
  if (alloc_nodes(c, n)) return 1;
  if (n != c->n || nInternal != c->nInt) { c->sortValid = c->rowsValid = c->pairsValid = c->derivsValid = false; }
  c->nInt = nInternal; c->nGhost = nGhost; c->n = n;
  return 0;
}

static const double* state_ptr(const sphb200_host_state* s, int slot) {
  switch (slot) {
    case S_POS: return s->position; case S_VEL: return s->velocity; case S_H: return s->H; case S_MASS: return s->mass;
    case S_RHO: return s->massDensity; case S_EPS: return s->specificThermalEnergy; case S_P: return s->pressure;
    case S_CS: return s->soundSpeed; case S_OMEGA: return s->omegaGradh; case S_DVDXQ: return s->DvDxQ;
    case S_FCL: return s->fCl; case S_FCQ: return s->fCq;
    case S_VOLUME: return s->volume; case S_RKCORR: return s->rkCorrections;
  }
  return nullptr;
}

static int upload_state_impl(sphb200_ctx* c, unsigned mask, const sphb200_host_state* s, bool keepConnectivity);
int sphb200_upload_state(sphb200_ctx* c, unsigned mask, const sphb200_host_state* s) { return upload_state_impl(c, mask, s, false); }
int sphb200_upload_state_values(sphb200_ctx* c, unsigned mask, const sphb200_host_state* s) { return upload_state_impl(c, mask, s, true); }

static int upload_state_impl(sphb200_ctx* c, unsigned mask, const sphb200_host_state* s, bool keepConnectivity) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!s) return sphb200_fail(c, "upload_state: null state");
  CU_CHECK(c, cudaSetDevice(c->device));
  for (int slot = 0; slot < S_COUNT; ++slot)
    if ((mask & (1u << slot)) && !state_ptr(s, slot)) return sphb200_fail(c, "upload_state: field selected in mask but pointer is null");
  // the copy stream starts after everything already queued on the main stream (kernels there may still read the arrays)
  // and after its own earlier copies; positions and H go first so that build_pairs can start while the rest is in flight
  CU_CHECK(c, cudaEventRecord(c->evMainMark, c->stream));
  CU_CHECK(c, cudaStreamWaitEvent(c->copyStream, c->evMainMark, 0));
  for (int pass = 0; pass < 2; ++pass) {
    bool any = false;
    for (int slot = 0; slot < S_COUNT; ++slot) {
      if (!(mask & (1u << slot))) continue;
      const bool geom = (slot == S_POS || slot == S_H);
      if (geom != (pass == 0)) continue;
      const double* src = state_ptr(s, slot);
      if (!c->api[slot] && c->n) return sphb200_fail(c, "upload_state: volume / RK corrections exist only in CRKSPH contexts");
      const size_t bytes = c->n*(size_t)sphb200_state_width(c->ndim, slot)*sizeof(double);
      if (bytes) CU_CHECK(c, cudaMemcpyAsync(c->api[slot], src, bytes, cudaMemcpyHostToDevice, c->copyStream));
      any = true;
      c->have[slot] = true;
      if (geom && !keepConnectivity) { c->sortValid = false; c->pairsValid = false; }
      if (slot != S_EPS && slot != S_VOLUME && slot != S_RKCORR) c->rowsValid = false;
    }
    if (any && pass == 0) { CU_CHECK(c, cudaEventRecord(c->evGeomUp, c->copyStream)); c->pendGeomUp = true; }
    if (any && pass == 1) { CU_CHECK(c, cudaEventRecord(c->evRestUp, c->copyStream)); c->pendRestUp = true; }
  }
  return 0;
}

int sphb200_download_state(sphb200_ctx* c, unsigned mask, double* const* fields) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  for (int slot = 0; slot < S_COUNT; ++slot) {
    if (!(mask & (1u << slot))) continue;
    if (!fields[slot]) return sphb200_fail(c, "download_state: null destination");
    if (!c->have[slot]) return sphb200_fail(c, "download_state: field was never uploaded");
    const size_t bytes = c->n*(size_t)sphb200_state_width(c->ndim, slot)*sizeof(double);
    if (bytes) CU_CHECK(c, cudaMemcpyAsync(fields[slot], c->api[slot], bytes, cudaMemcpyDeviceToHost, c->stream));
  }
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

int sphb200_build_pairs(sphb200_ctx* c, size_t* npairs) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device));
  // only positions and H are needed here: the other state fields may still be in flight on the copy stream (k_pack then reads
  // stale values of those, which is harmless: the rows are marked stale and re-packed once the fields have landed)
  if (sphb200_join_uploads(c, false)) return 1;
  const bool restInFlight = c->pendRestUp;
  if (c->n == 0) { c->npairs = c->nEdges = c->nSlots = 0; c->nTiles = 0; c->pairsValid = true; c->sortValid = true; c->rowsValid = true; if (npairs) *npairs = 0; return 0; }
  cudaEventRecord(c->ev[0], c->stream);
  if (sphb200_sort_and_pack(c)) return 1;
  cudaEventRecord(c->ev[1], c->stream);
  int nrc = sphb200_neighbors(c);
  if (nrc == 2) {                                    // fine grid + wide stencil produced too many runs for a tile: wide cells instead
    c->forceR1 = true;
    if (sphb200_sort_and_pack(c)) return 1;
    cudaEventRecord(c->ev[1], c->stream);
    nrc = sphb200_neighbors(c);
  }
  if (nrc) return 1;
  cudaEventRecord(c->ev[2], c->stream);
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->stats.ms_build_pairs, c->ev[0], c->ev[2]);
  cudaEventElapsedTime(&c->stats.ms_neighbor_kernels, c->ev[1], c->ev[2]);
  c->derivsValid = false;
  if (restInFlight) c->rowsValid = false;
  if (npairs) *npairs = c->npairs;
  return 0;
}

int sphb200_connectivity_valid(const sphb200_ctx* c) { return (c && c->pairsValid) ? 1 : 0; }

unsigned sphb200_state_fields_present(const sphb200_ctx* c) {
  unsigned m = 0;
  if (c) for (int s = 0; s < S_COUNT; ++s) if (c->have[s] && c->api[s]) m |= 1u << s;
  return m;
}

int sphb200_download_pairs(sphb200_ctx* c, uint32_t* i, uint32_t* j, size_t cap) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!c->pairsValid) return sphb200_fail(c, "download_pairs: no valid pair list (call build_pairs after changing position/H)");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->npairs == 0) return 0;
  return sphb200_pairs_to_host(c, i, j, cap, nullptr, 0);
}

int sphb200_download_pair_accelerations(sphb200_ctx* c, double* pacc, size_t cap) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!c->pairsValid || !c->derivsValid) return sphb200_fail(c, "download_pair_accelerations: derivatives have not been evaluated for the current pair list");
  if (!c->opt.compatibleEnergy) return sphb200_fail(c, "download_pair_accelerations: pair-wise accelerations exist only with compatibleEnergyEvolution (SPH.cc:129-133)");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->npairs == 0) return 0;
  return sphb200_pairs_to_host(c, nullptr, nullptr, 0, pacc, cap);
}

int sphb200_download_neighbor_counts(sphb200_ctx* c, uint32_t* counts) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!c->pairsValid) return sphb200_fail(c, "download_neighbor_counts: no valid pair list");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->nInt == 0) return 0;
  if (ensure_stage(c, c->nInt*sizeof(uint32_t))) return 1;
  k_counts_by_orig<<<(unsigned)((c->n + RB - 1)/RB), RB, 0, c->stream>>>(c->nbrCount, c->perm, c->n, (uint32_t)c->nInt, (uint32_t*)c->stage);
  KERNEL_CHECK(c, "k_counts_by_orig");
  CU_CHECK(c, cudaMemcpyAsync(counts, c->stage, c->nInt*sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// checks of evaluateDerivatives + the row pack; returns 2 when there is nothing to evaluate (no nodes)
static int evaluate_prepare(sphb200_ctx* c) {
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->pairsValid) return sphb200_fail(c, "evaluateDerivatives: connectivity is stale or missing (requireConnectivity: call build_pairs first)");
  if (c->n == 0) { c->derivsValid = true; return 2; }
  const bool crk = c->opt.hydro == SPHB200_HYDRO_CRKSPH;
  for (int s : {S_POS, S_VEL, S_H, S_MASS, S_RHO, S_P, S_CS})
    if (!c->have[s]) return sphb200_fail(c, "evaluateDerivatives: required state field missing on device (position, velocity, H, mass, mass density, pressure, sound speed)");
  if (!crk && !c->have[S_OMEGA]) return sphb200_fail(c, "evaluateDerivatives: required state field missing on device (grad h corrections)");
  if (crk && (!c->have[S_VOLUME] || !c->have[S_RKCORR]))
    return sphb200_fail(c, "evaluateDerivatives: CRKSPH needs the volume and the RK corrections (call crk_compute_volume / crk_compute_corrections or upload them)");
  if (!c->W.set) return sphb200_fail(c, "evaluateDerivatives: kernel table not set");
  cudaEventRecord(c->ev[3], c->stream);
  if (!c->rowsValid && sphb200_pack_rows(c)) return 1;
  cudaEventRecord(c->ev[4], c->stream);
  return 0;
}
static int evaluate_finish(sphb200_ctx* c) {
  cudaEventRecord(c->ev[5], c->stream);
  // keep the sorted order these node-wise derivatives are stored in (see permEval)
  if (sphb200_ensure(c, c->permEval, c->permEvalCap, c->cap)) return 1;
  CU_CHECK(c, cudaMemcpyAsync(c->permEval, c->perm, c->n*sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  c->nEval = c->n; c->capEval = c->cap; c->nIntEval = c->nInt; c->derivNodeValid = true;
  return 0;
}

int sphb200_evaluate_derivatives(sphb200_ctx* c, double /*time*/, double /*dt*/) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  const int rc = evaluate_prepare(c);
  if (rc) return rc == 2 ? 0 : 1;
  if ((c->opt.hydro == SPHB200_HYDRO_CRKSPH) ? sphb200_launch_crk_derivs(c) : sphb200_launch_derivs(c)) return 1;
  if (c->opt.hEvolution == SPHB200_H_ASPH_CLASSIC && sphb200_launch_asph_classic(c)) return 1;
  return evaluate_finish(c);
}

static double* deriv_ptr(const sphb200_host_derivs* d, int slot) {
  switch (slot) {
    case DV_DXDT: return d->DxDt; case DV_DRHODT: return d->DrhoDt; case DV_DVDT: return d->DvDt; case DV_DEPSDT: return d->DepsDt;
    case DV_DVDX: return d->DvDx; case DV_LOCALDVDX: return d->localDvDx; case DV_GRADRHO: return d->gradRho; case DV_M: return d->M;
    case DV_LOCALM: return d->localM; case DV_RHOSUM: return d->rhoSum; case DV_NORM: return d->normalization;
    case DV_MAXQ: return d->maxViscousPressure; case DV_EFFQ: return d->effViscousPressure; case DV_XSPHW: return d->XSPHWeightSum;
    case DV_XSPHDV: return d->XSPHDeltaV; case DV_DHDT: return d->DHDt; case DV_HIDEAL: return d->Hideal;
    case DV_M0: return d->massZerothMoment; case DV_M1: return d->massFirstMoment;
  }
  return nullptr;
}

int sphb200_download_derivs(sphb200_ctx* c, unsigned mask, const sphb200_host_derivs* d) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!d) return sphb200_fail(c, "download_derivs: null destination struct");
  if (!c->derivsValid) return sphb200_fail(c, "download_derivs: derivatives have not been evaluated");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->n == 0) return 0;
  size_t total = 0;
  for (int s = 0; s < DV_COUNT; ++s) if (mask & (1u << s)) total += c->n*(size_t)sphb200_deriv_width(c->ndim, s);
  if (ensure_stage(c, total*sizeof(double))) return 1;
  size_t off = 0;
  const unsigned nb = (unsigned)((c->n + RB - 1)/RB);
  for (int s = 0; s < DV_COUNT; ++s) {
    if (!(mask & (1u << s))) continue;
    double* dst = deriv_ptr(d, s);
    if (!dst) return sphb200_fail(c, "download_derivs: field selected in mask but pointer is null");
    const int w = sphb200_deriv_width(c->ndim, s);
    k_unpermute<<<nb, RB, 0, c->stream>>>(c->deriv[s], c->cap, c->perm, c->n, w, c->stage + off);
    KERNEL_CHECK(c, "k_unpermute");
    CU_CHECK(c, cudaMemcpyAsync(dst, c->stage + off, c->n*(size_t)w*sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    off += c->n*(size_t)w;
  }
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// Tiles of each chunk of the host index range: chunk q = internal nodes with original index in [q*chunkSize, (q+1)*chunkSize) (the
// last chunk takes the remainder and the ghost nodes, whose derivative entries are zeros).  A tile that holds nodes of several chunks
// is listed in each of them.  The order inside a list is whatever the atomics give (roughly ascending); results do not depend on it.
constexpr int CL_THREADS = 1024;                 // 32 tiles per block: one global atomic per (block, chunk) instead of one per (tile, chunk)
__global__ void __launch_bounds__(CL_THREADS) k_chunk_lists(const uint32_t* __restrict__ perm, size_t n, uint32_t nInt, uint32_t chunkSize, int Q,
                                                            size_t nTiles, uint32_t* __restrict__ lists, uint32_t* __restrict__ counts) {
  __shared__ uint32_t cnt[SPHB200_MAX_CHUNKS], base[SPHB200_MAX_CHUNKS];
  if (threadIdx.x < SPHB200_MAX_CHUNKS) cnt[threadIdx.x] = 0u;
  __syncthreads();
  const size_t tile = ((size_t)blockIdx.x*CL_THREADS + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const size_t i = tile*SPHB200_TILE + lane;
  unsigned m = 0u;
  if (tile < nTiles && i < n) {
    const uint32_t o = perm[i];
    const uint32_t q = (o < nInt) ? min((uint32_t)(Q - 1), o/chunkSize) : (uint32_t)(Q - 1);
    m = 1u << q;
  }
  m = __reduce_or_sync(0xffffffffu, m);
  uint32_t pos[SPHB200_MAX_CHUNKS];
  if (lane == 0)
    for (unsigned r = m; r; r &= r - 1u) { const int q = __ffs(r) - 1; pos[q] = atomicAdd(&cnt[q], 1u); }
  __syncthreads();
  if (threadIdx.x < (unsigned)Q && cnt[threadIdx.x]) base[threadIdx.x] = atomicAdd(&counts[threadIdx.x], cnt[threadIdx.x]);
  __syncthreads();
  if (lane == 0)
    for (unsigned r = m; r; r &= r - 1u) { const int q = __ffs(r) - 1; lists[(size_t)q*nTiles + base[q] + pos[q]] = (uint32_t)tile; }
}
// sorted SoA (component-major) -> host AoS order for the original indices [lo, hi): one thread per output element (coalesced writes)
__global__ void __launch_bounds__(RB) k_unpermute_range(const double* __restrict__ src, size_t cap, const uint32_t* __restrict__ invPerm,
                                                        size_t lo, size_t hi, int width, double* __restrict__ dst) {
  const size_t t = (size_t)blockIdx.x*RB + threadIdx.x;
  if (t >= (hi - lo)*(size_t)width) return;
  const size_t o = lo + t/(size_t)width;
  const int q = (int)(t - (o - lo)*(size_t)width);
  dst[o*(size_t)width + q] = src[(size_t)q*cap + invPerm[o]];
}

// SPHB200_E2H_CHUNKS: 1 = no pipelining; default 4 chunks, 8 from 4 M internal nodes on (measured: 1 M 9.49 -> 8.45 ms with 4, 9.01
// with 8; 8 M 68.2 -> 59.2 with 4, 57.9 with 8 -- the download of the last chunk is what cannot be hidden)
static int e2h_chunks_wanted(size_t nInt) {
  const char* e = getenv("SPHB200_E2H_CHUNKS");
  const int q = e ? atoi(e) : (nInt >= ((size_t)1 << 22) ? 8 : 4);
  return q < 1 ? 1 : (q > SPHB200_MAX_CHUNKS ? SPHB200_MAX_CHUNKS : q);
}

int sphb200_evaluate_derivatives_to_host(sphb200_ctx* c, double time, double dt, unsigned mask, const sphb200_host_derivs* d) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!d) return sphb200_fail(c, "evaluate_derivatives_to_host: null destination struct");
  for (int s = 0; s < DV_COUNT; ++s)
    if ((mask & (1u << s)) && !deriv_ptr(d, s)) return sphb200_fail(c, "evaluate_derivatives_to_host: field selected in mask but pointer is null");
  int Q = e2h_chunks_wanted(c->nInt);
  const char* force = getenv("SPHB200_E2H_FORCE");          // tests: chunk whatever the size / the coherence of the host order
  const bool forced = force && atoi(force) != 0;
  // (the classic ASPH ideal H is a second loop over all nodes after the pair loop: no chunk is complete before it has run)
  if (c->opt.hydro == SPHB200_HYDRO_CRKSPH || c->opt.hEvolution == SPHB200_H_ASPH_CLASSIC || (!forced && c->nInt < (size_t)262144) || c->nInt < (size_t)Q) Q = 1;
  if (Q > 1) {
    const int rc = evaluate_prepare(c);
    if (rc == 1) return 1;
    if (rc == 2) return 0;
    const uint32_t chunkSize = (uint32_t)(c->nInt/(size_t)Q);
    if (!c->chunkListsValid || c->chunkQ != Q) {
      if (sphb200_ensure(c, c->chunkList, c->chunkListCap, (size_t)SPHB200_MAX_CHUNKS*(c->cap/SPHB200_TILE + 2))) return 1;
      CU_CHECK(c, cudaMemsetAsync(c->chunkCount, 0, SPHB200_MAX_CHUNKS*sizeof(uint32_t), c->stream));
      k_chunk_lists<<<(unsigned)((c->nTiles*32 + CL_THREADS - 1)/CL_THREADS), CL_THREADS, 0, c->stream>>>(c->perm, c->n, (uint32_t)c->nInt, chunkSize, Q, c->nTiles, c->chunkList, c->chunkCount);
      KERNEL_CHECK(c, "k_chunk_lists");
      CU_CHECK(c, cudaMemcpyAsync(c->chunkCountHost, c->chunkCount, SPHB200_MAX_CHUNKS*sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      CU_CHECK(c, cudaStreamSynchronize(c->stream));
      c->chunkListsValid = true; c->chunkQ = Q;
    }
    size_t listed = 0;
    for (int q = 0; q < Q; ++q) listed += c->chunkCountHost[q];
    // a host order without spatial coherence puts every tile into every list: Q times the work.  Then one launch and one download.
    if (!forced && listed > c->nTiles + c->nTiles/4 + (size_t)Q) Q = 1;
    if (Q == 1) {
      if (sphb200_launch_derivs(c)) return 1;
      if (evaluate_finish(c)) return 1;
      return sphb200_download_derivs(c, mask, d);
    }
    if (sphb200_inverse_perm(c)) return 1;
    size_t total = 0;
    for (int s = 0; s < DV_COUNT; ++s) if (mask & (1u << s)) total += c->n*(size_t)sphb200_deriv_width(c->ndim, s);
    if (ensure_stage(c, total*sizeof(double))) return 1;
    // Without XSPH DxDt of an internal node IS its velocity (SPH.cc:503-509), which sits on the device in host order already: it
    // leaves first, while the first chunk is computed (the ghost entries -- zeros -- follow with the last chunk)
    const bool dxdtEarly = (mask & (1u << DV_DXDT)) && !c->opt.XSPH && c->have[S_VEL];
    if (dxdtEarly) CU_CHECK(c, cudaEventRecord(c->evMainMark, c->stream));      // the velocity is final here (uploads joined, no kernel of this call writes it)
    // every chunk is enqueued before the first copy: a copy into pageable host memory blocks the calling thread until it is done, and
    // must not keep the next chunk from being launched
    for (int q = 0; q < Q; ++q) {
      const uint32_t lo = (uint32_t)q*chunkSize, hi = (q == Q - 1) ? 0xffffffffu : (uint32_t)(q + 1)*chunkSize;
      if (sphb200_launch_derivs_chunk(c, c->chunkList + (size_t)q*c->nTiles, c->chunkCountHost[q], lo, hi)) return 1;
      CU_CHECK(c, cudaEventRecord(c->evChunk[q], c->stream));
    }
    c->derivsValid = true;
    if (evaluate_finish(c)) return 1;
    if (dxdtEarly) {
      CU_CHECK(c, cudaStreamWaitEvent(c->copyStream, c->evMainMark, 0));
      CU_CHECK(c, cudaMemcpyAsync(deriv_ptr(d, DV_DXDT), c->api[S_VEL], c->nInt*(size_t)c->ndim*sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
    }
    for (int q = 0; q < Q; ++q) {
      const uint32_t lo = (uint32_t)q*chunkSize, hi = (q == Q - 1) ? 0xffffffffu : (uint32_t)(q + 1)*chunkSize;
      CU_CHECK(c, cudaStreamWaitEvent(c->copyStream, c->evChunk[q], 0));
      // this chunk's slice of every selected field: un-permute into the staging area and copy, while the next chunk is computed
      const size_t olo = lo, ohi = (q == Q - 1) ? c->n : (size_t)hi;          // the last chunk carries the ghost entries (zeros)
      size_t off = 0;
      for (int s = 0; s < DV_COUNT; ++s) {
        if (!(mask & (1u << s))) continue;
        const int w = sphb200_deriv_width(c->ndim, s);
        size_t flo = olo;
        if (s == DV_DXDT && dxdtEarly) flo = std::max(olo, std::min(ohi, c->nInt));     // internal entries already left
        const size_t cnt = (ohi - flo)*(size_t)w;
        if (cnt) {
          k_unpermute_range<<<(unsigned)((cnt + RB - 1)/RB), RB, 0, c->copyStream>>>(c->deriv[s], c->cap, c->invPerm, flo, ohi, w, c->stage + off);
          KERNEL_CHECK(c, "k_unpermute_range");
          CU_CHECK(c, cudaMemcpyAsync(deriv_ptr(d, s) + flo*(size_t)w, c->stage + off + flo*(size_t)w, cnt*sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
        }
        off += c->n*(size_t)w;
      }
    }
    CU_CHECK(c, cudaStreamSynchronize(c->copyStream));
    return 0;
  }
  if (sphb200_evaluate_derivatives(c, time, dt)) return 1;
  return sphb200_download_derivs(c, mask, d);
}

int sphb200_upload_derivs(sphb200_ctx* c, unsigned mask, const sphb200_host_derivs* d) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  if (!d) return sphb200_fail(c, "upload_derivs: null source struct");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->n == 0) return 0;
  size_t total = 0;
  for (int s = 0; s < DV_COUNT; ++s) if (mask & (1u << s)) {
    if (!deriv_ptr(d, s)) return sphb200_fail(c, "upload_derivs: field selected in mask but pointer is null");
    total += c->n*(size_t)sphb200_deriv_width(c->ndim, s);
  }
  if (ensure_stage(c, total*sizeof(double))) return 1;
  // the derivative arrays are (re)interpreted in original node order: every field not uploaded is zeroed so that no value of a
  // previous, differently ordered evaluation survives
  for (int s = 0; s < DV_COUNT; ++s)
    CU_CHECK(c, cudaMemsetAsync(c->deriv[s], 0, c->cap*(size_t)sphb200_deriv_width(c->ndim, s)*sizeof(double), c->stream));
  size_t off = 0;
  const unsigned nb = (unsigned)((c->n + RB - 1)/RB);
  for (int s = 0; s < DV_COUNT; ++s) {
    if (!(mask & (1u << s))) continue;
    const int w = sphb200_deriv_width(c->ndim, s);
    CU_CHECK(c, cudaMemcpyAsync(c->stage + off, deriv_ptr(d, s), c->n*(size_t)w*sizeof(double), cudaMemcpyHostToDevice, c->stream));
    k_derivs_in<<<nb, RB, 0, c->stream>>>(c->stage + off, c->cap, c->n, w, c->deriv[s]);
    KERNEL_CHECK(c, "k_derivs_in");
    off += c->n*(size_t)w;
  }
  if (sphb200_ensure(c, c->permEval, c->permEvalCap, c->cap)) return 1;
  k_iota<<<nb, RB, 0, c->stream>>>(c->permEval, c->n);
  KERNEL_CHECK(c, "k_iota");
  CU_CHECK(c, cudaStreamSynchronize(c->stream));                 // the host buffers may be pageable
  c->nEval = c->n; c->capEval = c->cap; c->nIntEval = c->nInt; c->derivNodeValid = true;
  c->derivsValid = false;                                        // the pair-wise data (pair accelerations) are not restored
  return 0;
}

int sphb200_copy_DvDx_to_Q(sphb200_ctx* c) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  // node-wise derivatives stay usable after a later build_pairs (permEval keeps the order they are stored in)
  if (!c->derivNodeValid) return sphb200_fail(c, "copy_DvDx_to_Q: derivatives have not been evaluated");
  if (c->nIntEval != c->nInt) return sphb200_fail(c, "copy_DvDx_to_Q: the internal node count changed since the derivatives were evaluated");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (c->n == 0) return 0;
  // same ghost set as at the evaluation: every node is written (ghost derivatives are zeros); otherwise the internal nodes only,
  // the ghost entries being the boundary conditions' / halo exchange's to fill (ArtificialViscosityHandle.cc:165-180 + boundaries)
  const size_t limit = (c->nEval == c->n) ? c->n : c->nInt;
  if (!c->have[S_DVDXQ]) CU_CHECK(c, cudaMemsetAsync(c->api[S_DVDXQ], 0, c->n*(size_t)(c->ndim*c->ndim)*sizeof(double), c->stream));
  k_copy_dvdx<<<(unsigned)((c->nEval + RB - 1)/RB), RB, 0, c->stream>>>(c->deriv[DV_DVDX], c->capEval, c->permEval, c->nEval, limit, c->ndim*c->ndim, c->api[S_DVDXQ]);
  KERNEL_CHECK(c, "k_copy_dvdx");
  c->have[S_DVDXQ] = true;
  c->rowsValid = false;
  return 0;
}

int sphb200_update_energy_compatible(sphb200_ctx* c, double multiplier) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->opt.compatibleEnergy) return sphb200_fail(c, "update_energy_compatible: compatibleEnergyEvolution is off");
  if (!c->derivsValid || !c->pairsValid) return sphb200_fail(c, "update_energy_compatible: needs the derivatives and pair accelerations of the current pair list");
  for (int s : {S_VEL, S_MASS, S_EPS}) if (!c->have[s]) return sphb200_fail(c, "update_energy_compatible: velocity, mass and specific thermal energy must be on the device");
  if (c->n == 0) return 0;
  cudaEventRecord(c->ev[6], c->stream);
  if (sphb200_launch_energy(c, multiplier)) return 1;
  cudaEventRecord(c->ev[7], c->stream);
  c->energyTimed = true;
  return 0;
}

size_t sphb200_halo_bytes_per_node(const sphb200_ctx* c, unsigned mask) {
  if (!c) return 0;
  size_t w = 0;
  for (int s = 0; s < S_COUNT; ++s) if (mask & (1u << s)) w += (size_t)sphb200_state_width(c->ndim, s);
  return w*sizeof(double);
}

int sphb200_halo_pack(sphb200_ctx* c, unsigned mask, const uint32_t* nodes, size_t count, void* staging) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  HaloFields f{};
  for (int s = 0; s < S_COUNT; ++s) {
    if (!(mask & (1u << s))) continue;
    if (!c->have[s]) return sphb200_fail(c, "halo_pack: field not on device");
    f.src[f.nf] = c->api[s]; f.width[f.nf] = sphb200_state_width(c->ndim, s); f.off[f.nf] = f.total;
    f.total += count*(unsigned long long)f.width[f.nf];
    ++f.nf;
  }
  if (f.total) {
    k_halo_pack_all<<<(unsigned)((f.total + RB - 1)/RB), RB, 0, c->stream>>>(f, nodes, count, (double*)staging);
    KERNEL_CHECK(c, "k_halo_pack_all");
  }
  return 0;
}

static int halo_unpack_impl(sphb200_ctx* c, unsigned mask, size_t firstGhost, size_t count, const void* staging, bool keepConnectivity) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (firstGhost + count > c->n) return sphb200_fail(c, "halo_unpack: ghost range exceeds node count");
  // Late fields of a two-phase exchange (everything but positions and H, landing after the neighbour build): the sorted rows are
  // current except for these ghosts, so only their rows are packed again instead of all of them
  const bool patchRows = c->sortValid && c->rowsValid && !(mask & ((1u << S_POS) | (1u << S_H)));
  HaloFields f{};
  for (int s = 0; s < S_COUNT; ++s) {
    if (!(mask & (1u << s))) continue;
    if (!c->api[s] && c->n) return sphb200_fail(c, "halo_unpack: volume / RK corrections exist only in CRKSPH contexts");
    // ghosts are a contiguous tail of the host-ordered arrays: field block a lands at api[s] + firstGhost*width
    f.dst[f.nf] = c->api[s]; f.width[f.nf] = sphb200_state_width(c->ndim, s); f.off[f.nf] = f.total;
    f.total += count*(unsigned long long)f.width[f.nf];
    ++f.nf;
    c->have[s] = true;
    if (!keepConnectivity && (s == S_POS || s == S_H)) { c->sortValid = false; c->pairsValid = false; }
    if (s != S_EPS && s != S_VOLUME && s != S_RKCORR) c->rowsValid = false;
  }
  if (f.total) {
    k_halo_unpack_all<<<(unsigned)((f.total + RB - 1)/RB), RB, 0, c->stream>>>(f, firstGhost, (const double*)staging);
    KERNEL_CHECK(c, "k_halo_unpack_all");
  }
  if (patchRows && !c->rowsValid) {
    if (sphb200_pack_rows_range(c, firstGhost, count)) return 1;
    c->rowsValid = true;
  }
  return 0;
}
int sphb200_halo_unpack(sphb200_ctx* c, unsigned mask, size_t firstGhost, size_t count, const void* staging) {
  return halo_unpack_impl(c, mask, firstGhost, count, staging, false);
}
int sphb200_halo_unpack_values(sphb200_ctx* c, unsigned mask, size_t firstGhost, size_t count, const void* staging) {
  return halo_unpack_impl(c, mask, firstGhost, count, staging, true);
}

int sphb200_inverse_perm(sphb200_ctx* c) {
  if (sphb200_ensure(c, c->invPerm, c->invPermCap, c->cap)) return 1;
  if (c->n == 0) return 0;
  k_inverse_perm<<<(unsigned)((c->n + RB - 1)/RB), RB, 0, c->stream>>>(c->perm, c->n, c->invPerm);
  KERNEL_CHECK(c, "k_inverse_perm");
  return 0;
}

int sphb200_halo_pack_derivs(sphb200_ctx* c, const uint32_t* nodes, size_t count, void* staging) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->derivsValid || !c->pairsValid) return sphb200_fail(c, "halo_pack_derivs: derivatives have not been evaluated on the current connectivity");
  if (count == 0) return 0;
  if (!nodes || !staging) return sphb200_fail(c, "halo_pack_derivs: null argument");
  if (sphb200_inverse_perm(c)) return 1;
  k_halo_pack_derivs<<<(unsigned)((count + RB - 1)/RB), RB, 0, c->stream>>>(c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], c->cap, c->ndim, c->invPerm,
                                                                           nodes, count, (double*)staging);
  KERNEL_CHECK(c, "k_halo_pack_derivs");
  return 0;
}

int sphb200_halo_unpack_derivs(sphb200_ctx* c, size_t firstGhost, size_t count, const void* staging) {
  if (!c) return sphb200_fail(nullptr, "null ctx");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (!c->derivsValid || !c->pairsValid) return sphb200_fail(c, "halo_unpack_derivs: derivatives have not been evaluated on the current connectivity");
  if (firstGhost + count > c->n) return sphb200_fail(c, "halo_unpack_derivs: ghost range exceeds node count");
  if (count == 0) return 0;
  if (!staging) return sphb200_fail(c, "halo_unpack_derivs: null argument");
  if (sphb200_inverse_perm(c)) return 1;
  k_halo_unpack_derivs<<<(unsigned)((count + RB - 1)/RB), RB, 0, c->stream>>>(c->deriv[DV_DVDT], c->deriv[DV_DEPSDT], c->cap, c->ndim, c->invPerm,
                                                                             firstGhost, count, (const double*)staging);
  KERNEL_CHECK(c, "k_halo_unpack_derivs");
  return 0;
}

int sphb200_node_bounds(sphb200_ctx* c, size_t count, double lo[3], double hi[3], double maxExtent[3]) {
  if (!c || !lo || !hi || !maxExtent) return sphb200_fail(c, "node_bounds: null argument");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (count > c->n) return sphb200_fail(c, "node_bounds: count exceeds node count");
  if (!c->have[S_POS] || !c->have[S_H] || !c->W.set) return sphb200_fail(c, "node_bounds: position, H and the kernel table must be set first");
  for (int a = 0; a < 3; ++a) { lo[a] = 0.0; hi[a] = 0.0; maxExtent[a] = 0.0; }
  if (count == 0) return 0;
  if (sphb200_bounds_reduce(c, count)) return 1;
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  for (int a = 0; a < c->ndim; ++a) { lo[a] = c->reduceHost[a]; hi[a] = c->reduceHost[3 + a]; maxExtent[a] = c->reduceHost[6 + a]; }
  return 0;
}

int sphb200_halo_select(sphb200_ctx* c, int axis, size_t count, double lo, double hi, double width,
                        uint32_t* sendLow, size_t* nLow, uint32_t* sendHigh, size_t* nHigh, size_t cap) {
  if (!c || !nLow || !nHigh) return sphb200_fail(c, "halo_select: null argument");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (axis < 0 || axis >= c->ndim) return sphb200_fail(c, "halo_select: bad axis");
  if (count > c->n) return sphb200_fail(c, "halo_select: count exceeds node count");
  if (!c->have[S_POS]) return sphb200_fail(c, "halo_select: positions are not on the device");
  *nLow = *nHigh = 0;
  if (count == 0) return 0;
  if (ensure_stage(c, 2*(count + 1)*sizeof(uint32_t))) return 1;
  uint32_t* fLow = (uint32_t*)c->stage; uint32_t* fHigh = fLow + (count + 1);
  const unsigned nb = (unsigned)((count + RB - 1)/RB);
  k_halo_flags<<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->ndim, axis, count, lo + width, hi - width, fLow, fHigh);
  KERNEL_CHECK(c, "k_halo_flags");
  if (sphb200_scan_u32(c, fLow, fLow, count)) return 1;
  if (sphb200_scan_u32(c, fHigh, fHigh, count)) return 1;
  k_halo_scatter<<<nb, RB, 0, c->stream>>>(fLow, fHigh, count, cap, sendLow, sendHigh);
  KERNEL_CHECK(c, "k_halo_scatter");
  uint32_t tot[2];
  CU_CHECK(c, cudaMemcpyAsync(&tot[0], fLow + count, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaMemcpyAsync(&tot[1], fHigh + count, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  *nLow = tot[0]; *nHigh = tot[1];
  if (tot[0] > cap || tot[1] > cap) return sphb200_fail(c, "halo_select: send list capacity too small");
  return 0;
}

int sphb200_node_bounds_device(sphb200_ctx* c, size_t count, double* outDevice) {
  if (!c || !outDevice) return sphb200_fail(c, "node_bounds_device: null argument");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (count > c->n || count == 0) return sphb200_fail(c, "node_bounds_device: count must be in [1, node count]");
  if (!c->have[S_POS] || !c->have[S_H] || !c->W.set) return sphb200_fail(c, "node_bounds_device: position, H and the kernel table must be set first");
  if (sphb200_bounds_reduce(c, count)) return 1;
  CU_CHECK(c, cudaMemcpyAsync(outDevice, c->reduceBuf + 296*12, 9*sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

int sphb200_halo_select_device(sphb200_ctx* c, int axis, size_t count, double lo, double hi, const double* maxExtentDevice,
                               uint32_t* sendLow, uint32_t* sendHigh, long long* countsDevice, size_t cap) {
  if (!c || !maxExtentDevice || !countsDevice) return sphb200_fail(c, "halo_select_device: null argument");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  if (axis < 0 || axis >= c->ndim) return sphb200_fail(c, "halo_select_device: bad axis");
  if (count > c->n || count == 0) return sphb200_fail(c, "halo_select_device: count must be in [1, node count]");
  if (!c->have[S_POS]) return sphb200_fail(c, "halo_select_device: positions are not on the device");
  if (ensure_stage(c, 2*(count + 1)*sizeof(uint32_t))) return 1;
  uint32_t* fLow = (uint32_t*)c->stage; uint32_t* fHigh = fLow + (count + 1);
  const unsigned nb = (unsigned)((count + RB - 1)/RB);
  k_halo_flags_dev<<<nb, RB, 0, c->stream>>>(c->api[S_POS], c->ndim, axis, count, lo, hi, maxExtentDevice, fLow, fHigh);
  KERNEL_CHECK(c, "k_halo_flags_dev");
  if (sphb200_scan_u32(c, fLow, fLow, count)) return 1;
  if (sphb200_scan_u32(c, fHigh, fHigh, count)) return 1;
  k_halo_scatter<<<nb, RB, 0, c->stream>>>(fLow, fHigh, count, cap, sendLow, sendHigh);
  KERNEL_CHECK(c, "k_halo_scatter");
  k_halo_counts<<<1, 32, 0, c->stream>>>(fLow, fHigh, count, cap, countsDevice);
  KERNEL_CHECK(c, "k_halo_counts");
  return 0;
}

int sphb200_get_stats(sphb200_ctx* c, sphb200_stats* out) {
  if (!c || !out) return sphb200_fail(c, "get_stats: null argument");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  CU_CHECK(c, cudaStreamSynchronize(c->stream));
  if (c->derivsValid && c->n) {
    if (cudaEventElapsedTime(&c->stats.ms_evaluate, c->ev[3], c->ev[5]) != cudaSuccess) c->stats.ms_evaluate = 0;
    if (cudaEventElapsedTime(&c->stats.ms_pair_kernel, c->ev[4], c->ev[5]) != cudaSuccess) c->stats.ms_pair_kernel = 0;
  }
  if (!c->energyTimed || cudaEventElapsedTime(&c->stats.ms_energy, c->ev[6], c->ev[7]) != cudaSuccess) c->stats.ms_energy = 0;   // events never recorded: no API error
  cudaGetLastError();
  c->stats.stencil_radius = (uint32_t)c->stencilR;
  c->stats.fine_walk = c->fineWalk ? 1u : 0u;
  *out = c->stats;
  return 0;
}

int sphb200_measure_fp64_peak(sphb200_ctx* c, double* tflops) {
  if (!c || !tflops) return sphb200_fail(c, "measure_fp64_peak: null argument");
  CU_CHECK(c, cudaSetDevice(c->device)); if (sphb200_join_uploads(c, true)) return 1;
  cudaDeviceProp prop;
  CU_CHECK(c, cudaGetDeviceProperties(&prop, c->device));
  const int blocks = prop.multiProcessorCount*8, threads = 256, iters = 20000;
  if (ensure_stage(c, (size_t)blocks*threads*sizeof(double))) return 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, c->stream);
    k_fp64_peak<<<blocks, threads, 0, c->stream>>>(c->stage, iters, 0.999999, 1.0e-7);
    cudaEventRecord(e1, c->stream);
    cudaStreamSynchronize(c->stream);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
    c->stats.launches++;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  CU_CHECK(c, cudaGetLastError());
  const double flops = 2.0*8.0*(double)iters*(double)blocks*(double)threads;
  *tflops = flops/(best*1e-3)/1e12;
  return 0;
}

}  // extern "C"
