// sphb200_internal.cuh -- context layout and device helpers shared by the kernels of libsphb200.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   api_*      : fields exactly as the host hands them (reference AoS, original node order) -- the landing
//                zone of uploads / halo receives and the source of downloads.
//   rows       : nodes re-ordered by Morton cell key, one 128-byte (3-D) / 96-byte (2-D) record per node
//                   3-D: x y z vx vy vz Hxx Hxy Hxz Hyy Hyz Hzz m rho Prho cs
//                   2-D: x y vx vy Hxx Hxy Hyy m rho Prho cs (pad)
//                Prho = safeInv(omega)*P/(rho*rho) (SPH.cc:310,425 with epsTensile == 0) is folded per node.
//   nbr        : neighbour lists of internal nodes, sliced-ELL: tile t = sorted slots [32t,32t+32),
//                entry k of lane l at nbr[tileOff[t] + 32*k + l] = sorted slot of the neighbour.
//   d_*        : derivative fields, SoA per component, sorted order.
//   pacc       : the pair force of every directed edge (deltaDvDt of SPH.cc:427 in i's orientation), blocked like nbr: word c of edge
//                `slot` at pacc_at(W, slot, c) = W*(slot & ~31) + 32*c + (slot & 31), one 32-lane line per (list row, word).  W words per
//                edge, by storage mode (ctx::paccMode): FULL the DIM components; ISO one scalar sd with delta = sd r_ij (every H a
//                multiple of the identity); TENSOR two scalars (a, b) with delta = a H_i.(H_i.r_ij) + b H_j.(H_j.r_ij).  The compressed
//                modes are expanded by their readers (k_energy, k_emit_pacc) from the node rows of the evaluation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "sphb200.h"

#define SPHB200_TILE 32
#define SPHB200_MAX_PLANES 6
#define SPHB200_MAX_STENCIL 4      /* largest stencil radius in cells (9^3 cells; run codes of k_nbr_build hold 2048 runs per tile) */
#define SPHB200_DIL 32768          /* entries per axis of the dilated-coordinate table (max cells per axis) */

enum StateSlot { S_POS = 0, S_VEL, S_H, S_MASS, S_RHO, S_EPS, S_P, S_CS, S_OMEGA, S_DVDXQ, S_FCL, S_FCQ, S_VOLUME, S_RKCORR, S_COUNT };
enum DerivSlot { DV_DXDT = 0, DV_DRHODT, DV_DVDT, DV_DEPSDT, DV_DVDX, DV_LOCALDVDX, DV_GRADRHO, DV_M, DV_LOCALM,
                 DV_RHOSUM, DV_NORM, DV_MAXQ, DV_EFFQ, DV_XSPHW, DV_XSPHDV, DV_DHDT, DV_HIDEAL, DV_M0, DV_M1, DV_COUNT };

// width (doubles per node) of each field
__host__ __device__ inline int sphb200_state_width(int ndim, int slot) {
  switch (slot) {
    case S_POS: case S_VEL: return ndim;
    case S_H: return ndim == 3 ? 6 : 3;
    case S_DVDXQ: return ndim*ndim;
    case S_RKCORR: return (ndim + 1)*(ndim + 1);
    default: return 1;
  }
}
__host__ __device__ inline int sphb200_deriv_width(int ndim, int slot) {
  switch (slot) {
    case DV_DXDT: case DV_DVDT: case DV_GRADRHO: case DV_XSPHDV: case DV_M1: return ndim;
    case DV_DVDX: case DV_LOCALDVDX: case DV_M: case DV_LOCALM: return ndim*ndim;
    case DV_DHDT: case DV_HIDEAL: return ndim == 3 ? 6 : 3;
    default: return 1;
  }
}

struct TableDev {
  double kext = 0, xmin = 0, xstep = 0;
  uint32_t n1 = 0;
  double* coef = nullptr;          // interleaved per interval: W a0 a1 a2, gradW g0 g1 g2  (6*(n1+1) doubles)
  bool set = false;
  // CubicHermite nperh lookup (Wsum -> nperh)
  uint32_t nperhN = 0; double nperhXmin = 0, nperhXmax = 0, nperhXstep = 0; double* nperhVals = nullptr;
  std::vector<double> hostW, hostG;   // host copies (W0, WnPerh evaluation, equality test)
};

struct GridDev {                    // uniform cell grid, Morton keyed
  double lo[3], cs[3], invcs_unused[3];
  int nc[3];
  int bits[3];
  uint32_t mask[3];                 // dilated bit masks per axis
  uint8_t bitpos[3][16];            // output bit position of bit l of axis a
  uint32_t tableSize;               // 2^(sum bits)
};

struct sphb200_ctx {
  int device = 0;
  int ndim = 3;
  sphb200_options opt{};
  cudaStream_t stream = nullptr;
  // host->device uploads run on their own stream so that the neighbour build (which needs positions and H only) overlaps the
  // transfer of the other state fields; every entry point joins the pending uploads it depends on (sphb200_join_uploads)
  cudaStream_t copyStream = nullptr;
  cudaEvent_t evGeomUp = nullptr, evRestUp = nullptr, evMainMark = nullptr;
  bool pendGeomUp = false, pendRestUp = false;
  bool ghostRefillPending = false;  // plane ghosts exist whose non-geometric values predate an upload still in flight (sphb200_join_uploads refills them)
  std::string err;

  size_t nInt = 0, nGhost = 0, n = 0, cap = 0;
  double* api[S_COUNT] = {nullptr};
  bool have[S_COUNT] = {false};

  TableDev W, WQ;
  bool oneKernel = true;

  // grid + sort
  GridDev grid{};
  GridDev gridFine{};               // two-level walk (experimental, SPHB200_FINE_WALK=1): the sort grid, cells of half the width; grid stays the coarse one
  bool nbrV2 = false;               // the current rows were packed for k_nbr_build2 (tile-frame coordinates: wider FP32 error bands)
  bool fineWalk = false;            // the current sort used gridFine (keys, cellStart, dilTab refer to it; coarse key = fine key >> ndim)
  uint32_t* dilTab = nullptr;       // 3*SPHB200_DIL dilated cell coordinates
  uint32_t* cellKeyApi = nullptr;   // key per node, original order
  uint32_t ghostBase = 0;           // ghost split of the current sort: offset of the ghost copy of the cell table (0: ghosts sorted among the internal nodes)
  uint32_t* cellStart = nullptr;    // tableSize+1
  uint32_t* cellCursor = nullptr;
  size_t cellCap = 0, cellCursorCap = 0;
  uint32_t* perm = nullptr;         // sorted slot -> original index
  uint32_t* skey = nullptr;         // sorted slot -> cell key
  double* reduceBuf = nullptr;      // bbox partials
  double* reduceHost = nullptr;     // pinned

  // sorted rows + aux
  double* rows = nullptr;
  float* frows = nullptr; size_t frowsCap = 0;   // FP32 pre-filter rows of K2 (relpos, H, band thresholds)
  double* aux2 = nullptr;           // {det H, 1/rho} per node (sorted): the per-node part of the pair arithmetic
  double* auxPneg = nullptr;        // max(-P,0)                      (tensile, SPH.cc:417)
  double* auxSomr2 = nullptr;       // safeInv(omega)/(rho*rho)        (tensile)
  double* auxDvDxQ = nullptr;       // ndim*ndim per node (sorted)     (LimitedMG / Balsara)
  double* auxfCl = nullptr; double* auxfCq = nullptr;
  double* crkVolS = nullptr;        // CRKSPH: volume per node (sorted)
  double* crkCorrS = nullptr;       // CRKSPH: (1+ndim)^2 RK coefficients per node (sorted, AoS)
  double* crkQS = nullptr;          // CRKSPH: {volume, Q velocity gradient} per node (sorted, stride 10 / 6)
  double* crkAux = nullptr;         // CRKSPH: {det H, volume} per node (sorted)
  int stencilR = 1;                 // stencil radius (cells) covering the largest kernel extent; 1 unless the extents have a heavy tail
  int coarseHold = 0;               // builds left before the fine grid is tried again (0 with forceR1: never, run-code overflow)
  bool forceR1 = false;             // fall back to cells as wide as the largest extent (set when a tile overflows its run codes)
  uint32_t* cellReach = nullptr; size_t cellReachCap = 0;   // per cell: stencil radius tiles of that cell must walk
  uint32_t* tileRadius = nullptr; size_t tileRadiusCap = 0;
  bool allIsotropic = false;        // every H packed at the last build_pairs was a multiple of the identity (k_pack; read back with the counters)
  bool isoHint = false;             // allIsotropic of the last completed build: tuning hint for the next one (never a correctness input)
  bool rowsValid = false;           // rows reflect current api state for the current sort
  bool sortValid = false;

  // neighbour lists
  size_t nTiles = 0;
  uint32_t* nbrCount = nullptr;     // per sorted slot
  uint32_t* tileRows = nullptr;     // per tile: max count
  unsigned long long* tileOff = nullptr;  // nTiles+1, in entries
  uint32_t* nbr = nullptr; size_t nbrCap = 0;
  uint4* runs = nullptr; size_t runsCap = 0;   // candidate runs of all tiles: {first sorted slot, length, sx|sy<<16, sz}
  uint32_t* tileRunStart = nullptr; // per tile: first run
  uint32_t* tileRunCount = nullptr; // per tile: number of runs
  int listRows = 0;                 // rows of the shared-memory list staging of k_nbr_build (adapts to the longest list)
  int nbrChunks = 0;                // k_nbr_build2: chunks of 32 surviving candidates a tile may record (adapts to the busiest tile)
  unsigned long long* counters = nullptr; // see NbrArgs::counters (8 entries)
  unsigned long long* countersHost = nullptr; // pinned
  size_t npairs = 0, nEdges = 0, nSlots = 0;
  bool pairsValid = false;
  void* scanTmp = nullptr; size_t scanTmpBytes = 0;

  // derivatives (sorted, SoA per component): deriv[slot] points at width*cap doubles, component c at + c*cap
  double* deriv[DV_COUNT] = {nullptr};
  double* pacc = nullptr; size_t paccCap = 0;   // paccWidth * nSlots
  int paccMode = 0, paccWidth = 0;              // PACC_FULL | PACC_ISO | PACC_TENSOR of the last evaluation, words per directed edge
  bool rowsAtEval = false;                      // `rows` still hold what the last evaluation read (the compressed modes expand against them)
  bool derivsValid = false;
  // node-wise derivatives outlive the connectivity they were computed on (CheapSynchronousRK2 advances the next step's trial
  // state with them after the neighbour update): the sorted order of the evaluation is kept beside them
  uint32_t* permEval = nullptr; size_t permEvalCap = 0;
  size_t nEval = 0, capEval = 0, nIntEval = 0;
  bool derivNodeValid = false;
  // state0 of the integrator (State::copyState, CheapSynchronousRK2.cc:70-71)
  double* api0[S_COUNT] = {nullptr}; size_t cap0[S_COUNT] = {0}; bool have0[S_COUNT] = {false}; size_t n0 = 0;
  // reflecting planes (boundary.cu): {point[3], unit normal[3]} per plane, the ghost range and the control list of each
  int nPlanes = 0; double planes[12*SPHB200_MAX_PLANES] = {0}; int planeKind[SPHB200_MAX_PLANES] = {0};   // enter {p, n}, exit {p, n}; 1 = periodic
  size_t planeFirst[SPHB200_MAX_PLANES] = {0}, planeCount[SPHB200_MAX_PLANES] = {0};
  uint32_t* planeCtl[SPHB200_MAX_PLANES] = {nullptr}; size_t planeCtlCap[SPHB200_MAX_PLANES] = {0};
  uint32_t* invPerm = nullptr; size_t invPermCap = 0;      // original index -> sorted slot (built on demand)
  uint32_t* hDone = nullptr; size_t hDoneCap = 0;           // iterateIdealH: nodes whose H has converged
  // time-step reduction scratch
  unsigned long long* dtCand = nullptr; size_t dtCandCap = 0;
  double* dtAux = nullptr; size_t dtAuxCap = 0;
  double* stage = nullptr; size_t stageBytes = 0;   // download staging (device)
  // chunked evaluation with pipelined download (sphb200_evaluate_derivatives_to_host): per chunk of the host index range the tiles that
  // hold its nodes; valid for the current sort
  uint32_t* chunkList = nullptr; size_t chunkListCap = 0;     // SPHB200_MAX_CHUNKS lists of nTiles entries
  uint32_t* chunkCount = nullptr; uint32_t* chunkCountHost = nullptr;   // device / pinned
  bool chunkListsValid = false; int chunkQ = 0;
  cudaEvent_t evChunk[8] = {nullptr};

  // instrumentation
  sphb200_stats stats{};
  cudaEvent_t ev[8] = {nullptr};
  bool energyTimed = false;          // ev[6], ev[7] have been recorded (sphb200_update_energy_compatible ran)
};

int  sphb200_fail(sphb200_ctx* c, const std::string& msg);
constexpr int SPHB200_MAX_CHUNKS = 8;
#define CU_CHECK(c, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) \
  return sphb200_fail((c), std::string(#call) + ": " + cudaGetErrorString(e__)); } while (0)
#define KERNEL_CHECK(c, name) do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) \
  return sphb200_fail((c), std::string("launch ") + name + ": " + cudaGetErrorString(e__)); (c)->stats.launches++; } while (0)

template <typename T> int sphb200_ensure(sphb200_ctx* c, T*& p, size_t& cap, size_t need);

// implemented in scan.cu: exclusive scans with the total written at out[n]
int sphb200_scan_u32(sphb200_ctx* c, const uint32_t* in, uint32_t* out, size_t n);
int sphb200_scan_tiles(sphb200_ctx* c, const uint32_t* rows, unsigned long long* out, size_t n);   // out = 32*rows prefix

// implemented in neighbors.cu / derivs.cu / energy.cu
int sphb200_join_uploads(sphb200_ctx* c, bool all);   // main stream waits for pending uploads: positions + H only, or everything
int sphb200_sort_and_pack(sphb200_ctx* c);
extern "C" int sphb200_inverse_perm(sphb200_ctx* c);        // (re)builds c->invPerm from the current sort (api.cu)
int sphb200_bounds_reduce(sphb200_ctx* c, size_t count);   // bbox + max extents of nodes [0,count) -> reduceHost[0..8] (async copy)
int sphb200_pack_rows(sphb200_ctx* c);
int sphb200_pack_rows_range(sphb200_ctx* c, size_t first, size_t count);   // rows of the nodes [first, first + count) only (sort unchanged)
int sphb200_neighbors(sphb200_ctx* c);
int sphb200_launch_derivs(sphb200_ctx* c);
int sphb200_launch_derivs_chunk(sphb200_ctx* c, const uint32_t* tileList, uint32_t nList, uint32_t origLo, uint32_t origHi);
int sphb200_launch_energy(sphb200_ctx* c, double multiplier);
int sphb200_launch_asph_classic(sphb200_ctx* c);        // steps.cu: ASPHClassicSmoothingScale ideal H, after the pair loop
int sphb200_launch_crk_derivs(sphb200_ctx* c);
int sphb200_pairs_to_host(sphb200_ctx* c, uint32_t* pi, uint32_t* pj, size_t cap, double* pacc, size_t paccCap);

// ---- device helpers ---------------------------------------------------------------------------------------
template <int DIM> struct Dm;
template <> struct Dm<3> { static constexpr int NS = 6, NT = 9, ROW = 16, R_POS = 0, R_VEL = 3, R_H = 6, R_M = 12, R_RHO = 13, R_PRHO = 14, R_CS = 15; };
template <> struct Dm<2> { static constexpr int NS = 3, NT = 4, ROW = 12, R_POS = 0, R_VEL = 2, R_H = 4, R_M = 7, R_RHO = 8, R_PRHO = 9, R_CS = 10; };

__host__ __device__ __forceinline__ bool sphb200_is_asph(int hEvolution) { return hEvolution == SPHB200_H_ASPH || hEvolution == SPHB200_H_ASPH_CLASSIC; }
enum { PACC_FULL = 0, PACC_ISO = 1, PACC_TENSOR = 2 };
__host__ __device__ __forceinline__ int pacc_width(int mode, int ndim) { return mode == PACC_ISO ? 1 : (mode == PACC_TENSOR ? 2 : ndim); }
__host__ __device__ __forceinline__ unsigned long long pacc_at(int width, unsigned long long slot, int word) {
  return (unsigned long long)width*(slot & ~31ull) + 32ull*(unsigned)word + (slot & 31ull);
}

__device__ __forceinline__ double d_sgn(double x) { return x < 0.0 ? -1.0 : 1.0; }

// Exactly-rounded, never-contracted H.r and |.|^2 for the pair predicate (ConnectivityMap.cc:912-925;
// GeomSymmetricTensorInline.hh:1762-1777, GeomVectorInline.hh:1045-1055): the reference is built for generic
// x86-64, i.e. without FMA, so every product and sum is rounded separately, left to right.
template <int DIM> __device__ __forceinline__ double eta2_exact(const double* H, const double* r) {
  if (DIM == 3) {
    const double ex = __dadd_rn(__dadd_rn(__dmul_rn(H[0], r[0]), __dmul_rn(H[1], r[1])), __dmul_rn(H[2], r[2]));
    const double ey = __dadd_rn(__dadd_rn(__dmul_rn(H[1], r[0]), __dmul_rn(H[3], r[1])), __dmul_rn(H[4], r[2]));
    const double ez = __dadd_rn(__dadd_rn(__dmul_rn(H[2], r[0]), __dmul_rn(H[4], r[1])), __dmul_rn(H[5], r[2]));
    return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
  } else {
    const double ex = __dadd_rn(__dmul_rn(H[0], r[0]), __dmul_rn(H[1], r[1]));
    const double ey = __dadd_rn(__dmul_rn(H[1], r[0]), __dmul_rn(H[2], r[1]));
    return __dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey));
  }
}

// det H by cofactors of the first row: 9 FP64 instructions instead of the 17 of the six-term form (GeomSymmetricTensorInline.hh:
// 1740-1749); the two agree to round-off, which is all the 1e-10 bar asks of a pair-loop factor.
template <int DIM> __device__ __forceinline__ double sym_det(const double* H) {
  if (DIM == 3) return fma(H[0], fma(H[3], H[5], -H[4]*H[4]), fma(H[2], fma(H[1], H[4], -H[2]*H[3]), -H[1]*fma(H[1], H[5], -H[2]*H[4])));
  return H[0]*H[2] - H[1]*H[1];
}
template <int DIM> __device__ __forceinline__ void sym_dot(const double* H, const double* r, double* o) {
  if (DIM == 3) {
    o[0] = H[0]*r[0] + H[1]*r[1] + H[2]*r[2];
    o[1] = H[1]*r[0] + H[3]*r[1] + H[4]*r[2];
    o[2] = H[2]*r[0] + H[4]*r[1] + H[5]*r[2];
  } else {
    o[0] = H[0]*r[0] + H[1]*r[1];
    o[1] = H[1]*r[0] + H[2]*r[1];
  }
}
template <int DIM> __device__ __forceinline__ double vdot(const double* a, const double* b) {
  double s = a[0]*b[0];
#pragma unroll
  for (int k = 1; k < DIM; ++k) s += a[k]*b[k];
  return s;
}

// cell coordinate of a position on the grid (shared by every kernel that needs it, so it is bit-identical)
__device__ __forceinline__ int cell_coord(double x, double lo, double cs, int nc) {
  int k = (int)floor((x - lo)/cs);
  return k < 0 ? 0 : (k >= nc ? nc - 1 : k);
}
__device__ __forceinline__ uint32_t dilate(const GridDev& g, int axis, int c) {
  uint32_t key = 0;
  for (int l = 0; l < g.bits[axis]; ++l) key |= ((uint32_t)(c >> l) & 1u) << g.bitpos[axis][l];
  return key;
}
