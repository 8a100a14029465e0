// derivs.cu -- K3 (SPH pair loop), K4 (per-node finalize) and K5 (smoothing-scale derivatives), fused.
//
// Replaces SPH<Dim>::evaluateDerivativesImpl (SPH/SPH.cc:165-555) followed by
// SPHSmoothingScale::evaluateDerivatives (SmoothingScale/SPHSmoothingScale.cc:101-275) or
// ASPHSmoothingScale::evaluateDerivatives (SmoothingScale/ASPHSmoothingScale.cc:110-147).
//
// Formulation: i-centric gather.  The reference walks each pair (i<j) once and scatters into both nodes through
// per-thread full-size scratch copies (SPH.cc:271-292, 473).  Here every internal node i walks its complete neighbour
// list and accumulates only its own sums in registers, so there is no scatter, no atomic and the summation order of
// a node is fixed by the list order (bitwise reproducible run to run).  Each pair is therefore evaluated twice; the
// terms are the reference's (Appendix A of SURVEY.md) with i and j exchanged for the second visit.
#include "sphb200_internal.cuh"
#include "pair_common.cuh"
#include <algorithm>
#include <cmath>

namespace {

struct DerivArgs {
  const double* rows; const double* aux2; const uint32_t* perm;
  const uint32_t* nbrCount; const uint32_t* tileRows; const unsigned long long* tileOff; const uint32_t* nbr;
  const double *auxPneg, *auxSomr2, *auxDvDxQ, *auxfCl, *auxfCq;
  const double* tabW; const double* tabQ;       // interleaved coefficient tables (6 per interval)
  double kextW, xminW, xstepW; uint32_t n1W;
  double kextQ, xminQ, xstepQ; uint32_t n1Q;
  const double* nperhVals; uint32_t nperhN; double nperhXmin, nperhXmax, nperhXstep;
  double W0, WnPerh;
  size_t n, cap; uint32_t nInt;
  sphb200_options o;
  int oneKernel;
  double* deriv[DV_COUNT];
  double* pacc; size_t nSlots;
  // chunked evaluation (sphb200_evaluate_derivatives_to_host): this launch visits the tiles of `tileList` only and evaluates the
  // internal nodes whose ORIGINAL index lies in [origLo, origHi); nodes of other chunks are left untouched.  tileList == nullptr:
  // every tile, every internal node.
  const uint32_t* tileList; uint32_t nList; uint32_t origLo, origHi;
};

// Asynchronous global->shared copies (LDGSTS): a warp streams the 128-byte rows of its lanes' upcoming neighbours into a
// shared-memory ring, PAIR_STAGES iterations ahead of their use, so the L2 latency of the gather is off the critical path
// although only 2 warps per scheduler are resident (the accumulators cost ~200 registers per thread).
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

__device__ __forceinline__ void cp_async16_s(unsigned smemDst, const void* gmemSrc) {
#if defined(SPHB200_CP_CG) && SPHB200_CP_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smemDst), "l"(gmemSrc) : "memory");     // L2 only
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smemDst), "l"(gmemSrc) : "memory");
#endif
}

#ifndef SPHB200_COPY_LANES
#define SPHB200_COPY_LANES 8
#endif
#ifndef SPHB200_PAIR_STAGES
#define SPHB200_PAIR_STAGES 4
#endif
// One CTA of 8 warps per SM (register-limited to 8 warps either way): the TableKernel table is staged once per SM instead of twice,
// which leaves 24 KB more L1 for the neighbour-row gather (noh8m pair kernel 16.75 -> 16.44 ms; profiles/r02_notes.md).
#ifndef SPHB200_PAIR_WARPS
#define SPHB200_PAIR_WARPS 8
#endif
#ifndef SPHB200_PAIR_CTAS
#define SPHB200_PAIR_CTAS 1
#endif
// 1: stream the per-node {det H, 1/rho} record of every neighbour through the ring as well; 0: recompute both from the row.
// The per-lane 16-byte copy costs 32 shared-memory wavefronts per iteration, as many as the 32 rows together, and the LSU data
// pipe is the unit this kernel saturates first (profiles/r01_notes.md); 23 extra FP64 instructions per edge are cheaper.
#ifndef SPHB200_PAIR_CTAS_ISO
#define SPHB200_PAIR_CTAS_ISO 1
#endif
#ifndef SPHB200_PAIR_AUX
#define SPHB200_PAIR_AUX 0
#endif
// 1: on the isotropic path every pair term is a scalar times r_ij (both kernel gradients are parallel to r_ij), so the viscosity
// projections, the pair force, the work terms and the moments are formed from v_ij.r_ij and scalar factors instead of component
// by component: ~20 FP64 instructions less per directed edge.  Same terms, re-associated (parity 1e-10).
#ifndef SPHB200_ISO_SCALAR
#define SPHB200_ISO_SCALAR 1
#endif
__device__ __forceinline__ const unsigned char* mad_wide(uint32_t a, uint32_t b, const unsigned char* c) {   // c + a*b in one IMAD.WIDE
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long)c));
  return reinterpret_cast<const unsigned char*>(r);
}

#ifndef SPHB200_TILE_SYNC
#define SPHB200_TILE_SYNC 5     // 0: none (formally a race), 1 / 4: warp barrier after the tile's loop, 2 / 3: before the next tile's prologue, 5: rotating ring base (no barrier)
#endif
#ifndef SPHB200_PAIR_UNROLL
#define SPHB200_PAIR_UNROLL 1
#endif
constexpr int PAIR_UNROLL = SPHB200_PAIR_UNROLL;   // unroll factor of the neighbour loop (1: measured best, see profiles/r02_notes.md)
constexpr int PAIR_STAGES = SPHB200_PAIR_STAGES;   // depth of the neighbour-row ring
constexpr int PAIR_WARPS = SPHB200_PAIR_WARPS;     // warps (= tiles in flight) per CTA
constexpr int PAIR_CTAS = SPHB200_PAIR_CTAS;       // resident CTAs per SM the register budget is sized for
template <int DIM> struct RingGeom {
  static constexpr int ROWB = Dm<DIM>::ROW*8 + 16;          // +16 B pad: conflict-free 128-bit LDS
  static constexpr int STAGEB = 32*ROWB + (SPHB200_PAIR_AUX ? 32*16 : 0);   // 32 rows (+ 32 aux records {det H, 1/rho})
};

// The pair-loop kernel: one warp per tile of 32 Morton-consecutive nodes, lane <-> node i; persistent CTAs stride over tiles.
//   GEN    : general path -- every option of the reference is honoured at run time (LimitedMG, Balsara, Cl/Cq multipliers,
//            tensile correction, separate Pi kernel, linear/quadraticInExpansion, any XSPH/compatible/smoothing-scale choice).
//   !GEN   : the plain MonaghanGingold path named by BASELINE.json, with XSPH / SPH-moments / pair-acceleration storage
//            fixed at compile time so that unused accumulators cost no registers.
//   ISO    : every H on the device is a multiple of the identity (checked in k_pack; the normal state of an SPH -- as opposed to
//            ASPH -- run).  Then eta = r/h, both kernel gradients are parallel to r_ij and the tensor products collapse to
//            scalar factors: ~40 % fewer FP64 instructions per edge, 3 accumulators less (M is symmetric).  Same terms as the
//            general path, re-associated (parity 1e-10).
template <int DIM, bool GEN, bool XSPH_, bool HSPH_, bool COMPAT_, bool ISO>
__global__ void __launch_bounds__(32*PAIR_WARPS, ISO ? SPHB200_PAIR_CTAS_ISO : PAIR_CTAS) k_sph_derivs(DerivArgs a) {
  static_assert(!(GEN && ISO), "the isotropic fast path exists for the plain MonaghanGingold configuration only");
  using D = Dm<DIM>;
  constexpr int NS = D::NS, NT = D::NT, ROW = D::ROW;
  constexpr int ROWB = RingGeom<DIM>::ROWB;
  constexpr bool ISC = ISO && (SPHB200_ISO_SCALAR != 0);
  extern __shared__ __align__(16) double smem[];
  // stage the interleaved W/gradW table(s) in shared memory
  // n1+1 interval records of 6 coefficients plus one all-zero record (eta >= kext)
  const uint32_t nW = 6u*(a.n1W + 2u), nQ = (GEN && !a.oneKernel) ? 6u*(a.n1Q + 2u) : 0u;
  for (uint32_t k = threadIdx.x; k < nW; k += blockDim.x) smem[k] = (k + 6u < nW) ? a.tabW[k] : 0.0;
  for (uint32_t k = threadIdx.x; k < nQ; k += blockDim.x) smem[nW + k] = (k + 6u < nQ) ? a.tabQ[k] : 0.0;
  __syncthreads();
  const unsigned tW = (unsigned)__cvta_generic_to_shared(smem);
  const unsigned tQ = tW + 8u*nW;
  const double rxW = 1.0/a.xstepW, rxQ = (GEN && !a.oneKernel) ? 1.0/a.xstepQ : 0.0;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  // ring of this warp: PAIR_STAGES stages x 32 lane slots of ROWB bytes
  constexpr int STAGEB = RingGeom<DIM>::STAGEB;
  const unsigned warpRing = tW + 8u*(nW + nQ) + (unsigned)warp*(PAIR_STAGES*STAGEB);
  const sphb200_options& o = a.o;
  const size_t nTiles = a.tileList ? (size_t)a.nList : (a.n + SPHB200_TILE - 1)/SPHB200_TILE;
  // Ring stage of list position 0 of the current tile.  It rotates from tile to tile so that the prologue of a tile never refills the
  // stage the previous tile read LAST: a lane that leaves a tile early (ghost, node of another chunk, short list) may issue the next
  // prologue while a slower lane is still reading that stage, and the first __syncwarp both have passed orders every earlier read only
  // (compute-sanitizer racecheck reports this tile-to-tile write-after-read).  A warp barrier between tiles cures it too but costs
  // 2 % of the kernel (15.27 -> 15.6-15.9 ms at 8 M, profiles/r02_notes.md); the rotation costs nothing.
  uint32_t stageBase = 0u;
  for (size_t tl = (size_t)blockIdx.x*PAIR_WARPS + warp; tl < nTiles; tl += (size_t)gridDim.x*PAIR_WARPS) {
  const size_t tile = a.tileList ? (size_t)a.tileList[tl] : tl;
  const size_t i = tile*SPHB200_TILE + lane;
  const bool inRange = i < a.n;
  const uint32_t origi = inRange ? a.perm[i] : 0xffffffffu;
  const bool internal = inRange && origi < a.nInt;
  const bool active = internal && origi >= a.origLo && origi < a.origHi;
  const double tiny = 1.0e-30;
  const bool xsph = GEN ? (o.XSPH != 0) : XSPH_;
  const bool hsph = GEN ? (o.hEvolution == SPHB200_H_SPH) : HSPH_;
  const bool compat = GEN ? (o.compatibleEnergy != 0) : COMPAT_;
  const bool tens = GEN && (o.epsTensile != 0.0);
  const bool needQ = GEN && ((o.Qkind == SPHB200_Q_LIMITED_MG) || o.balsara);
  const bool mult = GEN && (a.auxfCl != nullptr);
  const bool twoK = GEN && !a.oneKernel;

  // ---- node i state
  double ri[DIM], vi[DIM], Hi[NS];
  double mi = 0, rhoi = 1, Prhoi0 = 0, ci = 0;
  if (inRange) {
    const double* r = a.rows + i*ROW;
#pragma unroll
    for (int k = 0; k < DIM; ++k) { ri[k] = r[D::R_POS + k]; vi[k] = r[D::R_VEL + k]; }
#pragma unroll
    for (int k = 0; k < NS; ++k) Hi[k] = r[D::R_H + k];
    mi = r[D::R_M]; rhoi = r[D::R_RHO]; Prhoi0 = r[D::R_PRHO]; ci = r[D::R_CS];
  } else {
#pragma unroll
    for (int k = 0; k < DIM; ++k) { ri[k] = 0; vi[k] = 0; }
#pragma unroll
    for (int k = 0; k < NS; ++k) Hi[k] = 0;
  }
  const double Hdeti = sym_det<DIM>(Hi);
  const double hiInv = Hi[0], hi2 = hiInv*hiInv;           // ISO: H_i = hiInv * I
  const double rhoiInv = 1.0/rhoi;
  const double mi_over_rhoi = mi/rhoi;
  const double Pnegi = (tens && inRange) ? a.auxPneg[i] : 0.0;
  const double somr2i = (tens && inRange) ? a.auxSomr2[i] : 0.0;
  double DvDxQi[GEN ? NT : 1];
  if (GEN) {
#pragma unroll
    for (int k = 0; k < NT; ++k) DvDxQi[k] = (needQ && inRange) ? a.auxDvDxQ[i*NT + k] : 0.0;
  }
  const double fCli = (mult && inRange) ? a.auxfCl[i] : 1.0, fCqi = (mult && inRange) ? a.auxfCq[i] : 1.0;
  const double balsi = (GEN && o.balsara && inRange) ? balsara<DIM>(o, DvDxQi, Hdeti, ci) : 1.0;
  const double Cl0 = o.Cl, Cq0 = o.Cq, eps2 = o.eps2;

  // ---- accumulators
  double rhoSum = 0, norm = 0, DepsDt = 0, maxQ = 0, effQ = 0, XW = 0, m0 = 0;
  double DvDt[DIM], gradRho[DIM], XdV[DIM], m1[DIM], DvDx[NT], M[NT];
#pragma unroll
  for (int k = 0; k < DIM; ++k) { DvDt[k] = 0; gradRho[k] = 0; XdV[k] = 0; m1[k] = 0; }
#pragma unroll
  for (int k = 0; k < NT; ++k) { DvDx[k] = 0; M[k] = 0; }

  const uint32_t cnt = active ? a.nbrCount[i] : 0u;
  const uint32_t rows = a.tileRows[tile];
  const unsigned long long base = a.tileOff[tile] + lane;

  // storage mode of the pair force (sphb200_internal.cuh): one scalar on the scalar-factor isotropic path, two on the tensor paths
  constexpr int PW = ISC ? 1 : (ISO ? DIM : 2);
  double* const paccTile = compat ? a.pacc + (size_t)PW*a.tileOff[tile] + lane : nullptr;      // pacc_at(PW, slot, word)
  // ---- software pipeline over the neighbour list: indices one iteration ahead of the row copies, row copies
  //      PAIR_STAGES-1 iterations ahead of the arithmetic.  The copy of the 32 rows of one iteration is warp-cooperative:
  //      CH lanes fetch the CH 16-byte chunks of one row, so a LDGSTS instruction touches 32/CH full lines instead of 32
  //      partial ones (the per-lane gather was L1-wavefront bound: profiles/r01_notes.md).
  constexpr int CH = ROW*8/16;                       // 16-byte chunks per row: 8 (3-D) / 6 (2-D)
  // positions past the end of a lane's list fetch row 0 (never read back): no predication in the copy
  const unsigned char* const rowsB = reinterpret_cast<const unsigned char*>(a.rows);
  const unsigned char* const srcLane = rowsB + 16*(lane & (SPHB200_COPY_LANES - 1));
  auto issue_rows = [&](uint32_t p, uint32_t jraw) {  // jraw: this lane's list entry at position p (0 if none)
    const uint32_t jrow = jraw;
    const unsigned stage = warpRing + ((p + stageBase) % PAIR_STAGES)*(unsigned)STAGEB;
#if SPHB200_PAIR_AUX
    cp_async16_s(stage + 32u*ROWB + 16u*lane, a.aux2 + 2*(size_t)jrow);        // this lane's own neighbour: {det H, 1/rho}
#endif
#if SPHB200_COPY_LANES == 8
    if (CH == 8) {
      // 8 lanes per row (one 16-byte chunk each): an LDGSTS touches 4 full lines
      const unsigned dst = stage + (unsigned)(lane >> 3)*ROWB + 16u*(lane & 7);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t jr = __shfl_sync(0xffffffffu, jrow, 4*q + (lane >> 3));
        cp_async16_s(dst + (unsigned)q*(4u*ROWB), mad_wide(jr, (uint32_t)(ROW*8), srcLane));
      }
    } else
#endif
    if (CH == 8) {
      // 4 lanes per row (two 16-byte chunks each): 4 shuffles + 4 address computations feed 8 LDGSTS
      const unsigned dst = stage + (unsigned)(lane >> 2)*ROWB + 16u*(lane & 3);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint32_t jr = __shfl_sync(0xffffffffu, jrow, 8*g + (lane >> 2));
        const unsigned char* src = mad_wide(jr, (uint32_t)(ROW*8), srcLane);
        cp_async16_s(dst + (unsigned)g*(8u*ROWB), src);
        cp_async16_s(dst + (unsigned)g*(8u*ROWB) + 64u, src + 64);
      }
    } else {
#pragma unroll
      for (int q = 0; q < CH; ++q) {
        const int t = q*32 + lane;
        const int row = t/CH, chunk = t - row*CH;
        const uint32_t jr = __shfl_sync(0xffffffffu, jrow, row);
        cp_async16_s(stage + (unsigned)row*ROWB + 16u*chunk, rowsB + (size_t)jr*(ROW*8) + 16*chunk);
      }
    }
    cp_async_commit();
  };
  auto load_idx = [&](uint32_t p) -> uint32_t {
    return (p < cnt) ? a.nbr[base + (unsigned long long)p*SPHB200_TILE] : 0u;
  };
  uint32_t jn = load_idx(0u);
  // every lane is past its reads of the previous tile's ring stages before this tile's prologue refills them (a lane that left the
  // previous tile early -- a ghost, a node of another chunk -- must not run ahead of the others' last reads; racecheck reports the
  // tile-to-tile write-after-read otherwise)
#if SPHB200_TILE_SYNC == 2
  __syncwarp();
#elif SPHB200_TILE_SYNC == 3
  asm volatile("bar.warp.sync 0xffffffff;");      // execution barrier only: no compiler memory fence (the ring accesses are volatile asm themselves)
#endif
#pragma unroll
  for (uint32_t p = 0; p < (uint32_t)PAIR_STAGES - 1u; ++p) {
    issue_rows(p, jn);
    jn = load_idx(p + 1u);
  }

#pragma unroll PAIR_UNROLL
  for (uint32_t k = 0; k < rows; ++k) {
    cp_async_wait<PAIR_STAGES - 2>();                 // this lane's copies for position k have landed ...
    __syncwarp();                                     // ... and so have every other lane's
    // ---- node j state: one 128-byte (3-D) / 96-byte (2-D) row, from this lane's ring slot
    double rw[ROW];
    double2 aux;
    {
      const unsigned st = warpRing + ((k + stageBase) % PAIR_STAGES)*(unsigned)STAGEB;
      const unsigned rp = st + (unsigned)lane*ROWB;
#pragma unroll
      for (int q = 0; q < ROW/2; ++q) { const double2 v = lds128v(rp + 16u*q); rw[2*q] = v.x; rw[2*q + 1] = v.y; }
#if SPHB200_PAIR_AUX
      aux = lds128v(st + 32u*ROWB + 16u*lane);
#else
      if constexpr (ISO) { const double hj = rw[D::R_H]; aux.x = (DIM == 3) ? hj*hj*hj : hj*hj; }
      else aux.x = sym_det<DIM>(rw + D::R_H);
      aux.y = fast_rcp(rw[D::R_RHO]);
#endif
    }
    {
      // refill the stage consumed in the previous iteration (every lane is past its reads: they precede the __syncwarp above)
      const uint32_t p = k + (uint32_t)PAIR_STAGES - 1u;
      issue_rows(p, jn);
      jn = load_idx(p + 1u);
    }
    double* const paccRow = paccTile + (size_t)k*(PW*32);
    if (k < cnt) {
    const unsigned long long slot = base + (unsigned long long)k*SPHB200_TILE;
    const uint32_t j = GEN ? a.nbr[slot] : 0u;
    const double* rj = rw + D::R_POS; const double* vj = rw + D::R_VEL; const double* Hj = rw + D::R_H;
    const double mj = rw[D::R_M], rhoj = rw[D::R_RHO], cj = rw[D::R_CS];
    const double Hdetj = aux.x, rhojInv = aux.y;          // per-node values, computed once in k_pack

    // SPH.cc:363-369 : rij, eta = H.rij, |eta|, unit vectors (safeInvVar: 0 for coincident nodes)
    double rij[DIM], etai[DIM], etaj[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) rij[q] = ri[q] - rj[q];
    double e2i, e2j, Wi, gWi, Wj, gWj, gWiRaw;
    double gradWi[DIM], gradWj[DIM];
    double gi = 0.0, gj = 0.0, hjInv = 0.0;               // ISO: gradW_i = gi * rij, gradW_j = gj * rij
    double sWi = 0.0, sWj = 0.0, sQi = 0.0, sQj = 0.0;    // tensor paths: gradW_x = sW_x H_x.eta_x, gradWQ_x = sQ_x H_x.eta_x
    if constexpr (ISO) {
      hjInv = Hj[0];
      const double r2 = vdot<DIM>(rij, rij);
      // coincident nodes: r = 0 multiplies the finite reciprocal below, as safeInvVar's 0 * 1e30 does in the reference
      const double rinv = fast_rsqrt(r2 + 1.0e-300);
      const double r = r2*rinv;
      e2i = hi2*r2; e2j = (hjInv*hjInv)*r2;
      table_eval_raw(tW, a.kextW, a.xminW, a.xstepW, rxW, a.n1W, hiInv*r, Wi, gWi);
      if (xsph) table_eval_raw(tW, a.kextW, a.xminW, a.xstepW, rxW, a.n1W, hjInv*r, Wj, gWj);
      else { Wj = 0.0; table_eval_grad(tW, a.kextW, a.xminW, a.xstepW, rxW, a.n1W, hjInv*r, gWj); }
      gWiRaw = gWi;
      Wi *= Hdeti; Wj *= Hdetj;
      gi = (gWi*Hdeti)*(hiInv*rinv);                       // gWi * Hi * etaiUnit = gWi * hiInv * rij/r
      gj = (gWj*Hdetj)*(hjInv*rinv);
      if constexpr (!ISC) {
#pragma unroll
        for (int q = 0; q < DIM; ++q) { etai[q] = hiInv*rij[q]; etaj[q] = hjInv*rij[q]; gradWi[q] = gi*rij[q]; gradWj[q] = gj*rij[q]; }
      }
    } else {
    sym_dot<DIM>(Hi, rij, etai);
    sym_dot<DIM>(Hj, rij, etaj);
    e2i = vdot<DIM>(etai, etai); e2j = vdot<DIM>(etaj, etaj);
    // coincident nodes (eta == 0): 1e-300 keeps the reciprocal finite; it multiplies exact zeros (H.eta, eta^2) below, which
    // reproduces safeInvVar's 0 * 1e30 = 0 (SPH.cc:368-369); for every other pair the addend is below half an ulp
    const double invi = fast_rsqrt(e2i + 1.0e-300);
    const double invj = fast_rsqrt(e2j + 1.0e-300);
    const double etaMagi = e2i*invi, etaMagj = e2j*invj;

    // SPH.cc:374-377 : W, gradW (table values carry no Hdet yet).  W_j only enters the XSPH weight and the tensile term: without
    // them the j side reads the gradient coefficients only (two of the record's three 16-byte chunks)
    table_eval_raw(tW, a.kextW, a.xminW, a.xstepW, rxW, a.n1W, etaMagi, Wi, gWi);
    if (GEN || xsph) table_eval_raw(tW, a.kextW, a.xminW, a.xstepW, rxW, a.n1W, etaMagj, Wj, gWj);
    else { Wj = 0.0; table_eval_grad(tW, a.kextW, a.xminW, a.xstepW, rxW, a.n1W, etaMagj, gWj); }
    gWiRaw = gWi;
    Wi *= Hdeti; gWi *= Hdeti; Wj *= Hdetj; gWj *= Hdetj;
    double Hei[DIM], Hej[DIM];
    sym_dot<DIM>(Hi, etai, Hei);                 // gWi*Hi*etaiUnit == (gWi/|etai|) * (Hi.etai)
    sym_dot<DIM>(Hj, etaj, Hej);
    { const double si = gWi*invi, sj = gWj*invj;
      sWi = sQi = si; sWj = sQj = sj;
#pragma unroll
      for (int q = 0; q < DIM; ++q) { gradWi[q] = si*Hei[q]; gradWj[q] = sj*Hej[q]; } }
    }
    double WQi = Wi, WQj = Wj, gradWQi[DIM], gradWQj[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { gradWQi[q] = gradWi[q]; gradWQj[q] = gradWj[q]; }
    if constexpr (GEN) {
    if (twoK) {                                   // SPH.cc:383-388
      double gWQi, gWQj;
      const double invi = fast_rsqrt(e2i + 1.0e-300), invj = fast_rsqrt(e2j + 1.0e-300);
      double Hei[DIM], Hej[DIM];
      sym_dot<DIM>(Hi, etai, Hei);
      sym_dot<DIM>(Hj, etaj, Hej);
      table_eval_raw(tQ, a.kextQ, a.xminQ, a.xstepQ, rxQ, a.n1Q, e2i*invi, WQi, gWQi);
      table_eval_raw(tQ, a.kextQ, a.xminQ, a.xstepQ, rxQ, a.n1Q, e2j*invj, WQj, gWQj);
      WQi *= Hdeti; WQj *= Hdetj;
      const double si = gWQi*Hdeti*invi, sj = gWQj*Hdetj*invj;
      sQi = si; sQj = sj;
#pragma unroll
      for (int q = 0; q < DIM; ++q) { gradWQi[q] = si*Hei[q]; gradWQj[q] = sj*Hej[q]; }
    }
    }

    // SPH.cc:391-396 (i side)
    rhoSum = fma(mj, Wi, rhoSum);
    norm = fma(mi_over_rhoi, Wi, norm);

    // ---- artificial viscosity: MonaghanGingoldViscosity.cc:69-100 / LimitedMonaghanGingoldViscosity.cc:140-218
    double vij[DIM], vijQ[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) { vij[q] = vi[q] - vj[q]; vijQ[q] = vij[q]; }
    double Clij = Cl0, Cqij = Cq0;
    if (GEN) {
      double fshear = 1.0, fClj = 1.0, fCqj = 1.0;
      if (mult) { fClj = a.auxfCl[j]; fCqj = a.auxfCq[j]; }
      if (needQ) {
        double DvDxQj[NT];
#pragma unroll
        for (int q = 0; q < NT; ++q) DvDxQj[q] = a.auxDvDxQ[(size_t)j*NT + q];
        if (o.balsara) fshear = 0.5*(balsi + balsara<DIM>(o, DvDxQj, Hdetj, cj));
        if (o.Qkind == SPHB200_Q_LIMITED_MG) {
          const double etaCrit = o.etaCritFrac/o.nPerh, etaFold = o.etaFoldFrac/o.nPerh;
          double xij[DIM], t1[DIM], t2[DIM];
#pragma unroll
          for (int q = 0; q < DIM; ++q) xij[q] = 0.5*rij[q];
          ten_dot<DIM>(DvDxQi, xij, t1); const double gradi = vdot<DIM>(t1, xij);
          ten_dot<DIM>(DvDxQj, xij, t2); const double gradj = vdot<DIM>(t2, xij);
          const double rri = gradi/(d_sgn(gradj)*fmax(1.0e-30, fabs(gradj)));
          const double rrj = gradj/(d_sgn(gradi)*fmax(1.0e-30, fabs(gradi)));
          const double x = fmin(rri, rrj);
          double phi = (x > 0.0 ? 2.0/(1.0 + x)*2.0*x/(1.0 + x) : 0.0);       // van Leer
          const double e2m = e2i < e2j ? e2i : e2j;
          const double etaij = e2m*fast_rsqrt(e2m + 1.0e-300);
          if (etaij < etaCrit) { const double z = (etaij - etaCrit)/etaFold; phi *= exp(-z*z); }
#pragma unroll
          for (int q = 0; q < DIM; ++q) vijQ[q] = (vi[q] - phi*t1[q]) - (vj[q] + phi*t2[q]);
        }
      }
      Clij = 0.5*(fCli + fClj)*fshear*Cl0;
      Cqij = 0.5*(fCqi + fCqj)*fshear*Cq0;
    }
    if constexpr (ISC) {
      // every pair vector below is a scalar times r_ij
      const double vr = vdot<DIM>(vij, rij);
      const double mui = (hiInv*vr)*fast_rcp(e2i + eps2), muj = (hjInv*vr)*fast_rcp(e2j + eps2);
      const double mui0 = mui < 0.0 ? mui : 0.0, muj0 = muj < 0.0 ? muj : 0.0;
      const double ei = fma(Cqij*mui0, mui0, -(Clij*ci)*mui0), ej = fma(Cqij*muj0, muj0, -(Clij*cj)*muj0);
      const double hQPiij = 0.5*(ei*rhoiInv), hQPiji = 0.5*(ej*rhojInv);     // 0.5*QPi (SPH.cc:405-406)
      const double Qi = rhoi*ei;
      maxQ = Qi > maxQ ? Qi : maxQ;
      effQ = fma(mj*Qi, Wi*rhojInv, effQ);
      // SPH.cc:405-426 : delta = (Prho_i + QPi_ij/2) gradW_i + (Prho_j + QPi_ji/2) gradW_j = sd r_ij
      const double ai = Prhoi0 + hQPiij, aj = rw[D::R_PRHO] + hQPiji;
      const double sd = fma(ai, gi, aj*gj);
      const double msd = mj*sd;
#pragma unroll
      for (int q = 0; q < DIM; ++q) DvDt[q] = fma(-msd, rij[q], DvDt[q]);
      if (compat) paccRow[0] = sd;                                           // delta = sd r_ij, expanded by the readers
      // SPH.cc:434 : mj (Prho_i vij.gradW_i + workQ_i) = mj gi (vij.rij) (Prho_i + QPi_ij/2)
      const double mg = mj*gi;
      DepsDt = fma(mg*vr, ai, DepsDt);
      // SPH.cc:438-445, 457-468
      { double g[DIM];
#pragma unroll
        for (int q = 0; q < DIM; ++q) g[q] = mg*rij[q];
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int c2 = 0; c2 < DIM; ++c2) DvDx[r*DIM + c2] = fma(-vij[r], g[c2], DvDx[r*DIM + c2]);
        int t = 0;
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int c2 = r; c2 < DIM; ++c2) { M[t] = fma(-rij[r], g[c2], M[t]); ++t; }
        const double f = rhoj - rhoi;
#pragma unroll
        for (int q = 0; q < DIM; ++q) gradRho[q] = fma(f, g[q], gradRho[q]);
      }
      if (xsph) {                                                            // SPH.cc:448-454
        const double w = 0.5*fma(mi_over_rhoi, Wi, mj*rhojInv*Wj);
        XW += w;
#pragma unroll
        for (int q = 0; q < DIM; ++q) XdV[q] = fma(-w, vij[q], XdV[q]);
      }
      if (hsph) {                                                            // SPHSmoothingScale.cc:186-222
        const double WSPHi = fabs(gWiRaw);
        m0 += WSPHi;
        const double wh = WSPHi*hiInv;
#pragma unroll
        for (int q = 0; q < DIM; ++q) m1[q] = fma(-wh, rij[q], m1[q]);
      }
    } else {
    const double mui = vdot<DIM>(vijQ, etai)*fast_rcp(e2i + eps2);
    const double muj = vdot<DIM>(vijQ, etaj)*fast_rcp(e2j + eps2);
    const double mui0 = mui < 0.0 ? mui : 0.0, muj0 = muj < 0.0 ? muj : 0.0;
    double ei, ej;
    if (GEN && (o.linearInExpansion || o.quadraticInExpansion)) {
      ei = -Clij*ci*(o.linearInExpansion ? mui : mui0) + Cqij*(o.quadraticInExpansion ? -d_sgn(mui)*mui*mui : mui0*mui0);
      ej = -Clij*cj*(o.linearInExpansion ? muj : muj0) + Cqij*(o.quadraticInExpansion ? -d_sgn(muj)*muj*muj : muj0*muj0);
    } else {
      ei = fma(Cqij*mui0, mui0, -(Clij*ci)*mui0);
      ej = fma(Cqij*muj0, muj0, -(Clij*cj)*muj0);
    }
    const double hQPiij = 0.5*(ei*rhoiInv), hQPiji = 0.5*(ej*rhojInv);     // 0.5*QPi (SPH.cc:405-406)
    const double Qi = rhoi*ei;

    // SPH.cc:405-414
    double Qacc[DIM];                                                     // Qacci + Qaccj
    double workQi = 0.0;
#pragma unroll
    for (int q = 0; q < DIM; ++q) {
      const double qi = hQPiij*gradWQi[q];
      workQi = fma(vij[q], qi, workQi);
      Qacc[q] = fma(hQPiji, gradWQj[q], qi);
    }
    maxQ = Qi > maxQ ? Qi : maxQ;
    effQ = fma(mj*Qi, WQi*rhojInv, effQ);

    // SPH.cc:417-426
    double Prhoi = Prhoi0, Prhoj = rw[D::R_PRHO];
    if (tens) {
      const double t = Wi/(Hdeti*a.WnPerh), u = Wj/(Hdetj*a.WnPerh);
      Prhoi += somr2i*(o.epsTensile*(t*t*t*t)*Pnegi);
      Prhoj += a.auxSomr2[j]*(o.epsTensile*(u*u*u*u)*a.auxPneg[j]);
    }
    double delta[DIM];
#pragma unroll
    for (int q = 0; q < DIM; ++q) delta[q] = fma(Prhoi, gradWi[q], fma(Prhoj, gradWj[q], Qacc[q]));
#pragma unroll
    for (int q = 0; q < DIM; ++q) DvDt[q] = fma(-mj, delta[q], DvDt[q]);
    if (compat) {
      if constexpr (ISO) {
#pragma unroll
        for (int q = 0; q < DIM; ++q) paccRow[32*q] = delta[q];
      } else {
        // delta = (Prho_i sW_i + QPi_ij/2 sQ_i) H_i.eta_i + (Prho_j sW_j + QPi_ji/2 sQ_j) H_j.eta_j: two scalars, expanded by the readers
        paccRow[0] = fma(Prhoi, sWi, hQPiij*sQi);
        paccRow[32] = fma(Prhoj, sWj, hQPiji*sQj);
      }
    }

    // SPH.cc:434
    DepsDt = fma(mj, fma(Prhoi, vdot<DIM>(vij, gradWi), workQi), DepsDt);

    // SPH.cc:438-445, 463-468 : DvDx -= mj*vij (x) gradWi ; M -= mj*rij (x) gradWi
    { double g[DIM];
#pragma unroll
      for (int q = 0; q < DIM; ++q) g[q] = mj*gradWi[q];
#pragma unroll
      for (int r = 0; r < DIM; ++r)
#pragma unroll
        for (int c2 = 0; c2 < DIM; ++c2) DvDx[r*DIM + c2] = fma(-vij[r], g[c2], DvDx[r*DIM + c2]);
      if constexpr (ISO) {
        // rij (x) gradWi = gi rij (x) rij is symmetric: upper triangle only (mirrored in the finalize)
        int t = 0;
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int c2 = r; c2 < DIM; ++c2) { M[t] = fma(-rij[r], g[c2], M[t]); ++t; }
      } else {
#pragma unroll
        for (int r = 0; r < DIM; ++r)
#pragma unroll
          for (int c2 = 0; c2 < DIM; ++c2) M[r*DIM + c2] = fma(-rij[r], g[c2], M[r*DIM + c2]);
      }
      // SPH.cc:457-460
      const double f = rhoj - rhoi;
#pragma unroll
      for (int q = 0; q < DIM; ++q) gradRho[q] = fma(f, g[q], gradRho[q]);
    }

    // SPH.cc:448-454
    if (xsph) {
      const double w = 0.5*fma(mi_over_rhoi, Wi, mj*rhojInv*Wj);
      XW += w;
#pragma unroll
      for (int q = 0; q < DIM; ++q) XdV[q] = fma(-w, vij[q], XdV[q]);
    }

    // SPHSmoothingScale.cc:186-222 (i side; same NodeList, Cartesian => fweightij = 1): WSPHi = |gradW table(eta_i)|
    if (hsph) {
      const double WSPHi = fabs(gWiRaw);
      m0 += WSPHi;
#pragma unroll
      for (int q = 0; q < DIM; ++q) m1[q] = fma(-WSPHi, etai[q], m1[q]);
    }
    }
    }
  }

  cp_async_wait<0>();
#if SPHB200_TILE_SYNC == 1
  __syncwarp();
#elif SPHB200_TILE_SYNC == 4
  asm volatile("bar.warp.sync 0xffffffff;");
#elif SPHB200_TILE_SYNC == 5
  stageBase = (stageBase + rows) % (uint32_t)PAIR_STAGES;      // the next prologue fills every stage but (rows - 1 + old base) % PAIR_STAGES
#endif
  if (!inRange) continue;
  // ---- K4: per-node finalize (SPH.cc:480-552); ghost nodes get zeros
  const size_t cap = a.cap;
  auto put = [&](int slot, int comp, double v) { a.deriv[slot][(size_t)comp*cap + i] = v; };
  if (!active) {
    if (internal) continue;                          // a node of another chunk: its launch writes it
    for (int s = 0; s < DV_COUNT; ++s) { const int w = sphb200_deriv_width(DIM, s); for (int q = 0; q < w; ++q) put(s, q, 0.0); }
    continue;
  }
  rhoSum += mi*a.W0*Hdeti;
  norm += mi_over_rhoi*a.W0*Hdeti;
  if constexpr (ISO) {                                   // unpack the upper triangle accumulated above
    double S[NS];
#pragma unroll
    for (int q = 0; q < NS; ++q) S[q] = M[q];
    if (DIM == 3) { M[0] = S[0]; M[1] = S[1]; M[2] = S[2]; M[3] = S[1]; M[4] = S[3]; M[5] = S[4]; M[6] = S[2]; M[7] = S[4]; M[8] = S[5]; }
    else { M[0] = S[0]; M[1] = S[1]; M[2] = S[1]; M[3] = S[2]; }
  }
  double Minv[NT], DvDxF[NT];
  const uint32_t pownu2 = (DIM == 3) ? 8u : 4u;
  if (o.correctVelocityGradient && fabs(ten_det<DIM>(M)) > 1.0e-10 && cnt > pownu2) {
    ten_inverse<DIM>(M, Minv);
    ten_mul<DIM>(DvDx, Minv, DvDxF);
  } else {
#pragma unroll
    for (int q = 0; q < NT; ++q) { Minv[q] = M[q]; DvDxF[q] = DvDx[q]*rhoiInv; }
  }
#pragma unroll
  for (int q = 0; q < DIM; ++q) put(DV_GRADRHO, q, gradRho[q]*rhoiInv);
  put(DV_DRHODT, 0, -rhoi*ten_trace<DIM>(DvDxF));
  if (o.evolveTotalEnergy) DepsDt = mi*(vdot<DIM>(vi, DvDt) + DepsDt);
  put(DV_DEPSDT, 0, DepsDt);
  if (xsph) {
    XW += Hdeti*mi/rhoi*a.W0;
    const double deninv = 1.0/fmax(tiny, XW);
#pragma unroll
    for (int q = 0; q < DIM; ++q) put(DV_DXDT, q, vi[q] + XdV[q]*deninv);
  } else {
#pragma unroll
    for (int q = 0; q < DIM; ++q) put(DV_DXDT, q, vi[q]);
  }
  put(DV_RHOSUM, 0, rhoSum); put(DV_NORM, 0, norm); put(DV_MAXQ, 0, maxQ); put(DV_EFFQ, 0, effQ); put(DV_XSPHW, 0, XW);
#pragma unroll
  for (int q = 0; q < DIM; ++q) { put(DV_DVDT, q, DvDt[q]); put(DV_XSPHDV, q, XdV[q]); }
#pragma unroll
  for (int q = 0; q < NT; ++q) { put(DV_DVDX, q, DvDxF[q]); put(DV_LOCALDVDX, q, DvDxF[q]); put(DV_M, q, Minv[q]); put(DV_LOCALM, q, Minv[q]); }

  // ---- K5: smoothing scale
  if (hsph) {
    const double z0 = rootnu<DIM>(fmax(0.0, m0));
    put(DV_M0, 0, z0);
#pragma unroll
    for (int q = 0; q < DIM; ++q) put(DV_M1, q, m1[q]);
    const double tr = ten_trace<DIM>(DvDxF);
    const double dinv = 1.0/(double)DIM;
#pragma unroll
    for (int q = 0; q < NS; ++q) put(DV_DHDT, q, ((-Hi[q])*dinv)*tr);
    const bool isolated = fabs(z0 - 0.0) <= 1.0e-15*fmax(1.0, fabs(z0));                 // fuzzyEqual(z0, 0)
    const double cur = isolated ? 0.5*o.nPerh : fmax(0.0, hermite_eval(a.nperhVals, a.nperhN, a.nperhXmin, a.nperhXmax, a.nperhXstep, z0));
    const double sv = fmin(4.0, fmax(0.25, o.nPerh/(cur + 1.0e-30)));
    const double aa = (sv < 1.0 ? 0.4*(1.0 + sv*sv) : 0.4*(1.0 + 1.0/(sv*sv*sv)));
    const double hi0 = 1.0/Hi[0];
    const double hi1 = fmin(o.hmax, fmax(o.hmin, hi0*(1.0 - aa + aa*sv)));
    const double hinv = 1.0/hi1;
    if (DIM == 3) { put(DV_HIDEAL, 0, hinv); put(DV_HIDEAL, 1, 0.0); put(DV_HIDEAL, 2, 0.0); put(DV_HIDEAL, 3, hinv); put(DV_HIDEAL, 4, 0.0); put(DV_HIDEAL, 5, hinv); }
    else { put(DV_HIDEAL, 0, hinv); put(DV_HIDEAL, 1, 0.0); put(DV_HIDEAL, 2, hinv); }
  } else {
    put(DV_M0, 0, 0.0);
#pragma unroll
    for (int q = 0; q < DIM; ++q) put(DV_M1, q, 0.0);
    double dh[NS];
    if (o.hEvolution == SPHB200_H_ASPH || o.hEvolution == SPHB200_H_ASPH_CLASSIC) asph_DHDt<DIM>(Hi, DvDxF, dh);   // the classic package's ideal H follows in k_asph_classic
    else {
#pragma unroll
      for (int q = 0; q < NS; ++q) dh[q] = 0.0;
    }
#pragma unroll
    for (int q = 0; q < NS; ++q) { put(DV_DHDT, q, dh[q]); put(DV_HIDEAL, q, 0.0); }
  }
  }   // tile loop
}

template <int DIM, bool GEN, bool X, bool H, bool C, bool ISO = false>
int launch_one(sphb200_ctx* c, const DerivArgs& a, unsigned nb, int threads, size_t shm) {
  CU_CHECK(c, cudaFuncSetAttribute(k_sph_derivs<DIM, GEN, X, H, C, ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
  k_sph_derivs<DIM, GEN, X, H, C, ISO><<<nb, threads, shm, c->stream>>>(a);
  return 0;
}
template <int DIM>
int launch_dim(sphb200_ctx* c, const DerivArgs& a, unsigned nb, int threads, size_t shm, bool gen) {
  const bool x = a.o.XSPH != 0, h = a.o.hEvolution == SPHB200_H_SPH, p = a.o.compatibleEnergy != 0;
  if (gen) return launch_one<DIM, true, true, true, true>(c, a, nb, threads, shm);
  const int code = (x ? 4 : 0) | (h ? 2 : 0) | (p ? 1 : 0);
  if (c->allIsotropic && !sphb200_is_asph(a.o.hEvolution)) {      // every H = h^-1 I (k_pack checked the rows this launch reads)
    switch (code) {
      case 0: return launch_one<DIM, false, false, false, false, true>(c, a, nb, threads, shm);
      case 1: return launch_one<DIM, false, false, false, true, true>(c, a, nb, threads, shm);
      case 2: return launch_one<DIM, false, false, true, false, true>(c, a, nb, threads, shm);
      case 3: return launch_one<DIM, false, false, true, true, true>(c, a, nb, threads, shm);
      case 4: return launch_one<DIM, false, true, false, false, true>(c, a, nb, threads, shm);
      case 5: return launch_one<DIM, false, true, false, true, true>(c, a, nb, threads, shm);
      case 6: return launch_one<DIM, false, true, true, false, true>(c, a, nb, threads, shm);
      default: return launch_one<DIM, false, true, true, true, true>(c, a, nb, threads, shm);
    }
  }
  switch (code) {
    case 0: return launch_one<DIM, false, false, false, false>(c, a, nb, threads, shm);
    case 1: return launch_one<DIM, false, false, false, true>(c, a, nb, threads, shm);
    case 2: return launch_one<DIM, false, false, true, false>(c, a, nb, threads, shm);
    case 3: return launch_one<DIM, false, false, true, true>(c, a, nb, threads, shm);
    case 4: return launch_one<DIM, false, true, false, false>(c, a, nb, threads, shm);
    case 5: return launch_one<DIM, false, true, false, true>(c, a, nb, threads, shm);
    case 6: return launch_one<DIM, false, true, true, false>(c, a, nb, threads, shm);
    default: return launch_one<DIM, false, true, true, true>(c, a, nb, threads, shm);
  }
}

}  // namespace

static double host_table_eval(const TableDev& t, double eta, bool grad) {
  if (!(eta < t.kext)) return 0.0;
  double q = std::max(0.0, eta - t.xmin)/t.xstep;
  size_t k = std::min<size_t>((size_t)q, t.n1);
  const std::vector<double>& c = grad ? t.hostG : t.hostW;
  return c[3*k] + (c[3*k + 1] + c[3*k + 2]*eta)*eta;
}

int sphb200_launch_derivs(sphb200_ctx* c) { return sphb200_launch_derivs_chunk(c, nullptr, 0u, 0u, 0xffffffffu); }

// One launch of the pair loop over the tiles of `tileList` (nullptr: all), internal nodes with original index in [origLo, origHi).
int sphb200_launch_derivs_chunk(sphb200_ctx* c, const uint32_t* tileList, uint32_t nList, uint32_t origLo, uint32_t origHi) {
  DerivArgs a{};
  a.tileList = tileList; a.nList = nList; a.origLo = origLo; a.origHi = origHi;
  a.rows = c->rows; a.aux2 = c->aux2; a.perm = c->perm; a.nbrCount = c->nbrCount; a.tileRows = c->tileRows; a.tileOff = c->tileOff; a.nbr = c->nbr;
  const bool tens = c->opt.epsTensile != 0.0;
  const bool needQ = (c->opt.Qkind == SPHB200_Q_LIMITED_MG) || c->opt.balsara;
  const bool mult = c->have[S_FCL] && c->have[S_FCQ];
  a.auxPneg = tens ? c->auxPneg : nullptr; a.auxSomr2 = tens ? c->auxSomr2 : nullptr;
  a.auxDvDxQ = needQ ? c->auxDvDxQ : nullptr; a.auxfCl = mult ? c->auxfCl : nullptr; a.auxfCq = mult ? c->auxfCq : nullptr;
  a.tabW = c->W.coef; a.kextW = c->W.kext; a.xminW = c->W.xmin; a.xstepW = c->W.xstep; a.n1W = c->W.n1;
  a.oneKernel = c->oneKernel ? 1 : 0;
  const TableDev& Q = c->oneKernel ? c->W : c->WQ;
  a.tabQ = Q.coef; a.kextQ = Q.kext; a.xminQ = Q.xmin; a.xstepQ = Q.xstep; a.n1Q = Q.n1;
  a.nperhVals = c->W.nperhVals; a.nperhN = c->W.nperhN; a.nperhXmin = c->W.nperhXmin; a.nperhXmax = c->W.nperhXmax; a.nperhXstep = c->W.nperhXstep;
  if (c->opt.hEvolution == SPHB200_H_SPH && (!a.nperhVals || a.nperhN < 2))
    return sphb200_fail(c, "evaluateDerivatives: SPHSmoothingScale needs the TableKernel nperh lookup (nperhVals) but none was set");
  a.W0 = host_table_eval(c->W, 0.0, false);                       // SPH.cc:189
  a.WnPerh = host_table_eval(c->W, 1.0/c->opt.nPerh, false);      // SPH.cc:264-266
  a.n = c->n; a.cap = c->cap; a.nInt = (uint32_t)c->nInt; a.o = c->opt;
  for (int s = 0; s < DV_COUNT; ++s) a.deriv[s] = c->deriv[s];
  const bool isoPath0 = c->allIsotropic && !sphb200_is_asph(c->opt.hEvolution);
  const bool gen0 = (c->opt.Qkind != SPHB200_Q_MG) || c->opt.balsara || mult || tens || !c->oneKernel || c->opt.linearInExpansion || c->opt.quadraticInExpansion;
  c->paccMode = (isoPath0 && !gen0) ? (SPHB200_ISO_SCALAR ? PACC_ISO : PACC_FULL) : PACC_TENSOR;
  c->paccWidth = pacc_width(c->paccMode, c->ndim);
  if (c->opt.compatibleEnergy) {
    if (sphb200_ensure(c, c->pacc, c->paccCap, c->nSlots*(size_t)c->paccWidth)) return 1;
  }
  a.pacc = c->pacc; a.nSlots = c->nSlots;
  const int wpb = PAIR_WARPS;
  const size_t ringBytes = (size_t)PAIR_WARPS*PAIR_STAGES*(c->ndim == 3 ? RingGeom<3>::STAGEB : RingGeom<2>::STAGEB);
  const size_t shm = (size_t)6*(c->W.n1 + 2)*sizeof(double) + (c->oneKernel ? 0 : (size_t)6*(c->WQ.n1 + 2)*sizeof(double)) + ringBytes;
  if (shm > 226*1024) return sphb200_fail(c, "kernel table too large for shared memory");
  // persistent CTAs: 2 per SM (register-limited), each striding over the tiles
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, c->device);
  const bool isoPath = c->allIsotropic && !sphb200_is_asph(c->opt.hEvolution);
  const size_t tilesHere = tileList ? (size_t)nList : c->nTiles;
  if (tilesHere == 0) return 0;
  const unsigned nb = (unsigned)std::min<size_t>((tilesHere + wpb - 1)/wpb, (size_t)nsm*(isoPath ? SPHB200_PAIR_CTAS_ISO : PAIR_CTAS));
  const bool gen = (c->opt.Qkind != SPHB200_Q_MG) || c->opt.balsara || mult || tens || !c->oneKernel ||
                   c->opt.linearInExpansion || c->opt.quadraticInExpansion;
  if (c->ndim == 3) { if (launch_dim<3>(c, a, nb, wpb*32, shm, gen)) return 1; }
  else              { if (launch_dim<2>(c, a, nb, wpb*32, shm, gen)) return 1; }
  KERNEL_CHECK(c, "k_sph_derivs");
  c->derivsValid = true;
  c->rowsAtEval = true;
  return 0;
}
