// nbr_ring.cuh -- warp-cooperative streaming of neighbour records into a shared-memory ring (cp.async / LDGSTS).
//
// Every pair-loop kernel of this library is an i-centric gather: lane <-> node i walks its own neighbour list, so in one
// iteration the 32 lanes of a warp need the records of 32 *different* nodes j.  A per-lane load of a 128-byte record touches 32
// different lines per instruction and the kernel becomes L1/TEX-wavefront bound (k_crk_* before this ring: 79-86 % L1TEX
// throughput at 12-19 % FP64 pipe, profiles/r01_notes.md).  Here the warp copies the 32 records cooperatively -- 8 lanes
// fetch the 8 16-byte chunks of one 128-byte record, so an LDGSTS touches 4 full lines -- into a ring of STAGES stages,
// STAGES-1 iterations ahead of the arithmetic, which also takes the L2 latency of the gather off the critical path.
//
// Stage layout (per warp):  32 x RB node rows | 32 x X1B first extra record | 32 x X2B second extra record | 32 x 16 B aux2 (optional)
// Record strides are padded to an odd number of 16-byte units so that the per-lane 128-bit reads are bank-conflict free.
#pragma once
#include "sphb200_internal.cuh"

namespace {

__device__ __forceinline__ void ring_cp16(unsigned smemDst, const void* gmemSrc) {
#if defined(SPHB200_CP_CG) && SPHB200_CP_CG
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smemDst), "l"(gmemSrc) : "memory");     // L2 only
#else
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smemDst), "l"(gmemSrc) : "memory");
#endif
}
__device__ __forceinline__ void ring_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void ring_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ const unsigned char* ring_mad_wide(uint32_t a, uint32_t b, const unsigned char* c) {   // c + a*b in one IMAD.WIDE
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"((unsigned long long)c));
  return reinterpret_cast<const unsigned char*>(r);
}
__device__ __forceinline__ double2 ring_lds128(unsigned addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}

#ifndef SPHB200_RING_ROTATE
#define SPHB200_RING_ROTATE 1
#endif
constexpr int ring_pad(int bytes) { return bytes == 0 ? 0 : (((bytes/16) & 1) ? bytes : bytes + 16); }

// the leading BYTES of 32 records (STRIDE bytes apart in memory, record index jrow of every lane) -> rows of stride RB at dst0
template <int BYTES, int RB, int STRIDE = BYTES>
__device__ __forceinline__ void ring_copy_records(unsigned dst0, const unsigned char* __restrict__ src, uint32_t jrow, int lane) {
  constexpr int CH = BYTES/16;
  static_assert(BYTES % 16 == 0 && STRIDE % 16 == 0, "records are multiples of 16 bytes");
  if (CH == 8) {
    const unsigned dst = dst0 + (unsigned)(lane >> 3)*RB + 16u*(lane & 7);
    const unsigned char* const s = src + 16*(lane & 7);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t jr = __shfl_sync(0xffffffffu, jrow, 4*q + (lane >> 3));
      ring_cp16(dst + (unsigned)q*(4u*RB), ring_mad_wide(jr, (uint32_t)STRIDE, s));
    }
  } else if (CH == 2 || CH == 4) {
    // CH lanes per record: 32/CH records per instruction
    constexpr int PER = 32/CH;
    const unsigned dst = dst0 + (unsigned)(lane/CH)*RB + 16u*(lane % CH);
    const unsigned char* const s = src + 16*(lane % CH);
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const uint32_t jr = __shfl_sync(0xffffffffu, jrow, PER*q + lane/CH);
      ring_cp16(dst + (unsigned)q*((unsigned)PER*RB), ring_mad_wide(jr, (uint32_t)STRIDE, s));
    }
  } else {
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const int t = q*32 + lane;
      const int row = t/CH, chunk = t - row*CH;
      const uint32_t jr = __shfl_sync(0xffffffffu, jrow, row);
      ring_cp16(dst0 + (unsigned)row*RB + 16u*chunk, ring_mad_wide(jr, (uint32_t)STRIDE, src + 16*chunk));
    }
  }
}

// ROWCOPY: leading bytes of the node row a kernel needs (the position leads the row, then velocity, H, m, rho, P/rho^2, cs)
// AUX: also copy this lane's own 16-byte per-node record (aux2).  A per-lane 16-byte copy costs 32 shared-memory wavefronts per
// iteration -- as many as the 32 node rows together -- and the pair loops are bound by exactly that (LSU data-pipe wavefronts
// 93 % busy, profiles/r01_notes.md), so kernels that can recompute those two numbers from the row leave it out.
template <int DIM, int X1BYTES, int X2BYTES, int STAGES, int ROWCOPY = Dm<DIM>::ROW*8, bool AUX = true>
struct NbrRing {
  static constexpr int ROWBYTES = Dm<DIM>::ROW*8;
  static constexpr int RB = ring_pad(ROWCOPY);
  static constexpr int X1B = ring_pad(X1BYTES), X2B = ring_pad(X2BYTES);
  static constexpr int X1OFF = 32*RB, X2OFF = X1OFF + 32*X1B, AUXOFF = X2OFF + 32*X2B;
  static constexpr int STAGEB = AUXOFF + (AUX ? 32*16 : 0);
  static constexpr int WARPB = STAGES*STAGEB;

  unsigned base;                                  // 32-bit shared address of this warp's ring
  const unsigned char *rows, *x1, *x2, *aux2;
  mutable unsigned rot = 0u;                      // stage of list position 0 of the current walk (ring_walk rotates it from walk to walk)

  __device__ __forceinline__ unsigned stage(uint32_t p) const { return base + ((p + rot) % STAGES)*(unsigned)STAGEB; }
  // jrow: this lane's list entry at position p (0 past the end of the lane's list: row 0 is fetched and never read back)
  __device__ __forceinline__ void issue(uint32_t p, uint32_t jrow, int lane) const {
    const unsigned st = stage(p);
    if (AUX) ring_cp16(st + AUXOFF + 16u*lane, aux2 + 16*(size_t)jrow);
    ring_copy_records<ROWCOPY, RB, ROWBYTES>(st, rows, jrow, lane);
    if (X1BYTES) ring_copy_records<X1BYTES == 0 ? 16 : X1BYTES, X1B == 0 ? 16 : X1B>(st + X1OFF, x1, jrow, lane);
    if (X2BYTES) ring_copy_records<X2BYTES == 0 ? 16 : X2BYTES, X2B == 0 ? 16 : X2B>(st + X2OFF, x2, jrow, lane);
    ring_commit();
  }
  template <int N> __device__ __forceinline__ void read(unsigned addr, double* out) const {
#pragma unroll
    for (int q = 0; q < N/2; ++q) { const double2 v = ring_lds128(addr + 16u*q); out[2*q] = v.x; out[2*q + 1] = v.y; }
  }
  __device__ __forceinline__ void read_row(uint32_t k, int lane, double* rw) const { read<ROWCOPY/8>(stage(k) + (unsigned)lane*RB, rw); }
  template <int N> __device__ __forceinline__ void read_x1(uint32_t k, int lane, double* o) const { read<N>(stage(k) + X1OFF + (unsigned)lane*X1B, o); }
  template <int N> __device__ __forceinline__ void read_x2(uint32_t k, int lane, double* o) const { read<N>(stage(k) + X2OFF + (unsigned)lane*X2B, o); }
  __device__ __forceinline__ double2 read_aux(uint32_t k, int lane) const { return ring_lds128(stage(k) + AUXOFF + 16u*lane); }
};

// The pipelined walk over one tile's neighbour list.  `body(k, j)` runs for every list position k < cnt of this lane with the
// records of neighbour j readable from ring stage k; it must read them before returning.
//   load_idx(p): this lane's list entry at position p (0 past the end)
// Two queues run ahead of the arithmetic: list entries are loaded IDX iterations before they are needed as copy addresses (the
// index load is an L2 round trip of its own; one iteration ahead left it on the critical path of the short loops), and record
// copies are issued STAGES-1 iterations before the body reads them.
#ifndef SPHB200_RING_IDX_AHEAD
#define SPHB200_RING_IDX_AHEAD 4
#endif
template <typename Ring, int STAGES, typename LoadIdx, typename Body>
__device__ __forceinline__ void ring_walk(const Ring& ring, int lane, uint32_t rowsT, uint32_t cnt, LoadIdx load_idx, Body body) {
  constexpr int IDX = SPHB200_RING_IDX_AHEAD;
  uint32_t ji[IDX];                               // entries of positions q .. q+IDX-1, q = next position to issue
#pragma unroll
  for (int s = 0; s < IDX; ++s) ji[s] = load_idx((uint32_t)s);
  uint32_t jq[STAGES];                            // entries whose copies are in flight (the body gets its own entry back)
#pragma unroll
  for (int s = 0; s < STAGES; ++s) jq[s] = 0u;
#pragma unroll
  for (uint32_t p = 0; p < (uint32_t)STAGES - 1u; ++p) {
    ring.issue(p, ji[0], lane);
    jq[p] = ji[0];
#pragma unroll
    for (int s = 0; s < IDX - 1; ++s) ji[s] = ji[s + 1];
    ji[IDX - 1] = load_idx(p + (uint32_t)IDX);
  }
  for (uint32_t k = 0; k < rowsT; ++k) {
    ring_wait<STAGES - 2>();                      // this lane's copies for position k have landed ...
    __syncwarp();                                 // ... and so have every other lane's; everyone is past its reads of stage k-1
    const uint32_t j = jq[0];
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) jq[s] = jq[s + 1];
    // refill the stage consumed in the previous iteration (it aliases position k+STAGES-1; every lane is past those reads)
    ring.issue(k + (uint32_t)STAGES - 1u, ji[0], lane);
    jq[STAGES - 2] = ji[0];
#pragma unroll
    for (int s = 0; s < IDX - 1; ++s) ji[s] = ji[s + 1];
    ji[IDX - 1] = load_idx(k + (uint32_t)(STAGES - 1 + IDX));
    if (k < cnt) body(k, j);
  }
  ring_wait<0>();
#if SPHB200_RING_ROTATE
  // The next walk of this warp (next tile, or the next pass over the same tile) must not refill a stage a slower lane still reads.
  // Every read but those of the last iteration precedes a __syncwarp all lanes have passed; the last iteration read stage
  // (rot + rowsT - 1) % STAGES.  Advancing rot by rowsT makes that the one stage the next prologue (positions 0 .. STAGES-2) leaves
  // alone, and its refill in iteration 0 of the next walk comes after that iteration's __syncwarp.  A __syncwarp here does the same
  // at +3 % kernel time (measured on k_sph_derivs, profiles/r02_notes.md 11e).
  ring.rot = (ring.rot + rowsT) % (unsigned)STAGES;
#else
  __syncwarp();
#endif
}

}  // namespace
