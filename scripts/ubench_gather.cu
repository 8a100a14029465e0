// ubench_gather.cu -- how should the pair loop fetch its neighbour rows on B200?
//
// k_sph_derivs is an i-centric gather: per warp iteration the 32 lanes need the 128-byte rows of 32 different nodes.  The round-1
// kernel streams them with warp-cooperative LDGSTS (cp.async) through the L1 into a per-warp shared-memory ring and is bound by
// LSU data-pipe wavefronts (profiles/r01_notes.md).  This microbenchmark times the same access pattern (Morton-local random rows,
// ~5x reuse inside a tile, 8 warps per SM, optional dependent FP64 work per row) for three fetch paths:
//   0  LDGSTS  : 8 lanes fetch the 8 chunks of one row, 8 instructions per iteration (the round-1 scheme)
//   1  GATHER4 : cp.async.bulk.tensor.2d.tile::gather4 -- lanes 0..7 each ask the TMA unit for 4 rows (SWIZZLE_128B destination)
//   2  BULK1D  : cp.async.bulk (1-D, 128 B) -- every lane asks the TMA unit for its own row (padded slot, no swizzle needed)
// The TMA paths bypass the LSU and the L1: they trade LSU wavefronts for L2 requests.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_gather scripts/ubench_gather.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

#ifndef UB_WARPS
#define UB_WARPS 4
#endif
#ifndef UB_STAGES
#define UB_STAGES 4
#endif
#ifndef UB_CTAS
#define UB_CTAS 2
#endif
constexpr int ROWD = 16, ROWB = 128, WARPS = UB_WARPS, STAGES = UB_STAGES, CTAS = UB_CTAS, ITERS = 96;

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src) { asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ double2 lds128v(unsigned a) { double2 v; asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void mbar_init(unsigned bar, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_gather4(unsigned dst, const CUtensorMap* tm, int c0, int r0, int r1, int r2, int r3, unsigned bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
               :: "r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk1d(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// idx[(tile*ITERS + it)*32 + lane]
template <int MODE, int NF, int MIX>
__global__ void __launch_bounds__(32*WARPS, CTAS) k_gather(const double* __restrict__ rows, const uint32_t* __restrict__ idx, int nTiles,
                                                        const __grid_constant__ CUtensorMap tmap, double* __restrict__ out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int SLOT = (MODE == 1) ? 128 : 144;                    // gather4 writes rows at 128 B stride (swizzled); others use a padded slot
  constexpr int STAGEB = (MODE == 1) ? 4096 : 32*SLOT;
  const unsigned base = ((unsigned)__cvta_generic_to_shared(smem) + 1023u) & ~1023u;      // SWIZZLE_128B destinations: 1024-byte aligned
  const unsigned ring = base + (unsigned)warp*(STAGES*STAGEB);
  const unsigned bars = base + WARPS*STAGES*STAGEB + (unsigned)warp*STAGES*8;
  if (MODE != 0) {
    if (lane == 0) for (int s = 0; s < STAGES; ++s) mbar_init(bars + 8*s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
  }
  double acc = 0.0;
  unsigned phaseBits = 0;                                              // parity of every stage's mbarrier (bit s); persists across tiles
  const unsigned char* rowsB = reinterpret_cast<const unsigned char*>(rows);
  for (int tile = blockIdx.x*WARPS + warp; tile < nTiles; tile += gridDim.x*WARPS) {
    const uint32_t* ti = idx + (size_t)tile*ITERS*32 + lane;
    auto issue = [&](int p, uint32_t j) {
      const int s = p % STAGES;
      const unsigned st = ring + s*STAGEB;
      if (MODE == 0) {
        const unsigned dst = st + (unsigned)(lane >> 3)*SLOT + 16u*(lane & 7);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t jr = __shfl_sync(0xffffffffu, j, 4*q + (lane >> 3));
          cp_async16(dst + q*(4*SLOT), rowsB + (size_t)jr*ROWB + 16*(lane & 7));
        }
        cp_commit();
      } else if (MODE == 1) {
        const int r0 = __shfl_sync(0xffffffffu, j, (4*lane) & 31), r1 = __shfl_sync(0xffffffffu, j, (4*lane + 1) & 31);
        const int r2 = __shfl_sync(0xffffffffu, j, (4*lane + 2) & 31), r3 = __shfl_sync(0xffffffffu, j, (4*lane + 3) & 31);
        if (lane == 0) mbar_expect(bars + 8*s, 32*ROWB);
        __syncwarp();
        if (lane < 8) tma_gather4(st + 512u*lane, &tmap, 0, r0, r1, r2, r3, bars + 8*s);
      } else {
        if (lane == 0) mbar_expect(bars + 8*s, 32*ROWB);
        __syncwarp();
        bulk1d(st + (unsigned)lane*SLOT, rowsB + (size_t)j*ROWB, ROWB, bars + 8*s);
      }
    };
    uint32_t jn = ti[0];
#pragma unroll
    if (MIX != 1) for (int p = 0; p < STAGES - 1; ++p) { issue(p, jn); jn = ti[(p + 1)*32]; }
    for (int k = 0; k < ITERS; ++k) {
      const int s = k % STAGES;
      if (MIX == 1) {} else if (MODE == 0) { cp_wait<STAGES - 2>(); __syncwarp(); }
      else { mbar_wait(bars + 8*s, (phaseBits >> s) & 1u); phaseBits ^= 1u << s; }
      double rw[ROWD];
      const unsigned rp = ring + s*STAGEB + (unsigned)lane*SLOT;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const unsigned off = (MODE == 1) ? 16u*(unsigned)(q ^ (lane & 7)) : 16u*q;     // SWIZZLE_128B: chunk ^= row & 7
        const double2 v = lds128v(rp + off); rw[2*q] = v.x; rw[2*q + 1] = v.y;
      }
      __syncwarp();
      if (MIX != 1) { const int p = k + STAGES - 1; if (p < ITERS + STAGES - 1) { issue(p, jn); jn = (p + 1 < ITERS) ? ti[(p + 1)*32] : 0u; } }
      // dependent FP64 work: NF DFMAs in 4 chains seeded by the row
      double c0 = rw[0] + rw[4], c1 = rw[1] + rw[5], c2 = rw[2] + rw[6], c3 = rw[3] + rw[7];
      if (MIX == 2) {                                                  // same instruction count on the FP32 pipe
        float f0 = (float)c0, f1 = (float)c1, f2 = (float)c2, f3 = (float)c3; const float a0 = (float)rw[8], b0 = (float)rw[12];
#pragma unroll
        for (int q = 0; q < NF/4; ++q) { f0 = fmaf(f0, a0, b0); f1 = fmaf(f1, a0, b0); f2 = fmaf(f2, a0, b0); f3 = fmaf(f3, a0, b0); }
        c0 = f0; c1 = f1; c2 = f2; c3 = f3;
      } else if (MIX == 4) {
        double d0 = c0 + 1.0, d1 = c1 + 1.0, d2 = c2 + 1.0, d3 = c3 + 1.0;
#pragma unroll
        for (int q = 0; q < NF/8; ++q) { c0 = fma(c0, rw[8], rw[12]); c1 = fma(c1, rw[9], rw[13]); c2 = fma(c2, rw[10], rw[14]); c3 = fma(c3, rw[11], rw[15]);
                                         d0 = fma(d0, rw[8], rw[12]); d1 = fma(d1, rw[9], rw[13]); d2 = fma(d2, rw[10], rw[14]); d3 = fma(d3, rw[11], rw[15]); }
        c0 += d0; c1 += d1; c2 += d2; c3 += d3;
      } else {
#pragma unroll
        for (int q = 0; q < NF/4; ++q) { c0 = fma(c0, rw[8], rw[12]); c1 = fma(c1, rw[9], rw[13]); c2 = fma(c2, rw[10], rw[14]); c3 = fma(c3, rw[11], rw[15]); }
      }
      acc += (c0 + c1) + (c2 + c3);
    }
    if (MODE == 0) cp_wait<0>();
    else {                                                             // drain the look-ahead copies of this tile
      for (int p = ITERS; p < ITERS + STAGES - 1; ++p) { const int s = p % STAGES; mbar_wait(bars + 8*s, (phaseBits >> s) & 1u); phaseBits ^= 1u << s; }
    }
    __syncwarp();
  }
  out[blockIdx.x*blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int NF, int MIX = 0>
double run(const char* name, const double* rows, const uint32_t* idx, int nTiles, const CUtensorMap& tm, double* out, int nsm, double* hsum) {
  const size_t shm = (size_t)WARPS*STAGES*((MODE == 1) ? 4096 : 32*144) + WARPS*STAGES*8 + 1024;
  cudaFuncSetAttribute(k_gather<MODE, NF, MIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = CTAS*nsm;
  k_gather<MODE, NF, MIX><<<grid, 32*WARPS, shm>>>(rows, idx, nTiles, tm, out);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k_gather<MODE, NF, MIX><<<grid, 32*WARPS, shm>>>(rows, idx, nTiles, tm, out);
  cudaEventRecord(e1);
  cudaError_t err = cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  std::vector<double> h((size_t)grid*32*WARPS);
  cudaMemcpy(h.data(), out, h.size()*8, cudaMemcpyDeviceToHost);
  double s = 0; for (double v : h) s += v;
  *hsum = s;
  const double iters = (double)nTiles*ITERS;
  printf("[%dx%d warps, %d stages] %-10s mix %d NF=%3d  %8.3f ms  %6.1f cycles/SM per warp-iteration  %6.1f G rows/s  %6.2f TB/s   sum=%.6e  %s\n", CTAS, WARPS, STAGES, name, MIX, NF, ms,
         ms*1e-3*1.965e9*nsm/iters, iters*32/ms*1e-6, iters*32*128/ms*1e-9, s, cudaGetErrorString(err));
  fflush(stdout);
  return ms;
}

int main(int argc, char** argv) {
  const size_t N = (argc > 1) ? (size_t)atol(argv[1]) : ((size_t)1 << 23);   // rows (8 Mi x 128 B = 1 GiB)
  const int W = (argc > 2) ? atoi(argv[2]) : 300;                              // half width of a tile's row window
  const int box1 = (argc > 3) ? atoi(argv[3]) : 1;                            // tensor-map box rows for gather4
  int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int nTiles = (int)(N/32);
  double* rows; cudaMalloc(&rows, N*ROWB);
  { std::vector<double> h(N*ROWD); for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 + 1e-3*(double)((i*2654435761ull) & 1023); cudaMemcpy(rows, h.data(), N*ROWB, cudaMemcpyHostToDevice); }
  uint32_t* idx; cudaMalloc(&idx, (size_t)nTiles*ITERS*32*4);
  { std::vector<uint32_t> h((size_t)nTiles*ITERS*32); uint64_t x = 88172645463325252ull;
    for (int t = 0; t < nTiles; ++t) for (int k = 0; k < ITERS*32; ++k) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      long long j = (long long)t*32 + (long long)(x % (uint64_t)(2*W)) - W; if (j < 0) j = 0; if (j >= (long long)N) j = (long long)N - 1;
      h[(size_t)t*ITERS*32 + k] = (uint32_t)j; }
    cudaMemcpy(idx, h.data(), h.size()*4, cudaMemcpyHostToDevice); }
  double* out; cudaMalloc(&out, (size_t)CTAS*nsm*32*WARPS*8);
  CUtensorMap tm{};
  { EncodeFn enc = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q);
    cuuint64_t dims[2] = {ROWD, (cuuint64_t)N}; cuuint64_t strides[1] = {ROWB}; cuuint32_t box[2] = {ROWD, (cuuint32_t)box1}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc ? enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, rows, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) : CUDA_ERROR_UNKNOWN;
    printf("tensor map encode: %d (box rows %d)  N=%zu rows  window +-%d  tiles %d  iterations/tile %d\n", (int)r, box1, N, W, nTiles, ITERS); }
  double s0, s1, s2;
  const int which = (argc > 4) ? atoi(argv[4]) : 7;
  if (which & 1) { run<0, 0>("LDGSTS", rows, idx, nTiles, tm, out, nsm, &s0); run<0, 96>("LDGSTS", rows, idx, nTiles, tm, out, nsm, &s0); run<0, 200>("LDGSTS", rows, idx, nTiles, tm, out, nsm, &s0); }
  if (which & 8) { run<0, 200, 1>("FP64only", rows, idx, nTiles, tm, out, nsm, &s0); run<0, 200, 2>("FFMA", rows, idx, nTiles, tm, out, nsm, &s0); run<0, 400, 2>("FFMA", rows, idx, nTiles, tm, out, nsm, &s0);
                   run<0, 200, 4>("8chains", rows, idx, nTiles, tm, out, nsm, &s0); run<0, 96, 4>("8chains", rows, idx, nTiles, tm, out, nsm, &s0); run<0, 96, 1>("FP64only", rows, idx, nTiles, tm, out, nsm, &s0); }
  if (which & 2) { run<1, 0>("GATHER4", rows, idx, nTiles, tm, out, nsm, &s1); run<1, 96>("GATHER4", rows, idx, nTiles, tm, out, nsm, &s1); run<1, 200>("GATHER4", rows, idx, nTiles, tm, out, nsm, &s1); }
  if (which & 4) { run<2, 0>("BULK1D", rows, idx, nTiles, tm, out, nsm, &s2); run<2, 96>("BULK1D", rows, idx, nTiles, tm, out, nsm, &s2); run<2, 200>("BULK1D", rows, idx, nTiles, tm, out, nsm, &s2); }
  return 0;
}
