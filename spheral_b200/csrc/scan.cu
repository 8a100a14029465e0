// scan.cu -- hand-written device-wide exclusive scans (warp-shuffle block scan, 3-phase reduce-then-scan).
// Used for the cell table (counts -> cell starts) and the sliced-ELL tile offsets.  HBM-bound integer work:
// each element is read twice and written once.
#include "sphb200_internal.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;                       // per thread
constexpr int SCAN_BLOCK = SCAN_THREADS*SCAN_ITEMS; // 2048 elements per block

template <typename T> __device__ __forceinline__ T warp_incl_scan(T v) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) >= d) v += t;
  }
  return v;
}

// exclusive scan of the per-thread totals of a block; returns exclusive prefix, total via smem
template <typename T> __device__ __forceinline__ T block_excl_scan(T v, T* total) {
  __shared__ T warpSums[SCAN_THREADS/32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  T inc = warp_incl_scan(v);
  if (lane == 31) warpSums[w] = inc;
  __syncthreads();
  if (w == 0) {
    T s = (lane < SCAN_THREADS/32) ? warpSums[lane] : T(0);
    T si = warp_incl_scan(s);
    if (lane < SCAN_THREADS/32) warpSums[lane] = si - s;   // exclusive warp offsets
    if (lane == SCAN_THREADS/32 - 1) *total = si;
  }
  __syncthreads();
  return inc - v + warpSums[w];
}

template <typename TIn, typename TOut, int MUL>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const TIn* __restrict__ in, TOut* __restrict__ blockSums, size_t n) {
  __shared__ TOut total;
  const size_t base = (size_t)blockIdx.x*SCAN_BLOCK + (size_t)threadIdx.x*SCAN_ITEMS;
  TOut s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) if (base + k < n) s += (TOut)in[base + k]*MUL;
  block_excl_scan<TOut>(s, &total);
  if (threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

template <typename TIn, typename TOut, int MUL>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const TIn* in, TOut* out, const TOut* __restrict__ blockOffsets,
                                                             size_t nIn, size_t nOut) {
  __shared__ TOut total;
  const size_t base = (size_t)blockIdx.x*SCAN_BLOCK + (size_t)threadIdx.x*SCAN_ITEMS;
  TOut v[SCAN_ITEMS];
  TOut s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { v[k] = (base + k < nIn) ? (TOut)in[base + k]*MUL : TOut(0); s += v[k]; }
  TOut ex = block_excl_scan<TOut>(s, &total) + (blockOffsets ? blockOffsets[blockIdx.x] : TOut(0));
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { if (base + k < nOut) out[base + k] = ex; ex += v[k]; }
}

// out[i] = sum_{k<i} in[k]*MUL for i in [0,nOut), inputs beyond nIn count as 0 (nOut = nIn+1 also yields the total).
// in == out (same type) is allowed.
template <typename TIn, typename TOut, int MUL>
int scan_rec(sphb200_ctx* c, const TIn* in, TOut* out, size_t nIn, size_t nOut, char* tmp) {
  const size_t nb = (nOut + SCAN_BLOCK - 1)/SCAN_BLOCK;
  if (nb <= 1) {
    k_scan_apply<TIn, TOut, MUL><<<1, SCAN_THREADS, 0, c->stream>>>(in, out, nullptr, nIn, nOut);
    KERNEL_CHECK(c, "k_scan_apply");
    return 0;
  }
  TOut* sums = (TOut*)tmp;
  tmp += ((nb*sizeof(TOut) + 255)/256)*256;
  k_scan_reduce<TIn, TOut, MUL><<<(unsigned)nb, SCAN_THREADS, 0, c->stream>>>(in, sums, nIn);
  KERNEL_CHECK(c, "k_scan_reduce");
  if (int rc = scan_rec<TOut, TOut, 1>(c, sums, sums, nb, nb, tmp)) return rc;
  k_scan_apply<TIn, TOut, MUL><<<(unsigned)nb, SCAN_THREADS, 0, c->stream>>>(in, out, sums, nIn, nOut);
  KERNEL_CHECK(c, "k_scan_apply");
  return 0;
}

template <typename TIn, typename TOut, int MUL>
int scan_impl(sphb200_ctx* c, const TIn* in, TOut* out, size_t n) {
  size_t need = 4096, lv = n + 1;
  while (lv > 1) { lv = (lv + SCAN_BLOCK - 1)/SCAN_BLOCK; need += ((lv*sizeof(TOut) + 255)/256)*256 + 256; }
  if (need > c->scanTmpBytes) {
    if (c->scanTmp) cudaFree(c->scanTmp);
    c->scanTmp = nullptr; c->scanTmpBytes = 0;
    CU_CHECK(c, cudaMalloc(&c->scanTmp, need*2));
    c->scanTmpBytes = need*2;
  }
  return scan_rec<TIn, TOut, MUL>(c, in, out, n, n + 1, (char*)c->scanTmp);
}

}  // namespace

int sphb200_scan_u32(sphb200_ctx* c, const uint32_t* in, uint32_t* out, size_t n) {
  return scan_impl<uint32_t, uint32_t, 1>(c, in, out, n);
}
int sphb200_scan_tiles(sphb200_ctx* c, const uint32_t* rows, unsigned long long* out, size_t n) {
  return scan_impl<uint32_t, unsigned long long, SPHB200_TILE>(c, rows, out, n);
}
