// ubench_dfma.cu -- FP64 pipe on B200: dependent-issue latency of DFMA and the throughput reached with W warps per scheduler x C
// independent chains per thread (k_sph_derivs runs 2 warps per scheduler; how much ILP does it need?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_dfma scripts/ubench_dfma.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int C>
__global__ void k(double* out, int iters, double a, double b, long long* cyc) {
  double x[C];
#pragma unroll
  for (int c = 0; c < C; ++c) x[c] = 1.0 + 1e-9*(threadIdx.x + c);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < C; ++c) x[c] = fma(x[c], a, b);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) s += x[c];
  out[blockIdx.x*blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int C> void run(int warpsPerSM, double* out, long long* cyc, int nsm) {
  const int iters = 4096;
  k<C><<<nsm, 32*warpsPerSM>>>(out, iters, 0.999999, 1e-7, cyc);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<C><<<nsm, 32*warpsPerSM>>>(out, iters, 0.999999, 1e-7, cyc);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double perInstr = (double)h/((double)iters*C);
  printf("warps/SM %2d (%.1f per scheduler)  chains %2d : %6.2f cycles per DFMA per warp, %5.2f warp-DFMA/clk/SM (peak 2), %5.1f TFLOP/s\n", warpsPerSM, warpsPerSM/4.0, C,
         perInstr, warpsPerSM/perInstr, 2.0*32*warpsPerSM*(double)iters*C*nsm/(ms*1e-3)/1e12);
}
int main() {
  int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, (size_t)nsm*1024*8); long long* cyc; cudaMalloc(&cyc, 8);
  run<1>(1, out, cyc, nsm); run<2>(1, out, cyc, nsm); run<4>(1, out, cyc, nsm); run<8>(1, out, cyc, nsm); run<16>(1, out, cyc, nsm);
  run<1>(4, out, cyc, nsm); run<2>(4, out, cyc, nsm); run<4>(4, out, cyc, nsm); run<8>(4, out, cyc, nsm);
  run<1>(8, out, cyc, nsm); run<2>(8, out, cyc, nsm); run<4>(8, out, cyc, nsm); run<8>(8, out, cyc, nsm); run<16>(8, out, cyc, nsm);
  run<1>(12, out, cyc, nsm); run<2>(12, out, cyc, nsm); run<4>(12, out, cyc, nsm); run<8>(12, out, cyc, nsm);
  run<1>(16, out, cyc, nsm); run<2>(16, out, cyc, nsm); run<4>(16, out, cyc, nsm); run<8>(16, out, cyc, nsm);
  run<4>(32, out, cyc, nsm); run<8>(32, out, cyc, nsm);
  return 0;
}
