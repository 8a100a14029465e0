// ubench_tex_lds.cu -- do texture fetches (TEX pipe) and shared-memory loads (LSU pipe) of the L1TEX unit run side by side?
//
// The pair loops are bound by LSU data-pipe wavefronts (profiles/r01_notes.md); ~30 % of them are the TableKernel look-ups
// (random 48-byte records, bank conflicts).  If the TEX path has its own throughput, moving the look-ups there relieves the LSU.
// Each warp runs ITER iterations of  NL random LDS.128 (24 KB table in shared memory)  +  NT random tex1Dfetch<int4> (same table
// as a linear texture), at the occupancy of k_sph_derivs (8 warps per SM).  Prints ns per warp-iteration for the three mixes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tex_lds scripts/ubench_tex_lds.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int RECS = 502, ITER = 2000;

template <int NL, int NT, int DEP>
__global__ void __launch_bounds__(128, 2) k(cudaTextureObject_t tex, const int4* __restrict__ tab, double* out) {
  extern __shared__ int4 s[];
  for (int i = threadIdx.x; i < RECS*3; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  unsigned x = threadIdx.x*2654435761u + blockIdx.x*40503u + 12345u;
  double acc = 0.0;
  for (int it = 0; it < ITER; ++it) {
    int4 a[NL > 0 ? NL : 1], b[NT > 0 ? NT : 1];
#pragma unroll
    for (int q = 0; q < NL; ++q) { x = x*1664525u + 1013904223u; const unsigned k = ((x >> 8) % RECS)*3u + (q % 3); a[q] = s[k]; }
#pragma unroll
    for (int q = 0; q < NT; ++q) { x = x*1664525u + 1013904223u; const unsigned k = ((x >> 8) % RECS)*3u + (q % 3); b[q] = tex1Dfetch<int4>(tex, (int)k); }
#pragma unroll
    for (int q = 0; q < NL; ++q) acc += __hiloint2double(a[q].y, a[q].x);
#pragma unroll
    for (int q = 0; q < NT; ++q) acc += __hiloint2double(b[q].y, b[q].x);
    if (DEP) x += (unsigned)__double2int_rz(acc) & 1u;            // next indices depend on this iteration's data (latency exposed)
  }
  out[blockIdx.x*blockDim.x + threadIdx.x] = acc;
}

template <int NL, int NT, int DEP>
void run(const char* name, cudaTextureObject_t tex, const int4* tab, double* out, int nsm) {
  const size_t shm = 100*1024;                                    // 2 CTAs per SM, like the pair kernel
  cudaFuncSetAttribute(k<NL, NT, DEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<NL, NT, DEP><<<2*nsm, 128, shm>>>(tex, tab, out);
  cudaEventRecord(e0);
  k<NL, NT, DEP><<<2*nsm, 128, shm>>>(tex, tab, out);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  // 8 warps per SM share the SM: time per (SM, iteration of all 8 warps)
  printf("%-28s %8.3f ms   %7.1f ns per warp-iteration (8 warps/SM -> %6.1f cycles/SM per 8 iterations at 1.965 GHz)  err=%s\n",
         name, ms, ms*1e6/ITER/8.0, ms*1e6/ITER*1.965, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int nsm = 148; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  int4* tab; cudaMalloc(&tab, RECS*3*sizeof(int4));
  int4* h = (int4*)malloc(RECS*3*sizeof(int4));
  for (int i = 0; i < RECS*3; ++i) h[i] = make_int4(i, 0x3ff00000, i, 0x3ff00000);
  cudaMemcpy(tab, h, RECS*3*sizeof(int4), cudaMemcpyHostToDevice);
  cudaResourceDesc rd{}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
  rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = RECS*3*sizeof(int4);
  cudaTextureDesc td{}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; cudaCreateTextureObject(&tex, &rd, &td, nullptr);
  double* out; cudaMalloc(&out, 2*nsm*128*sizeof(double));
  run<24, 0, 0>("24 LDS.128", tex, tab, out, nsm);
  run<18, 0, 0>("18 LDS.128", tex, tab, out, nsm);
  run<0, 6, 0>("6 TEX", tex, tab, out, nsm);
  run<18, 6, 0>("18 LDS.128 + 6 TEX", tex, tab, out, nsm);
  run<24, 6, 0>("24 LDS.128 + 6 TEX", tex, tab, out, nsm);
  run<0, 12, 0>("12 TEX", tex, tab, out, nsm);
  run<18, 0, 1>("18 LDS.128 dependent", tex, tab, out, nsm);
  run<0, 6, 1>("6 TEX dependent", tex, tab, out, nsm);
  run<18, 6, 1>("18 LDS.128 + 6 TEX dependent", tex, tab, out, nsm);
  return 0;
}
