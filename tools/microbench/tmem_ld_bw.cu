// Microbenchmark: sustained tcgen05.ld (TMEM -> registers) and tcgen05.st throughput of one SM, as a function of the number
// of warps issuing (4 = one per sub-partition, 8 = two per sub-partition).  Question behind it: is the d=64
// attention kernel bound by reading its fp32 scores out of TMEM (64 KB per 128x128 block)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_bw tmem_ld_bw.cu && ./tmem_ld_bw
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../mgld-vsr_b200/csrc/ptx.cuh"
using namespace mgld;

template <int kMode>   // 0: ld x32 + wait each;  1: 4 x ld x32 back to back, one wait;  2: st x32 + wait each
__global__ void __launch_bounds__(256, 1) k(int iters, unsigned long long* out, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  uint32_t r[32], acc = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  tmem_st_x32(base, r); tmem_st_x32(base + 32, r); tmem_st_x32(base + 64, r); tmem_st_x32(base + 96, r);
  tmem_st_wait();
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (kMode == 0) {
      tmem_ld_x32(base + (it & 3) * 32, r);
      tmem_ld_wait();
      acc += r[0] ^ r[13] ^ r[31];
    } else if (kMode == 1) {
      uint32_t a[32], b[32], c[32];
      tmem_ld_x32(base, r); tmem_ld_x32(base + 32, a); tmem_ld_x32(base + 64, b); tmem_ld_x32(base + 96, c);
      tmem_ld_wait();
      acc += (r[0] ^ a[7]) + (b[19] ^ c[31]) + (r[31] ^ a[0] ^ b[0] ^ c[0]);
    } else {
      r[0] += it; r[17] ^= it;
      tmem_st_x32(base + (it & 3) * 32, r);
      tmem_st_wait();
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[0] = t1 - t0;
  sink[threadIdx.x] = acc + r[0];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  unsigned long long* d; uint32_t* sink;
  cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<1, threads>>>(iters, d, sink);
        if (mode == 1) k<1><<<1, threads>>>(iters, d, sink);
        if (mode == 2) k<2><<<1, threads>>>(iters, d, sink);
        cudaDeviceSynchronize();
      }
      unsigned long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
      const double per_it_bytes = (mode == 1 ? 4.0 : 1.0) * threads * 32 * 4;   // bytes moved per iteration by the CTA
      printf("mode %d (%s) warps %d: %.1f cycles/iter, %.1f B/clk/SM  [%s]\n", mode,
             mode == 0 ? "ld.x32 + wait" : mode == 1 ? "4 x ld.x32, one wait" : "st.x32 + wait", threads / 32, (double)c / iters,
             per_it_bytes * iters / (double)c, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
