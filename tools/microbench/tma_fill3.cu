// Microbenchmark 3: is the per-SM TMA fill bound an L2->SM bandwidth cap for DISTINCT data, and does cluster multicast lift it?
//   mode 0: every CTA streams its own private region (distinct data, L2 resident after the first pass)
//   mode 1: all CTAs stream the same region (shared data)
//   mode 2: clusters of 2, each CTA loads half of every box and multicasts it to both (distinct per cluster)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "../../mgld-vsr_b200/csrc/ptx.cuh"
using namespace mgld;
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
               ::"r"(dst), "l"(tm), "r"(bar), "r"(c0), "r"(c1), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() { asm volatile("barrier.cluster.arrive.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
struct P { int stages, box_rows, iters, mode, region_rows; };
__global__ void __launch_bounds__(192, 1) fill_kernel(const __grid_constant__ CUtensorMap tm, const __grid_constant__ CUtensorMap tm_half, const P p, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int op_bytes = p.box_rows * 128;
  const int csize = p.mode == 2 ? 2 : 1;
  const uint32_t rank = p.mode == 2 ? cluster_rank() : 0;
  if (threadIdx.x == 0) { for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), csize); } fence_mbar_init(); }
  __syncthreads();
  if (csize == 2) cluster_sync_all();
  long long t0 = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long region0 = p.mode == 1 ? 0 : (long long)(blockIdx.x / csize) * p.region_rows;
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(smem_u32(&empty_bar[s]), ((it / p.stages) & 1) ^ 1);
      mbar_expect_tx(smem_u32(&full_bar[s]), op_bytes);
      const long long r = region0 + ((long long)it * p.box_rows) % p.region_rows;
      if (csize == 1) tma_load_2d(base + s * op_bytes, &tm, smem_u32(&full_bar[s]), 0, (int)r);
      else tma_load_2d_mc(base + s * op_bytes + rank * (op_bytes / 2), &tm_half, smem_u32(&full_bar[s]), 0, (int)(r + rank * (p.box_rows / 2)), 3);
    }
  } else if (warp == 5 && lane == 0) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(smem_u32(&full_bar[s]), (it / p.stages) & 1);
      if (csize == 1) mbar_arrive(smem_u32(&empty_bar[s]));
      else { mbar_arrive_remote(smem_u32(&empty_bar[s]), 0); mbar_arrive_remote(smem_u32(&empty_bar[s]), 1); }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  if (csize == 2) { __syncthreads(); cluster_sync_all(); }
}
int main() {
  cudaSetDevice(0);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  const long long rows = 1ll << 21;   // 256 MB
  uint8_t* buf; cudaMalloc(&buf, rows * 128); cudaMemset(buf, 0, rows * 128);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("mode boxrows stages region_KB | per-SM B/clk (smem fill)  agg TB/s@1.9GHz  L2-read B/clk/SM\n");
  struct C { int mode, box_rows, stages, region_rows; };
  std::vector<C> cs;
  for (int br : {128, 256}) for (int rr : {2048, 4096}) for (int mode : {0, 1, 2}) cs.push_back({mode, br, 6, rr});
  cs.push_back({0, 128, 6, 14000}); cs.push_back({2, 128, 6, 14000});   // ~265 MB > L2: DRAM streaming
  for (auto c : cs) {
    P p; p.mode = c.mode; p.box_rows = c.box_rows; p.stages = c.stages; p.iters = 6000; p.region_rows = c.region_rows;
    if ((long long)148 * c.region_rows > rows) p.region_rows = (int)(rows / 148 / 256 * 256);
    CUtensorMap tm, tmh; cuuint64_t dims[2] = {64, (cuuint64_t)rows}; cuuint64_t str[1] = {128}; cuuint32_t es[2] = {1, 1};
    cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows}, boxh[2] = {64, (cuuint32_t)c.box_rows / 2};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&tmh, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, boxh, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int smem = p.stages * c.box_rows * 128 + 1024;
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = c.mode == 2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; ++rep) { cudaLaunchKernelEx(&cfg, fill_kernel, tm, tmh, p, cyc); cudaDeviceSynchronize(); }
    std::vector<unsigned long long> h(148); cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto x : h) avg += x; avg /= 148;
    const double bpc = (double)p.iters * c.box_rows * 128 / avg;
    printf("%d %4d %d %6d | %7.1f  %6.2f  %7.1f  %s\n", c.mode, c.box_rows, c.stages, p.region_rows * 128 / 1024, bpc, bpc * 148 * 1.9e9 / 1e12,
           c.mode == 2 ? bpc / 2 : bpc, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
