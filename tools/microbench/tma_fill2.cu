// Microbenchmark 2: what limits TMA issue rate per SM?  Variants: number of issuing warps, box rows, tensor vs 1-D bulk.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "../../mgld-vsr_b200/csrc/ptx.cuh"
using namespace mgld;
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
struct P { int stages, ops_per_stage, op_bytes, box_rows, iters, producers, bulk1d; };
__global__ void __launch_bounds__(192, 1) fill_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* buf, const P p, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int stage_bytes = p.ops_per_stage * p.op_bytes;
  if (threadIdx.x == 0) { for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), p.producers); mbar_init(smem_u32(&empty_bar[s]), 1); } fence_mbar_init(); }
  __syncthreads();
  long long t0 = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < p.producers && lane == 0) {
    const int ops_mine = p.ops_per_stage / p.producers;
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(smem_u32(&empty_bar[s]), ((it / p.stages) & 1) ^ 1);
      mbar_expect_tx(smem_u32(&full_bar[s]), ops_mine * p.op_bytes);
      for (int b = 0; b < ops_mine; ++b) {
        const int op = warp * ops_mine + b;
        const long long r = ((long long)(it * p.ops_per_stage + op) * p.box_rows + blockIdx.x * 8192) % 65536;   // 8 MB window: L2 hits
        if (p.bulk1d) bulk_load_1d(base + s * stage_bytes + op * p.op_bytes, buf + r * 128, p.op_bytes, smem_u32(&full_bar[s]));
        else tma_load_2d(base + s * stage_bytes + op * p.op_bytes, &tm, smem_u32(&full_bar[s]), 0, (int)r);
      }
    }
  } else if (warp == 5 && lane == 0) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(smem_u32(&full_bar[s]), (it / p.stages) & 1);
      mbar_arrive(smem_u32(&empty_bar[s]));
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  cudaSetDevice(0);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  const long long rows = 1ll << 20;
  uint8_t* buf; cudaMalloc(&buf, rows * 128); cudaMemset(buf, 0, rows * 128);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("producers bulk1d boxrows ops/stage stages | per-SM B/clk   cycles/op   (grid 148, L2-resident)\n");
  struct C { int producers, bulk1d, box_rows, ops, stages; };
  std::vector<C> cs = {{1,0,128,2,4},{2,0,128,2,4},{1,0,64,4,4},{2,0,64,4,4},{4,0,64,4,4},{1,0,256,1,4},{1,0,256,2,3},{2,0,256,2,3},
                       {1,1,128,2,4},{2,1,128,2,4},{1,1,256,2,3},{2,1,256,2,3},{1,1,64,4,4},{1,0,32,8,4},{4,0,32,8,4},{1,1,512,1,3}};
  for (auto c : cs) {
    P p; p.producers = c.producers; p.bulk1d = c.bulk1d; p.box_rows = c.box_rows; p.ops_per_stage = c.ops; p.stages = c.stages;
    p.op_bytes = c.box_rows * 128; p.iters = 4000;
    CUtensorMap tm; cuuint64_t dims[2] = {64, (cuuint64_t)rows}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)(c.box_rows > 256 ? 256 : c.box_rows)}; cuuint32_t es[2] = {1, 1};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int smem = p.stages * p.ops_per_stage * p.op_bytes + 1024;
    fill_kernel<<<148, 192, smem>>>(tm, buf, p, cyc); cudaDeviceSynchronize();
    fill_kernel<<<148, 192, smem>>>(tm, buf, p, cyc); cudaDeviceSynchronize();
    std::vector<unsigned long long> h(148); cudaMemcpy(h.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto x : h) avg += x; avg /= 148;
    printf("%d %d %4d %d %d | %7.1f  %8.1f   %s\n", c.producers, c.bulk1d, c.box_rows, c.ops, c.stages, (double)p.iters * p.ops_per_stage * p.op_bytes / avg,
           avg / ((double)p.iters * p.ops_per_stage), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
