// Microbenchmark: how fast can TMA fill shared memory from L2 / HBM on B200, per SM and in aggregate?
// Same producer/consumer ring as conv_gemm (no MMA: the consumer releases a stage as soon as it is full).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_fill tma_fill.cu -lcuda ; ./tma_fill
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include "../../mgld-vsr_b200/csrc/ptx.cuh"
using namespace mgld;

struct P { int stages, stage_bytes, box_rows, iters, rows_total, mode; };
// mode 0: each CTA streams its own disjoint region (HBM if > L2); mode 1: all CTAs stream the same 8 MB region (L2 hits)
__global__ void __launch_bounds__(64, 1) fill_kernel(const __grid_constant__ CUtensorMap tm, const P p, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { for (int s = 0; s < p.stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); } fence_mbar_init(); }
  __syncthreads();
  const int boxes_per_stage = p.stage_bytes / (p.box_rows * 128);
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    const int region_rows = p.mode == 0 ? p.rows_total / gridDim.x : 65536;
    const int row0 = p.mode == 0 ? blockIdx.x * region_rows : 0;
    for (int it = 0; it < p.iters; ++it) {
      const int s = it % p.stages;
      mbar_wait(smem_u32(&empty_bar[s]), ((it / p.stages) & 1) ^ 1);
      mbar_expect_tx(smem_u32(&full_bar[s]), p.stage_bytes);
      for (int b = 0; b < boxes_per_stage; ++b) {
        const int r = row0 + ((it * boxes_per_stage + b) * p.box_rows + (p.mode ? blockIdx.x * 4096 : 0)) % region_rows;
        tma_load_2d(base + s * p.stage_bytes + b * p.box_rows * 128, &tm, smem_u32(&full_bar[s]), 0, r);
      }
    }
  } else if (threadIdx.x == 32) {
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
  const long long rows = 32ll << 20;  // 32M rows x 128 B = 4 GB
  void* buf; cudaMalloc(&buf, rows * 128); cudaMemset(buf, 0, rows * 128);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  int sm_clock; cudaDeviceGetAttribute(&sm_clock, cudaDevAttrClockRate, 0);
  printf("mode grid stages stageKB boxrows | us  aggregate GB/s  per-SM B/clk(clock64)\n");
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {8, 30, 74, 148})
      for (int cfg = 0; cfg < 4; ++cfg) {
        P p; p.mode = mode; p.rows_total = (int)rows;
        const int stagesv[4] = {4, 6, 3, 12}; const int stageb[4] = {32768, 32768, 65536, 16384}; const int boxr[4] = {128, 128, 256, 128};
        p.stages = stagesv[cfg]; p.stage_bytes = stageb[cfg]; p.box_rows = boxr[cfg]; p.iters = 2000 * 32768 / p.stage_bytes;
        CUtensorMap tm; cuuint64_t dims[2] = {64, (cuuint64_t)rows}; cuuint64_t str[1] = {128}; cuuint32_t box[2] = {64, (cuuint32_t)p.box_rows}; cuuint32_t es[2] = {1, 1};
        enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        fill_kernel<<<grid, 64, p.stages * p.stage_bytes + 1024>>>(tm, p, cyc); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        fill_kernel<<<grid, 64, p.stages * p.stage_bytes + 1024>>>(tm, p, cyc);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<unsigned long long> h(grid); cudaMemcpy(h.data(), cyc, grid * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto c : h) avg += c; avg /= grid;
        const double bytes = (double)grid * p.iters * p.stage_bytes;
        printf("%d %4d %2d %3d %3d | %8.1f %9.1f %7.1f   %s\n", mode, grid, p.stages, p.stage_bytes / 1024, p.box_rows, ms * 1e3, bytes / ms / 1e6,
               (double)p.iters * p.stage_bytes / avg, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
