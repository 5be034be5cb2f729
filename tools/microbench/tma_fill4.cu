// Microbenchmark 4: separate TMA throughput from thread-to-thread mbarrier signalling.
//   mode 0: ONE thread issues and waits on its own full barriers (depth D in flight, re-issue on completion)
//   mode 1: producer thread + consumer thread (full/empty barriers), as in the conv pipeline
//   mode 2: as mode 0 with 1-D bulk copies (no tensor map)
//   mode 3: as mode 1, consumer = whole warp polling (lane 0 arrives)
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
struct P { int depth_log2, box_rows, iters, mode, region_rows; };
__global__ void __launch_bounds__(192, 1) fill_kernel(const __grid_constant__ CUtensorMap tm, const uint8_t* buf, const P p, unsigned long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int op_bytes = p.box_rows * 128;
  const int D = 1 << p.depth_log2, mask = D - 1;
  if (threadIdx.x == 0) { for (int s = 0; s < D; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); } fence_mbar_init(); }
  __syncthreads();
  long long t0 = clock64();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int region0 = blockIdx.x * p.region_rows;
  const int rmask = p.region_rows - 1;   // power of two
  if (p.mode == 0 || p.mode == 2) {
    if (threadIdx.x == 0) {
      for (int it = 0; it < p.iters + D; ++it) {
        const int s = it & mask;
        if (it >= D) mbar_wait(smem_u32(&full_bar[s]), ((it >> p.depth_log2) & 1) ^ 1);
        if (it < p.iters) {
          mbar_expect_tx(smem_u32(&full_bar[s]), op_bytes);
          const int r = region0 + ((it * p.box_rows) & rmask);
          if (p.mode == 2) bulk_load_1d(base + s * op_bytes, buf + (long long)r * 128, op_bytes, smem_u32(&full_bar[s]));
          else tma_load_2d(base + s * op_bytes, &tm, smem_u32(&full_bar[s]), 0, r);
        }
      }
      cycles[blockIdx.x] = clock64() - t0;
    }
    return;
  }
  if (warp == 0 && lane == 0) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it & mask;
      mbar_wait(smem_u32(&empty_bar[s]), ((it >> p.depth_log2) & 1) ^ 1);
      mbar_expect_tx(smem_u32(&full_bar[s]), op_bytes);
      const int r = region0 + ((it * p.box_rows) & rmask);
      tma_load_2d(base + s * op_bytes, &tm, smem_u32(&full_bar[s]), 0, r);
    }
  } else if (warp == 5 && (lane == 0 || p.mode == 3)) {
    for (int it = 0; it < p.iters; ++it) {
      const int s = it & mask;
      mbar_wait(smem_u32(&full_bar[s]), (it >> p.depth_log2) & 1);
      if (p.mode == 3) __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty_bar[s]));
    }
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
  }
}
int main() {
  cudaSetDevice(0);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  const long long rows = 1ll << 21;
  uint8_t* buf; cudaMalloc(&buf, rows * 128); cudaMemset(buf, 0, rows * 128);
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  printf("mode grid boxrows depth | cycles/op  per-SM B/clk  agg TB/s@1.9GHz\n");
  struct C { int mode, grid, box_rows, dl2; };
  std::vector<C> cs;
  for (int grid : {1, 148}) {
    for (int dl2 : {0, 1, 2, 3}) cs.push_back({0, grid, 128, dl2});
    cs.push_back({0, grid, 64, 3}); cs.push_back({0, grid, 64, 4}); cs.push_back({0, grid, 256, 2}); cs.push_back({0, grid, 32, 4});
    for (int dl2 : {1, 2, 3}) cs.push_back({1, grid, 128, dl2});
    cs.push_back({1, grid, 64, 4});
    cs.push_back({3, grid, 128, 3});
    cs.push_back({2, grid, 128, 2}); cs.push_back({2, grid, 128, 3});
  }
  for (auto c : cs) {
    P p; p.mode = c.mode; p.box_rows = c.box_rows; p.depth_log2 = c.dl2; p.iters = 4096; p.region_rows = 4096;
    CUtensorMap tm; cuuint64_t dims[2] = {64, (cuuint64_t)rows}; cuuint64_t str[1] = {128}; cuuint32_t es[2] = {1, 1};
    cuuint32_t box[2] = {64, (cuuint32_t)c.box_rows};
    enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int smem = (1 << c.dl2) * c.box_rows * 128 + 1024;
    for (int rep = 0; rep < 2; ++rep) { fill_kernel<<<c.grid, 192, smem>>>(tm, buf, p, cyc); cudaDeviceSynchronize(); }
    std::vector<unsigned long long> h(c.grid); cudaMemcpy(h.data(), cyc, c.grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (auto x : h) avg += x; avg /= c.grid;
    const double bpc = (double)p.iters * c.box_rows * 128 / avg;
    printf("%d %3d %4d %2d | %8.1f  %7.1f  %6.2f  %s\n", c.mode, c.grid, c.box_rows, 1 << c.dl2, avg / p.iters, bpc, bpc * c.grid * 1.9e9 / 1e12, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
