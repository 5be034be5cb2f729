// Host-side helpers shared by the C-ABI translation units.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

namespace mgld {

enum : int {
  MGLD_OK = 0,
  MGLD_ERR_ARG = -1,      // invalid argument / unsupported shape
  MGLD_ERR_CUDA = -2,     // a CUDA runtime / driver call failed
  MGLD_ERR_NOT_INIT = -3  // mgld_init() not called
};

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define MGLD_CHECK_ARG(cond, ...)     \
  do {                                \
    if (!(cond)) {                    \
      ::mgld::set_error(__VA_ARGS__); \
      return ::mgld::MGLD_ERR_ARG;    \
    }                                 \
  } while (0)

#define MGLD_CUDA(call)                                                \
  do {                                                                 \
    cudaError_t e__ = (call);                                          \
    if (e__ != cudaSuccess) return ::mgld::cuda_fail(e__, #call);      \
  } while (0)

#define MGLD_LAUNCH_CHECK(name)                                        \
  do {                                                                 \
    cudaError_t e__ = cudaGetLastError();                              \
    if (e__ != cudaSuccess) return ::mgld::cuda_fail(e__, name);       \
  } while (0)

// cuTensorMapEncodeTiled obtained through cudaGetDriverEntryPoint (no link-time libcuda dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();
int num_sms();
bool initialised();

// fp16 tensor map of rank `rank` (dims fastest-first, strides in BYTES for dims 1..rank-1), 128B swizzle unless
// the inner box is narrower (64B -> SWIZZLE_64B, 32B -> SWIZZLE_32B).
int make_tmap_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides = nullptr);
// same for fp32 tensors (inner box must be 32 elements = 128 bytes)
int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (PDL).  One DDPM tile-step is ~800 back-to-back kernels of 5-100 us in one stream / CUDA
// graph; with ordinary stream order each boundary costs the launch latency plus the next kernel's prologue (block
// scheduling, barrier init, TMEM allocation, tensor-map fetch) on an idle machine.  Kernels launched through launch_pdl()
// may start while their predecessor drains: every such kernel calls pdl_launch_dependents() at its top (lets ITS successor
// start early) and pdl_wait() before its first access to global memory (returns once all preceding grids have completed
// and their writes are visible, so data hazards are exactly those of ordinary stream order).  MGLD_PDL=0 switches the
// launch attribute off (the device instructions are then no-ops).
bool pdl_enabled();
#ifdef __CUDACC__
// Order-independent (bitwise repeatable) accumulation of fp32 partial sums across thread blocks: a value is split exactly
// into its integer part and a 40-bit fixed-point fraction, each accumulated with 64-bit INTEGER atomics (associative and
// commutative, unlike floating-point atomics).  One accumulator = two 8-byte words {integer part, fraction * 2^40}; zero
// bytes = zero.  Range: |sum of integer parts| < 2^63; resolution 2^-40.  Used for the GroupNorm / InstanceNorm sums.
__device__ __forceinline__ void fixsum_add(double* acc, float p) {
  const float fl = floorf(p);
  const long long hi = __float2ll_rd(p);
  const long long lo = __double2ll_rd((double)(p - fl) * 1099511627776.0);
  unsigned long long* a = reinterpret_cast<unsigned long long*>(acc);
  atomicAdd(a, (unsigned long long)hi);
  atomicAdd(a + 1, (unsigned long long)lo);
}
__device__ __forceinline__ double fixsum_load(const double* acc) {
  const long long* a = reinterpret_cast<const long long*>(acc);
  return (double)a[0] + (double)a[1] * (1.0 / 1099511627776.0);
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... P, typename... A>
static inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                     A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}
#endif

}  // namespace mgld
