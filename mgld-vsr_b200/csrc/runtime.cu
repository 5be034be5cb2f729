// Library state: error string, driver entry points, device properties.
#include <stdlib.h>
#include <string.h>

#include "../../include/mgld.h"
#include "common.h"

namespace mgld {

static thread_local char g_err[512] = "";
static EncodeTiledFn g_encode = nullptr;
static int g_num_sms = 0;
static bool g_init = false;
static int g_device = -1;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return MGLD_ERR_CUDA;
}
EncodeTiledFn encode_tiled_fn() { return g_encode; }
int num_sms() { return g_num_sms; }
bool initialised() { return g_init; }
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("MGLD_PDL"); return !e || atoi(e) != 0; }();
  return on;
}

int make_tmap_f16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides) {
  if (!g_encode) {
    set_error("mgld_init() has not been called");
    return MGLD_ERR_NOT_INIT;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides ? elem_strides[i] : 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  const uint32_t inner_bytes = box[0] * 2;
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  if (inner_bytes == 64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  else if (inner_bytes == 32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  else if (inner_bytes != 128) {
    set_error("make_tmap_f16: inner box of %u bytes is not a swizzle width", inner_bytes);
    return MGLD_ERR_ARG;
  }
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                        es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] stride0 %llu",
              (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
              (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0), bx[0],
              rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0,
              (unsigned long long)(rank > 1 ? gstr[0] : 0));
    return MGLD_ERR_CUDA;
  }
  return MGLD_OK;
}

int make_tmap_f32(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  if (!g_encode) {
    set_error("mgld_init() has not been called");
    return MGLD_ERR_NOT_INIT;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  if (box[0] != 32) { set_error("make_tmap_f32: inner box must be 32 elements"); return MGLD_ERR_ARG; }
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(f32) failed (%d)", (int)r); return MGLD_ERR_CUDA; }
  return MGLD_OK;
}

}  // namespace mgld

using namespace mgld;

extern "C" int mgld_abi_version(void) { return MGLD_ABI_VERSION; }

extern "C" const char* mgld_last_error(void) { return g_err; }

extern "C" int mgld_init(int device) {
  // One process drives one GPU (one rank per GPU under torchrun): the shared-memory opt-ins, the co-resident cluster count
  // and the SM count are cached per process, so a second device in the same process is refused instead of silently
  // running with the first device's capabilities.
  if (g_init && device != g_device) {
    set_error("mgld_init(%d): this process is already bound to device %d (libmgld is one-process-per-GPU)", device, g_device);
    return MGLD_ERR_ARG;
  }
  if (g_init) return MGLD_OK;
  MGLD_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MGLD_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error("mgld requires an sm_100a device (found sm_%d%d)", prop.major, prop.minor);
    return MGLD_ERR_ARG;
  }
  g_num_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  MGLD_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point not found");
    return MGLD_ERR_CUDA;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_device = device;
  g_init = true;
  return MGLD_OK;
}
