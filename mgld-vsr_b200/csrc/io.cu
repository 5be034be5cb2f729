// I/O edges of the path on the GPU: what the reference does per frame on the host with numpy / PIL /
// F.interpolate on a CPU tensor, as two fused HBM-bound kernels on device-resident uint8 frames.
//
//   mgld_frames_u8_to_f32_bicubic   uint8 HWC frame -> (x/255 - 0.5)/0.5 (read_image, script :124-130) -> bicubic resize to
//                                   (oh, ow) (F.interpolate(mode='bicubic'), script :349-357) -> [optional] clamp(-1, 1)
//                                   (:376) -> [optional] reflect pad to (oh + pad_h, ow + pad_w) (:383-387), NCHW fp32
//   mgld_frames_f32_to_u8_hwc       [0,1] NCHW fp32 -> x * 255 -> HWC -> crop -> truncate to uint8 (script :529-541:
//                                   `im_sr.cpu().numpy().transpose(0,2,3,1) * 255`, `[:, :ori_h, :ori_w]`, `.astype(np.uint8)`)
//
// The bicubic follows ATen's upsample_bicubic2d exactly (align_corners=False, scale = in/out, A = -0.75, indices clamped,
// four horizontal cubic interpolations then one vertical, fp32), so the result equals F.interpolate to fp32 rounding.
// One thread per output pixel (3 channels); reads are served by L1/L2 (every input pixel is read ~16x), writes are
// coalesced along x.  HBM traffic = 3 B per input pixel + 12 B per output pixel.
#include <stdint.h>

#include "../../include/mgld.h"
#include "common.h"

namespace mgld {

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }
__device__ __forceinline__ void cubic_coeffs(float t, float* c) {   // ATen get_cubic_upsample_coefficients
  const float A = -0.75f;
  c[0] = cubic2(t + 1.f, A);
  c[1] = cubic1(t, A);
  const float x2 = 1.f - t;
  c[2] = cubic1(x2, A);
  c[3] = cubic2(x2 + 1.f, A);
}
__device__ __forceinline__ float cubic_interp(float x0, float x1, float x2, float x3, const float* c) {
  return x0 * c[0] + x1 * c[1] + x2 * c[2] + x3 * c[3];
}

__global__ void u8_to_f32_bicubic_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int N, int H, int W,
                                         int OH, int OW, int PH, int PW, float scale_h, float scale_w, int clamp) {
  const int px = blockIdx.x * blockDim.x + threadIdx.x;
  const int py = blockIdx.y;
  const int n = blockIdx.z;
  if (px >= PW) return;
  // reflect pad (right / bottom only, script :387): padded index p >= O maps to 2 (O - 1) - p
  const int ox = px < OW ? px : 2 * (OW - 1) - px;
  const int oy = py < OH ? py : 2 * (OH - 1) - py;
  const float sx = scale_w * ((float)ox + 0.5f) - 0.5f;
  const float sy = scale_h * ((float)oy + 0.5f) - 0.5f;
  const float fx = floorf(sx), fy = floorf(sy);
  const int ix = (int)fx, iy = (int)fy;
  float cx[4], cy[4];
  cubic_coeffs(sx - fx, cx);
  cubic_coeffs(sy - fy, cy);
  const uint8_t* img = in + static_cast<long long>(n) * H * W * 3;
  float acc[3];
  float rows[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int y = min(max(iy - 1 + j, 0), H - 1);
    float v[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int x = min(max(ix - 1 + i, 0), W - 1);
      const uint8_t* p = img + (static_cast<long long>(y) * W + x) * 3;
#pragma unroll
      for (int c = 0; c < 3; ++c) v[i][c] = ((float)p[c] / 255.0f - 0.5f) / 0.5f;   // read_image: float32 arithmetic
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) rows[j][c] = cubic_interp(v[0][c], v[1][c], v[2][c], v[3][c], cx);
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    acc[c] = cubic_interp(rows[0][c], rows[1][c], rows[2][c], rows[3][c], cy);
    if (clamp) acc[c] = fminf(fmaxf(acc[c], -1.f), 1.f);
    out[((static_cast<long long>(n) * 3 + c) * PH + py) * PW + px] = acc[c];
  }
}

__global__ void f32_to_u8_hwc_kernel(const float* __restrict__ in, uint8_t* __restrict__ out, int N, int H, int W, int CH,
                                     int CW) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // over N * CH * CW pixels
  const long long total = static_cast<long long>(N) * CH * CW;
  if (i >= total) return;
  const int x = (int)(i % CW);
  const int y = (int)((i / CW) % CH);
  const int n = (int)(i / ((long long)CW * CH));
  uint8_t* o = out + i * 3;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = in[((static_cast<long long>(n) * 3 + c) * H + y) * W + x] * 255.0f;
    o[c] = (uint8_t)(int)v;          // numpy float32 -> uint8: truncation toward zero (inputs are clamped to [0,1] upstream)
  }
}

}  // namespace mgld

using namespace mgld;

extern "C" int mgld_frames_u8_to_f32_bicubic(const void* in, float* out, int n, int h, int w, int oh, int ow, int pad_h,
                                             int pad_w, int clamp, void* stream) {
  MGLD_CHECK_ARG(in && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && pad_h >= 0 && pad_w >= 0 && pad_h < oh &&
                     pad_w < ow, "frames_u8_to_f32_bicubic: bad arguments");
  dim3 grid(ceil_div(ow + pad_w, 128), oh + pad_h, n);
  u8_to_f32_bicubic_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)in, out, n, h, w, oh, ow, oh + pad_h,
                                                                   ow + pad_w, (float)h / (float)oh, (float)w / (float)ow,
                                                                   clamp);
  MGLD_LAUNCH_CHECK("u8_to_f32_bicubic_kernel");
  return MGLD_OK;
}

extern "C" int mgld_frames_f32_to_u8_hwc(const float* in, void* out, int n, int h, int w, int crop_h, int crop_w,
                                         void* stream) {
  MGLD_CHECK_ARG(in && out && n > 0 && crop_h > 0 && crop_w > 0 && crop_h <= h && crop_w <= w,
                 "frames_f32_to_u8_hwc: bad arguments");
  const long long total = 1LL * n * crop_h * crop_w;
  f32_to_u8_hwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, (uint8_t*)out, n, h, w, crop_h,
                                                                                         crop_w);
  MGLD_LAUNCH_CHECK("f32_to_u8_hwc_kernel");
  return MGLD_OK;
}
