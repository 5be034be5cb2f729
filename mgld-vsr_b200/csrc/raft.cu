// RAFT-specific kernels (everything that is not a convolution / GEMM): stem convolutions with 2-3 input channels,
// instance norm, all-pairs correlation pyramid + windowed bilinear lookup, SepConvGRU gating, convex 8x upsampling.
// The convolutions themselves run through mgld_conv_gemm (3x3, 1x1, 1x5, 5x1 taps; stride-2 via mgld_im2col_s2).
//
// Reference: basicsr/archs/raft_arch.py — CorrBlock :37-92, BasicEncoder :199-268, SepConvGRU :379-412,
// BasicMotionEncoder :426-445, upsample_flow :720-731, bilinear_sampler :517-532.
#include <math.h>

#include "../../include/mgld.h"
#include "common.h"

namespace mgld {

// direct convolution for tiny Cin (<= 4): (N,Cin,H,W) fp32 -> NHWC fp16 [N,Ho,Wo,ldo], any odd ks, stride 1|2, ReLU opt.
__global__ void conv_direct_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                   const float* __restrict__ bias, __half* __restrict__ out, int N, int Cin, int H,
                                   int W, int Cout, int ks, int stride, int pad, int Ho, int Wo, int ldo, int relu) {
  const long long total = static_cast<long long>(N) * Ho * Wo * Cout;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = i % Cout;
    long long p = i / Cout;
    const int ox = p % Wo; p /= Wo;
    const int oy = p % Ho;
    const int n = p / Ho;
    float acc = bias ? __ldg(bias + co) : 0.f;
    for (int c = 0; c < Cin; ++c)
      for (int ky = 0; ky < ks; ++ky) {
        const int iy = oy * stride + ky - pad;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < ks; ++kx) {
          const int ix = ox * stride + kx - pad;
          if (ix < 0 || ix >= W) continue;
          acc = fmaf(__ldg(in + ((static_cast<long long>(n) * Cin + c) * H + iy) * W + ix),
                     __ldg(w + ((static_cast<long long>(co) * Cin + c) * ks + ky) * ks + kx), acc);
        }
      }
    if (relu) acc = fmaxf(acc, 0.f);
    out[((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + co] = __float2half_rn(acc);
  }
}

// every second pixel (input of a stride-2 1x1 convolution), NHWC fp16, 16-byte vectors
__global__ void subsample2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int N, int H, int W, int vpr) {
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  const long long total = static_cast<long long>(N) * Ho * Wo * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = i % vpr;
    long long p = i / vpr;
    const int ox = p % Wo; p /= Wo;
    const int oy = p % Ho;
    const int n = p / Ho;
    out[i] = __ldg(in + ((static_cast<long long>(n) * H + 2 * oy) * W + 2 * ox) * vpr + v);
  }
}

// instance norm (no affine) from per-(n, channel) fixed-point sums [N, C, 2] x 16 bytes (mgld_gn_stats_f16 with groups = C) + optional ReLU
__global__ void inorm_apply_kernel(const __half* __restrict__ x, const double* __restrict__ sums, __half* __restrict__ out,
                                   int HW, int C, double eps, int relu) {
  extern __shared__ float sh[];  // [2*C] mean, rstd
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const double mean = fixsum_load(sums + (static_cast<long long>(n) * C + c) * 4) / HW;
    double var = fixsum_load(sums + (static_cast<long long>(n) * C + c) * 4 + 2) / HW - mean * mean;
    if (var < 0.0) var = 0.0;
    sh[2 * c] = (float)mean;
    sh[2 * c + 1] = (float)(1.0 / sqrt(var + eps));
  }
  __syncthreads();
  const long long total = static_cast<long long>(HW) * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long idx = static_cast<long long>(n) * total + i;
    float y = (__half2float(x[idx]) - sh[2 * c]) * sh[2 * c + 1];
    if (relu) y = fmaxf(y, 0.f);
    out[idx] = __float2half_rn(y);
  }
}

// 2x2 average pooling (floor) of a stack of fp32 maps: (n, h, w) -> (n, h/2, w/2)       raft_arch.py:50-52
__global__ void avgpool2_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, int h, int w) {
  const int ho = h / 2, wo = w / 2;
  const long long total = n * ho * wo;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ox = i % wo;
    const int oy = (i / wo) % ho;
    const long long m = i / (static_cast<long long>(wo) * ho);
    const float* b = in + (m * h + 2 * oy) * w + 2 * ox;
    out[i] = (b[0] + b[1] + b[w] + b[w + 1]) * 0.25f;
  }
}

// windowed lookup of the 4-level correlation pyramid (CorrBlock.__call__): for pixel p of pair b and level l the
// 9x9 window index (i,j) samples  x = cx/2^l + (i-4),  y = cy/2^l + (j-4)   — the reference adds `dy` to x and `dx` to y
// (raft_arch.py:66-72: delta = stack(meshgrid(dy, dx)) is added to (x, y) coordinates) — bilinear, zeros padding.
// out: NHWC fp16 [B, h*w, ldo], channel = l*81 + i*9 + j; columns [324, ldo) are left untouched (zero padding).
struct CorrLevels { const float* p[4]; int h[4]; int w[4]; };
__global__ void corr_lookup_kernel(CorrLevels lv, const float* __restrict__ coords, __half* __restrict__ out, int B,
                                   int H, int W, int ldo) {
  const int hw = H * W;
  const long long total = static_cast<long long>(B) * hw * 324;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = i % 324;
    const long long bp = i / 324;
    const int p = bp % hw, b = bp / hw;
    const int l = ch / 81, wi = (ch % 81) / 9, wj = ch % 9;
    const float cx = coords[(static_cast<long long>(b) * 2) * hw + p], cy = coords[(static_cast<long long>(b) * 2 + 1) * hw + p];
    const float sc = 1.f / (float)(1 << l);
    const float x = cx * sc + (float)(wi - 4), y = cy * sc + (float)(wj - 4);
    const int h = lv.h[l], w = lv.w[l];
    // bilinear_sampler: grid = 2*x/(W-1) - 1, align_corners=True  ->  source index = ((g+1)/2)*(W-1)
    const float gx = 2.f * x / (float)(w - 1) - 1.f, gy = 2.f * y / (float)(h - 1) - 1.f;
    const float ix = (gx + 1.f) * 0.5f * (float)(w - 1), iy = (gy + 1.f) * 0.5f * (float)(h - 1);
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wx1 = ix - fx0, wy1 = iy - fy0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
    const float* m = lv.p[l] + (static_cast<long long>(b) * hw + p) * h * w;
    float acc = 0.f;
    if (y0 >= 0 && y0 < h) {
      if (x0 >= 0 && x0 < w) acc = fmaf(m[y0 * w + x0], wx0 * wy0, acc);
      if (x0 + 1 >= 0 && x0 + 1 < w) acc = fmaf(m[y0 * w + x0 + 1], wx1 * wy0, acc);
    }
    if (y0 + 1 >= 0 && y0 + 1 < h) {
      if (x0 >= 0 && x0 < w) acc = fmaf(m[(y0 + 1) * w + x0], wx0 * wy1, acc);
      if (x0 + 1 >= 0 && x0 + 1 < w) acc = fmaf(m[(y0 + 1) * w + x0 + 1], wx1 * wy1, acc);
    }
    out[bp * ldo + ch] = __float2half_rn(acc);
  }
}

// SepConvGRU gating (raft_arch.py:398-412).  zr: [M, 2C] = sigmoid(convz | convr)
__global__ void gru_rh_kernel(const __half* __restrict__ zr, const __half* __restrict__ net, __half* __restrict__ rnet,
                              long long M, int C) {
  const long long total = M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / C;
    const int c = i % C;
    rnet[i] = __float2half_rn(__half2float(zr[m * 2 * C + C + c]) * __half2float(net[i]));
  }
}
__global__ void gru_update_kernel(const __half* __restrict__ zr, const __half* __restrict__ q, __half* __restrict__ net,
                                  long long M, int C) {
  const long long total = M * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / C;
    const int c = i % C;
    const float z = __half2float(zr[m * 2 * C + c]);
    net[i] = __float2half_rn((1.f - z) * __half2float(net[i]) + z * __half2float(q[i]));
  }
}

// write an (B, Cs, h, w) fp32 tensor into columns [col0, col0+Cs) of an NHWC fp16 buffer with row pitch ld
__global__ void set_channels_kernel(const float* __restrict__ src, __half* __restrict__ dst, int B, int Cs, int HW, int ld,
                                    int col0) {
  const long long total = static_cast<long long>(B) * Cs * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = i % HW;
    const int c = (i / HW) % Cs;
    const int b = i / (static_cast<long long>(HW) * Cs);
    dst[(static_cast<long long>(b) * HW + p) * ld + col0 + c] = __float2half_rn(src[i]);
  }
}

// convex 8x upsampling (upsample_flow, raft_arch.py:720-731): mask NHWC fp16 [B,h,w,576], channel = k*64 + (sy*8 + sx),
// softmax over the 9 neighbours k; out[b, c, 8y+sy, 8x+sx] = sum_k softmax_k * 8*flow[b, c, y+ky-1, x+kx-1] (zero pad)
__global__ void convex_upsample_kernel(const __half* __restrict__ mask, const float* __restrict__ flow,
                                       float* __restrict__ out, int B, int H, int W) {
  const long long total = static_cast<long long>(B) * H * W * 64;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int sub = i % 64;
    const long long bp = i / 64;
    const int x = bp % W, y = (bp / W) % H, b = bp / (static_cast<long long>(W) * H);
    const __half* mp = mask + bp * 576 + sub;
    float mv[9], mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < 9; ++k) { mv[k] = __half2float(mp[k * 64]); mx = fmaxf(mx, mv[k]); }
    float den = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) { mv[k] = __expf(mv[k] - mx); den += mv[k]; }
    const float inv = 1.f / den;
    float fx = 0.f, fy = 0.f;
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
        const float wk = mv[k] * inv;
        fx = fmaf(wk, 8.f * flow[((static_cast<long long>(b) * 2) * H + yy) * W + xx], fx);
        fy = fmaf(wk, 8.f * flow[((static_cast<long long>(b) * 2 + 1) * H + yy) * W + xx], fy);
      }
    }
    const int sy = sub / 8, sx = sub % 8;
    const long long o = (static_cast<long long>(b) * 2 * 8 * H + (8 * y + sy)) * 8 * W + 8 * x + sx;
    out[o] = fx;
    out[o + static_cast<long long>(8 * H) * 8 * W] = fy;
  }
}

}  // namespace mgld

using namespace mgld;

static inline int grid_1d(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int mgld_conv_direct_f32(const float* in, const float* w, const float* bias, void* out, int n, int cin, int h,
                                    int wd, int cout, int ks, int stride, int pad, int ldo, int relu, void* stream) {
  MGLD_CHECK_ARG(in && w && out && cin > 0 && cin <= 4 && ks % 2 == 1 && (stride == 1 || stride == 2), "conv_direct: bad arguments");
  const int ho = (h + 2 * pad - ks) / stride + 1, wo = (wd + 2 * pad - ks) / stride + 1;
  const long long total = 1LL * n * ho * wo * cout;
  conv_direct_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(in, w, bias, (__half*)out, n, cin, h, wd, cout, ks,
                                                                             stride, pad, ho, wo, ldo > 0 ? ldo : cout, relu);
  MGLD_LAUNCH_CHECK("conv_direct_kernel");
  return MGLD_OK;
}
extern "C" int mgld_subsample2_f16(const void* in, void* out, int n, int h, int w, int c, void* stream) {
  MGLD_CHECK_ARG(in && out && c % 8 == 0, "subsample2: bad arguments");
  const long long total = 1LL * n * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8);
  subsample2_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)in, (uint4*)out, n, h, w, c / 8);
  MGLD_LAUNCH_CHECK("subsample2_kernel");
  return MGLD_OK;
}
extern "C" int mgld_instance_norm_apply_f16(const void* x, const double* sums, void* out, int n, int hw, int c, double eps,
                                            int relu, void* stream) {
  MGLD_CHECK_ARG(x && sums && out && n > 0 && c > 0 && c <= 2048, "instance_norm_apply: bad arguments");
  dim3 grid(grid_1d(1LL * hw * c, 256 * 8), n);
  inorm_apply_kernel<<<grid, 256, 2 * c * sizeof(float), (cudaStream_t)stream>>>((const __half*)x, sums, (__half*)out, hw, c, eps, relu);
  MGLD_LAUNCH_CHECK("inorm_apply_kernel");
  return MGLD_OK;
}
extern "C" int mgld_avgpool2_f32(const float* in, float* out, long long n, int h, int w, void* stream) {
  MGLD_CHECK_ARG(in && out && n > 0 && h >= 2 && w >= 2, "avgpool2: bad arguments");
  avgpool2_kernel<<<grid_1d(n * (h / 2) * (w / 2), 256), 256, 0, (cudaStream_t)stream>>>(in, out, n, h, w);
  MGLD_LAUNCH_CHECK("avgpool2_kernel");
  return MGLD_OK;
}
extern "C" int mgld_corr_lookup_f32(const float* l0, const float* l1, const float* l2, const float* l3, const float* coords,
                                    void* out, int b, int h, int w, int ldo, void* stream) {
  MGLD_CHECK_ARG(l0 && l1 && l2 && l3 && coords && out && ldo >= 324, "corr_lookup: bad arguments");
  MGLD_CHECK_ARG(h / 8 >= 2 && w / 8 >= 2, "corr_lookup: the coarsest pyramid level must be at least 2x2 (1/8-res map %dx%d)", h, w);
  CorrLevels lv;
  lv.p[0] = l0; lv.p[1] = l1; lv.p[2] = l2; lv.p[3] = l3;
  int hh = h, ww = w;
  for (int l = 0; l < 4; ++l) { lv.h[l] = hh; lv.w[l] = ww; hh /= 2; ww /= 2; }
  corr_lookup_kernel<<<grid_1d(1LL * b * h * w * 324, 256), 256, 0, (cudaStream_t)stream>>>(lv, coords, (__half*)out, b, h, w, ldo);
  MGLD_LAUNCH_CHECK("corr_lookup_kernel");
  return MGLD_OK;
}
extern "C" int mgld_gru_rh_f16(const void* zr, const void* net, void* rnet, long long m, int c, void* stream) {
  MGLD_CHECK_ARG(zr && net && rnet && m > 0 && c > 0, "gru_rh: bad arguments");
  gru_rh_kernel<<<grid_1d(m * c, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)zr, (const __half*)net, (__half*)rnet, m, c);
  MGLD_LAUNCH_CHECK("gru_rh_kernel");
  return MGLD_OK;
}
extern "C" int mgld_gru_update_f16(const void* zr, const void* q, void* net, long long m, int c, void* stream) {
  MGLD_CHECK_ARG(zr && q && net && m > 0 && c > 0, "gru_update: bad arguments");
  gru_update_kernel<<<grid_1d(m * c, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)zr, (const __half*)q, (__half*)net, m, c);
  MGLD_LAUNCH_CHECK("gru_update_kernel");
  return MGLD_OK;
}
extern "C" int mgld_set_channels_f16(const float* src, void* dst, int b, int cs, int hw, int ld, int col0, void* stream) {
  MGLD_CHECK_ARG(src && dst && b > 0 && cs > 0, "set_channels: bad arguments");
  set_channels_kernel<<<grid_1d(1LL * b * cs * hw, 256), 256, 0, (cudaStream_t)stream>>>(src, (__half*)dst, b, cs, hw, ld, col0);
  MGLD_LAUNCH_CHECK("set_channels_kernel");
  return MGLD_OK;
}
extern "C" int mgld_convex_upsample8_f32(const void* mask, const float* flow, float* out, int b, int h, int w, void* stream) {
  MGLD_CHECK_ARG(mask && flow && out && b > 0, "convex_upsample: bad arguments");
  convex_upsample_kernel<<<grid_1d(1LL * b * h * w * 64, 256), 256, 0, (cudaStream_t)stream>>>((const __half*)mask, flow, out, b, h, w);
  MGLD_LAUNCH_CHECK("convex_upsample_kernel");
  return MGLD_OK;
}
