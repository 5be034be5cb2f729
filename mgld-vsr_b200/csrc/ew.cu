// Small HBM-bound helpers around the tensor-core kernels: layout conversion, nearest upsample, stride-2 im2col,
// tiny-channel direct convolutions (network stems / heads), GEMV for the timestep-embedding MLPs, the temporal
// (sequence length T) attention of the UNet middle block, and the VAE posterior sample.
//
// Reference call sites: Upsample openaimodel.py:178-188 / model.py:95-99; Downsample openaimodel.py:204-231 /
// model.py:104-121; conv_in / conv_out / out / quant_conv stems (openaimodel.py:2036-2040,2253-2257, model.py:496,
// 536, autoencoder.py:331-332); timestep_embedding util.py:151-171; time_embed / emb_layers
// openaimodel.py:2021-2025,418-424; TemporalAttention attention.py:124-143; DiagonalGaussianDistribution.sample
// distributions.py:24-37.
#include <math.h>

#include "../../include/mgld.h"
#include "common.h"

namespace mgld {

// ---------------------------------------------------------------------------------------------------------------
// layout conversion: (N,C,H,W) fp32 <-> (N,H,W,C) fp16, through a 32x32 smem transpose tile
// ---------------------------------------------------------------------------------------------------------------
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, __half* __restrict__ out, int C, int HW, int ldo,
                                    float scale) {
  pdl_launch_dependents();
  pdl_wait();

  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? in[(static_cast<long long>(n) * C + c) * HW + p] * scale : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[(static_cast<long long>(n) * HW + p) * ldo + c] = __float2half_rn(tile[threadIdx.x][i]);
  }
}
__global__ void nhwc_to_nchw_kernel(const __half* __restrict__ in, float* __restrict__ out, int C, int HW, int ldi,
                                    float scale) {
  pdl_launch_dependents();
  pdl_wait();

  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (p < HW && c < C) ? __half2float(in[(static_cast<long long>(n) * HW + p) * ldi + c]) * scale : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[(static_cast<long long>(n) * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// nearest 2x upsample, NHWC fp16 (one 16-byte vector per thread)
// ---------------------------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T, int H, int W, int vpr) {
  pdl_launch_dependents();
  pdl_wait();

  const long long total = static_cast<long long>(T) * 4 * H * W * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = i % vpr;
    long long p = i / vpr;
    const int ox = p % (2 * W); p /= (2 * W);
    const int oy = p % (2 * H);
    const int t = p / (2 * H);
    out[i] = __ldg(in + ((static_cast<long long>(t) * H + (oy >> 1)) * W + (ox >> 1)) * vpr + v);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// im2col for 3x3 stride-2 convolutions: out[t, oy, ox, tap*C + c] = in[t, 2*oy + ky - pad, 2*ox + kx - pad, c]
// (pad = 1: UNet Downsample conv(stride 2, padding 1); pad = 0 with zero fill on the far side: VAE Downsample,
//  F.pad(x, (0,1,0,1)) then conv(stride 2, padding 0))
// ---------------------------------------------------------------------------------------------------------------
__global__ void im2col_s2_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T, int H, int W, int Ho,
                                 int Wo, int vpr, int pad) {
  pdl_launch_dependents();
  pdl_wait();

  const long long total = static_cast<long long>(T) * Ho * Wo * 9 * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = i % vpr;
    long long p = i / vpr;
    const int tap = p % 9; p /= 9;
    const int ox = p % Wo; p /= Wo;
    const int oy = p % Ho;
    const int t = p / Ho;
    const int iy = 2 * oy + tap / 3 - pad, ix = 2 * ox + tap % 3 - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) val = __ldg(in + ((static_cast<long long>(t) * H + iy) * W + ix) * vpr + v);
    out[i] = val;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// direct conv with tiny Cin (<= 8): in (N,Cin,H,W) fp32 -> out NHWC fp16 [N,H,W,Cout] (+bias), k = 1 or 3 (pad k/2)
// ---------------------------------------------------------------------------------------------------------------
// blockIdx.y selects a chunk of 128 output channels (thread == channel, its <=72 weights live in registers);
// blockIdx.x a batch of 64 pixels whose input patches are staged once in shared memory and reused by all channels.
// Four pixels are accumulated at a time: four independent FMA chains per thread and one 16-byte broadcast read per 4 FMAs
// (the first version ran one 36-deep dependent chain per pixel and was 10x off the FMA rate).
// kK4 = number of 4-tap groups (compile time: 9 for 4x3x3, 7 for 3x3x3), 0 = run-time count up to 18.
constexpr int kScPix = 64;
template <int kK4>
__global__ void __launch_bounds__(128)
conv_small_cin_kernel(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                      __half* __restrict__ out, int N, int Cin, int H, int W, int Cout, int ks, int ldo) {
  pdl_launch_dependents();
  pdl_wait();

  __shared__ __align__(16) float patch[kScPix][76];   // 72 taps max, padded to float4 rows
  const int K = Cin * ks * ks, pad = ks / 2;
  const long long total = static_cast<long long>(N) * H * W;
  const long long p0 = static_cast<long long>(blockIdx.x) * kScPix;
  const int co = blockIdx.y * 128 + threadIdx.x;
  {
    // two threads per pixel: the pixel's coordinates are decoded once, each thread gathers every other tap
    const int pp = threadIdx.x >> 1, half = threadIdx.x & 1;
    const long long p = p0 + pp;
    const bool live = p < total;
    const int x = live ? static_cast<int>(p % W) : 0, y = live ? static_cast<int>((p / W) % H) : 0;
    const long long n = live ? p / (static_cast<long long>(W) * H) : 0;
    const float* base = in + n * Cin * H * W;
    int c = 0, ky = 0, kx = half;          // tap k = (c * ks + ky) * ks + kx, starting at k = half, stepping by 2
    if (kx >= ks) { kx -= ks; ky = 1; }    // ks == 1
    if (ky >= ks) { ky = 0; c = 1; }
    for (int k = half; k < 76; k += 2) {
      float v = 0.f;
      if (live && k < K) {
        const int iy = y + ky - pad, ix = x + kx - pad;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(base + (static_cast<long long>(c) * H + iy) * W + ix);
      }
      patch[pp][k] = v;
      kx += 2;
      while (kx >= ks) { kx -= ks; if (++ky >= ks) { ky = 0; ++c; } }
    }
  }
  __syncthreads();
  if (co >= Cout) return;
  constexpr int kMaxK4 = kK4 > 0 ? kK4 : 18;
  const int K4 = kK4 > 0 ? kK4 : (K + 3) >> 2;
  float wr[kMaxK4 * 4];
#pragma unroll
  for (int k = 0; k < kMaxK4 * 4; ++k) wr[k] = k < K ? __ldg(w + static_cast<long long>(co) * K + k) : 0.f;
  const float b = bias ? __ldg(bias + co) : 0.f;
  for (int pp = 0; pp < kScPix; pp += 4) {
    if (p0 + pp >= total) break;
    float acc[4] = {b, b, b, b};
#pragma unroll
    for (int k4 = 0; k4 < kMaxK4; ++k4) {
      if (kK4 > 0 || k4 < K4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 pv = *reinterpret_cast<const float4*>(&patch[pp + q][k4 * 4]);
          acc[q] = fmaf(pv.x, wr[k4 * 4], acc[q]);
          acc[q] = fmaf(pv.y, wr[k4 * 4 + 1], acc[q]);
          acc[q] = fmaf(pv.z, wr[k4 * 4 + 2], acc[q]);
          acc[q] = fmaf(pv.w, wr[k4 * 4 + 3], acc[q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (p0 + pp + q < total) out[(p0 + pp + q) * ldo + co] = __float2half_rn(acc[q]);
  }
}

// direct conv with tiny Cin and tiny Cout, NCHW fp32 -> NCHW fp32 (quant_conv / post_quant_conv, 1x1 or 3x3)
__global__ void conv_small_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ out, int N, int Cin, int H,
                                      int W, int Cout, int ks) {
  const long long total = static_cast<long long>(N) * Cout * H * W;
  const int pad = ks / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = i % W;
    long long r = i / W;
    const int y = r % H; r /= H;
    const int co = r % Cout;
    const int n = r / Cout;
    float acc = bias ? bias[co] : 0.f;
    for (int c = 0; c < Cin; ++c)
      for (int ky = 0; ky < ks; ++ky)
        for (int kx = 0; kx < ks; ++kx) {
          const int iy = y + ky - pad, ix = x + kx - pad;
          if (iy >= 0 && iy < H && ix >= 0 && ix < W)
            acc = fmaf(in[((static_cast<long long>(n) * Cin + c) * H + iy) * W + ix], w[((co * Cin + c) * ks + ky) * ks + kx], acc);
        }
    out[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// direct 3x3 conv with tiny Cout (<= 8): in NHWC fp16 [N,H,W,C] -> out (N,Cout,H,W) fp32; one warp per pixel,
// lanes stride the channel vectors, shuffle reduction.  w: [Cout][9][C] fp16 (packed like conv_gemm weights).
// ---------------------------------------------------------------------------------------------------------------
template <int COUT>
__global__ void conv_small_cout_kernel(const __half* __restrict__ in, const __half* __restrict__ w,
                                       const float* __restrict__ bias, float* __restrict__ out, int N, int H, int W,
                                       int C, int ldi) {
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long long total = static_cast<long long>(N) * H * W;
  if (warp >= total) return;
  const int x = warp % W, y = (warp / W) % H, n = warp / (static_cast<long long>(W) * H);
  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;
  const int vpr = C >> 3;
  for (int tap = 0; tap < 9; ++tap) {
    const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
    if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
    const __half* row = in + ((static_cast<long long>(n) * H + iy) * W + ix) * ldi;
    for (int v = lane; v < vpr; v += 32) {
      const uint4 xv = __ldg(reinterpret_cast<const uint4*>(row + v * 8));
      const __half2* xh = reinterpret_cast<const __half2*>(&xv);
#pragma unroll
      for (int o = 0; o < COUT; ++o) {
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + (static_cast<long long>(o) * 9 + tap) * C + v * 8));
        const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 a = __half22float2(xh[i]), b = __half22float2(wh[i]);
          acc[o] = fmaf(a.x, b.x, acc[o]);
          acc[o] = fmaf(a.y, b.y, acc[o]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float v = acc[o];
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0) out[((static_cast<long long>(n) * COUT + o) * H + y) * W + x] = v + (bias ? bias[o] : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// y[n] = act_out( sum_k act_in(x[k]) * W[n,k] + b[n] ),  x fp32 [K], W fp16 [N,K]; one warp per output row
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float v) { return v / (1.f + __expf(-v)); }
__global__ void gemv_kernel(const float* __restrict__ x, const __half* __restrict__ w, const float* __restrict__ bias,
                            const float* __restrict__ add, float* __restrict__ y, int N, int K, int silu_in,
                            int silu_out) {
  pdl_launch_dependents();
  pdl_wait();

  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= N) return;
  float acc = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + static_cast<long long>(row) * K + k));
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 b = __half22float2(wh[i]);
      float x0 = x[k + 2 * i], x1 = x[k + 2 * i + 1];
      if (silu_in) { x0 = silu_f(x0); x1 = silu_f(x1); }
      acc = fmaf(x0, b.x, acc);
      acc = fmaf(x1, b.y, acc);
    }
  }
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if (lane == 0) {
    float v = acc + (bias ? bias[row] : 0.f) + (add ? add[row] : 0.f);
    if (silu_out) v = silu_f(v);
    y[row] = v;
  }
}

// timestep_embedding (util.py:151-171): emb[i] = cos(t*f_i), emb[half+i] = sin(t*f_i), f_i = exp(-ln(max_period)*i/half)
__global__ void timestep_embedding_kernel(const float* __restrict__ tp, float* __restrict__ out, int dim, float max_period) {
  pdl_launch_dependents();
  pdl_wait();

  const float t = *tp;
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= half) return;
  const float freq = expf(-logf(max_period) * (float)i / (float)half);
  const float a = t * freq;
  out[i] = cosf(a);
  out[half + i] = sinf(a);
}

// ---------------------------------------------------------------------------------------------------------------
// temporal self-attention (sequence = T frames, batch = pixels x heads, head_dim 64): one warp per (pixel, head)
// qkv: fp16 [T, HW, 3C] (q | k | v, head h at h*64), out: fp16 [T, HW, C]
// ---------------------------------------------------------------------------------------------------------------
constexpr int kMaxT = 8;
__global__ void temporal_attention_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int T, int HW, int C,
                                          int heads, float scale) {
  pdl_launch_dependents();
  pdl_wait();

  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= static_cast<long long>(HW) * heads) return;
  const int h = warp % heads;
  const int pix = warp / heads;
  float2 q[kMaxT], k[kMaxT], v[kMaxT];
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
    if (t < T) {
      const __half* base = qkv + (static_cast<long long>(t) * HW + pix) * 3 * C + h * 64 + lane * 2;
      q[t] = __half22float2(*reinterpret_cast<const __half2*>(base));
      k[t] = __half22float2(*reinterpret_cast<const __half2*>(base + C));
      v[t] = __half22float2(*reinterpret_cast<const __half2*>(base + 2 * C));
    }
  }
#pragma unroll
  for (int i = 0; i < kMaxT; ++i) {
    if (i >= T) break;
    float s[kMaxT];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxT; ++j) {
      if (j < T) {
        float d = q[i].x * k[j].x + q[i].y * k[j].y;
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        s[j] = d * scale;
        mx = fmaxf(mx, s[j]);
      }
    }
    float sum = 0.f, ox = 0.f, oy = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxT; ++j) {
      if (j < T) {
        const float p = __expf(s[j] - mx);
        sum += p;
        ox = fmaf(p, v[j].x, ox);
        oy = fmaf(p, v[j].y, oy);
      }
    }
    const float inv = 1.f / sum;
    *reinterpret_cast<__half2*>(out + (static_cast<long long>(i) * HW + pix) * C + h * 64 + lane * 2) =
        __floats2half2_rn(ox * inv, oy * inv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// DiagonalGaussianDistribution.sample on moments (N, 2*Cz, H, W) fp32:  z = (mean + exp(0.5*clamp(logvar,-30,20)) * noise) * scale
// ---------------------------------------------------------------------------------------------------------------
__global__ void gaussian_sample_kernel(const float* __restrict__ moments, const float* __restrict__ noise,
                                       float* __restrict__ out, int N, int Cz, int HW, float scale) {
  const long long total = static_cast<long long>(N) * Cz * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int p = i % HW;
    const int c = (i / HW) % Cz;
    const int n = i / (static_cast<long long>(HW) * Cz);
    const float mean = moments[(static_cast<long long>(n) * 2 * Cz + c) * HW + p];
    float lv = moments[(static_cast<long long>(n) * 2 * Cz + Cz + c) * HW + p];
    lv = fminf(fmaxf(lv, -30.f), 20.f);
    const float std = expf(0.5f * lv);
    out[i] = (noise ? __fadd_rn(mean, __fmul_rn(std, noise[i])) : mean) * scale;
  }
}

// out = a*x + b*y (fp16, vectorised) — residual adds that are not fused into a GEMM epilogue
__global__ void axpby_f16_kernel(const uint4* __restrict__ x, const uint4* __restrict__ y, uint4* __restrict__ out,
                                 float a, float b, long long nvec, int relu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 xv = __ldg(x + i), yv = __ldg(y + i);
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
    const __half2* yh = reinterpret_cast<const __half2*>(&yv);
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 p = __half22float2(xh[k]), q = __half22float2(yh[k]);
      float r0 = a * p.x + b * q.x, r1 = a * p.y + b * q.y;
      if (relu) { r0 = fmaxf(r0, 0.f); r1 = fmaxf(r1, 0.f); }
      oh[k] = __floats2half2_rn(r0, r1);
    }
    out[i] = o;
  }
}

}  // namespace mgld

using namespace mgld;

static inline int grid_1d(long long total, int threads) {
  long long b = (total + threads - 1) / threads;
  const long long cap = 148LL * 32;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

extern "C" int mgld_nchw_f32_to_nhwc_f16(const float* in, void* out, int n, int c, int h, int w, int ldo, float scale,
                                         void* stream) {
  MGLD_CHECK_ARG(in && out && n > 0 && c > 0 && h > 0 && w > 0, "nchw_to_nhwc: bad arguments");
  dim3 grid(ceil_div(h * w, 32), ceil_div(c, 32), n), block(32, 8);
  MGLD_CUDA(launch_pdl(nchw_to_nhwc_kernel, grid, block, 0, (cudaStream_t)stream, in, (__half*)out, c, h * w, ldo > 0 ? ldo : c, scale));
  MGLD_LAUNCH_CHECK("nchw_to_nhwc_kernel");
  return MGLD_OK;
}
extern "C" int mgld_nhwc_f16_to_nchw_f32(const void* in, float* out, int n, int c, int h, int w, int ldi, float scale,
                                         void* stream) {
  MGLD_CHECK_ARG(in && out && n > 0 && c > 0 && h > 0 && w > 0, "nhwc_to_nchw: bad arguments");
  dim3 grid(ceil_div(h * w, 32), ceil_div(c, 32), n), block(32, 8);
  MGLD_CUDA(launch_pdl(nhwc_to_nchw_kernel, dim3(grid), dim3(block), 0, (cudaStream_t)stream, (const __half*)in, out, c, h * w, ldi > 0 ? ldi : c, scale));
  MGLD_LAUNCH_CHECK("nhwc_to_nchw_kernel");
  return MGLD_OK;
}
extern "C" int mgld_upsample_nearest2x_f16(const void* in, void* out, int t, int h, int w, int c, void* stream) {
  MGLD_CHECK_ARG(in && out && c % 8 == 0 && t > 0 && h > 0 && w > 0, "upsample2x: bad arguments");
  const long long total = 4LL * t * h * w * (c / 8);
  MGLD_CUDA(launch_pdl(upsample2x_kernel, dim3(grid_1d(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)in, (uint4*)out, t, h, w, c / 8));
  MGLD_LAUNCH_CHECK("upsample2x_kernel");
  return MGLD_OK;
}
extern "C" int mgld_im2col_s2_f16(const void* in, void* out, int t, int h, int w, int c, int ho, int wo, int pad,
                                  void* stream) {
  MGLD_CHECK_ARG(in && out && c % 8 == 0 && t > 0 && ho > 0 && wo > 0 && (pad == 0 || pad == 1), "im2col_s2: bad arguments");
  const long long total = 9LL * t * ho * wo * (c / 8);
  MGLD_CUDA(launch_pdl(im2col_s2_kernel, dim3(grid_1d(total, 256)), dim3(256), 0, (cudaStream_t)stream, (const uint4*)in, (uint4*)out, t, h, w, ho, wo, c / 8, pad));
  MGLD_LAUNCH_CHECK("im2col_s2_kernel");
  return MGLD_OK;
}
extern "C" int mgld_conv_small_cin_f32(const float* in, const float* w, const float* bias, void* out, int n, int cin,
                                       int h, int wd, int cout, int ks, int ldo, void* stream) {
  MGLD_CHECK_ARG(in && w && out && cin > 0 && cin <= 8 && (ks == 1 || ks == 3) && cout > 0, "conv_small_cin: bad arguments");
  const long long total = 1LL * n * h * wd;
  dim3 grid((unsigned)((total + kScPix - 1) / kScPix), (unsigned)ceil_div(cout, 128));
  const int K = cin * ks * ks, ld = ldo > 0 ? ldo : cout;
  cudaStream_t st = (cudaStream_t)stream;
  if (K == 36) MGLD_CUDA(launch_pdl(conv_small_cin_kernel<9>, dim3(grid), dim3(128), 0, st, in, w, bias, (__half*)out, n, cin, h, wd, cout, ks, ld));
  else if (K == 27) MGLD_CUDA(launch_pdl(conv_small_cin_kernel<7>, dim3(grid), dim3(128), 0, st, in, w, bias, (__half*)out, n, cin, h, wd, cout, ks, ld));
  else MGLD_CUDA(launch_pdl(conv_small_cin_kernel<0>, dim3(grid), dim3(128), 0, st, in, w, bias, (__half*)out, n, cin, h, wd, cout, ks, ld));
  MGLD_LAUNCH_CHECK("conv_small_cin_kernel");
  return MGLD_OK;
}
extern "C" int mgld_conv_small_f32(const float* in, const float* w, const float* bias, float* out, int n, int cin,
                                   int h, int wd, int cout, int ks, void* stream) {
  MGLD_CHECK_ARG(in && w && out && cin > 0 && cout > 0 && (ks == 1 || ks == 3), "conv_small_f32: bad arguments");
  const long long total = 1LL * n * cout * h * wd;
  conv_small_f32_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(in, w, bias, out, n, cin, h, wd, cout, ks);
  MGLD_LAUNCH_CHECK("conv_small_f32_kernel");
  return MGLD_OK;
}
extern "C" int mgld_conv3x3_small_cout_f16(const void* in, const void* w, const float* bias, float* out, int n, int h,
                                           int wd, int c, int cout, int ldi, void* stream) {
  MGLD_CHECK_ARG(in && w && out && c % 8 == 0 && n > 0, "conv_small_cout: bad arguments");
  const long long warps = 1LL * n * h * wd;
  const int threads = 256;
  const long long blocks = (warps * 32 + threads - 1) / threads;
  cudaStream_t s = (cudaStream_t)stream;
  const int ld = ldi > 0 ? ldi : c;
#define MGLD_SC(K) conv_small_cout_kernel<K><<<(unsigned)blocks, threads, 0, s>>>((const __half*)in, (const __half*)w, bias, out, n, h, wd, c, ld)
  switch (cout) {
    case 3: MGLD_SC(3); break;
    case 4: MGLD_SC(4); break;
    case 8: MGLD_SC(8); break;
    case 2: MGLD_SC(2); break;
    default: set_error("conv_small_cout: cout=%d unsupported (2,3,4,8)", cout); return MGLD_ERR_ARG;
  }
#undef MGLD_SC
  MGLD_LAUNCH_CHECK("conv_small_cout_kernel");
  return MGLD_OK;
}
extern "C" int mgld_gemv_f32(const float* x, const void* w, const float* bias, const float* add, float* y, int n, int k,
                             int silu_in, int silu_out, void* stream) {
  MGLD_CHECK_ARG(x && w && y && n > 0 && k > 0 && k % 8 == 0, "gemv: bad arguments");
  MGLD_CUDA(launch_pdl(gemv_kernel, dim3(ceil_div(n * 32, 256)), dim3(256), 0, (cudaStream_t)stream, x, (const __half*)w, bias, add, y, n, k, silu_in, silu_out));
  MGLD_LAUNCH_CHECK("gemv_kernel");
  return MGLD_OK;
}
extern "C" int mgld_timestep_embedding_f32(const float* t, float* out, int dim, float max_period, void* stream) {
  MGLD_CHECK_ARG(t && out && dim > 0 && dim % 2 == 0, "timestep_embedding: bad arguments");
  MGLD_CUDA(launch_pdl(timestep_embedding_kernel, dim3(ceil_div(dim / 2, 128)), dim3(128), 0, (cudaStream_t)stream, t, out, dim, max_period));
  MGLD_LAUNCH_CHECK("timestep_embedding_kernel");
  return MGLD_OK;
}
extern "C" int mgld_temporal_attention_f16(const void* qkv, void* out, int t, int hw, int c, int heads, float scale,
                                           void* stream) {
  MGLD_CHECK_ARG(qkv && out && t > 0 && t <= kMaxT && c == heads * 64, "temporal_attention: T=%d (max %d), C=%d, heads=%d (head_dim must be 64)", t, kMaxT, c, heads);
  const long long warps = 1LL * hw * heads;
  MGLD_CUDA(launch_pdl(temporal_attention_kernel, dim3((unsigned)((warps * 32 + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, (const __half*)qkv, (__half*)out, t, hw, c, heads, scale));
  MGLD_LAUNCH_CHECK("temporal_attention_kernel");
  return MGLD_OK;
}
extern "C" int mgld_gaussian_sample_f32(const float* moments, const float* noise, float* out, int n, int cz, int h, int w,
                                        float scale, void* stream) {
  MGLD_CHECK_ARG(moments && out && n > 0 && cz > 0, "gaussian_sample: bad arguments");
  const long long total = 1LL * n * cz * h * w;
  gaussian_sample_kernel<<<grid_1d(total, 256), 256, 0, (cudaStream_t)stream>>>(moments, noise, out, n, cz, h * w, scale);
  MGLD_LAUNCH_CHECK("gaussian_sample_kernel");
  return MGLD_OK;
}
extern "C" int mgld_axpby_f16(const void* x, const void* y, void* out, float a, float b, long long n, int relu, void* stream) {
  MGLD_CHECK_ARG(x && y && out && n > 0 && n % 8 == 0, "axpby: bad arguments");
  axpby_f16_kernel<<<grid_1d(n / 8, 256), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (const uint4*)y, (uint4*)out, a, b, n / 8, relu);
  MGLD_LAUNCH_CHECK("axpby_f16_kernel");
  return MGLD_OK;
}
