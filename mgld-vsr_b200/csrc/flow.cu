// Flow-guided latent kernels (fp32, NCHW — the layout the reference keeps latents and flows in).
//
//   mgld_flow_warp_f32            bilinear / nearest grid-sample warp      basicsr/archs/arch_util.py:156-194,
//                                                                          scripts/util_flow.py:64-111
//   mgld_flow_warp_bwd_input_f32  its adjoint w.r.t. the warped tensor     (autograd of F.grid_sample; test helper)
//   mgld_fb_consistency_f32       forward/backward occlusion masks         scripts/util_flow.py:114-136
//   mgld_motion_guidance_f32      fused motion-guided latent update        ldm/models/diffusion/ddpm.py:3538-3574 +
//                                                                          ddpm.py:4429-4435
//   mgld_resize_flow_f32          bilinear (align_corners=False) resize    basicsr/archs/arch_util.py:235-270
//   mgld_canvas_posterior_f32     eps-tile stitch + x0 + posterior + noise ddpm.py:4275-4316, 4404-4417
//
// All arithmetic mirrors the order of the PyTorch ops it replaces (normalise to [-1,1], un-normalise, floor, corner
// weights nw/ne/sw/se) so results agree with the reference to fp32 rounding.  These kernels are HBM-/latency-bound
// (a few MB per call); the point of fusing is launch count, not bandwidth.
#include <math.h>

#include "../../include/mgld.h"
#include "common.h"

namespace mgld {

struct Corner {
  int x0, y0;
  float nw, ne, sw, se;
  bool valid;  // false only for nearest-mode out-of-range (value is zero)
};

// PyTorch grid_sampler_compute_source_index for one axis
__device__ __forceinline__ float unnormalize(float g, int size, int align_corners) {
  if (align_corners) return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  return __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)size), -1.f), 0.5f);
}
__device__ __forceinline__ float clip_coord(float v, int size) { return fminf((float)(size - 1), fmaxf(v, 0.f)); }

// grid coordinate of output pixel (x,y) displaced by (fx,fy), normalised as the reference does:
//   2*(x+fx)/max(w-1,1) - 1      (arch_util.py:178-179; util_flow.py:76-77 divides by (w-1))
__device__ __forceinline__ void source_index(int x, int y, float fx, float fy, int h, int w, int align_corners,
                                             int padding_border, float* ix, float* iy) {
  const float vx = __fadd_rn((float)x, fx);
  const float vy = __fadd_rn((float)y, fy);
  const float gx = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, vx), (float)max(w - 1, 1)), -1.f);
  const float gy = __fadd_rn(__fdiv_rn(__fmul_rn(2.f, vy), (float)max(h - 1, 1)), -1.f);
  float sx = unnormalize(gx, w, align_corners);
  float sy = unnormalize(gy, h, align_corners);
  if (padding_border) { sx = clip_coord(sx, w); sy = clip_coord(sy, h); }
  *ix = sx;
  *iy = sy;
}

__device__ __forceinline__ Corner bilinear_corner(float ix, float iy) {
  Corner c;
  const float fx0 = floorf(ix), fy0 = floorf(iy);
  c.x0 = (int)fx0;
  c.y0 = (int)fy0;
  const float x1 = fx0 + 1.f, y1 = fy0 + 1.f;
  c.nw = __fmul_rn(x1 - ix, y1 - iy);
  c.ne = __fmul_rn(ix - fx0, y1 - iy);
  c.sw = __fmul_rn(x1 - ix, iy - fy0);
  c.se = __fmul_rn(ix - fx0, iy - fy0);
  c.valid = true;
  return c;
}

__device__ __forceinline__ float sample_bilinear(const float* __restrict__ img, int h, int w, const Corner& c) {
  float acc = 0.f;
  const bool xin0 = c.x0 >= 0 && c.x0 < w, xin1 = c.x0 + 1 >= 0 && c.x0 + 1 < w;
  const bool yin0 = c.y0 >= 0 && c.y0 < h, yin1 = c.y0 + 1 >= 0 && c.y0 + 1 < h;
  if (yin0 && xin0) acc = __fmaf_rn(img[c.y0 * w + c.x0], c.nw, acc);
  if (yin0 && xin1) acc = __fmaf_rn(img[c.y0 * w + c.x0 + 1], c.ne, acc);
  if (yin1 && xin0) acc = __fmaf_rn(img[(c.y0 + 1) * w + c.x0], c.sw, acc);
  if (yin1 && xin1) acc = __fmaf_rn(img[(c.y0 + 1) * w + c.x0 + 1], c.se, acc);
  return acc;
}

__device__ __forceinline__ void scatter_bilinear(float* __restrict__ g, int h, int w, const Corner& c, float v) {
  const bool xin0 = c.x0 >= 0 && c.x0 < w, xin1 = c.x0 + 1 >= 0 && c.x0 + 1 < w;
  const bool yin0 = c.y0 >= 0 && c.y0 < h, yin1 = c.y0 + 1 >= 0 && c.y0 + 1 < h;
  if (yin0 && xin0) atomicAdd(g + c.y0 * w + c.x0, c.nw * v);
  if (yin0 && xin1) atomicAdd(g + c.y0 * w + c.x0 + 1, c.ne * v);
  if (yin1 && xin0) atomicAdd(g + (c.y0 + 1) * w + c.x0, c.sw * v);
  if (yin1 && xin1) atomicAdd(g + (c.y0 + 1) * w + c.x0 + 1, c.se * v);
}

// Deterministic accumulation for the guidance gradient: every contribution is rounded once to a 2^-44 fixed-point
// integer and summed with 64-bit integer atomics.  Integer addition is associative, so the sum - and with it the whole
// DDPM trajectory - is bitwise repeatable whatever order the scheduler runs the scatter in (the fp32 atomicAdd form was
// not: the L1 sign() and the ~460x last step amplified its rounding differences).  |contribution| <=
// 1/(C*h*w) <= 1, at most a few dozen land on one element: no overflow; resolution 5.7e-14 << fp32 ulp of the result.
constexpr double kFixScale = 17592186044416.0;  // 2^44
__device__ __forceinline__ void fix_add(long long* acc, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)__double2ll_rn((double)v * kFixScale));
}
__device__ __forceinline__ void scatter_bilinear_fix(long long* __restrict__ g, int h, int w, const Corner& c, float v) {
  const bool xin0 = c.x0 >= 0 && c.x0 < w, xin1 = c.x0 + 1 >= 0 && c.x0 + 1 < w;
  const bool yin0 = c.y0 >= 0 && c.y0 < h, yin1 = c.y0 + 1 >= 0 && c.y0 + 1 < h;
  if (yin0 && xin0) fix_add(g + c.y0 * w + c.x0, c.nw * v);
  if (yin0 && xin1) fix_add(g + c.y0 * w + c.x0 + 1, c.ne * v);
  if (yin1 && xin0) fix_add(g + (c.y0 + 1) * w + c.x0, c.sw * v);
  if (yin1 && xin1) fix_add(g + (c.y0 + 1) * w + c.x0 + 1, c.se * v);
}

__device__ __forceinline__ void load_flow(const float* __restrict__ flow, int layout, int n, int hw, int p, float* fx,
                                          float* fy) {
  if (layout == 0) {  // (n,h,w,2)
    const float2 f = *reinterpret_cast<const float2*>(flow + (static_cast<long long>(n) * hw + p) * 2);
    *fx = f.x; *fy = f.y;
  } else {            // (n,2,h,w)
    *fx = flow[(static_cast<long long>(n) * 2) * hw + p];
    *fy = flow[(static_cast<long long>(n) * 2 + 1) * hw + p];
  }
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void flow_warp_kernel(const float* __restrict__ x, const float* __restrict__ flow, float* __restrict__ out,
                                 int N, int C, int H, int W, int layout, int nearest, int border, int align) {
  const int hw = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= hw) return;
  const int px = p % W, py = p / W;
  float fx, fy, ix, iy;
  load_flow(flow, layout, n, hw, p, &fx, &fy);
  source_index(px, py, fx, fy, H, W, align, border, &ix, &iy);
  const float* xb = x + static_cast<long long>(n) * C * hw;
  float* ob = out + static_cast<long long>(n) * C * hw;
  if (nearest) {
    const int sx = (int)nearbyintf(ix), sy = (int)nearbyintf(iy);
    const bool in = sx >= 0 && sx < W && sy >= 0 && sy < H;
    for (int c = 0; c < C; ++c) ob[c * hw + p] = in ? xb[c * hw + sy * W + sx] : 0.f;
  } else {
    const Corner cr = bilinear_corner(ix, iy);
    for (int c = 0; c < C; ++c) ob[c * hw + p] = sample_bilinear(xb + c * hw, H, W, cr);
  }
}

__global__ void flow_warp_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ flow,
                                     float* __restrict__ gin, int N, int C, int H, int W, int layout, int border,
                                     int align) {
  const int hw = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= hw) return;
  float fx, fy, ix, iy;
  load_flow(flow, layout, n, hw, p, &fx, &fy);
  source_index(p % W, p / W, fx, fy, H, W, align, border, &ix, &iy);
  const Corner cr = bilinear_corner(ix, iy);
  for (int c = 0; c < C; ++c)
    scatter_bilinear(gin + (static_cast<long long>(n) * C + c) * hw, H, W, cr,
                     gout[(static_cast<long long>(n) * C + c) * hw + p]);
}

// ---------------------------------------------------------------------------------------------------------------
// occlusion masks: occ_fwd = |f + warp(b, f)| > alpha (|f|+|b|) + beta ; symmetric for bwd (util_flow.py:123-134)
__global__ void fb_consistency_kernel(const float* __restrict__ fwd, const float* __restrict__ bwd,
                                      float* __restrict__ fwd_occ, float* __restrict__ bwd_occ, int B, int H, int W,
                                      float alpha, float beta) {
  const int hw = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = blockIdx.y;
  if (p >= hw) return;
  const int px = p % W, py = p / W;
  const float* f = fwd + static_cast<long long>(n) * 2 * hw;
  const float* b = bwd + static_cast<long long>(n) * 2 * hw;
  const float fx = f[p], fy = f[hw + p], bx = b[p], by = b[hw + p];
  const float mag = __fadd_rn(sqrtf(__fadd_rn(__fmul_rn(fx, fx), __fmul_rn(fy, fy))),
                              sqrtf(__fadd_rn(__fmul_rn(bx, bx), __fmul_rn(by, by))));
  float ix, iy;
  // util_flow.py:76-77 normalises with (w-1); identical to max(w-1,1) for w>1
  source_index(px, py, fx, fy, H, W, 1, 0, &ix, &iy);
  Corner c = bilinear_corner(ix, iy);
  const float wbx = sample_bilinear(b, H, W, c), wby = sample_bilinear(b + hw, H, W, c);
  source_index(px, py, bx, by, H, W, 1, 0, &ix, &iy);
  c = bilinear_corner(ix, iy);
  const float wfx = sample_bilinear(f, H, W, c), wfy = sample_bilinear(f + hw, H, W, c);
  const float dfx = __fadd_rn(fx, wbx), dfy = __fadd_rn(fy, wby);
  const float dbx = __fadd_rn(bx, wfx), dby = __fadd_rn(by, wfy);
  const float diff_f = sqrtf(__fadd_rn(__fmul_rn(dfx, dfx), __fmul_rn(dfy, dfy)));
  const float diff_b = sqrtf(__fadd_rn(__fmul_rn(dbx, dbx), __fmul_rn(dby, dby)));
  const float thr = __fadd_rn(__fmul_rn(alpha, mag), beta);
  fwd_occ[static_cast<long long>(n) * hw + p] = diff_f > thr ? 1.f : 0.f;
  bwd_occ[static_cast<long long>(n) * hw + p] = diff_b > thr ? 1.f : 0.f;
}

// ---------------------------------------------------------------------------------------------------------------
// Motion-guidance gradient.  blockIdx.y enumerates the 2(T-1) L1 terms of compute_temporal_condition_v4:
//   backward pass terms  j in [0,T-1):  j==0 -> zero-term on z_{T-2} with mask 1-fwd_occ[T-2]
//                                       else  i=T-2-j: (1-fwd_occ[i]) * (W(z_{i+1}, flow_bwd_prop[i+1]) - z_i)
//   forward  pass terms  j in [0,T-1):  j==0 -> zero-term on z_1 with mask 1-bwd_occ[0]
//                                       else  k=j+1: (1-bwd_occ[k-1]) * (W(z_{k-1}, flow_fwd_prop[k-2]) - z_k)
// (the off-by-one pairing and the comparison against zeros are the reference's behaviour).
__global__ void motion_guidance_grad_kernel(const float* __restrict__ z, const float* __restrict__ flow_fwd_prop,
                                            const float* __restrict__ flow_bwd_prop,
                                            const float* __restrict__ fwd_occ, const float* __restrict__ bwd_occ,
                                            long long* __restrict__ grad, long long* __restrict__ loss, int T,
                                            int C, int H, int W) {
  const int hw = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int term = blockIdx.y;
  float local = 0.f;
  if (p < hw) {
    const bool bwd_pass = term < (T - 1);
    const int j = bwd_pass ? term : term - (T - 1);
    int fa = -1, fb, fl = 0;  // frame warped (a), frame compared (b), flow index
    float m;
    const float* flow = nullptr;
    if (bwd_pass) {
      if (j == 0) { fb = T - 2; m = 1.f - fwd_occ[(T - 2) * hw + p]; }
      else { const int i = T - 2 - j; fa = i + 1; fb = i; fl = i + 1; m = 1.f - fwd_occ[i * hw + p]; flow = flow_bwd_prop; }
    } else {
      if (j == 0) { fb = 1; m = 1.f - bwd_occ[p]; }
      else { const int k = j + 1; fa = k - 1; fb = k; fl = k - 2; m = 1.f - bwd_occ[(k - 1) * hw + p]; flow = flow_fwd_prop; }
    }
    const float invN = 1.f / (float)(C * hw);
    Corner cr;
    if (fa >= 0) {
      float ix, iy;
      const float fx = flow[(fl * 2) * hw + p], fy = flow[(fl * 2 + 1) * hw + p];
      source_index(p % W, p / W, fx, fy, H, W, 1, 0, &ix, &iy);
      cr = bilinear_corner(ix, iy);
    }
    for (int c = 0; c < C; ++c) {
      const float prev = fa >= 0 ? sample_bilinear(z + (fa * C + c) * hw, H, W, cr) : 0.f;
      const float cur = z[(fb * C + c) * hw + p];
      const float d = __fadd_rn(__fmul_rn(m, prev), -__fmul_rn(m, cur));
      const float s = (d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f);
      local += fabsf(d);
      const float g = m * s * invN;
      if (g != 0.f) {
        fix_add(grad + (fb * C + c) * hw + p, -g);
        if (fa >= 0) scatter_bilinear_fix(grad + (fa * C + c) * hw, H, W, cr, g);
      }
    }
  }
  if (loss) {
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0 && local != 0.f) fix_add(loss, local / (float)(C * hw));
  }
}

// out = z - step * grad; the fixed-point sums are rounded to fp32 once here.  grad_f32 (optional) receives the gradient,
// loss_f32 (optional) the loss; both alias the head of the workspace, which is why thread 0 converts the loss last.
__global__ void axpy_update_kernel(const float* __restrict__ z, const long long* __restrict__ grad,
                                   float* __restrict__ out, float* __restrict__ grad_f32, float step, int n,
                                   const long long* __restrict__ loss_fix, float* __restrict__ loss_f32,
                                   const float* __restrict__ step_table, const int* __restrict__ step_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (step_table) step = step_table[*step_idx];    // graph-replayable form: the per-step scalar lives in device memory
  if (i < n) {
    const float g = (float)((double)grad[i] * (1.0 / kFixScale));
    out[i] = __fadd_rn(z[i], -__fmul_rn(step, g));
    if (grad_f32) grad_f32[i] = g;
  }
  if (i == 0 && loss_f32) *loss_f32 = (float)((double)*loss_fix * (1.0 / kFixScale));
}

// ---------------------------------------------------------------------------------------------------------------
// F.interpolate(bilinear, align_corners=False) of a (N,2,h,w) flow with the value scaling of resize_flow
__global__ void resize_flow_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int h, int w, int oh,
                                   int ow, float ratio_h, float ratio_w) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int nc = blockIdx.y;  // n*2 + channel
  if (p >= oh * ow) return;
  const int ox = p % ow, oy = p / ow;
  const float sh = (float)h / (float)oh, sw = (float)w / (float)ow;
  float sy = __fadd_rn(__fmul_rn(sh, __fadd_rn((float)oy, 0.5f)), -0.5f);
  float sx = __fadd_rn(__fmul_rn(sw, __fadd_rn((float)ox, 0.5f)), -0.5f);
  sy = fmaxf(sy, 0.f); sx = fmaxf(sx, 0.f);
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float hy = 1.f - ly, hx = 1.f - lx;
  const float sc = (nc & 1) ? ratio_h : ratio_w;
  const float* ib = in + static_cast<long long>(nc) * h * w;
  const float v00 = __fmul_rn(ib[y0 * w + x0], sc), v01 = __fmul_rn(ib[y0 * w + x1], sc);
  const float v10 = __fmul_rn(ib[y1 * w + x0], sc), v11 = __fmul_rn(ib[y1 * w + x1], sc);
  out[static_cast<long long>(nc) * oh * ow + p] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
}

// ---------------------------------------------------------------------------------------------------------------
// Canvas tail: eps = sum_tiles(eps_tile * w) / sum_tiles(w);  x0 = c_recip*x - c_recipm1*eps;
// mean = c1*x0 + c2*x;  out = mean + sigma*noise         (ddpm.py:4275-4316, 4404-4417; sigma = 0 at i == 0)
struct TileList {
  int n;
  int ox[64];
  int oy[64];
};
__global__ void canvas_posterior_kernel(const float* __restrict__ x, const float* const* __restrict__ eps_tiles,
                                        const double* __restrict__ tile_w, const float* __restrict__ noise,
                                        float* __restrict__ out, float* __restrict__ eps_out, TileList tl, int TC,
                                        int H, int W, int ts, float c_recip, float c_recipm1, float c1, float c2,
                                        float sigma, const float* __restrict__ coef_table,
                                        const int* __restrict__ step_idx, long long noise_step_stride, int noise_rep) {
  const int hw = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int tc = blockIdx.y;
  if (p >= hw) return;
  int tc_noise = tc;
  if (coef_table) {   // graph-replayable form: per-step scalars and this step's noise slice are picked on the device
    const int st = *step_idx;
    const float* c = coef_table + 5 * st;
    c_recip = c[0]; c_recipm1 = c[1]; c1 = c[2]; c2 = c[3]; sigma = c[4];
    noise += static_cast<long long>(st) * noise_step_stride;
    tc_noise = tc % noise_rep;     // clips batched in one canvas share one draw (the script re-seeds per clip, :428)
  }
  const int px = p % W, py = p / W;
  float acc = 0.f, cnt = 0.f;
  for (int i = 0; i < tl.n; ++i) {
    const int lx = px - tl.ox[i], ly = py - tl.oy[i];
    if (lx >= 0 && lx < ts && ly >= 0 && ly < ts) {
      // the reference accumulates fp32 += fp32 * fp64 weights (ddpm.py:4297-4298): type promotion makes each update
      // a double-precision multiply-add rounded back to fp32
      const double wgt = tile_w[ly * ts + lx];
      acc = (float)((double)acc + (double)eps_tiles[i][(tc * ts + ly) * ts + lx] * wgt);
      cnt = (float)((double)cnt + wgt);
    }
  }
  const float eps = __fdiv_rn(acc, cnt);
  const long long idx = static_cast<long long>(tc) * hw + p;
  const float xv = x[idx];
  const float x0 = __fadd_rn(__fmul_rn(c_recip, xv), -__fmul_rn(c_recipm1, eps));
  const float mean = __fadd_rn(__fmul_rn(c1, x0), __fmul_rn(c2, xv));
  if (eps_out) eps_out[idx] = eps;
  out[idx] = noise ? __fadd_rn(mean, __fmul_rn(sigma, noise[static_cast<long long>(tc_noise) * hw + p])) : mean;
}

}  // namespace mgld

using namespace mgld;

static inline dim3 pix_grid(int hw, int y) { return dim3((hw + 255) / 256, y, 1); }

extern "C" int mgld_flow_warp_f32(const float* x, const float* flow, float* out, int n, int c, int h, int w,
                                  int flow_layout, int interp_nearest, int padding_border, int align_corners,
                                  void* stream) {
  MGLD_CHECK_ARG(x && flow && out && n > 0 && c > 0 && h > 0 && w > 0, "flow_warp: bad arguments");
  flow_warp_kernel<<<pix_grid(h * w, n), 256, 0, (cudaStream_t)stream>>>(x, flow, out, n, c, h, w, flow_layout,
                                                                          interp_nearest, padding_border,
                                                                          align_corners);
  MGLD_LAUNCH_CHECK("flow_warp_kernel");
  return MGLD_OK;
}

extern "C" int mgld_flow_warp_bwd_input_f32(const float* grad_out, const float* flow, float* grad_in, int n, int c,
                                            int h, int w, int flow_layout, int padding_border, int align_corners,
                                            void* stream) {
  MGLD_CHECK_ARG(grad_out && flow && grad_in && n > 0 && c > 0 && h > 0 && w > 0, "flow_warp_bwd: bad arguments");
  MGLD_CUDA(cudaMemsetAsync(grad_in, 0, sizeof(float) * (size_t)n * c * h * w, (cudaStream_t)stream));
  flow_warp_bwd_kernel<<<pix_grid(h * w, n), 256, 0, (cudaStream_t)stream>>>(grad_out, flow, grad_in, n, c, h, w,
                                                                              flow_layout, padding_border,
                                                                              align_corners);
  MGLD_LAUNCH_CHECK("flow_warp_bwd_kernel");
  return MGLD_OK;
}

extern "C" int mgld_fb_consistency_f32(const float* fwd_flow, const float* bwd_flow, float* fwd_occ, float* bwd_occ,
                                       int b, int h, int w, float alpha, float beta, void* stream) {
  MGLD_CHECK_ARG(fwd_flow && bwd_flow && fwd_occ && bwd_occ && b > 0 && h > 1 && w > 1, "fb_consistency: bad arguments");
  fb_consistency_kernel<<<pix_grid(h * w, b), 256, 0, (cudaStream_t)stream>>>(fwd_flow, bwd_flow, fwd_occ, bwd_occ, b,
                                                                               h, w, alpha, beta);
  MGLD_LAUNCH_CHECK("fb_consistency_kernel");
  return MGLD_OK;
}

extern "C" int mgld_motion_guidance_f32(const float* latents, const float* flow_fwd_prop, const float* flow_bwd_prop,
                                        const float* fwd_occ, const float* bwd_occ, void* grad_ws, float* out,
                                        float* grad_out, float* loss, float step, int t, int c, int h, int w,
                                        void* stream) {
  MGLD_CHECK_ARG(latents && grad_ws && out && t >= 1 && c > 0 && h > 0 && w > 0, "motion_guidance: bad arguments");
  MGLD_CHECK_ARG(t == 1 || (flow_fwd_prop && flow_bwd_prop && fwd_occ && bwd_occ), "motion_guidance: null flows");
  MGLD_CHECK_ARG((reinterpret_cast<uintptr_t>(grad_ws) & 7) == 0, "motion_guidance: workspace must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = t * c * h * w;
  long long* acc = static_cast<long long*>(grad_ws);          // n gradient sums + 1 loss sum, 2^-44 fixed point
  MGLD_CUDA(cudaMemsetAsync(acc, 0, sizeof(long long) * ((size_t)n + 1), s));
  if (t >= 2) {
    motion_guidance_grad_kernel<<<pix_grid(h * w, 2 * (t - 1)), 256, 0, s>>>(latents, flow_fwd_prop, flow_bwd_prop,
                                                                              fwd_occ, bwd_occ, acc, loss ? acc + n : nullptr,
                                                                              t, c, h, w);
    MGLD_LAUNCH_CHECK("motion_guidance_grad_kernel");
  }
  axpy_update_kernel<<<(n + 255) / 256, 256, 0, s>>>(latents, acc, out, grad_out, step, n, acc + n, loss, nullptr, nullptr);
  MGLD_LAUNCH_CHECK("axpy_update_kernel");
  return MGLD_OK;
}

extern "C" int mgld_motion_guidance_dev_f32(const float* latents, const float* flow_fwd_prop, const float* flow_bwd_prop,
                                            const float* fwd_occ, const float* bwd_occ, void* grad_ws, float* out,
                                            const float* step_table, const int* step_idx, int t, int c, int h, int w,
                                            void* stream) {
  MGLD_CHECK_ARG(latents && grad_ws && out && step_table && step_idx && t >= 2 && c > 0 && h > 0 && w > 0,
                 "motion_guidance_dev: bad arguments");
  MGLD_CHECK_ARG(flow_fwd_prop && flow_bwd_prop && fwd_occ && bwd_occ, "motion_guidance_dev: null flows");
  MGLD_CHECK_ARG((reinterpret_cast<uintptr_t>(grad_ws) & 7) == 0, "motion_guidance_dev: workspace must be 8-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int n = t * c * h * w;
  long long* acc = static_cast<long long*>(grad_ws);
  MGLD_CUDA(cudaMemsetAsync(acc, 0, sizeof(long long) * ((size_t)n + 1), s));
  motion_guidance_grad_kernel<<<pix_grid(h * w, 2 * (t - 1)), 256, 0, s>>>(latents, flow_fwd_prop, flow_bwd_prop, fwd_occ,
                                                                            bwd_occ, acc, nullptr, t, c, h, w);
  MGLD_LAUNCH_CHECK("motion_guidance_grad_kernel");
  axpy_update_kernel<<<(n + 255) / 256, 256, 0, s>>>(latents, acc, out, nullptr, 0.f, n, acc + n, nullptr, step_table,
                                                     step_idx);
  MGLD_LAUNCH_CHECK("axpy_update_kernel");
  return MGLD_OK;
}

extern "C" int mgld_resize_flow_f32(const float* flow, float* out, int n, int h, int w, int oh, int ow, void* stream) {
  MGLD_CHECK_ARG(flow && out && n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0, "resize_flow: bad arguments");
  resize_flow_kernel<<<pix_grid(oh * ow, n * 2), 256, 0, (cudaStream_t)stream>>>(
      flow, out, n, h, w, oh, ow, (float)((double)oh / (double)h), (float)((double)ow / (double)w));
  MGLD_LAUNCH_CHECK("resize_flow_kernel");
  return MGLD_OK;
}

extern "C" int mgld_canvas_posterior_f32(const float* x, const float* const* eps_tiles_dev, const double* tile_w,
                                         const float* noise, float* out, float* eps_out, int n_tiles,
                                         const int* ofs_x, const int* ofs_y, int tc, int h, int w, int tile_size,
                                         float c_recip, float c_recipm1, float c1, float c2, float sigma,
                                         void* stream) {
  MGLD_CHECK_ARG(x && eps_tiles_dev && tile_w && out && n_tiles > 0 && n_tiles <= 64, "canvas_posterior: bad arguments");
  TileList tl;
  tl.n = n_tiles;
  for (int i = 0; i < n_tiles; ++i) { tl.ox[i] = ofs_x[i]; tl.oy[i] = ofs_y[i]; }
  canvas_posterior_kernel<<<pix_grid(h * w, tc), 256, 0, (cudaStream_t)stream>>>(
      x, eps_tiles_dev, tile_w, noise, out, eps_out, tl, tc, h, w, tile_size, c_recip, c_recipm1, c1, c2, sigma, nullptr,
      nullptr, 0, tc);
  MGLD_LAUNCH_CHECK("canvas_posterior_kernel");
  return MGLD_OK;
}

extern "C" int mgld_canvas_posterior_dev_f32(const float* x, const float* const* eps_tiles_dev, const double* tile_w,
                                             const float* noise_all, long long noise_step_stride, int noise_tc,
                                             float* out, int n_tiles, const int* ofs_x, const int* ofs_y, int tc, int h,
                                             int w, int tile_size, const float* coef_table, const int* step_idx,
                                             void* stream) {
  MGLD_CHECK_ARG(x && eps_tiles_dev && tile_w && out && noise_all && coef_table && step_idx && n_tiles > 0 && n_tiles <= 64 &&
                     noise_tc > 0 && tc % noise_tc == 0, "canvas_posterior_dev: bad arguments");
  TileList tl;
  tl.n = n_tiles;
  for (int i = 0; i < n_tiles; ++i) { tl.ox[i] = ofs_x[i]; tl.oy[i] = ofs_y[i]; }
  canvas_posterior_kernel<<<pix_grid(h * w, tc), 256, 0, (cudaStream_t)stream>>>(
      x, eps_tiles_dev, tile_w, noise_all, out, nullptr, tl, tc, h, w, tile_size, 0.f, 0.f, 0.f, 0.f, 0.f, coef_table,
      step_idx, noise_step_stride, noise_tc);
  MGLD_LAUNCH_CHECK("canvas_posterior_kernel");
  return MGLD_OK;
}
