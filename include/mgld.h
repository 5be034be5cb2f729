/*
 * mgld.h — C ABI of libmgld.so: the sm_100a kernels behind the MGLD-VSR hot path.
 *
 * The reference (IanYeung/MGLD-VSR) has no native code and no FFI of its own: its operator boundary is the set of
 * PyTorch library calls listed in SURVEY.md §2.1 / §8(b).4.  Every entry point below replaces one of those call sites
 * (cited per function as reference file:line, relative to the reference repository root).
 *
 * Conventions
 *   - plain C: raw device pointers, explicit sizes, POD descriptor structs; no torch types, no exceptions.
 *   - every function is asynchronous on the given stream (a cudaStream_t passed as void*), never synchronises, never
 *     allocates device memory and never takes ownership of a buffer.
 *   - return value: 0 = ok, negative = error (see mgld_last_error()).
 *   - activations are fp16 "NHWC": [T, H, W, C] row-major, which is also the token layout [T, H*W, C].
 *   - weights are fp16, K-major: [N, K] row-major with K = taps * C (tap-major, channel fastest).
 */
#ifndef MGLD_H_
#define MGLD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MGLD_ABI_VERSION 1

int mgld_abi_version(void);
/* Initialise per-device state (driver entry points, shared-memory opt-in).  Must be called once per process. */
int mgld_init(int device);
const char* mgld_last_error(void);

/* ------------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution / linear layer on tcgen05 tensor cores (TMA-fed, TMEM accumulators).
 *
 *   out[m, n] = epilogue( sum_{tap, c} A[pixel(m) + offset(tap), c] * W[n, tap*C + c] )
 *
 * Replaces: F.conv2d 3x3/1x1 (openaimodel.py:401-445, spade.py:83-88, model.py:134-161), nn.Linear
 * (attention.py:48-75,510-524; openaimodel.py:2021-2025,418-424), Conv3d (3,1,1) (util.py:291-310).
 * ------------------------------------------------------------------------------------------------------------------ */
enum mgld_taps { MGLD_TAPS_1 = 1, MGLD_TAPS_T3 = 3, MGLD_TAPS_3X3 = 9 };
enum mgld_epilogue {
  MGLD_EPI_LINEAR = 0, /* v = act(acc + bias[n]);           out = alpha*v + beta*res[m,n]                        */
  MGLD_EPI_GEGLU = 1,  /* W rows interleaved per 128: 64 value | 64 gate;  out = (acc_v+b_v) * gelu(acc_g+b_g)    */
  MGLD_EPI_SPADE = 2   /* W rows interleaved per 128: 64 gamma | 64 beta;
                          out = beta_res*res + GNaffine(h)[m,c] * (1 + gamma) + beta   (spade.py:100-109)        */
};
enum mgld_act { MGLD_ACT_NONE = 0, MGLD_ACT_RELU = 1, MGLD_ACT_SILU = 2, MGLD_ACT_LRELU02 = 3, MGLD_ACT_GELU = 4 };

typedef struct mgld_conv_gemm_desc {
  /* A operand: one or two NHWC fp16 tensors sharing (T,H,W); channels of a2 follow those of a (fused concat).      */
  const void* a;
  const void* a2;
  int32_t T, H, W;
  int32_t C1, C2; /* channels of a / a2 (C2 = 0 when a2 is NULL); both multiples of 64                                */
  int32_t lda, lda2; /* row pitch (elements) of a / a2; 0 = dense (= C1 / C2)                                         */
  /* B operand: packed weights [N, taps*(C1+C2)] fp16                                                                 */
  const void* w;
  int32_t N;       /* rows of w (for pair epilogues this is 2x the number of output columns)                           */
  int32_t taps;    /* enum mgld_taps                                                                                   */
  int32_t block_n; /* N tile (multiple of 16, <= 256; 128 for the pair epilogues); 0 = let the library choose          */
  /* epilogue                                                                                                          */
  int32_t epilogue; /* enum mgld_epilogue                                                                              */
  int32_t act;      /* enum mgld_act (LINEAR only)                                                                     */
  const float* bias; /* fp32 [N] or NULL                                                                               */
  float alpha, beta;
  const void* res; /* fp16 [M, ldres] or NULL                                                                          */
  int32_t ldres;
  /* SPADE extras: h fp16 [M, ldh] with n_out channels, stats fp32 [T, groups, 2] = (mean, rstd), GN affine fp32 [n_out] */
  const void* h;
  int32_t ldh;
  const float* gn_stats;
  const float* gn_weight;
  const float* gn_bias;
  int32_t groups;
  /* output: fp16 (or fp32 when out_f32) [M, ldout], written at column offset out_col0                                 */
  void* out;
  int32_t ldout;
  int32_t out_col0;
  int32_t out_f32;
} mgld_conv_gemm_desc;

int mgld_conv_gemm(const mgld_conv_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Fused softmax attention on tcgen05:  out = softmax(scale * Q K^T) V     (fp16 in/out, fp32 accumulate)
 *
 * Q, K, V are fp16 matrices [batch * n, ld] whose head h occupies columns [col0 + h*head_stride, +head_dim).
 * Replaces xformers.ops.memory_efficient_attention at attention.py:298 (self / temporal), attention.py:371 (cross;
 * kv_batched = 0 broadcasts the single context to every frame, as attention.py:336-337 does with repeat_interleave)
 * and QKVAttentionLegacy, openaimodel.py:554-590.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct mgld_attention_desc {
  const void* q; const void* k; const void* v;
  int32_t ldq, ldk, ldv;
  int32_t q_col0, k_col0, v_col0;
  int32_t q_head_stride, k_head_stride, v_head_stride;
  int32_t batch, heads, head_dim; /* head_dim 64 or 128 */
  int32_t nq, nkv;
  int32_t kv_batched;
  float scale;
  void* out; /* fp16 [batch*nq, ldo], head h at columns [h*head_dim, +head_dim) */
  int32_t ldo;
} mgld_attention_desc;

int mgld_attention(const mgld_attention_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Flow-guided latent ops (fp32, NCHW: the layout the reference keeps latents / flows in)
 * ------------------------------------------------------------------------------------------------------------------ */
/* basicsr/archs/arch_util.py:156 flow_warp (flow_layout 0: (n,h,w,2)) and scripts/util_flow.py:97 flow_warp
 * (flow_layout 1: (n,2,h,w)).  F.grid_sample semantics: bilinear|nearest, zeros|border padding.                     */
int mgld_flow_warp_f32(const float* x, const float* flow, float* out, int n, int c, int h, int w, int flow_layout,
                       int interp_nearest, int padding_border, int align_corners, void* stream);
/* adjoint of the bilinear warp w.r.t. x (what autograd computes for ddpm.py:4434)                                    */
int mgld_flow_warp_bwd_input_f32(const float* grad_out, const float* flow, float* grad_in, int n, int c, int h, int w,
                                 int flow_layout, int padding_border, int align_corners, void* stream);
/* scripts/util_flow.py:114 forward_backward_consistency_check: flows (b,2,h,w) -> float {0,1} masks (b,h,w)          */
int mgld_fb_consistency_f32(const float* fwd_flow, const float* bwd_flow, float* fwd_occ, float* bwd_occ, int b, int h,
                            int w, float alpha, float beta, void* stream);
/* ddpm.py:3538 compute_temporal_condition_v4 + the update of ddpm.py:4429-4435 in one call:
 *   out = latents - step * d(loss_b + loss_f)/d(latents),   step = guidance_scale * model_log_variance.
 * latents (t,c,h,w); flows (t-1,2,h,w); occlusion masks (t-1,h,w); grad_ws: t*c*h*w floats of workspace;
 * loss (optional, 1 float) receives loss_b + loss_f.                                                                 */
int mgld_motion_guidance_f32(const float* latents, const float* flow_fwd_prop, const float* flow_bwd_prop,
                             const float* fwd_occ, const float* bwd_occ, float* grad_ws, float* out, float* loss,
                             float step, int t, int c, int h, int w, void* stream);
/* basicsr/archs/arch_util.py:235 resize_flow (bilinear, align_corners=False, values scaled by the size ratio)        */
int mgld_resize_flow_f32(const float* flow, float* out, int n, int h, int w, int oh, int ow, void* stream);
/* ddpm.py:4275-4316 + 4404-4417: Gaussian-weighted stitch of eps tiles, x0, posterior mean, noise add.
 * eps_tiles_dev: device array of n_tiles pointers to (tc, tile, tile) fp32 tiles; ofs_x/ofs_y: host arrays.          */
int mgld_canvas_posterior_f32(const float* x, const float* const* eps_tiles_dev, const float* tile_w,
                              const float* noise, float* out, float* eps_out, int n_tiles, const int* ofs_x,
                              const int* ofs_y, int tc, int h, int w, int tile_size, float c_recip, float c_recipm1,
                              float c1, float c2, float sigma, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGLD_H_ */
