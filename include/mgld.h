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

#define MGLD_ABI_VERSION 2

int mgld_abi_version(void);
/* Initialise per-device state (driver entry points, shared-memory opt-in).  Must be called once per process; the process
 * is then bound to `device` (one process per GPU, as under torchrun): a later call with another device returns an error. */
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
enum mgld_taps { MGLD_TAPS_1 = 1, MGLD_TAPS_T3 = 3 /* (3,1,1) in time */, MGLD_TAPS_3X3 = 9,
                 MGLD_TAPS_1X5 = 5 /* RAFT SepConvGRU (1,5) */, MGLD_TAPS_5X1 = 6 /* (5,1): 5 taps along H */ };
enum mgld_epilogue {
  MGLD_EPI_LINEAR = 0, /* v = act(acc + bias[n]);           out = alpha*v + beta*res[m,n]                        */
  MGLD_EPI_GEGLU = 1,  /* W rows interleaved per 128: 64 value | 64 gate;  out = (acc_v+b_v) * gelu(acc_g+b_g)    */
  MGLD_EPI_SPADE = 2   /* W rows interleaved per 128: 64 gamma | 64 beta;
                          out = beta_res*res + GNaffine(h)[m,c] * (1 + gamma) + beta   (spade.py:100-109)        */
};
enum mgld_act { MGLD_ACT_NONE = 0, MGLD_ACT_RELU = 1, MGLD_ACT_SILU = 2, MGLD_ACT_LRELU02 = 3, MGLD_ACT_GELU = 4,
                MGLD_ACT_SIGMOID = 5, MGLD_ACT_TANH = 6 };

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
  /* optional fused GroupNorm statistics of the (fp16-rounded) output, for the GroupNorm that consumes it:
     stats_out[t][g] += (sum, sum of squares) over frame t and channel group g of `stats_groups` equal groups;
     fp64 [T, stats_groups, 2], zeroed by the caller.  Same layout mgld_gn_stats_f16 produces.                          */
  double* stats_out;
  int32_t stats_groups;
  /* optional scratch for the split-K path (layers with too few output tiles to fill the SMs: partial fp32 tiles are
     written per K-slice and reduced in a fixed order by a second kernel).  Size: mgld_conv_gemm_workspace_bytes().
     NULL / too small = single-pass execution.                                                                        */
  void* workspace;
  int64_t workspace_bytes;
} mgld_conv_gemm_desc;

int mgld_conv_gemm(const mgld_conv_gemm_desc* d, void* stream);
/* bytes of workspace with which mgld_conv_gemm would use split-K for this problem (0 = it would not)                   */
long long mgld_conv_gemm_workspace_bytes(const mgld_conv_gemm_desc* d);
/* Development hook: per-CTA cycle counters ([grid][16] int64, device memory) filled by the following conv_gemm launches;
   null switches it off (the default). */
void mgld_conv_gemm_set_debug_counters(void* dev_ptr);

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
  int32_t batch, heads, head_dim; /* head_dim 64, 128 or 512 (512: the VAE middle attention, model.py:247-305) */
  int32_t nq, nkv;
  int32_t kv_batched;
  float scale;
  void* out; /* fp16 [batch*nq, ldo], head h at columns [h*head_dim, +head_dim) */
  int32_t ldo;
} mgld_attention_desc;

int mgld_attention(const mgld_attention_desc* d, void* stream);
/* Development hook: per-CTA cycle counters ([CTAs][16] int64, device memory) filled by the following head-dim-64 attention
   launches (which role waits for which); null switches it off (the default). */
void mgld_attention_set_debug_counters(void* dev_ptr);

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
 * latents (t,c,h,w); flows (t-1,2,h,w); occlusion masks (t-1,h,w); grad_ws: 8*(t*c*h*w + 1) bytes of 8-byte aligned
 * workspace (64-bit fixed-point sums: the result is bitwise repeatable); grad_out (optional, t*c*h*w floats) receives
 * the gradient, loss (optional, 1 float) loss_b + loss_f.                                                            */
int mgld_motion_guidance_f32(const float* latents, const float* flow_fwd_prop, const float* flow_bwd_prop,
                             const float* fwd_occ, const float* bwd_occ, void* grad_ws, float* out, float* grad_out,
                             float* loss, float step, int t, int c, int h, int w, void* stream);
/* Graph-replayable form of mgld_motion_guidance_f32: the step scalar is read on the device as step_table[*step_idx], so
 * one captured launch sequence serves every DDPM step.                                                                */
int mgld_motion_guidance_dev_f32(const float* latents, const float* flow_fwd_prop, const float* flow_bwd_prop,
                                 const float* fwd_occ, const float* bwd_occ, void* grad_ws, float* out,
                                 const float* step_table, const int* step_idx, int t, int c, int h, int w, void* stream);
/* basicsr/archs/arch_util.py:235 resize_flow (bilinear, align_corners=False, values scaled by the size ratio)        */
int mgld_resize_flow_f32(const float* flow, float* out, int n, int h, int w, int oh, int ow, void* stream);
/* ddpm.py:4275-4316 + 4404-4417: Gaussian-weighted stitch of eps tiles, x0, posterior mean, noise add.
 * eps_tiles_dev: device array of n_tiles pointers to (tc, tile, tile) fp32 tiles; tile_w: (tile, tile) fp64 weights
 * (the reference keeps them in float64, ddpm.py:4610-4616); ofs_x/ofs_y: host arrays.                                */
int mgld_canvas_posterior_f32(const float* x, const float* const* eps_tiles_dev, const double* tile_w,
                              const float* noise, float* out, float* eps_out, int n_tiles, const int* ofs_x,
                              const int* ofs_y, int tc, int h, int w, int tile_size, float c_recip, float c_recipm1,
                              float c1, float c2, float sigma, void* stream);
/* Graph-replayable form: the five per-step scalars are read on the device from coef_table[5 * *step_idx + {0..4}]
 * (c_recip, c_recipm1, c1, c2, sigma), the noise of this step from noise_all + *step_idx * noise_step_stride, laid out
 * (noise_tc, h, w) and shared by the tc / noise_tc clips batched in the canvas (they are re-seeded identically).        */
int mgld_canvas_posterior_dev_f32(const float* x, const float* const* eps_tiles_dev, const double* tile_w,
                                  const float* noise_all, long long noise_step_stride, int noise_tc, float* out,
                                  int n_tiles, const int* ofs_x, const int* ofs_y, int tc, int h, int w, int tile_size,
                                  const float* coef_table, const int* step_idx, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Normalisation (fp16 NHWC activations, fp32 math)
 * ------------------------------------------------------------------------------------------------------------------ */
/* GroupNorm statistics over the virtual concat [x1 | x2] (x2 may be NULL): sums[t][g] += (sum, sum of squares).  `sums` is
 * an opaque array of [T, groups, 2] ACCUMULATORS OF 16 BYTES each (integer part + 2^-40 fixed-point fraction, summed with
 * integer atomics: the statistics are bitwise repeatable), i.e. T*groups*4 doubles' worth of memory, zeroed by the caller.
 * GroupNorm32 util.py:199-216 / Normalize model.py:80.                                                                */
int mgld_gn_stats_f16(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int T, int HW, int groups,
                      double* sums, void* stream);
/* (sum, sumsq) -> fp32 (mean, rstd) pairs, for the SPADE epilogue of mgld_conv_gemm                                  */
int mgld_gn_finalize(const double* sums, float* stats, int T, int groups, int HW, int C, double eps, void* stream);
/* y = [silu]( (x - mean) * rstd * gamma + beta ), written as one dense [T*HW, C1+C2] tensor                           */
int mgld_gn_apply_f16(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int T, int HW, int groups,
                      const double* sums, double eps, const float* gamma, const float* beta, int silu, void* out,
                      int ldo, void* stream);
/* GroupNorm [+SiLU] in ONE launch where the per-CTA patch fits in registers (UNet / struct-encoder shapes): statistics and
 * normalisation share a single read of the activation; CTAs of a frame reduce over distributed shared memory (thread-block
 * clusters).  out (dense [T*HW, C1+C2], row pitch ldo) and stats_out ((mean, rstd) fp32 [T, groups, 2], the SPADE
 * epilogue's operand) are each optional.  Shapes that do not fit (mgld_group_norm_fused_supported() == 0, e.g. the VAE's
 * 256^2+ maps) run gn_stats / gn_finalize / gn_apply internally and need `scratch`: [T, groups, 2] doubles, zeroed.
 * GroupNorm32 util.py:199-216 (eps 1e-5), Normalize model.py:80 (eps 1e-6).                                             */
int mgld_group_norm_f16(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int T, int HW, int groups,
                        double eps, const float* gamma, const float* beta, int silu, void* out, int ldo,
                        float* stats_out, double* scratch, void* stream);
int mgld_group_norm_fused_supported(int C, int T, int HW, int groups);
/* nn.LayerNorm over the last dim (attention.py:132,423-425)                                                          */
int mgld_layernorm_f16(const void* x, int ldx, int M, int C, const float* gamma, const float* beta, float eps,
                       void* out, int ldo, void* stream);
/* P = softmax(scale * S) row-wise, fp32 scores -> fp16 probabilities (VAE mid attention, model.py:294)               */
int mgld_softmax_rows_f32(const float* s, long long lds, int rows, int n, float scale, void* p, long long ldp,
                          void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Layout / resampling / stems / small ops
 * ------------------------------------------------------------------------------------------------------------------ */
int mgld_nchw_f32_to_nhwc_f16(const float* in, void* out, int n, int c, int h, int w, int ldo, float scale, void* stream);
int mgld_nhwc_f16_to_nchw_f32(const void* in, float* out, int n, int c, int h, int w, int ldi, float scale, void* stream);
/* F.interpolate(scale_factor=2, mode="nearest"), NHWC fp16 (openaimodel.py:178-188, model.py:95-99)                   */
int mgld_upsample_nearest2x_f16(const void* in, void* out, int t, int h, int w, int c, void* stream);
/* im2col of a 3x3 stride-2 conv: pad=1 (openaimodel.py:204-231) or pad=0 with far-side zero fill (model.py:104-121)   */
int mgld_im2col_s2_f16(const void* in, void* out, int t, int h, int w, int c, int ho, int wo, int pad, void* stream);
/* direct conv, Cin <= 8: (N,Cin,H,W) fp32 -> NHWC fp16; w fp32 [Cout,Cin,ks,ks] (network stems)                        */
int mgld_conv_small_cin_f32(const float* in, const float* w, const float* bias, void* out, int n, int cin, int h,
                            int wd, int cout, int ks, int ldo, void* stream);
/* direct conv fp32 NCHW -> NCHW, tiny channel counts (quant_conv / post_quant_conv, autoencoder.py:331-332)            */
int mgld_conv_small_f32(const float* in, const float* w, const float* bias, float* out, int n, int cin, int h, int wd,
                        int cout, int ks, void* stream);
/* direct 3x3 conv, Cout in {2,3,4,8}: NHWC fp16 -> (N,Cout,H,W) fp32; w fp16 [Cout, 9*C] (network heads)               */
int mgld_conv3x3_small_cout_f16(const void* in, const void* w, const float* bias, float* out, int n, int h, int wd,
                                int c, int cout, int ldi, void* stream);
/* y = [silu](W [silu](x) + bias + add): the M=1 linears of time_embed / emb_layers (openaimodel.py:2021-2025,418-424)  */
int mgld_gemv_f32(const float* x, const void* w, const float* bias, const float* add, float* y, int n, int k,
                  int silu_in, int silu_out, void* stream);
/* timestep_embedding, util.py:151-171; t is read from device memory (1 float) so the call is CUDA-graph replayable    */
int mgld_timestep_embedding_f32(const float* t, float* out, int dim, float max_period, void* stream);
/* TemporalAttention core (attention.py:135-141): softmax over the T frames of each pixel; qkv fp16 [T,HW,3C]           */
int mgld_temporal_attention_f16(const void* qkv, void* out, int t, int hw, int c, int heads, float scale, void* stream);
/* DiagonalGaussianDistribution.sample * scale (distributions.py:24-37, ddpm.py:3382-3390)                              */
int mgld_gaussian_sample_f32(const float* moments, const float* noise, float* out, int n, int cz, int h, int w,
                             float scale, void* stream);
/* out = [relu](a*x + b*y), fp16 */
int mgld_axpby_f16(const void* x, const void* y, void* out, float a, float b, long long n, int relu, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * RAFT optical flow (basicsr/archs/raft_arch.py) — the non-GEMM pieces; its convolutions use mgld_conv_gemm
 * ------------------------------------------------------------------------------------------------------------------ */
/* direct conv for Cin <= 4 (7x7 stems :215, :430): (N,Cin,H,W) fp32 -> NHWC fp16, stride 1|2, optional ReLU          */
int mgld_conv_direct_f32(const float* in, const float* w, const float* bias, void* out, int n, int cin, int h, int wd,
                         int cout, int ks, int stride, int pad, int ldo, int relu, void* stream);
/* every second pixel of an NHWC fp16 tensor (input of the stride-2 1x1 `downsample` convs, :124-126)                  */
int mgld_subsample2_f16(const void* in, void* out, int n, int h, int w, int c, void* stream);
/* nn.InstanceNorm2d (no affine, eps) [+ReLU] from the per-(n,channel) sums of mgld_gn_stats_f16(groups = C)            */
int mgld_instance_norm_apply_f16(const void* x, const double* sums, void* out, int n, int hw, int c, double eps,
                                 int relu, void* stream);
/* F.avg_pool2d(., 2, stride=2) on a stack of fp32 maps (correlation pyramid, :50-52)                                   */
int mgld_avgpool2_f32(const float* in, float* out, long long n, int h, int w, void* stream);
/* CorrBlock.__call__ (:57-81): 4 levels x 9x9 window bilinear lookup -> NHWC fp16 [b, h*w, ldo] (324 channels)          */
int mgld_corr_lookup_f32(const float* l0, const float* l1, const float* l2, const float* l3, const float* coords,
                         void* out, int b, int h, int w, int ldo, void* stream);
/* SepConvGRU gating (:398-412): rnet = r * net;  net = (1 - z) * net + z * q;  zr = [M, 2C] = sigmoid(z | r)            */
int mgld_gru_rh_f16(const void* zr, const void* net, void* rnet, long long m, int c, void* stream);
int mgld_gru_update_f16(const void* zr, const void* q, void* net, long long m, int c, void* stream);
/* copy a (B,Cs,h,w) fp32 tensor into columns [col0, col0+Cs) of an NHWC fp16 buffer (torch.cat([out, flow]), :445)     */
int mgld_set_channels_f16(const float* src, void* dst, int b, int cs, int hw, int ld, int col0, void* stream);
/* RAFT_SR.upsample_flow (:720-731): convex 8x upsampling, mask NHWC fp16 [b,h,w,576], flow (b,2,h,w) -> (b,2,8h,8w)     */
int mgld_convex_upsample8_f32(const void* mask, const float* flow, float* out, int b, int h, int w, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * I/O edges on the GPU; frames stay device-resident between the decoder of the input and the encoder
 * of the output.
 * read_image (script :124-130) + F.interpolate(mode='bicubic') (:349-357) [+ clamp(-1,1) :376, + reflect pad right/bottom
 * :383-387]: in uint8 [n, h, w, 3] -> out fp32 [n, 3, oh + pad_h, ow + pad_w].                                          */
int mgld_frames_u8_to_f32_bicubic(const void* in, float* out, int n, int h, int w, int oh, int ow, int pad_h, int pad_w,
                                  int clamp, void* stream);
/* script :529-541: in fp32 [n, 3, h, w] in [0,1] -> out uint8 [n, crop_h, crop_w, 3] = (uint8)(x * 255), top-left crop     */
int mgld_frames_f32_to_u8_hwc(const float* in, void* out, int n, int h, int w, int crop_h, int crop_w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MGLD_H_ */
