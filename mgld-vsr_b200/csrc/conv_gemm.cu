// Implicit-GEMM convolution / linear layer for sm_100a.
//
//   out[m, n] = epilogue( sum_{tap, c} A[pixel(m) + off(tap), c] * W[n, tap*C + c] )
//
// One CTA computes a 128 x BLOCK_N output tile.  The 128 rows of a tile are a (BW x BH x BT) box of pixels of the
// NHWC activation, so for every filter tap the A operand is ONE TMA box load at shifted coordinates — the TMA unit
// does the im2col, and its out-of-bounds zero fill is the convolution's zero padding (spatial and temporal).
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + tcgen05.mma issuer, warps 2-5 = epilogue
// (TMEM -> registers -> fp32 staging in the drained pipeline smem -> coalesced 16-byte global stores with the fused
// bias / activation / residual / GEGLU / SPADE math).
//
// Replaces (reference file:line): F.conv2d in ResBlockDual openaimodel.py:401-445, SPADE spade.py:83-88,
// VAE ResnetBlock model.py:134-161; nn.Linear in attention.py:48-75,510-524; Conv3d(3,1,1) util.py:291-310.
#include <string.h>

#include "../../include/mgld.h"
#include "common.h"
#include "ptx.cuh"

namespace mgld {

constexpr int kBlockM = 128;
constexpr int kKChunk = 64;  // fp16 elements per K step = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kKChunk * 2;
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;

struct ConvGemmParams {
  int T, H, W;
  int BW, BH, BT;
  int tiles_w, tiles_h, tiles_t;
  int kchunks1, kchunks;  // 64-wide chunks in source 1 / in both sources (per tap)
  int taps;
  int N, block_n, n_stage_cols, n_out_tile, n_out_total;
  int stages, tmem_cols;
  int epilogue, act;
  const float* bias;
  float alpha, beta;
  const __half* res;
  int ldres;
  const __half* h;
  int ldh;
  const float* gn_stats;
  const float* gn_weight;
  const float* gn_bias;
  int groups, ch_per_group;
  void* out;
  int ldout, out_col0, out_f32;
};

__device__ __forceinline__ float apply_act(float v, int act) {
  switch (act) {
    case MGLD_ACT_RELU: return fmaxf(v, 0.f);
    case MGLD_ACT_SILU: return v / (1.f + __expf(-v));
    case MGLD_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
    case MGLD_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

__global__ void __launch_bounds__(kThreads, 2)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t tmem_full_bar;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const int b_bytes = p.block_n * kKChunk * 2;
  const int stage_bytes = kABytes + b_bytes;

  // tile coordinates: blockIdx.x enumerates pixel boxes (w fastest), blockIdx.y the N tiles
  const int tile_m = blockIdx.x;
  const int tw = tile_m % p.tiles_w;
  const int th = (tile_m / p.tiles_w) % p.tiles_h;
  const int tt = tile_m / (p.tiles_w * p.tiles_h);
  const int x0 = tw * p.BW, y0 = th * p.BH, t0 = tt * p.BT;
  const int n0 = blockIdx.y * p.block_n;
  const int num_iters = p.taps * p.kchunks;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.kchunks1 < p.kchunks) tma_prefetch_desc(&tmA2);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&tmem_full_bar), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int it = 0;
      for (int tap = 0; tap < p.taps; ++tap) {
        int dx = 0, dy = 0, dt = 0;
        if (p.taps == 9) { dx = tap % 3 - 1; dy = tap / 3 - 1; }
        else if (p.taps == 3) { dt = tap - 1; }
        for (int kc = 0; kc < p.kchunks; ++kc, ++it) {
          const int s = it % p.stages;
          const uint32_t ph = (it / p.stages) & 1;
          mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1);
          const uint32_t fb = smem_u32(&full_bar[s]);
          const uint32_t sa = smem_base + s * stage_bytes;
          mbar_expect_tx(fb, stage_bytes);
          if (kc < p.kchunks1) tma_load_4d(sa, &tmA, fb, kc * kKChunk, x0 + dx, y0 + dy, t0 + dt);
          else tma_load_4d(sa, &tmA2, fb, (kc - p.kchunks1) * kKChunk, x0 + dx, y0 + dy, t0 + dt);
          tma_load_2d(sa + kABytes, &tmB, fb, it * kKChunk, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(kBlockM, p.block_n, 0, 0);
      for (int it = 0; it < num_iters; ++it) {
        const int s = it % p.stages;
        const uint32_t ph = (it / p.stages) & 1;
        mbar_wait(smem_u32(&full_bar[s]), ph);
        tc_fence_after();
        const uint32_t sa = smem_base + s * stage_bytes;
        const uint64_t adesc = umma_smem_desc(sa, 0, 1024, kSwz128);
        const uint64_t bdesc = umma_smem_desc(sa + kABytes, 0, 1024, kSwz128);
#pragma unroll
        for (int k = 0; k < kKChunk / 16; ++k) {
          // advance 32 bytes (16 fp16) along K inside the 128-byte swizzle row: +2 in the (addr>>4) field
          umma_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (it | k) != 0);
        }
        umma_commit(smem_u32(&empty_bar[s]));
      }
      umma_commit(smem_u32(&tmem_full_bar));
    }
  } else {
    // ============================ epilogue (warps 2..5) ============================
    const int q = warp & 3;           // TMEM lane quadrant this warp may access
    const int row = q * 32 + lane;    // tile row owned in phase 1
    const int P = p.n_stage_cols + 4; // fp32 staging pitch
    float* stage = reinterpret_cast<float*>(smem_gen);
    mbar_wait(smem_u32(&tmem_full_bar), 0);
    tc_fence_after();
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    // ---- phase 1: TMEM -> registers -> (+bias, act / pair op) -> fp32 staging ----
    if (p.epilogue == MGLD_EPI_GEGLU) {
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t rv[16], rg[16];
        tmem_ld_x16(trow + c0, rv);
        tmem_ld_x16(trow + 64 + c0, rg);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c0 + j + u;
            float v = __uint_as_float(rv[j + u]);
            float g = __uint_as_float(rg[j + u]);
            if (p.bias) { v += __ldg(p.bias + n0 + c); g += __ldg(p.bias + n0 + 64 + c); }
            o[u] = v * apply_act(g, MGLD_ACT_GELU);
          }
          *reinterpret_cast<float4*>(stage + row * P + c0 + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    } else {
      const bool lin = (p.epilogue == MGLD_EPI_LINEAR);
      for (int c0 = 0; c0 < p.block_n; c0 += 16) {
        uint32_t r[16];
        tmem_ld_x16(trow + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float o[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int n = n0 + c0 + j + u;
            float v = __uint_as_float(r[j + u]);
            if (p.bias && n < p.N) v += __ldg(p.bias + n);
            if (lin) v = apply_act(v, p.act);
            o[u] = v;
          }
          *reinterpret_cast<float4*>(stage + row * P + c0 + j) = make_float4(o[0], o[1], o[2], o[3]);
        }
      }
    }
    tc_fence_before();
    named_bar_sync(1, 128);

    // ---- phase 2: coalesced write-out with the tensor-valued epilogue terms ----
    const int e = threadIdx.x - 64;
    const int vpr = p.n_out_tile >> 3;  // 8-column vectors per row
    const int nout0 = blockIdx.y * p.n_out_tile;
    const int total = kBlockM * vpr;
    for (int idx = e; idx < total; idx += 128) {
      const int r = idx / vpr;
      const int v8 = idx - r * vpr;
      const int col = nout0 + v8 * 8;
      if (col >= p.n_out_total) continue;
      const int x = x0 + r % p.BW;
      const int y = y0 + (r / p.BW) % p.BH;
      const int t = t0 + r / (p.BW * p.BH);
      if (x >= p.W || y >= p.H || t >= p.T) continue;
      const long long m = (static_cast<long long>(t) * p.H + y) * p.W + x;
      float o[8];
      const float* sp = stage + r * P + v8 * 8;
      const float4 s0 = *reinterpret_cast<const float4*>(sp);
      const float4 s1 = *reinterpret_cast<const float4*>(sp + 4);
      o[0] = s0.x; o[1] = s0.y; o[2] = s0.z; o[3] = s0.w;
      o[4] = s1.x; o[5] = s1.y; o[6] = s1.z; o[7] = s1.w;
      float rs[8];
      const bool has_res = (p.res != nullptr);
      if (has_res) {
        const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + m * p.ldres + col));
        const __half2* hh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float2 f = __half22float2(hh[u]);
          rs[2 * u] = f.x; rs[2 * u + 1] = f.y;
        }
      }
      if (p.epilogue == MGLD_EPI_SPADE) {
        // staged: gamma at [0,64), beta at [64,128) of this tile; o[] currently holds gamma
        const float4 b0 = *reinterpret_cast<const float4*>(sp + 64);
        const float4 b1 = *reinterpret_cast<const float4*>(sp + 68);
        const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const uint4 hr = __ldg(reinterpret_cast<const uint4*>(p.h + m * p.ldh + col));
        const __half2* hh = reinterpret_cast<const __half2*>(&hr);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = col + u;
          const int g = c / p.ch_per_group;
          const float mean = __ldg(p.gn_stats + (t * p.groups + g) * 2);
          const float rstd = __ldg(p.gn_stats + (t * p.groups + g) * 2 + 1);
          const float2 f = __half22float2(hh[u >> 1]);
          const float hv = (u & 1) ? f.y : f.x;
          const float xn = (hv - mean) * rstd * __ldg(p.gn_weight + c) + __ldg(p.gn_bias + c);
          float val = xn * (1.f + o[u]) + bt[u];
          if (has_res) val += p.beta * rs[u];
          o[u] = val;
        }
      } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float val = p.alpha * o[u];
          if (has_res) val += p.beta * rs[u];
          o[u] = val;
        }
      }
      if (p.out_f32) {
        float* op = reinterpret_cast<float*>(p.out) + m * p.ldout + p.out_col0 + col;
        *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
      } else {
        uint4 pk;
        pk.x = pack_h2(o[0], o[1]); pk.y = pack_h2(o[2], o[3]);
        pk.z = pack_h2(o[4], o[5]); pk.w = pack_h2(o[6], o[7]);
        __half* op = reinterpret_cast<__half*>(p.out) + m * p.ldout + p.out_col0 + col;
        *reinterpret_cast<uint4*>(op) = pk;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

static int pick_block_n(int N) {
  if (N % 256 == 0) return 256;
  if (N % 160 == 0) return 160;
  if (N % 128 == 0) return 128;
  if (N % 192 == 0) return 192;
  if (N % 96 == 0) return 96;
  if (N % 64 == 0) return 64;
  if (N % 32 == 0) return 32;
  if (N % 16 == 0) return 16;
  return N <= 128 ? ((N + 15) / 16) * 16 : 128;
}

// choose the pixel box (BW,BH,BT), BW*BH*BT = 128, minimising padded work
static void pick_box(int T, int H, int W, int* BW, int* BH, int* BT) {
  long long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bt = 128 / (bw * bh);
      const long long cost = 1LL * ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(T, bt);
      // prefer wide boxes on ties (longer contiguous TMA rows)
      if (best < 0 || cost < best) { best = cost; *BW = bw; *BH = bh; *BT = bt; }
    }
  }
}

static bool g_attr_set = false;

}  // namespace mgld

using namespace mgld;

extern "C" int mgld_conv_gemm(const mgld_conv_gemm_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!initialised()) { set_error("mgld_init() has not been called"); return MGLD_ERR_NOT_INIT; }
  MGLD_CHECK_ARG(d && d->a && d->w && d->out, "conv_gemm: null pointer");
  MGLD_CHECK_ARG(d->T > 0 && d->H > 0 && d->W > 0, "conv_gemm: bad T/H/W %d/%d/%d", d->T, d->H, d->W);
  // K tail: with a single source and one tap the last 64-chunk may be partial (TMA zero-fills A and W alike)
  MGLD_CHECK_ARG(d->C1 > 0 && (d->C1 % 64 == 0 || (d->C1 % 8 == 0 && d->taps == 1 && d->C2 == 0)),
                 "conv_gemm: C1=%d must be a multiple of 64 (or of 8 for a single-source 1-tap GEMM)", d->C1);
  MGLD_CHECK_ARG(d->C2 >= 0 && d->C2 % 64 == 0 && ((d->C2 > 0) == (d->a2 != nullptr)),
                 "conv_gemm: C2=%d must be a multiple of 64 and match a2", d->C2);
  MGLD_CHECK_ARG(d->taps == 1 || d->taps == 3 || d->taps == 9, "conv_gemm: taps=%d", d->taps);
  MGLD_CHECK_ARG(d->N > 0, "conv_gemm: N=%d", d->N);
  const bool pair = d->epilogue == MGLD_EPI_GEGLU || d->epilogue == MGLD_EPI_SPADE;
  MGLD_CHECK_ARG(d->epilogue >= 0 && d->epilogue <= 2, "conv_gemm: epilogue=%d", d->epilogue);
  if (pair) MGLD_CHECK_ARG(d->N % 128 == 0, "conv_gemm: pair epilogue needs N %% 128 == 0 (N=%d)", d->N);
  if (d->epilogue == MGLD_EPI_SPADE)
    MGLD_CHECK_ARG(d->h && d->gn_stats && d->gn_weight && d->gn_bias && d->groups > 0 &&
                       (d->N / 2) % d->groups == 0,
                   "conv_gemm: SPADE epilogue operands missing");

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.T = d->T; p.H = d->H; p.W = d->W;
  pick_box(d->T, d->H, d->W, &p.BW, &p.BH, &p.BT);
  p.tiles_w = ceil_div(d->W, p.BW);
  p.tiles_h = ceil_div(d->H, p.BH);
  p.tiles_t = ceil_div(d->T, p.BT);
  p.kchunks1 = ceil_div(d->C1, 64);
  p.kchunks = ceil_div(d->C1 + d->C2, 64);
  p.taps = d->taps;
  p.N = d->N;
  p.block_n = pair ? 128 : (d->block_n > 0 ? d->block_n : pick_block_n(d->N));
  MGLD_CHECK_ARG(p.block_n % 16 == 0 && p.block_n >= 16 && p.block_n <= 256, "conv_gemm: block_n=%d", p.block_n);
  p.n_stage_cols = (d->epilogue == MGLD_EPI_GEGLU) ? 64 : p.block_n;
  p.n_out_tile = pair ? 64 : p.block_n;
  p.n_out_total = pair ? d->N / 2 : d->N;
  MGLD_CHECK_ARG(p.n_out_total % 8 == 0, "conv_gemm: output columns (%d) must be a multiple of 8", p.n_out_total);
  p.tmem_cols = 32;
  while (p.tmem_cols < p.block_n) p.tmem_cols <<= 1;
  p.epilogue = d->epilogue; p.act = d->act; p.bias = d->bias;
  p.alpha = d->alpha; p.beta = d->beta;
  p.res = reinterpret_cast<const __half*>(d->res); p.ldres = d->ldres;
  p.h = reinterpret_cast<const __half*>(d->h); p.ldh = d->ldh;
  p.gn_stats = d->gn_stats; p.gn_weight = d->gn_weight; p.gn_bias = d->gn_bias;
  p.groups = d->groups; p.ch_per_group = d->groups > 0 ? p.n_out_total / d->groups : 1;
  p.out = d->out; p.ldout = d->ldout; p.out_col0 = d->out_col0; p.out_f32 = d->out_f32;
  MGLD_CHECK_ARG(d->ldout % 8 == 0 && d->out_col0 % 8 == 0 && (!d->res || d->ldres % 8 == 0) &&
                     (!d->h || d->ldh % 8 == 0),
                 "conv_gemm: leading dimensions / column offset must be multiples of 8");

  const int stage_bytes = kABytes + p.block_n * 128;
  int stages = (109 * 1024) / stage_bytes;                      // aim for two CTAs per SM
  if (stages < 3) stages = (200 * 1024) / stage_bytes;          // otherwise one CTA with a deep pipeline
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const int staging_bytes = kBlockM * (p.n_stage_cols + 4) * 4;
  int smem = stages * stage_bytes;
  if (smem < staging_bytes) smem = staging_bytes;
  smem += 1024;

  // tensor maps
  const int lda = d->lda > 0 ? d->lda : d->C1;
  const int lda2 = d->lda2 > 0 ? d->lda2 : d->C2;
  CUtensorMap tmA, tmA2, tmB;
  {
    uint64_t dims[4] = {(uint64_t)d->C1, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->T};
    uint64_t str[3] = {(uint64_t)lda * 2, (uint64_t)lda * 2 * d->W, (uint64_t)lda * 2 * d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BT};
    int rc = make_tmap_f16(&tmA, d->a, 4, dims, str, box);
    if (rc) return rc;
    if (d->a2) {
      uint64_t dims2[4] = {(uint64_t)d->C2, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->T};
      uint64_t str2[3] = {(uint64_t)lda2 * 2, (uint64_t)lda2 * 2 * d->W, (uint64_t)lda2 * 2 * d->W * d->H};
      rc = make_tmap_f16(&tmA2, d->a2, 4, dims2, str2, box);
      if (rc) return rc;
    } else {
      tmA2 = tmA;
    }
    const uint64_t K = (uint64_t)d->taps * (d->C1 + d->C2);
    uint64_t dimsB[2] = {K, (uint64_t)d->N};
    uint64_t strB[1] = {K * 2};
    uint32_t boxB[2] = {64, (uint32_t)p.block_n};
    rc = make_tmap_f16(&tmB, d->w, 2, dimsB, strB, boxB);
    if (rc) return rc;
  }

  if (!g_attr_set) {
    MGLD_CUDA(cudaFuncSetAttribute(conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    g_attr_set = true;
  }
  dim3 grid(p.tiles_w * p.tiles_h * p.tiles_t, ceil_div(d->N, p.block_n), 1);
  conv_gemm_kernel<<<grid, kThreads, smem, stream>>>(tmA, tmA2, tmB, p);
  MGLD_LAUNCH_CHECK("conv_gemm_kernel");
  return MGLD_OK;
}
