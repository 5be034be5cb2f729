// Fused softmax attention on tcgen05 (sm_100a):  O = softmax(scale * Q K^T) V, fp16 in / fp32 accumulate / fp16 out.
//
// One CTA = 128 query rows of one (batch, head).  Per 128-key block:
//   warp 1 (one thread)  S = Q K_j^T        tcgen05.mma, SMEM x SMEM -> TMEM cols [0,128)
//   warps 2-5 (128 thr)  one query row per thread (TMEM lane == row, so row max/sum need no shuffles):
//                        online softmax, rescale of the O accumulator in TMEM, P_j (fp16) -> swizzled SMEM
//   warp 1               O += P_j V_j        V consumed straight from its row-major [key, d] layout as an
//                                            MN-major UMMA operand (no transpose anywhere)
//   warp 0 (one thread)  TMA producer for Q once and a 2-stage K/V ring.
// Two CTAs are resident per SM for head-dim 64 (112 KB smem, 256 TMEM columns each), so one CTA's softmax overlaps the
// other's MMAs.
//
// Replaces xformers.ops.memory_efficient_attention at ldm/modules/attention.py:298 (self), :371 (cross, K/V
// batch-broadcast as :336-337) and QKVAttentionLegacy at ldm/modules/diffusionmodules/openaimodel.py:554-590.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/mgld.h"
#include "common.h"
#include "ptx.cuh"

namespace mgld {

constexpr int kQTile = 128;
constexpr int kKVTile = 128;
constexpr int kKVStages = 2;
constexpr int kAttThreads = 192;

struct AttnParams {
  int nq, nkv, heads, batch;
  int q_col0, k_col0, v_col0;        // column of head 0 in each matrix
  int q_hstride, k_hstride, v_hstride;  // column stride between heads
  int kv_batched;                    // 0: K/V shared by all batches (cross-attention context)
  float scale_log2e;                 // scale * log2(e)
  __half* out;
  int ldo;                           // out row pitch (elements); out[(b*nq + i)*ldo + h*DH + d]
  int seq;                           // v3: the two softmax warpgroups take turns on the exponential section
  int stagger;                       // v3: cycles by which query tile 1 starts behind tile 0 (0 = none)
  int nomax;                         // v3: skip the row max after the first key block (overflow-checked, see the kernel)
  long long* dbg;                    // optional per-CTA cycle counters [CTAs][16] (mgld_attention_set_debug_counters); null in production
};

template <int DH>
__global__ void __launch_bounds__(kAttThreads, DH == 64 ? 2 : 1)
attention_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int kHalves = DH / 64;                 // 64-column (128-byte) slabs per operand row
  constexpr int kQBytes = kQTile * DH * 2;
  constexpr int kKBytes = kKVTile * DH * 2;
  constexpr int kSlab = kKVTile * 128;              // bytes of one [128 rows x 64 cols] swizzled slab
  constexpr int kPBytes = kQTile * kKVTile * 2;
  constexpr int kTmemCols = 256;
  constexpr uint32_t kOCol = 128;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, kv_full[kKVStages], kv_empty[kKVStages], s_full, p_full, o_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sQ = smem_base;
  const uint32_t sKV = sQ + kQBytes;                       // stage s: K at sKV + s*2*kKBytes, V right after K
  const uint32_t sP = sKV + kKVStages * 2 * kKBytes;
  uint8_t* sP_gen = smem_gen + (sP - smem_base);

  const int q0 = blockIdx.x * kQTile;
  const int head = blockIdx.y;
  const int b = blockIdx.z;
  const int nblk = (p.nkv + kKVTile - 1) / kKVTile;
  const int kvb = p.kv_batched ? b : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < kKVStages; ++s) { mbar_init(smem_u32(&kv_full[s]), 1); mbar_init(smem_u32(&kv_empty[s]), 1); }
    mbar_init(smem_u32(&s_full), 1);
    mbar_init(smem_u32(&p_full), 128);
    mbar_init(smem_u32(&o_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      mbar_expect_tx(smem_u32(&q_full), kQBytes);
#pragma unroll
      for (int hf = 0; hf < kHalves; ++hf)
        tma_load_3d(sQ + hf * kSlab, &tmQ, smem_u32(&q_full), p.q_col0 + head * p.q_hstride + hf * 64, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kKVStages;
        const uint32_t ph = (j / kKVStages) & 1;
        mbar_wait(smem_u32(&kv_empty[s]), ph ^ 1);
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * kKBytes);
        const uint32_t sk = sKV + s * 2 * kKBytes, sv = sk + kKBytes;
#pragma unroll
        for (int hf = 0; hf < kHalves; ++hf) {
          tma_load_3d(sk + hf * kSlab, &tmK, fb, p.k_col0 + head * p.k_hstride + hf * 64, j * kKVTile, kvb);
          tma_load_3d(sv + hf * kSlab, &tmV, fb, p.v_col0 + head * p.v_hstride + hf * 64, j * kKVTile, kvb);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc_s = umma_idesc_f16(kQTile, kKVTile, 0, 0);  // S[128 x 128] = Q (K-major) * K^T (K-major)
      const uint32_t idesc_o = umma_idesc_f16(kQTile, DH, 0, 1);       // O[128 x DH] += P (K-major) * V (MN-major)
      auto issue_s = [&](int j) {
        const int s = j % kKVStages;
        mbar_wait(smem_u32(&kv_full[s]), (j / kKVStages) & 1);
        tc_fence_after();
        const uint32_t sk = sKV + s * 2 * kKBytes;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k) {
          const uint32_t off = (k >> 2) * kSlab + (k & 3) * 32;
          umma_ss(tmem_base, umma_smem_desc(sQ + off, 0, 1024, kSwz128), umma_smem_desc(sk + off, 0, 1024, kSwz128),
                  idesc_s, k != 0);
        }
        umma_commit(smem_u32(&s_full));
      };
      mbar_wait(smem_u32(&q_full), 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kKVStages;
        mbar_wait(smem_u32(&p_full), j & 1);
        tc_fence_after();
        const uint32_t sv = sKV + s * 2 * kKBytes + kKBytes;
#pragma unroll
        for (int k = 0; k < kKVTile / 16; ++k) {
          // A: P rows are 2 slabs of 64 keys; B: V rows (keys) advance 16 * 128 B per step, LBO = next 64-col slab
          const uint64_t adesc = umma_smem_desc(sP + (k >> 2) * kSlab + (k & 3) * 32, 0, 1024, kSwz128);
          const uint64_t bdesc = umma_smem_desc(sv + k * 2048, kSlab, 1024, kSwz128);
          umma_ss(tmem_base + kOCol, adesc, bdesc, idesc_o, (j | k) != 0);
        }
        umma_commit(smem_u32(&kv_empty[s]));
        if (j + 1 < nblk) issue_s(j + 1);
      }
      umma_commit(smem_u32(&o_full));
    }
  } else {
    // ---------------- softmax / correction / output (warps 2..5; thread == query row) ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(smem_u32(&s_full), j & 1);
      tc_fence_after();
      const int kv_left = p.nkv - j * kKVTile;  // keys valid in this block (>= 1)
      // pass 1: row maximum of the raw scores
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < kKVTile; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(trow + c0, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float v = (c0 + i < kv_left) ? __uint_as_float(r[i]) : -INFINITY;
          mx = fmaxf(mx, v);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * p.scale_log2e);  // 0 on the first block (m_run = -inf)
      const float mk = m_new * p.scale_log2e;
      // pass 2: P = exp2(s*k - m*k) -> fp16 -> swizzled smem; row sum in fp32
      float rs = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < kKVTile; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(trow + c0, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float e0 = (c0 + i < kv_left) ? exp2f(__uint_as_float(r[i]) * p.scale_log2e - mk) : 0.f;
          float e1 = (c0 + i + 1 < kv_left) ? exp2f(__uint_as_float(r[i + 1]) * p.scale_log2e - mk) : 0.f;
          const __half2 h2 = __floats2half2_rn(e0, e1);
          // the row sum uses the fp16-rounded probabilities, i.e. exactly what multiplies V
          const float2 back = __half22float2(h2);
          rs += back.x + back.y;
          pk[i >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        // 32 keys = 4 chunks of 16 bytes; slab = c0/64, chunk index inside the 128-byte row XOR (row & 7)
        uint8_t* slab = sP_gen + (c0 >> 6) * kSlab + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (((c0 & 63) >> 3) + ch) ^ (row & 7);
          *reinterpret_cast<uint4*>(slab + chunk * 16) = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      // rescale O (written by PV_{j-1}, complete because s_full(j) was committed after it)
      if (j > 0) {
        const bool need = __any_sync(0xffffffffu, alpha != 1.f);
        if (need) {
#pragma unroll
          for (int c0 = 0; c0 < DH; c0 += 32) {
            uint32_t r[32];
            tmem_ld_x32(trow + kOCol + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st_x32(trow + kOCol + c0, r);
          }
          tmem_st_wait();
        }
      }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full));
    }
    // epilogue: O / l -> fp16 -> global
    mbar_wait(smem_u32(&o_full), 0);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int qi = q0 + row;
    __half* orow = p.out + (static_cast<long long>(b) * p.nq + qi) * p.ldo + head * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(trow + kOCol + c0, r);
      tmem_ld_wait();
      if (qi < p.nq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 v;
          v.x = pack_h2(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
          v.y = pack_h2(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
          v.z = pack_h2(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
          v.w = pack_h2(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + i) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <int DH>
static int launch_attention(const mgld_attention_desc* d, cudaStream_t stream) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.nq = d->nq; p.nkv = d->nkv; p.heads = d->heads; p.batch = d->batch;
  p.q_col0 = d->q_col0; p.k_col0 = d->k_col0; p.v_col0 = d->v_col0;
  p.q_hstride = d->q_head_stride; p.k_hstride = d->k_head_stride; p.v_hstride = d->v_head_stride;
  p.kv_batched = d->kv_batched;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(d->out); p.ldo = d->ldo;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)d->ldq, (uint64_t)d->nq, (uint64_t)d->batch};
    uint64_t str[2] = {(uint64_t)d->ldq * 2, (uint64_t)d->ldq * 2 * d->nq};
    uint32_t box[3] = {64, kQTile, 1};
    int rc = make_tmap_f16(&tmQ, d->q, 3, dims, str, box);
    if (rc) return rc;
    const int kvb = d->kv_batched ? d->batch : 1;
    uint64_t dimsk[3] = {(uint64_t)d->ldk, (uint64_t)d->nkv, (uint64_t)kvb};
    uint64_t strk[2] = {(uint64_t)d->ldk * 2, (uint64_t)d->ldk * 2 * d->nkv};
    uint32_t boxk[3] = {64, kKVTile, 1};
    rc = make_tmap_f16(&tmK, d->k, 3, dimsk, strk, boxk);
    if (rc) return rc;
    uint64_t dimsv[3] = {(uint64_t)d->ldv, (uint64_t)d->nkv, (uint64_t)kvb};
    uint64_t strv[2] = {(uint64_t)d->ldv * 2, (uint64_t)d->ldv * 2 * d->nkv};
    rc = make_tmap_f16(&tmV, d->v, 3, dimsv, strv, boxk);
    if (rc) return rc;
  }
  const int smem = kQTile * DH * 2 + kKVStages * 2 * kKVTile * DH * 2 + kQTile * kKVTile * 2 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MGLD_CUDA(cudaFuncSetAttribute(attention_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(d->nq, kQTile), d->heads, d->batch);
  attention_kernel<DH><<<grid, kAttThreads, smem, stream>>>(tmQ, tmK, tmV, p);
  MGLD_LAUNCH_CHECK("attention_kernel");
  return MGLD_OK;
}


// =====================================================================================================================
// v2 (head_dim 64): two 128-row query tiles per CTA in ping-pong.  While softmax warpgroup i works on S_i(j), the tensor
// core runs PV_{1-i}(j) and S_{1-i}(j+1).  P is written back to TMEM over the S columns (fp16, two per 32-bit column) and
// consumed as the A operand of the PV MMA straight from TMEM (tcgen05.mma A-from-TMEM form) — no shared-memory round trip,
// no proxy fence.  The O rescale is lazy: the running max is only advanced when it grew by more than 2^8 (in the exp2
// domain), which keeps fp16 P <= 256 * row-max-probability and makes the rescale rare; the final O / l is exact for any
// reference max.  One CTA per SM: 320 threads = TMA warp, MMA warp, 2 x 4 softmax warps; TMEM: S0 S1 (128 cols each),
// O0 O1 (64 each) of the 512 columns; smem: Q 32 KB + 3 x (K,V) 32 KB.
// =====================================================================================================================
constexpr int kV2Threads = 320;
constexpr int kV2Stages = 3;

__global__ void __launch_bounds__(kV2Threads, 1)
attention_v2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int DH = 64;
  constexpr int kTileBytes = 128 * DH * 2;  // 16 KB: one [128 x 64] fp16 operand tile
  constexpr uint32_t kSCol0 = 0, kSCol1 = 128, kOCol0 = 256, kOCol1 = 320;

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, kv_full[kV2Stages], kv_empty[kV2Stages], s_full[2], p_full[2], o_full;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                       // two tiles
  const uint32_t sKV = sQ + 2 * kTileBytes;            // stage s: K at +s*2*kTileBytes, V after K

  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.nkv + kKVTile - 1) / kKVTile;
  const int kvb = p.kv_batched ? b : 0;
  const bool tile1_live = q0 + 128 < p.nq;             // the second tile may be entirely out of range

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < kV2Stages; ++s) { mbar_init(smem_u32(&kv_full[s]), 1); mbar_init(smem_u32(&kv_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&p_full[i]), 128); }
    mbar_init(smem_u32(&o_full), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&q_full), 2 * kTileBytes);
      tma_load_3d(sQ, &tmQ, smem_u32(&q_full), p.q_col0 + head * p.q_hstride, q0, b);
      tma_load_3d(sQ + kTileBytes, &tmQ, smem_u32(&q_full), p.q_col0 + head * p.q_hstride, q0 + 128, b);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kV2Stages;
        mbar_wait(smem_u32(&kv_empty[s]), ((j / kV2Stages) & 1) ^ 1);
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * kTileBytes);
        tma_load_3d(sKV + s * 2 * kTileBytes, &tmK, fb, p.k_col0 + head * p.k_hstride, j * kKVTile, kvb);
        tma_load_3d(sKV + s * 2 * kTileBytes + kTileBytes, &tmV, fb, p.v_col0 + head * p.v_hstride, j * kKVTile, kvb);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc_s = umma_idesc_f16(128, kKVTile, 0, 0);
      const uint32_t idesc_o = umma_idesc_f16(128, DH, 0, 1);  // B = V, MN-major
      auto issue_s = [&](int i, int j) {
        const uint32_t sk = sKV + (j % kV2Stages) * 2 * kTileBytes;
        const uint32_t sq = sQ + i * kTileBytes;
#pragma unroll
        for (int k = 0; k < DH / 16; ++k)
          umma_ss(tmem_base + (i ? kSCol1 : kSCol0), umma_smem_desc(sq + k * 32, 0, 1024, kSwz128),
                  umma_smem_desc(sk + k * 32, 0, 1024, kSwz128), idesc_s, k != 0);
        umma_commit(smem_u32(&s_full[i]));
      };
      mbar_wait(smem_u32(&q_full), 0);
      mbar_wait(smem_u32(&kv_full[0]), 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kV2Stages;
        const uint32_t sv = sKV + s * 2 * kTileBytes + kTileBytes;
        if (j + 1 < nblk) {  // K_{j+1} must have landed before S_i(j+1) is issued below
          mbar_wait(smem_u32(&kv_full[(j + 1) % kV2Stages]), ((j + 1) / kV2Stages) & 1);
        }
        for (int i = 0; i < 2; ++i) {
          mbar_wait(smem_u32(&p_full[i]), j & 1);
          tc_fence_after();
          const uint32_t pcol = tmem_base + (i ? kSCol1 : kSCol0);
          const uint32_t ocol = tmem_base + (i ? kOCol1 : kOCol0);
#pragma unroll
          for (int k = 0; k < kKVTile / 16; ++k)  // A: 16 keys = 8 packed columns per step; B: 16 key rows = 2048 B
            umma_ts(ocol, pcol + k * 8, umma_smem_desc(sv + k * 2048, 0, 1024, kSwz128), idesc_o, (j | k) != 0);
          if (i == 1) umma_commit(smem_u32(&kv_empty[s]));
          if (j + 1 < nblk) issue_s(i, j + 1);
        }
      }
      umma_commit(smem_u32(&o_full));
    }
  } else {
    // softmax warpgroup i: warps 2-5 -> tile 0, warps 6-9 -> tile 1; thread == query row
    const int i = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t srow = tmem_base + lane_off + (i ? kSCol1 : kSCol0);
    const uint32_t orow = tmem_base + lane_off + (i ? kOCol1 : kOCol0);
    const float k2 = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      mbar_wait(smem_u32(&s_full[i]), j & 1);
      tc_fence_after();
      const int kv_left = p.nkv - j * kKVTile;
      float mx = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < kKVTile; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(srow + c0, r);
        tmem_ld_wait();
        if (kv_left >= kKVTile) {
#pragma unroll
          for (int u = 0; u < 32; ++u) mx = fmaxf(mx, __uint_as_float(r[u]));
        } else {
#pragma unroll
          for (int u = 0; u < 32; ++u) mx = fmaxf(mx, (c0 + u < kv_left) ? __uint_as_float(r[u]) : -INFINITY);
        }
      }
      // lazy max update: only move the reference max when it grew by more than 8 in the exp2 domain
      float alpha = 1.f;
      const bool grow = (mx - m_run) * k2 > 8.f;   // also true on the first block (m_run = -inf)
      if (grow) {
        alpha = exp2f((m_run - mx) * k2);           // 0 on the first block
        m_run = mx;
      }
      const float mk = m_run * k2;
      float rs = 0.f;
#pragma unroll
      for (int c0 = 0; c0 < kKVTile; c0 += 32) {
        uint32_t r[32];
        tmem_ld_x32(srow + c0, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int u = 0; u < 32; u += 2) {
          float e0 = exp2f(fmaf(__uint_as_float(r[u]), k2, -mk));
          float e1 = exp2f(fmaf(__uint_as_float(r[u + 1]), k2, -mk));
          if (kv_left < kKVTile) {
            if (c0 + u >= kv_left) e0 = 0.f;
            if (c0 + u + 1 >= kv_left) e1 = 0.f;
          }
          const __half2 h2 = __floats2half2_rn(e0, e1);
          const float2 back = __half22float2(h2);
          rs += back.x + back.y;
          pk[u >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        tmem_st_x16(srow + (c0 >> 1), pk);  // P(fp16 pairs) over the already-consumed S columns
      }
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // PV_i(j-1) is complete (s_full(j) was committed after it): rescale this row of O
#pragma unroll
        for (int c0 = 0; c0 < DH; c0 += 32) {
          uint32_t r[32];
          tmem_ld_x32(orow + c0, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) r[u] = __float_as_uint(__uint_as_float(r[u]) * alpha);
          tmem_st_x32(orow + c0, r);
        }
      }
      tmem_st_wait();
      l_run = l_run * alpha + rs;
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full[i]));
    }
    mbar_wait(smem_u32(&o_full), 0);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int qi = q0 + i * 128 + row;
    __half* optr = p.out + (static_cast<long long>(b) * p.nq + qi) * p.ldo + head * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(orow + c0, r);
      tmem_ld_wait();
      if (qi < p.nq) {
#pragma unroll
        for (int u = 0; u < 32; u += 8) {
          uint4 v;
          v.x = pack_h2(__uint_as_float(r[u]) * inv, __uint_as_float(r[u + 1]) * inv);
          v.y = pack_h2(__uint_as_float(r[u + 2]) * inv, __uint_as_float(r[u + 3]) * inv);
          v.z = pack_h2(__uint_as_float(r[u + 4]) * inv, __uint_as_float(r[u + 5]) * inv);
          v.w = pack_h2(__uint_as_float(r[u + 6]) * inv, __uint_as_float(r[u + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c0 + u) = v;
        }
      }
    }
  }
  (void)tile1_live;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

static int launch_attention_v2(const mgld_attention_desc* d, cudaStream_t stream) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.nq = d->nq; p.nkv = d->nkv; p.heads = d->heads; p.batch = d->batch;
  p.q_col0 = d->q_col0; p.k_col0 = d->k_col0; p.v_col0 = d->v_col0;
  p.q_hstride = d->q_head_stride; p.k_hstride = d->k_head_stride; p.v_hstride = d->v_head_stride;
  p.kv_batched = d->kv_batched;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(d->out); p.ldo = d->ldo;
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)d->ldq, (uint64_t)d->nq, (uint64_t)d->batch};
    uint64_t str[2] = {(uint64_t)d->ldq * 2, (uint64_t)d->ldq * 2 * d->nq};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_f16(&tmQ, d->q, 3, dims, str, box);
    if (rc) return rc;
    const int kvb = d->kv_batched ? d->batch : 1;
    uint64_t dimsk[3] = {(uint64_t)d->ldk, (uint64_t)d->nkv, (uint64_t)kvb};
    uint64_t strk[2] = {(uint64_t)d->ldk * 2, (uint64_t)d->ldk * 2 * d->nkv};
    rc = make_tmap_f16(&tmK, d->k, 3, dimsk, strk, box);
    if (rc) return rc;
    uint64_t dimsv[3] = {(uint64_t)d->ldv, (uint64_t)d->nkv, (uint64_t)kvb};
    uint64_t strv[2] = {(uint64_t)d->ldv * 2, (uint64_t)d->ldv * 2 * d->nkv};
    rc = make_tmap_f16(&tmV, d->v, 3, dimsv, strv, box);
    if (rc) return rc;
  }
  const int smem = 2 * 16384 + kV2Stages * 2 * 16384 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    MGLD_CUDA(cudaFuncSetAttribute(attention_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(d->nq, 256), d->heads, d->batch);
  attention_v2_kernel<<<grid, kV2Threads, smem, stream>>>(tmQ, tmK, tmV, p);
  MGLD_LAUNCH_CHECK("attention_v2_kernel");
  return MGLD_OK;
}


// =====================================================================================================================
// v3 (head_dim 64): v2's two ping-pong query tiles, restructured around what actually bounds d=64 attention on sm_100a.
// Per 128x128 score block the tensor core needs 512 cycles (QK^T + PV) but the SM's 16 MUFU lanes need 1024 cycles for
// the 16384 exponentials, so the softmax warps — not the MMA — set the pace, and every instruction they issue counts:
//   * a thread loads its WHOLE score row (128 fp32) into registers with four back-to-back tcgen05.ld and one wait, and
//     releases the S buffer at once (s_free): S_i(j+1) is computed while softmax_i(j) is still exponentiating.  P gets
//     its own TMEM columns (fp16 pairs) instead of overwriting S, so nothing in the softmax loop waits for an MMA
//     except the (long finished) PV of the previous block before P is overwritten;
//   * scale-subtract and the row sum use packed fp32x2 FMA/ADD, the exponential is a bare ex2.approx, max is 3-input;
//   * kEmu of every 8 exponentials are evaluated on the FMA pipe (Cody-Waite split + degree-3 minimax polynomial,
//     7.5e-5 relative error, below fp16 rounding of P) to take load off the MUFU unit;
//   * the row sum is taken over the unrounded fp32 probabilities (what PyTorch's softmax does).
// TMEM (512 columns): S0 S1 [0,256) | P0 P1 [256,384) | O0 O1 [384,512).  One CTA per SM, 384 threads; the softmax
// warpgroups take the registers the TMA / MMA warpgroup gives up (setmaxnreg) so a whole score row stays in registers.
// =====================================================================================================================
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for x in [-125, 125] on the FMA / ALU pipes: n = round(x) via the 1.5 * 2^23 trick, 2^(x - n) by a degree-3 minimax
// polynomial on [-0.5, 0.5], exponent patched in with an integer add.  Two lanes at a time.
__device__ __forceinline__ void exp2_poly2(float& y0, float& y1, float x0, float x1) {
  constexpr float kMagic = 12582912.f;  // 1.5 * 2^23
  constexpr float c0 = 0.9999280571937561f, c1 = 0.6932610273361206f, c2 = 0.2426111102104187f, c3 = 0.05517156794667244f;
  x0 = fmaxf(x0, -125.f);
  x1 = fmaxf(x1, -125.f);
  float t0, t1, n0, n1, r0, r1, p0, p1;
  fadd2(t0, t1, x0, x1, kMagic, kMagic);
  fadd2(n0, n1, t0, t1, -kMagic, -kMagic);
  ffma2(r0, r1, n0, n1, -1.f, -1.f, x0, x1);
  ffma2(p0, p1, r0, r1, c3, c3, c2, c2);
  ffma2(p0, p1, p0, p1, r0, r1, c1, c1);
  ffma2(p0, p1, p0, p1, r0, r1, c0, c0);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

constexpr int kV3Threads = 384;   // warpgroup 0: TMA warp, MMA warp, two idle warps; warpgroups 1, 2: softmax of tile 0, 1
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

template <int kEmu, bool kDbg>
__global__ void __launch_bounds__(kV3Threads, 1)
attention_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  constexpr int DH = 64;
  constexpr int kTileBytes = 128 * DH * 2;  // 16 KB: one [128 x 64] fp16 operand tile
  constexpr uint32_t kSCol = 0, kPCol = 256, kOCol = 384;   // + i * 128 / 64 / 64

  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, kv_full[kV2Stages], kv_empty[kV2Stages], s_full[2], s_free[2], p_full[2],
      pv_done[2], exp_turn[2];
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;                       // two tiles
  const uint32_t sKV = sQ + 2 * kTileBytes;            // stage s: K at +s*2*kTileBytes, V after K

  const int q0 = blockIdx.x * 256;
  const int head = blockIdx.y, b = blockIdx.z;
  const int nblk = (p.nkv + kKVTile - 1) / kKVTile;
  const int kvb = p.kv_batched ? b : 0;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < kV2Stages; ++s) { mbar_init(smem_u32(&kv_full[s]), 1); mbar_init(smem_u32(&kv_empty[s]), 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_free[i]), 4);     // one arrival per softmax warp (lane 0, after __syncwarp): 128 lanes
      mbar_init(smem_u32(&p_full[i]), 4);     // arriving on one barrier word serialise in the shared-memory atomic unit
      mbar_init(smem_u32(&pv_done[i]), 1);
      mbar_init(smem_u32(&exp_turn[i]), 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();   // after the TMEM allocation (see conv_gemm.cu)
  pdl_wait();                // q / k / v come from the preceding GEMM (see common.h)
  // kDbg = false folds every `if (dbg)` below away (the production instantiations)
  long long* const dbg = (kDbg && p.dbg) ? p.dbg + 16 * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
  const long long dbg_t0 = dbg ? clock64() : 0;
  // wait on a barrier, adding the cycles spent to `acc` when the development counters are on
  auto wait_t = [&](uint32_t bar, uint32_t parity, long long& acc) {
    if (dbg) {
      const long long t = clock64();
      mbar_wait_poll(bar, parity);
      acc += clock64() - t;
    } else {
      mbar_wait_poll(bar, parity);
    }
  };

  if (warp < 4) {
  setmaxnreg_dec<88>();
  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(smem_u32(&q_full), 2 * kTileBytes);
      tma_load_3d(sQ, &tmQ, smem_u32(&q_full), p.q_col0 + head * p.q_hstride, q0, b);
      tma_load_3d(sQ + kTileBytes, &tmQ, smem_u32(&q_full), p.q_col0 + head * p.q_hstride, q0 + 128, b);
      long long w_empty = 0;
      for (int j = 0; j < nblk; ++j) {
        const int s = j % kV2Stages;
        mbar_wait(smem_u32(&kv_empty[s]), ((j / kV2Stages) & 1) ^ 1);   // parked by the hardware, not polling
        (void)w_empty;
        const uint32_t fb = smem_u32(&kv_full[s]);
        mbar_expect_tx(fb, 2 * kTileBytes);
        tma_load_3d(sKV + s * 2 * kTileBytes, &tmK, fb, p.k_col0 + head * p.k_hstride, j * kKVTile, kvb);
        tma_load_3d(sKV + s * 2 * kTileBytes + kTileBytes, &tmV, fb, p.v_col0 + head * p.v_hstride, j * kKVTile, kvb);
      }
      if (dbg) dbg[14] = w_empty;
    }
  } else if (warp == 1) {
    {
      // Issue loop.  All 32 lanes run the loop and the waits; only the tcgen05 instructions are predicated on the elected
      // lane.  That keeps the loop state (descriptors, stage, barrier phases) warp-uniform, so it lives in uniform
      // registers and feeds UTCHMMA directly instead of going through a scalar chain of 64-bit adds + R2UR per operand
      // (this chain competes with two softmax warps for the scheduler; it set the latency from "P is ready" to "PV is
      // issued").  All state is carried incrementally; descriptors are base + compile-time offset (the 14-bit address
      // field never carries into the next field).
      const bool leader = elect_one();
      const uint32_t idesc_s = umma_idesc_f16(128, kKVTile, 0, 0);
      const uint32_t idesc_o = umma_idesc_f16(128, DH, 0, 1);  // B = V, MN-major
      const uint64_t dq = umma_smem_desc(sQ, 0, 1024, kSwz128);
      const uint64_t dkv0 = umma_smem_desc(sKV, 0, 1024, kSwz128);
      constexpr uint32_t kStageStep = (2 * kTileBytes) >> 4;
      const uint32_t bar_sfull = smem_u32(&s_full[0]), bar_sfree = smem_u32(&s_free[0]), bar_pfull = smem_u32(&p_full[0]);
      const uint32_t bar_pvdone = smem_u32(&pv_done[0]), bar_kvfull = smem_u32(&kv_full[0]), bar_kvempty = smem_u32(&kv_empty[0]);
      const uint32_t ts = tmem_base + kSCol, tp = tmem_base + kPCol, to = tmem_base + kOCol;
      auto issue_s = [&](const int i, const uint64_t dk) {
        if (leader) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma_ss(ts + i * 128, dq + ((i * kTileBytes + k * 32) >> 4), dk + ((k * 32) >> 4), idesc_s, k != 0);
          umma_commit(bar_sfull + 8 * i);
        }
      };
      mbar_wait_poll(smem_u32(&q_full), 0);
      mbar_wait_poll(bar_kvfull, 0);
      tc_fence_after();
      issue_s(0, dkv0);
      issue_s(1, dkv0);
      int st = 0, st_next = 1;                   // stage of block j / j + 1
      uint32_t ph_next = 0;                      // kv_full parity of block j + 1
      uint64_t dkv = dkv0;                       // K descriptor of block j (V = + kTileBytes)
      uint64_t dkv_next = dkv0 + kStageStep;
      uint32_t par = 0;                          // j & 1
      long long w_kv = 0, w_sfree = 0, w_pfull = 0;
      const long long t_loop = dbg ? clock64() : 0;
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) {
          wait_t(bar_kvfull + 8 * st_next, ph_next, w_kv);   // K_{j+1} has landed
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            wait_t(bar_sfree + 8 * i, par, w_sfree);      // softmax_i holds S_i(j) in registers: the buffer is free
            tc_fence_after();
            issue_s(i, dkv_next);
          }
        }
        const uint64_t dv = dkv + (kTileBytes >> 4);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          wait_t(bar_pfull + 8 * i, par, w_pfull);
          tc_fence_after();
          if (leader) {
#pragma unroll
            for (int k = 0; k < kKVTile / 16; ++k)  // A: 16 keys = 8 packed columns per step; B: 16 key rows = 2048 B
              umma_ts(to + i * 64, tp + i * 64 + k * 8, dv + ((k * 2048) >> 4), idesc_o, (j | k) != 0);
            umma_commit(bar_pvdone + 8 * i);
            if (i == 1) umma_commit(bar_kvempty + 8 * st);
          }
        }
        par ^= 1;
        st = st_next;
        dkv = dkv_next;
        if (++st_next == kV2Stages) { st_next = 0; ph_next ^= 1; dkv_next = dkv0; } else { dkv_next += kStageStep; }
      }
      if (dbg && leader) { dbg[0] = w_kv; dbg[1] = w_sfree; dbg[2] = w_pfull; dbg[3] = clock64() - t_loop; }
    }
  }
  } else {
    // softmax warpgroup i: warps 4-7 -> tile 0, warps 8-11 -> tile 1; thread == query row
    setmaxnreg_inc<208>();
    const int i = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t srow = tmem_base + lane_off + kSCol + i * 128;
    const uint32_t prow = tmem_base + lane_off + kPCol + i * 64;
    const uint32_t orow = tmem_base + lane_off + kOCol + i * 64;
    const float k2 = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    long long w_sfull = 0, w_pv = 0, w_seq = 0, c_exp = 0;
    const long long t_loop = dbg ? clock64() : 0;
    // The score row of the current block lives in registers.  Its load for block j+1 is issued as soon as the
    // exponentials of block j are done, so the S wait and the TMEM load latency hide behind the wait for the P store.
    float sc[kKVTile];
    auto load_scores = [&](const int j) {
      wait_t(smem_u32(&s_full[i]), j & 1, w_sfull);
      tc_fence_after();
#pragma unroll
      for (int c0 = 0; c0 < kKVTile; c0 += 32) tmem_ld_x32(srow + c0, reinterpret_cast<uint32_t*>(sc) + c0);
    };
    auto scores_loaded = [&]() {   // S_i is in registers: hand the buffer back for S_i(j+1)
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_free[i]));
    };
    // one 128-key block of this thread's query row.  kMasked: the last, partial block (keys beyond nkv are masked);
    // a separate instantiation so that the full blocks carry no per-element select
    auto block = [&](const int j, auto masked_tag, const bool has_next) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      if constexpr (kMasked) {
        const int kv_left = p.nkv - j * kKVTile;
#pragma unroll
        for (int u = 0; u < kKVTile; ++u)
          if (u >= kv_left) sc[u] = -INFINITY;
      }
      auto row_max = [&]() {
        float mx0 = sc[0], mx1 = sc[1], mx2 = sc[2], mx3 = sc[3];
#pragma unroll
        for (int u = 4; u < kKVTile; u += 8) {
          mx0 = fmaxf(mx0, fmaxf(sc[u], sc[u + 1]));
          mx1 = fmaxf(mx1, fmaxf(sc[u + 2], sc[u + 3]));
          if (u + 4 < kKVTile) {
            mx2 = fmaxf(mx2, fmaxf(sc[u + 4], sc[u + 5]));
            mx3 = fmaxf(mx3, fmaxf(sc[u + 6], sc[u + 7]));
          }
        }
        return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      };
      auto rescale_o = [&](const float a) {
#pragma unroll
        for (int o0 = 0; o0 < DH; o0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(orow + o0, r);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 16; ++u) r[u] = __float_as_uint(__uint_as_float(r[u]) * a);
          tmem_st_x16(orow + o0, r);
        }
      };
      // P = exp2(s * k2 - m_run * k2) as fp16 pairs into TMEM, row sum into rs0 + rs1.  `first_pass`: the wait for PV_i(j-1)
      // (P_i may only be overwritten and O_i rescaled once it is complete) sits behind the first chunk of exponentials —
      // by then the PV issued at the end of the previous block has long finished — followed by the pending rescale.
      float rs0 = 0.f, rs1 = 0.f;
      auto exp_store = [&](const float nmk, const bool first_pass, const bool rescale, const float a) {
        rs0 = 0.f; rs1 = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < kKVTile; c0 += 32) {
          uint32_t pk[16];
#pragma unroll
          for (int u = 0; u < 32; u += 2) {
            float x0, x1, e0, e1;
            ffma2(x0, x1, sc[c0 + u], sc[c0 + u + 1], k2, k2, nmk, nmk);
            if (((u >> 1) & 3) < kEmu / 2) {          // kEmu of every 8 exponentials on the FMA pipe
              exp2_poly2(e0, e1, x0, x1);
            } else {
              e0 = ex2_approx(x0);
              e1 = ex2_approx(x1);
            }
            fadd2(rs0, rs1, rs0, rs1, e0, e1);
            pk[u >> 1] = pack_h2(e0, e1);
          }
          if (c0 == 0 && first_pass && j > 0) {
            wait_t(smem_u32(&pv_done[i]), (j - 1) & 1, w_pv);
            tc_fence_after();
            if (__any_sync(0xffffffffu, rescale)) rescale_o(a);
          }
          tmem_st_x16(prow + (c0 >> 1), pk);
        }
      };
      float alpha = 1.f;
      if (p.seq && (i == 1 || j > 0)) wait_t(smem_u32(&exp_turn[i ^ 1]), (i == 1 ? j : j - 1) & 1, w_seq);
      const long long t_exp = dbg ? clock64() : 0;
      if (!p.nomax || j == 0) {
        // lazy max update: only move the reference max when it grew by more than 8 in the exp2 domain
        const float mx = row_max();
        const bool grow = (mx - m_run) * k2 > 8.f;   // also true on the first block (m_run = -inf)
        if (grow) {
          alpha = ex2_approx((m_run - mx) * k2);      // 0 on the first block
          m_run = mx;
        }
        exp_store(-m_run * k2, true, grow, alpha);
      } else {
        // No row max at all after the first block: any reference m_run gives the same softmax as long as no exponential
        // overflows fp16.  The exponentials are non-negative, so a row sum below 2^15 bounds every one of them; only when
        // the scores outgrew the reference by a factor > 2^15 (rare) is the block redone with the true max.
        exp_store(-m_run * k2, true, false, 1.f);
        const bool over = !(rs0 + rs1 < 32768.f);
        if (__any_sync(0xffffffffu, over)) {
          const float mx = row_max();
          if (mx > m_run) {
            alpha = ex2_approx((m_run - mx) * k2);
            m_run = mx;
          }
          rescale_o(alpha);                           // PV_i(j-1) is complete (waited for in the first pass)
          exp_store(-m_run * k2, false, false, 1.f);
        }
      }
      if (p.seq) { __syncwarp(); if (lane == 0) mbar_arrive(smem_u32(&exp_turn[i])); }
      if (dbg) c_exp += clock64() - t_exp;
      if (has_next) load_scores(j + 1);
      tmem_st_wait();
      l_run = l_run * alpha + (rs0 + rs1);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&p_full[i]));
      if (has_next) scores_loaded();
    };
    const int nfull = p.nkv / kKVTile;
    if (i == 1 && p.stagger > 0) {   // de-phase the two tiles: tile 1's block boundaries (PV / S bursts on the tensor pipe,
      const long long t0 = clock64();   // max + barrier phases) then fall into tile 0's exponential phase and vice versa
      while (clock64() - t0 < p.stagger) {}
    }
    load_scores(0);
    scores_loaded();
    for (int j = 0; j < nfull; ++j) block(j, std::false_type{}, j + 1 < nblk);
    if (nfull < nblk) block(nfull, std::true_type{}, false);
    if (dbg && (threadIdx.x & 127) == 0) {
      long long* d = dbg + 4 + 5 * i;
      d[0] = w_sfull; d[1] = w_pv; d[2] = w_seq; d[3] = c_exp; d[4] = clock64() - t_loop;
    }
    mbar_wait_poll(smem_u32(&pv_done[i]), (nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int qi = q0 + i * 128 + row;
    __half* optr = p.out + (static_cast<long long>(b) * p.nq + qi) * p.ldo + head * DH;
#pragma unroll
    for (int c0 = 0; c0 < DH; c0 += 32) {
      uint32_t r[32];
      tmem_ld_x32(orow + c0, r);
      tmem_ld_wait();
      if (qi < p.nq) {
#pragma unroll
        for (int u = 0; u < 32; u += 8) {
          uint4 v;
          v.x = pack_h2(__uint_as_float(r[u]) * inv, __uint_as_float(r[u + 1]) * inv);
          v.y = pack_h2(__uint_as_float(r[u + 2]) * inv, __uint_as_float(r[u + 3]) * inv);
          v.z = pack_h2(__uint_as_float(r[u + 4]) * inv, __uint_as_float(r[u + 5]) * inv);
          v.w = pack_h2(__uint_as_float(r[u + 6]) * inv, __uint_as_float(r[u + 7]) * inv);
          *reinterpret_cast<uint4*>(optr + c0 + u) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[15] = clock64() - dbg_t0;
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}



// attention_kv80.cu: cross-attention against <= 80 shared keys (the 77 text tokens) on warp-level MMAs
bool attention_kv80_supported(const mgld_attention_desc* d);
int launch_attention_kv80(const mgld_attention_desc* d, cudaStream_t stream);
// attention_hd512.cu: head dim 512 (the VAEs' single-head middle attention), split-D flash kernel
int launch_attention_hd512(const mgld_attention_desc* d, cudaStream_t stream);

static long long* g_attn_dbg = nullptr;
static bool nkv_blocks_for_stagger(int nkv) { return nkv > 4 * kKVTile; }   // pointless for a handful of key blocks

static int attn_v3_emu() {   // exponentials per 8 evaluated on the FMA pipe: MGLD_ATTN_EMU = 0, 2 or 4
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MGLD_ATTN_EMU");
    v = e ? atoi(e) : 2;
    if (v != 0 && v != 2 && v != 4) v = 2;
  }
  return v;
}

static int launch_attention_v3(const mgld_attention_desc* d, cudaStream_t stream) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  p.nq = d->nq; p.nkv = d->nkv; p.heads = d->heads; p.batch = d->batch;
  p.q_col0 = d->q_col0; p.k_col0 = d->k_col0; p.v_col0 = d->v_col0;
  p.q_hstride = d->q_head_stride; p.k_hstride = d->k_head_stride; p.v_hstride = d->v_head_stride;
  p.kv_batched = d->kv_batched;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(d->out); p.ldo = d->ldo;
  p.dbg = g_attn_dbg;
  {
    static int seq = -1;
    if (seq < 0) { const char* e = getenv("MGLD_ATTN_SEQ"); seq = e ? (atoi(e) != 0) : 0; }   // off: measured slower (profiles/r01_dev_run21*)
    p.seq = seq;
    static int stagger = -1;
    if (stagger < 0) { const char* e = getenv("MGLD_ATTN_STAGGER"); stagger = e ? atoi(e) : 0; }
    p.stagger = nkv_blocks_for_stagger(d->nkv) ? stagger : 0;
    static int nomax = -1;
    if (nomax < 0) { const char* e = getenv("MGLD_ATTN_NOMAX"); nomax = e ? (atoi(e) != 0) : 1; }
    p.nomax = nomax;
  }
  CUtensorMap tmQ, tmK, tmV;
  {
    uint64_t dims[3] = {(uint64_t)d->ldq, (uint64_t)d->nq, (uint64_t)d->batch};
    uint64_t str[2] = {(uint64_t)d->ldq * 2, (uint64_t)d->ldq * 2 * d->nq};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_f16(&tmQ, d->q, 3, dims, str, box);
    if (rc) return rc;
    const int kvb = d->kv_batched ? d->batch : 1;
    uint64_t dimsk[3] = {(uint64_t)d->ldk, (uint64_t)d->nkv, (uint64_t)kvb};
    uint64_t strk[2] = {(uint64_t)d->ldk * 2, (uint64_t)d->ldk * 2 * d->nkv};
    rc = make_tmap_f16(&tmK, d->k, 3, dimsk, strk, box);
    if (rc) return rc;
    uint64_t dimsv[3] = {(uint64_t)d->ldv, (uint64_t)d->nkv, (uint64_t)kvb};
    uint64_t strv[2] = {(uint64_t)d->ldv * 2, (uint64_t)d->ldv * 2 * d->nkv};
    rc = make_tmap_f16(&tmV, d->v, 3, dimsv, strv, box);
    if (rc) return rc;
  }
  const int smem = 2 * 16384 + kV2Stages * 2 * 16384 + 1024;
  using Fn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, AttnParams);
  static const Fn kFns[4] = {attention_v3_kernel<0, false>, attention_v3_kernel<2, false>, attention_v3_kernel<4, false>,
                             attention_v3_kernel<2, true>};
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 4; ++i) MGLD_CUDA(cudaFuncSetAttribute(kFns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(ceil_div(d->nq, 256), d->heads, d->batch);
  MGLD_CUDA(launch_pdl(kFns[p.dbg ? 3 : attn_v3_emu() / 2], grid, dim3(kV3Threads), smem, stream, tmQ, tmK, tmV, p));
  MGLD_LAUNCH_CHECK("attention_v3_kernel");
  return MGLD_OK;
}

}  // namespace mgld

using namespace mgld;

// Development hook: per-CTA cycle counters of the next attention (v3) launches ([CTAs][16] int64, device memory):
// 0-3 MMA thread (wait kv_full, wait s_free, wait p_full, loop total); 4-8 / 9-13 softmax warpgroup 0 / 1 (wait s_full,
// wait pv_done, wait for the exponential turn, exponentials + P store issue, loop total); 14 TMA wait kv_empty; 15 kernel total.
extern "C" void mgld_attention_set_debug_counters(void* dev_ptr) { g_attn_dbg = reinterpret_cast<long long*>(dev_ptr); }

extern "C" int mgld_attention(const mgld_attention_desc* d, void* stream) {
  if (!initialised()) { set_error("mgld_init() has not been called"); return MGLD_ERR_NOT_INIT; }
  MGLD_CHECK_ARG(d && d->q && d->k && d->v && d->out, "attention: null pointer");
  MGLD_CHECK_ARG(d->nq > 0 && d->nkv > 0 && d->heads > 0 && d->batch > 0, "attention: bad sizes");
  MGLD_CHECK_ARG(d->head_dim == 64 || d->head_dim == 128 || d->head_dim == 512,
                 "attention: head_dim %d not supported (64, 128, 512)", d->head_dim);
  MGLD_CHECK_ARG(d->ldq % 8 == 0 && d->ldk % 8 == 0 && d->ldv % 8 == 0 && d->ldo % 8 == 0,
                 "attention: row pitches must be multiples of 8 elements");
  MGLD_CHECK_ARG(d->q_col0 % 8 == 0 && d->k_col0 % 8 == 0 && d->v_col0 % 8 == 0, "attention: column offsets");
  if (d->head_dim == 512) return launch_attention_hd512(d, (cudaStream_t)stream);
  if (attention_kv80_supported(d)) return launch_attention_kv80(d, (cudaStream_t)stream);   // short shared context
  if (d->head_dim == 64) {
    // v2 (two query tiles per CTA, P in TMEM) pays off once there are >= 256 queries; MGLD_ATTN_V1=1 forces v1
    static const bool force_v1 = getenv("MGLD_ATTN_V1") != nullptr;
    static const bool force_v2 = getenv("MGLD_ATTN_V2") != nullptr;
    if (!force_v1 && !force_v2 && d->nq >= 256) return launch_attention_v3(d, (cudaStream_t)stream);
    if (!force_v1 && d->nq >= 256) return launch_attention_v2(d, (cudaStream_t)stream);
    return launch_attention<64>(d, (cudaStream_t)stream);
  }
  return launch_attention<128>(d, (cudaStream_t)stream);
}
