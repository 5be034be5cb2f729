// Fused softmax attention for head dim 512 (the single-head middle attention of both VAEs) on tcgen05:
//   O = softmax(scale * Q K^T) V,  fp16 in / fp32 accumulate / fp16 out,  no N x N tensor in memory.
//
// Why this is its own kernel.  At d = 512 a 128-row O accumulator alone is 128 x 512 fp32 = all 512 TMEM columns, and a
// 128-row Q tile is 128 KB of shared memory, so neither the score tile nor a K/V double buffer of the d = 64 / 128 kernels
// (attention.cu) fits beside them.  The CTA therefore owns (128 query rows) x (HALF of the output columns):
//   TMEM   S0 S1 [0,128) (two 64-key score blocks, fp32) | P0 P1 [128,192) (fp16 pairs) | O [256,512) (256 columns)
//   smem   Q tile 128 KB (eight [128 x 64] swizzled slabs, loaded once) + a 96 KB ring through which the K slabs
//          ([64 keys x 64 d], 8 KB) and the V slabs of this CTA's column half stream like the operands of a GEMM mainloop:
//          a slab is released the moment the four K=16 MMAs that read it have completed.
// Both CTAs of a query tile compute the full S = Q K^T (1.5x the algorithmic tensor work) - the price of not exchanging
// scores between SMs.  Per 64-key block the CTA moves 64 KB of K + 32 KB of V from L2 for 1536 tensor cycles with about
// 64 KB in flight; measured 2690 cycles per block (655 us per 14400-token frame, 649 TFLOP/s algorithmic; ncu: tensor pipe
// 61 %, L2 throughput 19 %, DRAM 19 MB against 59 MB of Q + K + V + O - the inputs are L2-resident).  It replaces GEMM ->
// fp32 scores -> row softmax -> GEMM: 1.9 - 4.2 GB of DRAM traffic per frame and 2.2x the time of the whole block.
//
// Roles (192 threads): warp 0 = TMA producer (one elected thread), warp 1 = tcgen05.mma issuer, warps 2-5 = softmax,
// thread == query row (TMEM lane).  MMA program order  S(0) S(1) PV(0) S(2) PV(1) ...  so the tensor core computes
// S(j+1) while the softmax warps turn S(j) into P(j); P is the A operand of the PV MMA straight from TMEM.  The running
// max is lazy (advanced only when it grew by 2^8 in the exp2 domain, as in attention.cu), which makes the 256-column
// rescale of O rare; O / l at the end is exact for any reference max.
//
// Replaces xformers.ops.memory_efficient_attention at ldm/modules/diffusionmodules/model.py:294 (MemoryEfficientAttnBlock,
// :247-305; one head of width C = 512, scale C^-0.5), used by Encoder.mid.attn_1 / Decoder.mid.attn_1 (model.py:473-572,
// 926-1056) of AutoencoderKL and VideoAutoencoderKLResi.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/mgld.h"
#include "common.h"
#include "ptx.cuh"

namespace mgld {

constexpr int kHdThreads = 192;
constexpr int kHdDim = 512;
constexpr int kHdChunks = kHdDim / 64;     // 64-column (128-byte) slabs per Q / K row
constexpr int kHdOutCols = 256;            // output columns per CTA
constexpr int kHdOutChunks = kHdOutCols / 64;
constexpr int kHdKeys = 64;                // keys per block
constexpr int kHdQSlab = 128 * 128;        // [128 rows x 64 cols] fp16, 128B-swizzled
constexpr int kHdKSlab = kHdKeys * 128;    // [64 keys x 64 cols]
constexpr int kHdRingSlabs = 12;           // 96 KB
constexpr int kHdSmem = kHdChunks * kHdQSlab + kHdRingSlabs * kHdKSlab + 1024;
constexpr uint32_t kHdSCol = 0, kHdPCol = 128, kHdOCol = 256;

struct Hd512Params {
  int nq, nkv;
  int q_col0, k_col0, v_col0;   // first column of this head in each matrix (multiples of 64)
  int kv_batched;
  float scale_log2e;
  __half* out;
  int ldo, out_col0;
};

template <int G>   // slabs per TMA operation / ring slot (4 or 2): compile time, so that the MMA issuer's operand descriptors are
__global__ void __launch_bounds__(kHdThreads, 1)   // base + constant (the scalar chain of the single issuing thread was the bound)
attention_hd512_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                       const __grid_constant__ CUtensorMap tmV, const Hd512Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, full_bar[kHdRingSlabs], empty_bar[kHdRingSlabs], s_full[2], p_full[2], pv_done;
  __shared__ uint32_t tmem_base_slot;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sQ = smem_base;
  const uint32_t sRing = sQ + kHdChunks * kHdQSlab;

  const int q0 = blockIdx.x * 128;
  const int half = blockIdx.y;               // which 256 output columns
  const int b = blockIdx.z;
  const int kvb = p.kv_batched ? b : 0;
  const int nblk = (p.nkv + kHdKeys - 1) / kHdKeys;
  constexpr int nslots = kHdRingSlabs / G;
  constexpr uint32_t slot_bytes = G * kHdKSlab;
  constexpr int k_units = kHdChunks / G, v_units = kHdOutChunks / G;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < kHdRingSlabs; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&s_full[i]), 1); mbar_init(smem_u32(&p_full[i]), 128); }
    mbar_init(smem_u32(&pv_done), 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer: Q once, then K(0) K(1) V(0) K(2) V(1) ... through the ring ----------------
      mbar_expect_tx(smem_u32(&q_full), kHdChunks * kHdQSlab);
      tma_load_4d(sQ, &tmQ, smem_u32(&q_full), 0, q0, p.q_col0 / 64, b);
      tma_load_4d(sQ + 4 * kHdQSlab, &tmQ, smem_u32(&q_full), 0, q0, p.q_col0 / 64 + 4, b);
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      auto load_slot = [&](const CUtensorMap* tm, int col0, int chunk0, int row0) {
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        mbar_expect_tx(full0 + 8 * s, slot_bytes);
        const uint32_t dst = sRing + s * slot_bytes;
        tma_load_4d(dst, tm, full0 + 8 * s, 0, row0, col0 / 64 + chunk0, kvb);
        if (++s == nslots) { s = 0; ph ^= 1; }
      };
      auto load_k = [&](int j) { for (int u = 0; u < k_units; ++u) load_slot(&tmK, p.k_col0, u * G, j * kHdKeys); };
      auto load_v = [&](int j) {
        for (int u = 0; u < v_units; ++u) load_slot(&tmV, p.v_col0, half * kHdOutChunks + u * G, j * kHdKeys);
      };
      load_k(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc_s = umma_idesc_f16(128, kHdKeys, 0, 0);   // S[128 x 64] = Q (K-major) * K^T (K-major), K = 512
      const uint32_t idesc_o = umma_idesc_f16(128, 64, 0, 1);        // O[128 x 64] += P (TMEM) * V slab (MN-major), K = 64
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      // Descriptors differ only in their 14-bit (address >> 4) field (shared memory < 256 KB: no carry out of it): Q slabs
      // are base + compile-time offset, the ring slot's base is carried incrementally.
      const uint64_t dq0 = umma_smem_desc(sQ, 0, 1024, kSwz128);
      const uint64_t dring0 = umma_smem_desc(sRing, 0, 1024, kSwz128);
      uint64_t dslot = dring0;
      uint32_t bar_full = full0, bar_empty = empty0;
      auto advance = [&]() {
        dslot += slot_bytes >> 4; bar_full += 8; bar_empty += 8;
        if (++s == nslots) { s = 0; ph ^= 1; dslot = dring0; bar_full = full0; bar_empty = empty0; }
      };
      auto issue_s = [&](int j) {
        const uint32_t scol = tmem_base + kHdSCol + (j & 1) * kHdKeys;
#pragma unroll
        for (int u = 0; u < k_units; ++u) {
          mbar_wait(bar_full, ph);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < G; ++g) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_ss(scol, dq0 + (((u * G + g) * kHdQSlab + k * 32) >> 4), dslot + ((g * kHdKSlab + k * 32) >> 4), idesc_s,
                      (u | g | k) != 0);
          }
          umma_commit(bar_empty);
          advance();
        }
        umma_commit(smem_u32(&s_full[j & 1]));
      };
      auto issue_pv = [&](int j) {
        mbar_wait(smem_u32(&p_full[j & 1]), (j >> 1) & 1);
        tc_fence_after();
        const uint32_t pcol = tmem_base + kHdPCol + (j & 1) * (kHdKeys / 2);
        const uint32_t acc = j != 0;
#pragma unroll
        for (int u = 0; u < v_units; ++u) {
          mbar_wait(bar_full, ph);
          tc_fence_after();
#pragma unroll
          for (int g = 0; g < G; ++g) {   // 64-column slab u * G + g of this CTA's output half
#pragma unroll
            for (int k = 0; k < 4; ++k)   // A: 16 keys = 8 packed TMEM columns per step; B: 16 key rows = 2048 B
              umma_ts(tmem_base + kHdOCol + (u * G + g) * 64, pcol + k * 8, dslot + ((g * kHdKSlab + k * 2048) >> 4), idesc_o,
                      k != 0 ? 1u : acc);
          }
          umma_commit(bar_empty);
          advance();
        }
        umma_commit(smem_u32(&pv_done));
      };
      mbar_wait(smem_u32(&q_full), 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        issue_pv(j);
      }
    }
  } else {
    // ---------------- softmax / correction / output (warps 2..5; thread == query row) ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const float k2 = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    for (int j = 0; j < nblk; ++j) {
      const int bf = j & 1;
      mbar_wait(smem_u32(&s_full[bf]), (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[kHdKeys];
      tmem_ld_x32(trow + kHdSCol + bf * kHdKeys, r);
      tmem_ld_x32(trow + kHdSCol + bf * kHdKeys + 32, r + 32);
      tmem_ld_wait();
      const int kv_left = p.nkv - j * kHdKeys;   // keys valid in this block (>= 1)
      float mx = -INFINITY;
      if (kv_left >= kHdKeys) {
#pragma unroll
        for (int u = 0; u < kHdKeys; ++u) mx = fmaxf(mx, __uint_as_float(r[u]));
      } else {
#pragma unroll
        for (int u = 0; u < kHdKeys; ++u) mx = fmaxf(mx, u < kv_left ? __uint_as_float(r[u]) : -INFINITY);
      }
      // lazy max update: the reference max only moves when it grew by more than 8 in the exp2 domain (P <= 2^8)
      float alpha = 1.f;
      const bool grow = (mx - m_run) * k2 > 8.f;   // also true on the first block (m_run = -inf)
      if (grow) {
        alpha = exp2f((m_run - mx) * k2);           // 0 on the first block
        m_run = mx;
      }
      const float mk = m_run * k2;
      float rs = 0.f;
      uint32_t pk[kHdKeys / 2];
#pragma unroll
      for (int u = 0; u < kHdKeys; u += 2) {
        float e0 = exp2f(fmaf(__uint_as_float(r[u]), k2, -mk));
        float e1 = exp2f(fmaf(__uint_as_float(r[u + 1]), k2, -mk));
        if (kv_left < kHdKeys) {
          if (u >= kv_left) e0 = 0.f;
          if (u + 1 >= kv_left) e1 = 0.f;
        }
        const __half2 h2 = __floats2half2_rn(e0, e1);
        const float2 back = __half22float2(h2);   // the row sum uses the fp16-rounded probabilities that multiply V
        rs += back.x + back.y;
        pk[u >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      // P(j) buffer: last read by PV(j-2), complete because S(j) was issued after it and s_full(j) covers all prior MMAs
      tmem_st_x16(trow + kHdPCol + bf * (kHdKeys / 2), pk);
      tmem_st_x16(trow + kHdPCol + bf * (kHdKeys / 2) + 16, pk + 16);
      if (j > 0 && __any_sync(0xffffffffu, grow)) {
        // O is being accumulated by PV(j-1), issued after S(j): wait for it before rescaling this row
        mbar_wait(smem_u32(&pv_done), (j - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < kHdOutCols; c0 += 32) {
          uint32_t o[32];
          tmem_ld_x32(trow + kHdOCol + c0, o);
          tmem_ld_wait();
#pragma unroll
          for (int u = 0; u < 32; ++u) o[u] = __float_as_uint(__uint_as_float(o[u]) * alpha);
          tmem_st_x32(trow + kHdOCol + c0, o);
        }
      }
      tmem_st_wait();
      l_run = l_run * alpha + rs;
      tc_fence_before();
      mbar_arrive(smem_u32(&p_full[bf]));
    }
    // epilogue: O / l -> fp16 -> global
    mbar_wait(smem_u32(&pv_done), (nblk - 1) & 1);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const int qi = q0 + row;
    __half* orow = p.out + (static_cast<long long>(b) * p.nq + qi) * p.ldo + p.out_col0 + half * kHdOutCols;
#pragma unroll 1
    for (int c0 = 0; c0 < kHdOutCols; c0 += 32) {
      uint32_t o[32];
      tmem_ld_x32(trow + kHdOCol + c0, o);
      tmem_ld_wait();
      if (qi < p.nq) {
#pragma unroll
        for (int u = 0; u < 32; u += 8) {
          uint4 v;
          v.x = pack_h2(__uint_as_float(o[u]) * inv, __uint_as_float(o[u + 1]) * inv);
          v.y = pack_h2(__uint_as_float(o[u + 2]) * inv, __uint_as_float(o[u + 3]) * inv);
          v.z = pack_h2(__uint_as_float(o[u + 4]) * inv, __uint_as_float(o[u + 5]) * inv);
          v.w = pack_h2(__uint_as_float(o[u + 6]) * inv, __uint_as_float(o[u + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + c0 + u) = v;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// [batch][rows][ld] fp16 matrix viewed slab-major, {64 cols, rows, ld / 64 slabs, batch}: one TMA operation fills `group`
// consecutive 64-column slabs, each landing as its own [rows x 64] swizzled slab in shared memory (a thread issues one TMA
// operation per ~270 cycles whatever its size, so 8 KB operations could not feed the ring: measured 811 / 994 / 1305 us
// per 14400-token frame with 4 / 2 / 1 slabs per operation, profiles/r02_attention_hd512.log)
static int slab_map(CUtensorMap* m, const void* base, int ld, int rows, int batch, int box_rows, int group) {
  uint64_t dims[4] = {64, (uint64_t)rows, (uint64_t)(ld / 64), (uint64_t)batch};
  uint64_t str[3] = {(uint64_t)ld * 2, 128, (uint64_t)ld * 2 * rows};
  uint32_t box[4] = {64, (uint32_t)box_rows, (uint32_t)group, 1};
  return make_tmap_f16(m, base, 4, dims, str, box);
}

int launch_attention_hd512(const mgld_attention_desc* d, cudaStream_t stream) {
  MGLD_CHECK_ARG(d->head_dim == kHdDim, "attention_hd512: head_dim %d", d->head_dim);
  // slab-major views need every operand's head to start on a 64-column slab boundary of a slab-divisible row
  MGLD_CHECK_ARG(d->ldq % 64 == 0 && d->ldk % 64 == 0 && d->ldv % 64 == 0 && d->q_col0 % 64 == 0 && d->k_col0 % 64 == 0 &&
                     d->v_col0 % 64 == 0 && d->q_head_stride % 64 == 0 && d->k_head_stride % 64 == 0 &&
                     d->v_head_stride % 64 == 0,
                 "attention_hd512: row pitches, column offsets and head strides must be multiples of 64");
  const int kvb = d->kv_batched ? d->batch : 1;
  const char* eg = getenv("MGLD_HD512_GROUP");   // development switch, read per call (the tests vary it): 2 or 4
  const int group = (eg && atoi(eg) == 2) ? 2 : 4;
  CUtensorMap tmQ, tmK, tmV;
  int rc = slab_map(&tmQ, d->q, d->ldq, d->nq, d->batch, 128, 4);
  if (!rc) rc = slab_map(&tmK, d->k, d->ldk, d->nkv, kvb, kHdKeys, group);
  if (!rc) rc = slab_map(&tmV, d->v, d->ldv, d->nkv, kvb, kHdKeys, group);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    MGLD_CUDA(cudaFuncSetAttribute(attention_hd512_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHdSmem));
    MGLD_CUDA(cudaFuncSetAttribute(attention_hd512_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHdSmem));
    attr_set = true;
  }
  for (int h = 0; h < d->heads; ++h) {
    Hd512Params p;
    memset(&p, 0, sizeof(p));
    p.nq = d->nq; p.nkv = d->nkv;
    p.q_col0 = d->q_col0 + h * d->q_head_stride;
    p.k_col0 = d->k_col0 + h * d->k_head_stride;
    p.v_col0 = d->v_col0 + h * d->v_head_stride;
    p.kv_batched = d->kv_batched;
    p.scale_log2e = d->scale * 1.4426950408889634f;
    p.out = reinterpret_cast<__half*>(d->out); p.ldo = d->ldo; p.out_col0 = h * kHdDim;
    dim3 grid(ceil_div(d->nq, 128), kHdDim / kHdOutCols, d->batch);
    MGLD_CUDA(launch_pdl(group == 4 ? attention_hd512_kernel<4> : attention_hd512_kernel<2>, grid, dim3(kHdThreads),
                         (size_t)kHdSmem, stream, tmQ, tmK, tmV, p));
    MGLD_LAUNCH_CHECK("attention_hd512_kernel");
  }
  return MGLD_OK;
}

}  // namespace mgld
