// Cross-attention against a SHORT shared context (nkv <= 80 keys, head dim 64): O = softmax(scale * Q K^T) V with the
// K / V of one head resident in shared memory and everything else in registers.
//
// The UNet's 16 cross-attention layers attend to the 77 text tokens (ldm/modules/attention.py:371, K / V broadcast over
// the frames as :336-337).  That is 2 x 77 x 64 MACs per query row and head - 4 GFLOP per launch at the 64x64 level against
// 52 MB of Q + O traffic: a memory-bound op.  The tcgen05 flash kernel (attention.cu) runs it as 800 one-block CTAs, one
// per SM at a time, each a serial chain TMA -> MMA -> TMEM load -> softmax -> MMA -> store (39 us at T = 10, 4x the memory
// floor).  Here a warp owns 16 query rows at a time: Q fragments via ldmatrix, S = Q K^T and O = P V as warp-level
// mma.sync.m16n8k16 (the accumulator layout of S IS the A-operand layout of P, so P never leaves registers), row max / sum
// by two quad shuffles, 16 warps per SM hide each other's latencies.  The tensor work is too small for tcgen05 to matter.
#include <stdlib.h>
#include <string.h>

#include "../../include/mgld.h"
#include "common.h"
#include "ptx.cuh"

namespace mgld {

constexpr int kXThreads = 256;     // 8 warps, each looping over 16-row query tiles
constexpr int kXKeys = 80;         // keys held (10 n-tiles of 8; rows >= nkv are zero and masked)
constexpr int kXTileBytes = 16 * 128;  // one warp's [16 rows x 64 halfs] staging tile (Q in, O out)

struct CrossParams {
  const __half* q; const __half* k; const __half* v; __half* out;
  int ldq, ldk, ldv, ldo;
  int q_col0, k_col0, v_col0, q_hstride, k_hstride, v_hstride;
  int rows;       // batch * nq query rows (K / V are shared by the batches)
  int nkv;
  float scale_log2e;
};

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk `chunk` of row `row` in a [rows x 128 B] tile; the XOR spreads the 8 rows an ldmatrix
// phase reads over all banks
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(kXThreads, 2) cross_attention_kv80_kernel(const CrossParams p) {
  __shared__ __align__(128) uint8_t sK[kXKeys * 128];
  __shared__ __align__(128) uint8_t sV[kXKeys * 128];
  __shared__ __align__(128) uint8_t sT[(kXThreads / 32) * kXTileBytes];
  pdl_launch_dependents();
  pdl_wait();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  // K / V of this head -> shared memory (rows >= nkv zero: 0 * garbage in P V would poison O)
  for (int i = threadIdx.x; i < kXKeys * 8; i += kXThreads) {
    const int key = i >> 3, ch = i & 7;
    uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
    if (key < p.nkv) {
      kk = __ldg(reinterpret_cast<const uint4*>(p.k + static_cast<size_t>(key) * p.ldk + p.k_col0 + head * p.k_hstride + ch * 8));
      vv = __ldg(reinterpret_cast<const uint4*>(p.v + static_cast<size_t>(key) * p.ldv + p.v_col0 + head * p.v_hstride + ch * 8));
    }
    *reinterpret_cast<uint4*>(sK + swz(key, ch)) = kk;
    *reinterpret_cast<uint4*>(sV + swz(key, ch)) = vv;
  }
  __syncthreads();

  uint8_t* tile = sT + warp * kXTileBytes;
  const uint32_t tile_s = smem_u32(tile), sK_s = smem_u32(sK), sV_s = smem_u32(sV);
  const int ntiles = (p.rows + 15) >> 4;
  const int g = lane >> 2, c = lane & 3;
  const __half* qh = p.q + p.q_col0 + head * p.q_hstride;
  __half* oh = p.out + head * 64;

  // this lane's four 16-byte pieces of a [16 x 64] tile: piece i = row (i*32 + lane) / 8, chunk lane % 8
  auto load_q = [&](int t, uint4* r) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i * 4 + (lane >> 3);
      const long long m = static_cast<long long>(t) * 16 + row;
      r[i] = m < p.rows ? __ldg(reinterpret_cast<const uint4*>(qh + m * p.ldq + (lane & 7) * 8)) : make_uint4(0, 0, 0, 0);
    }
  };

  int t = blockIdx.x * (kXThreads / 32) + warp;
  const int tstep = gridDim.x * (kXThreads / 32);
  uint4 qn[4];
  if (t < ntiles) load_q(t, qn);
  for (; t < ntiles; t += tstep) {
    // ---- Q tile -> shared -> A fragments (4 k-steps of 16 head-dim columns)
    __syncwarp();   // the previous tile's output has left the staging tile
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(tile + swz(i * 4 + (lane >> 3), lane & 7)) = qn[i];
    __syncwarp();
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
      ldmatrix_x4(tile_s + swz((lane & 7) + ((lane >> 3) & 1) * 8, 2 * ks + (lane >> 4)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    if (t + tstep < ntiles) load_q(t + tstep, qn);   // in flight during the math below

    // ---- S = Q K^T: 10 n-tiles of 8 keys; thread holds (row g, keys 8j + 2c, +1) in s[j][0..1], row g + 8 in s[j][2..3]
    float s[kXKeys / 8][4];
#pragma unroll
    for (int j = 0; j < kXKeys / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int u = 0; u < 2; ++u) {   // one ldmatrix.x4 = the B fragments of two k-steps
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(sK_s + swz(8 * j + (lane & 7), 4 * u + (lane >> 3)), b0, b1, b2, b3);
        mma_16816(s[j], qa[2 * u], b0, b1);
        mma_16816(s[j], qa[2 * u + 1], b2, b3);
      }
    }
    // ---- softmax over the keys (fp32): scale, mask, row max, exponentials, row sum
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < kXKeys / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = 8 * j + 2 * c + (e & 1);
        s[j][e] = key < p.nkv ? s[j][e] * p.scale_log2e : -INFINITY;
      }
      m0 = fmaxf(m0, fmaxf(s[j][0], s[j][1]));
      m1 = fmaxf(m1, fmaxf(s[j][2], s[j][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[kXKeys / 16][4];   // P as the A operand of P V: k-step ks = keys 16 ks .. 16 ks + 15 = n-tiles 2 ks, 2 ks + 1
#pragma unroll
    for (int j = 0; j < kXKeys / 8; ++j) {
      const float e0 = ex2_approx(s[j][0] - m0), e1 = ex2_approx(s[j][1] - m0);
      const float e2 = ex2_approx(s[j][2] - m1), e3 = ex2_approx(s[j][3] - m1);
      l0 += e0 + e1; l1 += e2 + e3;
      pa[j >> 1][(j & 1) * 2] = pack_h2(e0, e1);
      pa[j >> 1][(j & 1) * 2 + 1] = pack_h2(e2, e3);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);

    // ---- O = P V: 8 n-tiles of 8 head-dim columns, 5 k-steps of 16 keys; V row-major [key, d] -> ldmatrix.trans
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < kXKeys / 16; ++ks) {
#pragma unroll
      for (int n = 0; n < 8; n += 2) {
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(sV_s + swz(16 * ks + (lane & 7) + ((lane >> 3) & 1) * 8, n + (lane >> 4)), b0, b1, b2, b3);
        mma_16816(o[n], pa[ks], b0, b1);
        mma_16816(o[n + 1], pa[ks], b2, b3);
      }
    }
    // ---- normalise, stage the [16 x 64] fp16 tile, write 16-byte pieces (whole 128-byte rows per 8 lanes)
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    __syncwarp();   // every lane's ldmatrix of the Q tile is done
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      *reinterpret_cast<uint32_t*>(tile + swz(g, n) + 4 * c) = pack_h2(o[n][0] * i0, o[n][1] * i0);
      *reinterpret_cast<uint32_t*>(tile + swz(g + 8, n) + 4 * c) = pack_h2(o[n][2] * i1, o[n][3] * i1);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i * 4 + (lane >> 3);
      const long long m = static_cast<long long>(t) * 16 + row;
      if (m < p.rows)
        *reinterpret_cast<uint4*>(oh + m * p.ldo + (lane & 7) * 8) = *reinterpret_cast<const uint4*>(tile + swz(row, lane & 7));
    }
  }
}

bool attention_kv80_supported(const mgld_attention_desc* d) {
  static const bool off = [] { const char* e = getenv("MGLD_ATTN_KV80"); return e && atoi(e) == 0; }();
  return !off && d->head_dim == 64 && d->nkv <= kXKeys && !d->kv_batched;
}

int launch_attention_kv80(const mgld_attention_desc* d, cudaStream_t stream) {
  CrossParams p;
  memset(&p, 0, sizeof(p));
  p.q = reinterpret_cast<const __half*>(d->q); p.k = reinterpret_cast<const __half*>(d->k);
  p.v = reinterpret_cast<const __half*>(d->v); p.out = reinterpret_cast<__half*>(d->out);
  p.ldq = d->ldq; p.ldk = d->ldk; p.ldv = d->ldv; p.ldo = d->ldo;
  p.q_col0 = d->q_col0; p.k_col0 = d->k_col0; p.v_col0 = d->v_col0;
  p.q_hstride = d->q_head_stride; p.k_hstride = d->k_head_stride; p.v_hstride = d->v_head_stride;
  p.rows = d->batch * d->nq; p.nkv = d->nkv;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  // 16-byte vector accesses: pitches and column offsets are multiples of 8 elements (checked by mgld_attention), head strides too
  MGLD_CHECK_ARG(d->q_head_stride % 8 == 0 && d->k_head_stride % 8 == 0 && d->v_head_stride % 8 == 0,
                 "attention: head strides must be multiples of 8 elements");
  const int ntiles = ceil_div(p.rows, 16), per_cta = kXThreads / 32;
  int gx = ceil_div(ntiles, per_cta);
  const int resident = 2 * num_sms() / d->heads;   // two CTAs per SM: one wave of persistent CTAs when there is more work
  if (resident >= 1 && gx > resident) gx = resident;
  MGLD_CUDA(launch_pdl(cross_attention_kv80_kernel, dim3(gx, d->heads), dim3(kXThreads), 0, stream, p));
  MGLD_LAUNCH_CHECK("cross_attention_kv80_kernel");
  return MGLD_OK;
}

}  // namespace mgld
