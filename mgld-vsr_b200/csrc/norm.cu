// Normalisation kernels (fp16 NHWC activations, fp32 math): GroupNorm statistics / apply(+SiLU), LayerNorm, and the
// row softmax used by the unfused head-dim-512 VAE attention.  All HBM-bound: one read + one write of the activation,
// 16-byte vector accesses, grids sized to cover the 148 SMs several times over.
//
// Replaces: GroupNorm32 (ldm/modules/diffusionmodules/util.py:199-216, eps 1e-5, fp32 math), Normalize
// (ldm/modules/diffusionmodules/model.py:80, ldm/modules/attention.py:87, eps 1e-6), nn.SiLU / nonlinearity
// (model.py:75-77), nn.LayerNorm (attention.py:132,423-425), the softmax inside memory_efficient_attention at
// model.py:294.
#include <math.h>
#include <string.h>

#include "../../include/mgld.h"
#include "common.h"
#include "ptx.cuh"

namespace mgld {

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x; f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}

// load 8 channels starting at channel c of row m from the (virtually concatenated) pair of sources
__device__ __forceinline__ uint4 load_cat8(const __half* x1, int C1, int ld1, const __half* x2, int ld2, long long m,
                                           int c) {
  if (c < C1) return __ldg(reinterpret_cast<const uint4*>(x1 + m * ld1 + c));
  return __ldg(reinterpret_cast<const uint4*>(x2 + m * ld2 + (c - C1)));
}

constexpr int kGnRowsPerBlock = 64;
constexpr int kGnApplyItems = 1024;  // 8-channel vectors per block in gn_apply

// sums[t][g] += (sum, sumsq) over rows [r0, r0+64) of frame t.  Bitwise repeatable: fixed-order reduction inside the block
// (per-thread channel partials -> shared table -> one thread per (group, moment) sums them in a fixed order) and
// order-independent fixed-point accumulation across blocks (fixsum_add, common.h).  `sums`: [T, G, 2] accumulators of two
// 8-byte words each.
__global__ void gn_stats_kernel(const __half* __restrict__ x1, int C1, int ld1, const __half* __restrict__ x2, int C2,
                                int ld2, int HW, int G, double* __restrict__ sums, int rows_per_block) {
  pdl_launch_dependents();
  pdl_wait();

  extern __shared__ float sh[];  // [rpi * vpr][16]: per-thread (sum[8] | sumsq[8]) of its channel vector
  const int C = C1 + C2, vpr = C >> 3, cpg = C / G;
  const int t = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_block;
  const int rows = min(rows_per_block, HW - r0);
  // thread -> fixed channel vector, strided rows: per-channel fp32 partials stay in registers
  const int rpi = blockDim.x / vpr;  // rows handled per iteration (>= 1 because blockDim >= vpr)
  const int v = threadIdx.x % vpr, rr = threadIdx.x / vpr;
  if (rr < rpi) {
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
    int r = rr;
    for (; r + 3 * rpi < rows; r += 4 * rpi) {   // 4 independent 16-byte loads in flight per thread
      uint4 u[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        u[j] = load_cat8(x1, C1, ld1, x2, ld2, static_cast<long long>(t) * HW + r0 + r + j * rpi, v * 8);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float f[8];
        unpack8(u[j], f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    for (; r < rows; r += rpi) {
      float f[8];
      unpack8(load_cat8(x1, C1, ld1, x2, ld2, static_cast<long long>(t) * HW + r0 + r, v * 8), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
    }
    float4* dst = reinterpret_cast<float4*>(sh + (rr * vpr + v) * 16);
    dst[0] = make_float4(s[0], s[1], s[2], s[3]);
    dst[1] = make_float4(s[4], s[5], s[6], s[7]);
    dst[2] = make_float4(q[0], q[1], q[2], q[3]);
    dst[3] = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {   // i = 2 * group + moment
    const int g = i >> 1, mo = (i & 1) * 8;
    float acc = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
      const float* col = sh + (c >> 3) * 16 + mo + (c & 7);
      for (int k = 0; k < rpi; ++k) acc += col[k * vpr * 16];
    }
    fixsum_add(sums + (static_cast<long long>(t) * 2 * G + i) * 2, acc);
  }
}

// (sum, sumsq) -> (mean, rstd) fp32
__global__ void gn_finalize_kernel(const double* __restrict__ sums, float* __restrict__ stats, int n, double count,
                                   double eps) {
  pdl_launch_dependents();
  pdl_wait();

  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mean = fixsum_load(sums + 4 * i) / count;
  double var = fixsum_load(sums + 4 * i + 2) / count - mean * mean;
  if (var < 0.0) var = 0.0;
  stats[2 * i] = (float)mean;
  stats[2 * i + 1] = (float)(1.0 / sqrt(var + eps));
}

// y = act( (x - mean) * rstd * gamma + beta )
__global__ void gn_apply_kernel(const __half* __restrict__ x1, int C1, int ld1, const __half* __restrict__ x2, int C2,
                                int ld2, int HW, int G, const double* __restrict__ sums, double eps,
                                const float* __restrict__ gamma, const float* __restrict__ beta, int silu,
                                __half* __restrict__ out, int ldo) {
  pdl_launch_dependents();
  pdl_wait();

  extern __shared__ float sh[];  // [2*G] mean, rstd
  const int C = C1 + C2, vpr = C >> 3, cpg = C / G;
  const int t = blockIdx.y;
  // each block covers kGnApplyItems consecutive 8-channel vectors of frame t.  All of a thread's loads are issued before
  // the first one is consumed (an early `break` on the bounds check inside the unrolled loop used to serialise them: one
  // 16-byte load in flight per thread, ~2.5 TB/s on L2-resident maps).
  // A block owns a contiguous range of the frame's vectors; the launcher sizes the ranges so that the whole grid is
  // resident at once (1600 blocks on 1184 slots ran as two waves, the second one a third full).
  const long long total = static_cast<long long>(HW) * vpr;
  const long long per = (total + gridDim.x - 1) / gridDim.x;
  const long long base0 = static_cast<long long>(blockIdx.x) * per;
  const long long end = base0 + per < total ? base0 + per : total;
  constexpr int kPer = kGnApplyItems / 256;
  for (long long base = base0; base < end; base += kGnApplyItems) {
  uint4 u[kPer];
  int rr[kPer], vv[kPer];
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    const long long idx = base + j * 256 + threadIdx.x;
    const bool live = idx < end;
    rr[j] = live ? static_cast<int>(idx / vpr) : -1;
    vv[j] = live ? static_cast<int>(idx - static_cast<long long>(rr[j]) * vpr) : 0;
    u[j] = live ? load_cat8(x1, C1, ld1, x2, ld2, static_cast<long long>(t) * HW + rr[j], vv[j] * 8) : make_uint4(0, 0, 0, 0);
  }
  if (base == base0) {   // (mean, rstd) of the frame's groups, while the first loads are in flight
    const double count = (double)HW * cpg;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
      const double mean = fixsum_load(sums + (static_cast<long long>(t) * G + g) * 4) / count;
      double var = fixsum_load(sums + (static_cast<long long>(t) * G + g) * 4 + 2) / count - mean * mean;
      if (var < 0.0) var = 0.0;
      sh[2 * g] = (float)mean;
      sh[2 * g + 1] = (float)(1.0 / sqrt(var + eps));
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < kPer; ++j) {
    if (rr[j] < 0) continue;
    const int v = vv[j];
    const long long m = static_cast<long long>(t) * HW + rr[j];
    float f[8];
    unpack8(u[j], f);
    float gm[8], bt[8];
    if (gamma) {
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), gb = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), bb = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      gm[0] = ga.x; gm[1] = ga.y; gm[2] = ga.z; gm[3] = ga.w; gm[4] = gb.x; gm[5] = gb.y; gm[6] = gb.z; gm[7] = gb.w;
      bt[0] = ba.x; bt[1] = ba.y; bt[2] = ba.z; bt[3] = ba.w; bt[4] = bb.x; bt[5] = bb.y; bt[6] = bb.z; bt[7] = bb.w;
    }
    if (cpg >= 4) {
      // a vector of 8 channels touches at most two groups: [0, nb) in g, the rest in g + 1
      const int g = (v * 8) / cpg, nb = (g + 1) * cpg - v * 8;
      const float m0 = sh[2 * g], r0s = sh[2 * g + 1];
      const float m1 = nb < 8 ? sh[2 * g + 2] : 0.f, r1s = nb < 8 ? sh[2 * g + 3] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = i < nb ? (f[i] - m0) * r0s : (f[i] - m1) * r1s;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const int g = (v * 8 + i) / cpg; f[i] = (f[i] - sh[2 * g]) * sh[2 * g + 1]; }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = f[i];
      if (gamma) y = fmaf(y, gm[i], bt[i]);
      if (silu) y = __fdividef(y, 1.f + __expf(-y));   // MUFU.RCP + FMUL instead of the ~8-instruction IEEE division
      f[i] = y;
    }
    *reinterpret_cast<uint4*>(out + m * ldo + v * 8) = pack8(f);
  }
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Single-launch GroupNorm (+SiLU): statistics and normalisation in ONE pass over HBM.  A CTA owns a (rows x channels)
// patch of one frame -- `cc` channels (whole groups, whole 16-byte vectors) of `rp` rows -- and keeps it in REGISTERS
// (thread = fixed 8-channel column, up to kItems rows).  The `cs` CTAs that share a channel range of a frame form a
// thread-block cluster: partial (sum, sumsq) per group go to each CTA's shared memory, every CTA sums all partials over
// distributed shared memory in the same fixed order (so all agree bit-for-bit), then normalises its registers and stores.
// Replaces zero-fill + gn_stats + (gn_finalize) + gn_apply: 3-4 launches and a second read of the activation.
// ---------------------------------------------------------------------------------------------------------------------
struct GnFusedParams {
  const __half* x1; int C1, ld1;
  const __half* x2; int C2, ld2;
  int HW, G, cpg;
  int cc, vpc, R, rp, cs;   // channels per CTA, 8-channel vectors per row of the patch, row lanes, rows per CTA, cluster size
  double eps;
  const float* gamma; const float* beta;
  int silu;
  __half* out; int ldo;     // may be null: statistics only
  float* stats_out;         // may be null; (mean, rstd) [T, G, 2] for the SPADE epilogue of conv_gemm
};
constexpr int kGnFusedThreads = 512;
constexpr int kGnFusedMaxGroups = 64;   // groups per CTA patch

__device__ __forceinline__ float ld_shared_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int kItems>
__global__ void __launch_bounds__(kGnFusedThreads, 1) gn_fused_kernel(const GnFusedParams p) {
  pdl_launch_dependents();
  pdl_wait();

  extern __shared__ float tab[];                  // [2][R][cc]: per-thread channel partials (sum | sumsq)
  __shared__ float chs[2 * kGnFusedThreads];      // per-channel (sum | sumsq) of this CTA's patch
  __shared__ float part[2 * kGnFusedMaxGroups];   // this CTA's partial (sum, sumsq) per local group
  __shared__ float stat[2 * kGnFusedMaxGroups];   // (mean, rstd) per local group
  const int tid = threadIdx.x;
  const int rank = p.cs > 1 ? static_cast<int>(cluster_ctarank()) : 0;
  const int t = blockIdx.z;
  const int c_lo = blockIdx.y * p.cc;
  const int r_lo = rank * p.rp;
  const int rows = max(0, min(p.rp, p.HW - r_lo));
  const int gpc = p.cc / p.cpg;
  const int v = tid % p.vpc, rr = tid / p.vpc;
  const bool active = rr < p.R;
  const int c0 = c_lo + v * 8;   // first of this thread's 8 channels (virtual concat space)
  uint4 reg[kItems];
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; q[i] = 0.f; }
  const long long m0 = static_cast<long long>(t) * p.HW + r_lo;
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int r = rr + j * p.R;
    if (active && r < rows) reg[j] = load_cat8(p.x1, p.C1, p.ld1, p.x2, p.ld2, m0 + r, c0);
    else reg[j] = make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    float f[8];
    unpack8(reg[j], f);   // rows beyond the patch were zero-filled: they add nothing
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
  }
  // Fixed-order tree (no atomics: hundreds of threads adding into one or two shared addresses serialise, and the result
  // would depend on the order): thread partials -> table -> per-channel sums over the row lanes -> per-group sums.
  float* ts = tab;
  float* tq = tab + p.R * p.cc;
  if (active) {
    float4* ds = reinterpret_cast<float4*>(ts + rr * p.cc + v * 8);
    float4* dq = reinterpret_cast<float4*>(tq + rr * p.cc + v * 8);
    ds[0] = make_float4(s[0], s[1], s[2], s[3]); ds[1] = make_float4(s[4], s[5], s[6], s[7]);
    dq[0] = make_float4(q[0], q[1], q[2], q[3]); dq[1] = make_float4(q[4], q[5], q[6], q[7]);
  }
  __syncthreads();
  for (int c = tid; c < p.cc; c += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
    int r = 0;
    for (; r + 1 < p.R; r += 2) {
      a0 += ts[r * p.cc + c]; a1 += ts[(r + 1) * p.cc + c];
      b0 += tq[r * p.cc + c]; b1 += tq[(r + 1) * p.cc + c];
    }
    if (r < p.R) { a0 += ts[r * p.cc + c]; b0 += tq[r * p.cc + c]; }
    chs[c] = a0 + a1; chs[kGnFusedThreads + c] = b0 + b1;
  }
  __syncthreads();
  if (tid < gpc) {
    float a = 0.f, b = 0.f;
    for (int i = 0; i < p.cpg; ++i) { a += chs[tid * p.cpg + i]; b += chs[kGnFusedThreads + tid * p.cpg + i]; }
    part[2 * tid] = a; part[2 * tid + 1] = b;
  }
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();   // every CTA's partials are complete and visible cluster-wide
  if (tid < gpc) {
    double sum = 0.0, sq = 0.0;
    if (p.cs > 1) {
      for (int k = 0; k < p.cs; ++k) {
        sum += (double)ld_shared_cluster_f32(mapa_shared(smem_u32(&part[2 * tid]), k));
        sq += (double)ld_shared_cluster_f32(mapa_shared(smem_u32(&part[2 * tid + 1]), k));
      }
    } else {
      sum = part[2 * tid]; sq = part[2 * tid + 1];
    }
    const double count = (double)p.HW * p.cpg;
    const double mean = sum / count;
    double var = sq / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float mf = (float)mean, rf = (float)(1.0 / sqrt(var + p.eps));
    stat[2 * tid] = mf; stat[2 * tid + 1] = rf;
    if (p.stats_out && rank == 0) {
      float* so = p.stats_out + (static_cast<long long>(t) * p.G + c_lo / p.cpg + tid) * 2;
      so[0] = mf; so[1] = rf;
    }
  }
  __syncthreads();
  if (p.cs > 1) cluster_sync_all();   // nobody leaves (or reuses `part`) while a peer may still be reading it
  if (!p.out || !active) return;

  float sc[8], sh[8];   // y = x * sc + sh
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (v * 8 + i) / p.cpg;
    const float mean = stat[2 * g], rstd = stat[2 * g + 1];
    const float ga = p.gamma ? __ldg(p.gamma + c0 + i) : 1.f, be = p.gamma ? __ldg(p.beta + c0 + i) : 0.f;
    sc[i] = rstd * ga;
    sh[i] = fmaf(-mean, sc[i], be);
  }
#pragma unroll
  for (int j = 0; j < kItems; ++j) {
    const int r = rr + j * p.R;
    if (r < rows) {
      float f[8];
      unpack8(reg[j], f);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float y = fmaf(f[i], sc[i], sh[i]);
        if (p.silu) y = __fdividef(y, 1.f + __expf(-y));
        f[i] = y;
      }
      *reinterpret_cast<uint4*>(p.out + (m0 + r) * p.ldo + c0) = pack8(f);
    }
  }
}

// LayerNorm over the last dim, one warp per row (C <= 2048, multiple of 8).  kVec = 16-byte vectors per lane, compile
// time: the row lives in kVec * 8 registers (C = 320 needs 2, not the generic 8), which is what lets enough warps be
// resident to cover the load -> reduce -> reduce -> store latency chain of each row.
constexpr int kLnMaxVec = 8;  // vectors of 8 per lane
template <int kVec>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, int ldx, int M, int C, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, __half* __restrict__ out, int ldo) {
  pdl_launch_dependents();
  pdl_wait();

  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= M) return;
  const int vpr = C >> 3;
  float f[kVec][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const int v = lane + 32 * k;
    if (v < vpr) {
      unpack8(__ldg(reinterpret_cast<const uint4*>(x + static_cast<long long>(warp) * ldx + v * 8)), f[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += f[k][i];
    }
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    if (lane + 32 * k < vpr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = f[k][i] - mean; q = fmaf(d, d, q); }
    }
  }
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
  for (int k = 0; k < kVec; ++k) {
    const int v = lane + 32 * k;
    if (v < vpr) {
      float y[8];
      const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + v * 8)), gb = __ldg(reinterpret_cast<const float4*>(gamma + v * 8 + 4));
      const float4 ba = __ldg(reinterpret_cast<const float4*>(beta + v * 8)), bb = __ldg(reinterpret_cast<const float4*>(beta + v * 8 + 4));
      const float gmv[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float btv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaf((f[k][i] - mean) * rstd, gmv[i], btv[i]);
      *reinterpret_cast<uint4*>(out + static_cast<long long>(warp) * ldo + v * 8) = pack8(y);
    }
  }
}

// P[r, :] = softmax(scale * S[r, :]) -> fp16; one block per row, fp32 scores (n up to ~16k)
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long lds, int n, float scale,
                                    __half* __restrict__ p, long long ldp) {
  __shared__ float red[32];
  const float* row = s + blockIdx.x * lds;
  __half* prow = p + blockIdx.x * ldp;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) mx = fmaxf(mx, row[i]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = red[0];
  for (int i = 1; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) sum += __expf((row[i] - mx) * scale);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  for (int i = threadIdx.x; i < n; i += blockDim.x) prow[i] = __float2half_rn(__expf((row[i] - mx) * scale) * inv);
}

}  // namespace mgld

using namespace mgld;

static int gn_threads(int C) {
  const int vpr = C / 8;
  int th = 256;
  while (th < vpr) th += 32;
  return th;
}

extern "C" int mgld_gn_stats_f16(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int T, int HW,
                                 int groups, double* sums, void* stream) {
  const int C = C1 + C2;
  MGLD_CHECK_ARG(x1 && sums && T > 0 && HW > 0 && groups > 0, "gn_stats: bad arguments");
  MGLD_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0 && C % groups == 0 && C / 8 <= 704, "gn_stats: C1=%d C2=%d G=%d", C1, C2,
                 groups);
  MGLD_CHECK_ARG((C2 > 0) == (x2 != nullptr), "gn_stats: x2/C2 mismatch");
  // ~2 resident blocks per SM over all frames: long enough per block to keep 4 x 16-byte loads per thread in flight for
  // many iterations (640 blocks of 64 rows were launch- / latency-bound at 1.6 TB/s on L2-resident maps), and few
  // cross-block accumulations
  int bpf = ceil_div(2 * num_sms(), T);
  const int max_bpf = ceil_div(HW, kGnRowsPerBlock);
  if (bpf > max_bpf) bpf = max_bpf;
  if (bpf < 1) bpf = 1;
  const int rpb = ceil_div(ceil_div(HW, bpf), 8) * 8;
  dim3 grid(ceil_div(HW, rpb), T);
  MGLD_CUDA(launch_pdl(gn_stats_kernel, grid, dim3(gn_threads(C)), gn_threads(C) * 16 * sizeof(float), (cudaStream_t)stream,
                       (const __half*)x1, C1, ld1 > 0 ? ld1 : C1, (const __half*)x2, C2, ld2 > 0 ? ld2 : C2, HW, groups, sums, rpb));
  MGLD_LAUNCH_CHECK("gn_stats_kernel");
  return MGLD_OK;
}

extern "C" int mgld_gn_finalize(const double* sums, float* stats, int T, int groups, int HW, int C, double eps,
                                void* stream) {
  MGLD_CHECK_ARG(sums && stats && T > 0 && groups > 0, "gn_finalize: bad arguments");
  const int n = T * groups;
  MGLD_CUDA(launch_pdl(gn_finalize_kernel, dim3(ceil_div(n, 128)), dim3(128), 0, (cudaStream_t)stream, sums, stats, n,
                       (double)HW * (C / groups), eps));
  MGLD_LAUNCH_CHECK("gn_finalize_kernel");
  return MGLD_OK;
}

extern "C" int mgld_gn_apply_f16(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int T, int HW,
                                 int groups, const double* sums, double eps, const float* gamma, const float* beta,
                                 int silu, void* out, int ldo, void* stream) {
  const int C = C1 + C2;
  MGLD_CHECK_ARG(x1 && sums && out && T > 0 && HW > 0 && groups > 0, "gn_apply: bad arguments");
  MGLD_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0 && C % groups == 0, "gn_apply: C1=%d C2=%d G=%d", C1, C2, groups);
  MGLD_CHECK_ARG((gamma != nullptr) == (beta != nullptr), "gn_apply: gamma/beta");
  long long bx = (static_cast<long long>(HW) * (C / 8) + kGnApplyItems - 1) / kGnApplyItems;
  static int occ = -1;   // resident blocks per SM (register-limited)
  if (occ < 0) {
    int n = 0;
    MGLD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, gn_apply_kernel, 256, 2 * groups * sizeof(float)));
    occ = n > 0 ? n : 1;
  }
  const long long resident = static_cast<long long>(num_sms()) * occ / T;   // one wave: T frames share the machine
  if (bx > resident && resident >= 1) bx = resident;
  dim3 grid((unsigned)bx, T);
  MGLD_CUDA(launch_pdl(gn_apply_kernel, grid, dim3(256), 2 * groups * sizeof(float), (cudaStream_t)stream,
                       (const __half*)x1, C1, ld1 > 0 ? ld1 : C1, (const __half*)x2, C2, ld2 > 0 ? ld2 : C2, HW, groups, sums, eps,
                       gamma, beta, silu, (__half*)out, ldo > 0 ? ldo : C));
  MGLD_LAUNCH_CHECK("gn_apply_kernel");
  return MGLD_OK;
}

// plan of the single-launch GroupNorm; false = the patch does not fit in registers (use the multi-pass kernels)
static bool gn_fused_plan(int C, int T, int HW, int G, GnFusedParams* p, int* items, int* csplit) {
  if (C % 8 || G <= 0 || C % G) return false;
  // Measured (profiles/r01_perf_norm.log): one CTA per SM holding its patch in registers serialises load -> reduce ->
  // cluster barrier -> store, so for the 32x32 / 64x64 maps the two streaming kernels win; for <= 16x16 maps (latency-
  // bound, 3 launches -> 1) the single launch is 25-45% faster.
  if (HW > 256) return false;
  const int cpg = C / G;
  int a = 8, b = cpg;
  while (b) { const int r = a % b; a = b; b = r; }
  const int L = 8 / a * cpg;   // lcm(8, cpg): whole groups and whole 16-byte vectors
  if (C % L) return false;
  int cc = L;
  while (cc < 64 && C % (cc * 2) == 0) cc *= 2;
  if (cc / cpg > kGnFusedMaxGroups || cc > kGnFusedThreads) return false;
  const int vpc = cc / 8, R = kGnFusedThreads / vpc;
  auto need = [&](int cs) { return ceil_div(ceil_div(HW, cs), R); };
  int cs = 1;
  while (cs < 8 && need(cs) > 12) cs *= 2;
  if (need(cs) > 12) return false;
  const int ctas = T * (C / cc);
  while (cs < 8 && ctas * cs < num_sms() && need(cs) > 1) cs *= 2;   // spread a small problem over the machine
  p->cpg = cpg; p->cc = cc; p->vpc = vpc; p->R = R; p->cs = cs; p->rp = ceil_div(HW, cs);
  *items = need(cs); *csplit = C / cc;
  return true;
}

extern "C" int mgld_group_norm_fused_supported(int C, int T, int HW, int groups) {
  GnFusedParams p;
  int items, csplit;
  return initialised() && gn_fused_plan(C, T, HW, groups, &p, &items, &csplit) ? 1 : 0;
}

extern "C" int mgld_group_norm_f16(const void* x1, int C1, int ld1, const void* x2, int C2, int ld2, int T, int HW,
                                   int groups, double eps, const float* gamma, const float* beta, int silu, void* out,
                                   int ldo, float* stats_out, double* scratch, void* stream) {
  const int C = C1 + C2;
  MGLD_CHECK_ARG(x1 && (out || stats_out) && T > 0 && HW > 0 && groups > 0, "group_norm: bad arguments");
  MGLD_CHECK_ARG(C1 % 8 == 0 && C2 % 8 == 0 && C % groups == 0 && (!out || ldo % 8 == 0), "group_norm: C1=%d C2=%d G=%d", C1, C2, groups);
  MGLD_CHECK_ARG((C2 > 0) == (x2 != nullptr) && ((gamma != nullptr) == (beta != nullptr)), "group_norm: x2/C2 or gamma/beta mismatch");
  GnFusedParams p;
  memset(&p, 0, sizeof(p));
  int items = 0, csplit = 0;
  if (!gn_fused_plan(C, T, HW, groups, &p, &items, &csplit)) {
    // multi-pass path: the caller's zeroed scratch ([T, groups, 2] accumulators of 16 bytes) carries the sums between the kernels
    MGLD_CHECK_ARG(scratch, "group_norm: this shape needs the [T, groups, 2] x 16-byte scratch (zeroed)");
    int rc = mgld_gn_stats_f16(x1, C1, ld1, x2, C2, ld2, T, HW, groups, scratch, stream);
    if (rc) return rc;
    if (stats_out) { rc = mgld_gn_finalize(scratch, stats_out, T, groups, HW, C, eps, stream); if (rc) return rc; }
    if (out) rc = mgld_gn_apply_f16(x1, C1, ld1, x2, C2, ld2, T, HW, groups, scratch, eps, gamma, beta, silu, out, ldo, stream);
    return rc;
  }
  p.x1 = (const __half*)x1; p.C1 = C1; p.ld1 = ld1 > 0 ? ld1 : C1;
  p.x2 = (const __half*)x2; p.C2 = C2; p.ld2 = ld2 > 0 ? ld2 : C2;
  p.HW = HW; p.G = groups; p.eps = eps; p.gamma = gamma; p.beta = beta; p.silu = silu;
  p.out = (__half*)out; p.ldo = ldo; p.stats_out = stats_out;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.gridDim = dim3(p.cs, csplit, T);
  cfg.blockDim = dim3(p.vpc * p.R, 1, 1);
  cfg.dynamicSmemBytes = 2 * p.R * p.cc * sizeof(float);   // <= 32 KB (R * cc <= 512 * 8)
  cfg.stream = (cudaStream_t)stream;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (items <= 2) MGLD_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel<2>, p));
  else if (items <= 4) MGLD_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel<4>, p));
  else if (items <= 8) MGLD_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel<8>, p));
  else MGLD_CUDA(cudaLaunchKernelEx(&cfg, gn_fused_kernel<12>, p));
  MGLD_LAUNCH_CHECK("gn_fused_kernel");
  return MGLD_OK;
}

extern "C" int mgld_layernorm_f16(const void* x, int ldx, int M, int C, const float* gamma, const float* beta,
                                  float eps, void* out, int ldo, void* stream) {
  MGLD_CHECK_ARG(x && out && gamma && beta && M > 0, "layernorm: bad arguments");
  MGLD_CHECK_ARG(C % 8 == 0 && C / 8 <= 32 * kLnMaxVec, "layernorm: C=%d unsupported", C);
  const int warps_per_block = 8;
  const int vec = ceil_div(C / 8, 32);
  const dim3 grid(ceil_div(M, warps_per_block)), block(warps_per_block * 32);
  cudaStream_t st = (cudaStream_t)stream;
  const __half* xp = (const __half*)x;
  __half* op = (__half*)out;
  const int lx = ldx > 0 ? ldx : C, lo = ldo > 0 ? ldo : C;
  switch (vec) {
    case 1: MGLD_CUDA(launch_pdl(layernorm_kernel<1>, grid, block, 0, st, xp, lx, M, C, gamma, beta, eps, op, lo)); break;
    case 2: MGLD_CUDA(launch_pdl(layernorm_kernel<2>, grid, block, 0, st, xp, lx, M, C, gamma, beta, eps, op, lo)); break;
    case 3: MGLD_CUDA(launch_pdl(layernorm_kernel<3>, grid, block, 0, st, xp, lx, M, C, gamma, beta, eps, op, lo)); break;
    case 4: case 5: MGLD_CUDA(launch_pdl(layernorm_kernel<5>, grid, block, 0, st, xp, lx, M, C, gamma, beta, eps, op, lo)); break;
    default: MGLD_CUDA(launch_pdl(layernorm_kernel<kLnMaxVec>, grid, block, 0, st, xp, lx, M, C, gamma, beta, eps, op, lo)); break;
  }
  MGLD_LAUNCH_CHECK("layernorm_kernel");
  return MGLD_OK;
}

extern "C" int mgld_softmax_rows_f32(const float* s, long long lds, int rows, int n, float scale, void* p,
                                     long long ldp, void* stream) {
  MGLD_CHECK_ARG(s && p && rows > 0 && n > 0, "softmax_rows: bad arguments");
  softmax_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(s, lds, n, scale, (__half*)p, ldp);
  MGLD_LAUNCH_CHECK("softmax_rows_kernel");
  return MGLD_OK;
}
