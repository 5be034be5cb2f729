// Implicit-GEMM convolution / linear layer for sm_100a — persistent, warp-specialised, TMA in / TMA out.
//
//   out[m, n] = epilogue( sum_{tap, c} A[pixel(m) + off(tap), c] * W[n, tap*C + c] )
//
// A tile is 128 output pixels x BLOCK_N channels.  The 128 rows are a (BW x BH x BT) box of pixels of the NHWC
// activation, so for every filter tap the A operand is ONE TMA box load at shifted coordinates: the TMA unit does the
// im2col and its out-of-bounds zero fill is the convolution's zero padding (spatial and temporal).  The same box
// geometry is used by the epilogue's TMA loads (residual / SPADE input) and TMA stores, which also clip partial tiles.
//
// Each CTA (one per SM) loops over tiles.  Warp roles:
//   warp 0      TMA producer of A (one elected thread): A ring (3-8 stages) + the residual tile of the current output tile
//   warp 1      tcgen05.mma issuer; the fp32 accumulator is double-buffered in TMEM (2 x BLOCK_N columns), so the
//               mainloop of tile i+1 overlaps the epilogue of tile i
//   warps 6-7   TMA producers of the weight tile (half of BLOCK_N rows each)
//   warps 2-5, 8-11  two epilogue warpgroups, thread == accumulator row, alternating over the 64-column staging panels:
//               tcgen05.ld 32 columns -> (+bias, activation, residual / GEGLU / SPADE math in fp32) -> fp16 ->
//               128B-swizzled staging panel in smem -> TMA store of the panel while the next one is converted
// CTA-pair mode (clusters of 2, tcgen05 cta_group::2): the two CTAs compute M tiles 2j, 2j+1 of one N tile with one
// M=256 MMA stream issued by the leader; each CTA stages its own A rows and only half of the weight tile.
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
constexpr int kPanelBytes = kBlockM * 128;  // one staging panel: 128 rows x 128 bytes (64 fp16 or 32 fp32 columns)
constexpr int kMaxStages = 8;
constexpr int kThreads = 384;   // warp 0: A producer, 1: MMA, 2-5: epilogue group 0, 6-7: B producers, 8-11: epilogue group 1

struct ConvGemmParams {
  int T, H, W;
  int BW, BH, BT;
  int tiles_w, tiles_h, tiles_t, tiles_m, tiles_n;
  int kchunks1, kchunks;  // 64-wide chunks in source 1 / in both sources (per tap)
  int taps, tap_mode;  // taps = number of filter taps; tap_mode = enum mgld_taps (geometry)
  int split_k, k_per_split, slab_frames;  // split-K: work unit = (tile, split); partial fp32 tiles go to slab `split`
  int N, block_n, n_out_tile, n_out_total, n_panels;
  int stages, tmem_cols, acc_stride;
  int panel_cols;  // fp16 output columns per staging panel: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B)
  int epilogue, act, has_res, out_f32;
  int raster;    // 0: consecutive work units walk the M tiles of one N tile; 1: they walk the N tiles of one M tile
  int cta_pair;  // 1: launched as clusters of 2; one cta_group::2 MMA computes the two M tiles of a pair, each CTA stages half of B
  const float* bias;
  float alpha, beta;
  const float* gn_stats;
  const float* gn_weight;
  const float* gn_bias;
  int groups, ch_per_group;
  uint32_t off_staging, off_hstage, off_bias, off_stats;  // byte offsets from the 1024-aligned smem base
  double* stats_out;           // variant 7: [T, stats_groups, 2] x 16-byte fixed-point (sum, sumsq) accumulators of the output
  int stats_groups, stats_cpg; // groups of the consumer's GroupNorm, channels per group
  long long* dbg;  // optional per-CTA cycle counters [grid][16] (mgld_conv_gemm_set_debug_counters); null in production
};

__device__ __forceinline__ void act_inplace32(float* v, int act) {
  switch (act) {
    case MGLD_ACT_RELU:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
      break;
    case MGLD_ACT_SILU:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __fdividef(v[i], 1.f + __expf(-v[i]));
      break;
    case MGLD_ACT_LRELU02:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
      break;
    case MGLD_ACT_GELU:
#pragma unroll
      for (int i = 0; i < 32; i += 2) gelu_poly2(v[i], v[i + 1], v[i], v[i + 1]);
      break;
    case MGLD_ACT_SIGMOID:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = 1.f / (1.f + __expf(-v[i]));
      break;
    case MGLD_ACT_TANH:
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = tanhf(v[i]);
      break;
    default: break;
  }
}

// 16-byte chunk `chunk` (0..3) of the 32 consecutive fp16 columns [c0, c0+32) of row `row` in the swizzled staging
// tile.  PANEL = 64: 128-byte rows, chunk ^= row & 7 (TMA SWIZZLE_128B); PANEL = 32: 64-byte rows, chunk ^= (row >> 1) & 3
// (TMA SWIZZLE_64B) — used when BLOCK_N is an odd multiple of 32 (e.g. 160 for the 320-channel layers).
__device__ __forceinline__ uint8_t* stage_ptr16(uint8_t* staging, int row, int c0, int chunk, int panel_cols) {
  if (panel_cols == 64)
    return staging + (c0 >> 6) * kPanelBytes + row * 128 + (((((c0 & 63) >> 3) + chunk) ^ (row & 7)) << 4);
  return staging + (c0 >> 5) * (kPanelBytes / 2) + row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4);
}
__device__ __forceinline__ void load_stage32(uint8_t* staging, int row, int c0, float* r, int panel_cols) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const uint4 v = *reinterpret_cast<const uint4*>(stage_ptr16(staging, row, c0, ch, panel_cols));
    const __half2* hh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 f = __half22float2(hh[u]);
      r[ch * 8 + 2 * u] = f.x;
      r[ch * 8 + 2 * u + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void store_stage32(uint8_t* staging, int row, int c0, const float* o, int panel_cols) {
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint4 pk;
    pk.x = pack_h2(o[ch * 8 + 0], o[ch * 8 + 1]);
    pk.y = pack_h2(o[ch * 8 + 2], o[ch * 8 + 3]);
    pk.z = pack_h2(o[ch * 8 + 4], o[ch * 8 + 5]);
    pk.w = pack_h2(o[ch * 8 + 6], o[ch * 8 + 7]);
    *reinterpret_cast<uint4*>(stage_ptr16(staging, row, c0, ch, panel_cols)) = pk;
  }
}
// fp32 output: panel = 32 columns (128 bytes) -> 8 chunks of 16 bytes
__device__ __forceinline__ void store_stage32_f32(uint8_t* staging, int row, int c0, const float* o) {
  uint8_t* base = staging + (c0 >> 5) * kPanelBytes + row * 128;
#pragma unroll
  for (int ch = 0; ch < 8; ++ch)
    *reinterpret_cast<float4*>(base + ((ch ^ (row & 7)) << 4)) =
        make_float4(o[ch * 4], o[ch * 4 + 1], o[ch * 4 + 2], o[ch * 4 + 3]);
}


// Variant 7 (GroupNorm statistics of the output accumulated by the epilogue): (sum, sumsq) accumulators in shared memory,
// two buffers (tile parity) of kStatsGroups groups x 2 moments x {integer part, 2^-40 fraction} (fixsum, common.h)
constexpr int kStatsGroups = 66;   // block_n <= 256 columns / >= 4 channels per group, + 1 for a tile that starts mid-group
constexpr int kStatsAccBytes = 2 * kStatsGroups * 2 * 16;
constexpr int kStatsBytes = kStatsAccBytes + 2 * 128 * 8;   // + per epilogue group: 128 per-thread (sum, sumsq) partials
__device__ __forceinline__ void fixsum_add_words(unsigned long long* w, float v) {
  const float fl = floorf(v);
  atomicAdd(w, (unsigned long long)__float2ll_rd(v));
  atomicAdd(w + 1, (unsigned long long)__double2ll_rd((double)(v - fl) * 1099511627776.0));
}

// kPair instantiation must be launched as clusters of 2 (a kernel containing cta_group::2 instructions cannot be launched
// without a cluster: cudaErrorInvalidClusterSize), hence two instantiations rather than a runtime flag.
// kVariant: 0 = every epilogue option is a runtime flag; 1..4 = the common LINEAR epilogues with the flags folded at
// compile time (1: +bias; 2: +bias, residual; 3: raw fp32 out (split-K partials); 4: +bias, SiLU; 5: GEGLU without
// residual; 6: SPADE; 7: +bias and the GroupNorm statistics of the output for its consumer, long-K convolutions only: each
// finished staging panel is summed column-wise by the epilogue group that converted it - the long mainloop hides the
// extra work, the epilogue of a short-K GEMM would not, profiles/r01_dev_run7*).  With runtime flags the
// per-32-column step hops through six distant code islands (parameter load -> branch), paying instruction-fetch and
// constant-load latency on every hop in a single-warp latency chain.
template <bool kPair, int kVariant>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmRes, const __grid_constant__ CUtensorMap tmH,
                 const ConvGemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages];
  __shared__ __align__(8) uint64_t empty_bar[kMaxStages];
  __shared__ __align__(8) uint64_t acc_full[2], acc_empty[2], res_full, staging_free;
  __shared__ uint32_t tmem_base_slot;

  const long long dbg_k0 = clock64();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // CTA pair: the two CTAs of a cluster own M tiles 2j and 2j+1 of the same N tile.  Each stages its own A rows and
  // block_n/2 rows of the weight tile; the leader's MMA thread issues cta_group::2 MMAs (M = 256) that read both CTAs'
  // shared memory and write both CTAs' TMEM.  L2 -> SM traffic per MMA flop drops by (128+bn/2)/(128+bn).
  constexpr bool pair = kPair;
  const int cs = pair ? 2 : 1;
  uint32_t rank = 0;
  if constexpr (pair) rank = cluster_ctarank();
  const int worker = blockIdx.x / cs, nworkers = gridDim.x / cs;
  const int tiles_mw = (p.tiles_m + cs - 1) / cs;
  const int b_rows = p.block_n / cs;                      // weight rows staged by THIS CTA
  const int stage_bytes = kABytes + b_rows * kKChunk * 2;
  const int num_k = p.taps * p.kchunks;
  const int total_units = tiles_mw * p.tiles_n * p.split_k;  // work units of a worker (CTA or CTA pair)
  constexpr bool kGen = kVariant == 0;
  constexpr int kEpiFixed = kVariant == 5 ? MGLD_EPI_GEGLU : kVariant == 6 ? MGLD_EPI_SPADE : MGLD_EPI_LINEAR;
  const int epi = kGen ? p.epilogue : kEpiFixed;
  const int act = kGen ? p.act : (kVariant == 4 ? MGLD_ACT_SILU : MGLD_ACT_NONE);
  const bool has_res = kGen ? (p.has_res != 0) : (kVariant == 2 || (kVariant == 6 && p.has_res != 0));
  const bool out_f32 = kGen ? (p.out_f32 != 0) : (kVariant == 3);
  const bool unit_alpha = kGen ? (p.alpha == 1.f) : true;   // variants 1, 3, 4, 7 require alpha == 1; 2 uses the residual form
  constexpr bool kStats = kVariant == 7;
  const bool pair_spade = epi == MGLD_EPI_SPADE;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmOut);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);   // one arrival (A producer; the pair leader's) carrying the stage's whole tx count
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&acc_full[b]), 1);
      mbar_init(smem_u32(&acc_empty[b]), 256 * cs);
    }
    mbar_init(smem_u32(&res_full), 1);
    mbar_init(smem_u32(&staging_free), 2);   // both epilogue group leaders
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (pair) tmem_alloc_pair(smem_u32(&tmem_base_slot), p.tmem_cols);
    else tmem_alloc(smem_u32(&tmem_base_slot), p.tmem_cols);
  }
  tc_fence_before();
  if constexpr (pair) cluster_sync_all();   // barrier inits visible to the peer before any remote arrive / TMA completion
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // Everything above (barrier init, TMEM allocation, descriptor prefetch) is independent of earlier kernels' output and
  // overlaps the predecessor's tail; from here on global memory written by earlier kernels is read.  The trigger for
  // the NEXT kernel comes only now, after this CTA owns its TMEM columns: a successor CTA that became co-resident and
  // allocated first would hold the columns while waiting for this grid to finish.
  pdl_launch_dependents();
  pdl_wait();
  if (p.dbg && threadIdx.x == 0) { p.dbg[blockIdx.x * 16 + 12] = clock64() - dbg_k0; }

  // tile -> coordinates; consecutive tiles share the weight tile (n) and walk the pixel boxes (L2-friendly)
  // returns false for the padding M tile of an odd tile count (pair mode): its box lies beyond T, loads are zero-filled
  auto tile_coords = [&](int unit, int& x0, int& y0, int& t0, int& nt) -> bool {
    const int tile = unit / p.split_k;   // the splits of one tile run on neighbouring workers
    int tmw;
    if (p.raster) { tmw = tile / p.tiles_n; nt = tile - tmw * p.tiles_n; }
    else { nt = tile / tiles_mw; tmw = tile - nt * tiles_mw; }
    const int tm = tmw * cs + static_cast<int>(rank);
    x0 = (tm % p.tiles_w) * p.BW;
    y0 = ((tm / p.tiles_w) % p.tiles_h) * p.BH;
    t0 = (tm / (p.tiles_w * p.tiles_h)) * p.BT;
    return tm < p.tiles_m;
  };

  if (warp == 0) {
    // ============================ TMA producer ============================
    // The three single-thread loops (this one, the MMA issuer, the B producers) are latency chains of scalar
    // instructions: ~110 dependent instructions per K chunk (runtime divisions, S2UR address rebuilds, parameter reloads)
    // cost ~800 cycles and WERE the mainloop bound.  So: all loop state is carried incrementally in registers.
    if (elect_one()) {
      int lt = 0;
      int s = 0;
      uint32_t ph = 0;
      uint32_t sa = smem_base;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t full0_lead = full0;
      if constexpr (pair) full0_lead = mapa_shared(full0, 0);
      const int stages = p.stages, kchunks = p.kchunks, kchunks1 = p.kchunks1, tap_mode = p.tap_mode;
      const uint32_t stage_tx = cs * stage_bytes;
      long long dbg_wait = 0;
      const long long dbg_t0 = p.dbg ? clock64() : 0;
      for (int tile = worker; tile < total_units; tile += nworkers, ++lt) {
        int x0, y0, t0, nt;
        tile_coords(tile, x0, y0, t0, nt);
        const int kb = (tile % p.split_k) * p.k_per_split, ke = min(kb + p.k_per_split, num_k);
        int tap = kb / kchunks, kc = kb - tap * kchunks;
        int cx = x0, cy = y0, ct = t0;   // box origin of the current tap
        auto set_tap = [&](int tp) {
          cx = x0; cy = y0; ct = t0;
          if (tap_mode == MGLD_TAPS_3X3) { cx += tp % 3 - 1; cy += tp / 3 - 1; }
          else if (tap_mode == MGLD_TAPS_T3) { ct += tp - 1; }
          else if (tap_mode == MGLD_TAPS_1X5) { cx += tp - 2; }
          else if (tap_mode == MGLD_TAPS_5X1) { cy += tp - 2; }
        };
        set_tap(tap);
        for (int k = kb; k < ke; ++k) {
          if (p.dbg) {
            const long long tw = clock64();
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            dbg_wait += clock64() - tw;
          } else {
            mbar_wait(empty0 + 8 * s, ph ^ 1);
          }
          // the stage's whole transaction count (A + B, of both CTAs of a pair) rides on this one arrival; the B
          // producers only issue.  Their bytes may land first (tx count transiently negative) - the phase cannot
          // complete before the pending arrival, and nobody issues into a stage before its empty barrier fired.
          if (rank == 0) mbar_expect_tx(full0 + 8 * s, stage_tx);
          const bool src1 = kc < kchunks1;
          const CUtensorMap* tm = src1 ? &tmA : &tmA2;
          const int c0 = (src1 ? kc : kc - kchunks1) * kKChunk;
          if constexpr (pair) tma_load_4d_pair(sa, tm, full0_lead + 8 * s, c0, cx, cy, ct);
          else tma_load_4d(sa, tm, full0 + 8 * s, c0, cx, cy, ct);
          if (++kc == kchunks) { kc = 0; set_tap(++tap); }
          sa += stage_bytes;
          if (++s == stages) { s = 0; ph ^= 1; sa = smem_base; }
        }
        if (has_res || pair_spade) {
          // residual (and SPADE's h) tile of THIS output tile, into the staging buffers the epilogue will overwrite
          mbar_wait(smem_u32(&staging_free), (lt & 1) ^ 1);
          const uint32_t rb = smem_u32(&res_full);
          const int c0 = nt * p.n_out_tile;
          const int np = has_res ? p.n_panels : 0;
          mbar_expect_tx(rb, np * (p.panel_cols == 32 ? kPanelBytes / 2 : kPanelBytes) + (pair_spade ? p.n_panels * kPanelBytes : 0));
          const int pbytes = p.panel_cols == 32 ? kPanelBytes / 2 : kPanelBytes;
          for (int q = 0; q < np; ++q)
            tma_load_4d(smem_base + p.off_staging + q * pbytes, &tmRes, rb, c0 + q * p.panel_cols, x0, y0, t0);
          if (pair_spade)
            for (int q = 0; q < p.n_panels; ++q)
              tma_load_4d(smem_base + p.off_hstage + q * kPanelBytes, &tmH, rb, c0 + q * 64, x0, y0, t0);
        }
      }
      if (p.dbg) { p.dbg[blockIdx.x * 16 + 0] = dbg_wait; p.dbg[blockIdx.x * 16 + 1] = clock64() - dbg_t0; }
    }
  } else if (warp == 1) {
    // ============================ MMA issuer ============================
    if (rank == 0 && elect_one()) {
      const uint32_t idesc = umma_idesc_f16(kBlockM * cs, p.block_n, 0, 0);
      int lt = 0;
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const int stages = p.stages;
      // descriptors differ only in the 14-bit (address >> 4) field: carry that field incrementally
      const uint64_t desc0 = umma_smem_desc(smem_base, 0, 1024, kSwz128);
      const uint32_t desc_step = stage_bytes >> 4;
      uint64_t adesc = desc0;
      long long dbg_wf = 0, dbg_wa = 0;
      const long long dbg_t0 = p.dbg ? clock64() : 0;
      for (int tile = worker; tile < total_units; tile += nworkers, ++lt) {
        const int buf = lt & 1;
        const long long twa = p.dbg ? clock64() : 0;
        mbar_wait(smem_u32(&acc_empty[buf]), ((lt >> 1) & 1) ^ 1);
        if (p.dbg) dbg_wa += clock64() - twa;
        tc_fence_after();
        const uint32_t dcol = tmem_base + buf * p.acc_stride;
        const int kb = (tile % p.split_k) * p.k_per_split, ke = min(kb + p.k_per_split, num_k);
        for (int k = kb; k < ke; ++k) {
          if (p.dbg) {
            const long long tw = clock64();
            mbar_wait(full0 + 8 * s, ph);
            dbg_wf += clock64() - tw;
          } else {
            mbar_wait(full0 + 8 * s, ph);
          }
          tc_fence_after();
          const uint64_t bdesc = adesc + (kABytes >> 4);
          if constexpr (pair) {
#pragma unroll
            for (int kk = 0; kk < kKChunk / 16; ++kk)
              umma_ss_pair(dcol, adesc + 2 * kk, bdesc + 2 * kk, idesc, ((k - kb) | kk) != 0);
            umma_commit_pair(empty0 + 8 * s, 3);   // frees the stage in both CTAs
          } else {
#pragma unroll
            for (int kk = 0; kk < kKChunk / 16; ++kk)
              umma_ss(dcol, adesc + 2 * kk, bdesc + 2 * kk, idesc, ((k - kb) | kk) != 0);  // +32 B along K per step
            umma_commit(empty0 + 8 * s);
          }
          adesc += desc_step;
          if (++s == stages) { s = 0; ph ^= 1; adesc = desc0; }
        }
        if constexpr (pair) umma_commit_pair(smem_u32(&acc_full[buf]), 3);
        else umma_commit(smem_u32(&acc_full[buf]));
      }
      if (p.dbg) { p.dbg[blockIdx.x * 16 + 2] = dbg_wf; p.dbg[blockIdx.x * 16 + 3] = dbg_wa; p.dbg[blockIdx.x * 16 + 4] = clock64() - dbg_t0; }
    }
  } else if (warp == 6 || warp == 7) {
    // ============================ B (weight) producers ============================
    // The tensor map's box is block_n/2 rows.  Single CTA: warps 6 and 7 each load one half (a TMA-issuing thread manages
    // one op per ~270 cycles, tools/microbench/tma_fill4.cu).  CTA pair: warp 6 loads this CTA's half, warp 7 idles.
    if (!(pair && warp == 7) && elect_one()) {
      const int half = pair ? static_cast<int>(rank) : warp - 6;
      const int hrows = p.block_n >> 1;
      const int hbytes = hrows * kKChunk * 2;
      const int slot = pair ? 0 : half;   // position of the half inside this CTA's B stage
      int s = 0;
      uint32_t ph = 0;
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t full0_lead = full0;
      if constexpr (pair) full0_lead = mapa_shared(full0, 0);
      const int stages = p.stages;
      const uint32_t dst0 = smem_base + kABytes + slot * hbytes;
      uint32_t dst = dst0;
      long long dbg_wb = 0;
      const long long dbg_t0b = p.dbg ? clock64() : 0;
      for (int tile = worker; tile < total_units; tile += nworkers) {
        int x0, y0, t0, nt;
        tile_coords(tile, x0, y0, t0, nt);
        const int n0 = nt * p.block_n + half * hrows;
        const int kb = (tile % p.split_k) * p.k_per_split, ke = min(kb + p.k_per_split, num_k);
        for (int k = kb; k < ke; ++k) {
          if (p.dbg) {
            const long long tw = clock64();
            mbar_wait(empty0 + 8 * s, ph ^ 1);
            dbg_wb += clock64() - tw;
          } else {
            mbar_wait(empty0 + 8 * s, ph ^ 1);
          }
          if constexpr (pair) tma_load_2d_pair(dst, &tmB, full0_lead + 8 * s, k * kKChunk, n0);
          else tma_load_2d(dst, &tmB, full0 + 8 * s, k * kKChunk, n0);
          dst += stage_bytes;
          if (++s == stages) { s = 0; ph ^= 1; dst = dst0; }
        }
      }
      if (p.dbg && warp == 6) { p.dbg[blockIdx.x * 16 + 8] = dbg_wb; p.dbg[blockIdx.x * 16 + 9] = clock64() - dbg_t0b; }
    }
  } else {
    // ============================ epilogue (warps 2..5 = group 0, warps 8..11 = group 1) ============================
    // One epilogue warp per SM sub-partition is a pure latency chain (tcgen05.ld -> math -> st.shared, IPC ~0.4), so two
    // warpgroups share a tile: group g converts every second staging panel (64 output columns; 32 for fp32 / odd tiles).
    // As soon as a panel is complete in shared memory its group leader issues the TMA store, which then drains while the
    // next panel is being converted; only the last store's drain is exposed at the end of the tile.
    const int q = warp & 3;         // TMEM lane quadrant of this warp
    const int row = q * 32 + lane;  // accumulator row == tile row
    const int grp = warp >= 8 ? 1 : 0;
    const int e = grp * 128 + row;  // 0..255
    const bool leader = row == 0;   // one per group: issues that group's TMA stores
    uint8_t* staging = smem_gen + p.off_staging;
    uint8_t* hstage = smem_gen + p.off_hstage;
    float* bias_s0 = reinterpret_cast<float*>(smem_gen + p.off_bias);   // two buffers of 256 floats (tile parity)
    constexpr int ngroups = 2;
    unsigned long long* sacc = reinterpret_cast<unsigned long long*>(smem_gen + p.off_stats);   // variant 7 only
    int prev_t = 0, prev_g0 = 0;
    if constexpr (kStats) {
      for (int i = e; i < kStatsAccBytes / 8; i += 256) sacc[i] = 0ull;   // ordered before any use by the first tile's opening barrier
    }
    // add the finished sums of buffer `b` (frame t, groups g0 ...) to the global accumulators and clear the buffer
    auto flush_stats = [&](int b, int t, int g0) {
      for (int i = e; i < kStatsGroups * 2; i += 256) {   // i = local group * 2 + moment
        unsigned long long* w = sacc + (b * kStatsGroups * 2 + i) * 2;
        const unsigned long long hi = w[0], lo = w[1];
        if (hi | lo) {
          unsigned long long* dst = reinterpret_cast<unsigned long long*>(p.stats_out) +
                                    ((static_cast<long long>(t) * p.stats_groups + g0 + (i >> 1)) * 2 + (i & 1)) * 2;
          atomicAdd(dst, hi);
          atomicAdd(dst + 1, lo);
          w[0] = 0ull; w[1] = 0ull;
        }
      }
    };
    const int cols_per_panel = out_f32 ? 32 : p.panel_cols;
    const int pbytes = (!out_f32 && p.panel_cols == 32) ? kPanelBytes / 2 : kPanelBytes;
    int lt = 0;
    long long dbg_we = 0, dbg_ec = 0, dbg_es = 0;
    const long long dbg_t0e = p.dbg ? clock64() : 0;
    // Without a residual tile nobody else touches the staging panels, so the drain of a tile's TMA stores is only awaited
    // at the start of the NEXT tile's conversion (it has long finished by then) and the closing barrier disappears.
    const bool defer_drain = !(has_res || pair_spade);
    auto bias_of = [&](int unit) {   // this thread's bias element of work unit `unit` (block_n <= 256 = one per thread)
      const int tl = unit / p.split_k;
      const int n = (p.raster ? tl % p.tiles_n : tl / tiles_mw) * p.block_n + e;
      return (p.bias && e < p.block_n && n < p.N) ? __ldg(p.bias + n) : 0.f;
    };
    if (worker < total_units) bias_s0[e] = bias_of(worker);
    for (int tile = worker; tile < total_units; tile += nworkers, ++lt) {
      int x0, y0, t0, nt;
      const bool tile_valid = tile_coords(tile, x0, y0, t0, nt);
      const int buf = lt & 1;
      const int out_c0 = nt * p.n_out_tile;
      const int t_store = t0 + (tile % p.split_k) * p.slab_frames;
      const float* bias_s = bias_s0 + (lt & 1) * 256;
      // the next tile's bias element travels in a register during this tile's conversion (global-load latency hidden)
      const float bias_next = (tile + nworkers < total_units) ? bias_of(tile + nworkers) : 0.f;
      const long long twe = (p.dbg && e == 0) ? clock64() : 0;
      mbar_wait(smem_u32(&acc_full[buf]), (lt >> 1) & 1);
      if (p.dbg && e == 0) dbg_we += clock64() - twe;
      tc_fence_after();
      if (has_res || pair_spade) mbar_wait(smem_u32(&res_full), lt & 1);
      if (defer_drain && leader) {
        const long long td = (p.dbg && e == 0) ? clock64() : 0;
        tma_store_wait_read();   // the previous tile's stores of this group have read their panels
        if (p.dbg && e == 0) dbg_es += clock64() - td;
      }
      named_bar_sync(1, 256);
      if constexpr (kStats) {
        if (lt > 0) flush_stats((lt - 1) & 1, prev_t, prev_g0);   // every thread is past the previous tile's panel sums
      }
      const long long tc0 = (p.dbg && e == 0) ? clock64() : 0;
      const uint32_t trow = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * p.acc_stride;
      // panel `pn` of this group is complete in smem -> its leader stores it
      auto store_panel = [&](int pn) {
        fence_proxy_async_smem();
        named_bar_sync(2 + grp, 128);
        if (leader && tile_valid && out_c0 + pn * cols_per_panel < p.n_out_total) {
          tma_store_4d(&tmOut, smem_base + p.off_staging + pn * pbytes, out_c0 + pn * cols_per_panel, x0, y0, t_store);
          tma_store_commit();
        }
      };

      // variant 7: column sums of panel `pn` (complete in shared memory: store_panel() synchronised its group).  Thread ->
      // one column and cols_per_panel of the 128 rows; a warp reads 32 consecutive columns of one row (conflict-free under
      // either swizzle).  The values are the fp16-rounded outputs, i.e. what a gn_stats pass over the stored tensor reads.
      auto panel_stats = [&](int pn) {
        const int cpp = cols_per_panel;
        const int cc = row & (cpp - 1), rp = row / cpp;
        const int col = pn * cpp + cc;                       // column within the tile
        const int c32 = col & ~31, chunk = (col & 31) >> 3, sub = (col & 7) * 2;
        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 4
        for (int r = rp * cpp; r < (rp + 1) * cpp; r += 2) {
          const float f0 = __half2float(*reinterpret_cast<const __half*>(stage_ptr16(staging, r, c32, chunk, p.panel_cols) + sub));
          const float f1 = __half2float(*reinterpret_cast<const __half*>(stage_ptr16(staging, r + 1, c32, chunk, p.panel_cols) + sub));
          s0 += f0; q0 = fmaf(f0, f0, q0);
          s1 += f1; q1 = fmaf(f1, f1, q1);
        }
        // per-thread partials -> table -> one thread per (group, moment) sums its columns in a fixed order.  (Adding the
        // partials straight into the shared accumulators was 40 threads per address: the epilogue took 1.5x the mainloop.)
        float2* tab = reinterpret_cast<float2*>(sacc + kStatsAccBytes / 8) + grp * 128;
        tab[row] = make_float2(s0 + s1, q0 + q1);            // index = rp * cpp + cc
        named_bar_sync(2 + grp, 128);   // the next write of `tab` comes after the next panel's store_panel() barrier
        const int cpg = p.stats_cpg;
        const int ch0 = out_c0 + pn * cpp;                   // first channel of the panel
        const int g_first = ch0 / cpg;
        const int gl = row >> 1, mo = row & 1;
        const int lo = max((g_first + gl) * cpg, ch0), hi = min(min((g_first + gl + 1) * cpg, ch0 + cpp), p.n_out_total);
        if (tile_valid && lo < hi) {
          float acc = 0.f;
          const float* tf = reinterpret_cast<const float*>(tab) + mo;
          for (int c = lo - ch0; c < hi - ch0; ++c)
            for (int k = 0; k < 128 / cpp; ++k) acc += tf[(k * cpp + c) * 2];
          fixsum_add_words(sacc + (((lt & 1) * kStatsGroups + (g_first + gl - out_c0 / cpg)) * 2 + mo) * 2, acc);
        }
      };

      if (epi == MGLD_EPI_LINEAR) {
        // Flat list of this group's 32-column steps, software-pipelined over two register buffers: the tcgen05.ld of
        // step i+1 is in flight while step i is converted (TMEM loads are slow while the MMA is accumulating).
        const int steps = cols_per_panel / 32;   // 32-column steps per panel (1 or 2)
        const int my_panels = (grp < ngroups && grp < p.n_panels) ? (p.n_panels - grp + ngroups - 1) / ngroups : 0;
        const int nsteps = my_panels * steps;
        auto col_of = [&](int i) {
          const int pi = steps == 2 ? (i >> 1) : i, st = steps == 2 ? (i & 1) : 0;
          return (grp + pi * ngroups) * cols_per_panel + st * 32;
        };
        auto process = [&](float* v, int i) {
          const int c0 = col_of(i);
          // packed fp32x2 adds / FMAs: the step is a single-warp instruction chain (~290 SASS instructions per 32 columns
          // before this), so every instruction saved shortens the epilogue of the short-K GEMMs directly
#pragma unroll
          for (int u = 0; u < 32; u += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + c0 + u);
            fadd2(v[u], v[u + 1], v[u], v[u + 1], bb.x, bb.y);
            fadd2(v[u + 2], v[u + 3], v[u + 2], v[u + 3], bb.z, bb.w);
          }
          act_inplace32(v, act);
          if (has_res) {
            float r[32];
            load_stage32(staging, row, c0, r, p.panel_cols);
            if (p.alpha == 1.f && p.beta == 1.f) {       // the plain residual add of every out-projection
#pragma unroll
              for (int u = 0; u < 32; u += 2) fadd2(v[u], v[u + 1], v[u], v[u + 1], r[u], r[u + 1]);
            } else {
              const float al = p.alpha, be = p.beta;
#pragma unroll
              for (int u = 0; u < 32; u += 2) {
                float b0, b1;
                fmul2(b0, b1, r[u], r[u + 1], be, be);
                ffma2(v[u], v[u + 1], v[u], v[u + 1], al, al, b0, b1);
              }
            }
          } else if (!unit_alpha) {
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] *= p.alpha;
          }
          if (out_f32) store_stage32_f32(staging, row, c0, v);
          else store_stage32(staging, row, c0, v, p.panel_cols);
          if (steps == 1 || (i & 1)) {   // panel complete
            if (i == nsteps - 1) {       // ... and it was the last: this thread's accumulator rows are drained
              tc_fence_before();
              if (!pair || rank == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
              else mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));   // the pair's MMA issuer lives in the leader
            }
            store_panel(c0 / cols_per_panel);
            if constexpr (kStats) panel_stats(c0 / cols_per_panel);
          }
        };
        float va[32], vb[32];
        if (nsteps > 0) tmem_ld_x32(trow + col_of(0), reinterpret_cast<uint32_t*>(va));
        for (int i = 0; i < nsteps; i += 2) {
          tmem_ld_wait();
          if (i + 1 < nsteps) tmem_ld_x32(trow + col_of(i + 1), reinterpret_cast<uint32_t*>(vb));
          process(va, i);
          if (i + 1 < nsteps) {
            tmem_ld_wait();
            if (i + 2 < nsteps) tmem_ld_x32(trow + col_of(i + 2), reinterpret_cast<uint32_t*>(va));
            process(vb, i + 1);
          }
        }
        if (nsteps == 0) {   // a group without panels still releases the accumulator
          tc_fence_before();
          if (!pair || rank == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
          else mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
        }
      } else {
        // pair epilogues: 64-column output panels, each from 128 accumulator columns (value | gate, gamma | beta in blocks
        // of 64).  block_n = 128: one panel, group g converts its 32-column half; block_n = 256: group g converts panel g.
        const int npan = p.n_panels;
        const int pn = npan == 2 ? grp : 0;
        if (grp < ngroups) {
          for (int c0 = (npan == 2 ? 0 : grp * 32); c0 < 64; c0 += (npan == 2 ? 32 : ngroups * 32)) {
            const int ac = pn * 128 + c0;   // accumulator column of the value / gamma slice; gate / beta at +64
            const int oc = pn * 64 + c0;    // output column within the tile
            if (epi == MGLD_EPI_GEGLU) {
              float v[32], g[32];
              tmem_ld_x32(trow + ac, reinterpret_cast<uint32_t*>(v));
              tmem_ld_x32(trow + ac + 64, reinterpret_cast<uint32_t*>(g));
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; i += 4) {   // 16-byte shared loads of the two bias slices, packed adds
                const float4 bv = *reinterpret_cast<const float4*>(bias_s + ac + i);
                const float4 bg = *reinterpret_cast<const float4*>(bias_s + ac + 64 + i);
                fadd2(v[i], v[i + 1], v[i], v[i + 1], bv.x, bv.y);
                fadd2(v[i + 2], v[i + 3], v[i + 2], v[i + 3], bv.z, bv.w);
                fadd2(g[i], g[i + 1], g[i], g[i + 1], bg.x, bg.y);
                fadd2(g[i + 2], g[i + 3], g[i + 2], g[i + 3], bg.z, bg.w);
              }
#pragma unroll
              for (int i = 0; i < 32; i += 2) {   // value * gelu(gate)
                float g0, g1;
                gelu_poly2(g0, g1, g[i], g[i + 1]);
                fmul2(v[i], v[i + 1], v[i], v[i + 1], g0, g1);
              }
              if (has_res) {
                float r[32];
                load_stage32(staging, row, oc, r, 64);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaf(p.alpha, v[i], p.beta * r[i]);
              }
              store_stage32(staging, row, oc, v, 64);
            } else {  // SPADE: out = beta*res + GNaffine(h) * (1 + gamma) + beta_s
              const int tt = min(t0 + row / (p.BW * p.BH), p.T - 1);
              const int cbase = nt * p.n_out_tile + oc;
              float gm[32], bt[32], hv[32];
              tmem_ld_x32(trow + ac, reinterpret_cast<uint32_t*>(gm));
              tmem_ld_x32(trow + ac + 64, reinterpret_cast<uint32_t*>(bt));
              tmem_ld_wait();
              load_stage32(hstage, row, oc, hv, 64);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int c = cbase + i;
                const int g = c / p.ch_per_group;
                const float2 st = __ldg(reinterpret_cast<const float2*>(p.gn_stats) + tt * p.groups + g);
                const float xn = fmaf((hv[i] - st.x) * st.y, __ldg(p.gn_weight + c), __ldg(p.gn_bias + c));
                gm[i] = fmaf(xn, 1.f + gm[i] + bias_s[ac + i], bt[i] + bias_s[ac + 64 + i]);
              }
              if (has_res) {
                float r[32];
                load_stage32(staging, row, oc, r, 64);
#pragma unroll
                for (int i = 0; i < 32; ++i) gm[i] = fmaf(p.beta, r[i], gm[i]);
              }
              store_stage32(staging, row, oc, gm, 64);
            }
          }
        }
        tc_fence_before();
        if (!pair || rank == 0) mbar_arrive(smem_u32(&acc_empty[buf]));
        else mbar_arrive_cluster(mapa_shared(smem_u32(&acc_empty[buf]), 0));
        if (npan == 2) {
          store_panel(grp);       // each group stores the panel it converted
        } else {
          fence_proxy_async_smem();
          named_bar_sync(1, 256);   // both halves of the panel are in smem
          if (e == 0 && tile_valid) {
            tma_store_4d(&tmOut, smem_base + p.off_staging, out_c0, x0, y0, t_store);
            tma_store_commit();
          }
        }
      }
      const long long tc1 = (p.dbg && e == 0) ? clock64() : 0;
      if constexpr (kStats) { prev_t = t0; prev_g0 = out_c0 / p.stats_cpg; }
      bias_s0[((lt + 1) & 1) * 256 + e] = bias_next;   // read after the next tile's opening barrier
      if (p.dbg && e == 0) dbg_ec += tc1 - tc0;
      if (!defer_drain) {
        if (leader) {
          tma_store_wait_read();  // this group's stores have read their panels: the residual of the next tile may land
          mbar_arrive(smem_u32(&staging_free));
          if (p.dbg && e == 0) dbg_es += clock64() - tc1;
        }
        named_bar_sync(1, 256);
      }
    }
    if constexpr (kStats) {
      named_bar_sync(1, 256);
      if (lt > 0) flush_stats((lt - 1) & 1, prev_t, prev_g0);
    }
    if (leader) tma_store_wait_all();
    if (p.dbg && e == 0) { p.dbg[blockIdx.x * 16 + 5] = dbg_we; p.dbg[blockIdx.x * 16 + 6] = clock64() - dbg_t0e; p.dbg[blockIdx.x * 16 + 7] = lt; p.dbg[blockIdx.x * 16 + 10] = dbg_ec; p.dbg[blockIdx.x * 16 + 11] = dbg_es; }
  }

  tc_fence_before();
  if constexpr (pair) cluster_sync_all();   // the peer may still signal this CTA's barriers / read its smem through the MMA
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (pair) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (p.dbg && threadIdx.x == 0) { p.dbg[blockIdx.x * 16 + 13] = clock64() - dbg_k0; }
}


// Split-K finalize: sum the S partial fp32 tiles of every output element in a fixed order (deterministic) and apply the
// epilogue the single-pass kernel would have applied.  ws: [S][slab_frames][H][W][ldws] fp32 raw accumulators.
struct SplitFinalizeParams {
  const float* ws;
  long long slab_stride;  // floats between slabs
  int S, T, H, W, ldws;
  int N, n_out_total, epilogue, act;
  const float* bias;
  float alpha, beta;
  const __half* res; int ldres;
  const __half* h; int ldh;
  const float* gn_stats; const float* gn_weight; const float* gn_bias; int groups, ch_per_group;
  __half* out; int ldout, out_col0;
};
__device__ __forceinline__ float act1(float v, int act) {
  switch (act) {
    case MGLD_ACT_RELU: return fmaxf(v, 0.f);
    case MGLD_ACT_SILU: return v / (1.f + __expf(-v));
    case MGLD_ACT_LRELU02: return v > 0.f ? v : 0.2f * v;
    case MGLD_ACT_GELU: return 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    case MGLD_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    case MGLD_ACT_TANH: return tanhf(v);
    default: return v;
  }
}
__global__ void splitk_finalize_kernel(const SplitFinalizeParams p) {
  pdl_launch_dependents();
  pdl_wait();

  const int vpr = (p.n_out_total + 7) >> 3;
  const long long M = static_cast<long long>(p.T) * p.H * p.W;
  const long long total = M * vpr;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / vpr;
    const int c = (int)(i - m * vpr) * 8;
    const bool pair = p.epilogue != MGLD_EPI_LINEAR;
    const int ca = pair ? (c >> 6) * 128 + (c & 63) : c;   // raw accumulator column of the first half of a pair
    float a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { a[u] = 0.f; b[u] = 0.f; }
    for (int s = 0; s < p.S; ++s) {
      const float* row = p.ws + s * p.slab_stride + m * p.ldws;
      const float4 x0 = *reinterpret_cast<const float4*>(row + ca), x1 = *reinterpret_cast<const float4*>(row + ca + 4);
      a[0] += x0.x; a[1] += x0.y; a[2] += x0.z; a[3] += x0.w; a[4] += x1.x; a[5] += x1.y; a[6] += x1.z; a[7] += x1.w;
      if (pair) {
        const float4 y0 = *reinterpret_cast<const float4*>(row + ca + 64), y1 = *reinterpret_cast<const float4*>(row + ca + 68);
        b[0] += y0.x; b[1] += y0.y; b[2] += y0.z; b[3] += y0.w; b[4] += y1.x; b[5] += y1.y; b[6] += y1.z; b[7] += y1.w;
      }
    }
    float r[8];
    if (p.res) {
      const uint4 rr = __ldg(reinterpret_cast<const uint4*>(p.res + m * p.ldres + c));
      const __half2* hh = reinterpret_cast<const __half2*>(&rr);
#pragma unroll
      for (int u = 0; u < 4; ++u) { const float2 f = __half22float2(hh[u]); r[2 * u] = f.x; r[2 * u + 1] = f.y; }
    }
    float o[8];
    if (p.epilogue == MGLD_EPI_LINEAR) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float v = a[u] + ((p.bias && c + u < p.N) ? __ldg(p.bias + c + u) : 0.f);
        v = p.alpha * act1(v, p.act);
        o[u] = p.res ? fmaf(p.beta, r[u], v) : v;
      }
    } else if (p.epilogue == MGLD_EPI_GEGLU) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float v = a[u] + (p.bias ? __ldg(p.bias + ca + u) : 0.f), g = b[u] + (p.bias ? __ldg(p.bias + ca + 64 + u) : 0.f);
        const float y = v * act1(g, MGLD_ACT_GELU);
        o[u] = p.res ? fmaf(p.alpha, y, p.beta * r[u]) : y;
      }
    } else {
      const int t = (int)(m / (static_cast<long long>(p.H) * p.W));
      const uint4 hr = __ldg(reinterpret_cast<const uint4*>(p.h + m * p.ldh + c));
      const __half2* hh = reinterpret_cast<const __half2*>(&hr);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int cc = c + u, g = cc / p.ch_per_group;
        const float2 st = __ldg(reinterpret_cast<const float2*>(p.gn_stats) + t * p.groups + g);
        const float2 f = __half22float2(hh[u >> 1]);
        const float hv = (u & 1) ? f.y : f.x;
        const float xn = fmaf((hv - st.x) * st.y, __ldg(p.gn_weight + cc), __ldg(p.gn_bias + cc));
        const float gm = a[u] + (p.bias ? __ldg(p.bias + ca + u) : 0.f), bt = b[u] + (p.bias ? __ldg(p.bias + ca + 64 + u) : 0.f);
        const float y = fmaf(xn, 1.f + gm, bt);
        o[u] = p.res ? fmaf(p.beta, r[u], y) : y;
      }
    }
    uint4 pk;
    pk.x = pack_h2(o[0], o[1]); pk.y = pack_h2(o[2], o[3]); pk.z = pack_h2(o[4], o[5]); pk.w = pack_h2(o[6], o[7]);
    *reinterpret_cast<uint4*>(p.out + m * p.ldout + p.out_col0 + c) = pk;
  }
}

// split-K plan.  The mainloop of one CTA is bound by its SM's L2->SMEM fill rate (~50 B/clk), so a layer's time is
// ~ waves x (K chunks per unit); splitting K helps whenever it shortens that product by more than the price of the second
// (finalize) pass, which is about 10 chunk-times.  Short-K GEMMs (1x1 / linear layers) never split.
static int pick_split_k(int tiles, int num_k, int sms) {
  if (num_k < 48) return 1;
  const double fin = 10.0;
  double best_cost = (double)((tiles + sms - 1) / sms) * num_k;
  int best = 1;
  for (int S = 2; S <= 8; ++S) {
    const int kper = (num_k + S - 1) / S;
    if (kper < 12) break;
    const int units = tiles * S;
    const double cost = (double)((units + sms - 1) / sms) * kper + fin;
    if (cost < best_cost * 0.85) { best_cost = cost; best = S; }
  }
  return best;
}

static int pick_block_n(int N, int tiles_m, int sms) {
  // widest tile that keeps the machine busy: cost ~ waves * (mainloop ~ bn + per-tile overhead)
  const int cands[] = {256, 192, 160, 128, 96, 64, 32};
  int best = 32;
  double best_cost = 1e30;
  for (int bn : cands) {
    const int tn = ceil_div(N, bn);
    const long long tiles = 1LL * tiles_m * tn;
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = (double)waves * (bn + 32.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// choose the pixel box (BW,BH,BT), BW*BH*BT = 128, minimising padded work
static void pick_box(int T, int H, int W, int* BW, int* BH, int* BT) {
  long long best = -1;
  for (int bw = 128; bw >= 1; bw >>= 1) {
    for (int bh = 128 / bw; bh >= 1; bh >>= 1) {
      const int bt = 128 / (bw * bh);
      const long long cost = 1LL * ceil_div(W, bw) * ceil_div(H, bh) * ceil_div(T, bt);
      if (best < 0 || cost < best) { best = cost; *BW = bw; *BH = bh; *BT = bt; }  // ties: widest box wins
    }
  }
}

static bool g_attr_set = false;
static int g_max_pairs = -1;   // co-resident 2-CTA clusters of conv_gemm_kernel (GPCs with an odd SM count strand one SM)

// CTA-pair mode: MGLD_CONV_PAIR=0 disables, =1 forces (where legal); default: on when both halves of the machine get work
static long long* g_dbg = nullptr;
static int pair_mode_env() {
  const char* e = getenv("MGLD_CONV_PAIR");
  return e ? atoi(e) : -1;
}

}  // namespace mgld

using namespace mgld;

// `fused_stats` (optional): in = the caller wants d->stats_out accumulated by the epilogue if this launch qualifies;
// out = whether it was (otherwise the caller runs the streaming gn_stats pass over the output)
static int conv_gemm_single(const mgld_conv_gemm_desc* d, void* stream_, int split_k, int slab_frames,
                            bool* fused_stats = nullptr) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (!initialised()) { set_error("mgld_init() has not been called"); return MGLD_ERR_NOT_INIT; }
  MGLD_CHECK_ARG(d && d->a && d->w && d->out, "conv_gemm: null pointer");
  MGLD_CHECK_ARG(d->T > 0 && d->H > 0 && d->W > 0, "conv_gemm: bad T/H/W %d/%d/%d", d->T, d->H, d->W);
  // K tail: with a single source and one tap the last 64-chunk may be partial (TMA zero-fills A and W alike)
  MGLD_CHECK_ARG(d->C1 > 0 && (d->C1 % 64 == 0 || (d->C1 % 8 == 0 && d->taps == 1 && d->C2 == 0)),
                 "conv_gemm: C1=%d must be a multiple of 64 (or of 8 for a single-source 1-tap GEMM)", d->C1);
  MGLD_CHECK_ARG(d->C2 >= 0 && d->C2 % 64 == 0 && ((d->C2 > 0) == (d->a2 != nullptr)),
                 "conv_gemm: C2=%d must be a multiple of 64 and match a2", d->C2);
  MGLD_CHECK_ARG(d->taps == MGLD_TAPS_1 || d->taps == MGLD_TAPS_T3 || d->taps == MGLD_TAPS_3X3 ||
                     d->taps == MGLD_TAPS_1X5 || d->taps == MGLD_TAPS_5X1, "conv_gemm: taps=%d", d->taps);
  const int ntaps = d->taps == MGLD_TAPS_5X1 ? 5 : d->taps;
  MGLD_CHECK_ARG(d->N > 0, "conv_gemm: N=%d", d->N);
  MGLD_CHECK_ARG(d->epilogue >= 0 && d->epilogue <= 2, "conv_gemm: epilogue=%d", d->epilogue);
  const bool pair = d->epilogue == MGLD_EPI_GEGLU || d->epilogue == MGLD_EPI_SPADE;
  if (pair)
    MGLD_CHECK_ARG(d->N % 128 == 0 && !d->out_f32, "conv_gemm: pair epilogue needs N %% 128 == 0 (N=%d), fp16 out", d->N);
  if (d->epilogue == MGLD_EPI_SPADE)
    MGLD_CHECK_ARG(d->h && d->gn_stats && d->gn_weight && d->gn_bias && d->groups > 0 &&
                       (d->N / 2) % d->groups == 0,
                   "conv_gemm: SPADE epilogue operands missing");
  MGLD_CHECK_ARG((d->out_f32 ? d->ldout % 4 == 0 && d->out_col0 % 4 == 0 : d->ldout % 8 == 0 && d->out_col0 % 8 == 0) &&
                     (!d->res || d->ldres % 8 == 0) &&
                     (!d->h || d->ldh % 8 == 0),
                 "conv_gemm: leading dimensions / column offset must be multiples of 8");
  MGLD_CHECK_ARG(!(d->out_f32 && d->res), "conv_gemm: residual with fp32 output is not supported");

  ConvGemmParams p;
  memset(&p, 0, sizeof(p));
  p.T = d->T; p.H = d->H; p.W = d->W;
  pick_box(d->T, d->H, d->W, &p.BW, &p.BH, &p.BT);
  p.tiles_w = ceil_div(d->W, p.BW);
  p.tiles_h = ceil_div(d->H, p.BH);
  p.tiles_t = ceil_div(d->T, p.BT);
  p.tiles_m = p.tiles_w * p.tiles_h * p.tiles_t;
  p.kchunks1 = ceil_div(d->C1, 64);
  p.kchunks = ceil_div(d->C1 + d->C2, 64);
  p.taps = ntaps;
  p.tap_mode = d->taps;
  p.N = d->N;
  const int sms = num_sms();
  {
    const int pm = pair_mode_env();
    // a pair needs two M tiles.  Measured: pairs win on long-K convolutions (K >= 18 chunks) and lose on short-K GEMMs
    // (profiles/r01_dev_run12*.log); with the M-tile threshold at 16 instead of 512 the batched UNet tile-step gains 2 % and
    // the VAE decode 3-6 % (profiles/r01_dev_run35*.log, r01_dev_run36*.log; K thresholds below 18 chunks lose again).
    static const int min_tiles = [] { const char* e = getenv("MGLD_CONV_PAIR_MIN_TILES"); return e ? atoi(e) : 16; }();
    static const int min_k = [] { const char* e = getenv("MGLD_CONV_PAIR_MIN_K"); return e ? atoi(e) : 18; }();
    const bool big = p.tiles_m >= min_tiles && ntaps * p.kchunks >= min_k;
    p.cta_pair = (p.tiles_m >= 2 && (pm == 1 || (pm < 0 && big))) ? 1 : 0;
  }
  const int workers = p.cta_pair ? sms / 2 : sms;
  const int tiles_mw = p.cta_pair ? ceil_div(p.tiles_m, 2) : p.tiles_m;
  // pair epilogues: 128 accumulator columns per 64-column output panel.  Two panels per tile (block_n = 256) halve the
  // per-tile costs of the short-K FF1 / SPADE layers (tile bookkeeping, barriers, the producers' per-chunk issue time):
  // -15 % on the FF1 GEMMs of the 64x64 .. 16x16 levels, +16 .. +40 % where the tiles no longer fill two waves
  // (profiles/r02_conv_gemm_wide_pair_epilogue_raster.log) - so a wide tile is charged 1.7 narrow ones and the waves decide.
  // MGLD_CONV_PAIR_EPI_BN=128|256 forces.
  int pair_bn = 128;
  if (pair && d->N % 256 == 0) {
    const char* e = getenv("MGLD_CONV_PAIR_EPI_BN");
    const int force = e ? atoi(e) : 0;
    const int w128 = ceil_div(tiles_mw * (d->N / 128), workers), w256 = ceil_div(tiles_mw * (d->N / 256), workers);
    if (force == 256 || (force == 0 && 17 * w256 < 10 * w128)) pair_bn = 256;
  }
  p.block_n = pair ? pair_bn : (d->block_n > 0 ? d->block_n : pick_block_n(d->N, tiles_mw, workers));
  if (d->out_f32 && p.block_n > 128) p.block_n = 128;  // fp32 staging tile: 128 x 128 x 4 B = 64 KB
  MGLD_CHECK_ARG(p.block_n % 32 == 0 && p.block_n >= 32 && p.block_n <= 256, "conv_gemm: block_n=%d", p.block_n);
  p.tiles_n = ceil_div(d->N, p.block_n);
  p.split_k = split_k; p.k_per_split = ceil_div(ntaps * p.kchunks, split_k); p.slab_frames = slab_frames;
  p.n_out_tile = pair ? p.block_n / 2 : p.block_n;
  p.n_out_total = pair ? d->N / 2 : d->N;
  // columns beyond n_out_total are clipped by the TMA store; only the row pitch needs 16-byte alignment
  MGLD_CHECK_ARG(d->out_f32 ? d->ldout % 4 == 0 : d->ldout % 8 == 0, "conv_gemm: output row pitch %d not 16-byte aligned", d->ldout);
  p.panel_cols = (p.n_out_tile % 64 == 0) ? 64 : 32;
  p.n_panels = d->out_f32 ? p.n_out_tile / 32 : p.n_out_tile / p.panel_cols;
  p.acc_stride = p.block_n;  // multiple of 32 columns
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.acc_stride) p.tmem_cols <<= 1;
  p.epilogue = d->epilogue; p.act = d->act; p.bias = d->bias;
  p.alpha = d->alpha; p.beta = d->beta;
  p.has_res = d->res != nullptr; p.out_f32 = d->out_f32;
  p.gn_stats = d->gn_stats; p.gn_weight = d->gn_weight; p.gn_bias = d->gn_bias;
  p.groups = d->groups; p.ch_per_group = d->groups > 0 ? p.n_out_total / d->groups : 1;
  p.dbg = g_dbg;
  {
    // Order of the work units.  N-major (consecutive units walk the M tiles of one weight tile) keeps the weight tile
    // hot and is right for the 3x3 convolutions; M-major (consecutive units = the N tiles of one pixel tile, running
    // concurrently on neighbouring SMs) reads the activations once and wins on the 1x1 GEMMs of the large levels (FF2 at
    // 64x64: 105 MB of activations, 49.6 -> 40.0 us; the C x C projections -5..8 %) while the 3x3 convolutions of the
    // 32x32 / 16x16 levels lose 6-11 % with it (profiles/r02_conv_gemm_wide_pair_epilogue_raster.log).
    // MGLD_CONV_RASTER=0|1 forces.
    const char* e = getenv("MGLD_CONV_RASTER");
    const double a_bytes = 2.0 * d->T * d->H * d->W * (d->C1 + d->C2);
    const bool m_major = p.tiles_n > 1 && (ntaps == 1 ? a_bytes >= 12e6 : a_bytes > 48e6);
    p.raster = e ? (atoi(e) != 0 && p.tiles_n > 1) : m_major;
  }

  // GroupNorm statistics of the output in the epilogue (variant 7): plain +bias epilogue, fp16 out, a long mainloop to hide
  // the column sums behind (>= 18 K chunks: the 3x3 convolutions), tiles that lie inside one frame and inside the map
  // (every row of a tile is a real pixel), whole groups of >= 4 channels.  MGLD_CONV_FUSED_STATS=0 switches it off.
  bool fuse = false;
  if (fused_stats && *fused_stats) {
    static const int min_k = [] { const char* e = getenv("MGLD_CONV_FUSED_STATS_MIN_K"); return e ? atoi(e) : 18; }();
    const char* ev = getenv("MGLD_CONV_VARIANT");
    const int cpg = d->stats_groups > 0 ? d->N / d->stats_groups : 0;
    fuse = d->epilogue == MGLD_EPI_LINEAR && d->act == MGLD_ACT_NONE && !d->res && !d->out_f32 && d->alpha == 1.f &&
           split_k == 1 && p.BT == 1 && d->W % p.BW == 0 && d->H % p.BH == 0 && ntaps * p.kchunks >= min_k &&
           d->stats_groups > 0 && d->N % d->stats_groups == 0 && cpg >= 4 && !(ev && atoi(ev) == 0);
    p.stats_out = d->stats_out; p.stats_groups = d->stats_groups; p.stats_cpg = cpg > 0 ? cpg : 1;
    *fused_stats = fuse;
  }

  // shared memory plan: [A/B ring][staging panels][h panel (SPADE)][bias][statistics (variant 7)]
  const int stage_bytes = kABytes + (p.block_n / (p.cta_pair ? 2 : 1)) * 128;
  const int staging_bytes = d->out_f32 ? p.n_panels * kPanelBytes : p.n_out_tile * kBlockM * 2;
  const int hstage_bytes = d->epilogue == MGLD_EPI_SPADE ? p.n_panels * kPanelBytes : 0;
  const int stats_bytes = fuse ? kStatsBytes : 0;
  const int fixed = staging_bytes + hstage_bytes + 2048 /*bias, two tiles*/ + stats_bytes + 1024 /*alignment slack*/;
  int stages = (224 * 1024 - fixed) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  MGLD_CHECK_ARG(stages >= 2, "conv_gemm: tile does not fit in shared memory (block_n=%d)", p.block_n);
  p.stages = stages;
  p.off_staging = stages * stage_bytes;
  p.off_hstage = p.off_staging + staging_bytes;
  p.off_bias = p.off_hstage + hstage_bytes;
  p.off_stats = p.off_bias + 2048;
  const int smem = p.off_stats + stats_bytes + 1024;

  // tensor maps
  const int lda = d->lda > 0 ? d->lda : d->C1;
  const int lda2 = d->lda2 > 0 ? d->lda2 : d->C2;
  CUtensorMap tmA, tmA2, tmB, tmOut, tmRes, tmH;
  {
    auto nhwc_map = [&](CUtensorMap* m, const void* base, int C, int ld, int box_cols = 64) {
      uint32_t box[4] = {(uint32_t)box_cols, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BT};
      uint64_t dims[4] = {(uint64_t)C, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->T};
      uint64_t str[3] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * d->W, (uint64_t)ld * 2 * d->W * d->H};
      return make_tmap_f16(m, base, 4, dims, str, box);
    };
    int rc = nhwc_map(&tmA, d->a, d->C1, lda);
    if (rc) return rc;
    if (d->a2) { rc = nhwc_map(&tmA2, d->a2, d->C2, lda2); if (rc) return rc; }
    else tmA2 = tmA;
    const uint64_t K = (uint64_t)ntaps * (d->C1 + d->C2);
    uint64_t dimsB[2] = {K, (uint64_t)d->N};
    uint64_t strB[1] = {K * 2};
    uint32_t boxB[2] = {64, (uint32_t)(p.block_n / 2)};   // each of the two B-producer threads loads half of the N tile
    rc = make_tmap_f16(&tmB, d->w, 2, dimsB, strB, boxB);
    if (rc) return rc;
    if (d->out_f32) {
      uint64_t dims[4] = {(uint64_t)p.n_out_total, (uint64_t)d->W, (uint64_t)d->H,
                          (uint64_t)(split_k > 1 ? split_k * slab_frames : d->T)};
      uint64_t str[3] = {(uint64_t)d->ldout * 4, (uint64_t)d->ldout * 4 * d->W, (uint64_t)d->ldout * 4 * d->W * d->H};
      uint32_t box32[4] = {32, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BT};
      rc = make_tmap_f32(&tmOut, reinterpret_cast<const float*>(d->out) + d->out_col0, 4, dims, str, box32);
    } else {
      rc = nhwc_map(&tmOut, reinterpret_cast<const __half*>(d->out) + d->out_col0, p.n_out_total, d->ldout, p.panel_cols);
    }
    if (rc) return rc;
    if (d->res) { rc = nhwc_map(&tmRes, d->res, p.n_out_total, d->ldres, p.panel_cols); if (rc) return rc; }
    else tmRes = tmA;
    if (d->epilogue == MGLD_EPI_SPADE) { rc = nhwc_map(&tmH, d->h, p.n_out_total, d->ldh); if (rc) return rc; }
    else tmH = tmA;
  }

  using KernelFn = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, CUtensorMap, ConvGemmParams);
  static const KernelFn kKernels[2][8] = {
      {conv_gemm_kernel<false, 0>, conv_gemm_kernel<false, 1>, conv_gemm_kernel<false, 2>, conv_gemm_kernel<false, 3>,
       conv_gemm_kernel<false, 4>, conv_gemm_kernel<false, 5>, conv_gemm_kernel<false, 6>, conv_gemm_kernel<false, 7>},
      {conv_gemm_kernel<true, 0>, conv_gemm_kernel<true, 1>, conv_gemm_kernel<true, 2>, conv_gemm_kernel<true, 3>,
       conv_gemm_kernel<true, 4>, conv_gemm_kernel<true, 5>, conv_gemm_kernel<true, 6>, conv_gemm_kernel<true, 7>}};
  if (!g_attr_set) {
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 8; ++j)
        MGLD_CUDA(cudaFuncSetAttribute(kKernels[i][j], cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    g_attr_set = true;
  }
  // epilogue variant with compile-time flags (see the kernel's comment); MGLD_CONV_VARIANT=0 forces the generic kernel
  int variant = 0;
  if (d->epilogue == MGLD_EPI_LINEAR && d->act == MGLD_ACT_NONE) {
    if (!d->res && !d->out_f32 && d->alpha == 1.f) variant = 1;
    else if (d->res && !d->out_f32) variant = 2;
    else if (!d->res && d->out_f32 && d->alpha == 1.f) variant = 3;
  } else if (d->epilogue == MGLD_EPI_LINEAR && d->act == MGLD_ACT_SILU && !d->res && !d->out_f32 && d->alpha == 1.f) {
    variant = 4;
  } else if (d->epilogue == MGLD_EPI_GEGLU && !d->res) {
    variant = 5;
  } else if (d->epilogue == MGLD_EPI_SPADE) {
    variant = 6;
  }
  { const char* ev = getenv("MGLD_CONV_VARIANT"); if (ev && atoi(ev) == 0) variant = 0; }
  if (fuse) variant = 7;
  const KernelFn kernel = kKernels[p.cta_pair ? 1 : 0][variant];
  const int total_units = tiles_mw * p.tiles_n * p.split_k;
  if (!p.cta_pair) {
    dim3 grid(total_units < sms ? total_units : sms, 1, 1);
    MGLD_CUDA(launch_pdl(kernel, grid, dim3(kThreads), smem, stream, tmA, tmA2, tmB, tmOut, tmRes, tmH, p));
  } else {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.blockDim = dim3(kThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (g_max_pairs < 0) {
      cfg.gridDim = dim3(2 * (sms / 2), 1, 1);
      cfg.dynamicSmemBytes = 226 * 1024;
      int n = 0;
      MGLD_CUDA(cudaOccupancyMaxActiveClusters(&n, kKernels[1][0], &cfg));
      g_max_pairs = n > 0 ? (n < sms / 2 ? n : sms / 2) : 1;
      cfg.dynamicSmemBytes = smem;
    }
    const int pairs = total_units < g_max_pairs ? total_units : g_max_pairs;
    cfg.gridDim = dim3(2 * pairs, 1, 1);
    MGLD_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmA2, tmB, tmOut, tmRes, tmH, p));
  }
  MGLD_LAUNCH_CHECK("conv_gemm_kernel");
  return MGLD_OK;
}

// ---- split-K path: pass 1 = raw fp32 partial tiles into workspace slabs, pass 2 = deterministic sum + epilogue ---------
static bool split_plan(const mgld_conv_gemm_desc* d, int* S, int* bn, size_t* bytes, int* ldws, int* slab_frames) {
  if (d->out_f32) return false;
  const bool pair = d->epilogue == MGLD_EPI_GEGLU || d->epilogue == MGLD_EPI_SPADE;
  int BW, BH, BT;
  pick_box(d->T, d->H, d->W, &BW, &BH, &BT);
  const int tiles_m = ceil_div(d->W, BW) * ceil_div(d->H, BH) * ceil_div(d->T, BT);
  const int ntaps = d->taps == MGLD_TAPS_5X1 ? 5 : d->taps;
  const int num_k = ntaps * ceil_div(d->C1 + d->C2, 64);
  const int block_n = pair ? 128 : (d->N % 128 == 0 ? 128 : (d->N % 64 == 0 ? 64 : 0));
  if (!block_n) return false;
  int s = pick_split_k(tiles_m * (d->N / block_n), num_k, num_sms());
  while (s > 1 && (s - 1) * ceil_div(num_k, s) >= num_k) --s;   // every K-slice must be non-empty
  if (s <= 1) return false;
  *S = s; *bn = block_n;
  *ldws = d->N;                                        // raw accumulator columns (multiple of 64)
  *slab_frames = ceil_div(d->T, BT) * BT;              // frames padded to whole boxes: slabs never overlap
  *bytes = (size_t)s * *slab_frames * d->H * d->W * *ldws * sizeof(float);
  return true;
}

extern "C" long long mgld_conv_gemm_workspace_bytes(const mgld_conv_gemm_desc* d) {
  int S, bn, ldws, sf;
  size_t bytes;
  if (!d || !initialised() || !split_plan(d, &S, &bn, &bytes, &ldws, &sf)) return 0;
  return (long long)bytes;
}

static int conv_gemm_dispatch(const mgld_conv_gemm_desc* d, void* stream_, bool* fused_stats);

extern "C" int mgld_conv_gemm(const mgld_conv_gemm_desc* d, void* stream_) {
  if (!initialised()) { set_error("mgld_init() has not been called"); return MGLD_ERR_NOT_INIT; }
  MGLD_CHECK_ARG(d && d->a && d->w && d->out, "conv_gemm: null pointer");
  if (d->stats_out) MGLD_CHECK_ARG(!d->out_f32 && d->stats_groups > 0, "conv_gemm: stats_out needs fp16 output and stats_groups > 0");
  bool fused = false;
  int rc = conv_gemm_dispatch(d, stream_, &fused);
  if (rc || !d->stats_out || fused) return rc;
  // GroupNorm statistics of the output for its consumer: a streaming pass over the (L2-resident) output.  Only the long-K
  // convolutions accumulate them in their epilogue (variant 7): for short-K GEMMs the epilogue is the critical path and the
  // extra work was measured slower than this pass (profiles/r01_dev_run7_fused_stats_regression.log).
  const int n_out = d->epilogue == MGLD_EPI_LINEAR ? d->N : d->N / 2;
  return mgld_gn_stats_f16(reinterpret_cast<const __half*>(d->out) + d->out_col0, n_out, d->ldout, nullptr, 0, 0, d->T,
                           d->H * d->W, d->stats_groups, d->stats_out, stream_);
}

static int conv_gemm_dispatch(const mgld_conv_gemm_desc* d, void* stream_, bool* fused_stats) {
  int S, bn, ldws, sf;
  size_t bytes;
  if (!d->workspace || !split_plan(d, &S, &bn, &bytes, &ldws, &sf) || (size_t)d->workspace_bytes < bytes) {
    const char* ef = getenv("MGLD_CONV_FUSED_STATS");   // read per call (the tests toggle it)
    const bool fuse_on = ef ? atoi(ef) != 0 : false;
    *fused_stats = fuse_on && d->stats_out != nullptr;
    return conv_gemm_single(d, stream_, 1, 0, fused_stats);
  }
  *fused_stats = false;
  // pass 1: raw partial sums.  The descriptor is rewritten to a plain fp32-output GEMM into the workspace.
  mgld_conv_gemm_desc r = *d;
  r.epilogue = MGLD_EPI_LINEAR; r.act = MGLD_ACT_NONE; r.bias = nullptr; r.alpha = 1.f; r.beta = 0.f; r.res = nullptr;
  r.h = nullptr; r.gn_stats = nullptr; r.out = d->workspace; r.ldout = ldws; r.out_col0 = 0; r.out_f32 = 1; r.block_n = bn;
  int rc = conv_gemm_single(&r, stream_, S, sf);
  if (rc) return rc;
  SplitFinalizeParams f;
  memset(&f, 0, sizeof(f));
  f.ws = reinterpret_cast<const float*>(d->workspace);
  f.slab_stride = (long long)sf * d->H * d->W * ldws;
  f.S = S; f.T = d->T; f.H = d->H; f.W = d->W; f.ldws = ldws;
  f.N = d->N; f.epilogue = d->epilogue; f.act = d->act; f.bias = d->bias; f.alpha = d->alpha; f.beta = d->beta;
  f.n_out_total = (d->epilogue == MGLD_EPI_LINEAR) ? d->N : d->N / 2;
  f.res = reinterpret_cast<const __half*>(d->res); f.ldres = d->ldres;
  f.h = reinterpret_cast<const __half*>(d->h); f.ldh = d->ldh;
  f.gn_stats = d->gn_stats; f.gn_weight = d->gn_weight; f.gn_bias = d->gn_bias; f.groups = d->groups;
  f.ch_per_group = d->groups > 0 ? f.n_out_total / d->groups : 1;
  f.out = reinterpret_cast<__half*>(d->out); f.ldout = d->ldout; f.out_col0 = d->out_col0;
  MGLD_CHECK_ARG(f.n_out_total % 8 == 0, "conv_gemm(split-K): output columns must be a multiple of 8");
  const long long total = (long long)d->T * d->H * d->W * (f.n_out_total / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  MGLD_CUDA(launch_pdl(splitk_finalize_kernel, dim3(blocks), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream_), f));
  MGLD_LAUNCH_CHECK("splitk_finalize_kernel");
  return MGLD_OK;
}

// Development hook: per-CTA cycle counters of the next conv_gemm launches ([grid][16] int64 in device memory: A-producer
// wait-on-empty, A-producer total, MMA wait-on-full, MMA wait-on-accumulator, MMA total, epilogue wait, epilogue total,
// tiles, B-producer wait-on-empty, B-producer total).  Pass null to switch off.
extern "C" void mgld_conv_gemm_set_debug_counters(void* dev_ptr) { g_dbg = reinterpret_cast<long long*>(dev_ptr); }
