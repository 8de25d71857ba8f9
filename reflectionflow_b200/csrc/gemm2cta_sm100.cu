// gemm2cta_sm100.cu — the big-GEMM path: CTA-pair (tcgen05 cta_group::2) persistent bf16 GEMM.
//
//   out[M, N] = epilogue( A[M, K] @ W[N, K]^T ),  N % 256 == 0,  K % 64 == 0
//
// Same contract and epilogues as gemm_sm100.cu (which keeps serving the narrow shapes); this
// kernel exists because a single-CTA 128x256 tile pulls 96 B/clk/SM out of L2 and a 4-stage ring
// cannot cover the loaded L2 latency (tools/trace_gemm.py: the MMA thread waits ~40 % of a
// K=3072 tile for TMA data).  Here two CTAs of a cluster share one 256x256 output tile:
//   * each CTA loads its own 128 rows of A and HALF of the W tile (128 of 256 rows): 64 B/clk/SM,
//     32 KB per stage -> a 5-stage ring;
//   * the leader CTA issues tcgen05.mma.cta_group::2 (M = 256 across the pair, N = 256); the
//     accumulator rows of each CTA land in its own TMEM (2 stages x 256 columns);
//   * TMA transaction bytes of both CTAs are credited to the leader's full barrier; smem slots and
//     accumulator stages are released with multicast tcgen05.commit.
// Epilogue (warps 4..7 of both CTAs): TMEM -> registers -> fused math -> 128B-swizzled smem box
// [128 rows x 64 cols] -> TMA store (coalesced, asynchronous, clips ragged M); the residual of the
// gate+residual epilogue is prefetched by TMA one chunk ahead into smem.  Tiles are rastered in
// 12-tile-wide column bands so a wave's W slice (19 MB at K=3072) stays L2-resident.
#include <cstdlib>
#include <cuda.h>

#include "rf_internal.h"
#include "rf_ptx.cuh"

namespace rf {

static constexpr int kThreads2 = 256;
static constexpr int kRows = 128;  // rows of A / of the output per CTA
static constexpr int kBK = 64;
static constexpr int kStageA = kRows * kBK * 2;        // 16 KB
static constexpr int kBox = kRows * 64 * 2;            // 16 KB epilogue box
static constexpr int kMaxGroups2 = 3;
template <int BN>  // pair-tile N: 256 (DiT projections, wide convs) or 128 (128-channel convs)
struct Cfg2 {
  static constexpr int kStageB = (BN / 2) * kBK * 2;   // this CTA's half of the W tile
  static constexpr int kStage = kStageA + kStageB;     // 32 KB / 24 KB
  static constexpr int kStages = (BN == 256) ? 5 : 6;
  static constexpr int kSmem = kStages * kStage + 4 * kBox + 1024 + 512;
};

struct alignas(64) Gemm2Group {
  CUtensorMap tmA, tmB, tmOut, tmRes;
  // fused peft LoRA (kLora kernels): T = bf16(x A^T) [M, 64 per section] and the B factor [N, 64];
  // a tile in column section s = n0 / lora_sec_cols reads T columns [64 s, 64 s + 64)
  CUtensorMap tmT, tmLB;
  int lora_sec_cols;
  // per-group tile raster
  int bn, n_tiles, band, lora;
  const bf16* bias;
  const bf16* addend;
  const bf16* gate;
  const float* rope_cos;
  const float* rope_sin;
  const bf16* norm_q;
  const bf16* norm_k;
  int M, ldadd, m_pairs, tile_begin;
  // implicit-GEMM convolution over zero-ringed NHWC images (conv_w > 0; tmA/tmOut/tmRes are 3-D maps
  // {C, W+2, H+2}): an M tile is conv_by rows x conv_bx columns of OUTPUT pixels (bx * by = 128);
  // tap (ky, kx) of a k-block reads input pixels (s*y + ky + s-1, s*x + kx + s-1), s = conv_stride
  int conv_w, conv_taps, conv_cin_blocks, conv_bx, conv_by, conv_stride;
};
struct PixTile { int x0, y0; };
__device__ __forceinline__ PixTile pix_tile(const Gemm2Group& G, int m) {
  const int mt = m >> 7;                 // 128-pixel tile index
  const int xt = G.conv_w / G.conv_bx;   // tiles per image row
  const int ty = mt / xt;
  return PixTile{(mt - ty * xt) * G.conv_bx, ty * G.conv_by};
}
struct alignas(64) Gemm2Params {
  Gemm2Group g[kMaxGroups2];
  int ngroups, N, K, n_tiles, total_tiles, num_kb, band, bn;
  long long* trace;  // dev-only per-tile timeline of pair 0 (rf_dbg_set_gemm_trace)
  int dbg_skip;      // dev-only: 1 = skip the W loads, 2 = skip the A loads (timing experiments; results are garbage)
};
struct Tile2 {
  int g, m0, n0;  // m0: first row of the 256-row pair tile, n0: first column
};
__device__ __forceinline__ Tile2 decode2(const Gemm2Params& p, int t) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroups2; ++i)
    if (i < p.ngroups && t >= p.g[i].tile_begin) g = i;
  const int local = t - p.g[g].tile_begin;
  const int mp = p.g[g].m_pairs;
  const int n_tiles = p.g[g].n_tiles, band = p.g[g].band;
  const int full = n_tiles / band;
  const int full_tiles = full * mp * band;
  int m, n;
  if (local < full_tiles) {
    const int per = mp * band;
    const int sc = local / per, r = local - sc * per;
    m = r / band;
    n = sc * band + (r - m * band);
  } else {
    const int rem = n_tiles - full * band;
    const int r = local - full_tiles;
    m = r / rem;
    n = full * band + (r - m * rem);
  }
  return Tile2{g, m * 2 * kRows, n * p.g[g].bn};
}

__device__ __forceinline__ void ld8(const bf16* p, float* v) {  // 8 bf16 (16 B aligned) -> fp32
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
// v[64] = bf16(acc + bias) [ = bf16(v + addend) ]: nn.Linear (+ peft LoRA term) rounding points
__device__ __forceinline__ void linear_round64(const uint32_t (&acc)[64], const bf16* bias,
                                               const bf16* addend, float (&v)[64]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    float b[8];
    if (bias != nullptr) ld8(bias + q * 8, b);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      v[q * 8 + i] = __uint_as_float(acc[q * 8 + i]) + (bias != nullptr ? b[i] : 0.f);
      v[q * 8 + i + 1] = __uint_as_float(acc[q * 8 + i + 1]) + (bias != nullptr ? b[i + 1] : 0.f);
      bf16_round2(v[q * 8 + i], v[q * 8 + i + 1]);
    }
  }
  if (addend != nullptr) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float a[8];
      ld8(addend + q * 8, a);
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        v[q * 8 + i] += a[i];
        v[q * 8 + i + 1] += a[i + 1];
        bf16_round2(v[q * 8 + i], v[q * 8 + i + 1]);
      }
    }
  }
}
// row r of a [128 x 64] bf16 box in the 128B-swizzled layout TMA expects
__device__ __forceinline__ void box_store_row(uint8_t* box, int r, const float (&v)[64]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    uint4 u;
    u.x = pack_bf16x2(v[q * 8 + 0], v[q * 8 + 1]);
    u.y = pack_bf16x2(v[q * 8 + 2], v[q * 8 + 3]);
    u.z = pack_bf16x2(v[q * 8 + 4], v[q * 8 + 5]);
    u.w = pack_bf16x2(v[q * 8 + 6], v[q * 8 + 7]);
    *reinterpret_cast<uint4*>(box + r * 128 + ((q ^ (r & 7)) << 4)) = u;
  }
}
__device__ __forceinline__ void box_load_row(const uint8_t* box, int r, float (&v)[64]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const uint4 u = *reinterpret_cast<const uint4*>(box + r * 128 + ((q ^ (r & 7)) << 4));
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    v[q * 8 + 0] = a.x; v[q * 8 + 1] = a.y; v[q * 8 + 2] = b.x; v[q * 8 + 3] = b.y;
    v[q * 8 + 4] = c.x; v[q * 8 + 5] = c.y; v[q * 8 + 6] = d.x; v[q * 8 + 7] = d.y;
  }
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&acc)[64]) {
  uint32_t(&lo)[32] = *reinterpret_cast<uint32_t(*)[32]>(&acc[0]);
  uint32_t(&hi)[32] = *reinterpret_cast<uint32_t(*)[32]>(&acc[32]);
  tmem_ld_32x32(taddr, lo);
  tmem_ld_32x32(taddr + 32, hi);
  tmem_ld_wait();
}

// kLora: the tile also accumulates L = T B^T (ONE extra 64-deep k-block, its own TMEM accumulator) and the
// epilogue forms bf16(bf16(acc + bias) + bf16(L)) — peft's unfused LoRA arithmetic (lora_controller.py:5-42)
// without materialising L in HBM.  kBN = 128: two accumulator stages of (128 base + 128 L); kBN = 256: one stage
// of (256 base + 256 L) — these launches are bound by L2 -> SM operand traffic, and the 256-wide tile needs a
// third less of it per FLOP (ncu, K = 15360 condition stream: 141 us at 128 wide).
// kEG = 2 (kBN == 256): TWO epilogue warpgroups (warps 4..7 and 8..11, 384 threads).  A TMEM lane quadrant can
// only be read by warps of one SM sub-partition, so with one warpgroup every scheduler sees a single epilogue
// warp and nothing hides its dependent-issue latency: the K = 3072 GEMMs were epilogue-bound (30-35 k cycles per
// tile against a 27 k mainloop).  Group g owns columns [128 g, 128 g + 128) of every tile (one attention head of
// the QKV epilogue) with its own staging box, residual box and named barrier.
template <int EPI, int kBN, bool kLora = false, int kEG = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads2 + 128 * (kEG - 1), 1)
gemm2_kernel(const __grid_constant__ Gemm2Params p) {
  static_assert(kEG == 1 || (kEG == 2 && kBN == 256), "two epilogue groups split a 256-wide tile");
  // TMEM: kLora/128: 2 stages x (128 base + 128 L); kLora/256: ONE stage of 256 base + 256 L (the 256-wide tile
  // halves the L2 traffic per FLOP of the 128-wide one, which is what bounds these launches; the price is
  // that a tile's epilogue no longer overlaps the next mainloop); otherwise 2 stages x kBN
  constexpr int kAccStages = (kLora && kBN == 256) ? 1 : 2;
  constexpr int kTmemCols = kLora ? 512 : 2 * kBN;
  constexpr uint32_t kLoraOff = (kBN == 256) ? 256u : 2u * kBN;  // column offset of the low-rank accumulator
  constexpr int kStages = Cfg2<kBN>::kStages;
  constexpr int kStage = Cfg2<kBN>::kStage;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* out_box = smem + kStages * kStage;  // 2 boxes
  uint8_t* res_box = out_box + 2 * kBox;       // 2 boxes
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(res_box + 2 * kBox);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.ngroups; ++g) {
      tma_prefetch_desc(&p.g[g].tmA);
      tma_prefetch_desc(&p.g[g].tmB);
      tma_prefetch_desc(&p.g[g].tmOut);
      if constexpr (kLora) {
        tma_prefetch_desc(&p.g[g].tmT);
        tma_prefetch_desc(&p.g[g].tmLB);
      }
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 8 * kEG);  // every epilogue warp of both CTAs arrives on the leader's barrier
      mbar_init(&res_bar[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2cta<kTmemCols>(tmem_slot);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();  // barrier inits + TMEM allocation of BOTH CTAs visible before any remote op
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // everything above overlapped the previous kernel's tail; operands are ready from here on

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer (both CTAs) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int t = pair; t < p.total_tiles; t += npairs) {
      const Tile2 tc = decode2(p, t);
      const Gemm2Group& G = p.g[tc.g];
      const int my_m = tc.m0 + rank * kRows;
      constexpr int bnt = kBN;
      constexpr bool lora_tile = kLora;
      const uint32_t stage_tx = 2u * (kStageA + (bnt / 2) * kBK * 2);  // both CTAs: A rows + their half of W
      const int my_n = tc.n0 + rank * (bnt / 2);
      PixTile pt{0, 0};
      if (G.conv_w != 0) pt = pix_tile(G, my_m);
      if (lora_tile) {  // the low-rank k-block first: T rows of this CTA, B rows of its half tile
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStage;
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
        const int tcol = G.lora_sec_cols > 0 ? (tc.n0 / G.lora_sec_cols) * 64 : 0;
        tma_load_2d_2cta(sa, &G.tmT, &full_bar[stage], tcol, my_m);
        tma_load_2d_2cta(sa + kStageA, &G.tmLB, &full_bar[stage], 0, my_n);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStage;
#ifdef RF_DEV_HOOKS  // timing experiments of DESIGN.md §9 (make DEV=1): drop one operand stream
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], p.dbg_skip == 1 ? 2 * kStageA : p.dbg_skip == 2 ? 2 * Cfg2<kBN>::kStageB : 2 * kStage);
        if (p.dbg_skip == 2) {
        } else
#else
        if (leader) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
#endif
        if (G.conv_w != 0) {
          const int tap = kb / G.conv_cin_blocks;
          const int c0 = (kb - tap * G.conv_cin_blocks) * kBK;
          const int ky = G.conv_taps == 9 ? tap / 3 : 1, kx = G.conv_taps == 9 ? tap % 3 : 1;
          const int sft = G.conv_stride - 1;
          tma_load_3d_2cta(sa, &G.tmA, &full_bar[stage], c0, G.conv_stride * pt.x0 + kx + sft,
                           G.conv_stride * pt.y0 + ky + sft);
        } else {
          tma_load_2d_2cta(sa, &G.tmA, &full_bar[stage], kb * kBK, my_m);
        }
#ifdef RF_DEV_HOOKS
        if (p.dbg_skip != 1)
#endif
        tma_load_2d_2cta(sa + kStageA, &G.tmB, &full_bar[stage], kb * kBK, my_n);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && leader) {
    // ===================== MMA issuer (leader CTA only) =====================
    constexpr uint32_t idesc = make_idesc_bf16(256, kBN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    const bool tr = p.trace != nullptr && pair == 0;
    int ti = 0;
    for (int t = pair; t < p.total_tiles; t += npairs, ++ti) {
      if (tr && ti < 16) p.trace[ti * 8 + 0] = clock64();
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      if (tr && ti < 16) p.trace[ti * 8 + 1] = clock64();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * kBN;
      long long stall = 0;
      constexpr bool lora_tile = kLora;
      constexpr uint32_t idesc_t = idesc;
      constexpr uint32_t l_off = kLoraOff;  // column offset of the low-rank accumulator
      if (lora_tile) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStage);
        const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sa + kStageA, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          mma_ss_2cta(d_tmem + l_off, adesc + 2 * k, bdesc + 2 * k, idesc_t, k != 0 ? 1u : 0u);
        tc_commit_2cta(&empty_bar[stage], 3);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      for (int kb = 0; kb < p.num_kb; ++kb) {
        if (tr) {
          const long long c0 = clock64();
          mbar_wait(&full_bar[stage], phase);
          stall += clock64() - c0;
        } else {
          mbar_wait(&full_bar[stage], phase);
        }
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kStage);
        const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sa + kStageA, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k)
          mma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc_t, (kb | k) != 0 ? 1u : 0u);
        tc_commit_2cta(&empty_bar[stage], 3);  // slot free in both CTAs
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      tc_commit_2cta(&tfull_bar[as], 3);  // accumulator complete in both CTAs
      if (tr && ti < 16) { p.trace[ti * 8 + 2] = stall; p.trace[ti * 8 + 3] = clock64(); }
      if (++as == kAccStages) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs) =====================
    const int ew = warp & 3;
    const int eg = (kEG == 2) ? ((warp - 4) >> 2) : 0;  // epilogue group: columns [128 eg, 128 eg + 128) when kEG == 2
    const int r_in = ew * 32 + lane;          // row inside this CTA's 128 rows == TMEM lane
    const bool issuer = (ew == 0 && lane == 0);  // one per group
    const uint32_t bar_id = 1 + eg;
    const uint32_t lane_off = static_cast<uint32_t>(ew * 32) << 16;
    int as = 0;
    uint32_t aphase = 0;
    // running 64-column chunk counter of this group.  kEG == 1: two staging / residual boxes alternate by
    // cc & 1; kEG == 2: each group has ONE of each (box index eg), the other group is what overlaps the waits.
    uint32_t cc = 0;
    const int c_first = (kEG == 2) ? 2 * eg : 0;

    // residual prefetch of chunk `nc` of tile `nt` into box `nb` (EPI_GATE_RES)
    auto prefetch_res = [&](int nt, int nc, uint32_t nb) {
      if (nt >= p.total_tiles) return;
      const Tile2 tn = decode2(p, nt);
      mbar_arrive_expect_tx(&res_bar[nb], kBox);
      if (p.g[tn.g].conv_w != 0) {
        const PixTile q = pix_tile(p.g[tn.g], tn.m0 + rank * kRows);
        tma_load_3d(res_box + nb * kBox, &p.g[tn.g].tmRes, &res_bar[nb], tn.n0 + nc * 64, q.x0 + 1, q.y0 + 1);
      } else {
        tma_load_2d(res_box + nb * kBox, &p.g[tn.g].tmRes, &res_bar[nb], tn.n0 + nc * 64, tn.m0 + rank * kRows);
      }
    };
    if constexpr (EPI == EPI_GATE_RES) {
      if (issuer) prefetch_res(pair, c_first, kEG == 2 ? eg : 0);  // first chunk of the first tile
    }
    const bool etr = p.trace != nullptr && pair == 0 && leader && issuer && eg == 0;
    int eti = 0;
    for (int t = pair; t < p.total_tiles; t += npairs, ++eti) {
      const Tile2 tc = decode2(p, t);
      const Gemm2Group& G = p.g[tc.g];
      const int my_m = tc.m0 + rank * kRows;
      const int row = my_m + r_in;
      const int row_c = row < G.M ? row : G.M - 1;  // clamped row for direct global reads
      if (etr && eti < 16) p.trace[eti * 8 + 7] = clock64();
      mbar_wait(&tfull_bar[as], aphase);
      if (etr && eti < 16) p.trace[eti * 8 + 4] = clock64();
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_off + as * kBN;
      constexpr int bnt = kBN;
      constexpr uint32_t l_off = kLoraOff;
      // v = bf16(v + bf16(L)): the low-rank term of this 64-column chunk from its own accumulator
      auto lora_add = [&](uint32_t col, float (&v)[64]) {
        if constexpr (kLora) {
          uint32_t accl[64];
          tmem_ld64(taddr + l_off + col, accl);
#pragma unroll
          for (int i = 0; i < 64; i += 2) {
            float l0 = __uint_as_float(accl[i]), l1 = __uint_as_float(accl[i + 1]);
            bf16_round2(l0, l1);
            v[i] += l0;
            v[i + 1] += l1;
            bf16_round2(v[i], v[i + 1]);
          }
        }
      };

      // issue the box store of chunk `cc` (all 128 threads of the group call this); `nt`/`nc`: the group's next
      // chunk, whose residual (kEG == 2) is fetched into the single residual box once everybody has read it
      auto publish = [&](const float (&v)[64], int col, int nt, int nc) {
        uint8_t* ob = out_box + (kEG == 2 ? eg : (cc & 1)) * kBox;
        if (issuer) {  // the store that last used this box has drained
          if constexpr (kEG == 2) tma_store_wait_read<0>(); else tma_store_wait_read<1>();
        }
        named_bar_sync(bar_id, 128);
        if constexpr (EPI == EPI_GATE_RES && kEG == 2) {
          if (issuer) prefetch_res(nt, nc, eg);
        }
        box_store_row(ob, r_in, v);
        fence_proxy_async_smem();
        named_bar_sync(bar_id, 128);
        if (issuer) {
          if (G.conv_w != 0) {
            const PixTile q = pix_tile(G, my_m);
            tma_store_3d(&G.tmOut, ob, col, q.x0 + 1, q.y0 + 1);
          } else {
            tma_store_2d(&G.tmOut, ob, col, my_m);
          }
          tma_store_commit();
        }
      };

      if constexpr (EPI == EPI_QKV) {
        static_assert(kBN % 128 == 0, "QKV epilogue: a head is 128 columns");
        const int inner = p.N / 3;
        const float* cosr = G.rope_cos + static_cast<size_t>(row_c) * 64;
        const float* sinr = G.rope_sin + static_cast<size_t>(row_c) * 64;
#pragma unroll 1
        for (int hc = (kEG == 2 ? eg : 0); hc < (kEG == 2 ? eg + 1 : bnt / 128); ++hc) {
          const int col_h = tc.n0 + hc * 128;
          const int section = col_h / inner;  // 0 q, 1 k, 2 v
          const bf16* bias_h = G.bias ? G.bias + col_h : nullptr;
          const bf16* add_h = G.addend ? G.addend + static_cast<size_t>(row_c) * G.ldadd + col_h : nullptr;
          float rinv = 0.f;
          // ONE pass over TMEM per head: the bf16-rounded linear output of the 128 columns stays in
          // registers as packed bf16 pairs (exact: the values are already bf16) between the
          // sum-of-squares pass and the normalise / RoPE pass
          uint32_t keep[2][32];
          if (section != 2) {
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t acc[64];
              tmem_ld64(taddr + hc * 128 + c * 64, acc);
              float v[64];
              linear_round64(acc, bias_h ? bias_h + c * 64 : nullptr, add_h ? add_h + c * 64 : nullptr, v);
              lora_add(hc * 128 + c * 64, v);
#pragma unroll
              for (int i = 0; i < 64; i += 2) {
                ss = __fmaf_rn(v[i], v[i], ss);
                ss = __fmaf_rn(v[i + 1], v[i + 1], ss);
                keep[c][i >> 1] = pack_bf16x2(v[i], v[i + 1]);
              }
            }
            const float var = __fdiv_rn(ss, 128.0f);
            rinv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-6f)));
          }
          const bf16* nw = (section == 0) ? G.norm_q : G.norm_k;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[64];
            if (section != 2) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                float w[8];
                ld8(nw + c * 64 + q * 8, w);
                // this row's 4 (cos, sin) pairs of the 8 columns
                const float4 cq = __ldg(reinterpret_cast<const float4*>(cosr + c * 32) + q);
                const float4 sq = __ldg(reinterpret_cast<const float4*>(sinr + c * 32) + q);
                const float cs[4] = {cq.x, cq.y, cq.z, cq.w}, sn[4] = {sq.x, sq.y, sq.z, sq.w};
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                  const float2 x = unpack_bf16x2(keep[c][(q * 8 + i) >> 1]);
                  float y0 = __fmul_rn(x.x, rinv), y1 = __fmul_rn(x.y, rinv);
                  bf16_round2(y0, y1);  // RMSNorm -> bf16
                  y0 = __fmul_rn(y0, w[i]);
                  y1 = __fmul_rn(y1, w[i + 1]);
                  bf16_round2(y0, y1);  // * weight -> bf16
                  // interleaved-pair RoPE, fp32, one rounding at the pack
                  v[q * 8 + i] = __fadd_rn(__fmul_rn(y0, cs[i >> 1]), __fmul_rn(-y1, sn[i >> 1]));
                  v[q * 8 + i + 1] = __fadd_rn(__fmul_rn(y1, cs[i >> 1]), __fmul_rn(y0, sn[i >> 1]));
                }
              }
            } else {
              uint32_t acc[64];
              tmem_ld64(taddr + hc * 128 + c * 64, acc);
              linear_round64(acc, bias_h ? bias_h + c * 64 : nullptr, add_h ? add_h + c * 64 : nullptr, v);
              lora_add(hc * 128 + c * 64, v);
            }
            publish(v, col_h + c * 64, 0, 0);
            ++cc;
          }
        }
      } else {
        const bf16* add_r = G.addend ? G.addend + static_cast<size_t>(row_c) * G.ldadd + tc.n0 : nullptr;
        const int c_end = (kEG == 2) ? c_first + 2 : bnt / 64;
#pragma unroll 1
        for (int c = c_first; c < c_end; ++c) {
          float r[64];
          int nt = t, nc = c + 1;  // this group's next chunk
          if (nc == c_end) { nt = t + npairs; nc = c_first; }
          if constexpr (EPI == EPI_GATE_RES) {
            if constexpr (kEG == 1) {
              // prefetch the residual of the NEXT chunk into the other box (everybody finished reading
              // it before the barriers inside publish() of the previous chunk)
              if (issuer) prefetch_res(nt, nc, (cc + 1) & 1);
              mbar_wait(&res_bar[cc & 1], (cc >> 1) & 1);
              box_load_row(res_box + (cc & 1) * kBox, r_in, r);
            } else {
              mbar_wait(&res_bar[eg], cc & 1);
              box_load_row(res_box + eg * kBox, r_in, r);
            }
          }
          uint32_t acc[64];
          tmem_ld64(taddr + c * 64, acc);
          float v[64];
          linear_round64(acc, G.bias ? G.bias + tc.n0 + c * 64 : nullptr,
                         add_r ? add_r + c * 64 : nullptr, v);
          lora_add(c * 64, v);
          if constexpr (EPI == EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 64; i += 2) {
              const float2 g2 = gelu_tanh2(make_float2(v[i], v[i + 1]));
              v[i] = g2.x;
              v[i + 1] = g2.y;
            }
          }
          if constexpr (EPI == EPI_GATE_RES) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float gt[8];
              ld8(G.gate + tc.n0 + c * 64 + q * 8, gt);
#pragma unroll
              for (int i = 0; i < 8; i += 2) {
                float g0 = __fmul_rn(gt[i], v[q * 8 + i]), g1 = __fmul_rn(gt[i + 1], v[q * 8 + i + 1]);
                bf16_round2(g0, g1);
                v[q * 8 + i] = __fadd_rn(r[q * 8 + i], g0);
                v[q * 8 + i + 1] = __fadd_rn(r[q * 8 + i + 1], g1);
              }
            }
          }
          publish(v, tc.n0 + c * 64, nt, nc);
          ++cc;
        }
      }
      // release this accumulator stage: every epilogue warp of both CTAs arrives on the leader
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], 0);
      if (etr && eti < 16) p.trace[eti * 8 + 5] = clock64();
      if (++as == kAccStages) { as = 0; aphase ^= 1; }
    }
    if (issuer) tma_store_wait_all<0>();  // all boxes written to global before the kernel ends
  }

  tc_fence_before();
  cluster_sync_all();  // no CTA may exit (or free TMEM) while its pair can still touch it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2cta<kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------ host
template <int EPI, int BN, bool LORA = false, int EG = 1>
static int set_attr2() {
  static bool done = false;
  if (!done) {
    RF_CHECK_CUDA(cudaFuncSetAttribute(gemm2_kernel<EPI, BN, LORA, EG>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg2<BN>::kSmem));
    done = true;
  }
  return 0;
}
int gemm2_init() {
  return (set_attr2<EPI_BIAS, 256>() | set_attr2<EPI_GELU, 256>() | set_attr2<EPI_GATE_RES, 256>() |
          set_attr2<EPI_QKV, 256>() | set_attr2<EPI_BIAS, 128>() | set_attr2<EPI_GATE_RES, 128>() |
          set_attr2<EPI_GELU, 128, true>() | set_attr2<EPI_GATE_RES, 128, true>() |
          set_attr2<EPI_QKV, 128, true>() | set_attr2<EPI_GELU, 256, true>() | set_attr2<EPI_GATE_RES, 256, true>() |
          set_attr2<EPI_QKV, 256, true>())
             ? -2
             : 0;
}

long long* dbg_get_gemm_trace();
// Epilogue warpgroups of the 256-wide kernels.  The two-group form (kEG = 2) is a measured dead end and is only
// instantiated in DEV builds (make DEV=1; RF_GEMM_EPI_GROUPS=2 / RF_LORA_EPI_GROUPS=2, read at graph-capture time):
// the epilogues of the double-buffered kernels (6-11 k cycles; QKV 25 k — tools/trace_gemm.py) already hide under
// the 30 k-cycle mainloop of even the K = 3072 tiles, and a second group only adds register spills and issue
// pressure — entry A 64.3 vs 63.3 ms, entry B 91.5 vs 89.6 ms; on the single-stage LoRA kernels alone: no gain
// either (tools/step_ab.py, profiles/r02_ab_runs.md).
#ifdef RF_DEV_HOOKS
static int epi_groups(bool lora) {
  const char* e = getenv(lora ? "RF_LORA_EPI_GROUPS" : "RF_GEMM_EPI_GROUPS");
  return (e && e[0] == '2') ? 2 : 1;
}
#endif
template <int EPI, int BN, bool LORA = false, int EG = 1>
static int launch2(const Gemm2Params& p_in, int pairs, double rows, cudaStream_t stream) {
  Gemm2Params p = p_in;
  p.trace = dbg_get_gemm_trace();
#ifdef RF_DEV_HOOKS
  {
    static const int skip = getenv("RF_DBG_GEMM_SKIP") ? atoi(getenv("RF_DBG_GEMM_SKIP")) : 0;
    p.dbg_skip = skip;
  }
#endif
  if (int rc = set_attr2<EPI, BN, LORA, EG>()) return rc;
  static const char* kNames[4] = {"gemm_bias", "gemm_gelu", "gemm_gate_res", "gemm_qkv_rms_rope"};
  const char* name = p.g[0].conv_w ? (EPI == EPI_GATE_RES ? "conv_res" : "conv_bias") : kNames[EPI];
  ProfScope prof(name, 2.0 * rows * p.N * p.K,
                 2.0 * (rows * p.K / (p.g[0].conv_w ? p.g[0].conv_taps : 1) +
                        static_cast<double>(p.ngroups) * p.N * p.K + rows * p.N),
                 stream);
  RF_CHECK_CUDA(launch_pdl(gemm2_kernel<EPI, BN, LORA, EG>, dim3(2 * pairs), dim3(kThreads2 + 128 * (EG - 1)), Cfg2<BN>::kSmem, stream, p));
  count_launch();
  return 0;
}

bool gemm2_eligible(int epi, int N, int K, int ngroups, const GemmGroupArgs* groups) {
  if (N % 256 != 0 || K % kBK != 0 || ngroups > kMaxGroups2) return false;
  if (epi == EPI_QKV && N % 384 != 0) return false;
  for (int g = 0; g < ngroups; ++g) {
    if (groups[g].M < 128) return false;  // tiny problems stay on the single-CTA kernel
    if ((groups[g].ldo * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(groups[g].out) & 15)) return false;
    if (epi == EPI_GATE_RES &&
        ((groups[g].ldr * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(groups[g].res) & 15)))
      return false;
  }
  return true;
}

static thread_local int g_reserve_pairs = 0;
int gemm2_reserve_pairs(int pairs) {
  const int prev = g_reserve_pairs;
  g_reserve_pairs = pairs < 0 ? 0 : pairs;
  return prev;
}
// persistent grid: one CTA pair per TPC, minus the reserved ones when the tiles still take the same number of waves
static int gemm2_grid_pairs(int tiles) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  }
  int pairs = sms / 2;
  const int fewer = pairs - g_reserve_pairs;
  if (g_reserve_pairs > 0 && fewer > 0 && (tiles + fewer - 1) / fewer == (tiles + pairs - 1) / pairs) pairs = fewer;
  if (pairs > tiles) pairs = tiles;
  return pairs;
}

static int gemm2_dispatch(int epi, Gemm2Params& p, int tiles, double rows, cudaStream_t stream) {
  p.total_tiles = tiles;
  const int pairs = gemm2_grid_pairs(tiles);
  if (p.bn == 128) {
    if (epi == EPI_BIAS) return launch2<EPI_BIAS, 128>(p, pairs, rows, stream);
    if (epi == EPI_GATE_RES) return launch2<EPI_GATE_RES, 128>(p, pairs, rows, stream);
    set_error("gemm2: 128-wide tiles support the bias and residual epilogues only");
    return -1;
  }
#ifdef RF_DEV_HOOKS
  if (epi_groups(false) == 2) {
    switch (epi) {
      case EPI_BIAS: return launch2<EPI_BIAS, 256, false, 2>(p, pairs, rows, stream);
      case EPI_GELU: return launch2<EPI_GELU, 256, false, 2>(p, pairs, rows, stream);
      case EPI_GATE_RES: return launch2<EPI_GATE_RES, 256, false, 2>(p, pairs, rows, stream);
      case EPI_QKV: return launch2<EPI_QKV, 256, false, 2>(p, pairs, rows, stream);
      default: break;
    }
  }
#endif
  switch (epi) {
    case EPI_BIAS: return launch2<EPI_BIAS, 256>(p, pairs, rows, stream);
    case EPI_GELU: return launch2<EPI_GELU, 256>(p, pairs, rows, stream);
    case EPI_GATE_RES: return launch2<EPI_GATE_RES, 256>(p, pairs, rows, stream);
    case EPI_QKV: return launch2<EPI_QKV, 256>(p, pairs, rows, stream);
    default: break;
  }
  set_error("gemm2_launch: unknown epilogue");
  return -1;
}

int gemm2_launch(int epi, int N, int K, int ngroups, const GemmGroupArgs* groups,
                 cudaStream_t stream) {
  Gemm2Params p;
  memset(&p, 0, sizeof(p));
  p.ngroups = ngroups;
  p.N = N;
  p.K = K;
  p.bn = 256;
  p.n_tiles = N / p.bn;
  p.num_kb = K / kBK;
  // raster: tiles run m-major inside bands of `band` n-tiles.  Up to 12 n-tiles (N <= 3072) one band
  // covers N, so every A row block is read once.  Wider outputs (QKV 36, MLP-in 48 n-tiles) use
  // narrow bands: the W slice of a band (4 x 256 x K) then stays in L2 for all 18 m-tiles and A
  // (28 MB) survives between bands; measured DRAM reads per launch, band 12 -> 4 (ncu, B200):
  // MLP-in 176 -> 112 MB (operands: 104 MB), N = 9216 133 -> 88 MB (85 MB).  Same run time.
  p.band = p.n_tiles <= 12 ? p.n_tiles : 4;
#ifdef RF_DEV_HOOKS
  {
    static const int band_env = getenv("RF_GEMM_BAND") ? atoi(getenv("RF_GEMM_BAND")) : 0;  // raster experiments
    if (band_env > 0 && band_env < p.n_tiles) p.band = band_env;
  }
#endif
  int tiles = 0;
  double rows = 0;
  for (int g = 0; g < ngroups; ++g) {
    const GemmGroupArgs& a = groups[g];
    Gemm2Group& d = p.g[g];
    int rc = make_tmap_2d(&d.tmA, a.A, a.M, K, a.lda, kRows);
    if (rc) return rc;
    rc = make_tmap_2d(&d.tmB, a.W, N, K, K, p.bn / 2);
    if (rc) return rc;
    rc = make_tmap_2d(&d.tmOut, a.out, a.M, N, a.ldo, kRows);
    if (rc) return rc;
    if (epi == EPI_GATE_RES) {
      if (a.res == nullptr || a.gate == nullptr) {
        set_error("gemm2_launch: EPI_GATE_RES needs res and gate");
        return -1;
      }
      rc = make_tmap_2d(&d.tmRes, a.res, a.M, N, a.ldr, kRows);
      if (rc) return rc;
    }
    if (epi == EPI_QKV && (!a.rope_cos || !a.rope_sin || !a.norm_q || !a.norm_k)) {
      set_error("gemm2_launch: EPI_QKV needs rope tables and norm weights");
      return -1;
    }
    d.bias = a.bias; d.addend = a.addend; d.gate = a.gate;
    d.rope_cos = a.rope_cos; d.rope_sin = a.rope_sin; d.norm_q = a.norm_q; d.norm_k = a.norm_k;
    d.M = a.M; d.ldadd = a.ldadd;
    d.m_pairs = (a.M + 2 * kRows - 1) / (2 * kRows);
    d.bn = p.bn; d.n_tiles = p.n_tiles; d.band = p.band; d.lora = 0;
    d.tile_begin = tiles;
    tiles += d.m_pairs * p.n_tiles;
    rows += a.M;
  }
  return gemm2_dispatch(epi, p, tiles, rows, stream);
}

// One token stream (the condition tokens) with peft LoRA fused: out = epi(bf16(bf16(A W^T + b) + bf16(T B^T))),
// T = bf16(A lora_A^T) computed beforehand ([M, ldT]; columns [64 s, 64 s + 64) belong to output section
// s = n / sec_cols when sec_cols > 0 — the stacked q|k|v projection), lora_B [N, 64] (rank zero-padded to 64).
bool gemm2_lora_eligible(int epi, int N, int K, const GemmGroupArgs& a) {
  if (epi != EPI_GELU && epi != EPI_GATE_RES && epi != EPI_QKV) return false;
  if (N % 128 != 0 || K % kBK != 0 || a.M < 128 || a.addend != nullptr) return false;
  if (epi == EPI_QKV && N % 384 != 0) return false;
  if ((a.ldo * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.out) & 15)) return false;
  if (epi == EPI_GATE_RES && ((a.ldr * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(a.res) & 15))) return false;
  return true;
}
int gemm2_lora_launch(int epi, int N, int K, const GemmGroupArgs& a, const bf16* T, int ldT,
                      const bf16* loraB, int sec_cols, cudaStream_t stream) {
  Gemm2Params p;
  memset(&p, 0, sizeof(p));
  p.ngroups = 1;
  p.N = N;
  p.K = K;
  // 256-wide tiles (single accumulator stage) unless N only divides by 128 or RF_LORA_BN=128 asks for the
  // double-buffered 128-wide form (A/B runs)
  const char* bnenv = getenv("RF_LORA_BN");
  p.bn = (N % 256 == 0 && !(bnenv && atoi(bnenv) == 128)) ? 256 : 128;
  p.n_tiles = N / p.bn;
  p.num_kb = K / kBK;
  p.band = p.bn == 256 ? (p.n_tiles <= 12 ? p.n_tiles : 4) : (p.n_tiles <= 24 ? p.n_tiles : 8);
  Gemm2Group& d = p.g[0];
  int rc = make_tmap_2d(&d.tmA, a.A, a.M, K, a.lda, kRows);
  if (rc) return rc;
  rc = make_tmap_2d(&d.tmB, a.W, N, K, K, p.bn / 2);
  if (rc) return rc;
  rc = make_tmap_2d(&d.tmOut, a.out, a.M, N, a.ldo, kRows);
  if (rc) return rc;
  const int t_cols = sec_cols > 0 ? (N / sec_cols) * 64 : 64;
  rc = make_tmap_2d(&d.tmT, T, a.M, t_cols, ldT, kRows);
  if (rc) return rc;
  rc = make_tmap_2d(&d.tmLB, loraB, N, 64, 64, p.bn / 2);
  if (rc) return rc;
  if (epi == EPI_GATE_RES) {
    if (a.res == nullptr || a.gate == nullptr) {
      set_error("gemm2_lora_launch: EPI_GATE_RES needs res and gate");
      return -1;
    }
    rc = make_tmap_2d(&d.tmRes, a.res, a.M, N, a.ldr, kRows);
    if (rc) return rc;
  }
  if (epi == EPI_QKV && (!a.rope_cos || !a.rope_sin || !a.norm_q || !a.norm_k)) {
    set_error("gemm2_lora_launch: EPI_QKV needs rope tables and norm weights");
    return -1;
  }
  d.bias = a.bias; d.addend = nullptr; d.gate = a.gate;
  d.rope_cos = a.rope_cos; d.rope_sin = a.rope_sin; d.norm_q = a.norm_q; d.norm_k = a.norm_k;
  d.M = a.M; d.ldadd = 0;
  d.lora_sec_cols = sec_cols;
  d.m_pairs = (a.M + 2 * kRows - 1) / (2 * kRows);
  d.bn = p.bn; d.n_tiles = p.n_tiles; d.band = p.band; d.lora = 1;
  d.tile_begin = 0;
  const int tiles = d.m_pairs * p.n_tiles;
  p.total_tiles = tiles;
  const int pairs = gemm2_grid_pairs(tiles);
#ifdef RF_DEV_HOOKS
  if (p.bn == 256 && epi_groups(true) == 2) {
    switch (epi) {
      case EPI_GELU: return launch2<EPI_GELU, 256, true, 2>(p, pairs, a.M, stream);
      case EPI_GATE_RES: return launch2<EPI_GATE_RES, 256, true, 2>(p, pairs, a.M, stream);
      case EPI_QKV: return launch2<EPI_QKV, 256, true, 2>(p, pairs, a.M, stream);
      default: break;
    }
  }
#endif
  if (p.bn == 256) {
    switch (epi) {
      case EPI_GELU: return launch2<EPI_GELU, 256, true>(p, pairs, a.M, stream);
      case EPI_GATE_RES: return launch2<EPI_GATE_RES, 256, true>(p, pairs, a.M, stream);
      case EPI_QKV: return launch2<EPI_QKV, 256, true>(p, pairs, a.M, stream);
      default: break;
    }
  } else {
    switch (epi) {
      case EPI_GELU: return launch2<EPI_GELU, 128, true>(p, pairs, a.M, stream);
      case EPI_GATE_RES: return launch2<EPI_GATE_RES, 128, true>(p, pairs, a.M, stream);
      case EPI_QKV: return launch2<EPI_QKV, 128, true>(p, pairs, a.M, stream);
      default: break;
    }
  }
  set_error("gemm2_lora_launch: unsupported epilogue");
  return -1;
}

// 3x3 (taps = 9) or 1x1 (taps = 1) convolution, stride 1 (zero padding 1) or stride 2 (diffusers
// Downsample2D: pad right/bottom by one, no other padding), NHWC bf16 images with a one-pixel zero
// ring: in [(Hin+2)(Win+2), Cin], out [(H+2)(W+2), Cout] (interior written, ring untouched), weights
// [Cout, taps * Cin] (tap-major K), optional residual res (same geometry as out) added after the bias
// rounding (ResnetBlock2D: x + conv2(...)).  H, W are the OUTPUT size; Hin = stride * H.
// Needs: Cin % 64 == 0, Cout % 128 == 0, W a power of two >= 8 (or a multiple of 128), H % (128/bx) == 0.
int conv_launch(const bf16* in, const bf16* weight, const bf16* bias, bf16* out, const bf16* res,
                const bf16* ones, int H, int W, int Cin, int Cout, int taps, int stride,
                cudaStream_t stream) {
  const int bx = W >= 128 ? 128 : W;
  const int by = 128 / bx;
  if (Cin % 64 != 0 || Cout % 128 != 0 || (taps != 9 && taps != 1) || (stride != 1 && stride != 2) ||
      W % bx != 0 || 128 % bx != 0 || H % by != 0 || (static_cast<long long>(H) * W) % 256 != 0) {
    set_error("conv_launch: unsupported geometry (Cin % 64, Cout % 128, W = 2^k or multiple of 128, "
              "H divisible by 128 / min(W, 128))");
    return -1;
  }
  Gemm2Params p;
  memset(&p, 0, sizeof(p));
  p.ngroups = 1;
  p.N = Cout;
  p.K = taps * Cin;
  p.bn = (Cout % 256 == 0) ? 256 : 128;
  p.n_tiles = Cout / p.bn;
  p.num_kb = p.K / kBK;
  p.band = p.n_tiles;
  Gemm2Group& d = p.g[0];
  const int Hin = stride * H, Win = stride * W;
  int rc = make_tmap_3d(&d.tmA, in, Cin, Win + 2, Hin + 2, bx, by, stride);
  if (rc) return rc;
  rc = make_tmap_2d(&d.tmB, weight, Cout, p.K, p.K, p.bn / 2);
  if (rc) return rc;
  rc = make_tmap_3d(&d.tmOut, out, Cout, W + 2, H + 2, bx, by, 1);
  if (rc) return rc;
  if (res) {
    rc = make_tmap_3d(&d.tmRes, res, Cout, W + 2, H + 2, bx, by, 1);
    if (rc) return rc;
  }
  d.bias = bias;
  d.gate = ones;  // residual epilogue = bf16(res + bf16(1 * y))
  d.M = H * W;
  d.m_pairs = (d.M + 2 * kRows - 1) / (2 * kRows);
  d.bn = p.bn; d.n_tiles = p.n_tiles; d.band = p.band; d.lora = 0;
  d.conv_w = W;
  d.conv_taps = taps;
  d.conv_cin_blocks = Cin / kBK;
  d.conv_bx = bx;
  d.conv_by = by;
  d.conv_stride = stride;
  return gemm2_dispatch(res ? EPI_GATE_RES : EPI_BIAS, p, d.m_pairs * p.n_tiles,
                        static_cast<double>(d.M), stream);
}

}  // namespace rf
