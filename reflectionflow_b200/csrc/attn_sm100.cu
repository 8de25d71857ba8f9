// attn_sm100.cu — non-causal joint attention for sm_100a, head_dim 128.
//
//   O = softmax(Q K^T / sqrt(128)) V        over the joint [txt | img | cond] token sequence
//
// Replaces F.scaled_dot_product_attention at train_flux/flux/block.py:123-125 and the
// torch.cat / transpose traffic around it (block.py:31-33,70-72,102-104,126-128): Q, K, V are
// read by TMA straight out of the token-major [n_tok, heads*128] buffers the QKV GEMM epilogue
// wrote, and O is written token-major so it is the A operand of the out-projection GEMM.
//
// One CTA = two 128-row query tiles of one head; two softmax warpgroups ping-pong so that the
// tensor pipe works on one tile's QK^T / PV while the other tile's rows are in their softmax
// (see the kernel comment).  K/V tiles stream through a 2-stage TMA ring.
// Online softmax runs in the log2 domain and keeps a per-row reference max that is only
// refreshed when it grows by more than 2^8 (lazy rescale), so the O correction pass is rare.
// A masked (q tile, kv tile) pair of cond_mode 2 writes P = 0 and leaves the row statistics alone.
#include <cuda.h>

#include "rf_internal.h"
#include "rf_ptx.cuh"

namespace rf {

static constexpr int kAttnThreads = 320;  // 8 softmax warps (2 warpgroups) + TMA warp + MMA warp
static constexpr int kTile = 128;
static constexpr int kHalfBytes = kTile * 128;   // one [128 x 64] bf16 box = 16 KB
static constexpr int kTileBytes = 2 * kHalfBytes;  // [128 x 128] bf16 = 32 KB
static constexpr int kKVStages = 2;
static constexpr int kPolyEvery = 4;  // every n-th exp2 pair on the FMA pipe (0 = all on MUFU)
static constexpr int kAttnSmem = (2 + 2 * kKVStages) * kTileBytes + 1024 + 256;

struct alignas(64) AttnParamsDev {
  CUtensorMap tmQ, tmK, tmV;
  bf16* out;
  int ldo, n_tok, heads, batch, q_pairs, kv_tiles;
  int n_main, cond_mode;
  float scale_log2;  // log2(e) / sqrt(128)
  float bias_log2;   // log2(e) * cond_bias
  long long* trace;  // dev-only timeline of CTA 0 (rf_dbg_set_attn_trace); nullptr in production
};

#ifdef RF_DEV_HOOKS
#define RF_TR(id, j)                                                                   \
  do {                                                                                 \
    if (p.trace != nullptr && blockIdx.x == 0 && (j) < 24 && (threadIdx.x & 31) == 0)  \
      p.trace[(j) * 16 + (id)] = clock64();                                            \
  } while (0)
#else
#define RF_TR(id, j) do { } while (0)
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One CTA = TWO 128-row query tiles of one head (FlashAttention-4 style ping-pong):
//   warps 0..3  softmax warpgroup 0 (query tile 0), warps 4..7 softmax warpgroup 1 (tile 1);
//               thread r of a warpgroup owns score row r (tcgen05.ld 32x32b: no shuffles)
//   warp 8      TMA producer: Q0,Q1 once, then K_j / V_j through a 2-stage ring
//   warp 9      MMA issuer, interleaved so the tensor pipe always has the OTHER tile's work while one
//               warpgroup is in its softmax:   QK0_0 QK1_0 | PV0_j QK0_{j+1} PV1_j QK1_{j+1} | ...
// TMEM (512 columns): S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512); P_t (bf16) aliases S_t.
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_kernel(const __grid_constant__ AttnParamsDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // 2 tiles
  uint8_t* sK = smem + 2 * kTileBytes;                 // kKVStages tiles
  uint8_t* sV = sK + kKVStages * kTileBytes;           // kKVStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + kKVStages * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per query tile
  uint64_t* p_lo = bars + 11;     // [2] per query tile: P of kv rows 0..63 written
  uint64_t* p_hi = bars + 13;     // [2] per query tile: P of kv rows 64..127 written
  uint64_t* o_full = bars + 15;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);
  float* s_scale = reinterpret_cast<float*>(bars + 17);  // scale_log2, re-read after the p_lo arrive (see softmax)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // q pair fastest so that all CTAs of one head run together (its K/V stay in L2)
  int bid = blockIdx.x;
  const int qp = bid % p.q_pairs;
  bid /= p.q_pairs;
  const int head = bid % p.heads;
  const int b = bid / p.heads;
  const int row_base = b * p.n_tok;
  const int q0 = qp * 2 * kTile;
  const int col0 = head * 128;
  const int n_kv = p.kv_tiles;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 9 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_lo[i], 4);  // one arrive per softmax warp
      mbar_init(&p_hi[i], 4);
    }
    mbar_init(o_full, 1);
    *s_scale = p.scale_log2;
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc<512>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 8) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      for (int t = 0; t < 2; ++t) {
        tma_load_2d(sQ + t * kTileBytes, &p.tmQ, q_full, col0, row_base + q0 + t * kTile);
        tma_load_2d(sQ + t * kTileBytes + kHalfBytes, &p.tmQ, q_full, col0 + 64,
                    row_base + q0 + t * kTile);
      }
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const uint32_t par = ((j >> 1) & 1) ^ 1;
      uint8_t* kd = sK + st * kTileBytes;
      uint8_t* vd = sV + st * kTileBytes;
      mbar_wait_backoff(&k_empty[st], par, 200);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[st], kTileBytes);
        tma_load_2d(kd, &p.tmK, &k_full[st], col0, row_base + j * kTile);
        tma_load_2d(kd + kHalfBytes, &p.tmK, &k_full[st], col0 + 64, row_base + j * kTile);
      }
      __syncwarp();
      mbar_wait_backoff(&v_empty[st], par, 200);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], kTileBytes);
        tma_load_2d(vd, &p.tmV, &v_full[st], col0, row_base + j * kTile);
        tma_load_2d(vd + kHalfBytes, &p.tmV, &v_full[st], col0 + 64, row_base + j * kTile);
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    // The WHOLE warp runs this code (warp-uniform control flow keeps descriptors in uniform
    // registers: issuing from inside `if (lane == 0)` costs ~15 SASS instructions and an
    // ELECT/BRA loop per tcgen05.mma); one elected lane issues the MMAs and commits.
    constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);  // A, B K-major
    constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // B (V) MN-major
    const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
    auto issue_qk = [&](int t, int st) {
      const uint32_t q = aQ + t * kTileBytes, k = aK + st * kTileBytes;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * kHalfBytes + (kk & 3) * 32;
        mma_ss(tmem_base + t * 128, make_smem_desc(q + off, 16, 1024, 2),
               make_smem_desc(k + off, 16, 1024, 2), idesc_qk, kk != 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int t, int st, bool acc, int k0, int k1) {  // k-steps [k0, k1)
      const uint32_t v = aV + st * kTileBytes;
#pragma unroll
      for (int kk = k0; kk < k1; ++kk) {
        // 16 kv rows per step: P columns advance by 8 (16 bf16), V by 16 rows x 128 B
        mma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + kk * 8,
               make_smem_desc(v + kk * 2048, kHalfBytes, 1024, 2), idesc_pv,
               (acc || kk != 0) ? 1u : 0u);
      }
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
      issue_qk(0, 0);
      tc_commit(&s_full[0]);
      issue_qk(1, 0);
      tc_commit(&s_full[1]);
      tc_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const uint32_t pj = j & 1;
      const bool more = (j + 1 < n_kv);
      const int st2 = (j + 1) & 1;
      const uint32_t ph2 = ((j + 1) >> 1) & 1;
      mbar_wait(&v_full[st], ph);
      mbar_wait(&p_lo[0], pj);
      RF_TR(0, j);
      tc_fence_after();
      if (elect_one()) issue_pv(0, st, j != 0, 0, 6);  // p_lo arrives once the stores of P chunks 0..2 have landed: 6 of the 8 k-steps run under the last chunk's exponentials
      __syncwarp();
      mbar_wait(&p_hi[0], pj);
      if (more) mbar_wait(&k_full[st2], ph2);
      tc_fence_after();
      if (elect_one()) {
        issue_pv(0, st, true, 6, 8);
        if (more) {
          issue_qk(0, st2);
          tc_commit(&s_full[0]);
        }
      }
      __syncwarp();
      RF_TR(1, j);
      mbar_wait(&p_lo[1], pj);
      RF_TR(2, j);
      tc_fence_after();
      if (elect_one()) issue_pv(1, st, j != 0, 0, 6);
      __syncwarp();
      mbar_wait(&p_hi[1], pj);
      tc_fence_after();
      if (elect_one()) {
        issue_pv(1, st, true, 6, 8);
        tc_commit(&v_empty[st]);
        if (more) {
          issue_qk(1, st2);
          tc_commit(&s_full[1]);
          tc_commit(&k_empty[st2]);
        }
      }
      __syncwarp();
      RF_TR(3, j);
    }
    if (elect_one()) tc_commit(o_full);
    __syncwarp();
  } else if (warp < 8) {
    // ===================== softmax / correction / epilogue (two warpgroups) =====================
    const int t = warp >> 2;  // query tile of this warpgroup
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_off;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_off;
    const int qt0 = q0 + t * kTile;
    const int row = qt0 + (warp & 3) * 32 + lane;
    const bool q_is_cond = (p.cond_mode != 0) && (qt0 >= p.n_main);
    float m_used = -INFINITY;  // running reference max, log2-scaled units
    float l_sum = 0.f;
    bool have = false;  // at least one unmasked tile seen
    for (int j = 0; j < n_kv; ++j) {
      const bool cross = (p.cond_mode != 0) && ((j * kTile >= p.n_main) != q_is_cond);
      const float bias = (p.cond_mode == 1 && cross) ? p.bias_log2 : 0.f;
      const int kv_valid = p.n_tok - j * kTile;
      mbar_wait(&s_full[t], j & 1);
      if ((warp & 3) == 0 && lane == 0) RF_TR(4 + 4 * t, j);
      tc_fence_after();
      if (p.cond_mode == 2 && cross) {
        // fully masked tile: P = 0, statistics untouched
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_st_32x16(tS + c * 16, z);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_lo[t]);
      } else {
        // the whole 128-value score row lives in registers: one TMEM read per tile
        uint32_t s[4][32];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(tS + c * 32, s[c]);
        tmem_ld_wait();
        if ((warp & 3) == 0 && lane == 0) RF_TR(5 + 4 * t, j);
        // Keys past n_tok (last, partial tile): TMA zero-fills those K and V rows, so their scores
        // are exactly 0 and their P V contribution vanishes; only the row sum needs a correction
        // (below).  No per-element masking: that cost 254 ISETP/SEL per tile on every tile.
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            mx0 = fmaxf(mx0, __uint_as_float(s[c][i]));
            mx1 = fmaxf(mx1, __uint_as_float(s[c][i + 1]));
            mx2 = fmaxf(mx2, __uint_as_float(s[c][i + 2]));
            mx3 = fmaxf(mx3, __uint_as_float(s[c][i + 3]));
          }
        const float m_tile = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        const float m_new = fmaxf(m_used, fmaf(m_tile, p.scale_log2, bias));
        const bool need = (!have) || (m_new - m_used > 8.0f);
        if (__any_sync(0xffffffffu, need)) {
          const float alpha = have ? ex2_approx(m_used - m_new) : 0.f;
          m_used = m_new;
          l_sum *= alpha;
          if (have) {  // `have` is warp-uniform: it only depends on the tile index
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t o[32];
              tmem_ld_32x32(tO + c * 32, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              tmem_st_32x32(tO + c * 32, o);
            }
          }
        }
        have = true;
        // P = exp2(s * scale + bias - m_used), two elements per instruction (FFMA2 / FADD2).
        // One warp can issue a MUFU.EX2 every 8 cycles at best (tools/mufu_rate.cu), 1024 cycles for
        // the 128 values of a row, and this phase is on the critical path of the tile
        // (QK^T -> softmax -> PV -> next QK^T).  Every kPolyEvery-th pair is therefore evaluated on
        // the FMA pipe instead: round-to-nearest split x = n + f (magic-number add), 2^f by a cubic
        // (rel. error 7.7e-5, 50x below the bf16 rounding of P), 2^n by an integer add into the
        // exponent field.
        const float neg_m = bias - m_used;
        float2 sc2 = make_float2(p.scale_log2, p.scale_log2);
        const float2 nm2 = make_float2(neg_m, neg_m);
        float2 l2 = make_float2(0.f, 0.f);
        auto exp_chunk = [&](int c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 x = __ffma2_rn(
                make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sc2, nm2);
            float2 e;
            if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == (kPolyEvery - 1)) {
              x.x = fmaxf(x.x, -125.0f);
              x.y = fmaxf(x.y, -125.0f);
              const float2 magic = make_float2(12582912.0f, 12582912.0f);
              const float2 t2 = __fadd2_rn(x, magic);                       // low mantissa bits = round(x)
              const float2 n2 = __fadd2_rn(t2, make_float2(-12582912.0f, -12582912.0f));
              const float2 f2 = __ffma2_rn(n2, make_float2(-1.0f, -1.0f), x);  // [-0.5, 0.5]
              float2 q = __ffma2_rn(make_float2(0.05508868396282196f, 0.05508868396282196f), f2,
                                    make_float2(0.24260404706001282f, 0.24260404706001282f));
              q = __ffma2_rn(q, f2, make_float2(0.6932762265205383f, 0.6932762265205383f));
              q = __ffma2_rn(q, f2, make_float2(0.9999289512634277f, 0.9999289512634277f));
              e.x = __uint_as_float(__float_as_uint(q.x) + (__float_as_uint(t2.x) << 23));
              e.y = __uint_as_float(__float_as_uint(q.y) + (__float_as_uint(t2.y) << 23));
            } else {
              e.x = ex2_approx(x.x);
              e.y = ex2_approx(x.y);
            }
            l2 = __fadd2_rn(l2, e);
            pk[i] = pack_bf16x2(e.x, e.y);
          }
          tmem_st_32x16(tS + c * 16, pk);
        };
        if (warp == 0 && lane == 0) RF_TR(12, j);
        exp_chunk(0);
        exp_chunk(1);
        if (warp == 0 && lane == 0) RF_TR(13, j);
        exp_chunk(2);
        // P of kv rows 0..63 was stored two chunks ago: the wait below no longer stalls the MUFU
        // stream, and PV can start on the first half while the last chunk is computed
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_lo[t]);
        {
          // keep the last chunk's exponentials BEHIND the arrive: ptxas otherwise hoists every
          // MUFU.EX2 above the first STTM and the early signal is lost.  A shared-memory load
          // cannot move above the arrive, and the last chunk's FFMA2 depend on its value.
          float sc_b;
          asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(sc_b) : "r"(smem_u32(s_scale)) : "memory");
          sc2 = make_float2(sc_b, sc_b);
        }
        if (warp == 0 && lane == 0) RF_TR(14, j);
        exp_chunk(3);
        if (warp == 0 && lane == 0) RF_TR(15, j);
        float l_tile = l2.x + l2.y;
        if (kv_valid < kTile)  // zero-score padding keys: each added exp2(0 * scale + neg_m) to the sum
          l_tile -= static_cast<float>(kTile - kv_valid) * ex2_approx(neg_m);
        l_sum += l_tile;
      }
      if ((warp & 3) == 0 && lane == 0) RF_TR(6 + 4 * t, j);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_hi[t]);
      if ((warp & 3) == 0 && lane == 0) RF_TR(7 + 4 * t, j);
    }
    // ---- epilogue: O / l -> bf16 -> HBM (token-major)
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = have ? __fdiv_rn(1.0f, l_sum) : 0.f;
    bf16* orow = p.out + static_cast<size_t>(row_base + row) * p.ldo + col0;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(tO + c * 32, o);
      tmem_ld_wait();
      if (row < p.n_tok) {
        uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u;
          u.x = pack_bf16x2(__uint_as_float(o[q * 8 + 0]) * inv_l, __uint_as_float(o[q * 8 + 1]) * inv_l);
          u.y = pack_bf16x2(__uint_as_float(o[q * 8 + 2]) * inv_l, __uint_as_float(o[q * 8 + 3]) * inv_l);
          u.z = pack_bf16x2(__uint_as_float(o[q * 8 + 4]) * inv_l, __uint_as_float(o[q * 8 + 5]) * inv_l);
          u.w = pack_bf16x2(__uint_as_float(o[q * 8 + 6]) * inv_l, __uint_as_float(o[q * 8 + 7]) * inv_l);
          dst[q] = u;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

static long long* g_attn_trace = nullptr;
void dbg_set_attn_trace(long long* p) { g_attn_trace = p; }

int attention_init() {
  static bool attr_set = false;
  if (!attr_set) {
    RF_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       kAttnSmem));
    attr_set = true;
  }
  return 0;
}

int attention_launch(const AttnArgs& a, cudaStream_t stream) {
  if (a.n_tok <= 0 || a.heads <= 0 || a.batch <= 0) {
    set_error("attention_launch: empty problem");
    return -1;
  }
  if (a.cond_mode != 0 && (a.n_main % kTile != 0 || a.n_main <= 0 || a.n_main >= a.n_tok)) {
    set_error("attention_launch: cond split must be a positive multiple of 128 below n_tok");
    return -1;
  }
  if (a.batch > 1 && a.n_tok % kTile != 0) {
    set_error("attention_launch: batched call needs n_tok % 128 == 0");
    return -1;
  }
  AttnParamsDev p;
  memset(&p, 0, sizeof(p));
  const uint64_t rows = static_cast<uint64_t>(a.n_tok) * a.batch;
  const uint64_t cols = static_cast<uint64_t>(a.heads) * 128;
  int rc = make_tmap_2d(&p.tmQ, a.q, rows, cols, a.ld_qkv, kTile);
  if (rc) return rc;
  rc = make_tmap_2d(&p.tmK, a.k, rows, cols, a.ld_qkv, kTile);
  if (rc) return rc;
  rc = make_tmap_2d(&p.tmV, a.v, rows, cols, a.ld_qkv, kTile);
  if (rc) return rc;
  p.out = a.out;
  p.ldo = a.ldo;
  p.n_tok = a.n_tok;
  p.heads = a.heads;
  p.batch = a.batch;
  p.q_pairs = (a.n_tok + 2 * kTile - 1) / (2 * kTile);
  p.kv_tiles = (a.n_tok + kTile - 1) / kTile;
  p.n_main = a.cond_mode ? a.n_main : a.n_tok;
  p.cond_mode = a.cond_mode;
  const float kLog2e = 1.4426950408889634f;
  p.scale_log2 = kLog2e * 0.08838834764831845f;  // 1/sqrt(128)
  p.bias_log2 = kLog2e * a.cond_bias;
  p.trace = g_attn_trace;
  if (int rc = attention_init()) return rc;
  const int grid = p.q_pairs * a.heads * a.batch;
  // algorithmic FLOPs: QK^T and PV only (4 * n^2 * 128 per head); bytes: q,k,v read + o written
  const double n = a.n_tok;
  ProfScope prof("attention", 4.0 * n * n * 128.0 * a.heads * a.batch,
                 2.0 * 4.0 * n * 128.0 * a.heads * a.batch, stream);
  RF_CHECK_CUDA(launch_pdl(attn_kernel, dim3(grid), dim3(kAttnThreads), kAttnSmem, stream, p));
  count_launch();
  return 0;
}

}  // namespace rf
