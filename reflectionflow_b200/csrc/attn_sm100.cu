// attn_sm100.cu — non-causal joint attention for sm_100a, head_dim 128.
//
//   O = softmax(Q K^T / sqrt(128)) V        over the joint [txt | img | cond] token sequence
//
// Replaces F.scaled_dot_product_attention at train_flux/flux/block.py:123-125 and the
// torch.cat / transpose traffic around it (block.py:31-33,70-72,102-104,126-128): Q, K, V are
// read by TMA straight out of the token-major [n_tok, heads*128] buffers the QKV GEMM epilogue
// wrote, and O is written token-major so it is the A operand of the out-projection GEMM.
//
// One CTA = one 128-row query tile of one head; two CTAs are co-resident per SM so one CTA's
// softmax overlaps the other's tensor-core work.
//   warps 0..3  softmax / correction / epilogue: thread r owns score row r (tcgen05.ld 32x32b)
//   warp 4      TMA producer (Q once, then K_j, V_j per 128-row KV tile)
//   warp 5      MMA issuer: S = Q K_j^T (SS, both K-major), O += P V_j (A = P from TMEM,
//               B = V MN-major from smem)
// TMEM columns: [0,128) S (fp32), aliased by P (bf16, 64 columns) once a row has been read;
//               [128,256) O (fp32).
// Online softmax keeps a per-row running max that is only refreshed when it grows by more
// than 2^8 (lazy rescale), so the O correction pass is rare.
#include <cuda.h>

#include "rf_internal.h"
#include "rf_ptx.cuh"

namespace rf {

static constexpr int kAttnThreads = 192;
static constexpr int kTile = 128;
static constexpr int kHalfBytes = kTile * 128;   // one [128 x 64] bf16 box = 16 KB
static constexpr int kTileBytes = 2 * kHalfBytes;  // [128 x 128] bf16 = 32 KB
static constexpr int kAttnSmem = 3 * kTileBytes + 1024 + 128;

struct alignas(64) AttnParamsDev {
  CUtensorMap tmQ, tmK, tmV;
  bf16* out;
  int ldo, n_tok, heads, batch, q_tiles, kv_tiles;
  int n_main, cond_mode;
  float scale_log2;  // log2(e) / sqrt(128)
  float bias_log2;   // log2(e) * cond_bias
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 2)
attn_kernel(const __grid_constant__ AttnParamsDev p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTileBytes;
  uint8_t* sV = smem + 2 * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work decomposition: q tile fastest so that all q tiles of one head run together (K/V in L2)
  int bid = blockIdx.x;
  const int qt = bid % p.q_tiles;
  bid /= p.q_tiles;
  const int head = bid % p.heads;
  const int b = bid / p.heads;
  const int row_base = b * p.n_tok;  // first row of this batch element in the buffers
  const int q0 = qt * kTile;
  const int col0 = head * 128;
  const bool q_is_cond = (p.cond_mode != 0) && (q0 >= p.n_main);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
  }
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc<256>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;        // S / P
  const uint32_t tO = tmem_base + 128;  // O

  if (warp == 4) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_2d(sQ, &p.tmQ, q_full, col0, row_base + q0);
      tma_load_2d(sQ + kHalfBytes, &p.tmQ, q_full, col0 + 64, row_base + q0);
      int it = 0;
      for (int j = 0; j < p.kv_tiles; ++j) {
        const bool cross = (p.cond_mode != 0) && ((j * kTile >= p.n_main) != q_is_cond);
        if (p.cond_mode == 2 && cross) continue;
        const uint32_t par = (it & 1) ^ 1;
        mbar_wait(k_empty, par);
        mbar_arrive_expect_tx(k_full, kTileBytes);
        tma_load_2d(sK, &p.tmK, k_full, col0, row_base + j * kTile);
        tma_load_2d(sK + kHalfBytes, &p.tmK, k_full, col0 + 64, row_base + j * kTile);
        mbar_wait(v_empty, par);
        mbar_arrive_expect_tx(v_full, kTileBytes);
        tma_load_2d(sV, &p.tmV, v_full, col0, row_base + j * kTile);
        tma_load_2d(sV + kHalfBytes, &p.tmV, v_full, col0 + 64, row_base + j * kTile);
        ++it;
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);  // A, B K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // B (V) MN-major
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
      mbar_wait(q_full, 0);
      int it = 0;
      for (int j = 0; j < p.kv_tiles; ++j) {
        const bool cross = (p.cond_mode != 0) && ((j * kTile >= p.n_main) != q_is_cond);
        if (p.cond_mode == 2 && cross) continue;
        const uint32_t par = it & 1;
        mbar_wait(k_full, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t off = (k >> 2) * kHalfBytes + (k & 3) * 32;
          mma_ss(tS, make_smem_desc(aQ + off, 16, 1024, 2), make_smem_desc(aK + off, 16, 1024, 2),
                 idesc_qk, k != 0 ? 1u : 0u);
        }
        tc_commit(k_empty);
        tc_commit(s_full);
        mbar_wait(p_full, par);
        mbar_wait(v_full, par);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          // 16 kv rows per step: P columns advance by 8 (16 bf16), V by 16 rows x 128 B
          mma_ts(tO, tS + k * 8, make_smem_desc(aV + k * 2048, kHalfBytes, 1024, 2), idesc_pv,
                 (it | k) != 0 ? 1u : 0u);
        }
        tc_commit(v_empty);
        ++it;
      }
      tc_commit(o_full);
    }
  } else {
    // ===================== softmax / correction / epilogue =====================
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const int row = q0 + warp * 32 + lane;
    float m_used = -INFINITY;  // running max the exponentials are referenced to (raw score units)
    float l_sum = 0.f;
    int it = 0;
    for (int j = 0; j < p.kv_tiles; ++j) {
      const bool cross = (p.cond_mode != 0) && ((j * kTile >= p.n_main) != q_is_cond);
      if (p.cond_mode == 2 && cross) continue;
      const float bias = (p.cond_mode == 1 && cross) ? p.bias_log2 : 0.f;
      const int kv_valid = p.n_tok - j * kTile;  // < 128 only on a ragged last tile
      const uint32_t par = it & 1;
      mbar_wait(s_full, par);
      tc_fence_after();
      // ---- pass 1: row max (in log2-scaled units)
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t s[32];
        tmem_ld_32x32(tS + lane_off + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(s[i]);
          if (c * 32 + i >= kv_valid) x = -INFINITY;
          m_tile = fmaxf(m_tile, x);
        }
      }
      const float m_tile_l2 = fmaf(m_tile, p.scale_log2, bias);
      const float m_new = fmaxf(m_used, m_tile_l2);
      const bool need = (it == 0) || (m_new - m_used > 8.0f);
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = (it == 0) ? 0.f : ex2_approx(m_used - m_new);
        m_used = m_new;
        l_sum *= alpha;
        if (it != 0) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(tO + lane_off + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st_32x32(tO + lane_off + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      // ---- pass 2: P = exp2(s * scale + bias - m_used) -> bf16 into TMEM (aliasing S)
      const float neg_m = bias - m_used;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t s[32];
        tmem_ld_32x32(tS + lane_off + c * 32, s);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x0 = __uint_as_float(s[2 * i]), x1 = __uint_as_float(s[2 * i + 1]);
          float p0 = ex2_approx(fmaf(x0, p.scale_log2, neg_m));
          float p1 = ex2_approx(fmaf(x1, p.scale_log2, neg_m));
          if (c * 32 + 2 * i >= kv_valid) p0 = 0.f;
          if (c * 32 + 2 * i + 1 >= kv_valid) p1 = 0.f;
          l_sum += p0 + p1;
          pk[i] = pack_bf16x2(p0, p1);
        }
        tmem_st_32x16(tS + lane_off + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_full);
      ++it;
    }
    // ---- epilogue: O / l -> bf16 -> HBM (token-major)
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = (it > 0) ? __fdiv_rn(1.0f, l_sum) : 0.f;
    bf16* orow = p.out + static_cast<size_t>(row_base + row) * p.ldo + col0;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      if (it > 0) {
        tmem_ld_32x32(tO + lane_off + c * 32, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0;
      }
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
    tmem_dealloc<256>(tmem_base);
  }
}

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
  p.q_tiles = (a.n_tok + kTile - 1) / kTile;
  p.kv_tiles = p.q_tiles;
  p.n_main = a.cond_mode ? a.n_main : a.n_tok;
  p.cond_mode = a.cond_mode;
  const float kLog2e = 1.4426950408889634f;
  p.scale_log2 = kLog2e * 0.08838834764831845f;  // 1/sqrt(128)
  p.bias_log2 = kLog2e * a.cond_bias;
  if (int rc = attention_init()) return rc;
  const int grid = p.q_tiles * a.heads * a.batch;
  // algorithmic FLOPs: QK^T and PV only (4 * n^2 * 128 per head); bytes: q,k,v read + o written
  const double n = a.n_tok;
  ProfScope prof("attention", 4.0 * n * n * 128.0 * a.heads * a.batch,
                 2.0 * 4.0 * n * 128.0 * a.heads * a.batch, stream);
  attn_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(p);
  RF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace rf
