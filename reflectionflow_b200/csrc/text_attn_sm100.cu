// text_attn_sm100.cu — T5 / CLIP self-attention (head_dim 64, S <= 512) on the tensor cores.
//
//   scores = bf16(q k^T) [ = bf16(scores * scale) ] [ = bf16(scores + bias[h, i, j]) ] ; causal: j <= i
//   p      = bf16(softmax_fp32(scores)) ;  out = bf16(p v)
//
// the rounding points of transformers' eager attention in bf16 (T5Attention.forward: no scaling, additive
// relative-position bias; CLIPAttention: scaling, causal mask) — the text encoders diffusers' encode_prompt runs
// for every candidate prompt (train_flux/flux/generate.py:148-161).  Replaces the CUDA-core shared-memory kernel
// (text.cu::small_attn_kernel, 8.8 of the 15.8 ms of a T5-XXL encode in round 1).
//
// One CTA = 128 query rows of one head.  The whole score row block S[128, S_pad <= 512] is ONE set of
// tcgen05.mma (K = 64) into the 512 TMEM columns; thread r owns row r and walks it three times in place
// (round + bias + max | exp + sum | normalise -> bf16 P), P is written packed over the already consumed low
// columns [0, S_pad / 2), and O = P V is a TS-MMA (A = P from TMEM, B = V MN-major straight from the
// token-major buffer) into the freed columns [256, 320).  Probabilities are normalised BEFORE the bf16 rounding
// and before P V, exactly like the eager reference (an online-softmax kernel would round un-normalised P).
#include <cuda.h>

#include "rf_internal.h"
#include "rf_ptx.cuh"

namespace rf {

static constexpr int kTaThreads = 160;  // 4 softmax warps (thread = query row) + 1 TMA / MMA warp
static constexpr int kTaBox = 128 * 64 * 2;  // one [128 rows x 64 cols] bf16 box = 16 KB
static constexpr int kTaSmem = (1 + 4 + 4) * kTaBox + 1024 + 128;

struct alignas(64) TextAttnParams {
  CUtensorMap tmQ, tmK, tmV;
  bf16* out;
  const bf16* bias;  // [heads, S, S] or nullptr
  int ldo, S, nb, use_scale, causal, bias_vec;
  float scale;
};

__global__ void __launch_bounds__(kTaThreads, 1) text_attn_kernel(const __grid_constant__ TextAttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTaBox;
  uint8_t* sV = sK + 4 * kTaBox;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 4 * kTaBox);
  uint64_t* ld_bar = bars + 0;
  uint64_t* qk_bar = bars + 1;
  uint64_t* p_bar = bars + 2;
  uint64_t* o_bar = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int base = blockIdx.z * p.S;
  const int S = p.S, nb = p.nb, S_pad = nb * 128;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tmQ);
    tma_prefetch_desc(&p.tmK);
    tma_prefetch_desc(&p.tmV);
    mbar_init(ld_bar, 1);
    mbar_init(qk_bar, 1);
    mbar_init(p_bar, 4);
    mbar_init(o_bar, 1);
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

  if (warp == 4) {
    if (lane == 0) {
      mbar_arrive_expect_tx(ld_bar, kTaBox * (1 + 2 * nb));
      tma_load_2d(sQ, &p.tmQ, ld_bar, h * 64, base + q0);
      for (int b = 0; b < nb; ++b) {
        tma_load_2d(sK + b * kTaBox, &p.tmK, ld_bar, h * 64, base + b * 128);
        tma_load_2d(sV + b * kTaBox, &p.tmV, ld_bar, h * 64, base + b * 128);
      }
      mbar_wait(ld_bar, 0);
      tc_fence_after();
      constexpr uint32_t idesc_qk = make_idesc_bf16(128, 128, 0, 0);
      const uint64_t adesc = make_smem_desc(smem_u32(sQ), 16, 1024, 2);
      for (int b = 0; b < nb; ++b) {
        const uint64_t bdesc = make_smem_desc(smem_u32(sK + b * kTaBox), 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(tmem_base + b * 128, adesc + 2 * k, bdesc + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
      }
      tc_commit(qk_bar);
      mbar_wait(p_bar, 0);
      tc_fence_after();
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 64, 0, 1);  // B (V) MN-major
      const uint32_t aV = smem_u32(sV);
      for (int kk = 0; kk < S_pad / 16; ++kk)  // 16 keys per step: P advances 8 columns, V 16 rows x 128 B
        mma_ts(tmem_base + 256, tmem_base + kk * 8, make_smem_desc(aV + kk * 2048, kTaBox, 1024, 2), idesc_pv,
               kk != 0 ? 1u : 0u);
      tc_commit(o_bar);
    }
    __syncwarp();
  } else {
    const int r = warp * 32 + lane;     // row of the tile == TMEM lane
    const int row = q0 + r;             // query index inside the sequence
    const int row_c = row < S ? row : S - 1;
    const uint32_t tS = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const bf16* brow = p.bias ? p.bias + (static_cast<size_t>(h) * S + row_c) * S : nullptr;
    const int nch = S_pad / 32;
    mbar_wait(qk_bar, 0);
    tc_fence_after();
    // ---- pass 1: the reference's score roundings, mask, running max; rounded scores go back in place
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tS + c * 32, v);
      float bv[32];
      if (brow != nullptr) {
        if (p.bias_vec && c * 32 + 32 <= S) {
          const uint4* bp = reinterpret_cast<const uint4*>(brow + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 u = __ldg(bp + q);
            const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
            bv[q * 8 + 0] = a.x; bv[q * 8 + 1] = a.y; bv[q * 8 + 2] = b.x; bv[q * 8 + 3] = b.y;
            bv[q * 8 + 4] = cc.x; bv[q * 8 + 5] = cc.y; bv[q * 8 + 6] = d.x; bv[q * 8 + 7] = d.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) bv[i] = (c * 32 + i < S) ? __bfloat162float(brow[c * 32 + i]) : 0.f;
        }
      }
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int j = c * 32 + i;
        float x = bf16_round(__uint_as_float(v[i]));
        if (p.use_scale) x = bf16_round(x * p.scale);
        if (brow != nullptr) x = bf16_round(x + bv[i]);
        if (j >= S || (p.causal && j > row_c)) x = -INFINITY;
        mx = fmaxf(mx, x);
        v[i] = __float_as_uint(x);
      }
      tmem_st_32x32(tS + c * 32, v);
    }
    tmem_st_wait();
    // ---- pass 2: e = exp(x - max) in place, row sum
    float sum = 0.f;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tS + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = expf(__uint_as_float(v[i]) - mx);
        sum += e;
        v[i] = __float_as_uint(e);
      }
      tmem_st_32x32(tS + c * 32, v);
    }
    tmem_st_wait();
    // ---- pass 3: p = bf16(e / sum), packed over the consumed low columns (chunk c -> columns [16 c, 16 c + 16))
    const float inv = 1.0f / sum;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tS + c * 32, v);
      tmem_ld_wait();
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * inv, __uint_as_float(v[2 * i + 1]) * inv);
      tmem_st_32x16(tS + c * 16, pk);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(p_bar);
    // ---- epilogue: O -> bf16 -> token-major output
    mbar_wait(o_bar, 0);
    tc_fence_after();
    uint32_t o[2][32];
    tmem_ld_32x32(tS + 256, o[0]);
    tmem_ld_32x32(tS + 256 + 32, o[1]);
    tmem_ld_wait();
    if (row < S) {
      uint4* dst = reinterpret_cast<uint4*>(p.out + static_cast<size_t>(base + row) * p.ldo + h * 64);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t* s = &o[q >> 2][(q & 3) * 8];
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(s[0]), __uint_as_float(s[1]));
        u.y = pack_bf16x2(__uint_as_float(s[2]), __uint_as_float(s[3]));
        u.z = pack_bf16x2(__uint_as_float(s[4]), __uint_as_float(s[5]));
        u.w = pack_bf16x2(__uint_as_float(s[6]), __uint_as_float(s[7]));
        dst[q] = u;
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

bool text_attn_tc_eligible(int S, int ld, int ldo) {
  return S >= 1 && S <= 512 && (ld * 2) % 16 == 0 && (ldo * 2) % 16 == 0;
}

int text_attn_tc_launch(const bf16* q, const bf16* k, const bf16* v, int ld, bf16* out, int ldo, int B, int S,
                        int heads, const bf16* bias, float scale, int use_scale, int causal, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    RF_CHECK_CUDA(cudaFuncSetAttribute(text_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTaSmem));
    attr = true;
  }
  TextAttnParams p;
  memset(&p, 0, sizeof(p));
  const uint64_t rows = static_cast<uint64_t>(B) * S, cols = static_cast<uint64_t>(heads) * 64;
  int rc = make_tmap_2d(&p.tmQ, q, rows, cols, ld, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&p.tmK, k, rows, cols, ld, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&p.tmV, v, rows, cols, ld, 128);
  if (rc) return rc;
  p.out = out;
  p.bias = bias;
  p.ldo = ldo;
  p.S = S;
  p.nb = (S + 127) / 128;
  p.use_scale = use_scale;
  p.causal = causal;
  p.scale = scale;
  p.bias_vec = (S % 8 == 0 && (reinterpret_cast<uintptr_t>(bias) & 15) == 0) ? 1 : 0;
  dim3 grid((S + 127) / 128, heads, B);
  ProfScope prof("text_attention", 4.0 * B * heads * static_cast<double>(S) * S * 64, 0, stream);
  RF_CHECK_CUDA(launch_pdl(text_attn_kernel, grid, dim3(kTaThreads), kTaSmem, stream, p));
  count_launch();
  return 0;
}

}  // namespace rf
