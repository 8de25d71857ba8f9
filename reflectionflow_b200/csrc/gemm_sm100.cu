// gemm_sm100.cu — persistent, warp-specialised bf16 GEMM for sm_100a.
//
//   out[M, N] = epilogue( A[M, K] @ W[N, K]^T )        (torch.nn.Linear weight layout)
//
// Replaces every nn.Linear call site on the reference's DiT hot path
// (train_flux/flux/block.py:27-29,46-48,81-83,148-154,252-259,296,320-328;
//  train_flux/flux/transformer.py:92-93,115,244) together with the element-wise work the
// reference runs after it (bias, GELU-tanh, gate*y + residual, per-head RMSNorm + RoPE).
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      TMA producer   : cp.async.bulk.tensor 128B-swizzled A/B k-blocks -> smem ring
//   warp 1      MMA issuer     : one thread issues tcgen05.mma (128 x BN x 16), fp32 acc in TMEM
//   warp 2      TMEM allocator
//   warps 4..7  epilogue       : tcgen05.ld accumulator -> registers -> fused math -> HBM
// TMEM holds two accumulator stages so the epilogue of tile i overlaps the mainloop of i+1.
#include <cuda.h>

#include "rf_internal.h"
#include "rf_ptx.cuh"

namespace rf {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int kGemmThreads = 256;
static constexpr int kMaxGroups = 3;

struct alignas(64) GemmGroupDev {
  CUtensorMap tmA;
  CUtensorMap tmB;
  const bf16* bias;
  bf16* out;
  const bf16* addend;
  const bf16* res;
  const bf16* gate;
  const float* rope_cos;
  const float* rope_sin;
  const bf16* norm_q;
  const bf16* norm_k;
  int M, ldo, ldadd, ldr;
  int m_tiles, tile_begin;
};

struct alignas(64) GemmParamsDev {
  GemmGroupDev g[kMaxGroups];
  int ngroups, N, K, n_tiles, total_tiles, num_kb;
  long long* trace;  // dev-only per-tile timeline of CTA 0 (rf_dbg_set_gemm_trace)
};

template <int BN>
struct GemmCfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (power of two >= 32)
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct TileCoord {
  int g, m0, n0;
};
__device__ __forceinline__ TileCoord decode_tile(const GemmParamsDev& p, int t) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < kMaxGroups; ++i)
    if (i < p.ngroups && t >= p.g[i].tile_begin) g = i;
  int local = t - p.g[g].tile_begin;
  int m = local / p.n_tiles;
  int n = local - m * p.n_tiles;
  return TileCoord{g, m * BM, n};
}

// load 32 consecutive bf16 (64 B, 16B-aligned) as fp32
__device__ __forceinline__ void load32_bf16(const bf16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u = __ldg(q + i);
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z),
           d = unpack_bf16x2(u.w);
    v[i * 8 + 0] = a.x; v[i * 8 + 1] = a.y; v[i * 8 + 2] = b.x; v[i * 8 + 3] = b.y;
    v[i * 8 + 4] = c.x; v[i * 8 + 5] = c.y; v[i * 8 + 6] = d.x; v[i * 8 + 7] = d.y;
  }
}
__device__ __forceinline__ void store32_bf16(bf16* p, const float (&v)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 u;
    u.x = pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
    u.y = pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
    u.z = pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
    u.w = pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
    q[i] = u;
  }
}

// acc chunk (32 fp32 from TMEM) -> v = bf16(acc + bias) [-> v = bf16(v + addend)]
// These are exactly the rounding points of nn.Linear (+ peft LoRA add) in bf16.
__device__ __forceinline__ void linear_round(const uint32_t (&acc)[32], const bf16* bias_ptr,
                                             const bf16* addend_ptr, float (&v)[32]) {
  if (bias_ptr != nullptr) {
    float b[32];
    load32_bf16(bias_ptr, b);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(__uint_as_float(acc[i]) + b[i]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(__uint_as_float(acc[i]));
  }
  if (addend_ptr != nullptr) {
    float a[32];
    load32_bf16(addend_ptr, a);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = bf16_round(v[i] + a[i]);
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ GemmParamsDev p) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tfull_bar = empty_bar + Cfg::kStages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int g = 0; g < p.ngroups; ++g) {
      tma_prefetch_desc(&p.g[g].tmA);
      tma_prefetch_desc(&p.g[g].tmB);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();  // PDL: the prologue above overlapped the previous kernel's tail

  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    int ti = 0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
      TileCoord tc = decode_tile(p, t);
      const GemmGroupDev& G = p.g[tc.g];
      long long stall = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const long long w0 = tr ? clock64() : 0;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (tr) stall += clock64() - w0;
        uint8_t* sa = smem + stage * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kStageBytesA;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
        tma_load_2d(sa, &G.tmA, &full_bar[stage], kb * BK, tc.m0);
        tma_load_2d(sb, &G.tmB, &full_bar[stage], kb * BK, tc.n0 * BN);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if (tr && ti < 16) p.trace[ti * 8 + 6] = stall;
    }
  } else if (warp == 1 && lane == 0) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    int ti = 0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
      if (tr && ti < 16) p.trace[ti * 8 + 0] = clock64();
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      if (tr && ti < 16) p.trace[ti * 8 + 1] = clock64();
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      long long stall = 0;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        const long long w0 = tr ? clock64() : 0;
        mbar_wait(&full_bar[stage], phase);
        if (tr) stall += clock64() - w0;
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
        const uint32_t sb = sa + Cfg::kStageBytesA;
        const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // +32 B per 16-element k-step inside the 128 B swizzle row (encoded >> 4)
          mma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        tc_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      if (tr && ti < 16) { p.trace[ti * 8 + 2] = stall; p.trace[ti * 8 + 3] = clock64(); }
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int ew = warp & 3;  // TMEM lane quadrant this warp may access
    int as = 0;
    uint32_t aphase = 0;
    int ti = 0;
    const bool tr = p.trace != nullptr && blockIdx.x == 0 && warp == 4 && lane == 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++ti) {
      TileCoord tc = decode_tile(p, t);
      const GemmGroupDev& G = p.g[tc.g];
      const int row = tc.m0 + ew * 32 + lane;
      const bool row_ok = row < G.M;
      const int ncol0 = tc.n0 * BN;
      mbar_wait(&tfull_bar[as], aphase);
      if (tr && ti < 16) p.trace[ti * 8 + 4] = clock64();
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;

      if constexpr (EPI == EPI_QKV) {
        // Column space is [q | k | v] x heads x 128.  A head never straddles a tile (BN % 128 == 0).
        const int inner = p.N / 3;
        const float* cosr = G.rope_cos + static_cast<size_t>(row_ok ? row : 0) * 64;
        const float* sinr = G.rope_sin + static_cast<size_t>(row_ok ? row : 0) * 64;
#pragma unroll 1
        for (int hc = 0; hc < BN / 128; ++hc) {
          const int col_h = ncol0 + hc * 128;
          const int section = col_h / inner;  // 0 q, 1 k, 2 v
          const bf16* bias_h = G.bias ? G.bias + col_h : nullptr;
          const bf16* add_h =
              (G.addend && row_ok) ? G.addend + static_cast<size_t>(row) * G.ldadd + col_h : nullptr;
          bf16* out_h = G.out + static_cast<size_t>(row_ok ? row : 0) * G.ldo + col_h;
          if (section == 2) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t acc[32];
              tmem_ld_32x32(taddr + hc * 128 + c * 32, acc);
              tmem_ld_wait();
              float v[32];
              linear_round(acc, bias_h ? bias_h + c * 32 : nullptr,
                           add_h ? add_h + c * 32 : nullptr, v);
              if (row_ok) store32_bf16(out_h + c * 32, v);
            }
          } else {
            const bf16* nw = (section == 0) ? G.norm_q : G.norm_k;
            float ss = 0.f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t acc[32];
              tmem_ld_32x32(taddr + hc * 128 + c * 32, acc);
              tmem_ld_wait();
              float v[32];
              linear_round(acc, bias_h ? bias_h + c * 32 : nullptr,
                           add_h ? add_h + c * 32 : nullptr, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) ss = __fmaf_rn(v[i], v[i], ss);
            }
            const float var = __fdiv_rn(ss, 128.0f);
            const float rinv = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(var, 1e-6f)));
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
              uint32_t acc[32];
              tmem_ld_32x32(taddr + hc * 128 + c * 32, acc);
              tmem_ld_wait();
              float v[32], w[32];
              linear_round(acc, bias_h ? bias_h + c * 32 : nullptr,
                           add_h ? add_h + c * 32 : nullptr, v);
              load32_bf16(nw + c * 32, w);
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float y = bf16_round(__fmul_rn(v[i], rinv));  // RMSNorm result cast to bf16
                v[i] = bf16_round(__fmul_rn(y, w[i]));        // * weight (bf16 op)
              }
              // interleaved-pair RoPE in fp32, one rounding to bf16
              const float4* c4 = reinterpret_cast<const float4*>(cosr + c * 16);
              const float4* s4 = reinterpret_cast<const float4*>(sinr + c * 16);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                float4 cc = __ldg(c4 + q4), sn = __ldg(s4 + q4);
                float cs[4] = {cc.x, cc.y, cc.z, cc.w};
                float sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int i0 = (q4 * 4 + j) * 2;
                  const float x0 = v[i0], x1 = v[i0 + 1];
                  v[i0] = __fadd_rn(__fmul_rn(x0, cs[j]), __fmul_rn(-x1, sv[j]));
                  v[i0 + 1] = __fadd_rn(__fmul_rn(x1, cs[j]), __fmul_rn(x0, sv[j]));
                }
              }
              if (row_ok) store32_bf16(out_h + c * 32, v);
            }
          }
        }
      } else {
        const bf16* add_r =
            (G.addend && row_ok) ? G.addend + static_cast<size_t>(row) * G.ldadd + ncol0 : nullptr;
        bf16* out_r = G.out + static_cast<size_t>(row_ok ? row : 0) * G.ldo + ncol0;
        const bf16* res_r = nullptr;
        if constexpr (EPI == EPI_GATE_RES)
          res_r = G.res + static_cast<size_t>(row_ok ? row : 0) * G.ldr + ncol0;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          float r[32];
          if constexpr (EPI == EPI_GATE_RES) {
            if (row_ok) load32_bf16(res_r + c * 32, r);
          }
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          float v[32];
          linear_round(acc, G.bias ? G.bias + ncol0 + c * 32 : nullptr,
                       add_r ? add_r + c * 32 : nullptr, v);
          if constexpr (EPI == EPI_GELU) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_tanh(v[i]);
          }
          if constexpr (EPI == EPI_GATE_RES) {
            float gt[32];
            load32_bf16(G.gate + ncol0 + c * 32, gt);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float gy = bf16_round(__fmul_rn(gt[i], v[i]));
              v[i] = __fadd_rn(r[i], gy);
            }
          }
          if (row_ok) store32_bf16(out_r + c * 32, v);
        }
      }
      // release this accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (tr && ti < 16) p.trace[ti * 8 + 5] = clock64();
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------ host
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e =
      cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
  fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  return fn;
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                 uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return -3;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || ((ld * 2) & 15) != 0) {
    set_error("make_tmap_2d: base/pitch must be 16-byte aligned");
    return -1;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed, CUresult=" + std::to_string(static_cast<int>(r)));
    return -3;
  }
  return 0;
}

static long long* g_gemm_trace = nullptr;
void dbg_set_gemm_trace(long long* p) { g_gemm_trace = p; }
long long* dbg_get_gemm_trace() { return g_gemm_trace; }
int make_tmap_3d(CUtensorMap* out, const void* base, uint64_t C, uint64_t Wp, uint64_t Hp,
                 uint32_t box_x, uint32_t box_y, uint32_t stride) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return -3;
  }
  cuuint64_t gdim[3] = {C, Wp, Hp};
  cuuint64_t gstride[2] = {C * 2, Wp * C * 2};
  cuuint32_t box[3] = {64, box_x * stride, box_y * stride};
  cuuint32_t estr[3] = {1, stride, stride};
  if (box[1] > 256 || box[2] > 256) {
    set_error("make_tmap_3d: box dimension exceeds 256");
    return -1;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed, CUresult=" + std::to_string(static_cast<int>(r)));
    return -3;
  }
  return 0;
}

static int g_num_sms = 0;
static int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <int BN, int EPI>
static int set_attr() {
  static bool attr_set = false;
  if (!attr_set) {
    RF_CHECK_CUDA(cudaFuncSetAttribute(gemm_kernel<BN, EPI>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       GemmCfg<BN>::kSmemBytes));
    attr_set = true;
  }
  return 0;
}

int gemm2_init();
int gemm_init() {
  int rc = gemm2_init();
  rc |= set_attr<256, EPI_BIAS>(); rc |= set_attr<256, EPI_GELU>();
  rc |= set_attr<256, EPI_GATE_RES>(); rc |= set_attr<256, EPI_QKV>();
  rc |= set_attr<128, EPI_BIAS>(); rc |= set_attr<128, EPI_GELU>();
  rc |= set_attr<128, EPI_GATE_RES>(); rc |= set_attr<128, EPI_QKV>();
  rc |= set_attr<64, EPI_BIAS>(); rc |= set_attr<64, EPI_GELU>(); rc |= set_attr<64, EPI_GATE_RES>();
  return rc ? -2 : 0;
}

template <int BN, int EPI>
static int launch_cfg(const GemmParamsDev& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  if (int rc = set_attr<BN, EPI>()) return rc;
  int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  double rows = 0;
  for (int g = 0; g < p.ngroups; ++g) rows += p.g[g].M;
  static const char* kNames[4] = {"gemm_bias", "gemm_gelu", "gemm_gate_res", "gemm_qkv_rms_rope"};
  // algorithmic FLOPs: 2 M N K; algorithmic bytes: A + W + out once
  ProfScope prof(kNames[EPI], 2.0 * rows * p.N * p.K,
                 2.0 * (rows * p.K + static_cast<double>(p.ngroups) * p.N * p.K + rows * p.N), stream);
  RF_CHECK_CUDA(launch_pdl(gemm_kernel<BN, EPI>, dim3(grid), dim3(kGemmThreads), Cfg::kSmemBytes, stream, p));
  count_launch();
  return 0;
}

template <int BN>
static int launch_bn(int epi, const GemmParamsDev& p, cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch_cfg<BN, EPI_BIAS>(p, stream);
    case EPI_GELU: return launch_cfg<BN, EPI_GELU>(p, stream);
    case EPI_GATE_RES: return launch_cfg<BN, EPI_GATE_RES>(p, stream);
    default: break;
  }
  set_error("gemm_launch: epilogue not available for this tile width");
  return -1;
}

bool gemm2_eligible(int epi, int N, int K, int ngroups, const GemmGroupArgs* groups);
int gemm2_launch(int epi, int N, int K, int ngroups, const GemmGroupArgs* groups, cudaStream_t stream);
int gemm2_init();
static bool g_force_v1 = false;
void dbg_force_gemm_v1(bool on) { g_force_v1 = on; }

int gemm_launch(int epi, int N, int K, int ngroups, const GemmGroupArgs* groups,
                cudaStream_t stream) {
  if (ngroups >= 1 && !g_force_v1 && gemm2_eligible(epi, N, K, ngroups, groups))
    return gemm2_launch(epi, N, K, ngroups, groups, stream);  // CTA-pair kernel for the big shapes
  if (ngroups < 1 || ngroups > kMaxGroups) {
    set_error("gemm_launch: ngroups must be 1..3");
    return -1;
  }
  if (K % BK != 0 || K <= 0) {
    set_error("gemm_launch: K must be a positive multiple of 64");
    return -1;
  }
  int bn = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : (N % 64 == 0 ? 64 : 0));
  if (bn == 0) {
    set_error("gemm_launch: N must be a multiple of 64");
    return -1;
  }
  if (epi == EPI_QKV && (bn < 128 || N % 384 != 0)) {
    set_error("gemm_launch: EPI_QKV needs N = 3 * heads * 128");
    return -1;
  }
  GemmParamsDev p;
  memset(&p, 0, sizeof(p));
  p.ngroups = ngroups;
  p.N = N;
  p.K = K;
  p.n_tiles = N / bn;
  p.num_kb = K / BK;
  int tiles = 0;
  for (int g = 0; g < ngroups; ++g) {
    const GemmGroupArgs& a = groups[g];
    GemmGroupDev& d = p.g[g];
    if (a.M <= 0) {
      set_error("gemm_launch: empty group");
      return -1;
    }
    int rc = make_tmap_2d(&d.tmA, a.A, a.M, K, a.lda, BM);
    if (rc) return rc;
    rc = make_tmap_2d(&d.tmB, a.W, N, K, K, bn);
    if (rc) return rc;
    d.bias = a.bias; d.out = a.out; d.addend = a.addend; d.res = a.res; d.gate = a.gate;
    d.rope_cos = a.rope_cos; d.rope_sin = a.rope_sin; d.norm_q = a.norm_q; d.norm_k = a.norm_k;
    d.M = a.M; d.ldo = a.ldo; d.ldadd = a.ldadd; d.ldr = a.ldr;
    d.m_tiles = (a.M + BM - 1) / BM;
    d.tile_begin = tiles;
    tiles += d.m_tiles * p.n_tiles;
    if (epi == EPI_GATE_RES && (a.res == nullptr || a.gate == nullptr)) {
      set_error("gemm_launch: EPI_GATE_RES needs res and gate");
      return -1;
    }
    if (epi == EPI_QKV && (!a.rope_cos || !a.rope_sin || !a.norm_q || !a.norm_k)) {
      set_error("gemm_launch: EPI_QKV needs rope tables and norm weights");
      return -1;
    }
  }
  p.total_tiles = tiles;
  p.trace = g_gemm_trace;
  if (epi == EPI_QKV) {
    return bn == 256 ? launch_cfg<256, EPI_QKV>(p, stream) : launch_cfg<128, EPI_QKV>(p, stream);
  }
  switch (bn) {
    case 256: return launch_bn<256>(epi, p, stream);
    case 128: return launch_bn<128>(epi, p, stream);
    default: return launch_bn<64>(epi, p, stream);
  }
}

}  // namespace rf
