// lora_down_sm100.cu — the LoRA down-projection of the condition tokens as a split-K skinny GEMM.
//
//   T[M, NT] = bf16( X[M, K] @ A[NT, K]^T ),   NT = 64 (one target) | 192 (stacked q|k|v) | 256
//
// peft computes lora_A(x) as its own bf16 Linear (lora_controller.py:5-42 toggles the scaling of
// unfused LoraLayers), so T is rounded to bf16 before the up-projection; the up-projection itself is
// fused into the condition stream's main GEMM (gemm2cta_sm100.cu, kLora).  The product is tiny
// (M = 1024 condition tokens, NT <= 256) but K runs up to 15360: one CTA per 128-row tile would chain
// 240 k-blocks on 8 SMs.  Here every (m tile, K split) pair is a CTA — tcgen05.mma 128 x NT x 16 from a
// TMA ring, fp32 partial to a workspace — and once every split of an m tile has arrived each of its CTAs
// sums a share of the rows over the partials in split order (deterministic) and writes the bf16 rows.
// One launch, no atomics on the data.
#include <cuda.h>

#include "rf_internal.h"
#include "rf_ptx.cuh"

namespace rf {

static constexpr int kLdThreads = 256;
template <int NT> struct LdStages { static constexpr int v = (NT <= 64) ? 8 : 4; };  // 24 KB stages: 8 deep covers the L2 / HBM latency of a lone CTA stream
static constexpr int kLdMaxSplits = 16;

struct alignas(64) LoraDownParams {
  CUtensorMap tmX, tmA;
  float* ws;            // [splits][m_tiles * 128][NT] fp32 partials
  unsigned* counters;   // [m_tiles] arrivals, [m_tiles] finished shares: zero between launches (re-armed in-kernel)
  unsigned* done;
  bf16* T;
  int ldT, M, num_kb, splits, kb_per, m_tiles;
  int n_blocks;  // > 1: full-K CTAs per 64-column block of T (no split-K, no workspace): NT == 64, splits == 1
};

template <int NT>
struct LdCfg {
  static constexpr int kStageA = 128 * 64 * 2;
  static constexpr int kStageB = NT * 64 * 2;
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kTmemCols = NT <= 64 ? 64 : 256;
  static constexpr int kStages = LdStages<NT>::v;
  static constexpr int kSmem = kStages * kStage + 1024 + 256;
};

template <int NT>
__global__ void __launch_bounds__(kLdThreads, 1) lora_down_kernel(const __grid_constant__ LoraDownParams p) {
  constexpr int kStage = LdCfg<NT>::kStage;
  constexpr int kLdStages = LdCfg<NT>::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kLdStages * kStage);
  uint64_t* empty_bar = full_bar + kLdStages;
  uint64_t* tfull_bar = empty_bar + kLdStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tile = blockIdx.x / p.splits;
  const int sp = blockIdx.x - tile * p.splits;
  const int mt = tile / p.n_blocks;
  const int nb = tile - mt * p.n_blocks;
  const int kb0 = sp * p.kb_per;
  const int kb1 = min(p.num_kb, kb0 + p.kb_per);
  const int nkb = kb1 - kb0;  // >= 1 by construction of the launch

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kLdStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<LdCfg<NT>::kTmemCols>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* sa = smem + stage * kStage;
      mbar_arrive_expect_tx(&full_bar[stage], kStage);
      tma_load_2d(sa, &p.tmX, &full_bar[stage], kb * 64, mt * 128);
      tma_load_2d(sa + LdCfg<NT>::kStageA, &p.tmA, &full_bar[stage], kb * 64, nb * NT);
      if (++stage == kLdStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, NT, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < nkb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + stage * kStage);
      const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
      const uint64_t bdesc = make_smem_desc(sa + LdCfg<NT>::kStageA, 16, 1024, 2);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma_ss(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
      tc_commit(&empty_bar[stage]);
      if (++stage == kLdStages) { stage = 0; phase ^= 1; }
    }
    tc_commit(tfull_bar);
  } else if (warp >= 4) {
    const int ew = warp & 3;
    const int r_in = ew * 32 + lane;
    const int row = mt * 128 + r_in;
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const size_t rows_pad = static_cast<size_t>(p.m_tiles) * 128;
    float* wrow = p.ws + (static_cast<size_t>(sp) * rows_pad + row) * NT;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    if (p.splits == 1) {  // the whole K ran in this CTA: round and store the bf16 rows directly
      if (row < p.M) {
        bf16* trow = p.T + static_cast<size_t>(row) * p.ldT + nb * NT;
#pragma unroll 1
        for (int c = 0; c < NT / 32; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32(taddr + c * 32, acc);
          tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(trow + c * 32);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = pack_bf16x2(__uint_as_float(acc[8 * q]), __uint_as_float(acc[8 * q + 1]));
            o.y = pack_bf16x2(__uint_as_float(acc[8 * q + 2]), __uint_as_float(acc[8 * q + 3]));
            o.z = pack_bf16x2(__uint_as_float(acc[8 * q + 4]), __uint_as_float(acc[8 * q + 5]));
            o.w = pack_bf16x2(__uint_as_float(acc[8 * q + 6]), __uint_as_float(acc[8 * q + 7]));
            dst[q] = o;
          }
        }
      }
    } else {
#pragma unroll 1
    for (int c = 0; c < NT / 32; ++c) {
      uint32_t acc[32];
      tmem_ld_32x32(taddr + c * 32, acc);
      tmem_ld_wait();
      float4* dst = reinterpret_cast<float4*>(wrow + c * 32);
#pragma unroll
      for (int q = 0; q < 8; ++q)
        dst[q] = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                             __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3]));
    }
    // ---- every split CTA of this m tile reduces its share of the rows once all partials are written.  All
    // (m tile, split) CTAs are co-resident (grid <= one wave), so waiting on the arrival counter cannot
    // deadlock; the partials are summed in split order (deterministic).
    __threadfence();
    named_bar_sync(2, 128);
    if (warp == 4 && lane == 0) {
      atomicAdd(&p.counters[mt], 1u);
      const volatile unsigned* cnt = p.counters + mt;
      while (*cnt < static_cast<unsigned>(p.splits)) __nanosleep(32);
    }
    named_bar_sync(2, 128);
    __threadfence();
    {
      const int rps = (128 + p.splits - 1) / p.splits;  // rows per split CTA
      const int r0 = sp * rps, r1 = min(128, r0 + rps);
      const int items = (r1 - r0) * (NT / 8);           // (row, 8-column chunk) pairs
      for (int it = r_in; it < items; it += 128) {
        const int rr = r0 + it / (NT / 8), c = it % (NT / 8);
        const int grow = mt * 128 + rr;
        if (grow >= p.M) continue;
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = 0.f;
        // 4 splits' loads in flight at a time (the serial version spent one L2 round trip per split)
        for (int s0 = 0; s0 < p.splits; s0 += 4) {
          float4 u[4], v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int s2 = min(s0 + q, p.splits - 1);
            const float4* src = reinterpret_cast<const float4*>(p.ws + (static_cast<size_t>(s2) * rows_pad + grow) * NT + c * 8);
            u[q] = __ldcg(src);
            v[q] = __ldcg(src + 1);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (s0 + q < p.splits) {
              a[0] += u[q].x; a[1] += u[q].y; a[2] += u[q].z; a[3] += u[q].w;
              a[4] += v[q].x; a[5] += v[q].y; a[6] += v[q].z; a[7] += v[q].w;
            }
          }
        }
        uint4 o;
        o.x = pack_bf16x2(a[0], a[1]);
        o.y = pack_bf16x2(a[2], a[3]);
        o.z = pack_bf16x2(a[4], a[5]);
        o.w = pack_bf16x2(a[6], a[7]);
        *reinterpret_cast<uint4*>(p.T + static_cast<size_t>(grow) * p.ldT + c * 8) = o;
      }
    }
    // the last CTA to finish its share re-arms both counters (every CTA has left the wait loop by then)
    named_bar_sync(2, 128);
    if (warp == 4 && lane == 0) {
      const unsigned prev = atomicAdd(&p.done[mt], 1u);
      if (prev == static_cast<unsigned>(p.splits - 1)) {
        p.counters[mt] = 0u;
        p.done[mt] = 0u;
      }
    }
    }  // splits > 1
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<LdCfg<NT>::kTmemCols>(tmem_base);
  }
}

// ---- confined form: a few CTAs (2-CTA clusters, so that a cluster takes one TPC) walk every (m tile, 64-column
// block) item with the full K each.  It is meant to run on a forked stream UNDER the other token streams' GEMM of the
// same layer, which leaves `kLdSideClusters` TPCs free for it (gemm2_reserve_pairs): the 17-30 us a standalone
// lora_down launch takes on an otherwise idle GPU disappear from the step.  No split-K, no workspace.
static constexpr int kLdSideStages = 8;
static constexpr int kLdSideStage = 128 * 64 * 2 + 64 * 64 * 2;  // 24 KB
static constexpr int kLdSideSmem = kLdSideStages * kLdSideStage + 1024 + 256;
__global__ void __launch_bounds__(kLdThreads, 1) lora_down_side_kernel(const __grid_constant__ LoraDownParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kLdSideStages * kLdSideStage);
  uint64_t* empty_bar = full_bar + kLdSideStages;
  uint64_t* tfull_bar = empty_bar + kLdSideStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int items = p.m_tiles * p.n_blocks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmA);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kLdSideStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc<128>(tmem_slot);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x) {
      const int mt = it / p.n_blocks, nb = it - mt * p.n_blocks;
      for (int kb = 0; kb < p.num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kLdSideStage;
        mbar_arrive_expect_tx(&full_bar[stage], kLdSideStage);
        tma_load_2d(sa, &p.tmX, &full_bar[stage], kb * 64, mt * 128);
        tma_load_2d(sa + 128 * 64 * 2, &p.tmA, &full_bar[stage], kb * 64, nb * 64);
        if (++stage == kLdSideStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0) {
    constexpr uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int acc = n & 1;
      mbar_wait(&tempty_bar[acc], ((n >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d = tmem_base + acc * 64;
      for (int i = 0; i < p.num_kb; ++i) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + stage * kLdSideStage);
        const uint64_t adesc = make_smem_desc(sa, 16, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sa + 128 * 64 * 2, 16, 1024, 2);
#pragma unroll
        for (int k = 0; k < 4; ++k) mma_ss(d, adesc + 2 * k, bdesc + 2 * k, idesc, (i | k) != 0 ? 1u : 0u);
        tc_commit(&empty_bar[stage]);
        if (++stage == kLdSideStages) { stage = 0; phase ^= 1; }
      }
      tc_commit(&tfull_bar[acc]);
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    const int r_in = ew * 32 + lane;
    int n = 0;
    for (int it = blockIdx.x; it < items; it += gridDim.x, ++n) {
      const int mt = it / p.n_blocks, nb = it - mt * p.n_blocks;
      const int acc = n & 1;
      const int row = mt * 128 + r_in;
      const uint32_t taddr = tmem_base + acc * 64 + (static_cast<uint32_t>(ew * 32) << 16);
      mbar_wait(&tfull_bar[acc], (n >> 1) & 1);
      tc_fence_after();
      uint32_t lo[32], hi[32];
      tmem_ld_32x32(taddr, lo);
      tmem_ld_32x32(taddr + 32, hi);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (row < p.M) {
        uint4* dst = reinterpret_cast<uint4*>(p.T + static_cast<size_t>(row) * p.ldT + nb * 64);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(lo[8 * q]), __uint_as_float(lo[8 * q + 1]));
          o.y = pack_bf16x2(__uint_as_float(lo[8 * q + 2]), __uint_as_float(lo[8 * q + 3]));
          o.z = pack_bf16x2(__uint_as_float(lo[8 * q + 4]), __uint_as_float(lo[8 * q + 5]));
          o.w = pack_bf16x2(__uint_as_float(lo[8 * q + 6]), __uint_as_float(lo[8 * q + 7]));
          dst[q] = o;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(hi[8 * q]), __uint_as_float(hi[8 * q + 1]));
          o.y = pack_bf16x2(__uint_as_float(hi[8 * q + 2]), __uint_as_float(hi[8 * q + 3]));
          o.z = pack_bf16x2(__uint_as_float(hi[8 * q + 4]), __uint_as_float(hi[8 * q + 5]));
          o.w = pack_bf16x2(__uint_as_float(hi[8 * q + 6]), __uint_as_float(hi[8 * q + 7]));
          dst[4 + q] = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<128>(tmem_base);
  }
}

// T = bf16(X A^T) on `kLdSideClusters` TPCs (see the kernel comment).  Same result as lora_down_launch.
int lora_down_side_launch(const bf16* X, int ldx, int M, int K, const bf16* A, int NT, bf16* T, int ldT,
                          cudaStream_t stream) {
  if (NT % 64 != 0 || NT <= 0 || K % 64 != 0 || K <= 0 || M <= 0 || (ldT * 2) % 16 != 0) {
    set_error("lora_down_side: NT and K must be multiples of 64");
    return -1;
  }
  static bool attr_done = false;
  if (!attr_done) {
    RF_CHECK_CUDA(cudaFuncSetAttribute(lora_down_side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLdSideSmem));
    attr_done = true;
  }
  LoraDownParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.tmX, X, M, K, ldx, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&p.tmA, A, NT, K, K, 64);
  if (rc) return rc;
  p.m_tiles = (M + 127) / 128;
  p.num_kb = K / 64;
  p.n_blocks = NT / 64;
  p.splits = 1;
  p.kb_per = p.num_kb;
  p.T = T;
  p.ldT = ldT;
  p.M = M;
  ProfScope prof("lora_down", 2.0 * M * NT * static_cast<double>(K), 2.0 * (static_cast<double>(M) * K + static_cast<double>(NT) * K), stream);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * kLdSideClusters);
  cfg.blockDim = dim3(kLdThreads);
  cfg.dynamicSmemBytes = kLdSideSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, lora_down_side_kernel, p));
  count_launch();
  return 0;
}

template <int NT>
static int ld_set_attr() {
  static bool done = false;
  if (!done) {
    RF_CHECK_CUDA(cudaFuncSetAttribute(lora_down_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       LdCfg<NT>::kSmem));
    done = true;
  }
  return 0;
}
int lora_down_init() { return (ld_set_attr<64>() | ld_set_attr<192>() | ld_set_attr<256>()) ? -2 : 0; }

// layout: [counters: m_tiles x u32, padded to 256 B | partials].  The counters sit at a FIXED offset so that
// launches with different NT sharing one workspace never overwrite each other's counters with partials.
static size_t ld_counter_bytes(size_t m_tiles) { return (2 * m_tiles * sizeof(unsigned) + 255) / 256 * 256; }
size_t lora_down_workspace_bytes(int M, int NT) {
  const size_t m_tiles = (static_cast<size_t>(M) + 127) / 128;
  return ld_counter_bytes(m_tiles) + kLdMaxSplits * m_tiles * 128 * NT * sizeof(float);
}

// ws: lora_down_workspace_bytes(M, NT) bytes, ZERO-INITIALISED once by the caller (the counters live at
// its end and are re-armed by the kernel).
int lora_down_launch(const bf16* X, int ldx, int M, int K, const bf16* A, int NT, bf16* T, int ldT,
                     void* ws, cudaStream_t stream) {
  if ((NT != 64 && NT != 192 && NT != 256) || K % 64 != 0 || K <= 0 || M <= 0 || (ldT * 2) % 16 != 0) {
    set_error("lora_down: NT must be 64, 192 or 256, K a multiple of 64");
    return -1;
  }
  LoraDownParams p;
  memset(&p, 0, sizeof(p));
  int rc = make_tmap_2d(&p.tmX, X, M, K, ldx, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&p.tmA, A, NT, K, K, NT);
  if (rc) return rc;
  p.m_tiles = (M + 127) / 128;
  p.num_kb = K / 64;
  p.n_blocks = 1;
  // short K, wide T (the stacked q|k|v[|mlp] factors, K = 3072): one full-K CTA per (128 rows, 64 columns) —
  // x is re-read NT/64 times from L2 but nothing goes through the split-K workspace (the partial store, the
  // arrival wait and the L2 round trips of the reduction were ~2/3 of the 19 us such a launch took)
  if (NT >= 128 && p.num_kb <= 64) {
    int rc2 = make_tmap_2d(&p.tmA, A, NT, K, K, 64);
    if (rc2) return rc2;
    p.n_blocks = NT / 64;
    p.splits = 1;
    p.kb_per = p.num_kb;
    p.ws = nullptr;
    p.counters = p.done = nullptr;
    p.T = T;
    p.ldT = ldT;
    p.M = M;
    ProfScope prof("lora_down", 2.0 * M * NT * static_cast<double>(K), 2.0 * (static_cast<double>(M) * K + static_cast<double>(NT) * K), stream);
    if ((rc2 = ld_set_attr<64>())) return rc2;
    RF_CHECK_CUDA(launch_pdl(lora_down_kernel<64>, dim3(p.m_tiles * p.n_blocks), dim3(kLdThreads), LdCfg<64>::kSmem, stream, p));
    count_launch();
    return 0;
  }
  // >= 6 k-blocks per CTA, at most kLdMaxSplits splits, and no more CTAs than fit one wave
  int splits = p.num_kb / 6;
  if (splits < 1) splits = 1;
  if (splits > kLdMaxSplits) splits = kLdMaxSplits;
  while (splits > 1 && splits * p.m_tiles > 148) --splits;
  p.kb_per = (p.num_kb + splits - 1) / splits;
  p.splits = (p.num_kb + p.kb_per - 1) / p.kb_per;  // every split non-empty
  const size_t rows_pad = static_cast<size_t>(p.m_tiles) * 128;
  p.counters = static_cast<unsigned*>(ws);
  p.done = p.counters + p.m_tiles;
  p.ws = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + ld_counter_bytes(p.m_tiles));
  (void)rows_pad;
  p.T = T;
  p.ldT = ldT;
  p.M = M;
  const int grid = p.m_tiles * p.splits;
  ProfScope prof("lora_down", 2.0 * M * NT * static_cast<double>(K), 2.0 * (static_cast<double>(M) * K + static_cast<double>(NT) * K), stream);
  switch (NT) {
    case 64:
      if ((rc = ld_set_attr<64>())) return rc;
      RF_CHECK_CUDA(launch_pdl(lora_down_kernel<64>, dim3(grid), dim3(kLdThreads), LdCfg<64>::kSmem, stream, p));
      break;
    case 192:
      if ((rc = ld_set_attr<192>())) return rc;
      RF_CHECK_CUDA(launch_pdl(lora_down_kernel<192>, dim3(grid), dim3(kLdThreads), LdCfg<192>::kSmem, stream, p));
      break;
    default:
      if ((rc = ld_set_attr<256>())) return rc;
      RF_CHECK_CUDA(launch_pdl(lora_down_kernel<256>, dim3(grid), dim3(kLdThreads), LdCfg<256>::kSmem, stream, p));
      break;
  }
  RF_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return 0;
}

}  // namespace rf
