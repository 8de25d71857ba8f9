// dev microbenchmark: issue rate of tcgen05.mma variants on one SM (cycles per instruction).
#include <cstdio>
#include <cuda_runtime.h>
#include "../reflectionflow_b200/csrc/rf_ptx.cuh"
using namespace rf;

template <int MODE>  // 0: SS 128x128, 1: SS 128x256, 2: TS 128x128 (A from TMEM, B MN-major), 3: SS 128x64
__global__ void __launch_bounds__(128, 1) k(long long* out, int n) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc<512>(&slot); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) ((uint32_t*)smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = slot;
  if (threadIdx.x == 0) {
    constexpr int N = MODE == 1 ? 256 : (MODE == 3 ? 64 : 128);
    constexpr uint32_t idesc = make_idesc_bf16(128, N, 0, MODE == 2 ? 1 : 0);
    uint32_t a = smem_u32(smem), b = smem_u32(smem + 32768);
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint32_t off = ((i & 7) >> 2) * 16384 + (i & 3) * 32;
      if (MODE == 2)
        mma_ts(tm + 256, tm + (i & 7) * 8, make_smem_desc(b + (i & 7) * 2048, 16384, 1024, 2), idesc, i != 0);
      else
        mma_ss(tm, make_smem_desc(a + off, 16, 1024, 2), make_smem_desc(b + off, 16, 1024, 2), idesc, i != 0);
    }
    long long t1 = clock64();
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

template <int MODE> void run(const char* name, int ctas) {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int n = 1024;
  k<MODE><<<ctas, 128, 100 * 1024>>>(d, n); cudaDeviceSynchronize();
  k<MODE><<<ctas, 128, 100 * 1024>>>(d, n);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-28s ctas=%3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma  (%s)\n", name, ctas, h[0] / (double)n,
         h[1] / (double)n, cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  for (int ctas : {1, 148}) {
    run<0>("SS 128x128x16 (QK)", ctas);
    run<1>("SS 128x256x16 (GEMM 1cta)", ctas);
    run<2>("TS 128x128x16 (PV)", ctas);
    run<3>("SS 128x64x16", ctas);
  }
  return 0;
}
