// Micro-benchmark: issue rate of tcgen05.mma cta_group::1 kind::f16 (SS mode, K-major SW128 operands already in smem).
// Reports cycles per MMA for N = 64/128/256 with 1 or 2 accumulators, with and without a concurrent TMA-like
// smem write stream (plain st.shared from other warps) to expose shared-memory bandwidth contention.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../pillarnet-lts_b200/csrc/tc_common.cuh"
using namespace pn_tc;

template <int BN>
__device__ __forceinline__ constexpr uint32_t idesc_k() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int BN, int NACC, int STORM>
__global__ void __launch_bounds__(256, 1) k_rate(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* a = base;                 // 2 x 128 rows x 128 B
  uint8_t* b = base + 2 * 16384;     // 256 rows x 128 B
  uint8_t* junk = b + 32768;         // 64 KB scratch for the store storm
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (2 * 16384 + 32768) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc<512>(&tbase);
  }
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tm = tbase;
  if (warp == 0 && lane == 0) {
    const uint64_t ad = make_kmajor_sw128_desc(smem_u32(a)), bd = make_kmajor_sw128_desc(smem_u32(b));
    constexpr uint32_t id = idesc_k<BN>();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int m = 0; m < NACC; ++m)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm + m * BN, ad + m * 1024 + 2 * k, bd + 2 * k, id, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = (t1 - t0);
  } else if (STORM && warp >= 4) {
    // 4 warps hammering st.shared.v4: ~ TMA fill traffic
    uint4* j = (uint4*)junk;
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (int i = 0; i < iters * STORM; ++i) j[(i * 128 + (threadIdx.x - 128)) & 4095] = v;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(tm); }
}

template <int BN, int NACC, int STORM>
void run(int iters) {
  long long* d; cudaMalloc(&d, 8);
  const size_t smem = 2 * 16384 + 32768 + 65536 + 1024;
  cudaFuncSetAttribute(k_rate<BN, NACC, STORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_rate<BN, NACC, STORM><<<148, 256, smem>>>(iters, d);
  k_rate<BN, NACC, STORM><<<148, 256, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double per = (double)h / ((double)iters * NACC * 4);
  printf("N=%3d acc=%d storm=%d : %.1f clk/MMA (floor %d)  [%s]\n", BN, NACC, STORM, per, BN / 2, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  const int it = 2000;
  run<16, 1, 0>(it); run<16, 2, 0>(it); run<32, 1, 0>(it); run<32, 2, 0>(it); run<64, 1, 0>(it); run<64, 2, 0>(it); run<128, 1, 0>(it); run<256, 1, 0>(it);
  run<128, 2, 0>(it); run<256, 2, 0>(it);
  run<128, 2, 4>(it); run<256, 2, 4>(it);
  run<128, 2, 16>(it); run<256, 2, 16>(it);
  return 0;
}
