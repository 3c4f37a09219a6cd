// Microbenchmark: latency of the hand-offs the attention kernels are built from (sm_100a).
//   A: issuer-only chain   MMA x nmma -> commit -> wait                      (MMA issue-to-complete latency)
//   C: issuer-only chain with the A operand in tensor memory (mode 2)
//   B: ping-pong           MMA x nmma -> commit -> 128 threads: wait, tcgen05.ld x32, tcgen05.st x32, arrive -> issuer wait
// Operands are whatever is in shared memory; only the timing matters.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../datatunerx_b200/csrc mma_latency.cu -o mma_latency
#include <cstdio>
#include <cuda_runtime.h>
#include "common.cuh"

using namespace dtx;

__global__ void __launch_bounds__(160, 1) k(int mode, int nmma, int n_cols, int iters, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 128);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t idesc = umma_idesc_bf16(128, n_cols, 0, 0);
  const uint32_t loA = umma_desc_lo(smem_u32(smem), 16), loB = umma_desc_lo(smem_u32(smem + 32768), 16);
  if (warp == 4) {
    const bool leader = elect_one();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (leader) {
        if (mode == 2) {  // A operand from tensor memory (columns 256..), B from shared memory
          for (int i = 0; i < nmma; ++i) umma_bf16_ts(tmem, tmem + 256 + (i & 7) * 8, umma_desc_pack(loB + i * 2), idesc, i > 0);
        } else {
          for (int i = 0; i < nmma; ++i) umma_bf16(tmem, umma_desc_pack(loA + i * 2), umma_desc_pack(loB + i * 2), idesc, i > 0);
        }
        umma_commit(&bars[0]);
      }
      if (mode == 0 || mode == 2) {
        mbar_wait(&bars[0], it & 1);
        tc_fence_after();
      } else {
        mbar_wait(&bars[1], it & 1);
        tc_fence_after();
      }
    }
    const long long t1 = clock64();
    if (leader && blockIdx.x == 0) out[0] = t1 - t0;
  } else if (mode == 1) {
    const uint32_t t_lane = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int it = 0; it < iters; ++it) {
      mbar_wait(&bars[0], it & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(t_lane, v);
      tmem_ld_wait();
#pragma unroll
      for (int e = 0; e < 32; ++e) v[e] += 1;
      tmem_st32(t_lane + 64, v);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[1]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int iters = 2000;
  for (int mode = 0; mode < 3; ++mode)
    for (int n_cols : {64, 128})
      for (int nmma : {1, 8, 16, 64}) {
        k<<<1, 160, 70000>>>(mode, nmma, n_cols, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("mode %d (%s) N=%d nmma=%2d: %.1f cycles / iteration (MMA floor %d)%s\n", mode, mode == 1 ? "ping-pong" : (mode == 2 ? "issuer chain, A in TMEM" : "issuer chain"), n_cols, nmma,
               double(h) / iters, nmma * n_cols / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
      }
  return 0;
}
