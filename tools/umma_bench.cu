// Micro-benchmark: sustained tcgen05.mma issue/execute rate of one CTA per SM for 128 x N x 16 bf16 MMAs whose
// operands sit in shared memory (SS mode), optionally with other warps hammering shared memory with stores.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/umma_bench tools/umma_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../clip-dplm_b200/csrc/ptx.cuh"

template <int N>
__global__ void __launch_bounds__(192, 1) k(int iters, int writers, long long* out, int flags) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t barmem;
  const uint32_t base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  const uint32_t bar = ptx::smem_u32(&barmem);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(ptx::smem_u32(&tmem_ptr), 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
  const uint32_t tm = tmem_ptr;
  if (warp == 1 && writers >= 0) {
    long long t0 = clock64();
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
      for (int it = 0; it < iters; ++it) {
        const uint32_t a = base + (it & 3) * 16384;            // 4 A boxes [128][64]
        const uint32_t b = base + 65536 + (it & 1) * 32768;    // B boxes [N][64]
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::mma_f16(tm, ptx::smem_desc_k_sw128(a + kk * 32), ptx::smem_desc_k_sw128(b + kk * 32), idesc, 1);
      }
      ptx::mma_commit(bar);
    }
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 5 && writers < 0 && writers > -50) {
    // mode B: replicate the kernel's issue-loop structure: warp-uniform loop, a (always-passing) barrier wait,
    // tcgen05 fence, elected lane issues 4*boxes MMAs and commits to a (never waited-on) barrier each iteration
    __shared__ __align__(8) uint64_t dummy[8];
    const int boxes = -writers;
    if (lane == 0) for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&dummy[i]), 1);
    ptx::fence_mbar_init();
    __syncwarp();
    long long t0 = clock64();
    const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
    int stage = 0; uint32_t phase = 0;
    for (int it = 0; it < iters; it += boxes) {
      if (!(flags & 1)) ptx::mbar_wait(ptx::smem_u32(&dummy[4 + (stage & 3)]), 1);     // fresh barrier: parity-1 wait passes immediately
      if (!(flags & 2)) ptx::tc_fence_after();
      if (ptx::elect_one()) {
        for (int sub = 0; sub < boxes; ++sub) {
          const uint32_t a = base + ((stage * boxes + sub) & 3) * 16384;
          const uint32_t b = base + 65536 + (sub & 1) * 32768;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            ptx::mma_f16(tm, ptx::smem_desc_k_sw128(a + kk * 32), ptx::smem_desc_k_sw128(b + kk * 32), idesc, 1);
        }
        if (!(flags & 4)) ptx::mma_commit(ptx::smem_u32(&dummy[stage & 3]));
      }
      if (!(flags & 8)) __syncwarp();
      if (++stage == 4) { stage = 0; phase ^= 1u; }
    }
    if (ptx::elect_one()) ptx::mma_commit(bar);
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 4 && writers == -100) {
    // mode C: ONE elected lane runs the whole loop including the barrier waits
    __shared__ __align__(8) uint64_t dummy2[8];
    if (lane == 0) for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&dummy2[i]), 1);
    ptx::fence_mbar_init();
    __syncwarp();
    long long t0 = clock64();
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::idesc_bf16_f32(128, N);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < iters; ++it) {
        if (!(flags & 1)) ptx::mbar_wait(ptx::smem_u32(&dummy2[4 + (stage & 3)]), 1);
        if (!(flags & 2)) ptx::tc_fence_after();
        const uint32_t a = base + (stage & 3) * 16384;
        const uint32_t b = base + 65536 + (stage & 1) * 32768;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          ptx::mma_f16(tm, ptx::smem_desc_k_sw128(a + kk * 32), ptx::smem_desc_k_sw128(b + kk * 32), idesc, 1);
        if (!(flags & 4)) ptx::mma_commit(ptx::smem_u32(&dummy2[stage & 3]));
        if (++stage == 4) { stage = 0; phase ^= 1u; }
      }
      ptx::mma_commit(bar);
    }
    __syncwarp();
    ptx::mbar_wait(bar, 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp >= 2 && warp < 2 + writers) {
    // store stream into a disjoint 32 KiB region: ~1 STS.128 per warp per few clocks
    uint4 v = make_uint4(1, 2, 3, 4);
    uint8_t* dst = smem + 131072 + (warp - 2) * 4096 + lane * 16;
    for (int it = 0; it < iters * 8; ++it) {
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(ptx::smem_u32(dst + (it & 7) * 512)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  ptx::tc_fence_before(); __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tm, 512);
}

template <int N> void run(int iters, int writers, int flags = 0) {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rep = 0; rep < 2; ++rep) k<N><<<148, 192, 200 * 1024>>>(iters, writers, d, flags);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("N=%3d writers=%d flags=%d: %s  %.1f clk per MMA (ideal %d)  -> %.0f%% of tensor peak\n", N, writers, flags, cudaGetErrorString(e),
         avg / (4.0 * iters), N / 2, 100.0 * (N / 2) / (avg / (4.0 * iters)));
  cudaFree(d);
}

int main() {
  const int iters = 20000;
  for (int f : {0, 1, 4, 5}) run<64>(iters, -100, f);
  run<128>(iters, -100, 0);
  return 0;
}
