// Probe of the sm_100a primitives the pair (cta_group::2) kernels rely on.  Run on a B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/umma_probe tools/umma_probe.cu -lcuda
//   gpurun_out/umma_probe
// Part A  ex2 throughput: MUFU.EX2 vs FMA-pipe polynomials (elements / clk / SM).
// Part B  numerics of tcgen05.mma through TMA-written operands: cta_group 1/2, M 128/256, K-major and MN-major A/B
//         (SWIZZLE_128B), including the TMEM accumulator layouts (4x1 for M=256, 2x2 for M=128 over a CTA pair).
// Part C  sustained MMA rate of the same shapes from resident shared memory.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../clip-dplm_b200/csrc/ptx.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

// ------------------------------------------------------------------------------------------------ part A
__device__ __forceinline__ float ex2_mufu(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int DEG>
__device__ __forceinline__ float ex2_poly(float x) {
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float q;
  if (DEG == 5) {
    q = 0.001327646430581808f;
    q = fmaf(q, f, 0.009675540961325169f);
    q = fmaf(q, f, 0.05550713464617729f);
    q = fmaf(q, f, 0.24022120237350464f);
    q = fmaf(q, f, 0.6931469440460205f);
    q = fmaf(q, f, 1.0000001192092896f);
  } else {
    q = 0.05550410866f;
    q = fmaf(q, f, 0.2402265070f);
    q = fmaf(q, f, 0.6931471806f);
    q = fmaf(q, f, 1.0f);
  }
  return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
template <int KIND>
__global__ void __launch_bounds__(512, 1) exp_rate(int iters, float seed, long long* clk, float* sink) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed * (threadIdx.x + 1) * (i + 1) * 1e-4f - 3.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float e = KIND == 0 ? ex2_mufu(v[i]) : (KIND == 1 ? ex2_poly<5>(v[i]) : ex2_poly<3>(v[i]));
      v[i] = e - 3.5f;   // keeps the argument in [-3.5, -2.5]
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += v[i];
  if (s == 12345.f) sink[0] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

// ------------------------------------------------------------------------------------------------ part B / C
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_enc;
static void make_tmap(CUtensorMap* m, const void* base, int inner, int outer, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)inner * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_full(int m, int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit_all(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
  if (CG == 1)
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)m), "r"(bar), "r"(c0), "r"(c1) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)m), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t dst, uint32_t n) {
  if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(n) : "memory");
  else asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(n) : "memory");
  if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_free_cg(uint32_t a, uint32_t n) {
  if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(n) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(n) : "memory");
}

struct GemmArgs {
  int M, N;          // UMMA shape (whole instruction, both CTAs)
  int a_mn, b_mn;    // 1 = MN-major operand
  int iters;         // 0: numerics (TMA load, K = 64, dump TMEM); > 0: rate loop
};

// numerics: one K=64 slab.  rate: iters x 4 MMAs on whatever shared memory holds.
template <int CG>
__global__ void __launch_bounds__(192, 1) gemm_probe(const __grid_constant__ CUtensorMap ta, const __grid_constant__ CUtensorMap tb,
                                                     GemmArgs g, float* out, long long* clk) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_ptr;
  __shared__ __align__(8) uint64_t bars[2];
  const uint32_t base = ptx::smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 1 ? 0u : ptx::cluster_ctarank();
  const uint32_t bar_full = ptx::smem_u32(&bars[0]), bar_done = ptx::smem_u32(&bars[1]);
  const int MA = g.M / CG, NB = g.N / CG;
  const uint32_t a_smem = base, b_smem = base + 65536;
  if (g.iters > 0)
    for (int i = threadIdx.x; i < (160 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(bar_full, 1);
    ptx::mbar_init(bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) tmem_alloc_cg<CG>(ptx::smem_u32(&tmem_ptr), 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  if (CG == 1) __syncthreads(); else { __syncthreads(); ptx::cluster_sync(); }
  ptx::tc_fence_after();
  const uint32_t tm = tmem_ptr;

  if (warp == 1 && g.iters == 0) {
    if (ptx::elect_one()) {
      // leader arms the barrier for the bytes of BOTH CTAs; every CTA issues its own loads
      const uint32_t bytes_cta = (uint32_t)(MA + NB) * 128u;
      if (rank == 0) ptx::mbar_arrive_expect_tx(bar_full, bytes_cta * CG);
      if (!g.a_mn) tma2d<CG>(a_smem, &ta, bar_full, 0, rank * MA);
      else for (int gi = 0; gi < MA / 64; ++gi) tma2d<CG>(a_smem + gi * 8192, &ta, bar_full, rank * MA + gi * 64, 0);
      if (!g.b_mn) tma2d<CG>(b_smem, &tb, bar_full, 0, rank * NB);
      else for (int gi = 0; gi < NB / 64; ++gi) tma2d<CG>(b_smem + gi * 8192, &tb, bar_full, rank * NB + gi * 64, 0);
    }
    __syncwarp();
  }
  if (warp == 2 && rank == 0) {
    long long t0 = 0;
    if (g.iters == 0) ptx::mbar_wait(bar_full, 0);
    ptx::tc_fence_after();
    t0 = clock64();
    if (ptx::elect_one()) {
      const uint32_t idesc = idesc_full(g.M, g.N, g.a_mn, g.b_mn);
      const uint32_t a_step = g.a_mn ? 2048u : 32u, b_step = g.b_mn ? 2048u : 32u;
      const int n_it = g.iters > 0 ? g.iters : 1;
      for (int it = 0; it < n_it; ++it) {
        const uint32_t ao = (g.iters > 0) ? (uint32_t)(it & 1) * 16384u : 0u;
        const uint32_t bo = (g.iters > 0) ? (uint32_t)(it & 1) * 32768u : 0u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t ad = g.a_mn ? desc_sw128(a_smem + ao + kk * a_step, 8192, 1024) : desc_sw128(a_smem + ao + kk * a_step, 16, 1024);
          const uint64_t bd = g.b_mn ? desc_sw128(b_smem + bo + kk * b_step, 8192, 1024) : desc_sw128(b_smem + bo + kk * b_step, 16, 1024);
          mma<CG>(tm, ad, bd, idesc, (g.iters > 0) ? 1u : (kk > 0 ? 1u : 0u));
        }
      }
      commit_all<CG>(bar_done);
    }
    __syncwarp();
    ptx::mbar_wait(bar_done, 0);
    const long long t1 = clock64();
    if (lane == 0 && clk) clk[blockIdx.x / CG] = t1 - t0;
  }
  if (warp >= 2 && g.iters == 0) {
    // warps 2..5 -> lane quarters (warp & 3): dump this CTA's TMEM, N columns
    ptx::mbar_wait(bar_done, 0);
    ptx::tc_fence_after();
    const int q = warp & 3;
    for (int c0 = 0; c0 < g.N; c0 += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tm + ((uint32_t)(q * 32) << 16) + c0, r);
      ptx::tmem_ld_wait();
      for (int x = 0; x < 32; ++x) out[((size_t)rank * 128 + q * 32 + lane) * 256 + c0 + x] = __uint_as_float(r[x]);
    }
  }
  ptx::tc_fence_before();
  if (CG == 1) __syncthreads(); else { __syncthreads(); ptx::cluster_sync(); }
  if (warp == 0) tmem_free_cg<CG>(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int CG>
static void launch_probe(const CUtensorMap& ta, const CUtensorMap& tb, GemmArgs g, float* out, long long* clk, int nclusters) {
  auto kern = gemm_probe<CG>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(nclusters * CG);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaLaunchKernelEx(&cfg, kern, ta, tb, g, out, clk));
}

static bool numerics(int cg, int M, int N, int a_mn, int b_mn) {
  const int K = 64;
  std::vector<float> A((size_t)M * K), B((size_t)N * K);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[(size_t)m * K + k] = bf((float)(((m * 7 + k * 3) % 17) - 8) * 0.125f);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[(size_t)n * K + k] = bf((float)(((n * 5 + k * 11) % 13) - 6) * 0.25f);
  // device storage: K-major = [rows][64]; MN-major = [64 k][rows]
  std::vector<__nv_bfloat16> Ah((size_t)M * K), Bh((size_t)N * K);
  for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) Ah[a_mn ? (size_t)k * M + m : (size_t)m * K + k] = __float2bfloat16(A[(size_t)m * K + k]);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bh[b_mn ? (size_t)k * N + n : (size_t)n * K + k] = __float2bfloat16(B[(size_t)n * K + k]);
  __nv_bfloat16 *dA, *dB; float* dout;
  CK(cudaMalloc(&dA, Ah.size() * 2)); CK(cudaMalloc(&dB, Bh.size() * 2)); CK(cudaMalloc(&dout, 2 * 128 * 256 * 4));
  CK(cudaMemcpy(dA, Ah.data(), Ah.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Bh.data(), Bh.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xff, 2 * 128 * 256 * 4));
  CUtensorMap ta, tb;
  if (!a_mn) make_tmap(&ta, dA, K, M, M / cg); else make_tmap(&ta, dA, M, K, 64);
  if (!b_mn) make_tmap(&tb, dB, K, N, N / cg); else make_tmap(&tb, dB, N, K, 64);
  GemmArgs g{M, N, a_mn, b_mn, 0};
  if (cg == 1) launch_probe<1>(ta, tb, g, dout, nullptr, 1); else launch_probe<2>(ta, tb, g, dout, nullptr, 1);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("numerics cg=%d M=%d N=%d a_mn=%d b_mn=%d: CUDA error %s\n", cg, M, N, a_mn, b_mn, cudaGetErrorString(e)); exit(1); }
  std::vector<float> out(2 * 128 * 256);
  CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
  // try the candidate TMEM layouts and report which matches
  const char* names[3] = {"4x1 (lane = row of this CTA)", "2x2 (lane%64 = row, lane/64 selects the N half)", "none"};
  int match = 2;
  for (int lay = 0; lay < 2 && match == 2; ++lay) {
    double maxerr = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
      int rank, lane, col;
      const int rows_cta = M / cg;
      rank = m / rows_cta;
      const int ml = m % rows_cta;
      if (lay == 0) { lane = ml; col = n; if (rows_cta != 128) { maxerr = 1e9; break; } }
      else { if (rows_cta != 64) { maxerr = 1e9; break; } lane = ml + 64 * (n / (N / 2)); col = n % (N / 2); }
      const double got = out[((size_t)rank * 128 + lane) * 256 + col];
      const double err = fabs(got - ref);
      if (!(err <= maxerr)) maxerr = err;
    }
    if (maxerr < 1e-3) match = lay;
  }
  printf("numerics cg=%d M=%3d N=%3d A %s B %s : layout %s %s\n", cg, M, N, a_mn ? "MN" : "K ", b_mn ? "MN" : "K ", names[match],
         match == 2 ? "  <-- MISMATCH" : "OK");
  if (match == 2) {
    printf("   sample out[rank0][lane 0][0..7]:");
    for (int x = 0; x < 8; ++x) printf(" %g", out[x]);
    printf("\n   sample out[rank0][lane 64][0..7]:");
    for (int x = 0; x < 8; ++x) printf(" %g", out[64 * 256 + x]);
    double ref0 = 0; for (int k = 0; k < K; ++k) ref0 += (double)A[k] * B[k];
    printf("\n   ref D[0][0] = %g\n", ref0);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dout);
  return match != 2;
}

static void rate(int cg, int M, int N, int a_mn, int b_mn) {
  long long* dclk; CK(cudaMalloc(&dclk, 148 * sizeof(long long)));
  __nv_bfloat16* dummy; CK(cudaMalloc(&dummy, 1 << 20)); CK(cudaMemset(dummy, 0, 1 << 20));
  CUtensorMap ta, tb; make_tmap(&ta, dummy, 64, 256, 64); make_tmap(&tb, dummy, 64, 256, 64);
  const int iters = 4000;
  GemmArgs g{M, N, a_mn, b_mn, iters};
  const int ncl = 148 / cg;
  for (int rep = 0; rep < 2; ++rep) { if (cg == 1) launch_probe<1>(ta, tb, g, nullptr, dclk, ncl); else launch_probe<2>(ta, tb, g, nullptr, dclk, ncl); }
  cudaError_t e = cudaDeviceSynchronize();
  std::vector<long long> h(148); CK(cudaMemcpy(h.data(), dclk, ncl * sizeof(long long), cudaMemcpyDeviceToHost));
  double avg = 0; for (int i = 0; i < ncl; ++i) avg += h[i]; avg /= ncl;
  const double per = avg / (4.0 * iters);
  const double ideal = (double)M * N * 16 / (4096.0 * cg);
  printf("rate cg=%d M=%3d N=%3d A %s B %s : %s %.1f clk per MMA (ideal %.0f) -> %.0f%% of tensor peak\n", cg, M, N, a_mn ? "MN" : "K ",
         b_mn ? "MN" : "K ", cudaGetErrorString(e), per, ideal, 100.0 * ideal / per);
  cudaFree(dclk); cudaFree(dummy);
}

int main() {
  {
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    g_enc = reinterpret_cast<EncodeTiledFn>(fn);
  }
  // ---- part A
  {
    long long* dclk; float* sink; CK(cudaMalloc(&dclk, 148 * 8)); CK(cudaMalloc(&sink, 4));
    const int iters = 4096;
    const char* nm[3] = {"MUFU.EX2", "poly deg 5", "poly deg 3"};
    for (int kind = 0; kind < 3; ++kind) {
      for (int rep = 0; rep < 2; ++rep) {
        if (kind == 0) exp_rate<0><<<148, 512>>>(iters, 1.f, dclk, sink);
        if (kind == 1) exp_rate<1><<<148, 512>>>(iters, 1.f, dclk, sink);
        if (kind == 2) exp_rate<2><<<148, 512>>>(iters, 1.f, dclk, sink);
      }
      CK(cudaDeviceSynchronize());
      long long h[148]; CK(cudaMemcpy(h, dclk, sizeof h, cudaMemcpyDeviceToHost));
      double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
      printf("exp rate %-10s: %.2f elements / clk / SM (16 warps, 8 chains each; includes one FADD per element)\n", nm[kind], 512.0 * 8 * iters / avg);
    }
    cudaFree(dclk); cudaFree(sink);
  }
  // ---- part B
  bool ok = true;
  ok &= numerics(1, 128, 128, 0, 0);
  ok &= numerics(1, 128, 256, 0, 1);
  ok &= numerics(1, 128, 128, 1, 0);
  ok &= numerics(1, 128, 256, 1, 1);
  ok &= numerics(2, 256, 256, 0, 0);
  ok &= numerics(2, 128, 256, 0, 0);
  ok &= numerics(2, 128, 256, 0, 1);
  ok &= numerics(2, 128, 128, 0, 1);
  ok &= numerics(2, 256, 256, 0, 1);
  ok &= numerics(2, 256, 128, 1, 0);
  printf("numerics: %s\n", ok ? "ALL OK" : "FAILURES");
  // ---- part C
  rate(1, 128, 64, 0, 0);
  rate(1, 128, 128, 0, 0);
  rate(1, 128, 256, 0, 0);
  rate(1, 128, 256, 0, 1);
  rate(2, 256, 256, 0, 0);
  rate(2, 256, 128, 0, 0);
  rate(2, 128, 256, 0, 0);
  rate(2, 128, 256, 0, 1);
  rate(2, 128, 128, 0, 0);
  return ok ? 0 : 1;
}
