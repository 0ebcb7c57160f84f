// Projection-head tail fused in front of the loss:
//
//     e = LayerNorm(h W^T + b) * gamma + beta         nn.Linear(hidden, p) -> nn.LayerNorm(p), the last two layers of
//     rinv = 1 / max(|e|, eps)                        ProjectionHead (old/clip.py:26-33), OptimizedProjectionHead
//                                                     (old/clip_opt.py:16-44), the notebook heads
//                                                     (current/rna_clip_codes.ipynb:1901-1909); F.normalize of old/clip.py:63-64
//
// as ONE kernel: a tcgen05 GEMM whose epilogue is the LayerNorm and the row norm.  The output width p <= 512 fits the 512
// TMEM columns of one CTA, so a thread (= TMEM lane = row) sees its whole output row: mean / variance / the norm of the
// rounded row are thread-local, no cross-lane traffic.  The kernel writes the bf16 rows the contrastive kernels read,
// their 1/norm (the loss kernels scale the fp32 accumulator by it: the normalise stays fused), and what the LayerNorm
// backward needs (z_hat in bf16, rstd) -- the [N, p] fp32 pre-activation never exists in HBM.
//
// One CTA = 128 rows x p columns (cta_group::1, M = 128, N = 256 instructions, K = hidden streamed in 64-wide boxes);
// h [N, K] and W [p, K] are both K-major, exactly what nn.Linear stores.  The shapes this runs on (N = 4096 ... 65536,
// K = 1024 ... 1280: 5 - 90 GFLOP) are launch- and latency-bound rather than tensor-bound; the point of the kernel is the
// three [N, p] round trips it removes (Linear output, LayerNorm output, normalised copy).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "ptx.cuh"

namespace head {

constexpr int THREADS = 256;            // warp 0 TMA, 1 MMA, 2 TMEM alloc, 4..7 epilogue (lane quarter = warp & 3)
constexpr int ROWS = 128;
constexpr int A_BYTES = ROWS * 128;     // [128 rows][64 k] bf16
constexpr int SMALL = 8192;             // barriers | tmem ptr | bias, gamma, beta [3][512] f32 at +1024
constexpr int MAX_STAGES = 4;
constexpr int B_FULL = 0, B_EMPTY = MAX_STAGES, B_ACC = 2 * MAX_STAGES;

struct Params {
  int n, k, p;          // rows, hidden, output width (p % 128 == 0, p <= 512; k % 64 == 0)
  int nkb, stages;
  float ln_eps, norm_eps;
  const float* bias;    // [p] or nullptr
  const float* gamma;   // [p]
  const float* beta;    // [p]
  __nv_bfloat16* e;     // [n, p]
  __nv_bfloat16* zhat;  // [n, p] or nullptr
  float* rstd;          // [n] or nullptr
  float* rinv;          // [n]
};

__host__ __device__ constexpr int stage_bytes(int p) { return A_BYTES + p * 128; }
__host__ __device__ constexpr int smem_bytes(int p, int stages) { return stages * stage_bytes(p) + SMALL; }

__global__ void __launch_bounds__(THREADS, 1)
linear_ln_kernel(const __grid_constant__ CUtensorMap tmap_h,   // h box {64 k, 128 rows}
                 const __grid_constant__ CUtensorMap tmap_w,   // W box {64 k, 128 rows}
                 const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = ptx::smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int i0 = blockIdx.x * ROWS;
  const int sbytes = stage_bytes(p.p);
  const uint32_t small_off = p.stages * sbytes;
  const uint32_t bars = base + small_off;
  auto bar = [&](int i) -> uint32_t { return bars + 8u * i; };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem + small_off + 512);
  float* const vec = reinterpret_cast<float*>(smem + small_off + 1024);   // [3][512]: bias, gamma, beta

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_h);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      ptx::mbar_init(bar(B_FULL + s), 1);
      ptx::mbar_init(bar(B_EMPTY + s), 1);
    }
    ptx::mbar_init(bar(B_ACC), 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(ptx::smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), 512);
    ptx::tmem_relinquish();
  }
  for (int c = threadIdx.x; c < p.p; c += THREADS) {
    vec[c] = p.bias ? p.bias[c] : 0.f;
    vec[512 + c] = p.gamma[c];
    vec[1024 + c] = p.beta[c];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int nq = p.p >> 7;   // 128-row W boxes per stage

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        ptx::mbar_wait(bar(B_EMPTY + stage), phase ^ 1u);
        ptx::mbar_arrive_expect_tx(bar(B_FULL + stage), sbytes);
        const uint32_t st = base + stage * sbytes;
        ptx::tma_load_2d(st, &tmap_h, bar(B_FULL + stage), kb * 64, i0);
        for (int q = 0; q < nq; ++q) ptx::tma_load_2d(st + A_BYTES + q * 16384, &tmap_w, bar(B_FULL + stage), kb * 64, q * 128);
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t desc_hi = static_cast<uint32_t>(ptx::smem_desc_k_sw128(0) >> 32);
      auto desc = [&](uint32_t addr) -> uint64_t {
        return (static_cast<uint64_t>(desc_hi) << 32) | (((addr & 0x3FFFFu) >> 4) | (1u << 16));
      };
      int stage = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        ptx::mbar_wait(bar(B_FULL + stage), phase);
        ptx::tc_fence_after();
        const uint32_t st = base + stage * sbytes;
        // output columns in chunks of up to 256 (one instruction's N); W rows of a chunk are contiguous 128-row boxes
        for (int c0 = 0; c0 < p.p; c0 += 256) {
          const int wn = min(256, p.p - c0);
          const uint32_t idesc = ptx::idesc_bf16_f32(ROWS, wn);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::mma_f16(tmem_base + c0, desc(st + 32 * k), desc(st + A_BYTES + (c0 >> 7) * 16384 + 32 * k), idesc, (kb | k) != 0);
        }
        ptx::mma_commit(bar(B_EMPTY + stage));
        if (kb == p.nkb - 1) ptx::mma_commit(bar(B_ACC));
        if (++stage == p.stages) { stage = 0; phase ^= 1u; }
      }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue: thread = row
    const int q = warp & 3;
    const int row = i0 + 32 * q + lane;
    const bool ok = row < p.n;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    ptx::mbar_wait(bar(B_ACC), 0);
    ptx::tc_fence_after();
    // pass 1: mean and variance of z = acc + bias (two-pass over TMEM: shifted sums are not needed, the data is re-read)
    float sum = 0.f;
    for (int c = 0; c < p.p; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(t_lane + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) sum += __uint_as_float(r[x]) + vec[c + x];
    }
    const float mean = sum / (float)p.p;
    float var = 0.f;
    for (int c = 0; c < p.p; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(t_lane + c, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; ++x) {
        const float dz = __uint_as_float(r[x]) + vec[c + x] - mean;
        var = fmaf(dz, dz, var);
      }
    }
    const float rstd = rsqrtf(var / (float)p.p + p.ln_eps);
    // pass 2: z_hat, e = z_hat gamma + beta (rounded to bf16), |e|^2 of the ROUNDED row (what the tensor cores will see)
    float ss = 0.f;
    for (int c = 0; c < p.p; c += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(t_lane + c, r);
      ptx::tmem_ld_wait();
      uint32_t pe[16], pz[16];
#pragma unroll
      for (int x = 0; x < 32; x += 2) {
        const float z0 = (__uint_as_float(r[x]) + vec[c + x] - mean) * rstd;
        const float z1 = (__uint_as_float(r[x + 1]) + vec[c + x + 1] - mean) * rstd;
        const __nv_bfloat162 e2 = __floats2bfloat162_rn(fmaf(z0, vec[512 + c + x], vec[1024 + c + x]),
                                                        fmaf(z1, vec[512 + c + x + 1], vec[1024 + c + x + 1]));
        const float2 ef = __bfloat1622float2(e2);
        ss = fmaf(ef.x, ef.x, ss);
        ss = fmaf(ef.y, ef.y, ss);
        const __nv_bfloat162 z2 = __floats2bfloat162_rn(z0, z1);
        pe[x >> 1] = *reinterpret_cast<const uint32_t*>(&e2);
        pz[x >> 1] = *reinterpret_cast<const uint32_t*>(&z2);
      }
      if (ok) {
        uint4* de = reinterpret_cast<uint4*>(p.e + (long long)row * p.p + c);
#pragma unroll
        for (int v = 0; v < 4; ++v) de[v] = make_uint4(pe[4 * v], pe[4 * v + 1], pe[4 * v + 2], pe[4 * v + 3]);
        if (p.zhat != nullptr) {
          uint4* dz = reinterpret_cast<uint4*>(p.zhat + (long long)row * p.p + c);
#pragma unroll
          for (int v = 0; v < 4; ++v) dz[v] = make_uint4(pz[4 * v], pz[4 * v + 1], pz[4 * v + 2], pz[4 * v + 3]);
        }
      }
    }
    if (ok) {
      p.rinv[row] = 1.f / fmaxf(sqrtf(ss), p.norm_eps);
      if (p.rstd != nullptr) p.rstd[row] = rstd;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// LayerNorm backward of the tail, one warp per row:  dz = rstd (g - mean(g) - z_hat mean(g z_hat)),  g = de * gamma.
// de [n, p] (bf16 or f32): gradient with respect to e; writes dz [n, p] bf16 (the operand of the two weight/input GEMMs).
template <typename TG>
__global__ void ln_backward_rows(const TG* __restrict__ de, const __nv_bfloat16* __restrict__ zhat, const float* __restrict__ rstd,
                                 const float* __restrict__ gamma, int64_t n, int p, __nv_bfloat16* __restrict__ dz) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const TG* g = de + row * p;
  const __nv_bfloat16* zr = zhat + row * p;
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < p; c += 32) {
    const float gv = (float)g[c] * gamma[c];
    s1 += gv;
    s2 = fmaf(gv, __bfloat162float(zr[c]), s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  const float m1 = s1 / (float)p, m2 = s2 / (float)p, rs = rstd[row];
  for (int c = lane; c < p; c += 32) {
    const float gv = (float)g[c] * gamma[c];
    dz[row * p + c] = __float2bfloat16_rn(rs * (gv - m1 - __bfloat162float(zr[c]) * m2));
  }
}

}  // namespace head
