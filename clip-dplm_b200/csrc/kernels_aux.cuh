// HBM-bound helper kernels around the two contraction kernels: normalise (+ its backward),
// transpose, LSE plumbing on [N] vectors, deterministic reductions of per-CTA partials.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

namespace aux {

template <typename T> __device__ __forceinline__ float ld_f(const T* p);
template <> __device__ __forceinline__ float ld_f<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ld_f<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void st_f(T* p, float v);
template <> __device__ __forceinline__ void st_f<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void st_f<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr float kNormEps = 1e-12f;  // F.normalize default eps

// One lane's share of a row's sum of squares (the caller adds the lanes with warp_sum).  bf16 rows with d % 8 == 0 are read
// as 16-byte chunks.  Every kernel that needs 1/|x| goes through this function, so that a row's norm has the same bits
// wherever it is computed (normalize_rows on the owner, link::push_rows for the gathered copies).
template <typename TI>
__device__ __forceinline__ float row_sumsq_lane(const TI* __restrict__ xr, int d, int lane) {
  float ss = 0.f;
  if constexpr (sizeof(TI) == 2) {
    if ((d & 7) == 0 && (reinterpret_cast<uintptr_t>(xr) & 15u) == 0) {
      const uint4* xv = reinterpret_cast<const uint4*>(xr);
      for (int c = lane; c < (d >> 3); c += 32) {
        const uint4 v = xv[c];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = __bfloat1622float2(h[k]);
          ss = fmaf(f.x, f.x, ss);
          ss = fmaf(f.y, f.y, ss);
        }
      }
      return ss;
    }
  }
  for (int k = lane; k < d; k += 32) {
    const float v = ld_f(xr + k);
    ss = fmaf(v, v, ss);
  }
  return ss;
}

// One warp per row: rinv = 1 / max(|x|, eps), optionally x_hat = x / max(|x|, eps)   (old/clip.py:63-64).
template <typename TI, typename TO>
__global__ void normalize_rows(const TI* __restrict__ x, int64_t n, int d, TO* __restrict__ xh, float* __restrict__ rinv) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const TI* xr = x + row * d;
  const float ss = warp_sum(row_sumsq_lane(xr, d, lane));
  const float denom = fmaxf(sqrtf(ss), kNormEps);
  if (xh != nullptr) {
    TO* o = xh + row * d;
    for (int k = lane; k < d; k += 32) st_f(o + k, ld_f(xr + k) / denom);
  }
  if (lane == 0) rinv[row] = 1.f / denom;
}

// Operand staging: out_c = convert(in) [n,d] (optional) and out_t = convert(in)^T [d,ld_t] (optional).
// Padding columns n..ld_t of out_t are left untouched (TMA never reads them: the map's extent is n).
template <typename TI, typename TO>
__global__ void stage_operand(const TI* __restrict__ in, int64_t n, int d, TO* __restrict__ out_c, TO* __restrict__ out_t,
                              int64_t ld_t) {
  __shared__ TO tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int64_t rr = r0 + r;
    int cc = c0 + threadIdx.x;
    if (rr < n && cc < d) {
      TO v;
      st_f(&v, ld_f(in + rr * d + cc));
      tile[r][threadIdx.x] = v;
      if (out_c != nullptr) out_c[rr * d + cc] = v;
    }
  }
  if (out_t == nullptr) return;
  __syncthreads();
  for (int c = threadIdx.y; c < 32; c += blockDim.y) {
    int cc = c0 + c;
    int64_t rr = r0 + threadIdx.x;
    if (rr < n && cc < d) out_t[(int64_t)cc * ld_t + rr] = tile[threadIdx.x][c];
  }
}

// dx_i = rinv_i (g_i - xhat_i (xhat_i . g_i));  clamped rows (|x| < eps): dx_i = g_i * rinv_i.
template <typename TI, typename TO>
__global__ void normalize_rows_bwd(const TI* __restrict__ x, const float* __restrict__ rinv,
                                   const float* __restrict__ g, const float* __restrict__ grad_scale, int64_t n, int d,
                                   TO* __restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float ri = rinv[row];
  const float gs = grad_scale ? grad_scale[0] : 1.f;
  const TI* xr = x + row * d;
  const float* gr = g + row * d;
  TO* o = dx + row * d;
  const bool clamped = ri >= 0.5f / kNormEps;
  float dot = 0.f;
  if (!clamped) {
    for (int k = lane; k < d; k += 32) dot = fmaf(ld_f(xr + k) * ri, gr[k], dot);
    dot = warp_sum(dot);
  }
  for (int k = lane; k < d; k += 32) {
    float xh = ld_f(xr + k) * ri;
    st_f(o + k, (gr[k] - xh * dot) * (ri * gs));
  }
}

__global__ void softmax_weights(const float* __restrict__ l, int64_t n, float coef, float* __restrict__ w) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) w[i] = coef / l[i];   // l = +inf (column without positives) -> 0
}

__global__ void combine_lse(const float* __restrict__ m, const float* __restrict__ l, int64_t n, float* __restrict__ lse) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lse[i] = m[i] + logf(l[i]);
}

// Deterministic single-block loss reduction (fp64 accumulators, fixed order).
__global__ void loss_reduce(const float* __restrict__ row_m, const float* __restrict__ row_l,
                            const float* __restrict__ col_m, const float* __restrict__ col_l,
                            const float* __restrict__ diag, int64_t n_rows, int64_t diag_offset, double inv_denom,
                            int symmetric, float* __restrict__ loss) {
  __shared__ double sh[1024];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n_rows; i += blockDim.x) {
    // fp32 logf (1 ulp) + fp64 accumulation: the fp64 log cost 100 us on one SM at N = 65536 and buys nothing --
    // the sums l are fp32 quantities and ATen's log_softmax takes this log in fp32 as well
    const double dg = diag[i];
    acc += ((double)row_m[i] - dg) + (double)logf(row_l[i]);
    if (symmetric) acc += ((double)col_m[i + diag_offset] - dg) + (double)logf(col_l[i + diag_offset]);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = (float)(sh[0] * inv_denom);
}

// col_l[j] = sum_p part[p][j] (fixed order), col_m[j] = shift.   Used by the tensor-core forward
// whose partials all share the fixed shift s.
__global__ void reduce_col_partials(const float* __restrict__ part, int n_part, int64_t ld, int64_t n_cols, float shift,
                                    const float* __restrict__ shift_dev, float* __restrict__ col_m,
                                    float* __restrict__ col_l) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_cols) return;
  if (shift_dev != nullptr) shift = __ldg(shift_dev);
  float acc = 0.f;
  for (int p = 0; p < n_part; ++p) acc += part[(int64_t)p * ld + j];
  col_m[j] = shift;
  col_l[j] = acc;
}

// Sums taken relative to the offset shift p = s - off (the speculative large-scale forward): l = sum_p part[p][j] in fixed
// order, returned as an exactly rescaled pair (m = p + k ln 2, l 2^-k in [1, 2)) so that the soft-max weights coef / l of
// the backward stay normal numbers whatever the magnitude of the sum (up to n e^off).  A sum below `thr` means the
// entry's largest logit lies so far under the shift that terms were flushed to zero: *flag is raised and the exact
// sweeps that follow (gated on it) replace the statistics.
// n_pad > 0 (grouped launch): entry j belongs to a problem's padding when j % n_pad >= n_valid -- written, never flagged.
__global__ void reduce_shifted_partials(const float* __restrict__ part, int n_part, int64_t ld, int64_t n, float scale,
                                        const float* __restrict__ scale_dev, float off, float thr, float* __restrict__ out_m,
                                        float* __restrict__ out_l, int* __restrict__ flag, int64_t n_pad = 0,
                                        int64_t n_valid = 0) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  if (scale_dev != nullptr) scale = __ldg(scale_dev);
  float acc = 0.f;
  for (int p = 0; p < n_part; ++p) acc += part[(int64_t)p * ld + j];
  if (!(acc >= thr) || !(acc < INFINITY)) {
    if (n_pad == 0 || j % n_pad < n_valid) *flag = 1;   // every writer stores the same value
    out_m[j] = scale - off;
    out_l[j] = acc;
    return;
  }
  const int k = ilogbf(acc);
  out_m[j] = fmaf((float)k, 0.6931471805599453f, scale - off);
  out_l[j] = ldexpf(acc, -k);
}

// Combine (m,l) partial pairs: out over `n` entries, `n_part` partials with leading dimension ld.  gate: optional DEVICE
// flag, the kernel does nothing while it is 0.
__global__ void reduce_ml_partials(const float* __restrict__ pm, const float* __restrict__ pl, int n_part, int64_t ld,
                                   int64_t n, float* __restrict__ out_m, float* __restrict__ out_l,
                                   const int* __restrict__ gate = nullptr) {
  if (gate != nullptr && __ldg(gate) == 0) return;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float M = -INFINITY;
  for (int p = 0; p < n_part; ++p) M = fmaxf(M, pm[(int64_t)p * ld + i]);
  float L = 0.f;
  for (int p = 0; p < n_part; ++p) {
    float m = pm[(int64_t)p * ld + i];
    if (m > -INFINITY) L += pl[(int64_t)p * ld + i] * expf(m - M);
  }
  out_m[i] = M;
  out_l[i] = L;
}

// *dst += coef * sum_p part[p]   (single thread, fixed order -> deterministic).
__global__ void reduce_scalar_partials(const float* __restrict__ part, int n_part, float coef, float* __restrict__ dst) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double acc = 0.0;
    for (int p = 0; p < n_part; ++p) acc += (double)part[p];
    dst[0] += (float)(acc * (double)coef);
  }
}

// out = sum_s part[s] over n_split slabs of `n4` float4 each (fixed order).
__global__ void sum_splits(const float4* __restrict__ part, int n_split, int64_t n4, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 acc = part[i];
  for (int s = 1; s < n_split; ++s) {
    const float4 v = part[(int64_t)s * n4 + i];
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  out[i] = acc;
}

// Retrieval: merge `n_slot` candidate lists of KT entries per row into the row's k best (one warp per row; ties -> lower
// column index).  cand_* are [n_slot][n_rows][KT]; out_* are [n_rows][k].
__global__ void topk_merge(const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int n_slot, int64_t n_rows,
                           int kt, int k, float* __restrict__ out_score, int64_t* __restrict__ out_idx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int n_cand = n_slot * kt;                 // <= 32 * 64
  unsigned long long used = 0ull;                 // this lane's candidates c = lane + 32 u, u < 64
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    int best_i = 0x7fffffff, best_u = -1;
    for (int u = 0, c = lane; c < n_cand; ++u, c += 32) {
      if ((used >> u) & 1ull) continue;
      const int64_t o = ((int64_t)(c / kt) * n_rows + row) * kt + (c % kt);
      const float v = cand_score[o];
      const int ix = cand_idx[o];
      if (ix >= 0 && (v > best || (v == best && ix < best_i))) { best = v; best_i = ix; best_u = u; }
    }
    float wv = best;
    int wi = best_i;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; }
    }
    if (best_u >= 0 && best == wv && best_i == wi) used |= 1ull << best_u;   // (score, index) pairs are unique per row
    if (lane == 0) {
      out_score[row * k + r] = wi == 0x7fffffff ? -INFINITY : wv;
      out_idx[row * k + r] = wi == 0x7fffffff ? -1 : (int64_t)wi;
    }
  }
}

// part[block] = sum over the block's 8 rows of rinv_i <x_i, g_i>   (= <xhat_i, dxhat_i>; one warp per row)
template <typename TI>
__global__ void rowdot_partials(const TI* __restrict__ x, const float* __restrict__ rinv, const float* __restrict__ g,
                                int64_t n, int d, float* __restrict__ part) {
  __shared__ float sh[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  float dot = 0.f;
  if (row < n) {
    const TI* xr = x + row * d;
    const float* gr = g + row * d;
    for (int k = lane; k < d; k += 32) dot = fmaf(ld_f(xr + k), gr[k], dot);
    dot = warp_sum(dot) * rinv[row];
  }
  if (lane == 0) sh[w] = dot;
  __syncthreads();
  if (threadIdx.x == 0) part[blockIdx.x] = ((sh[0] + sh[1]) + (sh[2] + sh[3])) + ((sh[4] + sh[5]) + (sh[6] + sh[7]));
}

// The tail of one backward side in ONE pass over the gradient (one warp per row, the row staged in shared memory):
//   g_i   = sum over the n_split partial slabs of the column-sweep work items (fixed order; n_split = 1: the finished row)
//   part  = sum over the block's 8 rows of rinv_i <xc_i, g_i>        (= <xhat_i, dxhat_i>: sum G.S, d logit_scale)
//   dx_i  = rinv_i (g_i - xhat_i (xhat_i . g_i)) * grad_scale        (normalise backward; clamped rows: g_i rinv_i)
// xc = the rows the contraction saw (compute type), xo = the caller's rows (their own type): the same pointer for bf16
// inputs.  Replaces sum_splits + rowdot_partials + normalize_rows_bwd (three passes over [n,d] fp32).
template <typename TC, typename TI, typename TO>
__global__ void finish_rows(const float* __restrict__ parts, int n_split, int64_t slab, const TC* __restrict__ xc,
                            const TI* __restrict__ xo, const float* __restrict__ rinv,
                            const float* __restrict__ grad_scale, int64_t n, int d, TO* __restrict__ dx,
                            float* __restrict__ ds_part) {
  extern __shared__ float g_sh[];   // [8][d]
  __shared__ float dot_sh[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  float* g = g_sh + (size_t)w * d;
  float dot_c = 0.f;
  if (row < n) {
    const float ri = rinv[row];
    const float gs = grad_scale ? grad_scale[0] : 1.f;
    const float* pr = parts + row * d;
    const TC* xcr = xc + row * d;
    const TI* xor_ = xo + row * d;
    const bool same = reinterpret_cast<const void*>(xc) == reinterpret_cast<const void*>(xo);
    float dot_o = 0.f;
    for (int k = lane; k < d; k += 32) {
      float acc = pr[k];
      for (int s = 1; s < n_split; ++s) acc += pr[(int64_t)s * slab + k];
      g[k] = acc;
      const float vc = ld_f(xcr + k);
      dot_c = fmaf(vc, acc, dot_c);
      if (!same) dot_o = fmaf(ld_f(xor_ + k), acc, dot_o);
    }
    dot_c = warp_sum(dot_c) * ri;
    dot_o = same ? dot_c : warp_sum(dot_o) * ri;
    const bool clamped = ri >= 0.5f / kNormEps;
    if (clamped) dot_o = 0.f;
    __syncwarp();
    for (int k = lane; k < d; k += 32) {
      const float xh = ld_f(xor_ + k) * ri;
      st_f(dx + row * d + k, (g[k] - xh * dot_o) * (ri * gs));
    }
  }
  if (ds_part != nullptr) {
    if (lane == 0) dot_sh[w] = dot_c;
    __syncthreads();
    if (threadIdx.x == 0)
      ds_part[blockIdx.x] = ((dot_sh[0] + dot_sh[1]) + (dot_sh[2] + dot_sh[3])) + ((dot_sh[4] + dot_sh[5]) + (dot_sh[6] + dot_sh[7]));
  }
}

// four consecutive elements (4-element aligned) as floats / from floats
template <typename T> __device__ __forceinline__ float4 ld4_f(const T* p);
template <> __device__ __forceinline__ float4 ld4_f<float>(const float* p) { return *reinterpret_cast<const float4*>(p); }
template <> __device__ __forceinline__ float4 ld4_f<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void st4_f(T* p, float4 v);
template <> __device__ __forceinline__ void st4_f<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void st4_f<__nv_bfloat16>(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// finish_rows with 16-byte accesses (d % 4 == 0, all row bases 4-element aligned): every lane owns four consecutive
// features per 128-feature stripe; the n_split partial rows are independent 16-byte loads (memory-level parallelism is
// what this pass lives on: at n_split = 8 the scalar version took 52 us for 8192 x 512 against 25 us of sum_splits alone).
template <typename TC, typename TI, typename TO>
__global__ void finish_rows_v4(const float* __restrict__ parts, int n_split, int64_t slab, const TC* __restrict__ xc,
                               const TI* __restrict__ xo, const float* __restrict__ rinv,
                               const float* __restrict__ grad_scale, int64_t n, int d, TO* __restrict__ dx,
                               float* __restrict__ ds_part) {
  extern __shared__ __align__(16) float g4_sh[];   // [8][d]
  __shared__ float dot_sh[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  float* g = g4_sh + (size_t)w * d;
  float dot_c = 0.f;
  if (row < n) {
    const float ri = rinv[row];
    const float gs = grad_scale ? grad_scale[0] : 1.f;
    const float* pr = parts + row * d;
    const TC* xcr = xc + row * d;
    const TI* xor_ = xo + row * d;
    const bool same = reinterpret_cast<const void*>(xc) == reinterpret_cast<const void*>(xo);
    float dot_o = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      float4 acc = ld4_f(pr + k);
#pragma unroll 4
      for (int s = 1; s < n_split; ++s) {
        const float4 v = ld4_f(pr + (int64_t)s * slab + k);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(g + k) = acc;
      const float4 vc = ld4_f(xcr + k);
      dot_c = fmaf(vc.x, acc.x, dot_c); dot_c = fmaf(vc.y, acc.y, dot_c);
      dot_c = fmaf(vc.z, acc.z, dot_c); dot_c = fmaf(vc.w, acc.w, dot_c);
      if (!same) {
        const float4 vo = ld4_f(xor_ + k);
        dot_o = fmaf(vo.x, acc.x, dot_o); dot_o = fmaf(vo.y, acc.y, dot_o);
        dot_o = fmaf(vo.z, acc.z, dot_o); dot_o = fmaf(vo.w, acc.w, dot_o);
      }
    }
    dot_c = warp_sum(dot_c) * ri;
    dot_o = same ? dot_c : warp_sum(dot_o) * ri;
    if (ri >= 0.5f / kNormEps) dot_o = 0.f;   // clamped row: dx = g * rinv
    const float k2 = ri * gs;
    for (int k = lane * 4; k < d; k += 128) {   // every lane re-reads only what it wrote: no barrier needed
      const float4 gv = *reinterpret_cast<const float4*>(g + k);
      const float4 xv = ld4_f(xor_ + k);
      float4 o;
      o.x = (gv.x - xv.x * ri * dot_o) * k2;
      o.y = (gv.y - xv.y * ri * dot_o) * k2;
      o.z = (gv.z - xv.z * ri * dot_o) * k2;
      o.w = (gv.w - xv.w * ri * dot_o) * k2;
      st4_f(dx + row * d + k, o);
    }
  }
  if (ds_part != nullptr) {
    if (lane == 0) dot_sh[w] = dot_c;
    __syncthreads();
    if (threadIdx.x == 0)
      ds_part[blockIdx.x] = ((dot_sh[0] + dot_sh[1]) + (dot_sh[2] + dot_sh[3])) + ((dot_sh[4] + dot_sh[5]) + (dot_sh[6] + dot_sh[7]));
  }
}

// Two finish_rows_v4 passes in ONE launch (blockIdx.y = side): the tails of the row-sharded two-sided backward -- side 0 sums
// the segment slabs of dA_hat, side 1 the ranks' slots of dB_hat -- share the GPU instead of queueing behind each other.
struct FinishSide {
  const float* parts;
  int n_split;
  int64_t slab;
  const void* xc;
  const void* xo;
  const float* rinv;
  void* dx;
  float* ds_part;
};
template <typename TC, typename TI, typename TO>
__global__ void finish_rows_v4_dual(FinishSide s0, FinishSide s1, const float* __restrict__ grad_scale, int64_t n, int d) {
  const FinishSide& s = blockIdx.y == 0 ? s0 : s1;
  extern __shared__ __align__(16) float g4_sh[];   // [8][d]
  __shared__ float dot_sh[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * 8 + w;
  float* g = g4_sh + (size_t)w * d;
  float dot_c = 0.f;
  if (row < n) {
    const float ri = s.rinv[row];
    const float gs = grad_scale ? grad_scale[0] : 1.f;
    const float* pr = s.parts + row * d;
    const TC* xcr = reinterpret_cast<const TC*>(s.xc) + row * d;
    const TI* xor_ = reinterpret_cast<const TI*>(s.xo) + row * d;
    const bool same = s.xc == s.xo;
    float dot_o = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      float4 acc = ld4_f(pr + k);
#pragma unroll 4
      for (int q = 1; q < s.n_split; ++q) {
        const float4 v = ld4_f(pr + (int64_t)q * s.slab + k);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(g + k) = acc;
      const float4 vc = ld4_f(xcr + k);
      dot_c = fmaf(vc.x, acc.x, dot_c); dot_c = fmaf(vc.y, acc.y, dot_c);
      dot_c = fmaf(vc.z, acc.z, dot_c); dot_c = fmaf(vc.w, acc.w, dot_c);
      if (!same) {
        const float4 vo = ld4_f(xor_ + k);
        dot_o = fmaf(vo.x, acc.x, dot_o); dot_o = fmaf(vo.y, acc.y, dot_o);
        dot_o = fmaf(vo.z, acc.z, dot_o); dot_o = fmaf(vo.w, acc.w, dot_o);
      }
    }
    dot_c = warp_sum(dot_c) * ri;
    dot_o = same ? dot_c : warp_sum(dot_o) * ri;
    if (ri >= 0.5f / kNormEps) dot_o = 0.f;
    const float k2 = ri * gs;
    TO* out = reinterpret_cast<TO*>(s.dx);
    for (int k = lane * 4; k < d; k += 128) {
      const float4 gv = *reinterpret_cast<const float4*>(g + k);
      const float4 xv = ld4_f(xor_ + k);
      float4 o;
      o.x = (gv.x - xv.x * ri * dot_o) * k2;
      o.y = (gv.y - xv.y * ri * dot_o) * k2;
      o.z = (gv.z - xv.z * ri * dot_o) * k2;
      o.w = (gv.w - xv.w * ri * dot_o) * k2;
      st4_f(out + row * d + k, o);
    }
  }
  if (s.ds_part != nullptr) {
    if (lane == 0) dot_sh[w] = dot_c;
    __syncthreads();
    if (threadIdx.x == 0)
      s.ds_part[blockIdx.x] = ((dot_sh[0] + dot_sh[1]) + (dot_sh[2] + dot_sh[3])) + ((dot_sh[4] + dot_sh[5]) + (dot_sh[6] + dot_sh[7]));
  }
}

// *dst += coef * sum_p part[p]: 256 threads, strided fp64 partial sums + a fixed-order tree (deterministic).
__global__ void reduce_scalar_partials_par(const float* __restrict__ part, int n_part, float coef, float* __restrict__ dst) {
  __shared__ double sh[256];
  double acc = 0.0;
  for (int p = threadIdx.x; p < n_part; p += 256) acc += (double)part[p];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dst[0] += (float)(sh[0] * (double)coef);
}

// ---- several pair problems in one launch (kernels_pair.cuh `Group`; the tri-modal model of tf_clip_codes (1).ipynb:13152-13165)
// Per-problem mean loss of the symmetric InfoNCE, problem k over the virtual rows / columns [k n_pad, k n_pad + n):
// loss[k] = [ sum_i (r_i - diag_i) + (c_i - diag_i) ] / (2 n), loss[n_prob] = their sum.  stat_* = [row statistics | column
// statistics], each [n_prob n_pad].  One block, fp64 accumulators, fixed order.
__global__ void loss_reduce_group(const float* __restrict__ stat_m, const float* __restrict__ stat_l,
                                  const float* __restrict__ diag, int n_prob, int64_t n, int64_t n_pad,
                                  float* __restrict__ loss) {
  __shared__ double sh[1024];
  const int64_t V = (int64_t)n_prob * n_pad;
  double total = 0.0;
  for (int k = 0; k < n_prob; ++k) {
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const int64_t r = (int64_t)k * n_pad + i;
      const double dg = diag[r];
      acc += ((double)stat_m[r] - dg) + (double)logf(stat_l[r]);
      acc += ((double)stat_m[V + r] - dg) + (double)logf(stat_l[V + r]);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x >> 1; s > 0; s >>= 1) {
      if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
      __syncthreads();
    }
    const double lk = sh[0] / (2.0 * (double)n);
    if (threadIdx.x == 0) loss[k] = (float)lk;
    total += lk;
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[n_prob] = (float)total;
}

// finish_rows_v4 for the stacked members of a group: the gradient of member m's normalised rows is the sum of the slabs of
// every virtual problem in which m was the resident operand (its X sides and its Y sides), times the splits -- summed here
// in fixed order, followed by the row dots and the normalise backward as in finish_rows.  grid = (ceil(n_pad / 8), members).
//   ds_part[m][block] = sum over the block's rows of <xhat_i, g_i>   (every pair's sum G.S appears once per side: halve)
//   sq_part[m][block] = sum over the block's rows of |dx_i|^2        (the embeddings' share of a gradient norm)
constexpr int FG_MEMBERS = 4, FG_CONTRIB = 8;
struct FinishGroup {
  int n_contrib[FG_MEMBERS];
  int vprob[FG_MEMBERS][FG_CONTRIB];   // virtual problems whose resident rows are member m
  long long n_pad, n_valid;
};
template <typename TC, typename TI, typename TO>
__global__ void finish_rows_group(const float* __restrict__ parts, int n_split, int64_t slab, FinishGroup fg,
                                  const TC* __restrict__ xc, const TI* __restrict__ xo, const float* __restrict__ rinv,
                                  int d, TO* __restrict__ dx, float* __restrict__ ds_part, float* __restrict__ sq_part) {
  extern __shared__ __align__(16) float g4_sh[];   // [8][d]
  __shared__ float dot_sh[8], sq_sh[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int m = blockIdx.y;
  const int64_t row_m = (int64_t)blockIdx.x * 8 + w;       // row within the member
  const int64_t row = (int64_t)m * fg.n_pad + row_m;       // stack row
  float* g = g4_sh + (size_t)w * d;
  float dot_c = 0.f, sq = 0.f;
  if (row_m < fg.n_valid) {
    const float ri = rinv[row];
    const TC* xcr = xc + row * d;
    const TI* xor_ = xo + row * d;
    const bool same = reinterpret_cast<const void*>(xc) == reinterpret_cast<const void*>(xo);
    const int nc = fg.n_contrib[m];
    float dot_o = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int c = 0; c < nc; ++c) {
        const float* pr = parts + ((int64_t)fg.vprob[m][c] * fg.n_pad + row_m) * d + k;
#pragma unroll 4
        for (int s = 0; s < n_split; ++s) {
          const float4 v = ld4_f(pr + (int64_t)s * slab);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      *reinterpret_cast<float4*>(g + k) = acc;
      const float4 vc = ld4_f(xcr + k);
      dot_c = fmaf(vc.x, acc.x, dot_c); dot_c = fmaf(vc.y, acc.y, dot_c);
      dot_c = fmaf(vc.z, acc.z, dot_c); dot_c = fmaf(vc.w, acc.w, dot_c);
      if (!same) {
        const float4 vo = ld4_f(xor_ + k);
        dot_o = fmaf(vo.x, acc.x, dot_o); dot_o = fmaf(vo.y, acc.y, dot_o);
        dot_o = fmaf(vo.z, acc.z, dot_o); dot_o = fmaf(vo.w, acc.w, dot_o);
      }
    }
    dot_c = warp_sum(dot_c) * ri;
    dot_o = same ? dot_c : warp_sum(dot_o) * ri;
    if (ri >= 0.5f / kNormEps) dot_o = 0.f;
    for (int k = lane * 4; k < d; k += 128) {
      const float4 gv = *reinterpret_cast<const float4*>(g + k);
      const float4 xv = ld4_f(xor_ + k);
      float4 o;
      o.x = (gv.x - xv.x * ri * dot_o) * ri;
      o.y = (gv.y - xv.y * ri * dot_o) * ri;
      o.z = (gv.z - xv.z * ri * dot_o) * ri;
      o.w = (gv.w - xv.w * ri * dot_o) * ri;
      sq = fmaf(o.x, o.x, sq); sq = fmaf(o.y, o.y, sq); sq = fmaf(o.z, o.z, sq); sq = fmaf(o.w, o.w, sq);
      st4_f(dx + row * d + k, o);
    }
    sq = warp_sum(sq);
  }
  if (lane == 0) { dot_sh[w] = dot_c; sq_sh[w] = sq; }
  __syncthreads();
  if (threadIdx.x == 0) {
    const int64_t o = (int64_t)m * gridDim.x + blockIdx.x;
    ds_part[o] = ((dot_sh[0] + dot_sh[1]) + (dot_sh[2] + dot_sh[3])) + ((dot_sh[4] + dot_sh[5]) + (dot_sh[6] + dot_sh[7]));
    sq_part[o] = ((sq_sh[0] + sq_sh[1]) + (sq_sh[2] + sq_sh[3])) + ((sq_sh[4] + sq_sh[5]) + (sq_sh[6] + sq_sh[7]));
  }
}

// block 0: *d_scale_sum += 0.5 sum ds_part[all];  block 1 + m: sumsq[m] = sum sq_part[m][:]   (fp64, fixed-order tree)
__global__ void reduce_group_scalars(const float* __restrict__ ds_part, const float* __restrict__ sq_part, int n_members,
                                     int n_blk, float* __restrict__ d_scale_sum, float* __restrict__ sumsq) {
  __shared__ double sh[256];
  const int b = blockIdx.x;
  const float* src = b == 0 ? ds_part : sq_part + (int64_t)(b - 1) * n_blk;
  const int cnt = b == 0 ? n_members * n_blk : n_blk;
  if ((b == 0 && d_scale_sum == nullptr) || (b > 0 && sumsq == nullptr)) return;
  double acc = 0.0;
  for (int p = threadIdx.x; p < cnt; p += 256) acc += (double)src[p];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (b == 0) d_scale_sum[0] += (float)(0.5 * sh[0]);
    else sumsq[b - 1] = (float)sh[0];
  }
}

}  // namespace aux
