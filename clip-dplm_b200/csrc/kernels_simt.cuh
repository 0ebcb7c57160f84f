// Exact CUDA-core kernels (fp32 math, true online max).  They serve
//   - the fp32 check mode (parity <= 1e-5 against the reference, north_star),
//   - bf16 inputs the tensor-core kernels do not take: 2*s > 86 (exp(S - s) would leave fp32,
//     e.g. logit_scale clamped at 100, old/clip_opt.py:100), d % 8 != 0 or d > 768.
// Same contract as the tcgen05 kernels: logits are formed tile by tile and never stored.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include "kernels_aux.cuh"

namespace simt {

constexpr int TILE = 64;   // logits tile (rows x cols) per block
constexpr int KT = 16;     // contraction chunk staged in shared memory
constexpr int THREADS = 256;

// S tile (64x64): S_ij = scale * rinv_x[i] rinv_y[j] <x_i, y_j>; thread (ty,tx) owns rows ty*4.., cols tx*4..
// The dot products are accumulated in fp64 so that the check mode is limited by the fp32 inputs and
// the fp32 soft-max arithmetic only (the reference's own fp32 noise is ~1e-5 on the gradients).
template <typename T>
__device__ __forceinline__ void logits_tile(const T* __restrict__ x, const T* __restrict__ y,
                                            const float* __restrict__ rinv_x, const float* __restrict__ rinv_y,
                                            int64_t n_rows, int64_t n_cols, int d, int64_t i0, int64_t j0, float scale,
                                            float (&s)[4][4], float (*Xs)[TILE + 1], float (*Ys)[TILE + 1]) {
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int k0 = 0; k0 < d; k0 += KT) {
    // 64 rows x 16 k per operand = 1024 elements, 4 per thread; stored [k][row] for conflict-free reads
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      int idx = threadIdx.x + e * THREADS;
      int r = idx >> 4, k = idx & 15;
      int64_t gi = i0 + r, gj = j0 + r;
      int kk = k0 + k;
      Xs[k][r] = (gi < n_rows && kk < d) ? aux::ld_f(x + gi * d + kk) : 0.f;
      Ys[k][r] = (gj < n_cols && kk < d) ? aux::ld_f(y + gj * d + kk) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < KT; ++k) {
      float xa[4], yb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) xa[a] = Xs[k][ty * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) yb[b] = Ys[k][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma((double)xa[a], (double)yb[b], acc[a][b]);
    }
    __syncthreads();
  }
  float rx[4], ry[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) rx[a] = (i0 + ty * 4 + a < n_rows) ? rinv_x[i0 + ty * 4 + a] : 0.f;
#pragma unroll
  for (int b = 0; b < 4; ++b) ry[b] = (j0 + tx * 4 + b < n_cols) ? rinv_y[j0 + tx * 4 + b] : 0.f;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) s[a][b] = (float)(acc[a][b] * ((double)scale * (double)rx[a] * (double)ry[b]));
}

// Forward statistics: per-tile (max, sumexp) partials for rows and columns + the diagonal.
// grid = (col tiles, row tiles).  row_pm/pl: [n_jt][n_rows];  col_pm/pl: [n_it][n_cols].
template <typename T>
__global__ void __launch_bounds__(THREADS)
fwd_stats(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ rinv_x,
          const float* __restrict__ rinv_y, int64_t n_rows, int64_t n_cols, int d, int64_t diag_offset, float scale, const float* __restrict__ scale_dev, float* __restrict__ row_pm, float* __restrict__ row_pl,
          float* __restrict__ col_pm, float* __restrict__ col_pl, float* __restrict__ diag) {
  __shared__ float Xs[KT][TILE + 1];
  __shared__ float Ys[KT][TILE + 1];
  __shared__ float St[TILE][TILE + 1];
  const int64_t j0 = (int64_t)blockIdx.x * TILE, i0 = (int64_t)blockIdx.y * TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (scale_dev != nullptr) scale = __ldg(scale_dev);   // device-resident s (no host read per step)
  float s[4][4];
  logits_tile<T>(x, y, rinv_x, rinv_y, n_rows, n_cols, d, i0, j0, scale, s, Xs, Ys);
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int64_t gi = i0 + ty * 4 + a, gj = j0 + tx * 4 + b;
      bool ok = gi < n_rows && gj < n_cols;
      St[ty * 4 + a][tx * 4 + b] = ok ? s[a][b] : -INFINITY;
      if (ok && gj == gi + diag_offset) diag[gi] = s[a][b];
    }
  __syncthreads();
  const int t = threadIdx.x;
  if (t < TILE) {  // row t of the tile
    int64_t gi = i0 + t;
    if (gi < n_rows) {
      float m = -INFINITY;
      for (int c = 0; c < TILE; ++c) m = fmaxf(m, St[t][c]);
      float l = 0.f;
      for (int c = 0; c < TILE; ++c) l += expf(St[t][c] - m);   // exp(-inf) = 0 for masked entries
      row_pm[(int64_t)blockIdx.x * n_rows + gi] = m;
      row_pl[(int64_t)blockIdx.x * n_rows + gi] = (m > -INFINITY) ? l : 0.f;
    }
  } else if (t < 2 * TILE) {  // column t-64 of the tile
    int c = t - TILE;
    int64_t gj = j0 + c;
    if (gj < n_cols) {
      float m = -INFINITY;
      for (int r = 0; r < TILE; ++r) m = fmaxf(m, St[r][c]);
      float l = 0.f;
      for (int r = 0; r < TILE; ++r) l += expf(St[r][c] - m);
      col_pm[(int64_t)blockIdx.y * n_cols + gj] = m;
      col_pl[(int64_t)blockIdx.y * n_cols + gj] = (m > -INFINITY) ? l : 0.f;
    }
  }
}

// Backward, one side: dx[i0:i0+64, dd0:dd0+64] = out_scale * sum_j G_ij rinv_y[j] y_j with
// G_ij = exp(S_ij - rm_i) rw_i + exp(S_ij - cm_j) cw_j - diag_w [j == i + diag_offset].
// grid = (d tiles, row tiles); the x == 0 column of blocks also emits sum G.S partials.
template <typename T>
__global__ void __launch_bounds__(THREADS)
bwd_side(const T* __restrict__ x, const T* __restrict__ y, const float* __restrict__ rinv_x,
         const float* __restrict__ rinv_y, int64_t n_rows, int64_t n_cols, int d, int64_t diag_offset, float scale, const float* __restrict__ scale_dev, const float* __restrict__ row_m, const float* __restrict__ row_w,
         const float* __restrict__ col_m, const float* __restrict__ col_w, float diag_w, float out_scale, float* __restrict__ dx, float* __restrict__ ds_part) {
  __shared__ float Xs[KT][TILE + 1];
  __shared__ float Ys[KT][TILE + 1];
  __shared__ float Gs[TILE][TILE + 1];
  __shared__ float Yd[TILE][TILE + 1];
  __shared__ float red[THREADS];
  const int64_t i0 = (int64_t)blockIdx.y * TILE;
  const int dd0 = blockIdx.x * TILE;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  if (scale_dev != nullptr) {   // out_scale was formed with the host's hint of s: swap in the device value
    const float sd = __ldg(scale_dev);
    out_scale = out_scale / scale * sd;
    scale = sd;
  }
  float rm[4], rw[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int64_t gi = i0 + ty * 4 + a;
    rm[a] = (gi < n_rows) ? row_m[gi] : 0.f;
    rw[a] = (gi < n_rows) ? row_w[gi] : 0.f;
  }
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  float ds = 0.f;
  for (int64_t j0 = 0; j0 < n_cols; j0 += TILE) {
    float s[4][4];
    logits_tile<T>(x, y, rinv_x, rinv_y, n_rows, n_cols, d, i0, j0, scale, s, Xs, Ys);
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int64_t gj = j0 + tx * 4 + b;
      const bool has_c = col_w != nullptr && gj < n_cols;
      const float cm = has_c ? col_m[gj] : 0.f;
      const float cw = has_c ? col_w[gj] : 0.f;
      const float ryj = (gj < n_cols) ? rinv_y[gj] : 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        int64_t gi = i0 + ty * 4 + a;
        float g = 0.f;
        if (gi < n_rows && gj < n_cols) {
          g = expf(s[a][b] - rm[a]) * rw[a];
          if (cw != 0.f) g = fmaf(expf(s[a][b] - cm), cw, g);
          if (gj == gi + diag_offset) g -= diag_w;
          ds = fmaf(g, s[a][b], ds);
        }
        Gs[ty * 4 + a][tx * 4 + b] = g * ryj;   // dXhat = s * sum_j G_ij (rinv_y[j] y_j)
      }
    }
    // stage Y[j0:j0+64, dd0:dd0+64]
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      int idx = threadIdx.x + e * THREADS;
      int r = idx >> 6, c = idx & 63;
      int64_t gj = j0 + r;
      int dd = dd0 + c;
      Yd[r][c] = (gj < n_cols && dd < d) ? aux::ld_f(y + gj * d + dd) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < TILE; ++j) {
      float ga[4], yb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ga[a] = Gs[ty * 4 + a][j];
#pragma unroll
      for (int b = 0; b < 4; ++b) yb[b] = Yd[j][tx * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ga[a], yb[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      int64_t gi = i0 + ty * 4 + a;
      int dd = dd0 + tx * 4 + b;
      if (gi < n_rows && dd < d) dx[gi * d + dd] = acc[a][b] * out_scale;
    }
  if (ds_part != nullptr && blockIdx.x == 0) {
    red[threadIdx.x] = ds;
    __syncthreads();
    for (int st = THREADS >> 1; st > 0; st >>= 1) {
      if ((int)threadIdx.x < st) red[threadIdx.x] += red[threadIdx.x + st];
      __syncthreads();
    }
    if (threadIdx.x == 0) ds_part[blockIdx.y] = red[0];
  }
}

}  // namespace simt
