// Exchange steps of the row-sharded batch as kernels over NVLink / NVSwitch PEER MEMORY (what the
// reference does with dist.all_gather + torch.cat: old/clip_opt.py:102-112, run1/full.py:77-84).
//
// Every rank owns one "symmetric" buffer of identical layout, mapped into all ranks' address spaces (the host exchanges
// the mappings once; here torch symmetric memory).  `Peers.base[r]` is rank r's buffer as seen from this GPU.  The data
// kernels STORE into the peers' buffers (push: fire-and-forget stores ride NVLink at line rate, loads would pay the
// round trip); a barrier kernel publishes them: st.release.sys of an epoch into every peer's flag word, ld.acquire.sys
// spin on the own flag words.  Epochs are counted on the device, so a CUDA graph replays the same kernels unchanged.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>

#include "kernels_aux.cuh"

namespace link {

constexpr int MAX_WORLD = 16;
constexpr int MAX_PHASE = 8;
constexpr int MAX_SCALARS = 8;
// control block at the start of every symmetric buffer (zero-initialised by the host before the first use)
constexpr int64_t OFF_FLAGS = 0;                                        // u32 [MAX_PHASE][MAX_WORLD]   written by peers
constexpr int64_t OFF_EPOCH = OFF_FLAGS + 4 * MAX_PHASE * MAX_WORLD;    // u32 [MAX_PHASE]              local only
constexpr int64_t OFF_STATUS = OFF_EPOCH + 4 * MAX_PHASE;               // u32 [4]: [0] != 0 -> a barrier timed out
constexpr int64_t OFF_SCALARS = OFF_STATUS + 16;                        // f32 [MAX_PHASE][2 parity][MAX_WORLD][MAX_SCALARS]
constexpr int64_t CONTROL_BYTES = ((OFF_SCALARS + 4 * MAX_PHASE * 2 * MAX_WORLD * MAX_SCALARS + 255) / 256) * 256;

struct Peers {
  void* base[MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_f32(float* p, float v) {
  asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

template <typename T>
__device__ __forceinline__ T* at(void* base, int64_t byte_off) {
  return reinterpret_cast<T*>(reinterpret_cast<char*>(base) + byte_off);
}

// Arrive at phase `phase` on every peer and wait for every peer's arrival (threads 0..world-1 of one block).  Returns
// the epoch.  Everything this GPU enqueued before the calling kernel -- its stores into peer memory included -- is
// ordered before the flag by the release (cumulative over the kernel boundary); peers' stores are visible to kernels
// launched after this one.  A peer that never arrives (a crashed rank) trips the timeout (minutes by default, like
// NCCL's watchdog): that is FATAL -- status[0] records the phase for the host's post-mortem and the kernel traps, so the
// step can never go on with unsynchronised peer buffers (the CUDA error surfaces at the next host synchronisation).
__device__ __forceinline__ uint32_t barrier_arrive_wait(const Peers& peers, int world, int rank, int phase,
                                                        unsigned long long timeout_ns, uint32_t* epoch_sh) {
  char* mine = reinterpret_cast<char*>(peers.base[rank]);
  uint32_t* status = at<uint32_t>(mine, OFF_STATUS);
  if (threadIdx.x == 0) {
    uint32_t* ep = at<uint32_t>(mine, OFF_EPOCH) + phase;
    const uint32_t e = *ep + 1u;
    *ep = e;
    *epoch_sh = e;
  }
  __syncthreads();
  const uint32_t e = *epoch_sh;
  if ((int)threadIdx.x < world) {
    const int r = threadIdx.x;
    __threadfence_system();
    st_release_sys(at<uint32_t>(peers.base[r], OFF_FLAGS) + phase * MAX_WORLD + rank, e);
    const uint32_t* flag = at<uint32_t>(mine, OFF_FLAGS) + phase * MAX_WORLD + r;
    const unsigned long long t0 = globaltimer_ns();
    while ((int32_t)(ld_acquire_sys(flag) - e) < 0) {
      if (globaltimer_ns() - t0 > timeout_ns) {
        atomicExch(status, 1u + (uint32_t)phase);
        __threadfence_system();
        __trap();
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  __threadfence_system();
  return e;
}

// The two halves of a barrier for data that is CONSUMED BLOCK BY BLOCK (the gather of the columns beside the forward sweep,
// clipnce_link_send_blocks / clipnce_forward_gathered): `epoch_advance` opens phase `phase` of a new step on this GPU,
// `signal` tells ONE peer that this rank's block has landed in its buffer (enqueued behind the copy that delivered it);
// the consumer kernel waits for flags[phase][src] >= epoch[phase] before it touches block src.
__global__ void epoch_advance(Peers peers, int rank, int phase) {
  if (threadIdx.x == 0 && blockIdx.x == 0) at<uint32_t>(peers.base[rank], OFF_EPOCH)[phase] += 1u;
}
__global__ void signal(Peers peers, int rank, int dst, int phase) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const uint32_t e = at<uint32_t>(peers.base[rank], OFF_EPOCH)[phase];
    __threadfence_system();
    st_release_sys(at<uint32_t>(peers.base[dst], OFF_FLAGS) + phase * MAX_WORLD + rank, e);
  }
}

__global__ void barrier(Peers peers, int world, int rank, int phase, unsigned long long timeout_ns) {
  __shared__ uint32_t epoch_sh;
  barrier_arrive_wait(peers, world, rank, phase, timeout_ns, &epoch_sh);
}

// Fused normalise + all-gather: one warp per local row computes rinv = 1 / max(|x|, eps) (F.normalize, old/clip.py:63-64)
// from the caller's rows and stores the row (in the compute type TO) and rinv into EVERY rank's gathered buffers at
// global row  row0 + i  (the own copy included: the gathered matrix is complete on every rank).  Destinations are
// visited starting behind the own rank, so that at any moment the ranks store to different peers (no ingress hot
// spot at the switch).  Grid-stride over rows: a launch with few (fat) blocks is a background variant that leaves most SMs
// to a contraction kernel running beside it -- the step itself uses the copy engines for that (clipnce_link_copy), which
// leave it all of them.
template <typename TI, typename TO>
__global__ void push_rows(const TI* __restrict__ x, int64_t n, int d, Peers peers, int world, int rank, int64_t rows_off,
                          int64_t rinv_off, int64_t row0) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  constexpr bool kVec = sizeof(TI) == 2 && sizeof(TO) == 2;   // bf16 -> bf16: 16-byte chunks (d % 8 == 0 checked by the host)
  for (int64_t row = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); row < n; row += (int64_t)gridDim.x * wpb) {
    const TI* xr = x + row * d;
    const float ss = aux::warp_sum(aux::row_sumsq_lane(xr, d, lane));   // the same bits as aux::normalize_rows
    if constexpr (kVec) {
      const uint4* xv = reinterpret_cast<const uint4*>(xr);
      const int nv = d >> 3;
      for (int c = lane; c < nv; c += 32) {
        const uint4 v = xv[c];
        for (int k = 1; k <= world; ++k) {
          int r = rank + k;
          r = r >= world ? r - world : r;
          at<uint4>(peers.base[r], rows_off)[(row0 + row) * nv + c] = v;
        }
      }
    } else {
      for (int k = lane; k < d; k += 32) {
        TO v;
        aux::st_f(&v, aux::ld_f(xr + k));
        for (int q = 1; q <= world; ++q) {
          int r = rank + q;
          r = r >= world ? r - world : r;
          at<TO>(peers.base[r], rows_off)[(row0 + row) * d + k] = v;
        }
      }
    }
    if (lane < world) at<float>(peers.base[lane], rinv_off)[row0 + row] = 1.f / fmaxf(sqrtf(ss), aux::kNormEps);
  }
}

// Copy up to four local f32 vectors into every rank's buffer (statistics exchange after the forward sweep: this rank's
// partial column (shift, sum) pairs into its slot, its complete row pairs into the gathered row statistics).
struct PushSegs {
  const float* src[4];
  int64_t n[4];
  int64_t dst_off[4];   // bytes from the buffer base
  int n_seg;
};
__global__ void push_f32(PushSegs s, Peers peers, int world, int rank) {
  const int seg = blockIdx.y;
  if (seg >= s.n_seg) return;
  const float* src = s.src[seg];
  const int64_t n = s.n[seg];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = src[i];
    for (int k = 1; k <= world; ++k) {
      int r = rank + k;
      r = r >= world ? r - world : r;
      at<float>(peers.base[r], s.dst_off[seg])[i] = v;
    }
  }
}

// All-reduce (sum) of up to MAX_SCALARS floats in ONE kernel: push the values into every peer's slot, barrier, sum the
// slots in rank order (every rank forms bit-identical results).  Every phase has its own slots, alternating with the
// phase's epoch parity: the kernel may be called back to back, and a fast rank that already entered the NEXT phase's sum
// cannot overwrite values a slow peer has not read yet.
__global__ void sum_scalars(const float* __restrict__ vals, int cnt, Peers peers, int world, int rank, int phase,
                            unsigned long long timeout_ns, float* __restrict__ out) {
  __shared__ uint32_t epoch_sh;
  __shared__ uint32_t next_sh;
  char* mine = reinterpret_cast<char*>(peers.base[rank]);
  if (threadIdx.x == 0) next_sh = at<uint32_t>(mine, OFF_EPOCH)[phase] + 1u;
  __syncthreads();
  const int par = (int)(next_sh & 1u);
  const int t = threadIdx.x;
  if (t < world * cnt) {
    const int r = t / cnt, c = t % cnt;
    st_relaxed_sys_f32(at<float>(peers.base[r], OFF_SCALARS) + ((phase * 2 + par) * MAX_WORLD + rank) * MAX_SCALARS + c, vals[c]);
  }
  __threadfence_system();
  __syncthreads();
  barrier_arrive_wait(peers, world, rank, phase, timeout_ns, &epoch_sh);
  if (t < cnt) {
    double acc = 0.0;
    for (int q = 0; q < world; ++q)
      acc += (double)ld_relaxed_sys_f32(at<float>(mine, OFF_SCALARS) + ((phase * 2 + par) * MAX_WORLD + q) * MAX_SCALARS + t);
    out[t] = (float)acc;
  }
}

}  // namespace link
