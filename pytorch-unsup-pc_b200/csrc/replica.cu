// Replica-aware projection support + point dropout on the device (next row f2).
//
// Reference: models/model_pc_to.py:47-56 (tf_repeat_0), :302-306 (the cloud of
// every sample is materialised step_size x num_candidates times before the
// projection), :254-258 + util/point_cloud_to.py:269-295 (pc_point_dropout: a
// host-side numpy sampler, np.random.choice(N, M, replace=False) per replica,
// then an advanced-indexing gather) and the autograd of both (index backward +
// the sum over replicas of `repeat`).
//
// Here the projection kernels read the UN-replicated cloud through
// PoseArgs::{replicas, sel, N_src} (common.cuh point_offset), so neither the
// replicated nor the dropped-out cloud ever exists in memory; this file holds
// what is left:
//   dropout_select_kernel   the sampler: a uniformly random M-subset per replica
//   select_points_kernel    the stand-alone gather (pc_point_dropout as an op)
//   invert_selection_kernel sel -> inverse map (cloud point -> slot or -1)
//   replica_reduce_kernel   per-replica point gradients -> gradient of the
//                           cloud, summed over the replicas in replica order
//                           (deterministic; no atomics)
#include "common.cuh"

namespace dpc {

// ---- sampler -----------------------------------------------------------------
// Every (replica, point) pair gets a 32-bit key from a counter-based hash of
// (seed, replica, point); the M points with the smallest keys are the sample
// (ties broken by point index).  Sorting i.i.d. keys gives a uniform random
// permutation, so its first M elements are a uniform M-subset -- the same
// distribution as np.random.choice(N, M, replace=False), whose ORDER inside the
// subset is irrelevant to the projection (a sum over points).  The M-th
// smallest key is found by a 4-pass radix select on shared-memory histograms
// (keys are recomputed, never stored); the sample is emitted in ascending point
// order with a block scan.  One CTA per replica, no global atomics.
constexpr int kSelThreads = 256;

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x7feb352du;
  h ^= h >> 15; h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}
__device__ __forceinline__ uint32_t dropout_key(uint32_t s0, uint32_t s1, uint32_t b, uint32_t n) {
  return mix32(mix32(n ^ s0) + mix32(b * 0x9E3779B9u + s1));
}

__global__ void __launch_bounds__(kSelThreads)
dropout_select_kernel(int N_src, int M, uint32_t s0, uint32_t s1, int *__restrict__ sel) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_need;
  __shared__ unsigned warp_tot[kSelThreads / 32];
  __shared__ unsigned run_lt, run_eq;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // radix select: after the pass on byte k, s_prefix holds the top bytes of the
  // M-th smallest key and s_need how many keys with that prefix are still wanted
  if (tid == 0) { s_prefix = 0; s_need = (unsigned)M; }
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    __syncthreads();
    const unsigned prefix = s_prefix, need = s_need;
    const unsigned himask = shift == 24 ? 0u : ~0u << (shift + 8);
    for (int n = tid; n < N_src; n += kSelThreads) {
      const uint32_t k = dropout_key(s0, s1, b, n);
      if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {
      // the bin in which the running count reaches `need`: 8 bins per lane
      unsigned v[8], sum = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) { v[i] = hist[lane * 8 + i]; sum += v[i]; }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      unsigned run = incl - sum;
      if (run < need && need <= incl) {          // exactly one lane
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (run < need && need <= run + v[i]) {
            s_prefix = prefix | ((unsigned)(lane * 8 + i) << shift);
            s_need = need - run;
          }
          run += v[i];
        }
      }
    }
    __syncthreads();
  }
  const unsigned T = s_prefix, need_eq = s_need;   // take keys < T and the first need_eq keys == T
  if (tid == 0) { run_lt = 0; run_eq = 0; }
  __syncthreads();
  int *out = sel + (size_t)b * M;
  for (int base = 0; base < N_src; base += kSelThreads) {
    const int n = base + tid;
    unsigned lt = 0, eq = 0;
    if (n < N_src) {
      const uint32_t k = dropout_key(s0, s1, b, n);
      lt = k < T;
      eq = k == T;
    }
    // block-exclusive scan of (lt, eq) packed into one word (counts <= 256 each)
    unsigned v = lt | (eq << 16), incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    unsigned before = incl - v, tot = 0;
#pragma unroll
    for (int w = 0; w < kSelThreads / 32; ++w) {
      const unsigned t = warp_tot[w];
      if (w < wid) before += t;
      tot += t;
    }
    const unsigned lt_before = run_lt + (before & 0xFFFFu), eq_before = run_eq + (before >> 16);
    if (lt || (eq && eq_before < need_eq))
      out[lt_before + min(eq_before, need_eq)] = n;
    __syncthreads();
    if (tid == 0) { run_lt += tot & 0xFFFFu; run_eq += tot >> 16; }
    __syncthreads();
  }
}

// out[b][m][:] = points[b / R][sel[b][m]][:]
__global__ void __launch_bounds__(256)
select_points_kernel(const float *__restrict__ points, const int *__restrict__ sel, int R,
                     int N_src, int M, int C, float *__restrict__ out) {
  const int b = blockIdx.y, m = blockIdx.x * 256 + threadIdx.x;
  if (m >= M) return;
  const int src = min(max(__ldg(sel + (size_t)b * M + m), 0), N_src - 1);   // see point_offset()
  const float *p = points + ((size_t)(b / R) * N_src + src) * C;
  float *o = out + ((size_t)b * M + m) * C;
  for (int c = 0; c < C; ++c) o[c] = __ldg(p + c);
}

// inv[b][sel[b][m]] = m   (inv pre-filled with -1)
__global__ void __launch_bounds__(256)
invert_selection_kernel(const int *__restrict__ sel, const int *__restrict__ bmap, int N_src, int M,
                        int *__restrict__ inv) {
  const int b = blockIdx.y, m = blockIdx.x * 256 + threadIdx.x;     // b: chain slot
  if (m >= M) return;
  const int src = __ldg(sel + (size_t)(bmap ? __ldg(bmap + b) : b) * M + m);
  if ((unsigned)src < (unsigned)N_src) inv[(size_t)b * N_src + src] = m;   // never outside inv
}

// g_cloud[c][n][:] = sum_{r < R} g_rep[c R + r][slot(c R + r, n)][:]   (slot = n without dropout;
// dropped points contribute nothing).  Replicas are added in index order.
__global__ void __launch_bounds__(256)
replica_reduce_kernel(const float *__restrict__ g_rep, const int *__restrict__ inv, int R,
                      int N_src, int M, int C, float *__restrict__ g_cloud) {
  pdl_wait();             // the per-replica gradients come from the pose adjoint (or inv from the inversion)
  const int c = blockIdx.y, n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N_src) return;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int r = 0; r < R; ++r) {
    const int b = c * R + r;
    const int slot = inv ? ld_dep(inv + (size_t)b * N_src + n) : n;
    if (slot < 0) continue;
    const float *g = g_rep + ((size_t)b * M + slot) * C;
    for (int k = 0; k < C; ++k) acc[k] += ld_dep(g + k);
  }
  float *o = g_cloud + ((size_t)c * N_src + n) * C;
  for (int k = 0; k < C; ++k) o[k] = acc[k];
}

// ---- launchers -----------------------------------------------------------------
int launch_dropout_select(int P, int N_src, int M, uint64_t seed, int *sel, cudaStream_t s) {
  dropout_select_kernel<<<P, kSelThreads, 0, s>>>(N_src, M, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                  sel);
  return check_launch("dropout_select");
}

int launch_select_points(const float *points, const int *sel, int P, int R, int N_src, int M, int C,
                         float *out, cudaStream_t s) {
  select_points_kernel<<<dim3((M + 255) / 256, P), 256, 0, s>>>(points, sel, R, N_src, M, C, out);
  return check_launch("select_points");
}

int launch_replica_reduce(const float *g_rep, const int *sel, int *inv, int P, int R, int N_src,
                          int M, int C, float *g_cloud, cudaStream_t s, const int *bmap) {
  if (sel) {
    if (cudaMemsetAsync(inv, 0xFF, (size_t)P * N_src * sizeof(int), s) != cudaSuccess)
      return check_launch("memset(inv)");
    invert_selection_kernel<<<dim3((M + 255) / 256, P), 256, 0, s>>>(sel, bmap, N_src, M, inv);
    if (int e = check_launch("invert_selection")) return e;
  }
  launch_dep(replica_reduce_kernel, dim3((N_src + 255) / 256, P / R), dim3(256), 0, s, g_rep,
             (const int *)(sel ? inv : nullptr), R, N_src, M, C, g_cloud);
  return check_launch("replica_reduce");
}

}  // namespace dpc
