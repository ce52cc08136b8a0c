// Brute-force nearest neighbour for the Chamfer evaluation.
//
// Reference: util/point_cloud_distance.py:25-40 point_cloud_distance (for every
// source point the closest target point: materialises two [VsN,VtN,3] copies,
// the difference, its square, the [VsN,VtN] distances and an argmin), driven in
// chunks by run/eval_chamfer_to.py:24-44 compute_distance.
//
// Here: a thread owns kSrcPerThread source points in registers; the targets of
// the CTA's slice stream through shared memory (SoA, every lane reads the same
// target: a broadcast, no bank conflicts) and each pair costs 3 subs, 3 muls,
// 2 adds -- rounded separately and in the reference's order ((dx^2 + dy^2) +
// dz^2, no FMA contraction), so the squared distances are bit-identical to the
// reference's.  torch.argmin runs on sqrt(d2) and returns the FIRST minimum; two
// different d2 can round to the same sqrt, so the running best is only replaced
// when d2 falls below the smallest float whose correctly-rounded sqrt equals
// the current best distance.  The target axis is split over blockIdx.y; slices
// merge with a 64-bit atomicMin on (bits(dist) << 32 | index): non-negative
// floats order like their bit patterns and equal distances fall to the lowest
// index -- exactly argmin's rule, independent of scheduling (deterministic).
#include "common.cuh"

namespace dpc {

constexpr int kNNThreads = 128;
constexpr int kSrcPerThread = 4;
constexpr int kTgtTile = 1024;

// smallest float x with sqrt_rn(x) == sqrt_rn(d2)
__device__ __forceinline__ float sqrt_class_floor(float d2, float m) {
  float x = d2;
  while (x > 0.f) {
    const float y = __uint_as_float(__float_as_uint(x) - 1u);   // next float below
    if (__fsqrt_rn(y) != m) break;
    x = y;
  }
  return x;
}

__global__ void __launch_bounds__(kNNThreads)
nn_search_kernel(const float *__restrict__ src, int N, const float *__restrict__ tgt, int M,
                 int tgt_per_slice, unsigned long long *__restrict__ keys) {
  __shared__ float tx[kTgtTile], ty[kTgtTile], tz[kTgtTile];
  const int tid = threadIdx.x;
  float sx[kSrcPerThread], sy[kSrcPerThread], sz[kSrcPerThread];
  float lo[kSrcPerThread], best[kSrcPerThread];   // class floor of the best d2, best distance
  int arg[kSrcPerThread];
#pragma unroll
  for (int k = 0; k < kSrcPerThread; ++k) {
    const int n = (blockIdx.x * kSrcPerThread + k) * kNNThreads + tid;
    const bool ok = n < N;
    sx[k] = ok ? __ldg(src + 3 * (size_t)n) : 0.f;
    sy[k] = ok ? __ldg(src + 3 * (size_t)n + 1) : 0.f;
    sz[k] = ok ? __ldg(src + 3 * (size_t)n + 2) : 0.f;
    lo[k] = __int_as_float(0x7f800000);   // +inf: the first target always wins
    best[k] = __int_as_float(0x7f800000);
    arg[k] = 0;
  }
  const int j_begin = blockIdx.y * tgt_per_slice, j_end = min(M, j_begin + tgt_per_slice);
  for (int j0 = j_begin; j0 < j_end; j0 += kTgtTile) {
    const int cnt = min(kTgtTile, j_end - j0);
    __syncthreads();
    for (int i = tid; i < cnt; i += kNNThreads) {
      tx[i] = __ldg(tgt + 3 * (size_t)(j0 + i));
      ty[i] = __ldg(tgt + 3 * (size_t)(j0 + i) + 1);
      tz[i] = __ldg(tgt + 3 * (size_t)(j0 + i) + 2);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
      const float ax = tx[i], ay = ty[i], az = tz[i];
#pragma unroll
      for (int k = 0; k < kSrcPerThread; ++k) {
        const float dx = __fsub_rn(ax, sx[k]), dy = __fsub_rn(ay, sy[k]), dz = __fsub_rn(az, sz[k]);
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        if (d2 < lo[k]) {   // a strictly smaller sqrt: new best (targets arrive in index order)
          best[k] = __fsqrt_rn(d2);
          lo[k] = sqrt_class_floor(d2, best[k]);
          arg[k] = j0 + i;
        }
      }
    }
  }
  if (j_begin < j_end) {
#pragma unroll
    for (int k = 0; k < kSrcPerThread; ++k) {
      const int n = (blockIdx.x * kSrcPerThread + k) * kNNThreads + tid;
      if (n < N)
        atomicMin(keys + n, ((unsigned long long)__float_as_uint(best[k]) << 32) | (unsigned)arg[k]);
    }
  }
}

__global__ void __launch_bounds__(256)
nn_finalize_kernel(const unsigned long long *__restrict__ keys, const float *__restrict__ tgt,
                   int N, float *__restrict__ proj, float *__restrict__ min_dist,
                   long long *__restrict__ idx) {
  const int n = blockIdx.x * 256 + threadIdx.x;
  if (n >= N) return;
  const unsigned long long k = keys[n];
  const unsigned j = (unsigned)(k & 0xffffffffull);
  min_dist[n] = __uint_as_float((unsigned)(k >> 32));
  idx[n] = (long long)j;
  proj[3 * (size_t)n] = __ldg(tgt + 3 * (size_t)j);
  proj[3 * (size_t)n + 1] = __ldg(tgt + 3 * (size_t)j + 1);
  proj[3 * (size_t)n + 2] = __ldg(tgt + 3 * (size_t)j + 2);
}

int launch_nn_search(const float *src, int N, const float *tgt, int M, unsigned long long *keys,
                     float *proj, float *min_dist, long long *idx, cudaStream_t s) {
  if (cudaMemsetAsync(keys, 0xff, (size_t)N * sizeof(unsigned long long), s) != cudaSuccess)
    return check_launch("nn_search memset");
  const int gx = (N + kNNThreads * kSrcPerThread - 1) / (kNNThreads * kSrcPerThread);
  // enough target slices to fill the GPU (~4 CTAs per SM), each a whole number of tiles
  int slices = (148 * 4 + gx - 1) / gx;
  const int max_slices = (M + kTgtTile - 1) / kTgtTile;
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  int per = (M + slices - 1) / slices;
  per = (per + kTgtTile - 1) / kTgtTile * kTgtTile;
  slices = (M + per - 1) / per;
  nn_search_kernel<<<dim3(gx, slices), kNNThreads, 0, s>>>(src, N, tgt, M, per, keys);
  nn_finalize_kernel<<<(N + 255) / 256, 256, 0, s>>>(keys, tgt, N, proj, min_dist, idx);
  return check_launch("nn_search");
}

}  // namespace dpc
