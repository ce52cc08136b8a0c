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
// reference's.  Two source points ride in one 64-bit register, so those eight
// subtractions and squares issue as packed add/mul.rn.f32x2 (the targets sit in
// shared memory already duplicated as (t, t) pairs; a - s is evaluated as
// a + (-s), which is the same IEEE result).  torch.argmin runs on sqrt(d2) and returns the FIRST minimum; two
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
static_assert(kSrcPerThread % 2 == 0, "sources are processed in packed pairs");

typedef unsigned long long u64;
__device__ __forceinline__ u64 nn_pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void nn_unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 nn_add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 nn_mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

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
  __shared__ float2 tx[kTgtTile], ty[kTgtTile], tz[kTgtTile];   // every target as (t, t)
  const int tid = threadIdx.x;
  u64 nsx[kSrcPerThread / 2], nsy[kSrcPerThread / 2], nsz[kSrcPerThread / 2];   // NEGATED sources
  float lo[kSrcPerThread], best[kSrcPerThread];   // class floor of the best d2, best distance
  int arg[kSrcPerThread];
#pragma unroll
  for (int k = 0; k < kSrcPerThread; k += 2) {
    float c[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int n = (blockIdx.x * kSrcPerThread + k + h) * kNNThreads + tid;
      const bool ok = n < N;
      c[h][0] = ok ? -__ldg(src + 3 * (size_t)n) : 0.f;
      c[h][1] = ok ? -__ldg(src + 3 * (size_t)n + 1) : 0.f;
      c[h][2] = ok ? -__ldg(src + 3 * (size_t)n + 2) : 0.f;
      lo[k + h] = __int_as_float(0x7f800000);   // +inf: the first target always wins
      best[k + h] = __int_as_float(0x7f800000);
      arg[k + h] = 0;
    }
    nsx[k / 2] = nn_pack2(c[0][0], c[1][0]);
    nsy[k / 2] = nn_pack2(c[0][1], c[1][1]);
    nsz[k / 2] = nn_pack2(c[0][2], c[1][2]);
  }
  const int j_begin = blockIdx.y * tgt_per_slice, j_end = min(M, j_begin + tgt_per_slice);
  for (int j0 = j_begin; j0 < j_end; j0 += kTgtTile) {
    const int cnt = min(kTgtTile, j_end - j0);
    __syncthreads();
    for (int i = tid; i < cnt; i += kNNThreads) {
      const float x = __ldg(tgt + 3 * (size_t)(j0 + i)), y = __ldg(tgt + 3 * (size_t)(j0 + i) + 1),
                  z = __ldg(tgt + 3 * (size_t)(j0 + i) + 2);
      tx[i] = make_float2(x, x);
      ty[i] = make_float2(y, y);
      tz[i] = make_float2(z, z);
    }
    __syncthreads();
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
      const u64 ax = *reinterpret_cast<const u64 *>(&tx[i]), ay = *reinterpret_cast<const u64 *>(&ty[i]),
                az = *reinterpret_cast<const u64 *>(&tz[i]);
#pragma unroll
      for (int k = 0; k < kSrcPerThread; k += 2) {
        const u64 dx = nn_add2(ax, nsx[k / 2]), dy = nn_add2(ay, nsy[k / 2]), dz = nn_add2(az, nsz[k / 2]);
        // (the three squares are packed; their sum is taken with scalar __fadd_rn because
        // ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even under -fmad=false)
        float px[2], py[2], pz[2], d2[2];
        nn_unpack2(nn_mul2(dx, dx), px[0], px[1]);
        nn_unpack2(nn_mul2(dy, dy), py[0], py[1]);
        nn_unpack2(nn_mul2(dz, dz), pz[0], pz[1]);
        d2[0] = __fadd_rn(__fadd_rn(px[0], py[0]), pz[0]);
        d2[1] = __fadd_rn(__fadd_rn(px[1], py[1]), pz[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (d2[h] < lo[k + h]) {   // a strictly smaller sqrt: new best (targets arrive in index order)
            best[k + h] = __fsqrt_rn(d2[h]);
            lo[k + h] = sqrt_class_floor(d2[h], best[k + h]);
            arg[k + h] = j0 + i;
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
