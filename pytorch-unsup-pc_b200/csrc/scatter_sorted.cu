// Deterministic trilinear scatter: sort-then-segment (bit-exact run to run).
//
// Reference: point_cloud_to.py:10-87.  On CUDA the reference's
// index_put_(accumulate=True) is ATen's sort-based kernel, i.e. deterministic
// but slow; the default path here uses fp32 reductions whose order varies
// between runs.  This mode restores run-to-run bit-exactness without atomics:
//
//  1. sort_points_kernel (one CTA per projection): every point gets the key
//     of its base grid row (iz*V + iy; out-of-frustum points get a sentinel)
//     packed with its index as key<<16 | n, then a stable 4-pass LSD radix
//     sort (4-bit digits) orders the points by row.  Each thread owns a
//     contiguous chunk of the array and private digit counters, so the sort
//     uses no atomics and equal keys keep ascending point index.
//  2. segment_rows_kernel (one thread per output grid row): the row (z, y)
//     receives contributions only from the four base rows (z-dz, y-dy); the
//     thread walks those four sorted segments in a fixed order and sums into
//     a private shared-memory row, which the CTA then stores coalesced.  Every
//     voxel is written exactly once (no memset, no atomics) and its summation
//     order is a pure function of the inputs.
#include "common.cuh"
#include "pose.cuh"

namespace dpc {

constexpr int kSortThreads = 256;
constexpr int kRowThreads = 128;

size_t sorted_workspace_bytes(int P, int N, int Vz, int V) {
  (void)Vz; (void)V;
  return (size_t)2 * P * N * sizeof(uint32_t);
}

struct PointSource {
  PoseArgs pose;          // used when pose.points != nullptr
  const float *tr_pc;     // otherwise
};

__device__ __forceinline__ Cell point_cell(const PointSource &src, int b, int n, int N, int Vz,
                                           int V, float *tr_out) {
  double u0, u1, u2;
  const size_t pi = ((size_t)b * N + n) * 3;
  if (src.pose.points) {
    const PoseArgs &a = src.pose;
    const Quat q = load_quat(a.quat + 4 * b);
    const bool has_t = a.trans != nullptr;
    const float t0 = has_t ? a.trans[3 * b] : 0.f, t1 = has_t ? a.trans[3 * b + 1] : 0.f,
                t2 = has_t ? a.trans[3 * b + 2] : 0.f;
    const double f = a.focal ? (double)a.focal[b] : a.focal_const;
    const size_t si = point_offset(a, b, n);
    const PosePoint pp = pose_point(q, a.points[si], a.points[si + 1], a.points[si + 2], has_t, t0,
                                    t1, t2, f, a.cam_dist);
    u0 = pp.u0; u1 = pp.u1; u2 = pp.u2;
    if (tr_out) {
      tr_out[pi] = (float)u0;
      tr_out[pi + 1] = (float)u1;
      tr_out[pi + 2] = (float)u2;
    }
  } else {
    u0 = src.tr_pc[pi]; u1 = src.tr_pc[pi + 1]; u2 = src.tr_pc[pi + 2];
  }
  return make_cell(u0, u1, u2, Vz, V);
}

__global__ void __launch_bounds__(kSortThreads)
sort_points_kernel(PointSource src, float *__restrict__ tr_out, uint32_t *__restrict__ bufA,
                   uint32_t *__restrict__ bufB, int N, int Vz, int V) {
  __shared__ uint32_t cnt[16 * kSortThreads];
  __shared__ uint32_t warp_tot[kSortThreads / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  uint32_t *A = bufA + (size_t)b * N, *B = bufB + (size_t)b * N;
  const uint32_t invalid_key = (uint32_t)(Vz * V);
  for (int n = tid; n < N; n += kSortThreads) {
    const Cell c = point_cell(src, b, n, N, Vz, V, tr_out);
    const uint32_t key = c.valid ? (uint32_t)(c.iz * V + c.iy) : invalid_key;
    A[n] = (key << 16) | (uint32_t)n;
  }
  __syncthreads();
  const int chunk = (N + kSortThreads - 1) / kSortThreads;
  const int lo = min(tid * chunk, N), hi = min(lo + chunk, N);
  uint32_t *from = A, *to = B;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 16 + 4 * pass;
#pragma unroll
    for (int d = 0; d < 16; ++d) cnt[d * kSortThreads + tid] = 0;
    for (int i = lo; i < hi; ++i) cnt[((from[i] >> shift) & 15u) * kSortThreads + tid]++;
    __syncthreads();
    // exclusive scan of cnt[] in flat (digit-major, thread-minor) order
    uint32_t loc[16], sum = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      loc[i] = cnt[tid * 16 + i];
      sum += loc[i];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    uint32_t base = incl - sum;
    for (int w = 0; w < (tid >> 5); ++w) base += warp_tot[w];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      cnt[tid * 16 + i] = base;
      base += loc[i];
    }
    __syncthreads();
    for (int i = lo; i < hi; ++i) {
      const uint32_t item = from[i];
      to[cnt[((item >> shift) & 15u) * kSortThreads + tid]++] = item;
    }
    __syncthreads();
    uint32_t *t = from; from = to; to = t;
  }
  // 4 passes: the sorted array ends in bufA
}

__device__ __forceinline__ int lower_bound_key(const uint32_t *__restrict__ items, int N,
                                               uint32_t key) {
  int lo = 0, hi = N;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if ((items[mid] >> 16) < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(kRowThreads)
segment_rows_kernel(PointSource src, const uint32_t *__restrict__ sorted, int N, int Vz, int V,
                    float *__restrict__ grid) {
  extern __shared__ float acc[];  // [V][kRowThreads + 1]
  constexpr int S = kRowThreads + 1;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int rows = Vz * V;
  const int row = blockIdx.x * kRowThreads + tid;
  const uint32_t *items = sorted + (size_t)b * N;
  for (int x = 0; x < V; ++x) acc[x * S + tid] = 0.f;
  if (row < rows) {
    const int oz = row / V, oy = row - oz * V;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int bz = oz - dz, by = oy - dy;
        if (bz < 0 || by < 0) continue;
        const uint32_t key = (uint32_t)(bz * V + by);
        const int i0 = lower_bound_key(items, N, key), i1 = lower_bound_key(items, N, key + 1);
        for (int i = i0; i < i1; ++i) {
          const int n = (int)(items[i] & 0xffffu);
          const Cell c = point_cell(src, b, n, N, Vz, V, nullptr);
          const double wzy = (dz ? c.rz : 1.0 - c.rz) * (dy ? c.ry : 1.0 - c.ry);
          acc[c.ix * S + tid] += (float)(wzy * (1.0 - c.rx));
          if (c.ix + 1 < V) acc[(c.ix + 1) * S + tid] += (float)(wzy * c.rx);
        }
      }
  }
  __syncthreads();
  // coalesced store of kRowThreads consecutive rows
  float *g = grid + (size_t)b * rows * V + (size_t)blockIdx.x * kRowThreads * V;
  const int nrow = min(kRowThreads, rows - blockIdx.x * kRowThreads);
  for (int e = tid; e < nrow * V; e += kRowThreads) {
    const int r = e / V, x = e - r * V;
    g[e] = acc[x * S + r];
  }
}

int launch_scatter_sorted(const PoseArgs *a, const float *tr_pc_in, int P, int N, int Vz, int V,
                          float *tr_pc_out, float *grid, void *ws, size_t ws_bytes,
                          cudaStream_t s) {
  if (N > 65536 || Vz * V >= 65535) {
    set_error("sorted scatter: needs N <= 65536 and Vz*V < 65535");
    return DPC_ERR_ARG;
  }
  if (ws_bytes < sorted_workspace_bytes(P, N, Vz, V) || !ws) {
    set_error("sorted scatter: workspace too small");
    return DPC_ERR_WORKSPACE;
  }
  PointSource src;
  if (a) {
    src.pose = *a;
    src.tr_pc = nullptr;
  } else {
    src.pose = PoseArgs{};
    src.pose.points = nullptr;
    src.tr_pc = tr_pc_in;
  }
  uint32_t *bufA = (uint32_t *)ws, *bufB = bufA + (size_t)P * N;
  sort_points_kernel<<<P, kSortThreads, 0, s>>>(src, a ? tr_pc_out : nullptr, bufA, bufB, N, Vz, V);
  if (int e = check_launch("sort_points")) return e;
  const size_t smem = (size_t)V * (kRowThreads + 1) * sizeof(float);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(segment_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         128 * (kRowThreads + 1) * (int)sizeof(float));
  }
  dim3 g((Vz * V + kRowThreads - 1) / kRowThreads, P);
  segment_rows_kernel<<<g, kRowThreads, smem, s>>>(src, bufA, N, Vz, V, grid);
  return check_launch("segment_rows");
}

}  // namespace dpc
