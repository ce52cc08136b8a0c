// Pose + trilinear scatter (forward) and gather + pose adjoint (backward).
//
// Reference: quaternion.py:110-132, point_cloud_to.py:118-178 (pose),
// point_cloud_to.py:10-87 (8x index_put_ scatter) and their autograd.
//
// Forward: one thread per point.  The pose math, perspective divide, cell
// index and the eight trilinear weights are computed in registers and go
// straight to eight fire-and-forget fp32 reductions (RED.E.ADD.F32) on the
// grid, which sits in L2 (64 projections x 1 MiB = 64 MiB < 126 MB L2):
// there is no compaction, no index tensor and no host sync.  Out-of-frustum
// points are predicated off; out-of-range corners (coordinate exactly +0.5)
// carry weight 0 and are dropped instead of faulting.
//
// Backward: one thread per point re-derives the identical cell (same device
// function, same roundings -- forward and backward can never disagree on the
// cell), gathers the eight grid gradients (atomic-free), applies the pose
// adjoint in fp64 and block-reduces the per-projection quaternion /
// translation / focal partial sums into a fixed-order two-stage reduction, so
// every gradient is deterministic.
#include <cooperative_groups.h>
#include <cstdlib>

#include "common.cuh"
#include "pose.cuh"

namespace dpc {

constexpr int kPoseThreads = 256;

// 1-point-per-thread kernels (forward pose / scatter / stand-alone gather)
static int point_blocks(int N) { return (N + kPoseThreads - 1) / kPoseThreads; }
// the pose adjoint: kBwdThreads threads x kBwdPts points each per CTA, one row of
// 8 fp64 partial sums per CTA
#ifndef DPC_BWD_PT_THREADS
#define DPC_BWD_PT_THREADS 128 // threads per CTA of the pose adjoint (A/B)
#endif
constexpr int kBwdThreads = DPC_BWD_PT_THREADS;
#ifndef DPC_BWD_PTS
#define DPC_BWD_PTS 4          // points per thread of the pose adjoint (A/B)
#endif
constexpr int kBwdPts = DPC_BWD_PTS;
int pose_partial_blocks(int N) { return (N + kBwdThreads * kBwdPts - 1) / (kBwdThreads * kBwdPts); }

// The normalised quaternion of the CTA's projection, computed by ONE thread (a double-precision
// square root and four IEEE divides: ~80 instructions that every thread would otherwise repeat)
// and broadcast through shared memory.  Ends with a barrier: call it before any early return.
__device__ __forceinline__ Quat block_quat(const float *__restrict__ q) {
  __shared__ Quat sq;
  if (threadIdx.x == 0) sq = load_quat(q);
  __syncthreads();
  return sq;
}

template <bool WRITE_TRPC, bool SCATTER>
__global__ void __launch_bounds__(kPoseThreads)
pose_scatter_kernel(PoseArgs a, float *__restrict__ tr_pc, float *__restrict__ grid) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * kPoseThreads + threadIdx.x;
  const Quat q = block_quat(a.quat + 4 * b);
  if (n >= a.N) return;
  const bool has_t = a.trans != nullptr;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  if (has_t) {
    t0 = a.trans[3 * b];
    t1 = a.trans[3 * b + 1];
    t2 = a.trans[3 * b + 2];
  }
  const double f = a.focal ? (double)a.focal[b] : a.focal_const;
  const size_t pi = ((size_t)b * a.N + n) * 3, si = point_offset(a, b, n);
  const float p0 = a.points[si], p1 = a.points[si + 1], p2 = a.points[si + 2];
  const PosePoint pp = pose_point(q, p0, p1, p2, has_t, t0, t1, t2, f, a.cam_dist);
  if (WRITE_TRPC) {
    tr_pc[pi] = (float)pp.u0;
    tr_pc[pi + 1] = (float)pp.u1;
    tr_pc[pi + 2] = (float)pp.u2;
  }
  if (SCATTER) {
    const Cell c = make_cell(pp.u0, pp.u1, pp.u2, a.Vz, a.V);
    if (!c.valid) return;
    float *g = grid + (size_t)b * a.Vz * a.V * a.V;
    const double wz[2] = {1.0 - c.rz, c.rz}, wy[2] = {1.0 - c.ry, c.ry},
                 wx[2] = {1.0 - c.rx, c.rx};
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int z = c.iz + dz, y = c.iy + dy;
        if (z >= a.Vz || y >= a.V) continue;
        const double wzy = wz[dz] * wy[dy];
        float *row = g + ((size_t)z * a.V + y) * a.V;
        atomicAdd(row + c.ix, (float)(wzy * wx[0]));
        if (c.ix + 1 < a.V) atomicAdd(row + c.ix + 1, (float)(wzy * wx[1]));
      }
  }
}

// Pose + cell records (the forward of the plane-local scatter path): tr_pc for
// the caller, {z cell, (iy, ix), fractions} for the blur-XY kernels that build
// each Z-plane in shared memory.  Threads n in [N, Npad) write the padding.
template <bool WRITE_TRPC>
__global__ void __launch_bounds__(kPoseThreads)
pose_cells_kernel(PoseArgs a, float *__restrict__ tr_pc, CellsView cells) {
  pdl_release();          // head of the forward chain: launched without a programmatic edge
  const int b = blockIdx.y;
  const int n = blockIdx.x * kPoseThreads + threadIdx.x;
  const Quat q = block_quat(a.quat + 4 * b);
  if (n >= cells.Npad) return;
  uint8_t *cz = cells.cellz + (size_t)b * cells.Npad;
  if (n >= a.N) {
    cz[n] = (uint8_t)kCellNone;
    return;
  }
  const bool has_t = a.trans != nullptr;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  if (has_t) {
    t0 = a.trans[3 * b];
    t1 = a.trans[3 * b + 1];
    t2 = a.trans[3 * b + 2];
  }
  const double f = a.focal ? (double)a.focal[b] : a.focal_const;
  const size_t pi = ((size_t)b * a.N + n) * 3, si = point_offset(a, b, n);
  const PosePoint pp = pose_point(q, a.points[si], a.points[si + 1], a.points[si + 2], has_t, t0,
                                  t1, t2, f, a.cam_dist);
  if (WRITE_TRPC) {
    tr_pc[pi] = (float)pp.u0;
    tr_pc[pi + 1] = (float)pp.u1;
    tr_pc[pi + 2] = (float)pp.u2;
  }
  const Cell c = make_cell(pp.u0, pp.u1, pp.u2, a.Vz, a.V);
  cz[n] = (uint8_t)(c.valid ? (unsigned)c.iz : kCellNone);
  if (c.valid)
    cells.rec[(size_t)b * a.N + n] =
        make_uint4(((unsigned)n << 16) | ((unsigned)c.iy << 8) | (unsigned)c.ix,
                   __float_as_uint((float)c.rz),
                   __float_as_uint((float)c.ry), __float_as_uint((float)c.rx));
}

// Counting sort of every projection's point records by z cell, one thread-block
// CLUSTER of kBinSplit CTAs per projection.  Each CTA histograms ITS quarter of
// the points in shared memory; after a cluster barrier every CTA reads the
// other CTAs' histograms through distributed shared memory, which gives it both
// the per-cell totals (scanned into the cell boundaries) and the number of
// same-cell points owned by lower-ranked CTAs (its own starting cursor inside
// each cell); it then places its points' records.  No global atomics, no
// second kernel, 4x fewer shared-memory atomics per CTA than one CTA per
// projection.  The order inside a cell is arbitrary (the consumers accumulate
// with atomics or write per-point results); the cell boundaries are exact.
constexpr int kBinThreads = 256;
#ifndef DPC_BIN_SPLIT
#define DPC_BIN_SPLIT 4
#endif
constexpr int kBinSplit = DPC_BIN_SPLIT;      // CTAs per projection (one cluster)
constexpr int kMaxBins = 192;
__global__ void __cluster_dims__(kBinSplit, 1, 1) __launch_bounds__(kBinThreads)
bin_points_kernel(CellsView cells, int N, int Vz) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned hist[kMaxBins];      // this CTA's counts, read by the whole cluster
  __shared__ unsigned ahead[kMaxBins];     // same-cell points owned by lower-ranked CTAs
  __shared__ unsigned cursor[kMaxBins + 1];
  const int b = blockIdx.y, tid = threadIdx.x;
  const unsigned rank = cluster.block_rank();
  pdl_wait();             // the z-cell bytes and records come from pose_cells
  pdl_release();
  const int n8 = cells.Npad / 8;                          // 8-byte groups of z-cell bytes
  const int per = (n8 + kBinSplit - 1) / kBinSplit;
  const int i_lo = rank * per, i_hi = min(n8, i_lo + per);
  const uint2 *cz = reinterpret_cast<const uint2 *>(cells.cellz + (size_t)b * cells.Npad);
  const uint4 *rec = cells.rec + (size_t)b * N;
  uint4 *srec = cells.srec + (size_t)b * N;
  for (int i = tid; i < kMaxBins; i += kBinThreads) hist[i] = 0;
  __syncthreads();
  for (int i = i_lo + tid; i < i_hi; i += kBinThreads) {
    const uint2 w = ld_dep(cz + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned c = ((k < 4 ? w.x : w.y) >> (8 * (k % 4))) & 0xFFu;
      if (c != kCellNone) atomicAdd(&hist[c], 1u);
    }
  }
  cluster.sync();
  // cell totals and this CTA's offset inside every cell, from the cluster's histograms
  for (int z = tid; z < kMaxBins; z += kBinThreads) {
    unsigned total = 0, before = 0;
    if (z < Vz) {
#pragma unroll
      for (unsigned r = 0; r < kBinSplit; ++r) {
        const unsigned v = *cluster.map_shared_rank(&hist[z], r);
        total += v;
        before += r < rank ? v : 0u;
      }
    }
    cursor[z] = total;          // scanned below
    ahead[z] = before;          // (hist itself stays untouched: other CTAs are still reading it)
  }
  cluster.sync();               // no CTA may exit while its histogram can still be read
  if (tid < 32) {
    // exclusive scan of the <= 192 totals by one warp, 6 cells per lane
    constexpr int PER = kMaxBins / 32;
    unsigned v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      v[k] = cursor[tid * PER + k];
      sum += v[k];
    }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += t;
    }
    unsigned run = incl - sum;
    uint32_t *bs = cells.binstart + (size_t)b * cells.zstride;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int z = tid * PER + k;
      if (rank == 0 && z <= Vz) bs[z] = run;              // z == Vz: the total
      cursor[z] = run + ahead[z];
      run += v[k];
    }
    if (rank == 0 && tid == 31 && Vz == kMaxBins) bs[Vz] = run;
  }
  __syncthreads();
  for (int i = i_lo + tid; i < i_hi; i += kBinThreads) {
    const uint2 w = ld_dep(cz + i);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned c = ((k < 4 ? w.x : w.y) >> (8 * (k % 4))) & 0xFFu;
      if (c != kCellNone) srec[atomicAdd(&cursor[c], 1u)] = ld_dep(rec + 8 * i + k);
    }
  }
}

// Pose + cell records + counting sort in ONE launch (the default forward point kernel).
// pose_cells_kernel -> bin_points_kernel is a chain of two latency-bound launches at the head of
// every forward pass, where nothing else of the step can overlap them: the records travel to
// global memory and back, and the second kernel starts with a cold read of them.  Here the cluster
// that sorts a projection's records also computes them: every thread derives the pose, tr_pc and
// cell of up to kFuseIter points, keeps their records in REGISTERS while the cluster builds and
// exchanges its histograms (same scheme as bin_points_kernel), and then places them.  The
// unsorted record array is never written.  The kernel is a latency chain (points -> fp64 pose
// -> histogram -> two cluster barriers -> placement), so a thread gets as FEW points as a
// portable cluster (<= 8 CTAs of 256 threads) allows: ITER = 2 up to 4096 points, 4 up to 8192,
// 8 up to 16384 (measured at N = 8000, whole batch: the two kernels 21.9 us; 8 points per thread
// in 4 CTAs of 256: 18.4 us; 4 points in 8 CTAs of 256: 15.5 us; 2 points in 8 CTAs of 512:
// 19.4 us); larger clouds take the two kernels above.  Same device functions, same roundings:
// identical records.
#ifndef DPC_FUSE_THREADS
#define DPC_FUSE_THREADS 256   // threads per CTA (512: slower, the barriers grow with the CTA)
#endif
constexpr int kFuseThreads = DPC_FUSE_THREADS;
constexpr int kFuseMaxSplit = 8;               // portable cluster size
template <bool WRITE_TRPC, int ITER>
__global__ void __launch_bounds__(kFuseThreads)
pose_bin_kernel(PoseArgs a, float *__restrict__ tr_pc, CellsView cells) {
  constexpr int kFuseIter = ITER;
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ unsigned hist[kMaxBins];
  __shared__ unsigned ahead[kMaxBins];
  __shared__ unsigned cursor[kMaxBins + 1];
  pdl_release();          // head of the forward chain: launched without a programmatic edge
  const int b = blockIdx.y, tid = threadIdx.x;
  const unsigned rank = cluster.block_rank(), nsplit = cluster.num_blocks();
  for (int i = tid; i < kMaxBins; i += kFuseThreads) hist[i] = 0;
  const Quat q = block_quat(a.quat + 4 * b);     // (ends with a barrier: hist is zero behind it)
  const bool has_t = a.trans != nullptr;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  if (has_t) {
    t0 = a.trans[3 * b];
    t1 = a.trans[3 * b + 1];
    t2 = a.trans[3 * b + 2];
  }
  const double f = a.focal ? (double)a.focal[b] : a.focal_const;
  uint8_t *czg = cells.cellz + (size_t)b * cells.Npad;
  const int lo = (int)rank * (kFuseIter * kFuseThreads);
  uint4 rec[kFuseIter];
  unsigned cz[kFuseIter];
#pragma unroll
  for (int it = 0; it < kFuseIter; ++it) {
    const int n = lo + it * kFuseThreads + tid;
    cz[it] = kCellNone;
    rec[it] = make_uint4(0u, 0u, 0u, 0u);
    if (n < a.N) {
      const size_t pi = ((size_t)b * a.N + n) * 3, si = point_offset(a, b, n);
      const PosePoint pp = pose_point(q, a.points[si], a.points[si + 1], a.points[si + 2], has_t,
                                      t0, t1, t2, f, a.cam_dist);
      if (WRITE_TRPC) {
        st_stream(tr_pc + pi, (float)pp.u0);
        st_stream(tr_pc + pi + 1, (float)pp.u1);
        st_stream(tr_pc + pi + 2, (float)pp.u2);
      }
      const Cell c = make_cell(pp.u0, pp.u1, pp.u2, a.Vz, a.V);
      if (c.valid) {
        cz[it] = (unsigned)c.iz;
        rec[it] = make_uint4(((unsigned)n << 16) | ((unsigned)c.iy << 8) | (unsigned)c.ix,
                             __float_as_uint((float)c.rz), __float_as_uint((float)c.ry),
                             __float_as_uint((float)c.rx));
        atomicAdd(&hist[c.iz], 1u);
      }
    }
    if (n < cells.Npad) czg[n] = (uint8_t)cz[it];     // the backward reads the z cell per point
  }
  cluster.sync();
  for (int z = tid; z < kMaxBins; z += kFuseThreads) {
    unsigned total = 0, before = 0;
    if (z < a.Vz) {
      for (unsigned r = 0; r < nsplit; ++r) {
        const unsigned v = *cluster.map_shared_rank(&hist[z], r);
        total += v;
        before += r < rank ? v : 0u;
      }
    }
    cursor[z] = total;
    ahead[z] = before;
  }
  cluster.sync();               // no CTA may exit while its histogram can still be read
  if (tid < 32) {
    constexpr int PER = kMaxBins / 32;
    unsigned v[PER], sum = 0;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      v[k] = cursor[tid * PER + k];
      sum += v[k];
    }
    unsigned incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += t;
    }
    unsigned run = incl - sum;
    uint32_t *bs = cells.binstart + (size_t)b * cells.zstride;
#pragma unroll
    for (int k = 0; k < PER; ++k) {
      const int z = tid * PER + k;
      if (rank == 0 && z <= a.Vz) bs[z] = run;            // z == Vz: the total
      cursor[z] = run + ahead[z];
      run += v[k];
    }
    if (rank == 0 && tid == 31 && a.Vz == kMaxBins) bs[a.Vz] = run;
  }
  __syncthreads();
  uint4 *srec = cells.srec + (size_t)b * a.N;
#pragma unroll
  for (int it = 0; it < kFuseIter; ++it)
    if (cz[it] != kCellNone) srec[atomicAdd(&cursor[cz[it]], 1u)] = rec[it];
}

__global__ void __launch_bounds__(kPoseThreads)
scatter_trpc_kernel(const float *__restrict__ tr_pc, int N, int Vz, int V,
                    float *__restrict__ grid) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * kPoseThreads + threadIdx.x;
  if (n >= N) return;
  const size_t pi = ((size_t)b * N + n) * 3;
  const Cell c = make_cell((double)tr_pc[pi], (double)tr_pc[pi + 1], (double)tr_pc[pi + 2], Vz, V);
  if (!c.valid) return;
  float *g = grid + (size_t)b * Vz * V * V;
  const double wz[2] = {1.0 - c.rz, c.rz}, wy[2] = {1.0 - c.ry, c.ry}, wx[2] = {1.0 - c.rx, c.rx};
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int z = c.iz + dz, y = c.iy + dy;
      if (z >= Vz || y >= V) continue;
      const double wzy = wz[dz] * wy[dy];
      float *row = g + ((size_t)z * V + y) * V;
      atomicAdd(row + c.ix, (float)(wzy * wx[0]));
      if (c.ix + 1 < V) atomicAdd(row + c.ix + 1, (float)(wzy * wx[1]));
    }
}

// dL/du from the grid gradient at the eight corners (adjoint of the trilinear
// weights): dL/dr_a = sum_corners G[corner] * (+-1) * prod of
// the other two axes' weights; dL/du_a = (V_a - 1) dL/dr_a.
__device__ __forceinline__ void gather_cell(const Cell &c, const float *__restrict__ g, int Vz,
                                            int V, double &gz, double &gy, double &gx) {
  double G[2][2][2];
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const int z = c.iz + dz, y = c.iy + dy;
      const bool ok = (z < Vz) && (y < V);
      const float *row = g + ((size_t)(ok ? z : 0) * V + (ok ? y : 0)) * V;
      G[dz][dy][0] = ok ? (double)ld_dep(row + c.ix) : 0.0;
      G[dz][dy][1] = (ok && c.ix + 1 < V) ? (double)ld_dep(row + c.ix + 1) : 0.0;
    }
  const double wz[2] = {1.0 - c.rz, c.rz}, wy[2] = {1.0 - c.ry, c.ry}, wx[2] = {1.0 - c.rx, c.rx};
  gz = gy = gx = 0.0;
#pragma unroll
  for (int d1 = 0; d1 < 2; ++d1)
#pragma unroll
    for (int d2 = 0; d2 < 2; ++d2) {
      gz += wy[d1] * wx[d2] * (G[1][d1][d2] - G[0][d1][d2]);
      gy += wz[d1] * wx[d2] * (G[d1][1][d2] - G[d1][0][d2]);
      gx += wz[d1] * wy[d2] * (G[d1][d2][1] - G[d1][d2][0]);
    }
  gz *= (double)(Vz - 1);
  gy *= (double)(V - 1);
  gx *= (double)(V - 1);
}

__global__ void __launch_bounds__(kPoseThreads)
gather_trpc_bwd_kernel(const float *__restrict__ tr_pc, int N, int Vz, int V,
                       const float *__restrict__ g_grid, float *__restrict__ g_trpc) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * kPoseThreads + threadIdx.x;
  if (n >= N) return;
  const size_t pi = ((size_t)b * N + n) * 3;
  const Cell c = make_cell((double)tr_pc[pi], (double)tr_pc[pi + 1], (double)tr_pc[pi + 2], Vz, V);
  double gz = 0, gy = 0, gx = 0;
  if (c.valid) gather_cell(c, g_grid + (size_t)b * Vz * V * V, Vz, V, gz, gy, gx);
  g_trpc[pi] = (float)gz;
  g_trpc[pi + 1] = (float)gy;
  g_trpc[pi + 2] = (float)gx;
}

// Final reduction for one projection by one warp: lanes stride over the
// per-block partials, a fixed shuffle tree combines them (same order every
// run), and lane 0 applies the quaternion-normalisation Jacobian
//   dL/dq = (dL/dq^ - q^ (q^ . dL/dq^)) / |q|         (quaternion.py:119-121)
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct FinalizeArgs {
  const double *pose_partials;   // [P][pose_blocks][8] or NULL
  const float *scale_partials;   // [P][scale_blocks]   or NULL
  int pose_blocks, scale_blocks;
  float *g_quat, *g_trans, *g_focal, *g_scale;
  // winner-only backward (render_loss): projection b is the winner among the `cands` candidates
  // of its view; its finalize also writes the exact zeros of the losers' pose gradients.  0: off
  int cands = 0;
};

// (bj: the chain slot whose partials belong to projection b; bj == b outside render_loss)
__device__ __forceinline__ void finalize_projection(const PoseArgs &a, const FinalizeArgs &f,
                                                    int b, int bj, int lane) {
  if (f.pose_partials) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = lane; k < f.pose_blocks; k += 32) {
      const double *p = f.pose_partials + ((size_t)bj * f.pose_blocks + k) * 8;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += __ldcg(p + i);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
    if (lane == 0) {
      if (f.g_quat) {
        const Quat q = load_quat(a.quat + 4 * b);
        const double dot = q.w * acc[0] + q.x * acc[1] + q.y * acc[2] + q.z * acc[3];
        f.g_quat[4 * b] = (float)((acc[0] - q.w * dot) * q.inv_norm);
        f.g_quat[4 * b + 1] = (float)((acc[1] - q.x * dot) * q.inv_norm);
        f.g_quat[4 * b + 2] = (float)((acc[2] - q.y * dot) * q.inv_norm);
        f.g_quat[4 * b + 3] = (float)((acc[3] - q.z * dot) * q.inv_norm);
      }
      if (f.g_trans) {
        f.g_trans[3 * b] = (float)acc[4];
        f.g_trans[3 * b + 1] = (float)acc[5];
        f.g_trans[3 * b + 2] = (float)acc[6];
      }
      if (f.g_focal) f.g_focal[b] = (float)acc[7];
    }
  }
  if (f.scale_partials && f.g_scale) {
    double v = 0;
    for (int k = lane; k < f.scale_blocks; k += 32)
      v += (double)__ldcg(f.scale_partials + (size_t)bj * f.scale_blocks + k);
    v = warp_sum(v);
    if (lane == 0) f.g_scale[b] = (float)v;
  }
  if (f.cands > 1 && lane < f.cands) {
    const int o = (b / f.cands) * f.cands + lane;      // a candidate of b's view
    if (o != b) {
      if (f.g_quat) *reinterpret_cast<float4 *>(f.g_quat + 4 * o) = make_float4(0.f, 0.f, 0.f, 0.f);
      if (f.g_trans) f.g_trans[3 * o] = f.g_trans[3 * o + 1] = f.g_trans[3 * o + 2] = 0.f;
      if (f.g_focal) f.g_focal[o] = 0.f;
      if (f.g_scale) f.g_scale[o] = 0.f;
    }
  }
}

__global__ void finalize_kernel(PoseArgs a, FinalizeArgs f) {
  finalize_projection(a, f, blockIdx.x, blockIdx.x, threadIdx.x);
}

// Sum of 8 per-lane fp64 values over the warp with 9 exchanges instead of 40: a
// butterfly that HALVES the number of live values at each of the first three
// steps (lanes whose bit is clear keep the lower half of the values, the others
// the upper half), then two plain steps on the one value left.  Lane L ends up
// with the warp total of value ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1); the
// combination order is fixed, so the result is reproducible.
__device__ __forceinline__ double warp_sum8(const double (&v)[8], int lane) {
  double a[4];
  const bool h4 = lane & 16;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double send = h4 ? v[i] : v[i + 4], keep = h4 ? v[i + 4] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  double c[2];
  const bool h3 = lane & 8;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const double send = h3 ? a[i] : a[i + 2], keep = h3 ? a[i + 2] : a[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  const bool h2 = lane & 4;
  double r = (h2 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, h2 ? c[0] : c[1], 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// partials[b][block][8] = {dq^_w, dq^_x, dq^_y, dq^_z, dt0, dt1, dt2, df}
__global__ void __launch_bounds__(kBwdThreads)
gather_pose_bwd_kernel(PoseArgs a, const float *__restrict__ g_grid,
                       const float *__restrict__ g_trpc, float *__restrict__ g_points,
                       double *__restrict__ partials, int *__restrict__ counters, FinalizeArgs fin,
                       CellsView cells, const float4 *__restrict__ part) {
  pdl_wait();             // the per-plane partial gathers come from the blur-XY adjoint
  pdl_release();
  // chain slot bj works on projection b (winner-only backward: b = bmap[bj]); inputs and the
  // saved records belong to b, the partial gathers, the point gradients and the partials to bj
  const int bj = blockIdx.y;
  const int b = a.bmap ? __ldg(a.bmap + bj) : bj;
  const Quat q = block_quat(a.quat + 4 * b);
  const bool has_t = a.trans != nullptr;
  float t0 = 0.f, t1 = 0.f, t2 = 0.f;
  if (has_t) {
    t0 = a.trans[3 * b];
    t1 = a.trans[3 * b + 1];
    t2 = a.trans[3 * b + 2];
  }
  const double f = a.focal ? (double)a.focal[b] : a.focal_const;
  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (part) {
    // Plane-local path: the cell comes from the saved records, so nothing here
    // needs the reference's exact fp64 rounding any more -- the adjoint is smooth
    // in p', zc and q, and fp32 keeps it ~1e-6 relative (tolerance 1e-4).  B200
    // issues fp64 at 1/8 of the fp32 rate; in fp64 this kernel was bound by it.
    const float w = q.w, vx = q.x, vy = q.y, vz = q.z, ff = (float)f, cd = (float)a.cam_dist;
    const float ww = w * w - (vx * vx + vy * vy + vz * vz);
    float fa[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int it = 0; it < kBwdPts; ++it) {
      const int n = (blockIdx.x * kBwdPts + it) * kBwdThreads + threadIdx.x;
      if (n >= a.N) continue;
      const size_t pi = ((size_t)bj * a.N + n) * 3, si = point_offset(a, b, n);
      const float p0 = a.points[si], p1 = a.points[si + 1], p2 = a.points[si + 2];
      // p' = q^ (0,p) q^* (+ t), zc = p'0 + camera distance
      const float aw = -(vx * p0 + vy * p1 + vz * p2);
      const float ax = w * p0 + vy * p2 - vz * p1;
      const float ay = w * p1 + vz * p0 - vx * p2;
      const float az = w * p2 + vx * p1 - vy * p0;
      const float r0 = -aw * vx + ax * w - ay * vz + az * vy + t0;
      const float r1 = -aw * vy + ay * w - az * vx + ax * vz + t1;
      const float r2 = -aw * vz + az * w - ax * vy + ay * vx + t2;
      // the blur-XY adjoint already gathered the corners plane by plane:
      // part[0] = plane iz, part[1] = plane iz + 1 (absent when iz + 1 == Vz)
      // (both partials are loaded unconditionally, next to the z-cell byte, and
      // SELECTED afterwards: an unwritten slot may hold anything, NaN included)
      const size_t idx = (size_t)bj * a.N + n;
      const unsigned iz = cells.cellz[(size_t)b * cells.Npad + n];
      const float4 s0 = ld_stream(part + idx), s1 = ld_stream(part + (size_t)a.P * a.N + idx);   // last reader
      const bool in0 = iz != kCellNone, in1 = in0 && (int)iz + 1 < a.Vz;
      float gu0 = ((in0 ? s0.x : 0.f) + (in1 ? s1.x : 0.f)) * (float)(a.Vz - 1);
      float gu1 = ((in0 ? s0.y : 0.f) + (in1 ? s1.y : 0.f)) * (float)(a.V - 1);
      float gu2 = ((in0 ? s0.z : 0.f) + (in1 ? s1.z : 0.f)) * (float)(a.V - 1);
      if (g_trpc) {
        gu0 += g_trpc[pi];
        gu1 += g_trpc[pi + 1];
        gu2 += g_trpc[pi + 2];
      }
      // perspective adjoint
      const float izc = 1.f / (r0 + cd);
      const float s12 = (r1 * gu1 + r2 * gu2) * izc;
      const float g0 = gu0 - ff * s12 * izc, g1 = ff * gu1 * izc, g2 = ff * gu2 * izc;
      // rotation adjoint for F(q^) = (w^2-|v|^2) p + 2 (v.p) v + 2 w (v x p)
      const float vg = vx * g0 + vy * g1 + vz * g2;
      const float vp = vx * p0 + vy * p1 + vz * p2;
      const float gp = g0 * p0 + g1 * p1 + g2 * p2;
      const float c0 = vy * g2 - vz * g1, c1 = vz * g0 - vx * g2, c2 = vx * g1 - vy * g0;   // v x g
      st_stream(g_points + pi, ww * g0 + 2.f * vg * vx - 2.f * w * c0);
      st_stream(g_points + pi + 1, ww * g1 + 2.f * vg * vy - 2.f * w * c1);
      st_stream(g_points + pi + 2, ww * g2 + 2.f * vg * vz - 2.f * w * c2);
      const float x0 = p1 * g2 - p2 * g1, x1 = p2 * g0 - p0 * g2, x2 = p0 * g1 - p1 * g0;   // p x g
      fa[0] += 2.f * w * gp + 2.f * (vx * x0 + vy * x1 + vz * x2);
      fa[1] += -2.f * gp * vx + 2.f * (vg * p0 + vp * g0) + 2.f * w * x0;
      fa[2] += -2.f * gp * vy + 2.f * (vg * p1 + vp * g1) + 2.f * w * x1;
      fa[3] += -2.f * gp * vz + 2.f * (vg * p2 + vp * g2) + 2.f * w * x2;
      fa[4] += g0 - gu0;
      fa[5] += g1;
      fa[6] += g2;
      fa[7] += s12;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (double)fa[i];   // cross-thread sums stay in fp64
  } else {
  const double w = q.w, vx = q.x, vy = q.y, vz = q.z;
  const double ww = w * w - (vx * vx + vy * vy + vz * vz);
  // kBwdPts points per thread, a CTA-wide stride apart (coalesced); the thread's
  // partial sums stay in registers, so the reductions below run once per 4 points
#pragma unroll
  for (int it = 0; it < kBwdPts; ++it) {
    const int n = (blockIdx.x * kBwdPts + it) * kBwdThreads + threadIdx.x;
    if (n >= a.N) continue;
    const size_t pi = ((size_t)bj * a.N + n) * 3, si = point_offset(a, b, n);
    const double p0 = a.points[si], p1 = a.points[si + 1], p2 = a.points[si + 2];
    const PosePoint pp = pose_point(q, (float)p0, (float)p1, (float)p2, has_t, t0, t1, t2, f,
                                    a.cam_dist);
    double gu0 = 0, gu1 = 0, gu2 = 0;
    if (g_grid) {
      const Cell c = make_cell(pp.u0, pp.u1, pp.u2, a.Vz, a.V);
      if (c.valid) gather_cell(c, g_grid + (size_t)b * a.Vz * a.V * a.V, a.Vz, a.V, gu0, gu1, gu2);
    }
    if (g_trpc) {
      gu0 += (double)g_trpc[pi];
      gu1 += (double)g_trpc[pi + 1];
      gu2 += (double)g_trpc[pi + 2];
    }
    // perspective adjoint
    const double izc = 1.0 / pp.zc;
    const double s12 = (pp.r1 * gu1 + pp.r2 * gu2) * izc;  // (p'1 g1 + p'2 g2)/zc
    const double g0 = gu0 - f * s12 * izc;
    const double g1 = f * gu1 * izc;
    const double g2 = f * gu2 * izc;
    // rotation adjoint for F(q^) = (w^2-|v|^2) p + 2 (v.p) v + 2 w (v x p)
    const double vg = vx * g0 + vy * g1 + vz * g2;
    const double vp = vx * p0 + vy * p1 + vz * p2;
    const double gp = g0 * p0 + g1 * p1 + g2 * p2;
    // v x g
    const double c0 = vy * g2 - vz * g1, c1 = vz * g0 - vx * g2, c2 = vx * g1 - vy * g0;
    g_points[pi] = (float)(ww * g0 + 2.0 * vg * vx - 2.0 * w * c0);
    g_points[pi + 1] = (float)(ww * g1 + 2.0 * vg * vy - 2.0 * w * c1);
    g_points[pi + 2] = (float)(ww * g2 + 2.0 * vg * vz - 2.0 * w * c2);
    // p x g
    const double x0 = p1 * g2 - p2 * g1, x1 = p2 * g0 - p0 * g2, x2 = p0 * g1 - p1 * g0;
    acc[0] += 2.0 * w * gp + 2.0 * (vx * x0 + vy * x1 + vz * x2);  // g.(v x p) = v.(p x g)
    acc[1] += -2.0 * gp * vx + 2.0 * (vg * p0 + vp * g0) + 2.0 * w * x0;
    acc[2] += -2.0 * gp * vy + 2.0 * (vg * p1 + vp * g1) + 2.0 * w * x1;
    acc[3] += -2.0 * gp * vz + 2.0 * (vg * p2 + vp * g2) + 2.0 * w * x2;
    acc[4] += g0 - gu0;  // dL/dt0 = dL/dp'0 - g_u0
    acc[5] += g1;
    acc[6] += g2;
    acc[7] += s12;       // dL/df
  }
  }
  // fixed-order block reduction: one butterfly per warp, then warps in index order
  __shared__ double red[kBwdThreads / 32][8];
  const int lane = threadIdx.x & 31;
  const double tot = warp_sum8(acc, lane);
  if ((lane & 3) == 0) red[threadIdx.x >> 5][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = tot;
  __syncthreads();
  if (threadIdx.x < 8) {
    double v = 0;
#pragma unroll
    for (int wdx = 0; wdx < kBwdThreads / 32; ++wdx) v += red[wdx][threadIdx.x];
    partials[((size_t)bj * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = v;
    // publish the row before the arrival counter is bumped; only the writers fence
    // (a CTA-wide fence also waits for every warp's g_points stores: 25 % of this kernel)
    if (counters) __threadfence();
  }
  // Fused finalize: the last block of a projection to arrive reduces all of the
  // projection's partials (in block-index order, so the result does not depend
  // on which block happens to be last).  counters[] is zeroed by the DRC
  // backward kernel earlier in the same pass.
  if (counters) {
    __shared__ int is_last;
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(counters + bj, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
      __threadfence();
      finalize_projection(a, fin, b, bj, threadIdx.x);
    }
  }
}

// ---- launchers ---------------------------------------------------------------
int launch_pose_scatter(const PoseArgs &a, float *tr_pc, float *grid, cudaStream_t s) {
  dim3 g(point_blocks(a.N), a.P), t(kPoseThreads);
  if (tr_pc && grid)
    pose_scatter_kernel<true, true><<<g, t, 0, s>>>(a, tr_pc, grid);
  else if (grid)
    pose_scatter_kernel<false, true><<<g, t, 0, s>>>(a, tr_pc, grid);
  else
    pose_scatter_kernel<true, false><<<g, t, 0, s>>>(a, tr_pc, grid);
  return check_launch("pose_scatter");
}

int launch_scatter_trpc(const float *tr_pc, int P, int N, int Vz, int V, float *grid,
                        cudaStream_t s) {
  dim3 g(point_blocks(N), P), t(kPoseThreads);
  scatter_trpc_kernel<<<g, t, 0, s>>>(tr_pc, N, Vz, V, grid);
  return check_launch("scatter_trpc");
}

int launch_gather_pose_bwd(const PoseArgs &a, const float *g_grid, const float *g_trpc,
                           float *g_points, double *partials, cudaStream_t s) {
  dim3 g(pose_partial_blocks(a.N), a.P), t(kBwdThreads);
  gather_pose_bwd_kernel<<<g, t, 0, s>>>(a, g_grid, g_trpc, g_points, partials, nullptr,
                                         FinalizeArgs{}, CellsView{nullptr, nullptr, nullptr, nullptr, 0, 0}, nullptr);
  return check_launch("gather_pose_bwd");
}

int launch_gather_pose_finalize(const PoseArgs &a, const float *g_grid, const float *g_trpc,
                                float *g_points, double *partials, int *counters,
                                const float *scale_partials, int scale_blocks, float *g_quat,
                                float *g_trans, float *g_focal, float *g_scale, cudaStream_t s) {
  dim3 g(pose_partial_blocks(a.N), a.P), t(kBwdThreads);
  FinalizeArgs f{partials, scale_partials, pose_partial_blocks(a.N), scale_blocks,
                 g_quat, g_trans, g_focal, g_scale};
  gather_pose_bwd_kernel<<<g, t, 0, s>>>(a, g_grid, g_trpc, g_points, partials, counters, f,
                                         CellsView{nullptr, nullptr, nullptr, nullptr, 0, 0}, nullptr);
  return check_launch("gather_pose_finalize");
}

int launch_pose_bwd_partials(const PoseArgs &a, const CellsView &cells, const float4 *part,
                             const float *g_trpc, float *g_points, double *partials,
                             int *counters, const float *scale_partials, int scale_blocks,
                             float *g_quat, float *g_trans, float *g_focal, float *g_scale,
                             cudaStream_t s, int winner_of_cands) {
  dim3 g(pose_partial_blocks(a.N), a.P), t(kBwdThreads);
  FinalizeArgs f{partials, scale_partials, pose_partial_blocks(a.N), scale_blocks,
                 g_quat, g_trans, g_focal, g_scale, winner_of_cands};
  if (winner_of_cands > 32) {
    set_error("pose_bwd_partials: at most 32 candidates per view");
    return DPC_ERR_ARG;
  }
  launch_dep(gather_pose_bwd_kernel, g, t, 0, s, a, (const float *)nullptr, g_trpc, g_points,
             partials, counters, f, cells, part);
  return check_launch("pose_bwd_partials");
}

static int pose_bin_iter(int N);
int launch_pose_cells(const PoseArgs &a, float *tr_pc, const CellsView &cells, cudaStream_t s) {
  dim3 g((cells.Npad + kPoseThreads - 1) / kPoseThreads, a.P), t(kPoseThreads);
  if (a.Vz > kMaxBins || a.N > 65535 || a.V > 256) {
    set_error("pose_cells: the plane-local path needs vox_size_z <= %d, N <= 65535, vox_size <= 256",
              kMaxBins);
    return DPC_ERR_ARG;
  }
  const int split = pose_bin_split(a.N);
  if (split > 0) {
    // one launch: a cluster of `split` CTAs per projection (runtime cluster dimension)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(split, a.P);
    cfg.blockDim = dim3(kFuseThreads);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = split;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    const int iter = pose_bin_iter(a.N);
#define DPC_POSE_BIN(TR, IT) cudaLaunchKernelEx(&cfg, pose_bin_kernel<TR, IT>, a, tr_pc, cells)
    if (tr_pc) {
      if (iter == 2) DPC_POSE_BIN(true, 2); else if (iter == 4) DPC_POSE_BIN(true, 4); else DPC_POSE_BIN(true, 8);
    } else {
      if (iter == 2) DPC_POSE_BIN(false, 2); else if (iter == 4) DPC_POSE_BIN(false, 4); else DPC_POSE_BIN(false, 8);
    }
#undef DPC_POSE_BIN
    return check_launch("pose_bin");
  }
  if (tr_pc)
    pose_cells_kernel<true><<<g, t, 0, s>>>(a, tr_pc, cells);
  else
    pose_cells_kernel<false><<<g, t, 0, s>>>(a, tr_pc, cells);
  launch_dep(bin_points_kernel, dim3(kBinSplit, a.P), dim3(kBinThreads), 0, s, cells, a.N, a.Vz);
  return check_launch("pose_cells");
}

// CTAs per projection of the fused pose + binning kernel; 0: the cloud is too large for it (or
// DPC_FUSED_BIN=0 in the environment, for A/B runs) and the two-kernel path runs
// points per thread of the fused kernel: the smallest of 2 / 4 / 8 that fits one portable cluster
static int pose_bin_iter(int N) {
  const int npad = cells_npad(N);
  for (int it = 2; it <= 8; it *= 2)
    if (npad <= it * kFuseThreads * kFuseMaxSplit) return it;
  return 0;
}
int pose_bin_split(int N) {
  static const bool off = getenv("DPC_FUSED_BIN") && atoi(getenv("DPC_FUSED_BIN")) == 0;
  const int it = pose_bin_iter(N);
  if (off || it == 0) return 0;
  const int need = (cells_npad(N) + it * kFuseThreads - 1) / (it * kFuseThreads);
  int split = 1;
  while (split < need) split *= 2;
  return split;
}

int launch_gather_trpc_bwd(const float *tr_pc, int P, int N, int Vz, int V, const float *g_grid,
                           float *g_trpc, cudaStream_t s) {
  dim3 g(point_blocks(N), P), t(kPoseThreads);
  gather_trpc_bwd_kernel<<<g, t, 0, s>>>(tr_pc, N, Vz, V, g_grid, g_trpc);
  return check_launch("gather_trpc_bwd");
}

int launch_finalize(const PoseArgs &a, const double *pose_partials, int pose_blocks,
                    const float *scale_partials, int scale_blocks, float *g_quat, float *g_trans,
                    float *g_focal, float *g_scale, cudaStream_t s) {
  FinalizeArgs f{pose_partials, scale_partials, pose_blocks, scale_blocks,
                 g_quat, g_trans, g_focal, g_scale};
  finalize_kernel<<<a.P, 32, 0, s>>>(a, f);
  return check_launch("finalize");
}

}  // namespace dpc
