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
//     sort (4-bit digits) orders the points by row -- in shared memory when the
//     cloud fits (<= 12288 points).  Each thread owns a contiguous chunk of the
//     array and private digit counters, so the sort uses no atomics and equal
//     keys keep ascending point index.  The cell record of every point
//     (ix + fp32 fractions) and the start of every row's segment are written
//     once, here.
//  2. segment_rows_kernel (one thread per output grid row): the row (z, y)
//     receives contributions only from the four base rows (z-dz, y-dy); the
//     thread walks those four sorted segments in a fixed order (dz = 0 then 1, inside it row
//     y - 1 then row y: ascending sorted position) -- segment
//     bounds from the table, weights from the records: no pose, no search --
//     and sums into a private shared-memory row, which the CTA then stores
//     coalesced.  Every voxel is written exactly once (no memset, no atomics)
//     and its summation order is a pure function of the inputs.
// Round 2: records + shared-memory sort + segment table (workload C5's scatter
// stage: 241 us -> see DESIGN.md section 5).
//
// Plane-local variant (launch_sort_cells; the whole-projection path when the caller passes a
// cell-record buffer): the first two kernels are the same, but they write the plane-local path's
// saved state -- z-cell bytes, records {n << 16 | iy << 8 | ix, fractions} sorted by row (hence by
// z cell), the z-cell boundaries -- and step 2 happens INSIDE the blur-XY kernel (blur_xy.cu,
// DET): the CTA that owns plane z builds it in shared memory, one thread per row walking the
// same four segments in the same order.  The raw grid never reaches HBM, the sums are the very
// sums of segment_rows_kernel (bit-identical planes), and the backward is the plane-local one.
#include "common.cuh"
#include "pose.cuh"

namespace dpc {

constexpr int kItemThreads = 256;        // sort_items_kernel: one thread per point
constexpr int kSortThreads = 1024;       // sort_points_kernel: one CTA per projection
constexpr int kRowThreads = 128;
constexpr int kSmemSortMax = 12288;     // points per projection the shared-memory sort holds (2 x 48 KB)

// Per projection: items A [N] | items B [N] | records [N] uint4 | sorted records [N] uint4 |
// rowstart [Vz*V + 2], 16-byte aligned
static size_t sorted_stride_bytes(int N, int Vz, int V) {
  const size_t items = (((size_t)2 * N * sizeof(uint32_t)) + 15) & ~(size_t)15;
  const size_t rows = (((size_t)(Vz * V + 2) * sizeof(uint32_t)) + 15) & ~(size_t)15;
  return items + (size_t)2 * N * sizeof(uint4) + rows;
}
size_t sorted_workspace_bytes(int P, int N, int Vz, int V) {
  return (size_t)P * sorted_stride_bytes(N, Vz, V);
}
struct SortedView {
  uint32_t *A, *B;
  uint4 *rec, *srec;       // records in point order / in sorted order
  uint32_t *rowstart;
};
__host__ __device__ inline SortedView sorted_view(void *ws, size_t stride, int b, int N) {
  char *base = (char *)ws + (size_t)b * stride;
  const size_t items = (((size_t)2 * N * sizeof(uint32_t)) + 15) & ~(size_t)15;
  SortedView v;
  v.A = (uint32_t *)base;
  v.B = v.A + N;
  v.rec = (uint4 *)(base + items);
  v.srec = v.rec + N;
  v.rowstart = (uint32_t *)(v.srec + N);
  return v;
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

// One thread per point: the cell of every point is derived ONCE (fp64 pose, as everywhere) and
// kept as a record {ix, rz, ry, rx} (fp32 fractions, as on the default path) next to its sort item
// key << 16 | n (key = base grid row iz * V + iy; Vz * V for out-of-frustum points).
template <bool CELLS>
__global__ void __launch_bounds__(kItemThreads)
sort_items_kernel(PointSource src, float *__restrict__ tr_out, void *ws, size_t stride, int N,
                  int Vz, int V, CellsView cells) {
  const int b = blockIdx.y, n = blockIdx.x * kItemThreads + threadIdx.x;
  if (CELLS && n >= N && n < cells.Npad) cells.cellz[(size_t)b * cells.Npad + n] = (uint8_t)kCellNone;
  if (n >= N) return;
  const SortedView sv = sorted_view(ws, stride, b, N);
  const Cell c = point_cell(src, b, n, N, Vz, V, tr_out);
  const uint32_t key = c.valid ? (uint32_t)(c.iz * V + c.iy) : (uint32_t)(Vz * V);
  sv.A[n] = (key << 16) | (uint32_t)n;
  if (CELLS) {
    // the plane-local path's record (common.cuh CellsView): the point index and the in-plane
    // cell ride in the first word; the z-cell byte tells the backward which partials exist
    cells.cellz[(size_t)b * cells.Npad + n] = (uint8_t)(c.valid ? (unsigned)c.iz : kCellNone);
    cells.rec[(size_t)b * N + n] =
        make_uint4(((uint32_t)n << 16) | (((uint32_t)c.iy & 0xFFu) << 8) | ((uint32_t)c.ix & 0xFFu),
                   __float_as_uint((float)c.rz), __float_as_uint((float)c.ry),
                   __float_as_uint((float)c.rx));
  } else {
    sv.rec[n] = make_uint4((uint32_t)c.ix, __float_as_uint((float)c.rz), __float_as_uint((float)c.ry),
                           __float_as_uint((float)c.rx));
  }
}

// One CTA per projection: the items are sorted in SHARED memory when the cloud fits (SMEM; else in
// the two global buffers), and the start of every grid row's segment is tabulated, so that the
// last kernel neither recomputes a pose nor searches.
//
// Fast path (the row table fits the counter region: bins >= rows + 2): ONE counting pass over the
// full row key -- histogram (shared-memory atomics: counts do not depend on the order), exclusive
// scan (= the row-segment table), placement through per-row cursors (atomics again: the order
// INSIDE a row is arbitrary at this point), then every item is ranked inside its row by counting
// the row's smaller items (a row holds a handful) and its record is written to that place.  The
// result is THE stable sort -- key ascending, point index ascending -- whatever order the atomics
// ran in, so it is the 4-pass radix sort's result bit for bit, at a quarter of its passes.  A row
// that holds more than kRowSortMax points (a degenerate cloud piled into one grid row: the
// ranking is quadratic in the row) sends the projection down the radix path below instead.
#ifndef DPC_SORT_FAST
#define DPC_SORT_FAST 1        // A/B: 0 = always the 4-pass radix sort
#endif
constexpr uint32_t kRowSortMax = 128;
constexpr int kSortBinsMax = 20480;      // counter-region words the fast path may ask for (80 KB)
static int sort_counter_words(int rows) {
  const int need = (rows + 2 + 3) & ~3;
  return (DPC_SORT_FAST && need <= kSortBinsMax && need > 16 * kSortThreads) ? need : 16 * kSortThreads;
}
template <bool SMEM, bool CELLS>
__global__ void __launch_bounds__(kSortThreads)
sort_points_kernel(void *ws, size_t stride, int N, int Vz, int V, CellsView cells, int cnt_words) {
  extern __shared__ uint32_t sm_sort[];           // cnt [cnt_words] | SMEM: A [N] | B [N]
  __shared__ uint32_t warp_tot[kSortThreads / 32];
  uint32_t *cnt = sm_sort, *sm_items = sm_sort + cnt_words;
  const int b = blockIdx.x, tid = threadIdx.x;
  const SortedView sv = sorted_view(ws, stride, b, N);
  uint32_t *A = SMEM ? sm_items : sv.A, *B = SMEM ? sm_items + N : sv.B;
  const uint32_t rows = (uint32_t)(Vz * V);       // the key of out-of-frustum points
  uint32_t *bs = CELLS ? cells.binstart + (size_t)b * cells.zstride : nullptr;
  if (SMEM)
    for (int n = tid; n < N; n += kSortThreads) A[n] = sv.A[n];
  const int nb = (int)rows + 2;                   // bins: rows, the out-of-frustum key, a zero
  if (DPC_SORT_FAST && nb <= cnt_words) {
    uint32_t *bins = cnt;
    for (int k = tid; k < nb; k += kSortThreads) bins[k] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += kSortThreads) atomicAdd(&bins[A[i] >> 16], 1u);
    __syncthreads();
    // exclusive scan: a contiguous chunk of bins per thread, warp scan, warp totals
    const int per = (nb + kSortThreads - 1) / kSortThreads;
    const int k0 = min(tid * per, nb), k1 = min(k0 + per, nb);
    uint32_t sum = 0;
    bool big = false;
    for (int k = k0; k < k1; ++k) {
      const uint32_t v = bins[k];
      sum += v;
      big |= (uint32_t)k < rows && v > kRowSortMax;
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if ((tid & 31) >= o) incl += v;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    if (!__syncthreads_or(big)) {
      uint32_t run = incl - sum;
      for (int w = 0; w < (tid >> 5); ++w) run += warp_tot[w];
      for (int k = k0; k < k1; ++k) {
        const uint32_t v = bins[k];
        bins[k] = run;                              // the row's cursor
        sv.rowstart[k] = run;                       // rowstart[rows] = valid points, [rows + 1] = N
        if (CELLS && (uint32_t)k <= rows && (uint32_t)k % (uint32_t)V == 0) bs[(uint32_t)k / (uint32_t)V] = run;
        run += v;
      }
      __syncthreads();
      for (int i = tid; i < N; i += kSortThreads) {
        const uint32_t item = A[i];
        B[atomicAdd(&bins[item >> 16], 1u)] = item;
      }
      __syncthreads();
      // bins[k] is now the END of row k.  Every item finds its place inside its row by counting
      // the row's smaller items (a row holds a handful; the items are distinct) -- item-parallel,
      // so lanes do not wait for the longest row of the warp as a sort per row made them (ncu:
      // 40 % of the kernel's warp instructions at 7 active lanes) -- and its record goes straight
      // to that place: the record loads, random 16-byte reads from L2, are all requested before
      // the first count.  (The out-of-frustum tail keeps its arbitrary order: nobody reads it.)
      {
        const uint4 *__restrict__ rec = CELLS ? cells.rec + (size_t)b * N : sv.rec;
        uint4 *__restrict__ srec = CELLS ? cells.srec + (size_t)b * N : sv.srec;
        constexpr int G = 4;
        for (int i0 = tid; i0 < N; i0 += G * kSortThreads) {
          uint32_t item[G];
          uint4 r[G];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int i = i0 + g * kSortThreads;
            item[g] = i < N ? B[i] : 0xffffffffu;
            if (i < N) r[g] = __ldcg(rec + (item[g] & 0xffffu));
          }
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const int i = i0 + g * kSortThreads;
            if (i >= N) continue;
            const uint32_t key = item[g] >> 16;
            uint32_t pos = (uint32_t)i;
            if (key < rows) {
              const uint32_t s0 = key ? bins[key - 1] : 0u, s1 = bins[key];
              pos = s0;
              for (uint32_t j = s0; j < s1; ++j) pos += B[j] < item[g] ? 1u : 0u;
            }
            srec[pos] = r[g];
          }
        }
      }
      return;
    }
    __syncthreads();
  } else {
    __syncthreads();
  }
  // ---- the stable 4-pass LSD radix sort (grids whose row table does not fit, piled-up rows) ----
  {
  // contiguous chunk per thread (stable); an ODD chunk length keeps the threads' walks on
  // different shared-memory banks
  int chunk = (N + kSortThreads - 1) / kSortThreads;
  chunk |= 1;
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
  }
  // (radix: 4 passes, the sorted array ends in A.)  Segment starts: rowstart[k] = first index
  // whose key is >= k, for k = 0 .. rows + 1 (rowstart[rows] = the first out-of-frustum point,
  // rowstart[rows + 1] = N); the records go to the global buffer the next kernel reads, in
  // sorted order
  const uint4 *__restrict__ rec = CELLS ? cells.rec + (size_t)b * N : sv.rec;
  uint4 *__restrict__ srec = CELLS ? cells.srec + (size_t)b * N : sv.srec;
  for (int i = tid; i < N; i += kSortThreads) {
    const uint32_t item = A[i], key = item >> 16;
    srec[i] = rec[item & 0xffffu];                // sequential reads later
    const uint32_t prev = i > 0 ? (A[i - 1] >> 16) + 1u : 0u;
    for (uint32_t k = prev; k <= key; ++k) {
      sv.rowstart[k] = (uint32_t)i;
      // rows are z-major: the first row of a z cell starts the cell (k == rows: the valid total)
      if (CELLS && k % (uint32_t)V == 0) bs[k / (uint32_t)V] = (uint32_t)i;
    }
    if (i == N - 1)
      for (uint32_t k = key + 1; k <= rows + 1; ++k) {
        sv.rowstart[k] = (uint32_t)N;
        if (CELLS && k <= rows && k % (uint32_t)V == 0) bs[k / (uint32_t)V] = (uint32_t)N;
      }
  }
}

__global__ void __launch_bounds__(kRowThreads)
segment_rows_kernel(const void *ws, size_t stride, int N, int Vz, int V, float *__restrict__ grid) {
  extern __shared__ float acc[];  // [V][kRowThreads + 1]
  constexpr int S = kRowThreads + 1;
  const int b = blockIdx.y, tid = threadIdx.x;
  const int rows = Vz * V;
  const int row = blockIdx.x * kRowThreads + tid;
  const SortedView sv = sorted_view(const_cast<void *>(ws), stride, b, N);
  for (int x = 0; x < V; ++x) acc[x * S + tid] = 0.f;
  if (row < rows) {
    const int oz = row / V, oy = row - oz * V;
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 1; dy >= 0; --dy) {         // ascending sorted position: row oy - 1, then row oy
        const int bz = oz - dz, by = oy - dy;
        if (bz < 0 || by < 0) continue;
        const int key = bz * V + by;
        const int i0 = (int)sv.rowstart[key], i1 = (int)sv.rowstart[key + 1];
        for (int i = i0; i < i1; ++i) {
          const uint4 r = sv.srec[i];
          const int ix = (int)r.x;
          const float rz = __uint_as_float(r.y), ry = __uint_as_float(r.z), rx = __uint_as_float(r.w);
          // the same fp32 weight products as the default plane scatter (blur_xy.cu)
          const float wzy = (dz ? rz : 1.f - rz) * (dy ? ry : 1.f - ry);
          // (explicit fused multiply-adds: the plane-local build in blur_xy.cu spells the same)
          acc[ix * S + tid] = fmaf(wzy, 1.f - rx, acc[ix * S + tid]);
          if (ix + 1 < V) acc[(ix + 1) * S + tid] = fmaf(wzy, rx, acc[(ix + 1) * S + tid]);
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

// the first two kernels: items + records, then the per-projection sort and the row-segment table
// (cells != NULL: records, z-cell bytes and z-cell boundaries in the plane-local layout)
static int launch_sort(const PoseArgs *a, const float *tr_pc_in, int P, int N, int Vz, int V,
                       float *tr_pc_out, const CellsView *cells, void *ws, size_t ws_bytes,
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
  const size_t stride = sorted_stride_bytes(N, Vz, V);
  const CellsView cv = cells ? *cells : CellsView{nullptr, nullptr, nullptr, nullptr, 0, 0};
  const int n_items = cells ? cv.Npad : N;       // the padding z-cell bytes are written too
  const dim3 gi((n_items + kItemThreads - 1) / kItemThreads, P);
  if (cells)
    sort_items_kernel<true><<<gi, kItemThreads, 0, s>>>(src, a ? tr_pc_out : nullptr, ws, stride, N,
                                                        Vz, V, cv);
  else
    sort_items_kernel<false><<<gi, kItemThreads, 0, s>>>(src, a ? tr_pc_out : nullptr, ws, stride,
                                                         N, Vz, V, cv);
  if (int e = check_launch("sort_items")) return e;
  const int cnt_words = sort_counter_words(Vz * V);
  const size_t cnt_smem = (size_t)cnt_words * sizeof(uint32_t);
  static DeviceOnce sort_once;
  if (sort_once.first()) {
    const int small = kSortBinsMax * (int)sizeof(uint32_t);
    const int big = small + 2 * kSmemSortMax * (int)sizeof(uint32_t);
    cudaFuncSetAttribute(sort_points_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(sort_points_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
    cudaFuncSetAttribute(sort_points_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
    cudaFuncSetAttribute(sort_points_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, small);
  }
  const bool in_smem = N <= kSmemSortMax;
  const size_t smem = cnt_smem + (in_smem ? (size_t)2 * N * sizeof(uint32_t) : 0);
  if (in_smem && cells) sort_points_kernel<true, true><<<P, kSortThreads, smem, s>>>(ws, stride, N, Vz, V, cv, cnt_words);
  else if (in_smem) sort_points_kernel<true, false><<<P, kSortThreads, smem, s>>>(ws, stride, N, Vz, V, cv, cnt_words);
  else if (cells) sort_points_kernel<false, true><<<P, kSortThreads, smem, s>>>(ws, stride, N, Vz, V, cv, cnt_words);
  else sort_points_kernel<false, false><<<P, kSortThreads, smem, s>>>(ws, stride, N, Vz, V, cv, cnt_words);
  return check_launch("sort_points");
}

int launch_scatter_sorted(const PoseArgs *a, const float *tr_pc_in, int P, int N, int Vz, int V,
                          float *tr_pc_out, float *grid, void *ws, size_t ws_bytes,
                          cudaStream_t s) {
  if (int e = launch_sort(a, tr_pc_in, P, N, Vz, V, tr_pc_out, nullptr, ws, ws_bytes, s)) return e;
  const size_t stride = sorted_stride_bytes(N, Vz, V);
  const size_t smem = (size_t)V * (kRowThreads + 1) * sizeof(float);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(segment_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         128 * (kRowThreads + 1) * (int)sizeof(float));
  }
  dim3 g((Vz * V + kRowThreads - 1) / kRowThreads, P);
  segment_rows_kernel<<<g, kRowThreads, smem, s>>>(ws, stride, N, Vz, V, grid);
  return check_launch("segment_rows");
}

// Plane-local variant: pose -> tr_pc (NULL ok) + the sorted cell records; the planes themselves
// are built by the blur-XY kernel from `rowstart` (row r of projection b: rowstart[b * stride + r],
// stride in 32-bit words)
int launch_sort_cells(const PoseArgs &a, float *tr_pc, const CellsView &cells, void *ws,
                      size_t ws_bytes, const uint32_t **rowstart, size_t *rowstart_stride,
                      cudaStream_t s) {
  if (int e = launch_sort(&a, nullptr, a.P, a.N, a.Vz, a.V, tr_pc, &cells, ws, ws_bytes, s)) return e;
  const size_t stride = sorted_stride_bytes(a.N, a.Vz, a.V);
  *rowstart = sorted_view(ws, stride, 0, a.N).rowstart;
  *rowstart_stride = stride / sizeof(uint32_t);
  return DPC_OK;
}

// the workspace of projections [b0, ...) inside a whole-batch workspace
void *sorted_workspace_at(void *ws, int b0, int N, int Vz, int V) {
  return (char *)ws + (size_t)b0 * sorted_stride_bytes(N, Vz, V);
}

}  // namespace dpc
