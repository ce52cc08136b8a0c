// C ABI (include/dpc_b200.h): argument validation, workspace carving and the
// stream-ordered launch sequences.  No allocation, no host synchronisation.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace dpc {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return DPC_ERR_CUDA;
  }
  return DPC_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// Profiling hook (dpc_project_profile): when set, an event is recorded after
// every stage of dpc_project_fwd / dpc_project_bwd on the launching stream.
static int g_pdl_override = -1;       // dpc_set_programmatic_launch(): 0 / 1, or -1 = DPC_PDL / default
bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("DPC_PDL");
    return e ? atoi(e) != 0 : true;
  }();
  return g_pdl_override >= 0 ? g_pdl_override != 0 : on;
}

// DPC_TAP_EPS (env): total tap magnitude effective_radius() may drop (default 1e-7; 0: none);
// dpc_set_tap_truncation() overrides it for the process (a negative value: back to the default)
static float g_tap_eps_override = -1.f;
float tap_truncation_eps() {
  static const float eps = [] {
    const char *e = getenv("DPC_TAP_EPS");
    const double v = e ? atof(e) : 1e-7;
    return (float)(v < 0 ? 0 : v);
  }();
  const float o = g_tap_eps_override;
  return o >= 0.f ? o : eps;
}

static thread_local cudaEvent_t *tl_stage_events = nullptr;
static thread_local int tl_stage_idx = 0;
static inline void stage_mark(cudaStream_t s) {
  if (tl_stage_events) cudaEventRecord(tl_stage_events[tl_stage_idx++], s);
}

struct Workspace {
  double *pose_partials;
  float *scale_partials;
  int *counters;          // P ints: per-projection arrival counters of the fused finalize
  void *sorted;
  size_t sorted_bytes;
  float4 *part;           // [2][P][N] per-plane partial gathers of the backward (plane-local path)
  size_t total;
};

static Workspace carve(const dpc_params *p, void *base) {
  Workspace w;
  char *c = (char *)base;
  size_t off = 0;
  w.pose_partials = (double *)(c + off);
  off += align256((size_t)p->P * pose_partial_blocks(p->N) * 8 * sizeof(double));
  w.scale_partials = (float *)(c + off);
  off += align256((size_t)p->P * drc_scale_partial_blocks(p->V) * sizeof(float));
  w.counters = (int *)(c + off);
  off += align256((size_t)p->P * sizeof(int));
  w.sorted = (void *)(c + off);
  w.sorted_bytes = sorted_workspace_bytes(p->P, p->N, p->Vz, p->V);
  off += align256(w.sorted_bytes);
  w.part = (float4 *)(c + off);
  off += align256((size_t)2 * p->P * p->N * sizeof(float4));
  w.total = off;
  return w;
}

static int check_params(const dpc_params *p, bool need_points) {
  if (!p) { set_error("params is NULL"); return DPC_ERR_ARG; }
  if (p->P < 1) { set_error("P=%d must be >= 1", p->P); return DPC_ERR_ARG; }
  if (need_points && p->N < 1) { set_error("N=%d must be >= 1", p->N); return DPC_ERR_ARG; }
  if (p->V != 32 && p->V != 64 && p->V != 128) {
    set_error("vox_size=%d unsupported (32, 64 or 128)", p->V);
    return DPC_ERR_ARG;
  }
  if (p->Vz < 2 || p->Vz > 192) {
    set_error("vox_size_z=%d unsupported (2..192)", p->Vz);
    return DPC_ERR_ARG;
  }
  if ((double)p->P * p->Vz * p->V * p->V >= 2147483648.0 * 4) {
    set_error("grid too large");
    return DPC_ERR_ARG;
  }
  return DPC_OK;
}

static int check_taps(const float *t, int k, const char *name) {
  if (k == 0) return DPC_OK;
  if (k < 0 || k % 2 == 0 || k > DPC_MAX_TAPS || !t) {
    set_error("%s: tap count %d must be odd, <= %d, with a non-NULL pointer", name, k,
              DPC_MAX_TAPS);
    return DPC_ERR_ARG;
  }
  return DPC_OK;
}

static int check_ws(const dpc_params *p, void *ws, size_t bytes) {
  if (!ws || bytes < carve(p, nullptr).total) {
    set_error("workspace too small: need %zu bytes, got %zu", carve(p, nullptr).total, bytes);
    return DPC_ERR_WORKSPACE;
  }
  return DPC_OK;
}

#define DPC_REQUIRE(ptr)                                  \
  do {                                                    \
    if (!(ptr)) {                                         \
      set_error("%s: %s is NULL", __func__, #ptr);        \
      return DPC_ERR_ARG;                                 \
    }                                                     \
  } while (0)
#define DPC_TRY(expr)            \
  do {                           \
    int _e = (expr);             \
    if (_e != DPC_OK) return _e; \
  } while (0)

// f2: how the projections address an un-replicated cloud tensor (replicas == 0: plain [P,N,3])
struct Replica {
  int replicas = 0, N_src = 0;
  const int *sel = nullptr;
};

static PoseArgs pose_args(const dpc_params *p, const float *points, const float *quat,
                          const float *trans, const float *focal, const Replica &rep = Replica()) {
  PoseArgs a;
  a.points = points; a.quat = quat; a.trans = trans; a.focal = focal;
  a.P = p->P; a.N = p->N; a.Vz = p->Vz; a.V = p->V;
  a.cam_dist = p->camera_distance;
  a.focal_const = p->focal_length;
  a.replicas = rep.replicas; a.N_src = rep.N_src; a.sel = rep.sel;
  return a;
}

static int check_replica(const dpc_params *p, int replicas, int N_src, const int32_t *sel) {
  if (replicas < 1 || p->P % replicas != 0) {
    set_error("replicas=%d must be >= 1 and divide P=%d", replicas, p->P);
    return DPC_ERR_ARG;
  }
  if (sel ? (p->N > N_src) : (p->N != N_src)) {
    set_error("N=%d points per projection do not fit a cloud of N_src=%d points (%s)", p->N, N_src,
              sel ? "with a selection N <= N_src" : "without a selection N == N_src");
    return DPC_ERR_ARG;
  }
  return DPC_OK;
}

static DrcArgs drc_args(const dpc_params *p, const float *grid, const float *scale) {
  DrcArgs a;
  a.grid = grid; a.scale = scale;
  a.P = p->P; a.P_total = p->P; a.Vz = p->Vz; a.V = p->V;
  a.cam_dist = (float)p->camera_distance;
  a.max_depth = (float)p->max_depth;
  a.clip = (float)p->drc_clip;
  a.logsum = p->drc_logsum;
  a.flip_y = p->flip_y;
  return a;
}

static size_t grid_bytes(const dpc_params *p) {
  return (size_t)p->P * p->Vz * p->V * p->V * sizeof(float);
}

}  // namespace dpc

using namespace dpc;

extern "C" {

int dpc_version(void) { return DPC_B200_VERSION; }

size_t dpc_cells_bytes(const dpc_params *p) {
  if (!p || p->P < 1 || p->N < 1) return 0;
  return ray_ck_offset(p->P, p->N, p->Vz) + ray_ck_bytes(p->P, p->Vz, p->V);
}
const char *dpc_last_error(void) { return g_err; }

size_t dpc_workspace_bytes(const dpc_params *p) {
  if (!p || p->P < 1 || p->N < 0 || p->V < 1) return 0;
  return carve(p, nullptr).total;
}

int dpc_pose_fwd(const dpc_params *p, const float *points, const float *quat, const float *trans,
                 const float *focal, float *tr_pc, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(tr_pc);
  return launch_pose_scatter(pose_args(p, points, quat, trans, focal), tr_pc, nullptr,
                             (cudaStream_t)stream);
}

int dpc_pose_bwd(const dpc_params *p, const float *points, const float *quat, const float *trans,
                 const float *focal, const float *g_tr_pc, float *g_points, float *g_quat,
                 float *g_trans, float *g_focal, void *workspace, size_t workspace_bytes,
                 void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(g_tr_pc); DPC_REQUIRE(g_points);
  DPC_TRY(check_ws(p, workspace, workspace_bytes));
  const Workspace w = carve(p, workspace);
  const PoseArgs a = pose_args(p, points, quat, trans, focal);
  cudaStream_t s = (cudaStream_t)stream;
  DPC_TRY(launch_gather_pose_bwd(a, nullptr, g_tr_pc, g_points, w.pose_partials, s));
  return launch_finalize(a, w.pose_partials, pose_partial_blocks(p->N), nullptr, 0, g_quat,
                         trans ? g_trans : nullptr, focal ? g_focal : nullptr, nullptr, s);
}

int dpc_scatter_fwd(const dpc_params *p, const float *tr_pc, float *grid, int mode,
                    void *workspace, size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(tr_pc); DPC_REQUIRE(grid);
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == DPC_SCATTER_SORTED) {
    DPC_TRY(check_ws(p, workspace, workspace_bytes));
    const Workspace w = carve(p, workspace);
    return launch_scatter_sorted(nullptr, tr_pc, p->P, p->N, p->Vz, p->V, nullptr, grid, w.sorted,
                                 w.sorted_bytes, s);
  }
  if (mode != DPC_SCATTER_ATOMIC) { set_error("unknown scatter mode %d", mode); return DPC_ERR_ARG; }
  if (cudaMemsetAsync(grid, 0, grid_bytes(p), s) != cudaSuccess) return check_launch("memset");
  return launch_scatter_trpc(tr_pc, p->P, p->N, p->Vz, p->V, grid, s);
}

int dpc_scatter_bwd(const dpc_params *p, const float *tr_pc, const float *g_grid, float *g_tr_pc,
                    void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(tr_pc); DPC_REQUIRE(g_grid); DPC_REQUIRE(g_tr_pc);
  return launch_gather_trpc_bwd(tr_pc, p->P, p->N, p->Vz, p->V, g_grid, g_tr_pc,
                                (cudaStream_t)stream);
}

int dpc_blur3d(const dpc_params *p, const float *src, float *dst, const float *tx, int kx,
               const float *ty, int ky, const float *tz, int kz, void *stream) {
  DPC_TRY(check_params(p, false));
  DPC_REQUIRE(src); DPC_REQUIRE(dst);
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  cudaStream_t s = (cudaStream_t)stream;
  BlurXYArgs a;
  a.src = src; a.dst = dst; a.bits_out = nullptr; a.bits_in = nullptr;
  a.planes = p->P * p->Vz; a.V = p->V; a.clamp_in = false;
  DPC_TRY(launch_blur_xy(a, tx, kx, ty, ky, s));
  return launch_blur_z(dst, dst, p->P, p->Vz, p->V, tz, kz, s);
}

int dpc_blur3d_clamped(const dpc_params *p, const float *src, float *dst, const float *tx, int kx,
                       const float *ty, int ky, const float *tz, int kz, void *stream) {
  DPC_TRY(check_params(p, false));
  DPC_REQUIRE(src); DPC_REQUIRE(dst);
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  cudaStream_t s = (cudaStream_t)stream;
  BlurXYArgs a;
  a.src = src; a.dst = dst; a.bits_out = nullptr; a.bits_in = nullptr;
  a.planes = p->P * p->Vz; a.V = p->V; a.clamp_in = true;
  DPC_TRY(launch_blur_xy(a, tx, kx, ty, ky, s));
  return launch_blur_z(dst, dst, p->P, p->Vz, p->V, tz, kz, s);
}

// ---- a14 / f3: point-feature (RGB) branch --------------------------------------
static int check_channels(int C) {
  if (C < 1 || C > feat_max_channels()) {
    set_error("feature channels=%d unsupported (1..%d)", C, feat_max_channels());
    return DPC_ERR_ARG;
  }
  return DPC_OK;
}

int dpc_feat_scatter_fwd(const dpc_params *p, int C, const float *tr_pc, const float *feat,
                         float *fgrid, void *stream) {
  DPC_TRY(check_params(p, true)); DPC_TRY(check_channels(C));
  DPC_REQUIRE(tr_pc); DPC_REQUIRE(feat); DPC_REQUIRE(fgrid);
  return launch_feat_scatter(tr_pc, feat, p->P, p->N, C, p->Vz, p->V, fgrid, (cudaStream_t)stream);
}

int dpc_feat_scatter_bwd(const dpc_params *p, int C, const float *tr_pc, const float *feat,
                         const float *g_fgrid, const float *raw, float *g_feat, float *g_tr_pc,
                         void *stream) {
  DPC_TRY(check_params(p, true)); DPC_TRY(check_channels(C));
  DPC_REQUIRE(tr_pc); DPC_REQUIRE(feat); DPC_REQUIRE(g_fgrid); DPC_REQUIRE(g_feat);
  return launch_feat_gather_bwd(tr_pc, feat, g_fgrid, raw, p->P, p->N, C, p->Vz, p->V, g_feat,
                                g_tr_pc, (cudaStream_t)stream);
}

int dpc_colour_fwd(const dpc_params *p, int C, const float *probs, const float *fgrid,
                   const float *div, float eps, int clip_after, float *proj_rgb, float *voxels_rgb,
                   void *stream) {
  DPC_TRY(check_params(p, false)); DPC_TRY(check_channels(C));
  DPC_REQUIRE(probs); DPC_REQUIRE(fgrid); DPC_REQUIRE(proj_rgb);
  return launch_colour_fwd(probs, fgrid, div, eps, clip_after, p->P, C, p->Vz, p->V, p->flip_y,
                           proj_rgb, voxels_rgb, (cudaStream_t)stream);
}

int dpc_colour_bwd(const dpc_params *p, int C, const float *probs, const float *fgrid,
                   const float *div, float eps, int clip_after, const float *g_proj_rgb,
                   float *g_probs, float *g_fgrid, void *stream) {
  DPC_TRY(check_params(p, false)); DPC_TRY(check_channels(C));
  DPC_REQUIRE(probs); DPC_REQUIRE(fgrid); DPC_REQUIRE(g_proj_rgb); DPC_REQUIRE(g_probs);
  DPC_REQUIRE(g_fgrid);
  return launch_colour_bwd(probs, fgrid, div, eps, clip_after, p->P, C, p->Vz, p->V, p->flip_y,
                           g_proj_rgb, g_probs, g_fgrid, (cudaStream_t)stream);
}

int dpc_drc_fwd(const dpc_params *p, const float *voxels, float *mask, float *depth, float *probs,
                void *stream) {
  DPC_TRY(check_params(p, false));
  DPC_REQUIRE(voxels); DPC_REQUIRE(mask);
  return launch_blurz_drc_fwd(drc_args(p, voxels, nullptr), nullptr, 0, nullptr, mask, depth,
                              nullptr, probs, (cudaStream_t)stream);
}

int dpc_drc_bwd(const dpc_params *p, const float *voxels, const float *g_mask, const float *g_depth,
                const float *g_probs, float *g_voxels, void *workspace, size_t workspace_bytes,
                void *stream) {
  (void)workspace; (void)workspace_bytes;
  DPC_TRY(check_params(p, false));
  DPC_REQUIRE(voxels); DPC_REQUIRE(g_voxels);
  return launch_drc_blurz_bwd(drc_args(p, voxels, nullptr), nullptr, 0, g_mask, g_depth, g_probs,
                              nullptr, g_voxels, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int dpc_depth_from_probs_fwd(const dpc_params *p, const float *probs, float *depth, void *stream) {
  DPC_TRY(check_params(p, false));
  DPC_REQUIRE(probs); DPC_REQUIRE(depth);
  return launch_depth_from_probs(probs, depth, p->P, p->Vz, p->V, (float)p->camera_distance,
                                 (float)p->max_depth, (cudaStream_t)stream);
}

int dpc_depth_from_probs_bwd(const dpc_params *p, const float *g_depth, float *g_probs,
                             void *stream) {
  DPC_TRY(check_params(p, false));
  DPC_REQUIRE(g_depth); DPC_REQUIRE(g_probs);
  return launch_depth_from_probs_bwd(g_depth, g_probs, p->P, p->Vz, p->V,
                                     (float)p->camera_distance, (float)p->max_depth,
                                     (cudaStream_t)stream);
}

// ---- two-stream half-batch pipeline -------------------------------------------
// The batch runs as chunks of projections on two internal streams, so one chunk's kernels overlap
// the other's.  History (B200, P=64, 64^3, CUDA-graph replay): with the first kernels chunking
// lost (1 chunk 218 us/step, 2 chunks 223, 4 chunks 264, 8 chunks 333: wave quantisation and
// per-kernel ramp); with the current grid kernels the latency-bound point kernels are a quarter
// of the step and two half-batches win (148.0 vs 153.6 us), so that is the default for P >= 64.
// DPC_CHUNK=<n> (env) sets the chunk size (n >= P: one pass on the caller's stream).
struct Pipeline {
  cudaStream_t side[2] = {nullptr, nullptr};
  cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
  bool ready = false;
};
static thread_local Pipeline tl_pipe[64];
static thread_local bool tl_force_single = false;

static Pipeline *get_pipeline() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  Pipeline &pl = tl_pipe[dev];
  if (!pl.ready) {
    for (int i = 0; i < 2; ++i) {
      if (cudaStreamCreateWithFlags(&pl.side[i], cudaStreamNonBlocking) != cudaSuccess) return nullptr;
      if (cudaEventCreateWithFlags(&pl.join[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    if (cudaEventCreateWithFlags(&pl.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    pl.ready = true;
  }
  return &pl;
}

// The only state the library keeps: per host thread and device, two non-blocking side streams and
// three events for the half-batch split, created on first use.  dpc_release() destroys the calling
// thread's set for the current device (after the caller has synchronised the work it issued).
static int release_pipeline() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return DPC_OK;
  Pipeline &pl = tl_pipe[dev];
  if (!pl.ready) return DPC_OK;
  for (int i = 0; i < 2; ++i) {
    cudaStreamDestroy(pl.side[i]);
    cudaEventDestroy(pl.join[i]);
    pl.side[i] = nullptr;
    pl.join[i] = nullptr;
  }
  cudaEventDestroy(pl.fork);
  pl.fork = nullptr;
  pl.ready = false;
  return check_launch("release");
}

static int chunk_size(const dpc_params *p) {
  if (tl_force_single) return p->P;
  static int env_chunk = -1;
  if (env_chunk < 0) {
    const char *e = getenv("DPC_CHUNK");
    env_chunk = e ? atoi(e) : 0;
  }
  if (env_chunk > 0) return env_chunk;
  // Two half-batches on two internal streams: the latency-bound point kernels (pose, binning,
  // pose adjoint: ~25 % of the step, a few hundred CTAs each) of one half overlap the FMA-bound
  // grid kernels of the other.  Measured at workload A: 148.0 vs 153.6 us per step; quarters
  // lose (177.5 us: too few CTAs per kernel).
  // Larger batches at 64^3 and below: chunks of at most 64 MiB of grid (64 projections at 64^3), so
  // that a chunk's producer -> consumer grids stay in L2 (config-3 shapes, 256 projections:
  // 522.5 us per step as two halves of 128, 511.9 as four chunks of 64, 518.1 as eight of 32).
  // At 128^3 a projection is 8 MiB and smaller chunks only lose (workload B: 2004 us as halves of
  // 64, 2021 / 2059 / 2179 us as chunks of 16 / 8 / 4).
  if (p->P >= 64 && p->P % 2 == 0) {
    int chunk = p->P / 2;
    if (p->V <= 64) {
      const long long per = (long long)p->Vz * p->V * p->V * 4;
      const int cap = (int)((64ll << 20) / per);
      if (cap >= 32 && chunk > cap) chunk = cap;
    }
    return chunk;
  }
  return p->P;
}

int dpc_project_chunks(const dpc_params *p) {
  if (!p || p->P < 1) return 0;
  const int c = chunk_size(p);
  return c < p->P ? (p->P + c - 1) / c : 1;
}

int dpc_release(void) { return release_pipeline(); }

int dpc_set_programmatic_launch(int on) {
  g_pdl_override = on < 0 ? -1 : (on != 0);
  return DPC_OK;
}

int dpc_set_tap_truncation(double eps) {
  g_tap_eps_override = eps >= 0.0 ? (float)eps : -1.f;
  return DPC_OK;
}

int dpc_tap_radius(const float *taps_host, int n) {
  if (!taps_host || n < 1 || n % 2 == 0) return -1;
  return effective_radius(taps_host, n);
}

int dpc_project_kernels_per_chunk(const dpc_params *p) {
  if (!p || p->N < 1) return 0;
  // forward: pose (+ binning) | [binning] | plane scatter + blur XY | blur Z + DRC;
  // backward: DRC + blur Z adjoint | blur XY adjoint + plane gather | pose adjoint
  return pose_bin_split(p->N) > 0 ? 6 : 7;
}

struct FwdPtrs {
  const float *points, *quat, *trans, *focal, *scale;
  float *tr_pc, *grid_b; uint32_t *bits; float *mask, *depth, *voxels, *probs;
  void *cells;
  Replica rep;
};

// Projections [b0, ...) of a batch: the cloud tensor and the selection rows they start at.  A
// replica-aware range starts at a multiple of `replicas` (replica_chunk), so that projection
// b0 + i of the batch is replica i % replicas of cloud b0 / replicas + i / replicas.
static const float *range_points(const float *points, const Replica &rep, int b0, int N) {
  if (rep.replicas == 0) return points + (size_t)b0 * N * 3;
  return points + (size_t)(b0 / rep.replicas) * rep.N_src * 3;
}
static Replica range_replica(const Replica &rep, int b0, int N) {
  Replica r = rep;
  if (r.sel) r.sel += (size_t)b0 * N;
  return r;
}
static int replica_chunk(const dpc_params *p, const Replica &rep) {
  const int c = chunk_size(p);
  return (rep.replicas > 0 && c % rep.replicas != 0) ? p->P : c;
}

// The plane-local path packs (n, iy, ix) into 32 bits and bins by a z-cell byte.
static bool plane_local_ok(const dpc_params *p) {
  return p->N <= 65535 && p->V <= 256 && p->Vz <= 192;
}

// The DRC kernels' fast saved state (drc.cu): decided from what BOTH passes know -- the cell
// records and the parameters; the scatter mode does not matter (both modes save the same state).
static bool fast_ray_state(const dpc_params *p, const void *cells) {
  // (clip value > 0: the saved v >= clip is then strictly positive and its sign bit is free)
  return cells && p->Vz == p->V && p->outputs == 0 &&
         p->drc_logsum != 0 && p->drc_clip > 0.0 && p->drc_clip < 0.5;
}
static void set_ray_state(DrcArgs &da, const dpc_params *p, void *cells, int b0) {
  da.ck_slots = ray_ck_slots(p->Vz);
  da.tck = (float *)((char *)cells + ray_ck_offset(p->P, p->N, p->Vz)) +
           (size_t)b0 * da.ck_slots * p->V * p->V;
}

// the records of projections [b0, b0 + n) inside a whole-batch cells buffer
static CellsView cells_range(void *base, const dpc_params *p, int b0) {
  CellsView v = cells_view(base, p->P, p->N, p->Vz);
  v.cellz += (size_t)b0 * v.Npad;
  v.rec += (size_t)b0 * p->N;
  v.srec += (size_t)b0 * p->N;
  v.binstart += (size_t)b0 * v.zstride;
  return v;
}

// projections [b0, b0+n) of the forward pass on stream s
static int project_fwd_range(const dpc_params *p, int b0, int n, const FwdPtrs &q, const float *tx,
                             int kx, const float *ty, int ky, const float *tz, int kz,
                             int scatter_mode, const Workspace &w, cudaStream_t s) {
  dpc_params sp = *p;
  sp.P = n;
  const size_t N3 = (size_t)p->N * 3, G = (size_t)p->Vz * p->V * p->V, I = (size_t)p->V * p->V;
  const PoseArgs pa = pose_args(&sp, range_points(q.points, q.rep, b0, p->N), q.quat + b0 * 4,
                                q.trans ? q.trans + b0 * 3 : nullptr,
                                q.focal ? q.focal + b0 : nullptr, range_replica(q.rep, b0, p->N));
  float *grid = q.grid_b + b0 * G;
  float *tr_pc = q.tr_pc ? q.tr_pc + b0 * N3 : nullptr;
  stage_mark(s);
  // clamp(raw,0,1) + raw<=1 mask + blur X + blur Y, in place (identity taps when kernel=None)
  BlurXYArgs b;
  b.src = grid; b.dst = grid; b.bits_out = q.bits + b0 * (G / 32); b.bits_in = nullptr;
  b.planes = n * p->Vz; b.V = p->V; b.clamp_in = true;
  if (q.cells && scatter_mode == DPC_SCATTER_ATOMIC) {
    // plane-local path: pose -> cell records; every Z-plane is then built in
    // shared memory by the blur kernel itself (no memset, no global atomics)
    stage_mark(s);
    b.cells = cells_range(q.cells, p, b0);
    b.Vz = p->Vz; b.N = p->N; b.P = n;
    DPC_TRY(launch_pose_cells(pa, tr_pc, b.cells, s));
  } else if (q.cells && scatter_mode == DPC_SCATTER_SORTED) {
    // deterministic plane-local path: pose -> records sorted by grid row + row-segment table;
    // the blur kernel sums every plane row in a fixed order (no atomics, no raw grid in HBM) and
    // the saved state is the plane-local one, so the backward does not depend on the mode
    stage_mark(s);
    b.cells = cells_range(q.cells, p, b0);
    b.Vz = p->Vz; b.N = p->N; b.P = n;
    DPC_TRY(launch_sort_cells(pa, tr_pc, b.cells, sorted_workspace_at(w.sorted, b0, p->N, p->Vz, p->V),
                              sorted_workspace_bytes(n, p->N, p->Vz, p->V), &b.rowstart,
                              &b.rowstart_stride, s));
  } else if (scatter_mode == DPC_SCATTER_SORTED) {
    stage_mark(s);
    DPC_TRY(launch_scatter_sorted(&pa, nullptr, n, p->N, p->Vz, p->V, tr_pc, grid,
                                  sorted_workspace_at(w.sorted, b0, p->N, p->Vz, p->V),
                                  sorted_workspace_bytes(n, p->N, p->Vz, p->V), s));
  } else {
    if (cudaMemsetAsync(grid, 0, (size_t)n * G * sizeof(float), s) != cudaSuccess)
      return check_launch("memset");
    stage_mark(s);
    DPC_TRY(launch_pose_scatter(pa, tr_pc, grid, s));
  }
  stage_mark(s);
  DPC_TRY(launch_blur_xy(b, tx, kx, ty, ky, s));
  stage_mark(s);
  // blur Z + scale + clip + DRC; the blurred occupancy overwrites grid in place (saved for bwd)
  DrcArgs da = drc_args(&sp, grid, q.scale ? q.scale + b0 : nullptr);
  da.P_total = p->P;
  if (fast_ray_state(p, q.cells)) set_ray_state(da, p, q.cells, b0);
  DPC_TRY(launch_blurz_drc_fwd(da, tz, kz, grid, q.mask + b0 * I, q.depth ? q.depth + b0 * I : nullptr,
                               q.voxels ? q.voxels + b0 * G : nullptr,
                               q.probs ? q.probs + b0 * I : nullptr, s));
  stage_mark(s);
  return DPC_OK;
}

static int project_fwd_impl(const dpc_params *p, const Replica &rep, const float *points,
                    const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, void *cells, float *mask,
                    float *depth, float *voxels, float *probs, void *workspace,
                    size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(grid_b); DPC_REQUIRE(clamp_bits);
  DPC_REQUIRE(mask);
  if ((voxels != nullptr) != ((p->outputs & DPC_OUT_VOXELS) != 0) ||
      (probs != nullptr) != ((p->outputs & DPC_OUT_PROBS) != 0)) {
    set_error("project_fwd: params.outputs=%d does not match the voxels / probs pointers", p->outputs);
    return DPC_ERR_ARG;
  }
  if (!plane_local_ok(p)) cells = nullptr;
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  if (scatter_mode != DPC_SCATTER_SORTED && scatter_mode != DPC_SCATTER_ATOMIC) {
    set_error("unknown scatter mode %d", scatter_mode);
    return DPC_ERR_ARG;
  }
  Workspace w{};
  if (scatter_mode == DPC_SCATTER_SORTED) {
    DPC_TRY(check_ws(p, workspace, workspace_bytes));
    w = carve(p, workspace);
  }
  cudaStream_t s = (cudaStream_t)stream;
  const FwdPtrs q{points, quat, trans, focal, scale, tr_pc, grid_b, clamp_bits, mask, depth, voxels,
                  probs, cells, rep};
  const int chunk = replica_chunk(p, rep);
  Pipeline *pl = chunk < p->P ? get_pipeline() : nullptr;
  if (!pl) return project_fwd_range(p, 0, p->P, q, tx, kx, ty, ky, tz, kz, scatter_mode, w, s);
  cudaEventRecord(pl->fork, s);
  for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(pl->side[i], pl->fork, 0);
  int rc = DPC_OK;
  for (int b0 = 0, c = 0; b0 < p->P && rc == DPC_OK; b0 += chunk, ++c) {
    const int n = p->P - b0 < chunk ? p->P - b0 : chunk;
    rc = project_fwd_range(p, b0, n, q, tx, kx, ty, ky, tz, kz, scatter_mode, w, pl->side[c & 1]);
  }
  for (int i = 0; i < 2; ++i) {
    cudaEventRecord(pl->join[i], pl->side[i]);
    cudaStreamWaitEvent(s, pl->join[i], 0);
  }
  return rc;
}

int dpc_project_fwd(const dpc_params *p, const float *points, const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, void *cells, float *mask,
                    float *depth, float *voxels, float *probs, void *workspace,
                    size_t workspace_bytes, void *stream) {
  return project_fwd_impl(p, Replica(), points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz,
                          scatter_mode, tr_pc, grid_b, clamp_bits, cells, mask, depth, voxels, probs,
                          workspace, workspace_bytes, stream);
}

int dpc_project_replicated_fwd(const dpc_params *p, int replicas, int N_src, const int32_t *sel,
                    const float *points, const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, void *cells, float *mask,
                    float *depth, float *voxels, float *probs, void *workspace,
                    size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_TRY(check_replica(p, replicas, N_src, sel));
  Replica rep;
  rep.replicas = replicas; rep.N_src = N_src; rep.sel = sel;
  return project_fwd_impl(p, rep, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz,
                          scatter_mode, tr_pc, grid_b, clamp_bits, cells, mask, depth, voxels, probs,
                          workspace, workspace_bytes, stream);
}

struct BwdPtrs {
  const float *points, *quat, *trans, *focal, *scale, *grid_b; const uint32_t *bits;
  const float *g_mask, *g_depth, *g_probs, *g_voxels, *g_tr_pc;
  float *g_grid, *g_points, *g_quat, *g_trans, *g_focal, *g_scale;
  void *cells;
  Replica rep;
  void *fast_rays = nullptr;   // the cells buffer when the forward saved the fast ray state
};

static int project_bwd_range(const dpc_params *p, int b0, int n, const BwdPtrs &q, const float *tx,
                             int kx, const float *ty, int ky, const float *tz, int kz,
                             const Workspace &w, cudaStream_t s) {
  dpc_params sp = *p;
  sp.P = n;
  const size_t N3 = (size_t)p->N * 3, G = (size_t)p->Vz * p->V * p->V, I = (size_t)p->V * p->V;
  const PoseArgs pa = pose_args(&sp, range_points(q.points, q.rep, b0, p->N), q.quat + b0 * 4,
                                q.trans ? q.trans + b0 * 3 : nullptr,
                                q.focal ? q.focal + b0 : nullptr, range_replica(q.rep, b0, p->N));
  float *g_grid = q.g_grid + b0 * G;
  DrcArgs da = drc_args(&sp, q.grid_b + b0 * G, q.scale ? q.scale + b0 : nullptr);
  da.P_total = p->P;
  if (q.fast_rays) set_ray_state(da, p, q.fast_rays, b0);
  stage_mark(s);
  DPC_TRY(launch_drc_blurz_bwd(da, tz, kz, q.g_mask ? q.g_mask + b0 * I : nullptr,
                               q.g_depth ? q.g_depth + b0 * I : nullptr,
                               q.g_probs ? q.g_probs + b0 * I : nullptr,
                               q.g_voxels ? q.g_voxels + b0 * G : nullptr, g_grid,
                               w.scale_partials + (size_t)b0 * drc_scale_partial_blocks(p->V),
                               w.counters + b0, n, s));
  stage_mark(s);
  BlurXYArgs b;
  b.src = g_grid; b.dst = g_grid; b.bits_out = nullptr; b.bits_in = q.bits + b0 * (G / 32);
  b.planes = n * p->Vz; b.V = p->V; b.clamp_in = false;
  // adjoint of a correlation = correlation with the reversed taps
  float rx[DPC_MAX_TAPS], ry[DPC_MAX_TAPS];
  for (int i = 0; i < kx; ++i) rx[i] = tx[kx - 1 - i];
  for (int i = 0; i < ky; ++i) ry[i] = ty[ky - 1 - i];
  float4 *part = nullptr;
  if (q.cells) {
    // plane-local path: the masked dL/draw plane is gathered at the touching
    // points' corners straight from shared memory; g_grid is not rewritten
    b.cells = cells_range(q.cells, p, b0);
    part = w.part + (size_t)2 * b0 * p->N;
    b.part = part; b.Vz = p->Vz; b.N = p->N; b.P = n;
  }
  DPC_TRY(launch_blur_xy(b, rx, kx, ry, ky, s));
  stage_mark(s);
  if (q.cells) {
    DPC_TRY(launch_pose_bwd_partials(
        pa, b.cells, part, q.g_tr_pc ? q.g_tr_pc + b0 * N3 : nullptr, q.g_points + b0 * N3,
        w.pose_partials + (size_t)b0 * pose_partial_blocks(p->N) * 8, w.counters + b0,
        q.scale ? w.scale_partials + (size_t)b0 * drc_scale_partial_blocks(p->V) : nullptr,
        drc_scale_partial_blocks(p->V), q.g_quat ? q.g_quat + b0 * 4 : nullptr,
        (q.trans && q.g_trans) ? q.g_trans + b0 * 3 : nullptr,
        (q.focal && q.g_focal) ? q.g_focal + b0 : nullptr,
        (q.scale && q.g_scale) ? q.g_scale + b0 : nullptr, s));
    stage_mark(s);
    return DPC_OK;
  }
  // gather + pose adjoint; the last block of each projection also reduces the partials
  DPC_TRY(launch_gather_pose_finalize(
      pa, g_grid, q.g_tr_pc ? q.g_tr_pc + b0 * N3 : nullptr, q.g_points + b0 * N3,
      w.pose_partials + (size_t)b0 * pose_partial_blocks(p->N) * 8, w.counters + b0,
      q.scale ? w.scale_partials + (size_t)b0 * drc_scale_partial_blocks(p->V) : nullptr,
      drc_scale_partial_blocks(p->V), q.g_quat ? q.g_quat + b0 * 4 : nullptr,
      (q.trans && q.g_trans) ? q.g_trans + b0 * 3 : nullptr,
      (q.focal && q.g_focal) ? q.g_focal + b0 : nullptr,
      (q.scale && q.g_scale) ? q.g_scale + b0 : nullptr, s));
  stage_mark(s);
  return DPC_OK;
}

static int project_bwd_impl(const dpc_params *p, const Replica &rep, const float *points,
                    const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, const float *grid_b,
                    const uint32_t *clamp_bits, const void *cells, const float *g_mask,
                    const float *g_depth, const float *g_probs, const float *g_voxels,
                    const float *g_tr_pc, float *g_grid, float *g_points, float *g_quat,
                    float *g_trans, float *g_focal, float *g_scale, void *workspace,
                    size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(grid_b); DPC_REQUIRE(clamp_bits);
  DPC_REQUIRE(g_grid); DPC_REQUIRE(g_points);
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  DPC_TRY(check_ws(p, workspace, workspace_bytes));
  const Workspace w = carve(p, workspace);
  cudaStream_t s = (cudaStream_t)stream;
  void *fast_rays = nullptr;
  if (plane_local_ok(p) && fast_ray_state(p, cells)) {
    if (g_probs || g_voxels) {
      set_error("project_bwd: g_probs / g_voxels given but params.outputs says the forward did "
                "not materialise them");
      return DPC_ERR_ARG;
    }
    fast_rays = const_cast<void *>(cells);
  }
  // The plane gather (backward of the plane-local scatter) runs at every grid size: with the
  // one-tile blur at 128^2 two CTAs share an SM and cover each other's gather (workload B:
  // blur XY adjoint 569 -> 601 us, pose adjoint 140 -> 44 us).  DPC_GATHER128=0 restores the
  // gather from the gradient grid at V = 128 for A/B runs.
  static const bool grid_gather128 = getenv("DPC_GATHER128") && atoi(getenv("DPC_GATHER128")) == 0;
  if (!plane_local_ok(p) || (p->V > 64 && grid_gather128)) cells = nullptr;
  const BwdPtrs q{points, quat, trans, focal, scale, grid_b, clamp_bits, g_mask, g_depth, g_probs,
                  g_voxels, g_tr_pc, g_grid, g_points, g_quat, g_trans, g_focal, g_scale,
                  const_cast<void *>(cells), rep, fast_rays};
  const int chunk = replica_chunk(p, rep);
  Pipeline *pl = chunk < p->P ? get_pipeline() : nullptr;
  int rc = DPC_OK;
  if (!pl) {
    rc = project_bwd_range(p, 0, p->P, q, tx, kx, ty, ky, tz, kz, w, s);
  } else {
    cudaEventRecord(pl->fork, s);
    for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(pl->side[i], pl->fork, 0);
    for (int b0 = 0, c = 0; b0 < p->P && rc == DPC_OK; b0 += chunk, ++c) {
      const int n = p->P - b0 < chunk ? p->P - b0 : chunk;
      rc = project_bwd_range(p, b0, n, q, tx, kx, ty, ky, tz, kz, w, pl->side[c & 1]);
    }
    for (int i = 0; i < 2; ++i) {
      cudaEventRecord(pl->join[i], pl->side[i]);
      cudaStreamWaitEvent(s, pl->join[i], 0);
    }
  }
  return rc;
}

int dpc_project_bwd(const dpc_params *p, const float *points, const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, const float *grid_b,
                    const uint32_t *clamp_bits, const void *cells, const float *g_mask,
                    const float *g_depth, const float *g_probs, const float *g_voxels,
                    const float *g_tr_pc, float *g_grid, float *g_points, float *g_quat,
                    float *g_trans, float *g_focal, float *g_scale, void *workspace,
                    size_t workspace_bytes, void *stream) {
  return project_bwd_impl(p, Replica(), points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz,
                          grid_b, clamp_bits, cells, g_mask, g_depth, g_probs, g_voxels, g_tr_pc,
                          g_grid, g_points, g_quat, g_trans, g_focal, g_scale, workspace,
                          workspace_bytes, stream);
}

// ---- f2: replica-aware projection + point dropout on the device ----------------
int dpc_project_replicated_bwd(const dpc_params *p, int replicas, int N_src, const int32_t *sel,
                    const float *points, const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, const float *grid_b,
                    const uint32_t *clamp_bits, const void *cells, const float *g_mask,
                    const float *g_depth, const float *g_probs, const float *g_voxels,
                    const float *g_tr_pc, float *g_grid, float *g_points_rep, int32_t *inv_scratch,
                    float *g_points, float *g_quat,
                    float *g_trans, float *g_focal, float *g_scale, void *workspace,
                    size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_TRY(check_replica(p, replicas, N_src, sel));
  DPC_REQUIRE(g_points_rep); DPC_REQUIRE(g_points);
  if (sel) DPC_REQUIRE(inv_scratch);
  Replica rep;
  rep.replicas = replicas; rep.N_src = N_src; rep.sel = sel;
  DPC_TRY(project_bwd_impl(p, rep, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz,
                           grid_b, clamp_bits, cells, g_mask, g_depth, g_probs, g_voxels, g_tr_pc,
                           g_grid, g_points_rep, g_quat, g_trans, g_focal, g_scale, workspace,
                           workspace_bytes, stream));
  // autograd of tf_repeat_0 (sum over the replicas) and of the dropout gather, in one pass
  return launch_replica_reduce(g_points_rep, sel, inv_scratch, p->P, replicas, N_src, p->N, 3,
                               g_points, (cudaStream_t)stream);
}

static int check_select_args(int P, int replicas, int N_src, int M, int C) {
  if (P < 1 || replicas < 1 || P % replicas != 0 || N_src < 1 || M < 1 || M > N_src || C < 1 || C > 4) {
    set_error("point selection: need P=%d >= 1, replicas=%d dividing P, 1 <= M=%d <= N_src=%d, "
              "1 <= channels=%d <= 4", P, replicas, M, N_src, C);
    return DPC_ERR_ARG;
  }
  return DPC_OK;
}

int dpc_point_dropout_indices(int P, int N_src, int M, uint64_t seed, int32_t *sel, void *stream) {
  DPC_TRY(check_select_args(P, 1, N_src, M, 1));
  DPC_REQUIRE(sel);
  return launch_dropout_select(P, N_src, M, seed, sel, (cudaStream_t)stream);
}

int dpc_select_points(int P, int replicas, int N_src, int M, int C, const float *points,
                      const int32_t *sel, float *out, void *stream) {
  DPC_TRY(check_select_args(P, replicas, N_src, M, C));
  DPC_REQUIRE(points); DPC_REQUIRE(sel); DPC_REQUIRE(out);
  return launch_select_points(points, sel, P, replicas, N_src, M, C, out, (cudaStream_t)stream);
}

int dpc_replica_reduce(int P, int replicas, int N_src, int M, int C, const float *g_rep,
                       const int32_t *sel, int32_t *inv_scratch, float *g_cloud, void *stream) {
  DPC_TRY(check_select_args(P, replicas, N_src, M, C));
  DPC_REQUIRE(g_rep); DPC_REQUIRE(g_cloud);
  if (sel) DPC_REQUIRE(inv_scratch);
  if (!sel && M != N_src) { set_error("replica_reduce: without a selection M must equal N_src"); return DPC_ERR_ARG; }
  return launch_replica_reduce(g_rep, sel, inv_scratch, P, replicas, N_src, M, C, g_cloud,
                               (cudaStream_t)stream);
}

// ---- f1: candidate-selection projection loss --------------------------------
static int check_loss_args(int BV, int C, int V, int G) {
  if (BV < 1 || C < 1 || C > candidate_loss_max_candidates()) {
    set_error("candidate_loss: BV=%d must be >= 1 and 1 <= num_candidates=%d <= %d", BV, C,
              candidate_loss_max_candidates());
    return DPC_ERR_ARG;
  }
  if (V < 1 || G < V || G % V != 0) {
    set_error("candidate_loss: GT size %d must be a multiple of (and not below) the prediction size %d",
              G, V);
    return DPC_ERR_ARG;
  }
  return DPC_OK;
}

int dpc_candidate_loss_fwd(int BV, int C, int V, int G, const float *gt, const float *pred,
                           const float *weights, float *all_loss, int64_t *min_idx,
                           float *view_loss, void *stream) {
  DPC_TRY(check_loss_args(BV, C, V, G));
  DPC_REQUIRE(gt); DPC_REQUIRE(pred); DPC_REQUIRE(all_loss); DPC_REQUIRE(min_idx);
  DPC_REQUIRE(view_loss);
  return launch_candidate_loss_fwd(gt, pred, weights, BV, C, V, G, all_loss, (long long *)min_idx,
                                   view_loss, (cudaStream_t)stream);
}

int dpc_candidate_loss_bwd(int BV, int C, int V, int G, const float *gt, const float *pred,
                           const float *weights, const int64_t *min_idx, const float *upstream,
                           float coeff, float *g_pred, void *stream) {
  DPC_TRY(check_loss_args(BV, C, V, G));
  DPC_REQUIRE(gt); DPC_REQUIRE(pred); DPC_REQUIRE(min_idx); DPC_REQUIRE(g_pred);
  return launch_candidate_loss_bwd(gt, pred, weights, (const long long *)min_idx, upstream, coeff,
                                   BV, C, V, G, g_pred, (cudaStream_t)stream);
}

// ---- fused renderer + candidate-selection loss (f2 -> the path -> f1) -----------
// models/model_pc_to.py:302-331 (replication, dropout, projection) + :339-385, 410-440 (loss).
// The gradient of the loss reaches the projection of the WINNING candidate of every view only
// (one-hot mask, :425-430); the losing candidates' gradients are exactly zero.  When the forward
// saved the fast ray state the backward chain therefore runs over the BV winners (chain slots,
// bmap = winners) and the ray kernel builds dL/dmask on the fly; otherwise dL/dmask is written
// out for all P projections and the general backward runs.
static bool render_winner_only(const dpc_params *p, const void *cells) {
  return plane_local_ok(p) && fast_ray_state(p, cells);
}

static int check_render_args(const dpc_params *p, int replicas, int C, int G) {
  if (C < 1 || replicas % C != 0) {
    set_error("render_loss: replicas=%d must be views x num_candidates=%d", replicas, C);
    return DPC_ERR_ARG;
  }
  if (p->outputs != 0) {
    set_error("render_loss: params.outputs must be 0 (no voxels / probs)");
    return DPC_ERR_ARG;
  }
  return check_loss_args(p->P / C, C, p->V, G);
}

int dpc_render_loss_slots(const dpc_params *p, int num_candidates, int have_cells, int scatter_mode) {
  if (!p || p->P < 1 || num_candidates < 1 || p->P % num_candidates) return 0;
  (void)scatter_mode;     // both scatter modes save the same state
  return render_winner_only(p, have_cells ? (const void *)p : nullptr)
             ? p->P / num_candidates : p->P;
}

int dpc_render_loss_fwd(const dpc_params *p, int replicas, int N_src, const int32_t *sel,
                        const float *points, const float *quat, const float *trans,
                        const float *focal, const float *scale, const float *tx, int kx,
                        const float *ty, int ky, const float *tz, int kz, int scatter_mode,
                        int num_candidates, int G, const float *gt, const float *weights,
                        float weight_scale, float *grid_b, uint32_t *clamp_bits, void *cells,
                        float *mask, float *all_loss, int64_t *min_idx, float *view_loss,
                        float *loss, int32_t *winners, float *kcoef, void *workspace,
                        size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_TRY(check_replica(p, replicas, N_src, sel));
  DPC_TRY(check_render_args(p, replicas, num_candidates, G));
  DPC_REQUIRE(gt); DPC_REQUIRE(all_loss); DPC_REQUIRE(min_idx); DPC_REQUIRE(view_loss);
  DPC_REQUIRE(loss); DPC_REQUIRE(winners); DPC_REQUIRE(kcoef);
  Replica rep;
  rep.replicas = replicas; rep.N_src = N_src; rep.sel = sel;
  DPC_TRY(project_fwd_impl(p, rep, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz,
                           scatter_mode, nullptr, grid_b, clamp_bits, cells, mask, nullptr, nullptr,
                           nullptr, workspace, workspace_bytes, stream));
  const int BV = p->P / num_candidates;
  const float coeff = weight_scale / (float)BV;
  cudaStream_t s = (cudaStream_t)stream;
  DPC_TRY(launch_candidate_loss_fwd(gt, mask, weights, BV, num_candidates, p->V, G, all_loss,
                                    (long long *)min_idx, view_loss, s, winners, kcoef, coeff));
  return launch_loss_total(view_loss, BV, coeff, loss, s);
}

// chain slots [j0, j0 + n) of the winner-only backward on stream s
static int render_bwd_range(const dpc_params *p, int j0, int n, const BwdPtrs &q, const LossGrad &lg0,
                            const float *tx, int kx, const float *ty, int ky, const float *tz,
                            int kz, const Workspace &w, cudaStream_t s) {
  dpc_params sp = *p;
  sp.P = n;                                        // slots of this range
  const size_t N3 = (size_t)p->N * 3, G = (size_t)p->Vz * p->V * p->V;
  const int spb = drc_scale_partial_blocks(p->V), ppb = pose_partial_blocks(p->N);
  LossGrad lg = lg0;
  lg.bmap += j0;
  // everything that belongs to a projection (inputs, saved state, pose gradients) is passed
  // whole-batch and addressed through bmap; everything that belongs to a slot is offset here
  PoseArgs pa = pose_args(&sp, q.points, q.quat, q.trans, q.focal, q.rep);
  pa.bmap = lg.bmap;
  float *g_grid = q.g_grid + j0 * G;
  DrcArgs da = drc_args(&sp, q.grid_b, q.scale);
  da.P_total = p->P;
  set_ray_state(da, p, q.fast_rays, 0);
  stage_mark(s);
  DPC_TRY(launch_drc_blurz_bwd(da, tz, kz, nullptr, nullptr, nullptr, nullptr, g_grid,
                               w.scale_partials + (size_t)j0 * spb, w.counters + j0, n, s, &lg));
  stage_mark(s);
  BlurXYArgs b;
  b.src = g_grid; b.dst = g_grid; b.bits_out = nullptr; b.bits_in = q.bits;
  b.planes = n * p->Vz; b.V = p->V; b.clamp_in = false;
  float rx[DPC_MAX_TAPS], ry[DPC_MAX_TAPS];
  for (int i = 0; i < kx; ++i) rx[i] = tx[kx - 1 - i];
  for (int i = 0; i < ky; ++i) ry[i] = ty[ky - 1 - i];
  b.cells = cells_view(q.cells, p->P, p->N, p->Vz);
  float4 *part = w.part + (size_t)2 * j0 * p->N;
  b.part = part; b.Vz = p->Vz; b.N = p->N; b.P = n; b.bmap = lg.bmap;
  DPC_TRY(launch_blur_xy(b, rx, kx, ry, ky, s));
  stage_mark(s);
  DPC_TRY(launch_pose_bwd_partials(
      pa, b.cells, part, nullptr, q.g_points + j0 * N3, w.pose_partials + (size_t)j0 * ppb * 8,
      w.counters + j0, q.scale ? w.scale_partials + (size_t)j0 * spb : nullptr, spb, q.g_quat,
      (q.trans && q.g_trans) ? q.g_trans : nullptr, (q.focal && q.g_focal) ? q.g_focal : nullptr,
      (q.scale && q.g_scale) ? q.g_scale : nullptr, s, lg.C));
  stage_mark(s);
  return DPC_OK;
}

int dpc_render_loss_bwd(const dpc_params *p, int replicas, int N_src, const int32_t *sel,
                        const float *points, const float *quat, const float *trans,
                        const float *focal, const float *scale, const float *tx, int kx,
                        const float *ty, int ky, const float *tz, int kz, int scatter_mode,
                        int num_candidates, int G, const float *gt, const float *weights,
                        float weight_scale, const float *grid_b, const uint32_t *clamp_bits,
                        const void *cells, const float *mask, const int64_t *min_idx,
                        const int32_t *winners, const float *kcoef, const float *upstream,
                        float *g_grid, float *g_points_rep, int32_t *inv_scratch,
                        float *g_mask_scratch, float *g_points, float *g_quat, float *g_trans,
                        float *g_focal, float *g_scale, void *workspace, size_t workspace_bytes,
                        void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_TRY(check_replica(p, replicas, N_src, sel));
  DPC_TRY(check_render_args(p, replicas, num_candidates, G));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(grid_b); DPC_REQUIRE(clamp_bits);
  DPC_REQUIRE(gt); DPC_REQUIRE(mask); DPC_REQUIRE(min_idx); DPC_REQUIRE(winners); DPC_REQUIRE(kcoef);
  DPC_REQUIRE(g_grid); DPC_REQUIRE(g_points_rep); DPC_REQUIRE(g_points); DPC_REQUIRE(g_quat);
  if (sel) DPC_REQUIRE(inv_scratch);
  const int C = num_candidates, BV = p->P / C;
  cudaStream_t s = (cudaStream_t)stream;
  (void)scatter_mode;     // the saved state does not depend on it (see fast_ray_state)
  if (!render_winner_only(p, cells)) {
    // general saved state: dL/dmask for all P projections (zeros for the losers), then the
    // replica-aware backward
    DPC_REQUIRE(g_mask_scratch);
    DPC_TRY(launch_candidate_loss_bwd(gt, mask, weights, (const long long *)min_idx, upstream,
                                      weight_scale / (float)BV, BV, C, p->V, G, g_mask_scratch, s));
    return dpc_project_replicated_bwd(p, replicas, N_src, sel, points, quat, trans, focal, scale,
                                      tx, kx, ty, ky, tz, kz, grid_b, clamp_bits, cells,
                                      g_mask_scratch, nullptr, nullptr, nullptr, nullptr, g_grid,
                                      g_points_rep, inv_scratch, g_points, g_quat, g_trans, g_focal,
                                      g_scale, workspace, workspace_bytes, stream);
  }
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  DPC_TRY(check_ws(p, workspace, workspace_bytes));
  const Workspace w = carve(p, workspace);
  // (the pose gradients of the losing candidates, exact zeros, are written by the winners'
  // finalize in the last kernel of the chain: no memset nodes on the critical path)
  Replica rep;
  rep.replicas = replicas; rep.N_src = N_src; rep.sel = sel;
  LossGrad lg;
  lg.bmap = winners; lg.gt = gt; lg.pred = mask; lg.kcoef = kcoef; lg.upstream = upstream;
  lg.G = G; lg.C = C;
  // one view per cloud and no dropout: slot j IS cloud j and the winner's point gradient is the
  // cloud's -- the pose adjoint writes it in place and the reduction kernel is not launched
  const bool direct = (replicas / C == 1) && !sel;
  const BwdPtrs q{points, quat, trans, focal, scale, grid_b, clamp_bits, nullptr, nullptr, nullptr,
                  nullptr, nullptr, g_grid, direct ? g_points : g_points_rep, g_quat, g_trans, g_focal,
                  g_scale, const_cast<void *>(cells), rep, const_cast<void *>(cells)};
  // the winners as two half-chains on the two internal streams from 64 slots on (as the batch is)
  dpc_params slots = *p;
  slots.P = BV;
  const int chunk = chunk_size(&slots);
  Pipeline *pl = chunk < BV ? get_pipeline() : nullptr;
  int rc = DPC_OK;
  if (!pl) {
    rc = render_bwd_range(p, 0, BV, q, lg, tx, kx, ty, ky, tz, kz, w, s);
  } else {
    cudaEventRecord(pl->fork, s);
    for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(pl->side[i], pl->fork, 0);
    for (int j0 = 0, c = 0; j0 < BV && rc == DPC_OK; j0 += chunk, ++c) {
      const int n = BV - j0 < chunk ? BV - j0 : chunk;
      rc = render_bwd_range(p, j0, n, q, lg, tx, kx, ty, ky, tz, kz, w, pl->side[c & 1]);
    }
    for (int i = 0; i < 2; ++i) {
      cudaEventRecord(pl->join[i], pl->side[i]);
      cudaStreamWaitEvent(s, pl->join[i], 0);
    }
  }
  DPC_TRY(rc);
  if (direct) return DPC_OK;
  // cloud gradient = sum over the views of a cloud of its winners' point gradients, routed
  // through their dropout selections (autograd of tf_repeat_0 and of the gather)
  return launch_replica_reduce(g_points_rep, sel, inv_scratch, BV, replicas / C, N_src, p->N, 3,
                               g_points, s, winners);
}

// ---- f4: nearest neighbour of the Chamfer evaluation ---------------------------
int dpc_point_cloud_distance(int N, int M, const float *src, const float *tgt, float *proj,
                             float *min_dist, int64_t *idx, void *workspace,
                             size_t workspace_bytes, void *stream) {
  if (N < 1 || M < 1) { set_error("point_cloud_distance: N=%d and M=%d must be >= 1", N, M); return DPC_ERR_ARG; }
  DPC_REQUIRE(src); DPC_REQUIRE(tgt); DPC_REQUIRE(proj); DPC_REQUIRE(min_dist); DPC_REQUIRE(idx);
  if (!workspace || workspace_bytes < (size_t)N * 8) {
    set_error("point_cloud_distance: workspace needs %zu bytes, got %zu", (size_t)N * 8, workspace_bytes);
    return DPC_ERR_WORKSPACE;
  }
  return launch_nn_search(src, N, tgt, M, (unsigned long long *)workspace, proj, min_dist,
                          (long long *)idx, (cudaStream_t)stream);
}

int dpc_project_profile(const dpc_params *p, const float *points, const float *quat,
                        const float *trans, const float *focal, const float *scale,
                        const float *tx, int kx, const float *ty, int ky, const float *tz, int kz,
                        int scatter_mode, float *tr_pc, float *grid_b, uint32_t *clamp_bits,
                        void *cells, float *mask, float *depth, const float *g_mask,
                        const float *g_depth,
                        float *g_grid, float *g_points, float *g_quat, float *g_trans,
                        float *g_focal, float *g_scale, void *workspace, size_t workspace_bytes,
                        void *stream, int iters, float *stage_ms_host) {
  if (iters < 1 || !stage_ms_host) { set_error("profile: iters >= 1 and stage_ms_host required"); return DPC_ERR_ARG; }
  cudaStream_t s = (cudaStream_t)stream;
  cudaEvent_t ev[DPC_PROFILE_STAGES + 2];
  for (auto &e : ev) cudaEventCreate(&e);
  double acc[DPC_PROFILE_STAGES] = {0};
  int rc = DPC_OK;
  for (int it = 0; it < iters && rc == DPC_OK; ++it) {
    tl_stage_events = ev;
    tl_stage_idx = 0;
    tl_force_single = true;   // whole batch per kernel, on the caller's stream
    rc = dpc_project_fwd(p, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz, scatter_mode,
                         tr_pc, grid_b, clamp_bits, cells, mask, depth, nullptr, nullptr, workspace,
                         workspace_bytes, stream);
    const int nf = tl_stage_idx;  // 5 events: start + 4 stages
    if (rc == DPC_OK)
      rc = dpc_project_bwd(p, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz, grid_b,
                           clamp_bits, cells, g_mask,
                           g_depth, nullptr, nullptr, nullptr, g_grid, g_points, g_quat, g_trans,
                           g_focal, g_scale, workspace, workspace_bytes, stream);
    const int nb = tl_stage_idx;  // + 5 events
    tl_stage_events = nullptr;
    tl_force_single = false;
    if (rc != DPC_OK) break;
    if (cudaStreamSynchronize(s) != cudaSuccess) { rc = check_launch("profile sync"); break; }
    int k = 0;
    for (int i = 1; i < nf; ++i, ++k) { float ms; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); acc[k] += ms; }
    for (int i = nf + 1; i < nb; ++i, ++k) { float ms; cudaEventElapsedTime(&ms, ev[i - 1], ev[i]); acc[k] += ms; }
  }
  for (auto &e : ev) cudaEventDestroy(e);
  for (int k = 0; k < DPC_PROFILE_STAGES; ++k) stage_ms_host[k] = (float)(acc[k] / iters);
  return rc;
}

}  // extern "C"
