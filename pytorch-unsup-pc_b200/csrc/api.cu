// C ABI (include/dpc_b200.h): argument validation, workspace carving and the
// stream-ordered launch sequences.  No allocation, no host synchronisation.
#include <cstdarg>
#include <cstdio>
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
static thread_local cudaEvent_t *tl_stage_events = nullptr;
static thread_local int tl_stage_idx = 0;
static inline void stage_mark(cudaStream_t s) {
  if (tl_stage_events) cudaEventRecord(tl_stage_events[tl_stage_idx++], s);
}

struct Workspace {
  double *pose_partials;
  float *scale_partials;
  void *sorted;
  size_t sorted_bytes;
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
  w.sorted = (void *)(c + off);
  w.sorted_bytes = sorted_workspace_bytes(p->P, p->N, p->Vz, p->V);
  off += align256(w.sorted_bytes);
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

static PoseArgs pose_args(const dpc_params *p, const float *points, const float *quat,
                          const float *trans, const float *focal) {
  PoseArgs a;
  a.points = points; a.quat = quat; a.trans = trans; a.focal = focal;
  a.P = p->P; a.N = p->N; a.Vz = p->Vz; a.V = p->V;
  a.cam_dist = p->camera_distance;
  a.focal_const = p->focal_length;
  return a;
}

static DrcArgs drc_args(const dpc_params *p, const float *grid, const float *scale) {
  DrcArgs a;
  a.grid = grid; a.scale = scale;
  a.P = p->P; a.Vz = p->Vz; a.V = p->V;
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
                              nullptr, g_voxels, nullptr, (cudaStream_t)stream);
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

int dpc_project_fwd(const dpc_params *p, const float *points, const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, float *mask, float *depth,
                    float *voxels, float *probs, void *workspace, size_t workspace_bytes,
                    void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(grid_b); DPC_REQUIRE(clamp_bits);
  DPC_REQUIRE(mask);
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  cudaStream_t s = (cudaStream_t)stream;
  const PoseArgs pa = pose_args(p, points, quat, trans, focal);
  stage_mark(s);
  if (scatter_mode == DPC_SCATTER_SORTED) {
    DPC_TRY(check_ws(p, workspace, workspace_bytes));
    const Workspace w = carve(p, workspace);
    stage_mark(s);
    DPC_TRY(launch_scatter_sorted(&pa, nullptr, p->P, p->N, p->Vz, p->V, tr_pc, grid_b, w.sorted,
                                  w.sorted_bytes, s));
  } else if (scatter_mode == DPC_SCATTER_ATOMIC) {
    if (cudaMemsetAsync(grid_b, 0, grid_bytes(p), s) != cudaSuccess) return check_launch("memset");
    stage_mark(s);
    DPC_TRY(launch_pose_scatter(pa, tr_pc, grid_b, s));
  } else {
    set_error("unknown scatter mode %d", scatter_mode);
    return DPC_ERR_ARG;
  }
  // clamp(raw,0,1) + raw<=1 mask + blur X + blur Y, in place (identity taps when kernel=None)
  BlurXYArgs b;
  b.src = grid_b; b.dst = grid_b; b.bits_out = clamp_bits; b.bits_in = nullptr;
  b.planes = p->P * p->Vz; b.V = p->V; b.clamp_in = true;
  stage_mark(s);
  DPC_TRY(launch_blur_xy(b, tx, kx, ty, ky, s));
  stage_mark(s);
  // blur Z + scale + clip + DRC; the blurred occupancy overwrites grid in place (saved for bwd)
  DPC_TRY(launch_blurz_drc_fwd(drc_args(p, grid_b, scale), tz, kz, grid_b, mask, depth, voxels,
                               probs, s));
  stage_mark(s);
  return DPC_OK;
}

int dpc_project_bwd(const dpc_params *p, const float *points, const float *quat, const float *trans,
                    const float *focal, const float *scale, const float *tx, int kx,
                    const float *ty, int ky, const float *tz, int kz, const float *grid_b,
                    const uint32_t *clamp_bits, const float *g_mask, const float *g_depth,
                    const float *g_probs, const float *g_voxels, const float *g_tr_pc,
                    float *g_grid, float *g_points, float *g_quat, float *g_trans, float *g_focal,
                    float *g_scale, void *workspace, size_t workspace_bytes, void *stream) {
  DPC_TRY(check_params(p, true));
  DPC_REQUIRE(points); DPC_REQUIRE(quat); DPC_REQUIRE(grid_b); DPC_REQUIRE(clamp_bits);
  DPC_REQUIRE(g_grid); DPC_REQUIRE(g_points);
  DPC_TRY(check_taps(tx, kx, "taps_x")); DPC_TRY(check_taps(ty, ky, "taps_y"));
  DPC_TRY(check_taps(tz, kz, "taps_z"));
  DPC_TRY(check_ws(p, workspace, workspace_bytes));
  const Workspace w = carve(p, workspace);
  cudaStream_t s = (cudaStream_t)stream;
  const PoseArgs pa = pose_args(p, points, quat, trans, focal);
  stage_mark(s);
  DPC_TRY(launch_drc_blurz_bwd(drc_args(p, grid_b, scale), tz, kz, g_mask, g_depth, g_probs,
                               g_voxels, g_grid, w.scale_partials, s));
  stage_mark(s);
  BlurXYArgs b;
  b.src = g_grid; b.dst = g_grid; b.bits_out = nullptr; b.bits_in = clamp_bits;
  b.planes = p->P * p->Vz; b.V = p->V; b.clamp_in = false;
  // adjoint of a correlation = correlation with the reversed taps
  float rx[DPC_MAX_TAPS], ry[DPC_MAX_TAPS];
  for (int i = 0; i < kx; ++i) rx[i] = tx[kx - 1 - i];
  for (int i = 0; i < ky; ++i) ry[i] = ty[ky - 1 - i];
  DPC_TRY(launch_blur_xy(b, rx, kx, ry, ky, s));
  stage_mark(s);
  DPC_TRY(launch_gather_pose_bwd(pa, g_grid, g_tr_pc, g_points, w.pose_partials, s));
  stage_mark(s);
  DPC_TRY(launch_finalize(pa, w.pose_partials, pose_partial_blocks(p->N),
                          scale ? w.scale_partials : nullptr, drc_scale_partial_blocks(p->V),
                          g_quat, trans ? g_trans : nullptr, focal ? g_focal : nullptr,
                          scale ? g_scale : nullptr, s));
  stage_mark(s);
  return DPC_OK;
}

int dpc_project_profile(const dpc_params *p, const float *points, const float *quat,
                        const float *trans, const float *focal, const float *scale,
                        const float *tx, int kx, const float *ty, int ky, const float *tz, int kz,
                        int scatter_mode, float *tr_pc, float *grid_b, uint32_t *clamp_bits,
                        float *mask, float *depth, const float *g_mask, const float *g_depth,
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
    rc = dpc_project_fwd(p, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz, scatter_mode,
                         tr_pc, grid_b, clamp_bits, mask, depth, nullptr, nullptr, workspace,
                         workspace_bytes, stream);
    const int nf = tl_stage_idx;  // 5 events: start + 4 stages
    if (rc == DPC_OK)
      rc = dpc_project_bwd(p, points, quat, trans, focal, scale, tx, kx, ty, ky, tz, kz, grid_b,
                           clamp_bits, g_mask, g_depth, nullptr, nullptr, nullptr, g_grid, g_points,
                           g_quat, g_trans, g_focal, g_scale, workspace, workspace_bytes, stream);
    const int nb = tl_stage_idx;  // + 5 events
    tl_stage_events = nullptr;
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
