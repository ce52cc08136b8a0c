/*
 * dpc_b200.h -- C ABI of the B200-native differentiable point-cloud projection.
 *
 * The reference (NiteshBharadwaj/pytorch-unsup-pc) has NO native layer: the
 * path is plain Python/torch (dpc/util/point_cloud_to.py, drc.py,
 * gauss_kernel.py, quaternion.py).  This header is therefore the boundary a
 * maintainer binds with ctypes; each entry point names
 * the reference function(s) (file:line under /root/reference/dpc) it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *    tensors are dense, row-major, fp32 unless stated;
 *  - the caller owns every buffer (outputs, saved state, workspace); the
 *    library never allocates, never synchronises the host and keeps no
 *    mutable global state besides the thread-local error string;
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*) of the
 *    CURRENT device;
 *  - return 0 on success, a negative dpc_status otherwise;
 *    dpc_last_error() describes the last failure on the calling thread;
 *  - "NULL ok" marks optional arguments.
 *
 * Layouts (P projections, N points, grid Vz x V x V, V in {32,64,128}):
 *   points/tr_pc/g_points  [P,N,3]      quat [P,4]   trans [P,3]
 *   focal/scale            [P]          grid [P,Vz,V,V]
 *   mask/depth             [P,V,V]      probs [Vz+1,P,V,V]
 *   clamp bits             [P,Vz,V,V/32] uint32, bit x%32 of word x/32
 *   taps                   HOST pointers, odd length <= DPC_MAX_TAPS
 */
#ifndef DPC_B200_H
#define DPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DPC_API __attribute__((visibility("default")))
#else
#define DPC_API
#endif

#define DPC_B200_VERSION 140
#define DPC_MAX_TAPS 21

typedef enum {
  DPC_OK = 0,
  DPC_ERR_ARG = -1,         /* bad shape / NULL / unsupported size  */
  DPC_ERR_CUDA = -2,        /* CUDA runtime error at launch         */
  DPC_ERR_WORKSPACE = -3    /* workspace too small                  */
} dpc_status;

/* Geometry + the config keys the path reads (default_config.yaml:72-83). */
typedef struct {
  int32_t P, N;             /* projections, points per cloud                    */
  int32_t Vz, V;            /* vox_size_z (or vox_size when -1), vox_size       */
  double camera_distance;   /* cfg.camera_distance                              */
  double focal_length;      /* cfg.focal_length, used when focal == NULL        */
  double max_depth;         /* cfg.max_depth                                    */
  double drc_clip;          /* cfg.drc_logsum_clip_val                          */
  int32_t drc_logsum;       /* cfg.drc_logsum: 1 = clip + exp(clip) end factors */
  int32_t flip_y;           /* 1 = apply the Y flips of point_cloud_to.py:239,242 */
  int32_t outputs;          /* DPC_OUT_* flags: the optional outputs dpc_project_fwd
                               materialises; dpc_project_bwd must get the same value
                               (it selects the layout of the saved ray state)        */
} dpc_params;

#define DPC_OUT_VOXELS 1    /* "voxels"    [P,Vz,V,V]   (point_cloud_to.py:222) */
#define DPC_OUT_PROBS 2     /* "drc_probs" [Vz+1,P,V,V] (point_cloud_to.py:239) */

/* Flags for dpc_project_fwd / scatter mode. */
#define DPC_SCATTER_ATOMIC 0
#define DPC_SCATTER_SORTED 1   /* deterministic sort-then-segment */

DPC_API int dpc_version(void);
DPC_API const char *dpc_last_error(void);

/* Bytes of workspace the calls below need for this geometry (one buffer is
 * shared by all of them; 256-byte aligned). */
DPC_API size_t dpc_workspace_bytes(const dpc_params *p);

/* Bytes of the state besides grid_b / clamp_bits that dpc_project_fwd saves for
 * dpc_project_bwd: the per-point cell records (z cell byte + two 16-byte
 * records per point) and the ray-transmittance checkpoints of the DRC kernels
 * (Vz/8 V x V planes per projection). */
DPC_API size_t dpc_cells_bytes(const dpc_params *p);

/* ---- a1+a2: quaternion.py:110-132 quaternion_rotate +
 *      point_cloud_to.py:118-178 pc_perspective_transform ------------------ */
DPC_API int dpc_pose_fwd(const dpc_params *p, const float *points, const float *quat,
                 const float *trans /*NULL ok*/, const float *focal /*NULL ok*/,
                 float *tr_pc, void *stream);
/* Adjoint: g_tr_pc -> g_points [P,N,3], g_quat [P,4], g_trans [P,3] (NULL ok),
 * g_focal [P] (NULL ok). */
DPC_API int dpc_pose_bwd(const dpc_params *p, const float *points, const float *quat,
                 const float *trans, const float *focal, const float *g_tr_pc,
                 float *g_points, float *g_quat, float *g_trans, float *g_focal,
                 void *workspace, size_t workspace_bytes, void *stream);

/* ---- a3: point_cloud_to.py:10-87 pointcloud2voxels3d_fast ----------------
 * Trilinear scatter of already-transformed points.  `grid` is overwritten.
 * mode DPC_SCATTER_SORTED is bit-exact run to run. */
DPC_API int dpc_scatter_fwd(const dpc_params *p, const float *tr_pc, float *grid, int mode,
                    void *workspace, size_t workspace_bytes, void *stream);
/* Adjoint: atomic-free gather of g_grid at the 8 corners -> g_tr_pc. */
DPC_API int dpc_scatter_bwd(const dpc_params *p, const float *tr_pc, const float *g_grid,
                    float *g_tr_pc, void *stream);

/* ---- a6: point_cloud_to.py:90-103 smoothen_voxels3d ----------------------
 * Zero-padded 'same' separable cross-correlation X, Y, Z.  src == dst allowed.
 * Self-adjoint for symmetric taps, so it is its own backward. */
DPC_API int dpc_blur3d(const dpc_params *p, const float *src, float *dst,
               const float *taps_x_host, int kx, const float *taps_y_host, int ky,
               const float *taps_z_host, int kz, void *stream);

/* ---- a8-a11: drc.py:48-129 drc_projection, :145-160 drc_depth_projection --
 * voxels [P,Vz,V,V] -> mask [P,V,V], probs [Vz+1,P,V,V] (NULL ok),
 * depth [P,V,V] (NULL ok).  No blur, no scaling; honours p->flip_y. */
DPC_API int dpc_drc_fwd(const dpc_params *p, const float *voxels, float *mask, float *depth,
                float *probs, void *stream);
DPC_API int dpc_drc_bwd(const dpc_params *p, const float *voxels, const float *g_mask /*NULL ok*/,
                const float *g_depth /*NULL ok*/, const float *g_probs /*NULL ok*/,
                float *g_voxels, void *workspace, size_t workspace_bytes, void *stream);
/* drc.py:152-160 on stored probabilities, and its adjoint. */
DPC_API int dpc_depth_from_probs_fwd(const dpc_params *p, const float *probs, float *depth, void *stream);
DPC_API int dpc_depth_from_probs_bwd(const dpc_params *p, const float *g_depth, float *g_probs,
                             void *stream);

/* ---- a12+a13: point_cloud_to.py:191-263 pointcloud_project_fast ----------
 * The whole path in one call: pose -> scatter -> clamp -> blur XY (in place)
 * -> blur Z (in place) + scale + clip + DRC ray march (+ Y flips).
 * Saved for backward (opaque to the caller): grid_b [P,Vz,V,V] (the blurred
 * occupancy B before scaling; or, on the plane-local path with a cubic grid, no
 * optional outputs and drc_logsum, the clipped ray occupancy v = clip(s B) with
 * the clamp gate in its sign bit), clamp_bits (raw <= 1 mask) and cells
 * (dpc_cells_bytes; NULL ok).
 * With cells != NULL the plane-local path runs: the raw grid never exists in
 * global memory -- a pose kernel writes the cell records and every Z-plane is
 * built in shared memory by the blur kernel from the points that touch it
 * (forward), and gathered from shared memory (backward).  DPC_SCATTER_ATOMIC
 * accumulates a plane with order-free fixed-point shared-memory atomics;
 * DPC_SCATTER_SORTED sorts the records by grid row first (it needs the
 * workspace in the forward too) and sums every plane row in a fixed order in
 * fp32 -- the same sums, bit for bit, as dpc_scatter_fwd's sorted mode.  Both
 * save the same state, so the backward does not depend on the mode.
 * With cells == NULL the grid is scattered in global memory first (atomics,
 * or sort-then-segment).  Pass the same cells pointer (or NULL) to
 * dpc_project_bwd.
 * voxels/probs are written only when non-NULL.
 * ntaps == 0 means kernel=None (no blur). */
DPC_API int dpc_project_fwd(const dpc_params *p, const float *points, const float *quat,
                    const float *trans /*NULL ok*/, const float *focal /*NULL ok*/,
                    const float *scale /*NULL ok*/,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, void *cells /*NULL ok*/,
                    float *mask, float *depth,
                    float *voxels /*NULL ok*/, float *probs /*NULL ok*/,
                    void *workspace, size_t workspace_bytes, void *stream);

/* How many chunks of projections dpc_project_fwd / dpc_project_bwd split this batch into (each
 * chunk is one set of kernel launches on one of two internal streams that fork from and join
 * `stream`): 2 half-batches for P >= 64, else 1; DPC_CHUNK=<projections> (environment)
 * overrides.  bench.py counts its launches with it.  Inside a chunk every kernel but the first
 * is a programmatic dependent launch (it waits for its predecessor on the device with
 * griddepcontrol.wait); DPC_PDL=0 (environment) launches them as plain stream-ordered kernels. */
DPC_API int dpc_project_chunks(const dpc_params *p);

/* Kernel launches of one chunk's forward + backward on the default (plane-local) path: 6 when
 * the pose and the z-binning of the points run as one cluster kernel (N <= 16384), 7 otherwise.
 * bench.py's gpu_launches = this x dpc_project_chunks x steps. */
DPC_API int dpc_project_kernels_per_chunk(const dpc_params *p);

/* Backward of the whole path.  g_grid is a [P,Vz,V,V] scratch buffer.
 * Upstream grads: g_mask, g_depth [P,V,V]; g_probs, g_voxels, g_tr_pc optional.
 * Outputs: g_points [P,N,3], g_quat [P,4]; g_trans [P,3], g_focal [P],
 * g_scale [P] written when the matching input was given and the pointer is
 * non-NULL. */
DPC_API int dpc_project_bwd(const dpc_params *p, const float *points, const float *quat,
                    const float *trans, const float *focal, const float *scale,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz,
                    const float *grid_b, const uint32_t *clamp_bits, const void *cells /*NULL ok*/,
                    const float *g_mask /*NULL ok*/, const float *g_depth /*NULL ok*/,
                    const float *g_probs /*NULL ok*/, const float *g_voxels /*NULL ok*/,
                    const float *g_tr_pc /*NULL ok*/,
                    float *g_grid, float *g_points, float *g_quat, float *g_trans,
                    float *g_focal, float *g_scale,
                    void *workspace, size_t workspace_bytes, void *stream);

/* ---- f1 (next row): models/model_pc_to.py:339-385 add_proj_loss (candidate
 * branch) + :410-440 proj_loss_pose_candidates ------------------------------
 * gt [BV,G,G] ground-truth masks (G a multiple of V: AvgPool2d(G/V) is fused),
 * pred [BV*C,V,V] candidate masks (candidates of a view adjacent), weights [BV]
 * valid_samples (NULL ok).  Outputs: all_loss [BV*C] = sum((gt-pred)^2) per
 * candidate, min_idx [BV] = argmin over the candidates (first minimum),
 * view_loss [BV] = weights^2 * all_loss[min]; proj_loss = sum(view_loss) / BV. */
DPC_API int dpc_candidate_loss_fwd(int BV, int C, int V, int G, const float *gt, const float *pred,
                           const float *weights /*NULL ok*/, float *all_loss, int64_t *min_idx,
                           float *view_loss, void *stream);
/* g_pred [BV*C,V,V] = coeff * upstream * d(sum view_loss)/d pred: zero for the
 * losing candidates.  upstream: device scalar (NULL = 1); coeff: host scalar
 * (weight_scale / BV). */
DPC_API int dpc_candidate_loss_bwd(int BV, int C, int V, int G, const float *gt, const float *pred,
                           const float *weights /*NULL ok*/, const int64_t *min_idx,
                           const float *upstream /*NULL ok*/, float coeff, float *g_pred,
                           void *stream);

/* ---- f2 (next row): replica-aware projection + point dropout on the device --
 * Reference: models/model_pc_to.py:47-56 tf_repeat_0 and :302-306 (every cloud
 * is materialised step_size x num_candidates times before the projection),
 * :254-258 + util/point_cloud_to.py:269-295 pc_point_dropout (host numpy
 * sampler np.random.choice(N, M, replace=False) per replica + index gather),
 * and the autograd of both (index backward, sum over replicas).
 *
 * points is the UN-replicated cloud tensor [P/replicas, N_src, 3]; projection b
 * reads cloud b / replicas (tf_repeat_0 order: the replicas of a cloud are
 * adjacent).  sel [P,N] int32 (NULL ok) lists, per projection, the N cloud
 * points that survive the dropout -- DISTINCT indices in [0, N_src), as the
 * reference samples them; without it N == N_src.  Everything else is as in
 * dpc_project_fwd (p->N = points per projection after dropout). */
DPC_API int dpc_project_replicated_fwd(const dpc_params *p, int replicas, int N_src,
                    const int32_t *sel /*NULL ok*/,
                    const float *points, const float *quat,
                    const float *trans /*NULL ok*/, const float *focal /*NULL ok*/,
                    const float *scale /*NULL ok*/,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, void *cells /*NULL ok*/,
                    float *mask, float *depth,
                    float *voxels /*NULL ok*/, float *probs /*NULL ok*/,
                    void *workspace, size_t workspace_bytes, void *stream);
/* As dpc_project_bwd, but g_points is the gradient of the cloud tensor
 * [P/replicas, N_src, 3]: the per-replica point gradients (g_points_rep
 * [P,N,3], scratch) are summed over the replicas in replica order and routed
 * through the selection (dropped points get 0) -- deterministic, no atomics.
 * inv_scratch: [P,N_src] int32, needed when sel != NULL. */
DPC_API int dpc_project_replicated_bwd(const dpc_params *p, int replicas, int N_src,
                    const int32_t *sel /*NULL ok*/,
                    const float *points, const float *quat,
                    const float *trans, const float *focal, const float *scale,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz,
                    const float *grid_b, const uint32_t *clamp_bits, const void *cells /*NULL ok*/,
                    const float *g_mask /*NULL ok*/, const float *g_depth /*NULL ok*/,
                    const float *g_probs /*NULL ok*/, const float *g_voxels /*NULL ok*/,
                    const float *g_tr_pc /*NULL ok*/,
                    float *g_grid, float *g_points_rep, int32_t *inv_scratch /*NULL ok*/,
                    float *g_points, float *g_quat, float *g_trans,
                    float *g_focal, float *g_scale,
                    void *workspace, size_t workspace_bytes, void *stream);
/* ---- the renderer and its loss as ONE step (f2 -> the path -> f1) ------------
 * Reference: ModelPointCloud.forward models/model_pc_to.py:302-331 (tf_repeat_0
 * of clouds and scales over views x candidates, pc_point_dropout,
 * pointcloud_project_fast) followed by add_proj_loss :339-385 and
 * proj_loss_pose_candidates :410-440, and the autograd of all of it.
 *
 * P = BV * num_candidates projections, BV = clouds * views, replicas = views *
 * num_candidates (candidates of a view adjacent, views of a cloud adjacent:
 * tf_repeat_0 order).  points / sel / pose inputs as in
 * dpc_project_replicated_fwd; gt [BV,G,G], weights [BV] (NULL ok), as in
 * dpc_candidate_loss_fwd; p->outputs must be 0.
 * Forward outputs: mask [P,V,V] (the candidate masks), all_loss [BV*C],
 * min_idx [BV], view_loss [BV], loss [1] = weight_scale / BV * sum(view_loss),
 * winners [BV] int32 (projection index of every view's winning candidate) and
 * kcoef [BV] (factor of its mask gradient) -- both state for the backward, as
 * are grid_b / clamp_bits / cells. */
DPC_API int dpc_render_loss_fwd(const dpc_params *p, int replicas, int N_src,
                    const int32_t *sel /*NULL ok*/,
                    const float *points, const float *quat,
                    const float *trans /*NULL ok*/, const float *focal /*NULL ok*/,
                    const float *scale /*NULL ok*/,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz, int scatter_mode,
                    int num_candidates, int G, const float *gt, const float *weights /*NULL ok*/,
                    float weight_scale,
                    float *grid_b, uint32_t *clamp_bits, void *cells /*NULL ok*/, float *mask,
                    float *all_loss, int64_t *min_idx, float *view_loss, float *loss,
                    int32_t *winners, float *kcoef,
                    void *workspace, size_t workspace_bytes, void *stream);
/* Backward of the step: g_points [P/replicas, N_src, 3] (summed over views and
 * candidates), g_quat [P,4], g_trans [P,3], g_focal [P], g_scale [P] (each
 * NULL when its input is) times the device scalar `upstream` (NULL = 1).
 * The one-hot candidate mask (model_pc_to.py:425-430) makes the gradient of
 * every losing candidate exactly zero: when the forward saved the fast ray
 * state (cells given, cubic grid, log-sum DRC; either scatter mode) the backward
 * chain runs over the BV winners only and dL/dmask is built inside the ray
 * kernel; otherwise it is written to g_mask_scratch [P,V,V] and the general
 * backward runs.  scatter_mode: the forward's.
 * Scratch, in projection SLOTS = dpc_render_loss_slots() (BV or P):
 * g_grid [slots,Vz,V,V], g_points_rep [slots,N,3], inv_scratch [slots,N_src]
 * int32 (needed when sel != NULL). */
DPC_API int dpc_render_loss_bwd(const dpc_params *p, int replicas, int N_src,
                    const int32_t *sel /*NULL ok*/,
                    const float *points, const float *quat,
                    const float *trans, const float *focal, const float *scale,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz, int scatter_mode,
                    int num_candidates, int G, const float *gt, const float *weights /*NULL ok*/,
                    float weight_scale,
                    const float *grid_b, const uint32_t *clamp_bits, const void *cells /*NULL ok*/,
                    const float *mask, const int64_t *min_idx, const int32_t *winners,
                    const float *kcoef, const float *upstream /*NULL ok*/,
                    float *g_grid, float *g_points_rep, int32_t *inv_scratch /*NULL ok*/,
                    float *g_mask_scratch /*NULL ok when slots == BV*/,
                    float *g_points, float *g_quat, float *g_trans,
                    float *g_focal, float *g_scale,
                    void *workspace, size_t workspace_bytes, void *stream);
/* Projection slots the backward's scratch must hold: BV when the winner-only
 * chain applies to this geometry / mode, else P (host-side query). */
DPC_API int dpc_render_loss_slots(const dpc_params *p, int num_candidates, int have_cells,
                          int scatter_mode);

/* The sampler of pc_point_dropout (point_cloud_to.py:275-283) on the device:
 * sel [P,M] = a uniformly random M-subset of [0, N_src) per projection, in
 * ascending order, a pure function of (seed, projection index).  Same
 * distribution as np.random.choice(N_src, M, replace=False) up to the order
 * inside the subset, which the projection does not depend on. */
DPC_API int dpc_point_dropout_indices(int P, int N_src, int M, uint64_t seed, int32_t *sel,
                              void *stream);
/* The gather of pc_point_dropout (select_3d, point_cloud_to.py:266-267) as a
 * stand-alone op: out [P,M,C] = points [P/replicas, N_src, C] at sel [P,M]
 * (C <= 4: xyz or rgb). */
DPC_API int dpc_select_points(int P, int replicas, int N_src, int M, int C, const float *points,
                      const int32_t *sel, float *out, void *stream);
/* Its adjoint fused with the adjoint of tf_repeat_0: g_cloud [P/replicas,
 * N_src, C] = sum over the replicas of g_rep [P,M,C] routed through sel
 * (sel == NULL: M == N_src, identity routing).  inv_scratch as above. */
DPC_API int dpc_replica_reduce(int P, int replicas, int N_src, int M, int C, const float *g_rep,
                       const int32_t *sel /*NULL ok*/, int32_t *inv_scratch /*NULL ok*/,
                       float *g_cloud, void *stream);

/* ---- a14 / f3 (next row): the point-feature (RGB) branch --------------------
 * Specification: the TF original util/point_cloud.py:99-129 (feature scatter),
 * :148-154 convolve_rgb, :244-262 (clip, division by the blurred occupancy, Y
 * flip) and project_volume_rgb_integral (util/drc.py:132-142 in the torch
 * port).  The torch port of this branch does not run (point_cloud_to.py:64,
 * drc.py:137), so parity is pinned to oracle/rgb.py's restatement of the TF file
 * only.  Feature grids are channel-planar: fgrid [P,C,Vz,V,V], C <= 4.
 *
 * fgrid[b,c,cell+corner] += w(corner) * feat[b,n,c] over the valid points of
 * tr_pc [P,N,3] (the same trilinear weights as dpc_scatter_fwd); feat [P,N,C]. */
DPC_API int dpc_feat_scatter_fwd(const dpc_params *p, int C, const float *tr_pc, const float *feat,
                         float *fgrid, void *stream);
/* Adjoint: g_feat [P,N,C] and g_tr_pc [P,N,3] (NULL ok: pc_rgb_stop_points_gradient).
 * raw (NULL ok): the forward's fgrid; when given, g_fgrid counts only where
 * 0 <= raw <= 1 -- the gate of the clip that precedes the blur (point_cloud.py:247). */
DPC_API int dpc_feat_scatter_bwd(const dpc_params *p, int C, const float *tr_pc, const float *feat,
                         const float *g_fgrid, const float *raw /*NULL ok*/, float *g_feat,
                         float *g_tr_pc /*NULL ok*/, void *stream);
/* dpc_blur3d with clamp(src, 0, 1) taken on the way in (point_cloud.py:246-248).
 * For feature grids pass params with P = P * C. */
DPC_API int dpc_blur3d_clamped(const dpc_params *p, const float *src, float *dst,
                       const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                       const float *taps_z_host, int kz, void *stream);
/* Colour integral: proj_rgb [P,V,V,C] = sum_{k<Vz} probs[k] f_c[k] + probs[Vz]
 * (white background), f = fgrid, / (div + eps) when div [P,Vz,V,V] is given
 * (pc_rgb_divide_by_occupancies), clipped to [0,1] when clip_after
 * (pc_rgb_clip_after_conv).  probs [Vz+1,P,V,V] is in output row order (as
 * dpc_project_fwd writes it); the grids are read with the Y flip of p->flip_y.
 * voxels_rgb (NULL ok): [P,Vz,V,V,C], the reference's "voxels_rgb" output. */
DPC_API int dpc_colour_fwd(const dpc_params *p, int C, const float *probs, const float *fgrid,
                   const float *div /*NULL ok*/, float eps, int clip_after, float *proj_rgb,
                   float *voxels_rgb /*NULL ok*/, void *stream);
/* Adjoint: g_probs [Vz+1,P,V,V] and g_fgrid [P,C,Vz,V,V] (both fully written). */
DPC_API int dpc_colour_bwd(const dpc_params *p, int C, const float *probs, const float *fgrid,
                   const float *div /*NULL ok*/, float eps, int clip_after,
                   const float *g_proj_rgb, float *g_probs, float *g_fgrid, void *stream);

/* ---- f4 (next row): util/point_cloud_distance.py:25-40 point_cloud_distance
 * (the kernel of run/eval_chamfer_to.py:24-44 compute_distance) -------------
 * For every source point src [N,3] the closest target point of tgt [M,3]:
 * proj [N,3] = that point, min_dist [N] = sqrt of the squared distance,
 * idx [N] = its index (first minimum, like torch.argmin).  workspace: 8 N bytes. */
DPC_API int dpc_point_cloud_distance(int N, int M, const float *src, const float *tgt, float *proj,
                             float *min_dist, int64_t *idx, void *workspace,
                             size_t workspace_bytes, void *stream);

/* ---- measurement aid (bench.py): runs dpc_project_fwd + dpc_project_bwd
 * `iters` times with a CUDA event recorded on `stream` after every stage and
 * returns the average duration of each stage in milliseconds.  Synchronises
 * the stream once per iteration -- for profiling only.  Stage order:
 * 0 memset(grid) (nothing on the plane-local path) | 1 pose + scatter (pose +
 * cell records on the plane-local path) | 2 [plane scatter +] blur XY fwd |
 * 3 blur Z + DRC fwd | 4 DRC bwd + blur Z adjoint | 5 blur XY adjoint [+ plane
 * gather] | 6 [gather +] pose adjoint + fused quaternion/translation/focal/scale
 * reduction. */
#define DPC_PROFILE_STAGES 7
DPC_API int dpc_project_profile(const dpc_params *p, const float *points, const float *quat,
                    const float *trans, const float *focal, const float *scale,
                    const float *taps_x_host, int kx, const float *taps_y_host, int ky,
                    const float *taps_z_host, int kz, int scatter_mode,
                    float *tr_pc, float *grid_b, uint32_t *clamp_bits, void *cells /*NULL ok*/,
                    float *mask, float *depth, const float *g_mask, const float *g_depth,
                    float *g_grid, float *g_points, float *g_quat, float *g_trans,
                    float *g_focal, float *g_scale,
                    void *workspace, size_t workspace_bytes, void *stream,
                    int iters, float *stage_ms_host);

/* The library allocates nothing and keeps no state -- with one exception: per host thread and
 * device, the two internal side streams (and three events) of the half-batch split (batches of
 * >= 64 projections), created on first use.  dpc_release() destroys the calling thread's set for
 * the current device; call it after synchronising the work issued through the library (it is
 * re-created on the next call that needs it).  Optional: the set is a few handles per thread. */
DPC_API int dpc_release(void);

/* Host-side query: the tap radius the blur kernels run with for these n (odd) host taps -- the
 * smallest radius outside of which the taps' total magnitude is <= 1e-7 (DPC_TAP_EPS in the
 * environment overrides; 0 keeps every non-zero tap).  The reference's Gaussian always has
 * pc_gauss_kernel_size taps (gauss_kernel.py:5-11) while sigma falls from 3.0 to 0.2 over training
 * (model_pc_to.py:59-63); kernels are instantiated for radii 2 / 5 / 7 / 10 (Z pass also 0). */
DPC_API int dpc_tap_radius(const float *taps_host, int n);
/* Process-wide override of that bound (0: every non-zero tap runs; negative: back to the
 * default / DPC_TAP_EPS).  Set it between calls, not while another thread is launching. */
DPC_API int dpc_set_tap_truncation(double eps);
/* Process-wide switch of the programmatic dependent launches of the kernel chain (1: on, the
 * default; 0: plain stream-ordered launches -- same results bit for bit; negative: back to the
 * default / DPC_PDL in the environment).  A/B and test helper; not for use inside a capture. */
DPC_API int dpc_set_programmatic_launch(int on);

/* Measurement helper (bench.py roofline.fp32_frac): one launch of an FFMA2-only kernel -- the
 * packed inner product of the blur kernels without its loads -- of blocks x 256 threads x iters
 * x 21 taps x 16 scalar FMAs; *fma_count_host (NULL ok) receives that count.  The caller times it
 * with CUDA events.  out: blocks * 256 floats of scratch. */
DPC_API int dpc_fma_rate_probe(int blocks, int iters, float *out, double *fma_count_host,
                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DPC_B200_H */
