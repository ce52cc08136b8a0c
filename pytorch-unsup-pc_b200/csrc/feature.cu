// Point-feature (RGB) branch of the projection.
//
// Specification: the TF original util/point_cloud.py:99-129 (feature scatter with the
// occupancy's trilinear weights), :148-154 convolve_rgb (per-channel separable blur),
// :244-262 (clip before or after the blur, optional division by the blurred raw
// occupancy, Y flip) and util/drc.py project_volume_rgb_integral (colour integral
// along the ray with a white background; torch port: drc.py:132-142).  The torch
// port of this branch does not run (point_cloud_to.py:64 AttributeError,
// drc.py:137 torch.float63), so parity is UNPINNED: the oracle (oracle/rgb.py)
// restates the TF file.
//
// Feature grids are channel-planar, F[P][C][Vz][V][V], so the per-channel blur is
// the occupancy blur run on P*C grids.
#include "common.cuh"
#include "pose.cuh"

namespace dpc {

constexpr int kFeatThreads = 256;
constexpr int kMaxFeat = 4;

// F[b][c][cell + corner] += w(corner) * feat[b][n][c]   (valid points only)
__global__ void __launch_bounds__(kFeatThreads)
feat_scatter_kernel(const float *__restrict__ tr_pc, const float *__restrict__ feat, int N, int C,
                    int Vz, int V, float *__restrict__ grid) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * kFeatThreads + threadIdx.x;
  if (n >= N) return;
  const size_t pi = ((size_t)b * N + n) * 3;
  const Cell c = make_cell((double)tr_pc[pi], (double)tr_pc[pi + 1], (double)tr_pc[pi + 2], Vz, V);
  if (!c.valid) return;
  float f[kMaxFeat];
  for (int k = 0; k < C; ++k) f[k] = __ldg(feat + ((size_t)b * N + n) * C + k);
  const size_t G = (size_t)Vz * V * V;
  float *g = grid + (size_t)b * C * G;
  const double wz[2] = {1.0 - c.rz, c.rz}, wy[2] = {1.0 - c.ry, c.ry}, wx[2] = {1.0 - c.rx, c.rx};
#pragma unroll
  for (int dz = 0; dz < 2; ++dz)
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int z = c.iz + dz, y = c.iy + dy, x = c.ix + dx;
        if (z >= Vz || y >= V || x >= V) continue;     // zero-weight corners of a +0.5 coordinate
        const float w = (float)(wz[dz] * wy[dy] * wx[dx]);
        const size_t o = ((size_t)z * V + y) * V + x;
        for (int k = 0; k < C; ++k) atomicAdd(g + k * G + o, w * f[k]);
      }
}

// Adjoint: g_feat[b][n][c] = sum_corners w G_c[corner];
// g_tr_pc[b][n][a] = (V_a - 1) sum_c feat_c sum_corners (dw/dr_a) G_c[corner]
// GATE: G_c[corner] counts only where 0 <= raw_c[corner] <= 1 (the clip that precedes the blur).
template <bool GATE>
__global__ void __launch_bounds__(kFeatThreads)
feat_gather_bwd_kernel(const float *__restrict__ tr_pc, const float *__restrict__ feat,
                       const float *__restrict__ g_grid, const float *__restrict__ raw, int N, int C,
                       int Vz, int V, float *__restrict__ g_feat, float *__restrict__ g_trpc) {
  const int b = blockIdx.y;
  const int n = blockIdx.x * kFeatThreads + threadIdx.x;
  if (n >= N) return;
  const size_t pi = ((size_t)b * N + n) * 3;
  const Cell c = make_cell((double)tr_pc[pi], (double)tr_pc[pi + 1], (double)tr_pc[pi + 2], Vz, V);
  const size_t G = (size_t)Vz * V * V;
  double gz = 0, gy = 0, gx = 0;
  const double wz[2] = {1.0 - c.rz, c.rz}, wy[2] = {1.0 - c.ry, c.ry}, wx[2] = {1.0 - c.rx, c.rx};
  for (int k = 0; k < C; ++k) {
    double gf = 0, cz = 0, cy = 0, cx = 0;
    if (c.valid) {
      const float *g = g_grid + ((size_t)b * C + k) * G;
      const float *r = GATE ? raw + ((size_t)b * C + k) * G : nullptr;
#pragma unroll
      for (int dz = 0; dz < 2; ++dz)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int z = c.iz + dz, y = c.iy + dy, x = c.ix + dx;
            if (z >= Vz || y >= V || x >= V) continue;
            const size_t o = ((size_t)z * V + y) * V + x;
            double v = (double)__ldg(g + o);
            if (GATE) {
              const float rv = __ldg(r + o);
              if (!(rv >= 0.f && rv <= 1.f)) v = 0.0;
            }
            gf += wz[dz] * wy[dy] * wx[dx] * v;
            cz += (dz ? 1.0 : -1.0) * wy[dy] * wx[dx] * v;
            cy += (dy ? 1.0 : -1.0) * wz[dz] * wx[dx] * v;
            cx += (dx ? 1.0 : -1.0) * wz[dz] * wy[dy] * v;
          }
    }
    const double f = (double)__ldg(feat + ((size_t)b * N + n) * C + k);
    g_feat[((size_t)b * N + n) * C + k] = (float)gf;
    gz += f * cz;
    gy += f * cy;
    gx += f * cx;
  }
  if (g_trpc) {
    g_trpc[pi] = (float)(gz * (double)(Vz - 1));
    g_trpc[pi + 1] = (float)(gy * (double)(V - 1));
    g_trpc[pi + 2] = (float)(gx * (double)(V - 1));
  }
}

// The feature value a ray sees in voxel (k, row, x) of channel c: the blurred grid, optionally
// divided by the blurred raw occupancy + eps, optionally clipped to [0, 1] afterwards.
__device__ __forceinline__ float feat_value(float fb, float div, float eps, bool has_div,
                                            bool clip_after, float &dfdF) {
  float f = fb;
  dfdF = 1.f;
  if (has_div) {
    dfdF = 1.f / (div + eps);
    f = fb * dfdF;
  }
  if (clip_after) {
    if (!(f >= 0.f && f <= 1.f)) dfdF = 0.f;    // torch/TF clip gradient: closed interval
    f = fminf(fmaxf(f, 0.f), 1.f);
  }
  return f;
}

// proj_rgb[b][y][x][c] = sum_{k<Vz} p_k f_c(k) + p_Vz * 1   (white background)
// probs [Vz+1][P][V][V] is in OUTPUT row order (already Y-flipped when flip_y); the feature
// grids are in grid row order, so row = flip_y ? V-1-y : y.
__global__ void __launch_bounds__(128)
colour_fwd_kernel(const float *__restrict__ probs, const float *__restrict__ fgrid,
                  const float *__restrict__ div, float eps, int clip_after, int P, int C, int Vz,
                  int V, int flip_y, float *__restrict__ proj_rgb) {
  const int VV = V * V;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= P * VV) return;
  const int b = i / VV, yx = i - b * VV, y = yx / V, x = yx - y * V;
  const int row = flip_y ? V - 1 - y : y;
  const size_t G = (size_t)Vz * VV, pstride = (size_t)P * VV;
  float acc[kMaxFeat] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < Vz; ++k) {
    const float p = __ldg(probs + k * pstride + i);
    const size_t o = (size_t)k * VV + row * V + x;
    const float dv = div ? __ldg(div + (size_t)b * G + o) : 0.f;
    for (int c = 0; c < C; ++c) {
      float d;
      acc[c] = fmaf(p, feat_value(__ldg(fgrid + ((size_t)b * C + c) * G + o), dv, eps, div != nullptr,
                                  clip_after != 0, d), acc[c]);
    }
  }
  const float pz = __ldg(probs + (size_t)Vz * pstride + i);
  for (int c = 0; c < C; ++c) proj_rgb[(size_t)i * C + c] = acc[c] + pz;
}

__global__ void __launch_bounds__(128)
colour_bwd_kernel(const float *__restrict__ probs, const float *__restrict__ fgrid,
                  const float *__restrict__ div, float eps, int clip_after, int P, int C, int Vz,
                  int V, int flip_y, const float *__restrict__ g_proj, float *__restrict__ g_probs,
                  float *__restrict__ g_fgrid) {
  const int VV = V * V;
  const int i = blockIdx.x * 128 + threadIdx.x;
  if (i >= P * VV) return;
  const int b = i / VV, yx = i - b * VV, y = yx / V, x = yx - y * V;
  const int row = flip_y ? V - 1 - y : y;
  const size_t G = (size_t)Vz * VV, pstride = (size_t)P * VV;
  float g[kMaxFeat], gsum = 0.f;
  for (int c = 0; c < C; ++c) {
    g[c] = __ldg(g_proj + (size_t)i * C + c);
    gsum += g[c];
  }
  for (int k = 0; k < Vz; ++k) {
    const float p = __ldg(probs + k * pstride + i);
    const size_t o = (size_t)k * VV + row * V + x;
    const float dv = div ? __ldg(div + (size_t)b * G + o) : 0.f;
    float gp = 0.f;
    for (int c = 0; c < C; ++c) {
      float d;
      const float f = feat_value(__ldg(fgrid + ((size_t)b * C + c) * G + o), dv, eps, div != nullptr,
                                 clip_after != 0, d);
      gp = fmaf(g[c], f, gp);
      g_fgrid[((size_t)b * C + c) * G + o] = p * g[c] * d;
    }
    g_probs[k * pstride + i] = gp;
  }
  g_probs[(size_t)Vz * pstride + i] = gsum;
}

// voxels_rgb[b][k][y][x][c] (channel-last, Y-flipped like the reference's output)
__global__ void __launch_bounds__(256)
feat_voxels_out_kernel(const float *__restrict__ fgrid, const float *__restrict__ div, float eps,
                       int clip_after, int P, int C, int Vz, int V, int flip_y,
                       float *__restrict__ out) {
  const size_t G = (size_t)Vz * V * V;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= (size_t)P * G) return;
  const int b = (int)(i / G);
  const size_t r = i - (size_t)b * G;
  const int k = (int)(r / (V * V)), yx = (int)(r - (size_t)k * V * V), y = yx / V, x = yx - y * V;
  const int row = flip_y ? V - 1 - y : y;
  const size_t o = (size_t)k * V * V + row * V + x;
  const float dv = div ? __ldg(div + (size_t)b * G + o) : 0.f;
  for (int c = 0; c < C; ++c) {
    float d;
    out[i * C + c] = feat_value(__ldg(fgrid + ((size_t)b * C + c) * G + o), dv, eps, div != nullptr,
                                clip_after != 0, d);
  }
}

// ---- launchers -----------------------------------------------------------------
int feat_max_channels() { return kMaxFeat; }

int launch_feat_scatter(const float *tr_pc, const float *feat, int P, int N, int C, int Vz, int V,
                        float *grid, cudaStream_t s) {
  if (cudaMemsetAsync(grid, 0, (size_t)P * C * Vz * V * V * sizeof(float), s) != cudaSuccess)
    return check_launch("memset(feature grid)");
  feat_scatter_kernel<<<dim3((N + kFeatThreads - 1) / kFeatThreads, P), kFeatThreads, 0, s>>>(
      tr_pc, feat, N, C, Vz, V, grid);
  return check_launch("feat_scatter");
}

int launch_feat_gather_bwd(const float *tr_pc, const float *feat, const float *g_grid,
                           const float *raw, int P, int N, int C, int Vz, int V, float *g_feat,
                           float *g_trpc, cudaStream_t s) {
  dim3 g((N + kFeatThreads - 1) / kFeatThreads, P);
  if (raw)
    feat_gather_bwd_kernel<true><<<g, kFeatThreads, 0, s>>>(tr_pc, feat, g_grid, raw, N, C, Vz, V,
                                                            g_feat, g_trpc);
  else
    feat_gather_bwd_kernel<false><<<g, kFeatThreads, 0, s>>>(tr_pc, feat, g_grid, nullptr, N, C, Vz,
                                                             V, g_feat, g_trpc);
  return check_launch("feat_gather_bwd");
}

int launch_colour_fwd(const float *probs, const float *fgrid, const float *div, float eps,
                      int clip_after, int P, int C, int Vz, int V, int flip_y, float *proj_rgb,
                      float *voxels_rgb, cudaStream_t s) {
  colour_fwd_kernel<<<(P * V * V + 127) / 128, 128, 0, s>>>(probs, fgrid, div, eps, clip_after, P, C,
                                                            Vz, V, flip_y, proj_rgb);
  if (int e = check_launch("colour_fwd")) return e;
  if (voxels_rgb) {
    const size_t n = (size_t)P * Vz * V * V;
    feat_voxels_out_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(fgrid, div, eps, clip_after, P,
                                                                       C, Vz, V, flip_y, voxels_rgb);
    return check_launch("feat_voxels_out");
  }
  return DPC_OK;
}

int launch_colour_bwd(const float *probs, const float *fgrid, const float *div, float eps,
                      int clip_after, int P, int C, int Vz, int V, int flip_y, const float *g_proj,
                      float *g_probs, float *g_fgrid, cudaStream_t s) {
  colour_bwd_kernel<<<(P * V * V + 127) / 128, 128, 0, s>>>(probs, fgrid, div, eps, clip_after, P, C,
                                                            Vz, V, flip_y, g_proj, g_probs, g_fgrid);
  return check_launch("colour_bwd");
}

}  // namespace dpc
