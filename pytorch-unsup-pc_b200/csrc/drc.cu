// Z pass of the Gaussian blur fused with the DRC ray march (forward), and the
// DRC reverse scan fused with the Z-pass adjoint (backward).
//
// Reference: point_cloud_to.py:97 (third conv3d, kernel [1,1,K,1,1]), :218-222
// (occupancy scaling + clamp), drc.py:48-106 (ray-termination probabilities,
// log-sum form with clip_val 'unity' padding), :114-129 (silhouette), :145-160
// (expected depth) and the two Y flips point_cloud_to.py:239, 242.
//
// One thread owns one ray (b, y, x); a warp owns 32 consecutive x, so every
// global access is a full 128-byte line.  The Z blur is a register ring of
// 2R+1 values indexed at compile time (the z loop is unrolled by the ring
// length), so the blurred occupancy never exists in memory: each value is
// consumed by the ray march as soon as it is produced.
//
//   vox_k = clamp(s * blurZ(grid)_k, 0, 1)        v_k = clamp(vox_k, c, 1-c)
//   p_k = e_k v_k T_k,  T_{k+1} = T_k (1 - v_k),  p_Z = e_Z T_Z,  e_0 = e_Z = exp(c)
//   mask = sum_{k<Z} p_k      depth = sum_k psi_k p_k
//
// The reference evaluates the same product as exp(cumsum(log(.))) in fp64; the
// product form needs no transcendentals and agrees to fp32 rounding.
//
// Backward (rewritten without cancellation): with a_k the
// upstream weight of p_k and D_j = dL/dT_j,
//   D_Z = a_Z e_Z,   D_k = a_k e_k v_k + (1 - v_k) D_{k+1},
//   dL/dv_k = T_k (a_k e_k - D_{k+1}).
// Sweep 1 (forward in z) recomputes blurZ and T_k into shared memory
// ([k][thread] layout, conflict-free); sweep 2 (reverse in z) runs the D
// recursion, applies the clip/clamp gates and the scale, and streams the
// result through the same register ring to apply the Z-blur adjoint on the
// way out.  dL/dscale is reduced per block in fixed order.
#include "common.cuh"

namespace dpc {

constexpr int kRayThreads = 128;

int drc_scale_partial_blocks(int V) { return V * V / kRayThreads; }

struct RayConst {
  int P, Vz, V, VV;
  float inv_z, depth0, max_depth, clip, one_minus_clip, exp_clip;
  int logsum, flip_y, has_scale;
};

static RayConst make_ray_const(const DrcArgs &a) {
  RayConst c;
  c.P = a.P; c.Vz = a.Vz; c.V = a.V; c.VV = a.V * a.V;
  c.inv_z = 1.0f / (float)a.Vz;
  c.depth0 = a.cam_dist - 0.5f;
  c.max_depth = a.max_depth;
  c.clip = a.clip;
  c.one_minus_clip = 1.0f - a.clip;
  c.exp_clip = a.logsum ? (float)exp((double)a.clip) : 1.0f;
  c.logsum = a.logsum;
  c.flip_y = a.flip_y;
  c.has_scale = a.scale != nullptr;
  return c;
}

template <int R>
__device__ __forceinline__ float ring_dot(const float (&ring)[2 * R + 1], const Taps<R> &taps,
                                          int first /*compile-time*/) {
  constexpr int W = 2 * R + 1;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int t = 0; t < W; ++t) {
    const float v = ring[(first + t) % W];
    if (t % 3 == 0) s0 = fmaf(taps.k[t], v, s0);
    else if (t % 3 == 1) s1 = fmaf(taps.k[t], v, s1);
    else s2 = fmaf(taps.k[t], v, s2);
  }
  return (s0 + s1) + s2;
}

// Streams blurZ(col)_z for z = 0..Vz-1 to `sink(z, value)`.
template <int R, typename Sink>
__device__ __forceinline__ void stream_blur_z(const float *__restrict__ col, int Vz, int VV,
                                              const Taps<R> &taps, Sink &&sink) {
  constexpr int W = 2 * R + 1;
  if (R == 0) {
#pragma unroll 8
    for (int z = 0; z < Vz; ++z) sink(z, taps.k[0] * __ldg(col + (size_t)z * VV));
    return;
  }
  float ring[W];
#pragma unroll
  for (int i = 0; i < W; ++i) ring[i] = 0.f;
#pragma unroll
  for (int i = 0; i < R; ++i) ring[i] = (i < Vz) ? __ldg(col + (size_t)i * VV) : 0.f;
#pragma unroll 1
  for (int z0 = 0; z0 < Vz; z0 += W) {
    float nxt[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int zin = z0 + j + R;
      nxt[j] = (zin < Vz) ? __ldg(col + (size_t)zin * VV) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int z = z0 + j;
      if (z < Vz) {
        ring[(j + R) % W] = nxt[j];
        sink(z, ring_dot<R>(ring, taps, (j + R + 1) % W));
      }
    }
  }
}

__device__ __forceinline__ void ray_index(const RayConst &c, int &b, int &yx, int &out_idx) {
  const int ray = blockIdx.x * kRayThreads + threadIdx.x;
  b = ray / c.VV;
  yx = ray - b * c.VV;
  const int y = yx / c.V, x = yx - y * c.V;
  const int yo = c.flip_y ? (c.V - 1 - y) : y;
  out_idx = b * c.VV + yo * c.V + x;
}

template <int R>
__global__ void __launch_bounds__(kRayThreads)
blurz_drc_fwd_kernel(const float *__restrict__ grid, const float *__restrict__ scale, RayConst c,
                     const Taps<R> kz, float *__restrict__ mask, float *__restrict__ depth,
                     float *__restrict__ voxels, float *__restrict__ probs) {
  int b, yx, oi;
  ray_index(c, b, yx, oi);
  const float *col = grid + (size_t)b * c.Vz * c.VV + yx;
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  float T = 1.f, m = 0.f, d = 0.f;
  const size_t pstride = (size_t)c.P * c.VV;
  stream_blur_z<R>(col, c.Vz, c.VV, kz, [&](int z, float bz) {
    float vox = bz;
    if (c.has_scale) vox = fminf(fmaxf(s * bz, 0.f), 1.f);
    if (voxels) voxels[(size_t)b * c.Vz * c.VV + (size_t)z * c.VV + yx] = vox;
    const float v = c.logsum ? fminf(fmaxf(vox, c.clip), c.one_minus_clip) : vox;
    float p = v * T;
    if (z == 0) p *= c.exp_clip;
    if (probs) probs[(size_t)z * pstride + oi] = p;
    m += p;
    d = fmaf((float)z * c.inv_z + c.depth0, p, d);
    T *= (1.f - v);
  });
  const float pz = c.exp_clip * T;
  if (probs) probs[(size_t)c.Vz * pstride + oi] = pz;
  mask[oi] = m;
  if (depth) depth[oi] = fmaf(c.max_depth, pz, d);
}

template <int R>
__global__ void __launch_bounds__(kRayThreads)
drc_blurz_bwd_kernel(const float *__restrict__ grid, const float *__restrict__ scale, RayConst c,
                     const Taps<R> kz, const float *__restrict__ g_mask,
                     const float *__restrict__ g_depth, const float *__restrict__ g_probs,
                     const float *__restrict__ g_voxels, float *__restrict__ g_grid,
                     float *__restrict__ scale_partials) {
  constexpr int W = 2 * R + 1;
  extern __shared__ float sm[];
  float *sB = sm + threadIdx.x;                        // [Vz][threads]  blurZ value
  float *sT = sm + c.Vz * kRayThreads + threadIdx.x;   // [Vz][threads]  transmittance T_k
  int b, yx, oi;
  ray_index(c, b, yx, oi);
  const float *col = grid + (size_t)b * c.Vz * c.VV + yx;
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  // ---- sweep 1: recompute blurZ and T_k ----
  float T = 1.f;
  stream_blur_z<R>(col, c.Vz, c.VV, kz, [&](int z, float bz) {
    sB[z * kRayThreads] = bz;
    sT[z * kRayThreads] = T;
    float vox = bz;
    if (c.has_scale) vox = fminf(fmaxf(s * bz, 0.f), 1.f);
    const float v = c.logsum ? fminf(fmaxf(vox, c.clip), c.one_minus_clip) : vox;
    T *= (1.f - v);
  });
  // ---- sweep 2: reverse scan + Z-blur adjoint ----
  const float gm = g_mask ? __ldg(g_mask + oi) : 0.f;
  const float gd = g_depth ? __ldg(g_depth + oi) : 0.f;
  const size_t pstride = (size_t)c.P * c.VV;
  float D = c.max_depth * gd;
  if (g_probs) D += __ldg(g_probs + (size_t)c.Vz * pstride + oi);
  D *= c.exp_clip;
  float ds = 0.f;
  float *gcol = g_grid + (size_t)b * c.Vz * c.VV + yx;
  const float *gvcol = g_voxels ? g_voxels + (size_t)b * c.Vz * c.VV + yx : nullptr;

  auto step = [&](int k) -> float {  // returns dL/d(grid_xy blurred in z)_k
    const float bz = sB[k * kRayThreads];
    const float Tk = sT[k * kRayThreads];
    const float sb = s * bz;
    float vox = bz;
    if (c.has_scale) vox = fminf(fmaxf(sb, 0.f), 1.f);
    const float v = c.logsum ? fminf(fmaxf(vox, c.clip), c.one_minus_clip) : vox;
    float a = fmaf((float)k * c.inv_z + c.depth0, gd, gm);
    if (g_probs) a += __ldg(g_probs + (size_t)k * pstride + oi);
    if (k == 0) a *= c.exp_clip;
    float gv = Tk * (a - D);
    D = fmaf(a, v, (1.f - v) * D);
    if (c.logsum) gv = (vox >= c.clip && vox <= c.one_minus_clip) ? gv : 0.f;
    if (gvcol) gv += __ldg(gvcol + (size_t)k * c.VV);
    if (c.has_scale) {
      gv = (sb >= 0.f && sb <= 1.f) ? gv : 0.f;
      ds = fmaf(gv, bz, ds);
      gv *= s;
    }
    return gv;
  };

  if (R == 0) {
#pragma unroll 4
    for (int k = c.Vz - 1; k >= 0; --k) gcol[(size_t)k * c.VV] = kz.k[0] * step(k);
  } else {
    float ring[W];
#pragma unroll
    for (int i = 0; i < W; ++i) ring[i] = 0.f;
    const int total = c.Vz + R;  // steps m = 0 .. Vz+R-1, input k = Vz-1-m, output z = k+R
#pragma unroll 1
    for (int m0 = 0; m0 < total; m0 += W) {
#pragma unroll
      for (int j = 0; j < W; ++j) {
        const int m = m0 + j;
        if (m < total) {
          const int k = c.Vz - 1 - m;
          ring[j] = (k >= 0) ? step(k) : 0.f;
          const int z = k + R;
          if (z < c.Vz) {
            // out[z] = sum_t kz[t] * in[z + t - R]; in[i] sits in slot (Vz-1-i) % W = (m - t) % W
            float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int t = 0; t < W; ++t) {
              // adjoint of a correlation = correlation with the reversed taps
              const float v = ring[(j - t + W) % W];
              if (t % 3 == 0) s0 = fmaf(kz.k[W - 1 - t], v, s0);
              else if (t % 3 == 1) s1 = fmaf(kz.k[W - 1 - t], v, s1);
              else s2 = fmaf(kz.k[W - 1 - t], v, s2);
            }
            gcol[(size_t)z * c.VV] = (s0 + s1) + s2;
          }
        }
      }
    }
  }
  // ---- dL/dscale: fixed-order block reduction ----
  if (scale_partials) {
    __shared__ float red[kRayThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ds += __shfl_down_sync(0xffffffffu, ds, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kRayThreads / 32; ++w) v += red[w];
      scale_partials[blockIdx.x] = v;  // blocks are projection-major
    }
  }
}

template <int R>
__global__ void __launch_bounds__(kRayThreads)
blur_z_kernel(const float *__restrict__ src, float *__restrict__ dst, int Vz, int VV,
              const Taps<R> kz) {
  const int ray = blockIdx.x * kRayThreads + threadIdx.x;
  const int b = ray / VV, yx = ray - b * VV;
  const float *col = src + (size_t)b * Vz * VV + yx;
  float *out = dst + (size_t)b * Vz * VV + yx;
  stream_blur_z<R>(col, Vz, VV, kz, [&](int z, float bz) { out[(size_t)z * VV] = bz; });
}

__global__ void __launch_bounds__(kRayThreads)
depth_from_probs_kernel(const float *__restrict__ probs, float *__restrict__ depth, int P, int Vz,
                        int VV, float inv_z, float depth0, float max_depth) {
  const int i = blockIdx.x * kRayThreads + threadIdx.x;
  if (i >= P * VV) return;
  const size_t stride = (size_t)P * VV;
  float d = 0.f;
#pragma unroll 8
  for (int k = 0; k < Vz; ++k) d = fmaf((float)k * inv_z + depth0, __ldg(probs + k * stride + i), d);
  depth[i] = fmaf(max_depth, __ldg(probs + (size_t)Vz * stride + i), d);
}

__global__ void __launch_bounds__(256)
depth_from_probs_bwd_kernel(const float *__restrict__ g_depth, float *__restrict__ g_probs, int P,
                            int Vz, int VV, float inv_z, float depth0, float max_depth) {
  const size_t stride = (size_t)P * VV;
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (i >= stride * (Vz + 1)) return;
  const int k = (int)(i / stride);
  const float psi = (k == Vz) ? max_depth : (float)k * inv_z + depth0;
  g_probs[i] = psi * __ldg(g_depth + (i - k * stride));
}

// ---- launchers ---------------------------------------------------------------
static int z_radius(const float *tz, int kz) { return kz > 0 ? effective_radius(tz, kz) : 0; }

template <int R>
static Taps<R> z_taps(const float *tz, int kz, int r) {
  if (kz <= 0) return make_taps<R>(nullptr, 0);
  const int off = kz / 2 - r;  // drop outer exact zeros
  return make_taps<R>(tz + off, 2 * r + 1);
}

static int check_ray_geometry(const DrcArgs &a) {
  if ((a.V * a.V) % kRayThreads != 0) {
    set_error("drc: V*V must be a multiple of %d", kRayThreads);
    return DPC_ERR_ARG;
  }
  return 0;
}

#define DPC_DISPATCH_R(r, ...)                                            \
  do {                                                                    \
    if (r == 0) { constexpr int R = 0; __VA_ARGS__; }                     \
    else if (r <= 2) { constexpr int R = 2; __VA_ARGS__; }                \
    else if (r <= 5) { constexpr int R = 5; __VA_ARGS__; }                \
    else if (r <= 10) { constexpr int R = 10; __VA_ARGS__; }              \
    else { set_error("z tap radius %d > 10 unsupported", r); return DPC_ERR_ARG; } \
  } while (0)

int launch_blurz_drc_fwd(const DrcArgs &a, const float *tz, int kz, float *mask, float *depth,
                         float *voxels, float *probs, cudaStream_t s) {
  if (int e = check_ray_geometry(a)) return e;
  const RayConst c = make_ray_const(a);
  const int r = z_radius(tz, kz);
  const int blocks = a.P * a.V * a.V / kRayThreads;
  DPC_DISPATCH_R(r, blurz_drc_fwd_kernel<R><<<blocks, kRayThreads, 0, s>>>(
                        a.grid, a.scale, c, z_taps<R>(tz, kz, r), mask, depth, voxels, probs));
  return check_launch("blurz_drc_fwd");
}

int launch_drc_blurz_bwd(const DrcArgs &a, const float *tz, int kz, const float *g_mask,
                         const float *g_depth, const float *g_probs, const float *g_voxels,
                         float *g_grid, float *scale_partials, cudaStream_t s) {
  if (int e = check_ray_geometry(a)) return e;
  const RayConst c = make_ray_const(a);
  const int r = z_radius(tz, kz);
  const int blocks = a.P * a.V * a.V / kRayThreads;
  const size_t smem = (size_t)2 * a.Vz * kRayThreads * sizeof(float);
  if (smem > 200 * 1024) {
    set_error("drc_bwd: vox_size_z %d too large", a.Vz);
    return DPC_ERR_ARG;
  }
  DPC_DISPATCH_R(r,
    cudaFuncSetAttribute(drc_blurz_bwd_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         200 * 1024);
    drc_blurz_bwd_kernel<R><<<blocks, kRayThreads, smem, s>>>(
        a.grid, a.scale, c, z_taps<R>(tz, kz, r), g_mask, g_depth, g_probs, g_voxels, g_grid,
        a.scale ? scale_partials : nullptr));
  return check_launch("drc_blurz_bwd");
}

int launch_blur_z(const float *src, float *dst, int P, int Vz, int V, const float *tz, int kz,
                  cudaStream_t s) {
  if ((V * V) % kRayThreads != 0) {
    set_error("blur_z: V*V must be a multiple of %d", kRayThreads);
    return DPC_ERR_ARG;
  }
  const int r = z_radius(tz, kz);
  const int blocks = P * V * V / kRayThreads;
  DPC_DISPATCH_R(r, blur_z_kernel<R><<<blocks, kRayThreads, 0, s>>>(src, dst, Vz, V * V,
                                                                    z_taps<R>(tz, kz, r)));
  return check_launch("blur_z");
}

int launch_depth_from_probs(const float *probs, float *depth, int P, int Vz, int V,
                            float cam_dist, float max_depth, cudaStream_t s) {
  const int n = P * V * V;
  depth_from_probs_kernel<<<(n + kRayThreads - 1) / kRayThreads, kRayThreads, 0, s>>>(
      probs, depth, P, Vz, V * V, 1.0f / (float)Vz, cam_dist - 0.5f, max_depth);
  return check_launch("depth_from_probs");
}

int launch_depth_from_probs_bwd(const float *g_depth, float *g_probs, int P, int Vz, int V,
                                float cam_dist, float max_depth, cudaStream_t s) {
  const size_t n = (size_t)P * V * V * (Vz + 1);
  depth_from_probs_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(
      g_depth, g_probs, P, Vz, V * V, 1.0f / (float)Vz, cam_dist - 0.5f, max_depth);
  return check_launch("depth_from_probs_bwd");
}

}  // namespace dpc
