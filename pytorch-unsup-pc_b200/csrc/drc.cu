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
// length) with the taps in uniform registers; each blurred value is consumed
// by the ray march as soon as it is produced and written back IN PLACE over
// the column it came from (an input is always loaded R steps before its slot
// is overwritten), so the buffer the backward needs costs no extra memory.
//
//   vox_k = clamp(s * B_k, 0, 1)   B = blurZ(grid_xy)     v_k = clamp(vox_k, c, 1-c)
//   p_k = e_k v_k T_k,  T_{k+1} = T_k (1 - v_k),  p_Z = e_Z T_Z,  e_0 = e_Z = exp(c)
//   mask = sum_{k<Z} p_k      depth = sum_k psi_k p_k
//
// The reference evaluates the same product as exp(cumsum(log(.))) in fp64; the
// product form needs no transcendentals and agrees to fp32 rounding.  Optional
// behaviour (no scale, product-form DRC) is folded into the clamp bounds
// (+-inf) instead of branches.
//
// Backward (rewritten without cancellation): with a_k the
// upstream weight of p_k and D_j = dL/dT_j,
//   D_Z = a_Z e_Z,   D_k = a_k e_k v_k + (1 - v_k) D_{k+1},
//   dL/dv_k = T_k (a_k e_k - D_{k+1}).
// Sweep 1 (forward in z) copies the saved B column to shared memory
// ([k][thread], conflict-free) and checkpoints T at every ring-length block;
// sweep 2 walks the blocks in reverse: it re-expands T inside the block from
// the checkpoint, runs the D recursion downwards, applies the clip/clamp
// gates and the scale, and pushes the result through the register ring to
// apply the Z-blur adjoint on the way out.  dL/dscale is reduced per block
// in fixed order.
#include <math_constants.h>

#include "common.cuh"

namespace dpc {

constexpr int kRayThreads = 128;

int drc_scale_partial_blocks(int V) { return V * V / kRayThreads; }

struct RayConst {
  int P, Vz, V, VV;
  float inv_z, depth0, max_depth, exp_clip;
  float lo_s, hi_s;   // clamp(s*B, 0, 1) bounds; +-inf when there is no scaling factor
  float lo_c, hi_c;   // DRC clip bounds;        +-inf for the product form
  int flip_y, has_scale;
};

static RayConst make_ray_const(const DrcArgs &a) {
  RayConst c;
  c.P = a.P; c.Vz = a.Vz; c.V = a.V; c.VV = a.V * a.V;
  c.inv_z = 1.0f / (float)a.Vz;
  c.depth0 = a.cam_dist - 0.5f;
  c.max_depth = a.max_depth;
  c.exp_clip = a.logsum ? (float)exp((double)a.clip) : 1.0f;
  c.has_scale = a.scale != nullptr;
  c.lo_s = c.has_scale ? 0.f : -INFINITY;
  c.hi_s = c.has_scale ? 1.f : INFINITY;
  c.lo_c = a.logsum ? a.clip : -INFINITY;
  c.hi_c = a.logsum ? 1.0f - a.clip : INFINITY;
  c.flip_y = a.flip_y;
  return c;
}

template <int R>
__device__ __forceinline__ float ring_dot(const float (&ring)[2 * R + 1], const Taps<R> &taps,
                                          int first /*compile-time*/, bool reversed) {
  constexpr int W = 2 * R + 1;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int t = 0; t < W; ++t) {
    const float v = ring[(first + t) % W];
    const float k = reversed ? taps.k[W - 1 - t] : taps.k[t];
    if (t % 3 == 0) s0 = fmaf(k, v, s0);
    else if (t % 3 == 1) s1 = fmaf(k, v, s1);
    else s2 = fmaf(k, v, s2);
  }
  return (s0 + s1) + s2;
}

// Streams blurZ(col)_z for z = 0..Vz-1 to `sink(j, z, value)`; j is the
// compile-time position inside the current block of W steps.
template <int R, typename Sink>
__device__ __forceinline__ void stream_blur_z(const float *col, int Vz, int VV,
                                              const Taps<R> &taps, Sink &&sink) {
  constexpr int W = 2 * R + 1;
  if (R == 0) {
#pragma unroll 8
    for (int z = 0; z < Vz; ++z) {
      sink(1, z, taps.k[0] * *col);
      col += VV;
    }
    return;
  }
  float ring[W];
#pragma unroll
  for (int i = 0; i < W; ++i) ring[i] = 0.f;
  const float *ld = col;
#pragma unroll
  for (int i = 0; i < R; ++i) {
    ring[i] = (i < Vz) ? *ld : 0.f;
    ld += VV;
  }
#pragma unroll 1
  for (int z0 = 0; z0 < Vz; z0 += W) {
    float nxt[W];
#pragma unroll
    for (int j = 0; j < W; ++j) {
      nxt[j] = (z0 + j + R < Vz) ? *ld : 0.f;
      ld += VV;
    }
#pragma unroll
    for (int j = 0; j < W; ++j) {
      const int z = z0 + j;
      if (z < Vz) {
        ring[(j + R) % W] = nxt[j];
        sink(j, z, ring_dot<R>(ring, taps, (j + R + 1) % W, false));
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

// EXTRA: the optional voxels / probs outputs exist.  `grid` and `bsave` may
// alias (in-place save), so neither is __restrict__.
template <int R, bool EXTRA>
__global__ void __launch_bounds__(kRayThreads)
blurz_drc_fwd_kernel(const float *grid, const float *__restrict__ scale, RayConst c,
                     const Taps<R> kz, float *bsave, float *__restrict__ mask,
                     float *__restrict__ depth, float *__restrict__ voxels,
                     float *__restrict__ probs) {
  int b, yx, oi;
  ray_index(c, b, yx, oi);
  const size_t col0 = (size_t)b * c.Vz * c.VV + yx;
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  float T = 1.f, m = 0.f, d = 0.f, kf = 0.f;
  float *bs = bsave ? bsave + col0 : nullptr;
  float *vx = (EXTRA && voxels) ? voxels + col0 : nullptr;
  float *pr = (EXTRA && probs) ? probs + oi : nullptr;
  const size_t pstride = (size_t)c.P * c.VV;
  stream_blur_z<R>(grid + col0, c.Vz, c.VV, kz, [&](int j, int z, float bz) {
    if (bs) { *bs = bz; bs += c.VV; }
    const float vox = fminf(fmaxf(s * bz, c.lo_s), c.hi_s);
    const float v = fminf(fmaxf(vox, c.lo_c), c.hi_c);
    float p = v * T;
    if ((R == 0 || j == 0) && z == 0) p *= c.exp_clip;   // only block position 0 can be z == 0
    if (EXTRA) {
      if (vx) { *vx = vox; vx += c.VV; }
      if (pr) { *pr = p; pr += pstride; }
    }
    m += p;
    d = fmaf(fmaf(kf, c.inv_z, c.depth0), p, d);
    kf += 1.f;
    T *= (1.f - v);
  });
  const float pz = c.exp_clip * T;
  if (EXTRA && pr) *pr = pz;
  mask[oi] = m;
  if (depth) depth[oi] = fmaf(c.max_depth, pz, d);
}

template <int R, bool EXTRA>
__global__ void __launch_bounds__(kRayThreads)
drc_blurz_bwd_kernel(const float *__restrict__ bgrid, const float *__restrict__ scale, RayConst c,
                     const Taps<R> kz, const float *__restrict__ g_mask,
                     const float *__restrict__ g_depth, const float *__restrict__ g_probs,
                     const float *__restrict__ g_voxels, float *__restrict__ g_grid,
                     float *__restrict__ scale_partials) {
  constexpr int W = 2 * R + 1;
  constexpr int BL = (R == 0) ? 16 : W;   // block length, a multiple of the ring length
  extern __shared__ float sm[];
  float *sB = sm + threadIdx.x;                        // [Vz][threads]     saved blurZ value
  float *sC = sm + c.Vz * kRayThreads + threadIdx.x;   // [nblk][threads]   T at block starts
  int b, yx, oi;
  ray_index(c, b, yx, oi);
  const size_t col0 = (size_t)b * c.Vz * c.VV + yx;
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  const int nblk = (c.Vz + BL - 1) / BL;

  auto occupancy = [&](float bz, float &sb, float &vox) -> float {
    sb = s * bz;
    vox = fminf(fmaxf(sb, c.lo_s), c.hi_s);
    return fminf(fmaxf(vox, c.lo_c), c.hi_c);
  };

  // ---- sweep 1: stage the column, checkpoint the transmittance ----
  {
    const float *ld = bgrid + col0;
    float T = 1.f;
#pragma unroll 1
    for (int bi = 0; bi < nblk; ++bi) {
      sC[bi * kRayThreads] = T;
      float vals[BL];
#pragma unroll
      for (int j = 0; j < BL; ++j) {
        vals[j] = (bi * BL + j < c.Vz) ? __ldg(ld) : 0.f;
        ld += c.VV;
      }
#pragma unroll
      for (int j = 0; j < BL; ++j) {
        const int z = bi * BL + j;
        if (z < c.Vz) {
          sB[z * kRayThreads] = vals[j];
          float sb, vox;
          T *= (1.f - occupancy(vals[j], sb, vox));
        }
      }
    }
  }
  // ---- sweep 2: reverse scan + Z-blur adjoint ----
  const float gm = g_mask ? __ldg(g_mask + oi) : 0.f;
  const float gd = g_depth ? __ldg(g_depth + oi) : 0.f;
  const size_t pstride = (size_t)c.P * c.VV;
  float D = c.max_depth * gd;
  if (EXTRA && g_probs) D += __ldg(g_probs + (size_t)c.Vz * pstride + oi);
  D *= c.exp_clip;
  float ds = 0.f;
  float ring[W];
#pragma unroll
  for (int i = 0; i < W; ++i) ring[i] = 0.f;

#pragma unroll 1
  for (int bi = nblk - 1; bi >= 0; --bi) {
    const int z0 = bi * BL;
    // re-expand T_k inside the block from its checkpoint
    float tseg[BL];
    {
      float T = sC[bi * kRayThreads];
#pragma unroll
      for (int j = 0; j < BL; ++j) {
        tseg[j] = T;
        if (z0 + j < c.Vz) {
          float sb, vox;
          T *= (1.f - occupancy(sB[(z0 + j) * kRayThreads], sb, vox));
        }
      }
    }
    // pointers at k = z0 + BL - 1 (outputs at z = k + R); they walk downwards
    float *gout = g_grid + col0 + (ptrdiff_t)(z0 + BL - 1 + R) * c.VV;
    const float *gpr =
        (EXTRA && g_probs) ? g_probs + (ptrdiff_t)(z0 + BL - 1) * (ptrdiff_t)pstride + oi : nullptr;
    const float *gvx =
        (EXTRA && g_voxels) ? g_voxels + col0 + (ptrdiff_t)(z0 + BL - 1) * c.VV : nullptr;
    float kf = (float)(z0 + BL - 1);
#pragma unroll
    for (int j = BL - 1; j >= 0; --j) {
      const int k = z0 + j;
      float gB = 0.f;
      if (k < c.Vz) {
        const float bz = sB[k * kRayThreads];
        float sb, vox;
        const float v = occupancy(bz, sb, vox);
        float a = fmaf(fmaf(kf, c.inv_z, c.depth0), gd, gm);
        if (EXTRA && gpr) a += __ldg(gpr);
        if (j == 0 && k == 0) a *= c.exp_clip;
        float gv = tseg[j] * (a - D);
        D = fmaf(a, v, (1.f - v) * D);
        gv = (vox >= c.lo_c && vox <= c.hi_c) ? gv : 0.f;
        if (EXTRA && gvx) gv += __ldg(gvx);
        gv = (sb >= c.lo_s && sb <= c.hi_s) ? gv : 0.f;
        ds = fmaf(gv, bz, ds);
        gB = gv * s;
      }
      ring[j % W] = gB;
      // out[z] = sum_t kz[2R-t] * in[k + t], z = k + R   (adjoint = reversed taps)
      if (k + R < c.Vz) *gout = ring_dot<R>(ring, kz, j % W, true);
      gout -= c.VV;
      if (EXTRA) {
        if (gpr) gpr -= pstride;
        if (gvx) gvx -= c.VV;
      }
      kf -= 1.f;
    }
  }
  if (R > 0) {
    // flush: inputs k = -1 .. -R are zero; they complete the outputs z = R-1 .. 0
    float *gout = g_grid + col0 + (ptrdiff_t)(R - 1) * c.VV;
#pragma unroll
    for (int j = BL - 1; j >= BL - R; --j) {
      ring[j % W] = 0.f;
      if (j - BL + R < c.Vz) *gout = ring_dot<R>(ring, kz, j % W, true);
      gout -= c.VV;
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
blur_z_kernel(const float *src, float *dst, int Vz, int VV, const Taps<R> kz) {
  const int ray = blockIdx.x * kRayThreads + threadIdx.x;
  const int b = ray / VV, yx = ray - b * VV;
  const float *col = src + (size_t)b * Vz * VV + yx;
  float *out = dst + (size_t)b * Vz * VV + yx;
  stream_blur_z<R>(col, Vz, VV, kz, [&](int, int, float bz) {
    *out = bz;
    out += VV;
  });
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

static int check_ray_geometry(int V) {
  if ((V * V) % kRayThreads != 0) {
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

int launch_blurz_drc_fwd(const DrcArgs &a, const float *tz, int kz, float *bsave, float *mask,
                         float *depth, float *voxels, float *probs, cudaStream_t s) {
  if (int e = check_ray_geometry(a.V)) return e;
  const RayConst c = make_ray_const(a);
  const int r = z_radius(tz, kz);
  const int blocks = a.P * a.V * a.V / kRayThreads;
  if (voxels || probs) {
    DPC_DISPATCH_R(r, blurz_drc_fwd_kernel<R, true><<<blocks, kRayThreads, 0, s>>>(
                          a.grid, a.scale, c, z_taps<R>(tz, kz, r), bsave, mask, depth, voxels,
                          probs));
  } else {
    DPC_DISPATCH_R(r, blurz_drc_fwd_kernel<R, false><<<blocks, kRayThreads, 0, s>>>(
                          a.grid, a.scale, c, z_taps<R>(tz, kz, r), bsave, mask, depth, nullptr,
                          nullptr));
  }
  return check_launch("blurz_drc_fwd");
}

template <int R, bool EXTRA>
static void launch_bwd_one(const DrcArgs &a, const RayConst &c, const Taps<R> &taps,
                           const float *g_mask, const float *g_depth, const float *g_probs,
                           const float *g_voxels, float *g_grid, float *scale_partials,
                           cudaStream_t s) {
  constexpr int BL = (R == 0) ? 16 : 2 * R + 1;
  const int nblk = (a.Vz + BL - 1) / BL;
  const size_t smem = (size_t)(a.Vz + nblk) * kRayThreads * sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(drc_blurz_bwd_kernel<R, EXTRA>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  const int blocks = a.P * a.V * a.V / kRayThreads;
  drc_blurz_bwd_kernel<R, EXTRA><<<blocks, kRayThreads, smem, s>>>(
      a.grid, a.scale, c, taps, g_mask, g_depth, g_probs, g_voxels, g_grid,
      a.scale ? scale_partials : nullptr);
}

int launch_drc_blurz_bwd(const DrcArgs &a, const float *tz, int kz, const float *g_mask,
                         const float *g_depth, const float *g_probs, const float *g_voxels,
                         float *g_grid, float *scale_partials, cudaStream_t s) {
  if (int e = check_ray_geometry(a.V)) return e;
  const RayConst c = make_ray_const(a);
  const int r = z_radius(tz, kz);
  if (g_probs || g_voxels) {
    DPC_DISPATCH_R(r, launch_bwd_one<R, true>(a, c, z_taps<R>(tz, kz, r), g_mask, g_depth, g_probs,
                                              g_voxels, g_grid, scale_partials, s));
  } else {
    DPC_DISPATCH_R(r, launch_bwd_one<R, false>(a, c, z_taps<R>(tz, kz, r), g_mask, g_depth, nullptr,
                                               nullptr, g_grid, scale_partials, s));
  }
  return check_launch("drc_blurz_bwd");
}

int launch_blur_z(const float *src, float *dst, int P, int Vz, int V, const float *tz, int kz,
                  cudaStream_t s) {
  if (int e = check_ray_geometry(V)) return e;
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
