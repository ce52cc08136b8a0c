// Z pass of the Gaussian blur fused with the DRC ray march (forward), and the
// DRC reverse scan fused with the Z-pass adjoint (backward).
//
// Reference: point_cloud_to.py:97 (third conv3d, kernel [1,1,K,1,1]), :218-222
// (occupancy scaling + clamp), drc.py:48-106 (ray-termination probabilities,
// log-sum form with clip_val 'unity' padding), :114-129 (silhouette), :145-160
// (expected depth) and the two Y flips point_cloud_to.py:239, 242.
//
// One thread owns TWO x-adjacent rays (b, y, 2i) and (b, y, 2i+1); a warp owns
// 64 consecutive x, so every global access is a 64-bit access in a fully used
// 256-byte span.  The pair rides in one 64-bit register, so the Z blur -- the
// bulk of the arithmetic -- issues as packed fma.rn.f32x2 (SASS FFMA2) with the
// tap as a scalar uniform-register operand: half the issue slots of scalar
// FFMA in kernels that ncu shows to be issue-bound, not pipe-bound.
//
// The blur is a register ring of L >= 2R+1 pairs indexed at compile time (the
// z loop is unrolled by L; V is a template parameter so every address inside a
// block is pointer + immediate).  The input for step z+R+D is loaded straight
// into its ring slot at step z (D = L-(2R+1) steps of prefetch distance), each
// blurred value is consumed by the ray march as soon as it is produced and is
// written back IN PLACE over the column it came from (an input is always
// loaded R+D steps before its slot is overwritten), so the tensor the backward
// needs costs no extra memory.
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
// ([k][thread], conflict-free) and checkpoints T at every block of L steps;
// sweep 2 walks the blocks in reverse: it re-expands T inside the block from
// the checkpoint, runs the D recursion downwards, applies the clip/clamp
// gates and the scale, and pushes the result through the register ring to
// apply the Z-blur adjoint on the way out.  dL/dscale is reduced per block
// in fixed order.
#include <math_constants.h>

#include "common.cuh"

namespace dpc {

// ---- A/B switches (scripts/build_variant.sh <name> -D<switch>=<value>) -------------------------
// Every default is the measured winner; DESIGN.md section 5 lists the numbers.
#ifndef DPC_VZ_FWD
#define DPC_VZ_FWD 1           // compile-time depth in the forward ray kernel (0: run-time range checks)
#endif
#ifndef DPC_VZ_BWD
#define DPC_VZ_BWD 1           // ... and in the general-layout backward
#endif
#ifndef DPC_FWD_MINB
#define DPC_FWD_MINB 1         // resident CTAs per SM the forward's registers are sized for (1: no cap)
#endif
#ifndef DPC_BWD_MINB
#define DPC_BWD_MINB 4         // ... and the backward's (4: 252 registers; 6: 168; 8: 128 + spills)
#endif
#ifndef DPC_FWD_L10
#define DPC_FWD_L10 32         // ring / ray-block length at tap radius 10 (24, 28, 32 measured; round 2: 32 is 0.6 us ahead in the backward, two ray blocks instead of three)
#endif
#ifndef DPC_BWD_THREADS
#define DPC_BWD_THREADS 64     // ray pairs per backward CTA (32, 64, 128: no difference)
#endif
#ifndef DPC_RING_NACC
#define DPC_RING_NACC 2        // independent accumulation chains of the ring dot product
#endif
// DPC_PROBE_FEW_TAPS / DPC_PROBE_NO_STORE: timing probes (WRONG results).
constexpr int kFwdThreads = 128;                // ray pairs per CTA (forward)
constexpr int kBwdThreads = DPC_BWD_THREADS;    // ray pairs per CTA (backward)

int drc_scale_partial_blocks(int V) { return V * V / (2 * kBwdThreads); }

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// clamp of both halves (min/max have no packed form)
__device__ __forceinline__ u64 clamp2(u64 x, float lo, float hi) {
  float a, b;
  unpack2(x, a, b);
  return pack2(fminf(fmaxf(a, lo), hi), fminf(fmaxf(b, lo), hi));
}
// keep a half of g where the matching halves of x and y are equal, else 0
__device__ __forceinline__ u64 keep_if_equal2(u64 g, u64 x, u64 y) {
  float g0, g1, x0, x1, y0, y1;
  unpack2(g, g0, g1);
  unpack2(x, x0, x1);
  unpack2(y, y0, y1);
  return pack2(x0 == y0 ? g0 : 0.f, x1 == y1 ? g1 : 0.f);
}

// Ring length for a tap radius.  Backward (fed from shared memory): 2R+1 taps
// + 3.  Forward (fed from global memory): 2R+1 taps + the load-ahead distance
// that keeps enough bytes in flight per SM to cover DRAM latency.
template <int R> struct RingLen { static constexpr int L = 2 * R + 1 + 3; };
template <> struct RingLen<0> { static constexpr int L = 8; };
template <> struct RingLen<5> { static constexpr int L = 16; };
// (the forward and the fast backward share the length: the forward checkpoints the
// transmittance at the ray-block starts the backward resumes from.  At radius 10: 32 up to 64^3
// -- two ray blocks instead of three, backward 35.5 -> 34.8 us -- and 28 at 128^3, where 32 gains
// 20 us per launch forward but loses 62 backward.)
template <int R, int V = 64> struct FwdRingLen { static constexpr int L = RingLen<R>::L; };
template <int V> struct FwdRingLen<0, V> { static constexpr int L = 16; };
template <int V> struct FwdRingLen<7, V> { static constexpr int L = 24; };    // 16 steps of load-ahead, as at radius 10
template <int V> struct FwdRingLen<10, V> { static constexpr int L = V <= 64 ? DPC_FWD_L10 : 28; };

struct RayConst {
  int P, Vz;          // P = projections in the whole batch (probs stride)
  float inv_z, depth0, max_depth, exp_clip;
  float lo_s, hi_s;   // clamp(s*B, 0, 1) bounds; +-inf when there is no scaling factor
  float lo_c, hi_c;   // DRC clip bounds;        +-inf for the product form
  // clamp(clamp(x, lo_s, hi_s), lo_c, hi_c) == clamp(x, lo, hi) with the nested
  // bounds below, and BOTH clamp gates are open exactly when the clamp left x
  // unchanged (torch's clamp backward passes the closed interval)
  float lo, hi;
  int flip_y, has_scale;
};

static RayConst make_ray_const(const DrcArgs &a) {
  RayConst c;
  c.P = a.P_total; c.Vz = a.Vz;
  c.inv_z = 1.0f / (float)a.Vz;
  c.depth0 = a.cam_dist - 0.5f;
  c.max_depth = a.max_depth;
  c.exp_clip = a.logsum ? (float)exp((double)a.clip) : 1.0f;
  c.has_scale = a.scale != nullptr;
  c.lo_s = c.has_scale ? 0.f : -INFINITY;
  c.hi_s = c.has_scale ? 1.f : INFINITY;
  c.lo_c = a.logsum ? a.clip : -INFINITY;
  c.hi_c = a.logsum ? 1.0f - a.clip : INFINITY;
  c.lo = fmaxf(c.lo_s, c.lo_c);
  c.hi = fminf(c.hi_s, c.hi_c);
  c.flip_y = a.flip_y;
  return c;
}

// sum_t k[t] * ring[(first + t) % L]  (reversed: k[2R - t]), packed pairs
template <int R, int L>
__device__ __forceinline__ u64 ring_dot2(const u64 (&ring)[L], const u64 (&k2)[2 * R + 1],
                                         int first /*compile-time*/, bool reversed) {
  constexpr int W = 2 * R + 1;
  // NACC independent accumulation chains (fixed association, so reproducible)
  constexpr int NACC = W >= 12 ? DPC_RING_NACC : (W >= 3 ? 3 : 1);
  u64 acc[NACC];
#pragma unroll
  for (int t = 0; t < W; ++t) {
    const u64 v = ring[(first + t) % L];
    const u64 k = reversed ? k2[W - 1 - t] : k2[t];
    // the first term of every chain is a plain product: no zeroed accumulator registers
#ifdef DPC_PROBE_FEW_TAPS
    if (t >= 2 * NACC) continue;      // timing probe: 2 * NACC taps of 2R+1
#endif
    acc[t % NACC] = t < NACC ? mul2(k, v) : fma2(k, v, acc[t % NACC]);
  }
#pragma unroll
  for (int n = NACC; n > 1; n = (n + 1) / 2)
#pragma unroll
    for (int i = 0; i < n / 2; ++i) acc[i] = add2(acc[i], acc[n - 1 - i]);
  return acc[0];
}

template <int V>
__device__ __forceinline__ void pair_index(const RayConst &c, int threads, int &b, int &yx,
                                           int &out_idx) {
  constexpr int VV = V * V;
  const int pair = blockIdx.x * threads + threadIdx.x;
  b = pair / (VV / 2);
  yx = 2 * (pair - b * (VV / 2));            // even x
  const int y = yx / V, x = yx - y * V;
  const int yo = c.flip_y ? (V - 1 - y) : y;
  out_idx = b * VV + yo * V + x;
}

// Streams the pair blurZ(col)_z, z = 0..Vz-1, to sink(j, z, pair); j is the
// compile-time position inside the current block of L steps.  SAVE writes the
// blurred pair back over the input column (bs may alias col).
// VZ > 0: the depth is a compile-time constant (== Vz).  The blocks whose loads
// and outputs are all in range run in a rolled loop WITHOUT any range check;
// the last block(s) are unrolled with a compile-time z0, so their checks fold
// away.  With a runtime depth (VZ == 0) every step carries two range checks --
// a branch per step, which also stops the scheduler from overlapping steps.
template <int V, int R, bool SAVE, int VZ, typename Sink>
__device__ __forceinline__ void stream_blur_z2(const float *col, float *bs, int Vz_rt,
                                               const Taps<R> &taps, Sink &&sink) {
  constexpr int W = 2 * R + 1, L = FwdRingLen<R, V>::L, AHEAD = R + (L - W);   // load-ahead in steps
  constexpr int VV = V * V;
  const int Vz = VZ ? VZ : Vz_rt;
  u64 k2[W];
#pragma unroll
  for (int t = 0; t < W; ++t) k2[t] = pack2(taps.k[t], taps.k[t]);
  u64 ring[L];
#pragma unroll
  for (int i = 0; i < L; ++i) ring[i] = 0;
#pragma unroll
  for (int i = 0; i < AHEAD; ++i)
    if (i < Vz) ring[i] = *reinterpret_cast<const u64 *>(col + (size_t)i * VV);
  // one step; `checked` folds to a constant in the unrolled tail blocks
  auto step = [&](const int j, const int z, const bool in_range, const bool load_ok) __attribute__((always_inline)) {
    if (in_range) {
      ring[(j + AHEAD) % L] =
          load_ok ? *reinterpret_cast<const u64 *>(col + (size_t)(j + AHEAD) * VV) : 0ull;
      const u64 b2 = ring_dot2<R, L>(ring, k2, (j + L - R) % L, false);
      const u64 keep2 = sink(j, z, b2);      // what the backward wants to find in this voxel pair
      // next read: the backward.  Streaming only where a half-batch of grids can live in L2
      // (64^3: 32 MiB); at 128^3 (512 MiB) the hint costs 10 us per launch and keeps nothing
      if (SAVE) {
        if (V <= 64) st_stream(reinterpret_cast<u64 *>(bs + (size_t)j * VV), keep2);
        else *reinterpret_cast<u64 *>(bs + (size_t)j * VV) = keep2;
      }
    }
  };
  if (VZ) {
    constexpr int NBLK = (VZ + L - 1) / L;
    constexpr int NSAFE = VZ > AHEAD ? (VZ - AHEAD) / L : 0;   // (bi+1) L - 1 + AHEAD < VZ
#pragma unroll 1
    for (int bi = 0; bi < NSAFE; ++bi) {
#pragma unroll
      for (int j = 0; j < L; ++j) step(j, bi * L + j, true, true);
      col += (size_t)L * VV;
      if (SAVE) bs += (size_t)L * VV;
    }
#pragma unroll
    for (int bi = NSAFE; bi < NBLK; ++bi) {
#pragma unroll
      for (int j = 0; j < L; ++j) step(j, bi * L + j, bi * L + j < VZ, bi * L + j + AHEAD < VZ);
      col += (size_t)L * VV;
      if (SAVE) bs += (size_t)L * VV;
    }
  } else {
#pragma unroll 1
    for (int z0 = 0; z0 < Vz; z0 += L) {
#pragma unroll
      for (int j = 0; j < L; ++j) step(j, z0 + j, z0 + j < Vz, z0 + j + AHEAD < Vz);
      col += (size_t)L * VV;
      if (SAVE) bs += (size_t)L * VV;
    }
  }
}

// EXTRA: the optional voxels / probs outputs exist.  `grid` and `bsave` may
// alias (in-place save), so neither is __restrict__.
// FAST (needs VZ, SAVE, !EXTRA, log-sum DRC): instead of B the kernel saves what the backward
// actually consumes -- v = clip(s B) with the sign bit set where the clip changed the value
// (v >= clip_val > 0, so the sign is free; the gate of torch's clamp backward is "unchanged") --
// and the transmittance T at the start of every block of L steps (tck), so that the backward
// needs neither a forward sweep nor any clamp arithmetic.
template <int V, int R, bool EXTRA, bool SAVE, int VZ, bool FAST>
__global__ void __launch_bounds__(kFwdThreads, VZ ? DPC_FWD_MINB : 1)
blurz_drc_fwd_kernel(const float *grid, const float *__restrict__ scale, RayConst c,
                     const Taps<R> kz, float *bsave, float *__restrict__ mask,
                     float *__restrict__ depth, float *__restrict__ voxels,
                     float *__restrict__ probs, float *__restrict__ tck, int ck_slots) {
  static_assert(!FAST || (VZ > 0 && SAVE && !EXTRA), "fast ray state: compile-time depth, saved, no extras");
  constexpr int VV = V * V;
  int b, yx, oi;
  pair_index<V>(c, kFwdThreads, b, yx, oi);
  const size_t col0 = (size_t)b * c.Vz * VV + yx;
  pdl_wait();             // the XY-blurred grid comes from blur_xy
  pdl_release();
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  const u64 s2 = pack2(s, s), one2 = pack2(1.f, 1.f), neg2 = pack2(-1.f, -1.f);
  const u64 ec2 = pack2(c.exp_clip, c.exp_clip);
  u64 T2 = one2, m2 = 0, d2 = 0;   // transmittance, mask and depth sums of the two rays
  float kf = 0.f;
  float *vx = (EXTRA && voxels) ? voxels + col0 : nullptr;
  float *pr = (EXTRA && probs) ? probs + oi : nullptr;
  const size_t pstride = (size_t)c.P * VV;
  u64 *ckp = FAST ? reinterpret_cast<u64 *>(tck + (size_t)b * ck_slots * VV + yx) : nullptr;
  stream_blur_z2<V, R, SAVE, VZ>(grid + col0, SAVE ? bsave + col0 : nullptr, c.Vz, kz,
                                 [&](int j, int z, u64 b2) -> u64 {
    if (FAST && j == 0 && z > 0) {   // block start: checkpoint the transmittance
      if (V <= 64) st_stream(ckp, T2); else *ckp = T2;      // next read: the backward
      ckp += VV / 2;
    }
    const float psi = fmaf(kf, c.inv_z, c.depth0);
    kf += 1.f;
    const u64 sb2 = mul2(s2, b2);
    u64 v2;
    if (EXTRA) {
      const u64 vox2 = clamp2(sb2, c.lo_s, c.hi_s);
      v2 = clamp2(vox2, c.lo_c, c.hi_c);
      if (vx) { *reinterpret_cast<u64 *>(vx) = vox2; vx += VV; }
    } else {
      v2 = clamp2(sb2, c.lo, c.hi);   // the two nested clamps in one
    }
    u64 p2 = mul2(v2, T2);
    if (j == 0 && z == 0) p2 = mul2(p2, ec2);   // only block position 0 can be z == 0
    m2 = add2(m2, p2);
    d2 = fma2(pack2(psi, psi), p2, d2);
    T2 = mul2(T2, fma2(v2, neg2, one2));         // T *= (1 - v)
    if (EXTRA && pr) { *reinterpret_cast<u64 *>(pr) = p2; pr += pstride; }
    if (FAST) {
      float v0, v1, x0, x1;
      unpack2(v2, v0, v1);
      unpack2(sb2, x0, x1);
      return pack2(v0 == x0 ? v0 : -v0, v1 == x1 ? v1 : -v1);
    }
    return b2;
  });
  const u64 pz2 = mul2(ec2, T2);
  if (EXTRA && pr) *reinterpret_cast<u64 *>(pr) = pz2;
  *reinterpret_cast<u64 *>(mask + oi) = m2;
  if (depth) *reinterpret_cast<u64 *>(depth + oi) = fma2(pack2(c.max_depth, c.max_depth), pz2, d2);
}

// VZ > 0: compile-time depth (== c.Vz), see stream_blur_z2: the blocks that lie
// entirely inside the column run in rolled loops without range checks, the last
// block(s) are unrolled with a compile-time z0.
template <int V, int R, bool EXTRA, int VZ>
__global__ void __launch_bounds__(kBwdThreads, VZ ? DPC_BWD_MINB : 1)
drc_blurz_bwd_kernel(const float *__restrict__ bgrid, const float *__restrict__ scale, RayConst c,
                     const Taps<R> kz, const float *__restrict__ g_mask,
                     const float *__restrict__ g_depth, const float *__restrict__ g_probs,
                     const float *__restrict__ g_voxels, float *__restrict__ g_grid,
                     float *__restrict__ scale_partials, int *__restrict__ zero_ints, int n_zero) {
  constexpr int W = 2 * R + 1, L = RingLen<R>::L;   // block length == ring length
  pdl_release();          // head of the backward chain: launched without a programmatic edge
  if (zero_ints && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kBwdThreads) zero_ints[i] = 0;
  constexpr int VV = V * V;
  const int Vz = VZ ? VZ : c.Vz;
  extern __shared__ float2 sm2[];
  float2 *sB = sm2 + threadIdx.x;                        // [Vz][threads]    saved blurZ pair
  float2 *sC = sm2 + Vz * kBwdThreads + threadIdx.x;     // [nblk][threads]  T at block starts
  int b, yx, oi;
  pair_index<V>(c, kBwdThreads, b, yx, oi);
  const size_t col0 = (size_t)b * Vz * VV + yx;
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  const int nblk = (Vz + L - 1) / L;
  // blocks [0, n_full) lie inside the column; blocks [0, n_store) also have every output row
  // z = k + R inside it.  Compile-time when VZ is; 0 otherwise (every step checked).
  constexpr int NFULL = VZ / L, NSTORE = VZ > R ? (VZ - R) / L : 0, NBLK = (VZ + L - 1) / L;

  const u64 s2 = pack2(s, s), one2 = pack2(1.f, 1.f), neg2 = pack2(-1.f, -1.f);
  const u64 ec2 = pack2(c.exp_clip, c.exp_clip);
  u64 *sB2 = reinterpret_cast<u64 *>(sB);
  u64 *sC2 = reinterpret_cast<u64 *>(sC);
  // 1 - v of the pair, v = clamp(clamp(s B, lo_s, hi_s), lo_c, hi_c) = clamp(s B, lo, hi)
  auto one_minus_v = [&](u64 b2) -> u64 {
    return fma2(clamp2(mul2(s2, b2), c.lo, c.hi), neg2, one2);
  };

  // ---- sweep 1: stage the column pair, checkpoint the transmittance ----
  {
    const float *ld = bgrid + col0;
    u64 T2 = one2;
    // block bi; `full`: every step is inside the column
    auto stage_block = [&](const int bi, const bool full) __attribute__((always_inline)) {
      sC2[bi * kBwdThreads] = T2;
      u64 vals[L];
#pragma unroll
      for (int j = 0; j < L; ++j)
        vals[j] = (full || bi * L + j < Vz) ? __ldg(reinterpret_cast<const u64 *>(ld + (size_t)j * VV)) : 0ull;
      ld += (size_t)L * VV;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        const int z = bi * L + j;
        if (full || z < Vz) {
          sB2[z * kBwdThreads] = vals[j];
          T2 = mul2(T2, one_minus_v(vals[j]));
        }
      }
    };
    if (VZ) {
#pragma unroll 1
      for (int bi = 0; bi < NFULL; ++bi) stage_block(bi, true);
#pragma unroll
      for (int bi = NFULL; bi < NBLK; ++bi) stage_block(bi, false);
    } else {
#pragma unroll 1
      for (int bi = 0; bi < nblk; ++bi) stage_block(bi, false);
    }
  }
  // ---- sweep 2: reverse scan + Z-blur adjoint ----
  const float2 gm = g_mask ? __ldg(reinterpret_cast<const float2 *>(g_mask + oi)) : make_float2(0.f, 0.f);
  const float2 gd = g_depth ? __ldg(reinterpret_cast<const float2 *>(g_depth + oi)) : make_float2(0.f, 0.f);
  const size_t pstride = (size_t)c.P * VV;
  u64 D2 = pack2(c.max_depth * gd.x, c.max_depth * gd.y);
  if (EXTRA && g_probs)
    D2 = add2(D2, __ldg(reinterpret_cast<const u64 *>(g_probs + (size_t)Vz * pstride + oi)));
  D2 = mul2(D2, ec2);
  u64 ds2 = 0;
  u64 k2[W];
#pragma unroll
  for (int t = 0; t < W; ++t) k2[t] = pack2(kz.k[t], kz.k[t]);
  u64 ring[L];
#pragma unroll
  for (int i = 0; i < L; ++i) ring[i] = 0;

  // block bi in reverse; `full`: every step inside the column, `store_all`: every output row too
  auto reverse_block = [&](const int bi, const bool full, const bool store_all) __attribute__((always_inline)) {
    const int z0 = bi * L;
    // re-expand T_k inside the block from its checkpoint
    u64 tseg[L];
    {
      u64 T2 = sC2[bi * kBwdThreads];
#pragma unroll
      for (int j = 0; j < L; ++j) {
        tseg[j] = T2;
        if (full || z0 + j < Vz) T2 = mul2(T2, one_minus_v(sB2[(z0 + j) * kBwdThreads]));
      }
    }
    // block-base pointers; inside the block every offset is an immediate
    float *gout = g_grid + col0 + (size_t)(z0 + R) * VV;               // row z = k + R at j = 0
    const float *gpr = (EXTRA && g_probs) ? g_probs + (size_t)z0 * pstride + oi : nullptr;
    const float *gvx = (EXTRA && g_voxels) ? g_voxels + col0 + (size_t)z0 * VV : nullptr;
    const float kf0 = (float)z0;
#pragma unroll
    for (int j = L - 1; j >= 0; --j) {
      const int k = z0 + j;
      u64 gB2 = 0;
      if (full || k < Vz) {
        const u64 b2 = sB2[k * kBwdThreads];
        const u64 sb2 = mul2(s2, b2);
        const float psi = fmaf(kf0 + (float)j, c.inv_z, c.depth0);
        u64 a2 = pack2(fmaf(psi, gd.x, gm.x), fmaf(psi, gd.y, gm.y));
        if (EXTRA && gpr) a2 = add2(a2, __ldg(reinterpret_cast<const u64 *>(gpr + (size_t)j * pstride)));
        if (j == 0 && k == 0) a2 = mul2(a2, ec2);
        u64 v2, gv2;
        if (EXTRA) {
          // g_voxels enters between the two clamp gates: keep them separate
          const u64 vox2 = clamp2(sb2, c.lo_s, c.hi_s);
          v2 = clamp2(vox2, c.lo_c, c.hi_c);
          gv2 = keep_if_equal2(mul2(tseg[j], fma2(D2, neg2, a2)), v2, vox2);
          if (gvx) gv2 = add2(gv2, __ldg(reinterpret_cast<const u64 *>(gvx + (size_t)j * VV)));
          gv2 = keep_if_equal2(gv2, vox2, sb2);
        } else {
          v2 = clamp2(sb2, c.lo, c.hi);
          gv2 = keep_if_equal2(mul2(tseg[j], fma2(D2, neg2, a2)), v2, sb2);   // T (a - D), gated
        }
        D2 = fma2(a2, v2, mul2(fma2(v2, neg2, one2), D2));   // D = a v + (1 - v) D
        ds2 = fma2(gv2, b2, ds2);
        gB2 = mul2(gv2, s2);
      }
      ring[j] = gB2;
      // out[z] = sum_t kz[2R-t] * in[k + t], z = k + R   (adjoint = reversed taps);
      // in[k + t] sits in slot (j + t) % L
      if (store_all || k + R < Vz)
        *reinterpret_cast<u64 *>(gout + (size_t)j * VV) = ring_dot2<R, L>(ring, k2, j, true);
    }
  };
  if (VZ) {
#pragma unroll
    for (int bi = NBLK - 1; bi >= NSTORE; --bi) reverse_block(bi, bi < NFULL, false);
#pragma unroll 1
    for (int bi = NSTORE - 1; bi >= 0; --bi) reverse_block(bi, true, true);
  } else {
#pragma unroll 1
    for (int bi = nblk - 1; bi >= 0; --bi) reverse_block(bi, false, false);
  }
  if (R > 0) {
    // flush: inputs k = -1 .. -R are zero; they complete the outputs z = R-1 .. 0
    float *gout = g_grid + col0;
#pragma unroll
    for (int j = L - 1; j >= L - R; --j) {
      ring[j] = 0;
      const int z = j - L + R;
      if (z < Vz) *reinterpret_cast<u64 *>(gout + (size_t)z * VV) = ring_dot2<R, L>(ring, k2, j, true);
    }
  }
  // ---- dL/dscale: fixed-order block reduction ----
  if (scale_partials) {
    __shared__ float red[kBwdThreads / 32];
    float ds, ds_hi;
    unpack2(ds2, ds, ds_hi);
    ds += ds_hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ds += __shfl_down_sync(0xffffffffu, ds, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kBwdThreads / 32; ++w) v += red[w];
      scale_partials[blockIdx.x] = v;  // blocks are projection-major
    }
  }
}

// ---- bulk asynchronous copy (TMA engine, no tensor map: rows are contiguous) + mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// The same copy with an L2 evict-first policy: the saved ray state is read exactly once, and at
// 32 MiB per half-batch it otherwise pushes the gradient grid this kernel WRITES out of L2 before
// the plane kernel reads it (ncu, steady state: the plane kernel read 37 MB per launch from DRAM).
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
#ifndef DPC_GGRID_KEEP
#define DPC_GGRID_KEEP 0       // A/B: gradient-grid stores with an L2 evict-last policy
#endif
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep(u64 *p, u64 v, uint64_t pol) {
  if (DPC_GGRID_KEEP)
    asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
  else
    *p = v;
}
__device__ __forceinline__ void bulk_g2s_once(void *dst, const void *src, uint32_t bytes, uint64_t *bar,
                                              uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// Backward on the fast ray state (see blurz_drc_fwd_kernel<FAST>): one sweep.
// The CTA's tile of the saved grid -- 64 ray pairs x VZ planes, every plane row 512
// contiguous bytes -- is brought to shared memory by the bulk-copy engine
// (cp.async.bulk, one instruction per row issued by warp 0, completion counted
// on one mbarrier per ray block, deepest block first): no thread spends an
// instruction or a register on the staging, the whole tile is in flight at once,
// and the reverse scan starts on the deepest block while the rest is landing.
// T at the block starts comes from the forward's checkpoints; |v| and the gate
// come straight from the saved value, so the scan has no clamp arithmetic;
// dL/dscale uses B = v / s on the open gates (the only voxels that count).
template <int V, int R>
__global__ void __launch_bounds__(kBwdThreads, DPC_BWD_MINB)
drc_blurz_bwd_fast_kernel(const float *__restrict__ vgrid, const float *__restrict__ tck,
                          int ck_slots, const float *__restrict__ scale, RayConst c,
                          const Taps<R> kz, const float *__restrict__ g_mask,
                          const float *__restrict__ g_depth, float *__restrict__ g_grid,
                          float *__restrict__ scale_partials, int *__restrict__ zero_ints,
                          int n_zero, const LossGrad lg) {
  constexpr int VZ = V, W = 2 * R + 1, L = FwdRingLen<R, V>::L, VV = V * V;
  constexpr int NBLK = (VZ + L - 1) / L, NFULL = VZ / L, NSTORE = VZ > R ? (VZ - R) / L : 0;
  constexpr uint32_t ROW_BYTES = kBwdThreads * sizeof(u64);
  pdl_release();          // head of the backward chain: launched without a programmatic edge
  if (zero_ints && blockIdx.x == 0)
    for (int i = threadIdx.x; i < n_zero; i += kBwdThreads) zero_ints[i] = 0;
  extern __shared__ __align__(128) unsigned char smraw[];
  u64 *tile = reinterpret_cast<u64 *>(smraw);                              // [VZ][threads]
  uint64_t *bars = reinterpret_cast<uint64_t *>(smraw + (size_t)VZ * ROW_BYTES);   // [NBLK]
  const int tid = threadIdx.x;
  // chain slot bj works on projection b (winner-only backward of render_loss: b = bmap[bj]); the
  // saved state, the scale and the upstream gradients belong to b, the gradient grid to the slot
  int bj, yx, oi;
  pair_index<V>(c, kBwdThreads, bj, yx, oi);
  const int b = lg.bmap ? __ldg(lg.bmap + bj) : bj;
  oi += (b - bj) * VV;
  const size_t col0 = (size_t)b * VZ * VV + yx, colo = (size_t)bj * VZ * VV + yx;
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < NBLK; ++i) mbar_init(bars + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid < 32) {
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < NBLK; ++i)
        mbar_expect_tx(bars + i, (uint32_t)((i < NFULL ? L : VZ - NFULL * L) * ROW_BYTES));
    }
    __syncwarp();
    // rows issued by the 32 lanes of warp 0, deepest rows first (measured: 35.2 us; all 64 rows
    // from one thread: 37.7 us -- the last rows are issued too late)
    const float *row0 = vgrid + (col0 - (size_t)(2 * tid));     // the CTA's first pair (tid == lane here)
    const uint64_t once = l2_evict_first_policy();
    for (int z = VZ - 1 - tid; z >= 0; z -= 32) {
      if (DPC_CACHE_HINTS && V <= 64)
        bulk_g2s_once(tile + (size_t)z * kBwdThreads, row0 + (size_t)z * VV, ROW_BYTES, bars + z / L, once);
      else
        bulk_g2s(tile + (size_t)z * kBwdThreads, row0 + (size_t)z * VV, ROW_BYTES, bars + z / L);
    }
  }
  const float s = c.has_scale ? __ldg(scale + b) : 1.f;
  const u64 s2 = pack2(s, s), one2 = pack2(1.f, 1.f), neg2 = pack2(-1.f, -1.f);
  const u64 ec2 = pack2(c.exp_clip, c.exp_clip);
  const uint64_t keep_pol = DPC_GGRID_KEEP ? l2_evict_last_policy() : 0ull;
  // transmittance at the block starts (block 0 starts at 1)
  u64 tstart[NBLK];
  tstart[0] = one2;
#pragma unroll
  for (int i = 1; i < NBLK; ++i)
    tstart[i] = __ldg(reinterpret_cast<const u64 *>(tck + ((size_t)b * ck_slots + (i - 1)) * VV + yx));
  float2 gm = g_mask ? __ldg(reinterpret_cast<const float2 *>(g_mask + oi)) : make_float2(0.f, 0.f);
  if (lg.gt) {
    // dL/dmask of the candidate-selection loss for the winner, built here instead of read:
    // k (pool(gt) - pred), rounded like candidate_loss_bwd_kernel (model_pc_to.py:430-437)
    const int bv = b / lg.C, pix = oi - b * VV, yo = pix / V, xo = pix - yo * V, n = lg.G / V;
    const float k = __ldg(lg.kcoef + bv) * (lg.upstream ? __ldg(lg.upstream) : 1.f);
    const float2 pr = __ldg(reinterpret_cast<const float2 *>(lg.pred + oi));
    const float *g = lg.gt + (size_t)bv * lg.G * lg.G;
    gm.x = k * (pooled_gt(g, lg.G, n, yo, xo) - pr.x);
    gm.y = k * (pooled_gt(g, lg.G, n, yo, xo + 1) - pr.y);
  }
  const float2 gd = g_depth ? __ldg(reinterpret_cast<const float2 *>(g_depth + oi)) : make_float2(0.f, 0.f);
  u64 D2 = mul2(pack2(c.max_depth * gd.x, c.max_depth * gd.y), ec2);
  u64 ds2 = 0;
  u64 k2[W];
#pragma unroll
  for (int t = 0; t < W; ++t) k2[t] = pack2(kz.k[t], kz.k[t]);
  u64 ring[L];
#pragma unroll
  for (int i = 0; i < L; ++i) ring[i] = 0;
  const u64 *col = tile + tid;
  auto abs2 = [](u64 w) -> u64 {
    float a, b;
    unpack2(w, a, b);
    return pack2(fabsf(a), fabsf(b));
  };

  auto reverse_block = [&](const int bi, const u64 T0, const bool full, const bool store_all)
                           __attribute__((always_inline)) {
    const int z0 = bi * L;
    mbar_wait(bars + bi, 0);
    u64 tseg[L];
    {
      u64 T2 = T0;
#pragma unroll
      for (int j = 0; j < L; ++j) {
        tseg[j] = T2;
        if (full || z0 + j < VZ) T2 = mul2(T2, fma2(abs2(col[(z0 + j) * kBwdThreads]), neg2, one2));
      }
    }
    float *gout = g_grid + colo + (size_t)(z0 + R) * VV;               // row z = k + R at j = 0
    const float kf0 = (float)z0;
#pragma unroll
    for (int j = L - 1; j >= 0; --j) {
      const int k = z0 + j;
      u64 gB2 = 0;
      if (full || k < VZ) {
        const u64 w2 = col[k * kBwdThreads];
        const u64 v2 = abs2(w2);
        const float psi = fmaf(kf0 + (float)j, c.inv_z, c.depth0);
        u64 a2 = pack2(fmaf(psi, gd.x, gm.x), fmaf(psi, gd.y, gm.y));
        if (j == 0 && k == 0) a2 = mul2(a2, ec2);
        float x0, x1, w0, w1;
        unpack2(mul2(tseg[j], fma2(D2, neg2, a2)), x0, x1);        // T (a - D)
        unpack2(w2, w0, w1);
        const u64 gv2 = pack2(w0 > 0.f ? x0 : 0.f, w1 > 0.f ? x1 : 0.f);   // sign bit = gate closed
        D2 = fma2(a2, v2, mul2(fma2(v2, neg2, one2), D2));             // D = a v + (1 - v) D
        ds2 = fma2(gv2, v2, ds2);
        gB2 = mul2(gv2, s2);
      }
      ring[j] = gB2;
      if (store_all || k + R < VZ) {
#ifdef DPC_PROBE_NO_STORE
        ds2 = add2(ds2, ring_dot2<R, L>(ring, k2, j, true));     // timing probe: no g_grid store
#else
        st_keep(reinterpret_cast<u64 *>(gout + (size_t)j * VV), ring_dot2<R, L>(ring, k2, j, true), keep_pol);
#endif
      }
    }
  };
#pragma unroll
  for (int bi = NBLK - 1; bi >= NSTORE; --bi) reverse_block(bi, tstart[bi], bi < NFULL, false);
  if (NSTORE > 0) {
    // the remaining blocks are all alike: a rolled loop (tstart is indexed at run time only here)
#pragma unroll 1
    for (int bi = NSTORE - 1; bi >= 0; --bi) {
      u64 T0 = one2;
#pragma unroll
      for (int i = 1; i < NSTORE; ++i) T0 = (i == bi) ? tstart[i] : T0;
      reverse_block(bi, T0, true, true);
    }
  }
  if (R > 0) {
    // flush: inputs k = -1 .. -R are zero; they complete the outputs z = R-1 .. 0
    float *gout = g_grid + colo;
#pragma unroll
    for (int j = L - 1; j >= L - R; --j) {
      ring[j] = 0;
      const int z = j - L + R;
      if (z < VZ) st_keep(reinterpret_cast<u64 *>(gout + (size_t)z * VV), ring_dot2<R, L>(ring, k2, j, true), keep_pol);
    }
  }
  if (scale_partials) {
    __shared__ float red[kBwdThreads / 32];
    float ds, ds_hi;
    unpack2(ds2, ds, ds_hi);
    ds += ds_hi;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ds += __shfl_down_sync(0xffffffffu, ds, o);
    if ((tid & 31) == 0) red[tid >> 5] = ds;
    __syncthreads();
    if (tid == 0) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kBwdThreads / 32; ++w) v += red[w];
      // sum gv * B with B = v / s on the open gates (s == 0 closes every gate: the sum is 0)
      scale_partials[blockIdx.x] = s != 0.f ? v / s : 0.f;
    }
  }
}

template <int V, int R>
__global__ void __launch_bounds__(kFwdThreads)
blur_z_kernel(const float *src, float *dst, int Vz, const Taps<R> kz) {
  constexpr int VV = V * V;
  const int pair = blockIdx.x * kFwdThreads + threadIdx.x;
  const int b = pair / (VV / 2), yx = 2 * (pair - b * (VV / 2));
  const size_t col0 = (size_t)b * Vz * VV + yx;
  stream_blur_z2<V, R, true, 0>(src + col0, dst + col0, Vz, kz, [](int, int, u64 b2) { return b2; });
}

__global__ void __launch_bounds__(128)
depth_from_probs_kernel(const float *__restrict__ probs, float *__restrict__ depth, int P, int Vz,
                        int VV, float inv_z, float depth0, float max_depth) {
  const int i = blockIdx.x * 128 + threadIdx.x;
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

#define DPC_DISPATCH_R(r, ...)                                            \
  do {                                                                    \
    if (r == 0) { constexpr int R = 0; __VA_ARGS__; }                     \
    else if (r <= 2) { constexpr int R = 2; __VA_ARGS__; }                \
    else if (r <= 5) { constexpr int R = 5; __VA_ARGS__; }                \
    else if (r <= 7) { constexpr int R = 7; __VA_ARGS__; }                \
    else if (r <= 10) { constexpr int R = 10; __VA_ARGS__; }              \
    else { set_error("z tap radius %d > 10 unsupported", r); return DPC_ERR_ARG; } \
  } while (0)
#define DPC_DISPATCH_V(v, ...)                                            \
  do {                                                                    \
    if (v == 32) { constexpr int V = 32; __VA_ARGS__; }                   \
    else if (v == 64) { constexpr int V = 64; __VA_ARGS__; }              \
    else if (v == 128) { constexpr int V = 128; __VA_ARGS__; }            \
    else { set_error("vox_size %d unsupported (32, 64, 128)", v); return DPC_ERR_ARG; } \
  } while (0)

template <int V, int R>
static void launch_fwd_vr(const DrcArgs &a, const RayConst &c, const Taps<R> &taps, float *bsave,
                          float *mask, float *depth, float *voxels, float *probs, cudaStream_t s) {
  const int blocks = a.P * (V * V / 2) / kFwdThreads;
  const bool extra = voxels || probs;
#define DPC_FWD(EX, SV, VZ, FS)                                                             \
  launch_dep(blurz_drc_fwd_kernel<V, R, EX, SV, VZ, FS>, dim3(blocks), dim3(kFwdThreads), 0, s, \
             a.grid, a.scale, c, taps, bsave, mask, depth, voxels, probs, a.tck, a.ck_slots)
  // the training configuration (cubic grid, no optional outputs, saved state) gets the
  // kernel with the depth as a compile-time constant
  // (a.tck: api.cu fast_ray_state() has checked cubic grid, log-sum DRC, no optional outputs)
  if (bsave) {
    if (extra) DPC_FWD(true, true, 0, false);
    else if (a.tck && a.Vz == V) DPC_FWD(false, true, V, true);
    else if (DPC_VZ_FWD && a.Vz == V) DPC_FWD(false, true, V, false);
    else DPC_FWD(false, true, 0, false);
  } else {
    if (extra) DPC_FWD(true, false, 0, false); else DPC_FWD(false, false, 0, false);
  }
#undef DPC_FWD
}

int launch_blurz_drc_fwd(const DrcArgs &a, const float *tz, int kz, float *bsave, float *mask,
                         float *depth, float *voxels, float *probs, cudaStream_t s) {
  const RayConst c = make_ray_const(a);
  const int r = z_radius(tz, kz);
  DPC_DISPATCH_V(a.V, DPC_DISPATCH_R(r, launch_fwd_vr<V, R>(a, c, z_taps<R>(tz, kz, r), bsave, mask,
                                                           depth, voxels, probs, s)));
  return check_launch("blurz_drc_fwd");
}

template <int V, int R, bool EXTRA, int VZ>
static void launch_bwd_vz(const DrcArgs &a, const RayConst &c, const Taps<R> &taps,
                           const float *g_mask, const float *g_depth, const float *g_probs,
                           const float *g_voxels, float *g_grid, float *scale_partials,
                           int *zero_ints, int n_zero, cudaStream_t s) {
  constexpr int L = RingLen<R>::L;
  const int nblk = (a.Vz + L - 1) / L;
  const size_t smem = (size_t)(a.Vz + nblk) * kBwdThreads * sizeof(float2);
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(drc_blurz_bwd_kernel<V, R, EXTRA, VZ>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  }
  const int blocks = a.P * (V * V / 2) / kBwdThreads;
  drc_blurz_bwd_kernel<V, R, EXTRA, VZ><<<blocks, kBwdThreads, smem, s>>>(
      a.grid, a.scale, c, taps, g_mask, g_depth, g_probs, g_voxels, g_grid,
      a.scale ? scale_partials : nullptr, zero_ints, n_zero);
}

template <int V, int R, bool EXTRA>
static void launch_bwd_one(const DrcArgs &a, const RayConst &c, const Taps<R> &taps,
                           const float *g_mask, const float *g_depth, const float *g_probs,
                           const float *g_voxels, float *g_grid, float *scale_partials,
                           int *zero_ints, int n_zero, cudaStream_t s) {
  // the training configuration (cubic grid, gradients through mask/depth only) gets the
  // kernel with the depth as a compile-time constant
  if (DPC_VZ_BWD && !EXTRA && a.Vz == V)
    launch_bwd_vz<V, R, false, V>(a, c, taps, g_mask, g_depth, nullptr, nullptr, g_grid,
                                  scale_partials, zero_ints, n_zero, s);
  else
    launch_bwd_vz<V, R, EXTRA, 0>(a, c, taps, g_mask, g_depth, g_probs, g_voxels, g_grid,
                                  scale_partials, zero_ints, n_zero, s);
}

template <int V, int R>
static void launch_bwd_fast(const DrcArgs &a, const RayConst &c, const Taps<R> &taps,
                            const float *g_mask, const float *g_depth, float *g_grid,
                            float *scale_partials, int *zero_ints, int n_zero, cudaStream_t s,
                            const LossGrad &lg) {
  constexpr int L = FwdRingLen<R, V>::L, NBLK = (V + L - 1) / L;
  static_assert(NBLK * sizeof(uint64_t) <= 128, "mbarrier area");
  const size_t smem = (size_t)V * kBwdThreads * sizeof(u64) + 128;
  static DeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(drc_blurz_bwd_fast_kernel<V, R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)smem);
  }
  const int blocks = a.P * (V * V / 2) / kBwdThreads;
  drc_blurz_bwd_fast_kernel<V, R><<<blocks, kBwdThreads, smem, s>>>(
      a.grid, a.tck, a.ck_slots, a.scale, c, taps, g_mask, g_depth, g_grid,
      a.scale ? scale_partials : nullptr, zero_ints, n_zero, lg);
}

int launch_drc_blurz_bwd(const DrcArgs &a, const float *tz, int kz, const float *g_mask,
                         const float *g_depth, const float *g_probs, const float *g_voxels,
                         float *g_grid, float *scale_partials, int *zero_ints, int n_zero,
                         cudaStream_t s, const LossGrad *lg) {
  const RayConst c = make_ray_const(a);
  const int r = z_radius(tz, kz);
  if (lg && !a.tck) {
    set_error("drc_blurz_bwd: the winner-only / fused-loss backward needs the fast ray state");
    return DPC_ERR_ARG;
  }
  if (a.tck) {
    if (g_probs || g_voxels || a.Vz != a.V || !a.logsum) {
      set_error("drc_blurz_bwd: fast ray state needs a cubic grid, log-sum DRC and no g_probs / g_voxels");
      return DPC_ERR_ARG;
    }
    DPC_DISPATCH_V(a.V, DPC_DISPATCH_R(r, launch_bwd_fast<V, R>(a, c, z_taps<R>(tz, kz, r), g_mask,
                                                                g_depth, g_grid, scale_partials,
                                                                zero_ints, n_zero, s,
                                                                lg ? *lg : LossGrad())));
    return check_launch("drc_blurz_bwd_fast");
  }
  if (g_probs || g_voxels) {
    DPC_DISPATCH_V(a.V, DPC_DISPATCH_R(r, launch_bwd_one<V, R, true>(
                                              a, c, z_taps<R>(tz, kz, r), g_mask, g_depth, g_probs,
                                              g_voxels, g_grid, scale_partials, zero_ints, n_zero, s)));
  } else {
    DPC_DISPATCH_V(a.V, DPC_DISPATCH_R(r, launch_bwd_one<V, R, false>(
                                              a, c, z_taps<R>(tz, kz, r), g_mask, g_depth, nullptr,
                                              nullptr, g_grid, scale_partials, zero_ints,
                                              n_zero, s)));
  }
  return check_launch("drc_blurz_bwd");
}

int launch_blur_z(const float *src, float *dst, int P, int Vz, int V, const float *tz, int kz,
                  cudaStream_t s) {
  const int r = z_radius(tz, kz);
  DPC_DISPATCH_V(V, DPC_DISPATCH_R(r, blur_z_kernel<V, R><<<P * (V * V / 2) / kFwdThreads,
                                                          kFwdThreads, 0, s>>>(
                                          src, dst, Vz, z_taps<R>(tz, kz, r))));
  return check_launch("blur_z");
}

int launch_depth_from_probs(const float *probs, float *depth, int P, int Vz, int V,
                            float cam_dist, float max_depth, cudaStream_t s) {
  const int n = P * V * V;
  depth_from_probs_kernel<<<(n + 127) / 128, 128, 0, s>>>(probs, depth, P, Vz, V * V,
                                                          1.0f / (float)Vz, cam_dist - 0.5f,
                                                          max_depth);
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
