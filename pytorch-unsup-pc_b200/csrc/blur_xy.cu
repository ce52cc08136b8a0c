// X and Y passes of the separable Gaussian blur, fused per Z-plane.
//
// Reference: point_cloud_to.py:90-103 smoothen_voxels3d -- three fp64
// F.conv3d calls with kernels [1,1,1,1,K], [1,1,1,K,1], [1,1,K,1,1] and zero
// 'same' padding (the X and Y ones are done here; Z is fused into the DRC ray
// march, drc.cu), plus the clamp(raw,0,1) that precedes them (:198-201).
//
// One CTA owns one V x V plane of one projection, so neither pass needs a
// halo: the zero padding lives in shared memory.  The plane is staged once
// (coalesced float4 loads, clamp + raw<=1 bit mask fused on the way in), the
// X pass writes its result transposed into a second shared tile, the Y pass
// reads that tile and stores coalesced rows.  Because the whole plane is in
// shared memory before the first store, src == dst is safe: the forward blurs
// the occupancy grid in place and only a 1-bit-per-voxel mask survives for the
// backward's clamp gate.
//
// Inner loop: each thread produces 16 consecutive outputs from a 16+2R window
// read with conflict-free LDS.128 (row stride = 4*odd floats); 16 independent
// accumulators, taps as constant-bank FFMA operands: 336 FFMA per 9 LDS.128.
// The backward (adjoint) is the same kernel: symmetric taps + zero padding
// make each pass self-adjoint and the passes commute; the clamp gate is
// applied on the final store.
#include "common.cuh"

namespace dpc {

template <int V, int R>
struct XYCfg {
  static constexpr int J = 16;                    // outputs per thread per pass
  static constexpr int W = J + 2 * R;             // window length
  static constexpr int W4 = (W + 3) / 4;          // LDS.128 per window
  static constexpr int RH = V < 64 ? V : 64;      // plane rows resident in tile A
  static constexpr int S0 = V - J + 4 * W4;       // floats a row must hold
  static constexpr int S = ((S0 / 4) % 2 == 1) ? S0 : S0 + 4;  // stride: 4*odd => no conflicts
  static constexpr int TASKS2 = V * V / J;
  static constexpr int THREADS = TASKS2 < 512 ? TASKS2 : 512;
  static constexpr int MINB = (V == 128) ? 2 : (V == 64 ? 4 : 8);
  static constexpr size_t SMEM = (size_t)(RH + V) * S * sizeof(float);
  static_assert(V % 32 == 0, "V must be a multiple of 32");
  static_assert((RH * V / 4) % THREADS == 0, "fill loop must be warp-uniform");
};

template <int R, int J, int W4>
__device__ __forceinline__ void window_fma(const float *__restrict__ win_base,
                                           const Taps<R> &taps, float (&acc)[J]) {
#pragma unroll
  for (int j = 0; j < J; ++j) acc[j] = 0.f;
#pragma unroll
  for (int i = 0; i < W4; ++i) {
    const float4 v4 = *reinterpret_cast<const float4 *>(win_base + 4 * i);
    const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int t = 4 * i + c - j;  // compile-time after unrolling
        if (t >= 0 && t <= 2 * R) acc[j] = fmaf(taps.k[t], vv[c], acc[j]);
      }
    }
  }
}

template <int V, int R, bool CLAMP_IN, bool WRITE_BITS, bool MASK_OUT>
__global__ void __launch_bounds__(XYCfg<V, R>::THREADS, XYCfg<V, R>::MINB)
blur_xy_kernel(const float *__restrict__ src, float *__restrict__ dst,
               uint32_t *__restrict__ bits_out, const uint32_t *__restrict__ bits_in,
               const Taps<R> kx, const Taps<R> ky) {
  using C = XYCfg<V, R>;
  extern __shared__ __align__(16) float smem[];
  float *A = smem;                 // [RH][S]  input rows, x padded by R zeros
  float *B = smem + C::RH * C::S;  // [V][S]   X-blurred, transposed: B[x][R + y]
  const int tid = threadIdx.x;
  const size_t plane = blockIdx.x;
  const float *sp = src + plane * V * V;

  // zero both tiles once (the pads stay zero for the whole kernel)
  for (int i = tid; i < (C::RH + V) * C::S / 4; i += C::THREADS)
    reinterpret_cast<float4 *>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

#pragma unroll 1
  for (int h = 0; h < V / C::RH; ++h) {
    // ---- stage rows [h*RH, (h+1)*RH) ----
    for (int i = tid; i < C::RH * V / 4; i += C::THREADS) {
      const int row = i / (V / 4), c4 = i % (V / 4);
      const int gy = h * C::RH + row;
      float4 v = __ldg(reinterpret_cast<const float4 *>(sp + gy * V) + c4);
      if (WRITE_BITS) {
        uint32_t nib = (v.x <= 1.f ? 1u : 0u) | (v.y <= 1.f ? 2u : 0u) | (v.z <= 1.f ? 4u : 0u) |
                       (v.w <= 1.f ? 8u : 0u);
        nib <<= 4 * (tid & 7);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 1);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 2);
        nib |= __shfl_xor_sync(0xffffffffu, nib, 4);
        if ((tid & 7) == 0) bits_out[plane * (V * V / 32) + (gy * V + 4 * c4) / 32] = nib;
      }
      if (CLAMP_IN) {
        v.x = fminf(fmaxf(v.x, 0.f), 1.f);
        v.y = fminf(fmaxf(v.y, 0.f), 1.f);
        v.z = fminf(fmaxf(v.z, 0.f), 1.f);
        v.w = fminf(fmaxf(v.w, 0.f), 1.f);
      }
      float *a = A + row * C::S + R + 4 * c4;
      if (R % 2 == 0) {
        *reinterpret_cast<float2 *>(a) = make_float2(v.x, v.y);
        *reinterpret_cast<float2 *>(a + 2) = make_float2(v.z, v.w);
      } else {
        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
      }
    }
    __syncthreads();
    // ---- X pass: lanes <-> 32 consecutive rows, one 16-wide x block ----
    for (int task = tid; task < C::RH * V / C::J; task += C::THREADS) {
      const int yl = task % C::RH, x0 = (task / C::RH) * C::J;
      float acc[C::J];
      window_fma<R, C::J, C::W4>(A + yl * C::S + x0, kx, acc);
      float *b = B + x0 * C::S + R + h * C::RH + yl;
#pragma unroll
      for (int j = 0; j < C::J; ++j) b[j * C::S] = acc[j];
    }
    __syncthreads();
  }
  // ---- Y pass: lanes <-> 32 consecutive x, one 16-tall y block ----
  float *dp = dst + plane * V * V;
  for (int task = tid; task < C::TASKS2; task += C::THREADS) {
    const int x = task % V, y0 = (task / V) * C::J;
    float acc[C::J];
    window_fma<R, C::J, C::W4>(B + x * C::S + y0, ky, acc);
#pragma unroll
    for (int j = 0; j < C::J; ++j) {
      float o = acc[j];
      if (MASK_OUT) {
        const uint32_t wbits = __ldg(bits_in + plane * (V * V / 32) + ((y0 + j) * V + x) / 32);
        o = ((wbits >> (x & 31)) & 1u) ? o : 0.f;
      }
      dp[(y0 + j) * V + x] = o;
    }
  }
}

template <int V, int R>
static int launch_vr(const BlurXYArgs &a, const float *tx, int kx, const float *ty, int ky,
                     cudaStream_t s) {
  using C = XYCfg<V, R>;
  const Taps<R> KX = make_taps<R>(tx, kx), KY = make_taps<R>(ty, ky);
  dim3 g(a.planes), t(C::THREADS);
#define DPC_LAUNCH_XY(CL, WB, MO)                                                              \
  do {                                                                                         \
    static bool attr_done = false;                                                             \
    if (!attr_done) {                                                                          \
      cudaFuncSetAttribute(blur_xy_kernel<V, R, CL, WB, MO>,                                   \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM);         \
      attr_done = true;                                                                        \
    }                                                                                          \
    blur_xy_kernel<V, R, CL, WB, MO><<<g, t, C::SMEM, s>>>(a.src, a.dst, a.bits_out, a.bits_in, \
                                                           KX, KY);                            \
  } while (0)
  if (a.bits_in) {
    DPC_LAUNCH_XY(false, false, true);
  } else if (a.bits_out) {
    if (!a.clamp_in) { set_error("blur_xy: bits_out requires clamp_in"); return DPC_ERR_ARG; }
    DPC_LAUNCH_XY(true, true, false);
  } else if (a.clamp_in) {
    DPC_LAUNCH_XY(true, false, false);
  } else {
    DPC_LAUNCH_XY(false, false, false);
  }
#undef DPC_LAUNCH_XY
  return check_launch("blur_xy");
}

template <int V>
static int launch_v(const BlurXYArgs &a, const float *tx, int kx, const float *ty, int ky,
                    cudaStream_t s) {
  int r = effective_radius(tx, kx), r2 = effective_radius(ty, ky);
  if (r2 > r) r = r2;
  // drop outer taps that are exactly zero: recentre on the shorter tap set
  const int ox = kx / 2 - r, oy = ky / 2 - r;
  const float one = 1.f;
  const float *px = kx > 0 ? tx + (ox > 0 ? ox : 0) : &one;
  const float *py = ky > 0 ? ty + (oy > 0 ? oy : 0) : &one;
  const int nx = kx > 0 ? (ox > 0 ? 2 * r + 1 : kx) : 1, ny = ky > 0 ? (oy > 0 ? 2 * r + 1 : ky) : 1;
  if (r <= 5) return launch_vr<V, 5>(a, px, nx, py, ny, s);
  if (r <= 10) return launch_vr<V, 10>(a, px, nx, py, ny, s);
  set_error("blur_xy: tap radius %d > 10 unsupported", r);
  return DPC_ERR_ARG;
}

int launch_blur_xy(const BlurXYArgs &a, const float *tx, int kx, const float *ty, int ky,
                   cudaStream_t s) {
  switch (a.V) {
    case 32: return launch_v<32>(a, tx, kx, ty, ky, s);
    case 64: return launch_v<64>(a, tx, kx, ty, ky, s);
    case 128: return launch_v<128>(a, tx, kx, ty, ky, s);
  }
  set_error("blur_xy: vox_size %d unsupported (32, 64, 128)", a.V);
  return DPC_ERR_ARG;
}

}  // namespace dpc
