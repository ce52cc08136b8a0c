// Roofline denominators measured on the box the bench runs on (bench.py `roofline.fp32_frac`).
//
// The grid kernels of the path (blur X / Y / Z and their adjoints) are bound by the FP32 FMA
// pipe, not by HBM, so their roofline needs the FMA rate this GPU sustains
// for the instruction they issue: fma.rn.f32x2 (SASS FFMA2) with the tap broadcast into both
// halves of a register pair and 8 independent packed accumulators per thread -- the inner product
// of blur_xy.cu / drc.cu without its loads.  The caller times the launch with CUDA events.
#include "common.cuh"

namespace dpc {

__global__ void __launch_bounds__(256)
fma2_probe_kernel(float *__restrict__ out, int iters, float seed) {
  unsigned long long acc[8], w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float a = threadIdx.x * 1e-3f + i, b = seed + i * 0.25f + threadIdx.x;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(a), "f"(a + 0.5f));
    asm("mov.b64 %0, {%1, %2};" : "=l"(w[i]) : "f"(b), "f"(b + 0.125f));
  }
  float kr[21];
#pragma unroll
  for (int t = 0; t < 21; ++t) kr[t] = 1e-3f * (float)(t + 1) + seed;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 21; ++t) {
      unsigned long long k2;
      asm("mov.b64 %0, {%1, %1};" : "=l"(k2) : "f"(kr[t]));
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[j]) : "l"(k2), "l"(w[(j + t) & 7]));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
    s += lo + hi;
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace dpc

extern "C" {

// One launch of `blocks` CTAs x 256 threads x iters x 21 taps x 8 packed FMAs (= 16 scalar FMAs):
// *fma_count_host receives the scalar-FMA count of the launch.  out: blocks * 256 floats.
int dpc_fma_rate_probe(int blocks, int iters, float *out, double *fma_count_host, void *stream) {
  using namespace dpc;
  if (blocks < 1 || iters < 1 || !out) {
    set_error("fma_rate_probe: blocks >= 1, iters >= 1 and an output buffer are required");
    return DPC_ERR_ARG;
  }
  fma2_probe_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.f);
  if (fma_count_host) *fma_count_host = (double)blocks * 256.0 * iters * 21.0 * 16.0;
  return check_launch("fma_rate_probe");
}

}  // extern "C"
