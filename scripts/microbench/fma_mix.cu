// Does packed fma.rn.f32x2 free issue slots for the non-FMA work of a mixed
// instruction stream?  Per inner step: 16 FMA (scalar) or 8 FFMA2 (packed) +
// 8 FMNMX/FADD "other" instructions.
#include <cstdio>
#include <cuda_runtime.h>
struct Taps { float k[21]; };

template <int MODE>
__global__ void __launch_bounds__(256) mix_kernel(float *out, Taps taps, int iters, float seed) {
  float acc[16], w[16], o[8];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x * 1e-3f + i; w[i] = seed + i * 0.25f + threadIdx.x; }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = seed + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < 21; ++t) {
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(taps.k[t], w[(j + t) & 15], acc[j]);
      } else {
        unsigned long long *a2 = reinterpret_cast<unsigned long long *>(acc);
        unsigned long long *w2 = reinterpret_cast<unsigned long long *>(w);
        unsigned long long k2;
        asm("mov.b64 %0, {%1, %1};" : "=l"(k2) : "f"(taps.k[t]));
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[j]) : "l"(w2[(j + t) & 7]), "l"(k2));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fminf(fmaxf(o[i] + 0.5f, w[i]), 7.f + o[(i + 1) & 7]);   // 3 ops each
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += o[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, float *out) {
  Taps t; for (int i = 0; i < 21; ++i) t.k[i] = 1e-3f * (i + 1);
  const int iters = 1000, blocks = 148 * 8;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  mix_kernel<MODE><<<blocks, 256>>>(out, t, 10, 0.f); cudaDeviceSynchronize();
  cudaEventRecord(e0); mix_kernel<MODE><<<blocks, 256>>>(out, t, iters, 0.f); cudaEventRecord(e1);
  cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * 256 * iters * 21 * 16;
  printf("%-34s %.3f ms  %.2f TFMA/s\n", name, ms, fma / ms * 1e-9);
}
int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  run<0>("scalar FFMA + 24 other/16 FMA", out);
  run<1>("packed FFMA2 + 24 other/16 FMA", out);
  return 0;
}
