// Microbenchmark: fp32 FMA issue rate on B200 for the operand forms the blur
// kernels can use.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a fma_rate.cu -o fma_rate
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float k[21]; };

template <int MODE>
__global__ void __launch_bounds__(256) fma_kernel(float *out, Taps taps, int iters, float seed) {
  float acc[16];
  float w[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { acc[i] = threadIdx.x * 1e-3f + i; w[i] = seed + i * 0.25f + threadIdx.x; }
  float kr[21];
#pragma unroll
  for (int t = 0; t < 21; ++t) kr[t] = taps.k[t] + seed;   // register copies (not foldable)
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {          // 3-register FFMA, taps in registers
#pragma unroll
      for (int t = 0; t < 21; ++t)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(kr[t], w[(j + t) & 15], acc[j]);
    } else if (MODE == 1) {   // constant-bank tap operand
#pragma unroll
      for (int t = 0; t < 21; ++t)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = fmaf(taps.k[t], w[(j + t) & 15], acc[j]);
    } else {                  // packed fma.rn.f32x2: 8 packed accumulators
      unsigned long long *a2 = reinterpret_cast<unsigned long long *>(acc);
      unsigned long long *w2 = reinterpret_cast<unsigned long long *>(w);
#pragma unroll
      for (int t = 0; t < 21; ++t) {
        unsigned long long k2;
        asm("mov.b64 %0, {%1, %1};" : "=l"(k2) : "f"(kr[t]));
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(a2[j]) : "l"(k2), "l"(w2[(j + t) & 7]));
      }
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, float *out, int blocks) {
  Taps t;
  for (int i = 0; i < 21; ++i) t.k[i] = 1e-3f * (i + 1);
  const int iters = 2000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  fma_kernel<MODE><<<blocks, 256>>>(out, t, 10, 0.f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  fma_kernel<MODE><<<blocks, 256>>>(out, t, iters, 0.f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double fma = (double)blocks * 256 * iters * 21 * 16;
  printf("%-28s blocks=%d  %.3f ms  %.2f TFMA/s (%.1f TFLOP/s)\n", name, blocks, ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
}

int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 256 * sizeof(float) * 4);
  for (int occ : {1, 2, 4, 8}) {
    int blocks = 148 * occ;
    run<0>("ffma 3-reg", out, blocks);
    run<1>("ffma const-bank tap", out, blocks);
    run<2>("fma.f32x2 packed", out, blocks);
  }
  return 0;
}
