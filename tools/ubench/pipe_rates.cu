// Per-SM issue rates of the instructions the attention softmax is made of (dev aid): ex2.approx, FMNMX3, F2FP pack, FFMA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float seed) {
  float a[8];
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(a[(i + 2) & 7]));
      if (OP == 3) { uint32_t h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(h); }
      if (OP == 4) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
    }
    if (OP >= 5) {   // the softmax mix per score: fma (scale - max), ex2, add (row sum), [OP 6: + half a cvt.f16x2 and half a max3]
      float b[8];
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(b[i]) : "f"(a[i]), "f"(seed), "f"(-seed));
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b[i]));
#pragma unroll
      for (int i = 0; i < 8; i++) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
      if (OP == 6) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          uint32_t h;
          asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b[i]), "f"(b[i + 1]));
          asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i + 1]), "f"(__uint_as_float(h)));
        }
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run_mix(const char* name, float* d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads = 128; threads <= 1024; threads *= 2) {   // 1, 2, 4, 8 warps per scheduler
    k<OP><<<sms, threads>>>(d, 100, 0.5f);
    cudaEventRecord(e0);
    k<OP><<<sms, threads>>>(d, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * threads * iters * 8;
    printf("%-10s %4d thr/SM (%d warps/scheduler): %.1f ex2/clk/SM at %d MHz nominal\n", name, threads, threads / 128, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  }
}
template <int OP> void run(const char* name, float* d) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks_per_sm = 1; blocks_per_sm <= 2; blocks_per_sm++) {
    k<OP><<<sms * blocks_per_sm, 1024>>>(d, 100, 0.5f);
    cudaEventRecord(e0);
    k<OP><<<sms * blocks_per_sm, 1024>>>(d, iters, 0.5f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * blocks_per_sm * 1024 * iters * 8;
    printf("%-10s %d x 1024 thr/SM: %.1f lane-ops/clk/SM (at %d MHz nominal), %.2f ms\n", name, blocks_per_sm, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, ms);
  }
}
int main() {
  float* d; cudaMalloc(&d, 1 << 26);
  run_mix<0>("ex2 alone", d); run_mix<5>("fma+ex2+add", d); run_mix<6>("softmax mix", d);
  run<0>("ex2", d); run<1>("ffma", d); run<2>("fmnmx3", d); run<3>("f2fp", d); run<4>("fadd", d);
  return 0;
}
