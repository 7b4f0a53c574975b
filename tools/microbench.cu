// microbench.cu -- FP64 pipe ceilings on the gpurun B200: DFMA, DMMA.8x8x4, exp/sqrt issue rates.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
// Output: one JSON object on stdout (committed under profiles/ as the FP64 roofline denominators).
#include <cuda_runtime.h>
#include <cstdio>
#include <cmath>

__global__ void dfma_kernel(double *out, int iters, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i * 1e-3 + threadIdx.x * 1e-6;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dmma_kernel(double *out, int iters, double seed) {
  double c0[NACC], c1[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c0[i] = c1[i] = 0.0;
  double a = seed + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-7;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_kernel(double *out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed * 0.1 + i * 1e-3 + threadIdx.x * 1e-6;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = exp(-a[i]) ;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void sqrt_kernel(double *out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + i + threadIdx.x * 1e-3;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = sqrt(a[i]) + 1.5;
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F launch, int reps) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = fminf(best, ms);
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double *out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  const int iters = 4096;
  for (int wps : {4, 8, 16, 32}) {  // warps per SM
    const int threads = 256, blocks = sms * wps * 32 / threads;
    float ms = time_ms([&] { dfma_kernel<<<blocks, threads>>>(out, iters, 1.0); }, 5);
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf(", \"dfma_tflops_w%d\": %.3f", wps, fl / ms * 1e-9);
  }
  for (int wps : {4, 8, 16}) {
    const int threads = 256, blocks = sms * wps * 32 / threads;
    float ms = time_ms([&] { dmma_kernel<8><<<blocks, threads>>>(out, iters, 1.0); }, 5);
    double fl = 2.0 * 256 * 8 * iters * (double)blocks * threads / 32;
    printf(", \"dmma884_tflops_acc8_w%d\": %.3f", wps, fl / ms * 1e-9);
    ms = time_ms([&] { dmma_kernel<32><<<blocks, threads>>>(out, iters / 4, 1.0); }, 5);
    fl = 2.0 * 256 * 32 * (iters / 4) * (double)blocks * threads / 32;
    printf(", \"dmma884_tflops_acc32_w%d\": %.3f", wps, fl / ms * 1e-9);
  }
  {
    const int threads = 256, blocks = sms * 8;
    float ms = time_ms([&] { dmma_kernel<1><<<blocks, threads>>>(out, iters, 1.0); }, 5);
    // dependent chain: latency per DMMA in ns
    printf(", \"dmma884_dep_latency_ns\": %.3f", ms * 1e6 / iters);
  }
  {
    const int threads = 256, blocks = sms * 8;
    float ms = time_ms([&] { exp_kernel<<<blocks, threads>>>(out, 1024, 1.0); }, 5);
    printf(", \"exp_gops\": %.2f", 8.0 * 1024 * blocks * threads / ms * 1e-6);
    ms = time_ms([&] { sqrt_kernel<<<blocks, threads>>>(out, 1024, 1.0); }, 5);
    printf(", \"sqrt_gops\": %.2f", 8.0 * 1024 * blocks * threads / ms * 1e-6);
  }
  printf("}\n");
  return 0;
}
