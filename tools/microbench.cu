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

// Do DFMA (FP64 vector pipe) and DMMA (FP64 tensor path) share execution resources?  Even warps issue DMMA,
// odd warps DFMA; if the combined rate exceeds either ceiling the two can overlap inside one SM.
__global__ void __launch_bounds__(256, 2) mixed_kernel(double *out, int iters_dmma, int iters_dfma, double seed) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp & 1) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i * 1e-3 + threadIdx.x * 1e-6;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters_dfma; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
  } else {
    double c0[16], c1[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c0[i] = c1[i] = 0.0;
    double a = seed + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-7;
    for (int it = 0; it < iters_dmma; ++it) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c0[i] + c1[i];
  }
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
  {
    // 16 warps per SM: 8 DMMA warps + 8 DFMA warps; each side sized to ~the same solo duration
    const int threads = 256, blocks = sms * 2;
    const int it_mma = 4096, it_fma = 4096 * 8;   // per warp: 2*256*16*it_mma vs 2*16*32*it_fma flop (equal)
    float ms_mma = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, it_mma, 0, 1.0); }, 5);
    float ms_fma = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, 0, it_fma, 1.0); }, 5);
    float ms_both = time_ms([&] { mixed_kernel<<<blocks, threads>>>(out, it_mma, it_fma, 1.0); }, 5);
    const double fl_mma = 2.0 * 256 * 16 * it_mma * (double)blocks * 4, fl_fma = 2.0 * 16 * 32 * (double)it_fma * blocks * 4;
    cudaError_t me = cudaGetLastError();
    if (me != cudaSuccess) printf(", \"mixed_error\": \"%s\"", cudaGetErrorString(me));
    printf(", \"mixed_dmma_only_tflops\": %.3f, \"mixed_dfma_only_tflops\": %.3f, \"mixed_both_tflops\": %.3f, "
           "\"mixed_ms\": [%.4f, %.4f, %.4f]",
           fl_mma / ms_mma * 1e-9, fl_fma / ms_fma * 1e-9, (fl_mma + fl_fma) / ms_both * 1e-9, ms_mma, ms_fma, ms_both);
  }
  printf("}\n");
  return 0;
}
