// dmma_order_test.cu -- (1) in which order does DMMA.8x8x4 accumulate its four products?  (2) dependent-issue latency of
// DFMA vs DMMA.  Decides whether a DFMA kernel can reproduce the tensor-core kernels' bits for tiny candidate batches.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_order_test.bin tools/dmma_order_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>

__global__ void dmma_once(const double *A, const double *B, const double *Cin, double *Cout) {
  // lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2*(T%4) + {0,1}]
  const int T = threadIdx.x;
  double a = A[(T / 4) * 4 + (T % 4)], b = B[(T % 4) * 8 + (T / 4)];
  double c0 = Cin[(T / 4) * 8 + 2 * (T % 4)], c1 = Cin[(T / 4) * 8 + 2 * (T % 4) + 1];
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  Cout[(T / 4) * 8 + 2 * (T % 4)] = c0;
  Cout[(T / 4) * 8 + 2 * (T % 4) + 1] = c1;
}

__global__ void dfma_chain(double *out, int iters, double seed) {
  double a = seed + threadIdx.x * 1e-6;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it) a = fma(a, b, c);
  out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void dmma_chain(double *out, int iters, double seed) {
  double c0 = 0, c1 = 0, a = seed + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-7;
  for (int it = 0; it < iters; ++it)
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1;
}

int main() {
  double hA[32], hB[32], hC[64], hD[64];
  double *dA, *dB, *dC, *dD;
  cudaMalloc(&dA, 256); cudaMalloc(&dB, 256); cudaMalloc(&dC, 512); cudaMalloc(&dD, 512);
  srand(12345);
  long n_asc = 0, n_desc = 0, n_pair = 0, n_pair_c_last = 0, n_exactish = 0, total = 0;
  for (int trial = 0; trial < 2000; ++trial) {
    for (int i = 0; i < 32; ++i) { hA[i] = (rand() / (double)RAND_MAX - 0.5) * pow(2.0, rand() % 20 - 10); hB[i] = (rand() / (double)RAND_MAX - 0.5) * pow(2.0, rand() % 20 - 10); }
    for (int i = 0; i < 64; ++i) hC[i] = (rand() / (double)RAND_MAX - 0.5) * pow(2.0, rand() % 20 - 10);
    cudaMemcpy(dA, hA, 256, cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, 256, cudaMemcpyHostToDevice); cudaMemcpy(dC, hC, 512, cudaMemcpyHostToDevice);
    dmma_once<<<1, 32>>>(dA, dB, dC, dD);
    cudaMemcpy(hD, dD, 512, cudaMemcpyDeviceToHost);
    for (int r = 0; r < 8; ++r)
      for (int c = 0; c < 8; ++c) {
        const double *a = hA + r * 4; double b[4] = {hB[0 * 8 + c], hB[1 * 8 + c], hB[2 * 8 + c], hB[3 * 8 + c]};
        const double c0 = hC[r * 8 + c], got = hD[r * 8 + c];
        double asc = fma(a[3], b[3], fma(a[2], b[2], fma(a[1], b[1], fma(a[0], b[0], c0))));
        double desc = fma(a[0], b[0], fma(a[1], b[1], fma(a[2], b[2], fma(a[3], b[3], c0))));
        double p01 = fma(a[1], b[1], a[0] * b[0]), p23 = fma(a[3], b[3], a[2] * b[2]);
        double pair = c0 + (p01 + p23);
        double pair2 = (c0 + p01) + p23;
        long double ex = (long double)c0 + (long double)a[0] * b[0] + (long double)a[1] * b[1] + (long double)a[2] * b[2] + (long double)a[3] * b[3];
        ++total;
        n_asc += got == asc; n_desc += got == desc; n_pair += got == pair; n_pair_c_last += got == pair2; n_exactish += got == (double)ex;
      }
  }
  printf("{\"dmma884_outputs\": %ld, \"equal_fma_chain_k_ascending\": %ld, \"equal_fma_chain_k_descending\": %ld, \"equal_pairwise\": %ld, "
         "\"equal_pairwise_c_first\": %ld, \"equal_long_double_sum_rounded\": %ld", total, n_asc, n_desc, n_pair, n_pair_c_last, n_exactish);
  // latencies: one warp per SM sub-partition (128 threads per block, one block per SM)
  double *out; cudaMalloc(&out, 148 * 128 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 20;
  float ms;
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dfma_chain<<<148, 128>>>(out, iters, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf(", \"dfma_dependent_latency_ns\": %.3f", ms * 1e6 / iters);
  for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0); dmma_chain<<<148, 128>>>(out, iters / 8, 1.0); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf(", \"dmma884_dependent_latency_ns\": %.3f}\n", ms * 1e6 / (iters / 8));
  return 0;
}
