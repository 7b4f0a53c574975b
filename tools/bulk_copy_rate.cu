// bulk_copy_rate.cu -- how fast can ONE CTA per SM stream global -> shared with 1-D TMA bulk copies (cp.async.bulk) of a
// given size through an R-slot mbarrier ring?  (Is a short-stage pipeline bound by the copy engine?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I boss.jl_b200/csrc -o tools/bulk_copy_rate.bin tools/bulk_copy_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include "common.cuh"
using namespace boss;

// thread 0 issues `n` copies of `bytes` each (2 per "stage" if pair), consumers (all 8 warps) wait for each and release it
__global__ void __launch_bounds__(384, 1) ring_kernel(const double *src, size_t stride_elems, int n, int bytes, int R, int per_stage, double *sink, int poll_mode, int NP, int wrap) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t *full = reinterpret_cast<uint64_t *>(smem), *empty = full + 64;
  unsigned char *ring = smem + 1024;
  const int tid = threadIdx.x, lane = tid & 31;
  if (tid < R) { mbar_init(smem_u32(&full[tid]), 1); mbar_init(smem_u32(&empty[tid]), 8); mbar_fence_init(); }
  __syncthreads();
  const double *base = src + (size_t)blockIdx.x * stride_elems;
  const int wid = tid >> 5;
  const bool dedicated_spin = NP <= 0 && poll_mode == 3;
  int issued = wid;   // next stage this producer issues (producers: lane 0 of warps 0..NP-1, stages round-robin)
  auto issue = [&](bool blocking) -> bool {
    const int slot = issued % R;
    if (issued >= R) {
      const uint32_t eb = smem_u32(&empty[slot]), par = (uint32_t)((issued / R - 1) & 1);
      if (blocking) mbar_wait(eb, par); else if (dedicated_spin) mbar_spin_wait(eb, par); else if (!mbar_test_wait(eb, par)) return false;
    }
    const uint32_t bar = smem_u32(&full[slot]);
    mbar_arrive_expect_tx(bar, (uint32_t)(bytes * per_stage));
    for (int q = 0; q < per_stage; ++q)
      bulk_g2s(smem_u32(ring + (size_t)slot * bytes * per_stage + (size_t)q * bytes), base + ((size_t)(issued % wrap) * per_stage + q) * (bytes / 8), bytes, bar);
    issued += NP;
    return true;
  };
  const bool dedicated = NP <= 0;
  if (dedicated) {
    const int nd = NP == 0 ? 1 : -NP;      // dedicated producer warps 8 .. 8 + nd - 1, stages round-robin
    NP = nd;
    if (wid >= 8) {
      if (lane == 0 && wid - 8 < nd) { issued = wid - 8; while (issued < n) issue(poll_mode != 3); if (false) return; }
      if (lane == 0 && wid - 8 < nd && poll_mode == 3) {}
      return;
    }
    issued = n;                // consumers never issue
  }
  const bool producer = !dedicated && lane == 0 && wid < NP;
  if (producer) while (issued < n && issued < R) issue(false);
  double acc = 0;
  for (int g = 0; g < n; ++g) {
    if (producer) while (issued < n && issued < g + R) { if (!issue(issued <= g)) break; }
    const int slot = g % R;
    if (poll_mode == 0) mbar_spin_wait(smem_u32(&full[slot]), (uint32_t)((g / R) & 1));
    else if (poll_mode == 1) mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / R) & 1));
    else if (poll_mode == 4) { if (lane == 0) mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / R) & 1)); __syncwarp(); }
    else { if (lane == 0) mbar_spin_wait(smem_u32(&full[slot]), (uint32_t)((g / R) & 1)); __syncwarp(); }
    acc += reinterpret_cast<const double *>(ring + (size_t)slot * bytes * per_stage)[tid % (bytes / 8)];
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
  }
  if (acc == 123.456) sink[0] = acc;
}

int main() {
  const size_t per_cta = 4 << 20;   // 4 MB per CTA
  double *src, *sink;
  cudaMalloc(&src, per_cta * 148); cudaMemset(src, 0, per_cta * 148); cudaMalloc(&sink, 8);
  cudaFuncSetAttribute(ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  printf("[");
  bool first = true;
  for (int resident = 0; resident < 2; ++resident)
    for (int ctas : {8, 148})
      for (int per_stage : {2, 8})
        for (int NP : {-2}) {
          const int bytes = 4096, R = 4, poll_mode = 1;
          // resident: every CTA loops over its own 512 KB window (148 x 512 KB = 74 MB: stays in the 126 MB L2)
          const size_t window = resident ? (512 << 10) : per_cta;
          const int wrap = (int)(window / ((size_t)bytes * per_stage));
          const int n = (int)(per_cta * 4 / ((size_t)bytes * per_stage));   // 16 MB per CTA
          float best = 1e9f;
          for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            ring_kernel<<<ctas, 384, 1024 + (size_t)bytes * per_stage * R>>>(src, per_cta / 8, n, bytes, R, per_stage, sink, poll_mode, NP, wrap);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
          }
          const double ns_per_stage = best * 1e6 / n;
          printf("%s\n {\"ctas\": %d, \"l2_resident\": %d, \"copy_bytes\": %d, \"copies_per_stage\": %d, \"ring_slots\": %d, \"stages\": %d, \"ns_per_stage\": %.1f, \"GBps_per_sm\": %.1f, \"TBps_total\": %.2f}",
                 first ? "" : ",", ctas, resident, bytes, per_stage, R, n, ns_per_stage, bytes * per_stage / ns_per_stage, ctas * (double)bytes * per_stage / ns_per_stage * 1e-3);
          first = false;
        }
  printf("\n]\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
