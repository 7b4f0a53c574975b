// common.cuh -- tile-packed FP64 matrix layout ("P-layout") + sm_100a PTX wrappers.
//
// P-layout.  Every matrix the GEMM core touches (K, its Cholesky factor L, the triangular inverse
// W = L^-1, W^T, the cross-covariance K*^T of a candidate chunk) is stored in HBM as
//
//   macro-tile  128 rows x 16 cols  = 2048 doubles = 16 KB, contiguous      -> one TMA bulk copy
//   micro-tile    8 rows x  8 cols  =   64 doubles = 512 B, 16x2 per macro   -> one DMMA fragment pair
//   inside a micro-tile element (r, c) sits at ((r*4 + c%4)*2 + c/4)
//
// so that lane T of a warp reading the 16 bytes at micro-tile + 16*T gets exactly its m8n8k4 FP64
// operand fragment (row T/4, k = T%4) for two consecutive k-steps: conflict-free LDS.128, no
// shuffles, no ldmatrix (which has no 64-bit form).  Macro-tiles of one 128-row block are
// consecutive along the column (k) axis, so the k-range a CTA streams is one contiguous run.
// Rows are padded to a multiple of 128 and columns to a multiple of 16.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace boss {

constexpr int TM = 128;              // macro-tile rows (= CTA tile M = CTA tile N)
constexpr int TK = 16;               // macro-tile cols (= k per pipeline stage)
constexpr int TILE_ELEMS = TM * TK;  // 2048 doubles
constexpr int TILE_BYTES = TILE_ELEMS * 8;
constexpr int KT_PER_BLOCK = TM / TK;  // 8 macro-tiles span one 128-wide block column

__host__ __device__ __forceinline__ size_t p_index(int r, int c, int ktiles) {
  size_t tile = (size_t)(r >> 7) * (size_t)ktiles + (size_t)(c >> 4);
  int rr = r & 127, cc = c & 15;
  int micro = ((rr >> 3) << 1) + (cc >> 3);
  int within = ((((rr & 7) << 2) + (cc & 3)) << 1) + ((cc & 7) >> 2);
  return tile * TILE_ELEMS + (size_t)micro * 64 + within;
}

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA bulk copy (cp.async.bulk -> UBLKCP), FP64 tensor-core MMA (DMMA)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// Re-initialise a barrier that already holds a valid mbarrier object (persistent kernels that reuse one set of
// barriers for rings of different depth): PTX leaves mbarrier.init on a live object undefined -- on B200 the arrival
// count of the old object survived -- so the object is invalidated first.
__device__ __forceinline__ void mbar_reinit(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Truly non-blocking probe (mbarrier.try_wait may suspend the thread up to a system time limit before it reports
// failure -- harmless in the long stages of the 128-wide mainloops, ruinous for a producer that shares its warp with
// consumers of 150-cycle stages).
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a lost TMA transaction traps (launch fails with an error) instead of hanging the GPU.  A failed
// try_wait suspends the thread for a few microseconds, so the bound is ~15 s -- far beyond any legitimate wait.
// -DBOSS_DEBUG_MBAR additionally prints which barrier timed out; it is off by default because the (never taken) call
// inside every wait loop cost the tensor-core mainloops 3 % (register allocation around the call site).
#ifdef BOSS_DEBUG_MBAR
static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {
  printf("boss_b200: mbarrier wait timed out: smem 0x%x parity %u block (%d,%d) thread %d\n", bar, parity, blockIdx.x,
         blockIdx.y, threadIdx.x);
  __trap();
}
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) {
#ifdef BOSS_DEBUG_MBAR
      mbar_timeout(bar, parity);
#else
      __trap();
#endif
    }
  }
}
// Spinning wait on test_wait (no suspension): for short pipeline stages, where the wake-up latency of a suspended
// try_wait is of the order of the stage itself.  Same bound as mbar_wait.
__device__ __forceinline__ void mbar_spin_wait(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++spins > (1u << 28)) __trap();
  }
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src_gmem), "r"(bytes), "r"(bar)
               : "memory");
}
// D(8x8) += A(8x4, row) * B(4x8, col), FP64 tensor core (SASS: DMMA.8x8x4).
// lane T holds A[T/4][T%4], B[T%4][T/4], C[T/4][2*(T%4) + {0,1}].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double2 lds128(const double *p) { return *reinterpret_cast<const double2 *>(p); }

// total order used by Julia's argmax (isless): NaN is maximal, -0.0 < +0.0; ties -> lowest index.
__device__ __forceinline__ bool acq_better(double va, long long ia, double vb, long long ib) {
  bool na = va != va, nb = vb != vb;
  if (na || nb) {
    if (na && nb) return ia < ib;
    return na;
  }
  if (va > vb) return true;
  if (va < vb) return false;
  // equal (incl. +-0): distinguish signed zeros, then index
  long long ba = __double_as_longlong(va), bb = __double_as_longlong(vb);
  if (ba != bb) return ba >= 0 && bb < 0;  // only possible for +0 vs -0
  return ia < ib;
}
#endif  // __CUDACC__

}  // namespace boss
