// gemm_core.cuh -- the one FP64 tensor-core mainloop every O(n^3)/O(n^2 M) stage runs on.
//
//   C(128x128) += A(128 x K) * B(128 x K)^T          ("NT": both operands K-major, P-layout)
//
// * operands arrive by 1-D TMA bulk copies (cp.async.bulk, one 16 KB macro-tile per operand per
//   stage) into a STAGES-deep shared-memory ring, completion on mbarriers;
// * 8 warps, warp tile 64x32 = 8x4 DMMA.8x8x4 fragments, accumulators in registers (FP64 has no
//   tcgen05/TMEM path on sm_100a; DMMA.8x8x4 is the native FP64 tensor instruction);
// * fragments are read with conflict-free LDS.128 straight from the P-layout (see common.cuh);
// * a CTA walks a *sequence* of output tiles (iterator `It`): the ring keeps streaming across tile
//   boundaries, the epilogue functor runs when a tile's last k-stage has been consumed.
//
// Users: Cholesky trailing update / panel solve, triangular inverse (cholesky.cuh) and the fused
// posterior-variance kernel (score.cuh).
#pragma once
#include "common.cuh"
#include <type_traits>

namespace boss {

constexpr int GEMM_THREADS = 256;
constexpr int GEMM_STAGES = 5;
constexpr int GEMM_STAGE_ELEMS = 2 * TILE_ELEMS;                  // A tile + B tile
constexpr int GEMM_STAGE_BYTES = GEMM_STAGE_ELEMS * 8;            // 32 KB
constexpr int GEMM_RING_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES;   // 160 KB
constexpr int GEMM_SCRATCH_BYTES = 4096;                          // epilogue scratch (after the ring)
constexpr int GEMM_SMEM_BYTES = GEMM_RING_BYTES + GEMM_SCRATCH_BYTES + 128;  // + mbarriers (full[5], empty[5])

// Accumulator fragment coordinates of this lane inside the 128x128 CTA tile:
//   row(fm)    = 64*wm + 8*fm + (lane>>2)
//   col(fn, e) = 32*wn + 8*fn + 2*(lane&3) + e
//   (il: interleaved row slabs, rslab(fm) = 2*fm + wm -- the triangular-operand kernels, see has_tri_stages)
struct FragCoord {
  int wm, wn, lane;
  bool il = false;
  __device__ __forceinline__ int rslab(int fm) const { return il ? 2 * fm + wm : 8 * wm + fm; }
  __device__ __forceinline__ int row(int fm) const { return 8 * rslab(fm) + (lane >> 2); }
  __device__ __forceinline__ int col(int fn, int e) const { return 32 * wn + 8 * fn + 2 * (lane & 3) + e; }
};

// Offset (in doubles) of local element (r, c) of a 128x128 block inside the P-layout run that
// starts at the block's first macro-tile (the 8 macro-tiles of a block are consecutive).
__host__ __device__ __forceinline__ int block_offset(int r, int c) {
  return (c >> 4) * TILE_ELEMS + ((((r >> 3) << 1) + ((c & 15) >> 3)) << 6) + ((((r & 7) << 2) + (c & 3)) << 1) +
         ((c & 7) >> 2);
}

// 16-byte P-layout store of one C fragment.  A lane holds (r, c) and (r, c + 1) of an 8x8 micro-tile, which sit
// 16 bytes apart in the P-layout; after swapping one value with lane ^ 2 it holds (r, k) and (r, k + 4), which are
// adjacent.  Halves the store instructions and, in shared memory, is bank-conflict free (the 8-byte scatter is not).
//   base: first macro-tile of the 128x128 block (shared or global); rslab / cslab: 8-row / 8-column slab of the tile
__device__ __forceinline__ void p_store_cfrag(double *base, int rslab, int cslab, int lane, double v0, double v1) {
  const int q = lane & 3;
  const double recv = __shfl_xor_sync(0xffffffffu, q < 2 ? v1 : v0, 2);
  const double2 pair = q < 2 ? make_double2(v0, recv) : make_double2(recv, v1);
  const int k3 = q < 2 ? 2 * q : 2 * q - 3;   // column (mod 4) of the pair's first element
  double *dst = base + (cslab >> 1) * TILE_ELEMS + (((rslab << 1) + (cslab & 1)) << 6) + ((((lane >> 2) << 2) + k3) << 1);
  *reinterpret_cast<double2 *>(dst) = pair;
}

// Transposed variant: element (r, c) of the fragment is stored at (c, r).  Rows r and r + 4 of one column are
// adjacent in the destination, so the exchange partner is lane ^ 16.
__device__ __forceinline__ void p_store_cfrag_t(double *base, int rslab, int cslab, int lane, double v0, double v1) {
  const int rr = lane >> 2, q = lane & 3;
  const bool hi = rr & 4;
  const double recv = __shfl_xor_sync(0xffffffffu, hi ? v0 : v1, 16);
  const double2 pair = hi ? make_double2(recv, v1) : make_double2(v0, recv);
  const int crow = hi ? 2 * q + 1 : 2 * q;   // destination row inside the micro-tile
  double *dst = base + (rslab >> 1) * TILE_ELEMS + (((cslab << 1) + (rslab & 1)) << 6) + (((crow << 2) + (rr & 3)) << 1);
  *reinterpret_cast<double2 *>(dst) = pair;
}

// The inverse: the lane's 16-byte pair of the P-layout (one load), then the lane-pair shuffle back to a C fragment.
__device__ __forceinline__ double2 p_load_cfrag_raw(const double *base, int rslab, int cslab, int lane) {
  const int q = lane & 3;
  const int k3 = q < 2 ? 2 * q : 2 * q - 3;
  return *reinterpret_cast<const double2 *>(base + (cslab >> 1) * TILE_ELEMS + (((rslab << 1) + (cslab & 1)) << 6) +
                                            ((((lane >> 2) << 2) + k3) << 1));
}
__device__ __forceinline__ void p_unpair_cfrag(double2 pair, int lane, double &v0, double &v1) {
  const int q = lane & 3;   // q < 2 holds (c, c + 4) with c = 2q: keeps .x, needs c + 1 = partner's .x; q >= 2 vice versa
  const double recv = __shfl_xor_sync(0xffffffffu, q < 2 ? pair.y : pair.x, 2);
  v0 = q < 2 ? pair.x : recv;
  v1 = q < 2 ? recv : pair.y;
}

// One k micro-step (8 wide) of a warp tile restricted to the live fragment rows LO..HI (triangular stages).
template <int LO, int HI, int STRIDE>
__device__ __forceinline__ void tri_micro_step(double (&acc)[8][4][2], const double *Am, const double2 (&b)[4]) {
  double2 a[8];
#pragma unroll
  for (int fm = LO; fm <= HI; ++fm) a[fm] = lds128(Am + fm * STRIDE);
#pragma unroll
  for (int fm = LO; fm <= HI; ++fm)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a[fm].x, b[fn].x);
#pragma unroll
  for (int fm = LO; fm <= HI; ++fm)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a[fm].y, b[fn].y);
}

// It must provide:  bool valid(); const double* A(); const double* B(); bool tile_end(); int tile(); void next();
//
// Synchronisation: per stage a `full` mbarrier (TMA transaction bytes) and an `empty` mbarrier (one arrival per
// consumer warp).  There is no CTA-wide barrier in the mainloop: a warp that finishes a stage releases the slot
// and moves on, so the eight warps drift apart by up to the ring depth and fill each other's bubbles.  Thread 0
// doubles as the producer: before consuming stage g it issues every later stage whose slot has already been
// released (non-blocking probe) and, if stage g itself has not been issued yet, waits for its slot (the other
// warps can always finish the stages already in flight, so this cannot deadlock).
// SINGLE: the iterator yields exactly one output tile; the epilogue then runs after the loop, so none of the
// pipeline state stays live across it (for epilogues that need the registers, e.g. chol_panel_kernel's phase 2).
// Optional iterator feature ("tail stages"): when It has is_tail() / tail_index(), the last 8 stages carry the 16 KB
// macro-tiles m = 0..7 of a tile C0 (A slot only) instead of operands, and the accumulators become C0 - A B^T as
// those stages are consumed: the tile arrives through the ring that is running anyway, with no exposed latency at
// either end of the kernel (chol_panel_kernel; clock64 showed 6-9 k cycles of accumulator-initialisation loads
// ahead of the first DMMA when the tile was read directly).
template <class T, class = void>
struct has_tail_stages : std::false_type {};
template <class T>
struct has_tail_stages<T, std::void_t<decltype(&T::is_tail)>> : std::true_type {};

// Optional iterator feature ("triangular stages"): when It has kTriMode / tri_diag() / tri_g(), the stages of a tile
// whose A operand is a DIAGONAL block of a triangular matrix (tri_diag()) skip the structurally zero fragments.
// kTriMode = 1: the block is lower triangular (row slab R is non-zero for k micro-steps kk <= R), 2: upper triangular
// (kk >= R); tri_g() = k-tile 0..7 inside the diagonal block.  Such iterators also switch the warp's row-slab ownership from
// blocked (8 wm + fm) to interleaved (2 fm + wm), which deals the live fragments of a triangular block evenly to
// the two warp rows: 72 + 64 of 256 (fm, kk) pairs instead of 100 + 36, so the skipped work (47 % of the block)
// actually comes off the critical path.  5.2 % fewer DMMAs in score_trmm_kernel at n = 2048.
template <class T, class = void>
struct has_tri_stages : std::false_type {};
template <class T>
struct has_tri_stages<T, std::void_t<decltype(&T::tri_diag)>> : std::true_type {};

// after_prologue: run by every thread once the first ring stages have been requested (further prefetches belong
// here: whatever is requested before stage 0 delays the first DMMA).
struct NoPrologueHook {
  __device__ __forceinline__ void operator()() const {}
};
template <bool SINGLE = false, class It, class Epi, class Pre = NoPrologueHook>
__device__ __forceinline__ void gemm_pipeline(It issue_it, It cons_it, Epi &&epi, Pre &&after_prologue = Pre()) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *ring = reinterpret_cast<double *>(smem_raw);
  uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + GEMM_RING_BYTES + GEMM_SCRATCH_BYTES);
  uint64_t *empty = full + GEMM_STAGES;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;

  if (tid < 2 * GEMM_STAGES) {   // one barrier per thread: the CTA waits on this before its first stage
    if (tid < GEMM_STAGES)
      mbar_init(smem_u32(&full[tid]), 1);
    else
      mbar_init(smem_u32(&empty[tid - GEMM_STAGES]), GEMM_THREADS / 32);
    mbar_fence_init();
  }
  __syncthreads();

  int issued = 0;
  // issue stage `issued` into its slot; returns false when the slot is still being read and !blocking
  auto try_issue = [&](bool blocking) -> bool {
    const int slot = issued % GEMM_STAGES;
    if (issued >= GEMM_STAGES) {
      const uint32_t eb = smem_u32(&empty[slot]);
      const uint32_t par = (uint32_t)((issued / GEMM_STAGES - 1) & 1);
      if (blocking)
        mbar_wait(eb, par);
      else if (!mbar_try_wait(eb, par))
        return false;
    }
    const uint32_t bar = smem_u32(&full[slot]);
    const uint32_t dst = smem_u32(ring + (size_t)slot * GEMM_STAGE_ELEMS);
    bool tail = false;
    if constexpr (has_tail_stages<It>::value) tail = issue_it.is_tail();
    mbar_arrive_expect_tx(bar, tail ? TILE_BYTES : GEMM_STAGE_BYTES);
    bulk_g2s(dst, issue_it.A(), TILE_BYTES, bar);
    if (!tail) bulk_g2s(dst + TILE_BYTES, issue_it.B(), TILE_BYTES, bar);
    issue_it.next();
    ++issued;
    return true;
  };

  // start the ring before anything else
  if (tid == 0) {
    while (issue_it.valid() && issued < GEMM_STAGES) try_issue(false);
  }
  after_prologue();

  double acc[8][4][2];
#pragma unroll
  for (int fm = 0; fm < 8; ++fm)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) acc[fm][fn][0] = acc[fm][fn][1] = 0.0;

  constexpr bool IL = has_tri_stages<It>::value;            // interleaved row slabs: 2 fm + wm instead of 8 wm + fm
  constexpr int A_FM_STRIDE = IL ? 256 : 128;
  const int a_off = (IL ? wm : wm * 8) * 128 + lane * 2;   // row slab R -> +R*128 ; micro-col mc -> +mc*64
  const int b_off = (wn * 4) * 128 + lane * 2;

  int g = 0;
  while (cons_it.valid()) {
    const int slot = g % GEMM_STAGES;
    if (tid == 0) {
      while (issue_it.valid() && issued < g + GEMM_STAGES) {
        if (!try_issue(issued <= g)) break;
      }
    }
    mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / GEMM_STAGES) & 1));

    const double *As = ring + (size_t)slot * GEMM_STAGE_ELEMS;
    const double *Bs = As + TILE_ELEMS;
    if constexpr (has_tail_stages<It>::value) {
      if (cons_it.is_tail()) {
        // macro-tile m of C0 holds column slabs 2m and 2m+1 = (wn, fn) with 4 wn + fn in {2m, 2m+1}
        const int m = cons_it.tail_index();
        if (wn == (m >> 1)) {
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) {
            if ((fn >> 1) == (m & 1)) {
              double2 pr[8];
#pragma unroll
              for (int fm = 0; fm < 8; ++fm) pr[fm] = p_load_cfrag_raw(As, 8 * wm + fm, fn & 1, lane);
#pragma unroll
              for (int fm = 0; fm < 8; ++fm) {
                double c0, c1;
                p_unpair_cfrag(pr[fm], lane, c0, c1);
                acc[fm][fn][0] = c0 - acc[fm][fn][0];
                acc[fm][fn][1] = c1 - acc[fm][fn][1];
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
        cons_it.next();
        ++g;
        continue;
      }
    }
    if constexpr (has_tri_stages<It>::value) {
      if (cons_it.tri_diag()) {
        const int tg = cons_it.tri_g();
#pragma unroll 1
        for (int mc = 0; mc < 2; ++mc) {
          const int kk = 2 * tg + mc;
          double2 b[4];
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) b[fn] = lds128(Bs + b_off + fn * 128 + mc * 64);
          const double *Am = As + a_off + mc * 64;
          // live row slabs R = 2 fm + wm.  Real (warp-uniform) branches: a predicated-off DMMA still takes its slot in
          // the tensor pipe (measured: the predicated form of this loop saved 0.6 % instead of 5 %).
          if constexpr (It::kTriMode == 1) {          // lower triangular block: R >= kk  <=>  fm >= lo
            switch ((kk - wm + 1) >> 1) {
              case 0: tri_micro_step<0, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 1: tri_micro_step<1, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 2: tri_micro_step<2, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 3: tri_micro_step<3, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 4: tri_micro_step<4, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 5: tri_micro_step<5, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 6: tri_micro_step<6, 7, A_FM_STRIDE>(acc, Am, b); break;
              case 7: tri_micro_step<7, 7, A_FM_STRIDE>(acc, Am, b); break;
              default: break;
            }
          } else {                                    // upper triangular block: R <= kk  <=>  fm <= hi
            switch ((kk - wm) >> 1) {                 // arithmetic shift: -1 when kk < wm (nothing live)
              case 0: tri_micro_step<0, 0, A_FM_STRIDE>(acc, Am, b); break;
              case 1: tri_micro_step<0, 1, A_FM_STRIDE>(acc, Am, b); break;
              case 2: tri_micro_step<0, 2, A_FM_STRIDE>(acc, Am, b); break;
              case 3: tri_micro_step<0, 3, A_FM_STRIDE>(acc, Am, b); break;
              case 4: tri_micro_step<0, 4, A_FM_STRIDE>(acc, Am, b); break;
              case 5: tri_micro_step<0, 5, A_FM_STRIDE>(acc, Am, b); break;
              case 6: tri_micro_step<0, 6, A_FM_STRIDE>(acc, Am, b); break;
              case 7: tri_micro_step<0, 7, A_FM_STRIDE>(acc, Am, b); break;
              default: break;
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
        const bool t_end = cons_it.tile_end();
        const int t_id = cons_it.tile();
        cons_it.next();
        ++g;
        if (!SINGLE && t_end) {
          FragCoord fc{wm, wn, lane, IL};
          epi(t_id, acc, fc);
#pragma unroll
          for (int fm = 0; fm < 8; ++fm)
#pragma unroll
            for (int fn = 0; fn < 4; ++fn) acc[fm][fn][0] = acc[fm][fn][1] = 0.0;
        }
        continue;
      }
    }
#pragma unroll
    for (int mc = 0; mc < 2; ++mc) {
      double2 a[8], b[4];
#pragma unroll
      for (int fm = 0; fm < 8; ++fm) a[fm] = lds128(As + a_off + fm * A_FM_STRIDE + mc * 64);
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) b[fn] = lds128(Bs + b_off + fn * 128 + mc * 64);
#pragma unroll
      for (int fm = 0; fm < 8; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a[fm].x, b[fn].x);
      if (mc == 1) {   // every LDS.128 of this stage has been consumed by a DMMA above: release the slot
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
      }
#pragma unroll
      for (int fm = 0; fm < 8; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) dmma884(acc[fm][fn][0], acc[fm][fn][1], a[fm].y, b[fn].y);
    }

    const bool tile_end = cons_it.tile_end();
    const int tile = cons_it.tile();
    cons_it.next();
    ++g;
    if (!SINGLE && tile_end) {
      FragCoord fc{wm, wn, lane, IL};
      epi(tile, acc, fc);
#pragma unroll
      for (int fm = 0; fm < 8; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) acc[fm][fn][0] = acc[fm][fn][1] = 0.0;
    }
  }
  if (SINGLE) {
    FragCoord fc{wm, wn, lane, IL};
    epi(0, acc, fc);
  }
}

// One output tile with tail stages: `nk` macro-tiles of A and of B, then the 8 macro-tiles of the tile C0 itself.
struct TailIt {
  const double *a, *b, *c0;
  int left;   // main stages + 8
  __device__ __forceinline__ bool valid() const { return left > 0; }
  __device__ __forceinline__ bool is_tail() const { return left <= KT_PER_BLOCK; }
  __device__ __forceinline__ int tail_index() const { return KT_PER_BLOCK - left; }
  __device__ __forceinline__ const double *A() const { return left > KT_PER_BLOCK ? a : c0; }
  __device__ __forceinline__ const double *B() const { return b; }
  __device__ __forceinline__ bool tile_end() const { return left == 1; }
  __device__ __forceinline__ int tile() const { return 0; }
  __device__ __forceinline__ void next() {
    if (left > KT_PER_BLOCK) {
      a += TILE_ELEMS;
      b += TILE_ELEMS;
    } else {
      c0 += TILE_ELEMS;
    }
    --left;
  }
};

// Simple iterator: one output tile, `nk` consecutive macro-tiles of A and of B.
struct LinearIt {
  const double *a;
  const double *b;
  int left;
  __device__ __forceinline__ bool valid() const { return left > 0; }
  __device__ __forceinline__ const double *A() const { return a; }
  __device__ __forceinline__ const double *B() const { return b; }
  __device__ __forceinline__ bool tile_end() const { return left == 1; }
  __device__ __forceinline__ int tile() const { return 0; }
  __device__ __forceinline__ void next() {
    a += TILE_ELEMS;
    b += TILE_ELEMS;
    --left;
  }
};

// Epilogue helper: write (or read-modify-write) the 128x128 accumulator tile to a block in P-layout.
//   dst      : pointer to the first macro-tile of the destination block
//   transpose: element (r, c) is stored at local (c, r)
//   out = (cin ? cin[...] : 0) + scale * acc
__device__ __forceinline__ void store_block(double *dst, bool transpose, double scale, const double *cin,
                                            const double (&acc)[8][4][2], const FragCoord &fc) {
  if (!cin) {   // plain store: 16-byte lane-pair stores
#pragma unroll
    for (int fm = 0; fm < 8; ++fm)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) {
        if (transpose)
          p_store_cfrag_t(dst, fc.rslab(fm), 4 * fc.wn + fn, fc.lane, scale * acc[fm][fn][0], scale * acc[fm][fn][1]);
        else
          p_store_cfrag(dst, fc.rslab(fm), 4 * fc.wn + fn, fc.lane, scale * acc[fm][fn][0], scale * acc[fm][fn][1]);
      }
    return;
  }
#pragma unroll
  for (int fm = 0; fm < 8; ++fm) {
    const int r = fc.row(fm);
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = fc.col(fn, e);
        const int off = transpose ? block_offset(c, r) : block_offset(r, c);
        double v = scale * acc[fm][fn][e];
        if (cin) v += cin[off];
        dst[off] = v;
      }
    }
  }
}

// In-place trailing update C -= acc.  All 64 loads are issued before the first store (acc is reused as the
// staging register file), so the tile's read-modify-write costs one memory round trip instead of 64.
__device__ __forceinline__ void rmw_sub_block(double *dst, double (&acc)[8][4][2], const FragCoord &fc) {
  // the accumulators are moved into the P-layout's lane-pair arrangement once (one shuffle per fragment); loads,
  // subtraction and stores then work on 16-byte pairs
  double2 pr[8][4];
#pragma unroll
  for (int fm = 0; fm < 8; ++fm)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) pr[fm][fn] = p_load_cfrag_raw(dst, fc.rslab(fm), 4 * fc.wn + fn, fc.lane);
#pragma unroll
  for (int fm = 0; fm < 8; ++fm)
#pragma unroll
    for (int fn = 0; fn < 4; ++fn) {
      double c0, c1;
      p_unpair_cfrag(pr[fm][fn], fc.lane, c0, c1);
      p_store_cfrag(dst, fc.rslab(fm), 4 * fc.wn + fn, fc.lane, c0 - acc[fm][fn][0], c1 - acc[fm][fn][1]);
    }
}

// ---------------------------------------------------------------------------------------------
// Triangular variants of the mainloop for the two places where half of a 128x128 tile's work is
// structurally zero:
//   syrk_diag_pipeline  C = A B^T where only the lower triangle of C is wanted (diagonal tile of the
//                       Cholesky trailing update):            micro-tile (R, C) is live iff R >= C
//   trsm_tri_pipeline   C = A B^T with B lower triangular (panel solve L_ij = A_ij Winv_jj^T, K = 128):
//                       k-micro-step kk contributes to column slab C iff kk <= C
// Warp w owns the two 8-column slabs c0 = w and c1 = 15 - w, which deals the 136 live micro-tiles
// (of 256) evenly: 17 per warp, 34 per SM sub-partition.  Same TMA ring, same P-layout fragment loads
// (LDS.128); the live set is enumerated branch-free so loads and DMMAs schedule freely.
// ---------------------------------------------------------------------------------------------
struct TmaRing {
  double *ring;
  uint64_t *full, *empty;
  int issued;
  __device__ __forceinline__ void init() {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    ring = reinterpret_cast<double *>(smem_raw);
    full = reinterpret_cast<uint64_t *>(smem_raw + GEMM_RING_BYTES + GEMM_SCRATCH_BYTES);
    empty = full + GEMM_STAGES;
    issued = 0;
    if (threadIdx.x == 0) {
      for (int s = 0; s < GEMM_STAGES; ++s) {
        mbar_init(smem_u32(&full[s]), 1);
        mbar_init(smem_u32(&empty[s]), GEMM_THREADS / 32);
      }
      mbar_fence_init();
    }
    __syncthreads();
  }
  // thread 0 only: issue stage `issued`; false when its slot is still being read and !blocking
  template <class It>
  __device__ __forceinline__ bool try_issue(It &it, bool blocking) {
    const int slot = issued % GEMM_STAGES;
    if (issued >= GEMM_STAGES) {
      const uint32_t eb = smem_u32(&empty[slot]);
      const uint32_t par = (uint32_t)((issued / GEMM_STAGES - 1) & 1);
      if (blocking)
        mbar_wait(eb, par);
      else if (!mbar_try_wait(eb, par))
        return false;
    }
    const uint32_t bar = smem_u32(&full[slot]);
    const uint32_t dst = smem_u32(ring + (size_t)slot * GEMM_STAGE_ELEMS);
    mbar_arrive_expect_tx(bar, GEMM_STAGE_BYTES);
    bulk_g2s(dst, it.A(), TILE_BYTES, bar);
    bulk_g2s(dst + TILE_BYTES, it.B(), TILE_BYTES, bar);
    it.next();
    ++issued;
    return true;
  }
  // thread 0 only, before consuming stage g: keep up to GEMM_STAGES stages in flight (see gemm_pipeline)
  template <class It>
  __device__ __forceinline__ void feed(It &it, int g) {
    while (it.valid() && issued < g + GEMM_STAGES) {
      if (!try_issue(it, issued <= g)) break;
    }
  }
  __device__ __forceinline__ const double *wait(int g) const {
    mbar_wait(smem_u32(&full[g % GEMM_STAGES]), (uint32_t)((g / GEMM_STAGES) & 1));
    return ring + (size_t)(g % GEMM_STAGES) * GEMM_STAGE_ELEMS;
  }
  // all lanes of a warp, after the stage's shared-memory reads have been consumed
  __device__ __forceinline__ void release(int g) const {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(smem_u32(&empty[g % GEMM_STAGES]));
  }
};

// Live micro-tile t (0..16) of warp w in the lower-triangular output: t < 16 - w -> (row slab w + t, slab c0 = w),
// else (row slab (15 - w) + (t - (16 - w)), slab c1 = 15 - w).
struct SyrkCoord {
  int w, lane;
  __device__ __forceinline__ int rslab(int t) const { return t < 16 - w ? w + t : t - 1; }   // (15-w)+(t-(16-w)) = t-1
  __device__ __forceinline__ int cslab(int t) const { return t < 16 - w ? w : 15 - w; }
  __device__ __forceinline__ int row(int t) const { return 8 * rslab(t) + (lane >> 2); }
  __device__ __forceinline__ int col(int t, int e) const { return 8 * cslab(t) + 2 * (lane & 3) + e; }
};

// `hook(st, g)` runs on every warp for every stage g (st = the stage's A tile) before the slot is released: the
// batched log-likelihood rides its forward substitution's GEMV on the row block this kernel streams anyway.
struct NoStageHook {
  __device__ __forceinline__ void operator()(const double *, int) const {}
};
template <class Epi, class Hook = NoStageHook>
__device__ __forceinline__ void syrk_diag_pipeline(LinearIt issue_it, LinearIt cons_it, Epi &&epi, Hook &&hook = Hook()) {
  TmaRing rg;
  rg.init();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  double acc[17][2];
#pragma unroll
  for (int t = 0; t < 17; ++t) acc[t][0] = acc[t][1] = 0.0;
  const SyrkCoord sc{w, lane};
  int aoff[17];
#pragma unroll
  for (int t = 0; t < 17; ++t) aoff[t] = sc.rslab(t) * 128 + 2 * lane;
  const int b0off = TILE_ELEMS + w * 128 + 2 * lane, b1off = TILE_ELEMS + (15 - w) * 128 + 2 * lane;
  const int n0 = 16 - w;
  int g = 0;
  while (cons_it.valid()) {
    if (tid == 0) rg.feed(issue_it, g);
    const double *st = rg.wait(g);
#pragma unroll
    for (int mc = 0; mc < 2; ++mc) {
      const double2 b0 = lds128(st + b0off + mc * 64), b1 = lds128(st + b1off + mc * 64);
      double2 a[17];
#pragma unroll
      for (int t = 0; t < 17; ++t) a[t] = lds128(st + aoff[t] + mc * 64);
#pragma unroll
      for (int t = 0; t < 17; ++t) dmma884(acc[t][0], acc[t][1], a[t].x, t < n0 ? b0.x : b1.x);
      if (mc == 1) {
        hook(st, g);
        rg.release(g);   // every load of the stage has been consumed by a DMMA / the hook's FMAs
      }
#pragma unroll
      for (int t = 0; t < 17; ++t) dmma884(acc[t][0], acc[t][1], a[t].y, t < n0 ? b0.y : b1.y);
    }
    cons_it.next();
    ++g;
  }
  epi(acc, sc);
}

struct TrsmCoord {
  int c0, c1, lane;
  __device__ __forceinline__ int row(int R) const { return 8 * R + (lane >> 2); }
  __device__ __forceinline__ int col(int h, int e) const { return 8 * (h ? c1 : c0) + 2 * (lane & 3) + e; }
};

// K = 128 exactly (8 stages): B = Winv_jj (lower triangular, K-major rows = output columns)
template <class Epi>
__device__ __forceinline__ void trsm_tri_pipeline(LinearIt issue_it, LinearIt cons_it, Epi &&epi) {
  TmaRing rg;
  rg.init();
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int c0 = w, c1 = 15 - w;   // c0 < c1
  double acc[16][2][2];
#pragma unroll
  for (int R = 0; R < 16; ++R) acc[R][0][0] = acc[R][0][1] = acc[R][1][0] = acc[R][1][1] = 0.0;
  int g = 0;
  while (cons_it.valid()) {
    if (tid == 0) rg.feed(issue_it, g);
    const double *As = rg.wait(g) + 2 * lane;
    const double *Bs = As + TILE_ELEMS;
#pragma unroll
    for (int mc = 0; mc < 2; ++mc) {
      const int kk = 2 * g + mc;   // k micro-step 0..15
      if (kk <= c0) {              // both column slabs live
        const double2 b0 = lds128(Bs + (c0 * 2 + mc) * 64), b1 = lds128(Bs + (c1 * 2 + mc) * 64);
#pragma unroll
        for (int Rg = 0; Rg < 16; Rg += 8) {
          double2 a[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = lds128(As + ((Rg + q) * 2 + mc) * 64);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            dmma884(acc[Rg + q][0][0], acc[Rg + q][0][1], a[q].x, b0.x);
            dmma884(acc[Rg + q][1][0], acc[Rg + q][1][1], a[q].x, b1.x);
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            dmma884(acc[Rg + q][0][0], acc[Rg + q][0][1], a[q].y, b0.y);
            dmma884(acc[Rg + q][1][0], acc[Rg + q][1][1], a[q].y, b1.y);
          }
        }
      } else if (kk <= c1) {       // only the far slab
        const double2 b1 = lds128(Bs + (c1 * 2 + mc) * 64);
#pragma unroll
        for (int Rg = 0; Rg < 16; Rg += 8) {
          double2 a[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) a[q] = lds128(As + ((Rg + q) * 2 + mc) * 64);
#pragma unroll
          for (int q = 0; q < 8; ++q) dmma884(acc[Rg + q][1][0], acc[Rg + q][1][1], a[q].x, b1.x);
#pragma unroll
          for (int q = 0; q < 8; ++q) dmma884(acc[Rg + q][1][0], acc[Rg + q][1][1], a[q].y, b1.y);
        }
      }
    }
    rg.release(g);   // every LDS of the stage fed a DMMA that has already issued (operands were ready)
    cons_it.next();
    ++g;
  }
  // the epilogue overwrites the A operand's tile in global memory: every warp must have passed its last wait
  __syncthreads();
  TrsmCoord tc{c0, c1, lane};
  epi(acc, tc);
}

}  // namespace boss
