// score_narrow.cuh -- the triangular products of the scoring path for SMALL candidate batches: 32 candidates per CTA.
//
// score_trmm_kernel / wtv_kernel (score.cuh, grad.cuh) give a CTA 128 candidates; a batch of a few dozen points -- the
// unfinished starts of the multi-start optimiser after its first rounds, single-point calls of the reference's
// one-x-at-a-time access pattern -- then still pays 128-wide DMMA work on every stage of the longest row block
// (0.56 ms per pass at n = 4096, whatever the batch size).  Here a CTA owns 32 candidates: a stage is the same 16 KB
// tile of W plus a 4 KB slice of K*^T, the tensor work per stage is a quarter, and four times as many CTAs share the
// row blocks.  The two kernels produce THE SAME BITS as their wide twins: identical accumulation order inside a tile
// (k ascending, the same structurally-zero fragments skipped) and, for the column sums of squares, the wide kernel's
// reduction order is replayed explicitly (rows of one parity in fragment order, then the 3-level butterfly over the
// eight rows of a slab) -- a candidate scores identically alone, in a 1 000-start batch or in a 40 000-point grid
// (tests/test_gpu_parity.py::test_small_batch_split_is_bitwise_invariant).
#pragma once
#include "grad.cuh"
#include "score.cuh"

namespace boss {

constexpr int NW_NB = 32;                                   // candidates per CTA
constexpr int NW_STAGES = 6;
constexpr int NW_STAGE_ELEMS = TILE_ELEMS + TILE_ELEMS / 4;  // A: 128 x 16 tile of W (16 KB) + B: 32 x 16 slice (4 KB)
constexpr int NW_V_ELEMS = 128 * NW_NB;                      // one tile of V (or U) staged for the ordered reduction
constexpr int NW_SMEM_BYTES = (NW_STAGES * NW_STAGE_ELEMS + NW_V_ELEMS + 2 * 8 * NW_NB) * 8 + 2 * NW_STAGES * 8;

struct NarrowParams {
  const double *A;      // W (MODE 0, lower triangular) or W^T (MODE 1, upper triangular), P-layout
  const double *B;      // K*^T chunk (MODE 0) or V^T chunk (MODE 1): rows = candidates, P-layout blocks of 128
  int nblk, ktiles;
  double *ss_part;      // MODE 0: [2*nblk][ld] partial column sums of squares (same layout as score_trmm_kernel)
  int ld;
  double *OT;           // MODE 0: optional V^T chunk (gradient mode); MODE 1: U^T chunk
};

// MODE 0:  V = W K*  (row block i uses k-tiles 0 .. 8(i+1)-1, the last 8 are the lower-triangular diagonal block)
// MODE 1:  U = W^T V (row block i uses k-tiles 8i .. ktiles-1, the first 8 are the upper-triangular diagonal block)
// grid = (32-candidate blocks, row-block splits); row blocks are dealt in the same zig-zag order as the wide kernels.
template <int MODE>
__global__ void __launch_bounds__(256, 1) score_narrow_kernel(const __grid_constant__ NarrowParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  double *ring = reinterpret_cast<double *>(smem_raw);
  double *Vst = ring + NW_STAGES * NW_STAGE_ELEMS;          // [16 slabs][8 rows][32 candidates]
  double *chn = Vst + NW_V_ELEMS;                           // [2 parities][8 rows][32 candidates]
  uint64_t *full = reinterpret_cast<uint64_t *>(chn + 2 * 8 * NW_NB), *empty = full + NW_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int cb32 = blockIdx.x, cb = cb32 >> 2, sub = cb32 & 3, y = blockIdx.y, ns2 = 2 * gridDim.y;
  if (tid < 2 * NW_STAGES) {
    mbar_init(smem_u32(tid < NW_STAGES ? &full[tid] : &empty[tid - NW_STAGES]), tid < NW_STAGES ? 1u : 8u);
    mbar_fence_init();
  }
  __syncthreads();

  // producer state (thread 0): the stream of (row block, k-tile) pairs of this CTA
  int pj = 0, pi = zigzag_row(0, y, ns2), pkt = MODE == 0 ? 0 : pi * KT_PER_BLOCK, issued = 0;
  const double *Bblk = p.B + (size_t)cb * p.ktiles * TILE_ELEMS + (size_t)sub * (TILE_ELEMS / 4);
  auto try_issue = [&](bool blocking) -> bool {
    const int slot = issued % NW_STAGES;
    if (issued >= NW_STAGES) {
      const uint32_t eb = smem_u32(&empty[slot]);
      const uint32_t par = (uint32_t)((issued / NW_STAGES - 1) & 1);
      if (blocking)
        mbar_wait(eb, par);
      else if (!mbar_try_wait(eb, par))
        return false;
    }
    const uint32_t bar = smem_u32(&full[slot]);
    double *dst = ring + (size_t)slot * NW_STAGE_ELEMS;
    mbar_arrive_expect_tx(bar, TILE_BYTES + TILE_BYTES / 4);
    bulk_g2s(smem_u32(dst), p.A + ((size_t)pi * p.ktiles + pkt) * TILE_ELEMS, TILE_BYTES, bar);
    bulk_g2s(smem_u32(dst + TILE_ELEMS), Bblk + (size_t)pkt * TILE_ELEMS, TILE_BYTES / 4, bar);
    ++issued;
    const int last = MODE == 0 ? (pi + 1) * KT_PER_BLOCK - 1 : p.ktiles - 1;
    if (pkt == last) {
      ++pj;
      pi = zigzag_row(pj, y, ns2);
      pkt = MODE == 0 ? 0 : pi * KT_PER_BLOCK;
    } else {
      ++pkt;
    }
    return true;
  };
  if (tid == 0)
    while (pi < p.nblk && issued < NW_STAGES) try_issue(false);

  const int R0 = 2 * w;   // this warp's row slabs R0, R0 + 1 of the current row block
  int g = 0;
  for (int j = 0;; ++j) {
    const int i = zigzag_row(j, y, ns2);
    if (i >= p.nblk) break;
    const int kt0 = MODE == 0 ? 0 : i * KT_PER_BLOCK, kt1 = MODE == 0 ? (i + 1) * KT_PER_BLOCK : p.ktiles;
    double acc[2][4][2];
#pragma unroll
    for (int fm = 0; fm < 2; ++fm)
#pragma unroll
      for (int fn = 0; fn < 4; ++fn) acc[fm][fn][0] = acc[fm][fn][1] = 0.0;
    for (int kt = kt0; kt < kt1; ++kt, ++g) {
      if (tid == 0) {
        while (pi < p.nblk && issued < g + NW_STAGES) {
          if (!try_issue(issued <= g)) break;
        }
      }
      const int slot = g % NW_STAGES;
      mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / NW_STAGES) & 1));
      const double *As = ring + (size_t)slot * NW_STAGE_ELEMS + 2 * lane, *Bs = As + TILE_ELEMS;
      const int dg = kt - i * KT_PER_BLOCK;          // k-tile inside the diagonal block when 0 <= dg < 8
      const bool diag = dg >= 0 && dg < KT_PER_BLOCK;
#pragma unroll 1
      for (int mc = 0; mc < 2; ++mc) {
        // live row slabs (warp-uniform, real branches): lower block R >= kk, upper block R <= kk
        bool l0 = true, l1 = true;
        if (diag) {
          const int kk = 2 * dg + mc;
          l0 = MODE == 0 ? R0 >= kk : R0 <= kk;
          l1 = MODE == 0 ? R0 + 1 >= kk : R0 + 1 <= kk;
        }
        if (!(l0 || l1)) continue;
        double2 b[4];
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) b[fn] = lds128(Bs + fn * 128 + mc * 64);
        if (l0) {
          const double2 a = lds128(As + R0 * 128 + mc * 64);
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) dmma884(acc[0][fn][0], acc[0][fn][1], a.x, b[fn].x);
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) dmma884(acc[0][fn][0], acc[0][fn][1], a.y, b[fn].y);
        }
        if (l1) {
          const double2 a = lds128(As + (R0 + 1) * 128 + mc * 64);
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) dmma884(acc[1][fn][0], acc[1][fn][1], a.x, b[fn].x);
#pragma unroll
          for (int fn = 0; fn < 4; ++fn) dmma884(acc[1][fn][0], acc[1][fn][1], a.y, b[fn].y);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&empty[slot]));
    }
    // ---- tile (row block i) finished ----
    if (p.OT) {   // transposed store: rows = candidates (block cb, slabs 4 sub + fn), columns = row index of block i
      double *ot = p.OT + (size_t)cb * p.ktiles * TILE_ELEMS + (size_t)i * KT_PER_BLOCK * TILE_ELEMS;
#pragma unroll
      for (int fm = 0; fm < 2; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) p_store_cfrag_t(ot, R0 + fm, 4 * sub + fn, lane, acc[fm][fn][0], acc[fm][fn][1]);
    }
    if (MODE == 0) {
      // column sums of squares in the wide kernel's order: for each parity pr of the row slab, the chain
      // v = fma(x, x, v) over the slabs R = pr, pr + 2, ..., pr + 14 (fixed row-in-slab r), then the butterfly over r
      const int r = lane >> 2, c2 = 2 * (lane & 3);
#pragma unroll
      for (int fm = 0; fm < 2; ++fm)
#pragma unroll
        for (int fn = 0; fn < 4; ++fn) {
          Vst[((R0 + fm) * 8 + r) * NW_NB + 8 * fn + c2] = acc[fm][fn][0];
          Vst[((R0 + fm) * 8 + r) * NW_NB + 8 * fn + c2 + 1] = acc[fm][fn][1];
        }
      __syncthreads();
      for (int e = tid; e < 2 * 8 * NW_NB; e += 256) {
        const int pr = e / (8 * NW_NB), rr = (e / NW_NB) & 7, cc = e % NW_NB;
        double v = 0.0;
#pragma unroll
        for (int fmw = 0; fmw < 8; ++fmw) {
          const double x = Vst[((2 * fmw + pr) * 8 + rr) * NW_NB + cc];
          v = fma(x, x, v);
        }
        chn[e] = v;
      }
      __syncthreads();
      if (tid < 2 * NW_NB) {
        const int pr = tid / NW_NB, cc = tid % NW_NB;
        const double *q = chn + pr * 8 * NW_NB + cc;
        const double s = ((q[0] + q[NW_NB]) + (q[2 * NW_NB] + q[3 * NW_NB])) + ((q[4 * NW_NB] + q[5 * NW_NB]) + (q[6 * NW_NB] + q[7 * NW_NB]));
        p.ss_part[(size_t)(2 * i + pr) * p.ld + (size_t)cb * 128 + 32 * sub + cc] = s;
      }
    }
  }
}

}  // namespace boss
