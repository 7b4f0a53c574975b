// potrf128.cuh -- Cholesky factor AND triangular inverse of one 128x128 diagonal block per CTA, on 8x8
// micro-tiles with FP64 tensor-core (DMMA.8x8x4) updates.
//
// This is the latency-critical step of the blocked factorisation (cholesky.cuh): every block column waits
// for it, and for batches of small matrices (BASELINE config C3: n = 512 -> 4 blocks per matrix) it is most
// of the run time.  Layout: the lower block triangle lives in shared memory as 136 packed 8x8 micro-tiles in
// the DMMA fragment order of the P-layout (common.cuh), 68 KB, so two CTAs share an SM and hide each other's
// dependency stalls.
//
//   factorisation  left-looking over the 16 tile columns s:
//                    update   C_Is = A_Is - sum_{t<s} L_It L_st^T        DMMA, all warps, accumulators in registers
//                    potf2    8x8 diagonal tile, one row per lane, pivots broadcast by warp shuffles
//                    trsm     rows of the panel below by forward substitution, one row per thread
//   inverse        8x8 diagonal inverses (one column per thread), then recursive doubling
//                    W21 = -W22 (L21 W11)   for block sizes 8, 16, 32, 64                DMMA
//
// Reference: LinearAlgebra.cholesky -> LAPACK dpotrf as called through AbstractGPs.posterior / logpdf
// (src/models/gaussian_process.jl:199-211, :269-280).
#pragma once
#include "common.cuh"

namespace boss {

struct PotrfParams {
  double *L;
  size_t L_stride;
  double *Winv;
  size_t Winv_stride;
  int nblk, ktiles, j;
  double *logdet_blk;  // S x nblk : sum_k log L_kk of this block
  int *status;         // S : BOSS_NOT_POSDEF on a non-positive / NaN pivot
  double *W, *WT;      // optional: also deposit Winv_jj into W(j,j), its transpose into WT(j,j)
  size_t W_stride;     // per matrix of the batch
  // in-line forward substitution of the batched log-likelihood (fwd_w null = off):
  //   w_j = Winv_jj (delta_j - r_j),  r_j = sum_{k<j} L_jk w_k from the diagonal-tile SYRK (chol_update_kernel)
  const double *fwd_ymm;   // delta = y - m(X): shared (fwd_ldy = 0) or one column per matrix
  long long fwd_ldy;
  int fwd_n, n_pad;
  const double *fwd_r;     // [S][n_pad]
  double *fwd_w;           // [S][n_pad]
  double *fwd_ssq;         // [S][nblk]  |w_j|^2
};

constexpr int PT_TILES = 136;                        // 16*17/2 lower micro-tiles
constexpr int PT_TMP_ELEMS = 64 * 64;                // scratch for the inverse products (one 64x64 block)
constexpr int PT_SMEM_BYTES = (PT_TILES * 64 + PT_TMP_ELEMS + 128 + 128 + 8 + 128 + 8) * 8;   // 105 600 B -> 2 CTAs / SM

// (I, J) of the t-th packed lower micro-tile
struct PtIJ {
  unsigned char I[PT_TILES], J[PT_TILES];
};
__host__ __device__ constexpr PtIJ pt_make_ij() {
  PtIJ r{};
  int t = 0;
  for (int I = 0; I < 16; ++I)
    for (int J = 0; J <= I; ++J) {
      r.I[t] = (unsigned char)I;
      r.J[t] = (unsigned char)J;
      ++t;
    }
  return r;
}
__constant__ PtIJ PT_IJ = pt_make_ij();

// packed lower-triangular micro-tile (I >= J)
__device__ __forceinline__ int pt_tile(int I, int J) { return ((I * (I + 1)) / 2 + J) * 64; }
// offset of element (r, c) inside a micro-tile (DMMA fragment order, see common.cuh)
__device__ __forceinline__ int pt_elem(int r, int c) { return (((r << 2) + (c & 3)) << 1) + (c >> 2); }
// offset of micro-tile (I, J) inside a 128x128 P-layout block (8 macro-tiles of 128 x 16)
__device__ __forceinline__ int pt_gtile(int I, int J) { return (J >> 1) * TILE_ELEMS + (((I << 1) + (J & 1)) << 6); }

// B operand of an "NN" product (B[k][n] = Q[k][n]): transposed fragment load of a stored tile
__device__ __forceinline__ double2 pt_ldT(const double *tile, int lane) {
  const int o = ((lane & 3) << 3) + (((lane >> 2) & 3) << 1) + (lane >> 4);
  return make_double2(tile[o], tile[o + 32]);
}

// C-fragment (lane holds (r = lane/4, c = 2(lane%4) + e)) <-> storage offsets
__device__ __forceinline__ int pt_cpos(int lane, int e) {
  const int r = lane >> 2, q = lane & 3;
  return (r << 3) + ((q & 1) << 2) + (e << 1) + (q >> 1);
}

__global__ void __launch_bounds__(256, 2) potrf_tile_kernel(PotrfParams p) {
  extern __shared__ __align__(16) double sm[];
  double *T = sm;                          // [136][64] packed lower tiles: A -> L -> W
  double *tmp = sm + PT_TILES * 64;        // [64][64] as 8x8 tiles (row-major tile grid, 8 tiles per row)
  double *dinv = tmp + PT_TMP_ELEMS;       // [128] 1 / L_kk
  double *lg = dinv + 128;                 // [128] log L_kk
  int *flag = reinterpret_cast<int *>(lg + 128);
  uint64_t *bar = reinterpret_cast<uint64_t *>(lg + 128 + 1);
  double *fv = lg + 128 + 8;               // [128] delta_j - r_j, then [8] warp partials of |w_j|^2
  const int s_mat = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *blk = p.L + (size_t)s_mat * p.L_stride + ((size_t)p.j * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS;

  // ---- load the lower block triangle: 136 TMA bulk copies of one 512 B micro-tile each, one mbarrier ----
  if (tid == 0) {
    *flag = 0;
    mbar_init(smem_u32(bar), 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) mbar_arrive_expect_tx(smem_u32(bar), PT_TILES * 512);
    __syncwarp();
    for (int t = lane; t < PT_TILES; t += 32)
      bulk_g2s(smem_u32(T + t * 64), blk + pt_gtile(PT_IJ.I[t], PT_IJ.J[t]), 512, smem_u32(bar));
  }
  mbar_wait(smem_u32(bar), 0);

  // ---- factorisation ----
  for (int s = 0; s < 16; ++s) {
    // (a) update tiles (I, s), I = s + warp, s + warp + 8   (4 independent accumulator pairs)
    if (s > 0) {
      for (int I = s + warp; I < 16; I += 8) {
        double c[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) c[q][0] = c[q][1] = 0.0;
        const double *ti = T + pt_tile(I, 0) + 2 * lane, *ts = T + pt_tile(s, 0) + 2 * lane;
        for (int t = 0; t < s; t += 4) {
          double2 a[4], b[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (t + q < s) {
              a[q] = lds128(ti + (t + q) * 64);
              b[q] = lds128(ts + (t + q) * 64);
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (t + q < s) dmma884(c[q][0], c[q][1], a[q].x, b[q].x);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (t + q < s) dmma884(c[q][0], c[q][1], a[q].y, b[q].y);
        }
        double *dst = T + pt_tile(I, s);
        dst[pt_cpos(lane, 0)] -= (c[0][0] + c[1][0]) + (c[2][0] + c[3][0]);
        dst[pt_cpos(lane, 1)] -= (c[0][1] + c[1][1]) + (c[2][1] + c[3][1]);
      }
      __syncthreads();
    }
    // (b) every participating thread factors the 8x8 diagonal tile redundantly in registers (no shuffles, no
    //     extra barrier), then solves its own row of the panel below:  x L_ss^T = c.  Thread 0 publishes L_ss.
    const int nrows = (15 - s) * 8;
    if (tid < (nrows > 32 ? nrows : 32)) {
      double *d = T + pt_tile(s, s);
      double l[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c <= i; ++c) l[i][c] = d[(i << 3) + ((c & 3) << 1) + (c >> 2)];
      double rsv[8];
      bool bad = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double piv = l[k][k];
        if (!(piv > 0.0)) {   // LAPACK dpotrf info > 0 (also catches NaN)
          bad = true;
          piv = 1.0;
        }
        const double rs = rsqrt(piv);
        rsv[k] = rs;
        l[k][k] = piv * rs;
#pragma unroll
        for (int i = k + 1; i < 8; ++i) l[i][k] *= rs;
#pragma unroll
        for (int jj = k + 1; jj < 8; ++jj)
#pragma unroll
          for (int i = jj; i < 8; ++i) l[i][jj] = fma(-l[i][k], l[jj][k], l[i][jj]);
      }
      if (tid < nrows) {
        const int I = s + 1 + (tid >> 3), r = tid & 7;
        double *row = T + pt_tile(I, s) + (r << 3);
        double x[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) x[c] = row[((c & 3) << 1) + (c >> 2)];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          double acc = x[k];
#pragma unroll
          for (int m = 0; m < k; ++m) acc = fma(-x[m], l[k][m], acc);
          x[k] = acc * rsv[k];
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) row[((c & 3) << 1) + (c >> 2)] = x[c];
      }
      if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int c = 0; c < 8; ++c) d[(i << 3) + ((c & 3) << 1) + (c >> 2)] = (c <= i) ? l[i][c] : 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) dinv[s * 8 + k] = rsv[k];
        if (bad) *flag = 1;
      }
    }
    __syncthreads();
  }

  // ---- log L_kk in parallel, L -> global (upper part zero); the fixed-order sum runs on an idle thread below ----
  if (tid < 128) lg[tid] = -log(dinv[tid]);
  for (int e = tid; e < TM * TM; e += 256) {
    const int kt = e >> 11, micro = (e >> 6) & 31, w = e & 63;
    const int I = micro >> 1, J = (kt << 1) + (micro & 1);
    blk[e] = (J <= I) ? T[pt_tile(I, J) + w] : 0.0;
  }
  __syncthreads();
  if (tid == 255) {
    if (*flag && p.status[s_mat] == 0) p.status[s_mat] = 1;
    double ld = 0.0;
    for (int k = 0; k < 128; ++k) ld += lg[k];
    p.logdet_blk[(size_t)s_mat * p.nblk + p.j] = ld;
  }

  // ---- inverse, level 0: 8x8 diagonal tiles, one column per thread (threads 0..127) -> tmp, then back ----
  if (tid < 128) {
    const int u = tid >> 3, c = tid & 7;
    const double *ld = T + pt_tile(u, u);
    double w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double acc = (i == c) ? 1.0 : 0.0;
#pragma unroll
      for (int m = 0; m < i; ++m) acc = fma(-ld[(i << 3) + ((m & 3) << 1) + (m >> 2)], w[m], acc);
      w[i] = (i >= c) ? acc * dinv[u * 8 + i] : 0.0;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) tmp[u * 64 + (i << 3) + ((c & 3) << 1) + (c >> 2)] = w[i];
  }
  __syncthreads();
  if (tid < 128) {
    const int u = tid >> 3, r = tid & 7;
    double *d = T + pt_tile(u, u) + (r << 3);
    const double *sv = tmp + u * 64 + (r << 3);
#pragma unroll
    for (int q = 0; q < 8; ++q) d[q] = sv[q];
  }
  __syncthreads();

  // ---- inverse, recursive doubling: diagonal blocks of mt tiles are inverted; merge pairs ----
  for (int mt = 1; mt < 16; mt <<= 1) {
    const int npair = 16 / (2 * mt), per = mt * mt;
    // step A: Tm(I, J) = sum_{t = J}^{mt-1} L21(I, t) W11(t, J)        -> tmp tile (pair, I, J)
    for (int o = warp; o < npair * per; o += 8) {
      const int pr = o / per, I = (o % per) / mt, J = o % mt;
      const int R0 = 2 * pr * mt + mt, C0 = 2 * pr * mt;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      int t = J;
      for (; t + 1 < mt; t += 2) {
        const double2 a = lds128(T + pt_tile(R0 + I, C0 + t) + 2 * lane), b = pt_ldT(T + pt_tile(C0 + t, C0 + J), lane);
        const double2 a2 = lds128(T + pt_tile(R0 + I, C0 + t + 1) + 2 * lane), b2 = pt_ldT(T + pt_tile(C0 + t + 1, C0 + J), lane);
        dmma884(c0, c1, a.x, b.x);
        dmma884(e0, e1, a2.x, b2.x);
        dmma884(c0, c1, a.y, b.y);
        dmma884(e0, e1, a2.y, b2.y);
      }
      if (t < mt) {
        const double2 a = lds128(T + pt_tile(R0 + I, C0 + t) + 2 * lane), b = pt_ldT(T + pt_tile(C0 + t, C0 + J), lane);
        dmma884(c0, c1, a.x, b.x);
        dmma884(c0, c1, a.y, b.y);
      }
      double *dst = tmp + (size_t)o * 64;
      dst[pt_cpos(lane, 0)] = c0 + e0;
      dst[pt_cpos(lane, 1)] = c1 + e1;
    }
    __syncthreads();
    // step B: W21(I, J) = - sum_{t = 0}^{I} W22(I, t) Tm(t, J)          -> overwrites L21
    for (int o = warp; o < npair * per; o += 8) {
      const int pr = o / per, I = (o % per) / mt, J = o % mt;
      const int R0 = 2 * pr * mt + mt, C0 = 2 * pr * mt;
      const double *tm = tmp + (size_t)pr * per * 64;
      double c0 = 0.0, c1 = 0.0, e0 = 0.0, e1 = 0.0;
      int t = 0;
      for (; t + 1 <= I; t += 2) {
        const double2 a = lds128(T + pt_tile(R0 + I, R0 + t) + 2 * lane), b = pt_ldT(tm + (t * mt + J) * 64, lane);
        const double2 a2 = lds128(T + pt_tile(R0 + I, R0 + t + 1) + 2 * lane), b2 = pt_ldT(tm + ((t + 1) * mt + J) * 64, lane);
        dmma884(c0, c1, a.x, b.x);
        dmma884(e0, e1, a2.x, b2.x);
        dmma884(c0, c1, a.y, b.y);
        dmma884(e0, e1, a2.y, b2.y);
      }
      if (t <= I) {
        const double2 a = lds128(T + pt_tile(R0 + I, R0 + t) + 2 * lane), b = pt_ldT(tm + (t * mt + J) * 64, lane);
        dmma884(c0, c1, a.x, b.x);
        dmma884(c0, c1, a.y, b.y);
      }
      double *dst = T + pt_tile(R0 + I, C0 + J);
      dst[pt_cpos(lane, 0)] = -(c0 + e0);
      dst[pt_cpos(lane, 1)] = -(c1 + e1);
    }
    __syncthreads();
  }

  // ---- Winv_jj -> global (and, for a posterior fit, into W(j,j) and transposed into WT(j,j)) ----
  double *wi = p.Winv + (size_t)s_mat * p.Winv_stride + (size_t)p.j * (TM * TM);
  double *wfull = p.W ? p.W + (size_t)s_mat * p.W_stride + ((size_t)p.j * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS : nullptr;
  double *wtfull = p.WT ? p.WT + (size_t)s_mat * p.W_stride + ((size_t)p.j * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS : nullptr;
  for (int e = tid; e < TM * TM; e += 256) {
    const int kt = e >> 11, micro = (e >> 6) & 31, w = e & 63;
    const int I = micro >> 1, J = (kt << 1) + (micro & 1);
    const double v = (J <= I) ? T[pt_tile(I, J) + w] : 0.0;
    wi[e] = v;
    if (wfull) wfull[e] = v;
    if (wtfull) {
      // element (r, c) of tile (I, J) of W^T is element (c, r) of tile (J, I) of W
      const int r = w >> 3, c = ((w >> 1) & 3) + ((w & 1) << 2);
      wtfull[e] = (I <= J) ? T[pt_tile(J, I) + pt_elem(c, r)] : 0.0;
    }
  }

  // ---- forward substitution step of the log-likelihood: w_j = Winv_jj (delta_j - r_j), |w_j|^2 ----
  if (p.fwd_w) {
    if (tid < 128) {
      const int row = p.j * 128 + tid;
      const double dl = (row < p.fwd_n) ? p.fwd_ymm[(size_t)s_mat * p.fwd_ldy + row] : 0.0;
      fv[tid] = dl - (p.j > 0 ? p.fwd_r[(size_t)s_mat * p.n_pad + row] : 0.0);
    }
    __syncthreads();
    // warp w: micro-rows 2w, 2w+1; lane T holds (row T/4, columns T%4 and 4 + T%4) of each 8x8 tile
    const int c4 = lane & 3;
    double sq = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int I = 2 * warp + m;
      double acc = 0.0;
      for (int J = 0; J <= I; ++J) {
        const double2 a = lds128(T + pt_tile(I, J) + 2 * lane);
        acc = fma(a.x, fv[J * 8 + c4], acc);
        acc = fma(a.y, fv[J * 8 + c4 + 4], acc);
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (c4 == 0) {
        p.fwd_w[(size_t)s_mat * p.n_pad + p.j * 128 + I * 8 + (lane >> 2)] = acc;
        sq = fma(acc, acc, sq);
      }
    }
    sq += __shfl_xor_sync(0xffffffffu, sq, 4);    // lanes with c4 == 0 hold the rows; fixed tree
    sq += __shfl_xor_sync(0xffffffffu, sq, 8);
    sq += __shfl_xor_sync(0xffffffffu, sq, 16);
    if (lane == 0) fv[128 + warp] = sq;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int q = 0; q < 8; ++q) t += fv[128 + q];
      p.fwd_ssq[(size_t)s_mat * p.nblk + p.j] = t;
    }
  }
}

}  // namespace boss
