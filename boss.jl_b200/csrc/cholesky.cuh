// cholesky.cuh -- batched blocked FP64 Cholesky, triangular inverse and log-likelihood pieces.
//
// Left-looking block algorithm on 128x128 blocks, S matrices per launch (one grid.y slice each):
//   for j:  UPDATE  A_ij -= L_i,0:j L_j,0:j^T   (i >= j)      gemm core, K = 128 j      [DMMA]
//           POTRF   L_jj = chol(A_jj), Winv_jj = L_jj^-1       one CTA per matrix, smem
//           TRSM    L_ij  = A_ij Winv_jj^T        (i >  j)      gemm core, K = 128        [DMMA]
// followed, for a posterior fit, by the full triangular inverse W = L^-1 (block recurrence, gemm core)
// and, for the log-likelihood, by a blocked forward substitution.
// Reference: AbstractGPs.posterior / logpdf(::FiniteGP) -> LinearAlgebra.cholesky (LAPACK dpotrf) as
// called from src/models/gaussian_process.jl:199-211 and :269-280.
#pragma once
#include "gemm_core.cuh"
#include "kernel_fn.cuh"
#include "potrf128.cuh"

namespace boss {

// ---------------------------------------------------------------------------------------------
// K2: fused ARD kernel-matrix construction (lower block triangle, P-layout), per hyper-parameter sample
// ---------------------------------------------------------------------------------------------
struct BuildKParams {
  const double *X;          // d x n raw training inputs (shared by all samples)
  int d, n, nblk, ktiles;   // nblk = n_pad/128, ktiles = n_pad/16
  const double *ls;         // d x S raw length-scales
  const double *amp;        // S raw amplitudes
  const double *noise;      // S raw noise std
  unsigned long long disc_bits;
  double *K;                // S matrices, P-layout, stride K_stride
  size_t K_stride;
  int *status;              // S; set to -1 on negative hyper-parameters (reference asserts)
};

template <int KID, int DP>
__global__ void __launch_bounds__(256) build_k_kernel(BuildKParams p) {
  __shared__ double xcol[128 * DP];
  __shared__ double invl[DP];
  const int s = blockIdx.y;
  // decode (rb >= cb) from the linear lower-triangle index
  int t = blockIdx.x, rb = 0;
  while (t >= rb + 1) {
    t -= rb + 1;
    ++rb;
  }
  const int cb = t;
  const int tid = threadIdx.x;
  if (tid < DP) {
    double l = (tid < p.d) ? p.ls[(size_t)s * p.d + tid] : 1.0;
    if (tid < p.d && !(l >= 0.0)) p.status[s] = -1;
    invl[tid] = (tid < p.d) ? 1.0 / (l + MIN_PARAM_VALUE) : 0.0;
  }
  __syncthreads();
  const double a_raw = p.amp[s], s_raw = p.noise[s];
  if (tid == 0 && (!(a_raw >= 0.0) || !(s_raw >= 0.0))) p.status[s] = -1;
  const double a = a_raw + MIN_PARAM_VALUE, sn = s_raw + MIN_PARAM_VALUE;
  const double a2 = a * a, s2 = sn * sn;

  // column points of this block -> smem (scaled)
  for (int e = tid; e < 128 * DP; e += 256) {
    const int jj = e / DP, i = e % DP;
    const int j = cb * 128 + jj;
    double v = 0.0;
    if (j < p.n && i < p.d) {
      v = p.X[(size_t)j * p.d + i];
      if ((p.disc_bits >> i) & 1ull) v = rint(v);
      v *= invl[i];
    }
    xcol[e] = v;
  }
  const int r = tid & 127, kh = tid >> 7;
  const int gi = rb * 128 + r;
  double xr[DP];
  load_scaled_point<DP>(xr, p.X + (size_t)gi * p.d, p.d, invl, p.disc_bits, gi < p.n);
  __syncthreads();

  double *blk = p.K + (size_t)s * p.K_stride + ((size_t)rb * p.ktiles + (size_t)cb * KT_PER_BLOCK) * TILE_ELEMS;
  for (int mcol = kh; mcol < 16; mcol += 2) {
    double v[8];
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int jj = mcol * 8 + kk;
      const int gj = cb * 128 + jj;
      double d2 = 0.0;
#pragma unroll
      for (int i = 0; i < DP; ++i) {
        const double df = xr[i] - xcol[jj * DP + i];
        d2 = fma(df, df, d2);
      }
      double val;
      if (gi < p.n && gj < p.n) {
        val = a2 * kappa<KID>(d2);
        if (gi == gj) val += s2;
      } else {
        val = (gi == gj) ? 1.0 : 0.0;  // identity padding: L_pad = I, log-det contribution 0
      }
      v[kk] = val;
    }
    double *dst = blk + (mcol >> 1) * TILE_ELEMS + ((((r >> 3) << 1) + (mcol & 1)) << 6) + ((r & 7) << 3);
#pragma unroll
    for (int q = 0; q < 4; ++q) *reinterpret_cast<double2 *>(dst + 2 * q) = make_double2(v[q], v[q + 4]);
  }
}

// ---------------------------------------------------------------------------------------------
// gemm-core users
// ---------------------------------------------------------------------------------------------
struct CholGemmParams {
  double *L;             // S matrices (P-layout)
  size_t L_stride;
  double *Winv;          // S x nblk diagonal-block inverses, each a 128x128 P-layout block (8 tiles)
  size_t Winv_stride;
  int nblk, ktiles;
  int j;                 // block column (UPDATE / TRSM) or block distance delta (TRTRI)
  double *W, *WT, *TT;   // full inverse, its transpose, per-task scratch (TRTRI; single matrix)
};

// A_ij -= L_i,0:j * L_j,0:j^T   for i = j + blockIdx.x
__global__ void __launch_bounds__(GEMM_THREADS, 1) chol_update_kernel(CholGemmParams p) {
  const int i = p.j + blockIdx.x;
  double *Lm = p.L + (size_t)blockIdx.y * p.L_stride;
  LinearIt it{Lm + (size_t)i * p.ktiles * TILE_ELEMS, Lm + (size_t)p.j * p.ktiles * TILE_ELEMS, p.j * KT_PER_BLOCK};
  double *dst = Lm + ((size_t)i * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS;
  gemm_pipeline(it, it, [&](int, double(&acc)[8][4][2], const FragCoord &fc) { rmw_sub_block(dst, acc, fc); });
}

// L_ij = A_ij * Winv_jj^T       for i = j + 1 + blockIdx.x
__global__ void __launch_bounds__(GEMM_THREADS, 1) chol_trsm_kernel(CholGemmParams p) {
  const int i = p.j + 1 + blockIdx.x;
  double *Lm = p.L + (size_t)blockIdx.y * p.L_stride;
  double *blk = Lm + ((size_t)i * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS;
  const double *wi = p.Winv + (size_t)blockIdx.y * p.Winv_stride + (size_t)p.j * (TM * TM);
  LinearIt it{blk, wi, KT_PER_BLOCK};
  gemm_pipeline(it, it, [&](int, const double(&acc)[8][4][2], const FragCoord &fc) {
    store_block(blk, false, 1.0, nullptr, acc, fc);
  });
}

// Triangular inverse, block distance delta = i - j (single matrix):
//   T_ij = sum_{k=j}^{i-1} L_ik W_kj   -> stored transposed in TT[task]
__global__ void __launch_bounds__(GEMM_THREADS, 1) trtri_t_kernel(CholGemmParams p) {
  const int jb = blockIdx.x, i = jb + p.j;
  LinearIt it{p.L + ((size_t)i * p.ktiles + (size_t)jb * KT_PER_BLOCK) * TILE_ELEMS,
              p.WT + ((size_t)jb * p.ktiles + (size_t)jb * KT_PER_BLOCK) * TILE_ELEMS, p.j * KT_PER_BLOCK};
  double *dst = p.TT + (size_t)blockIdx.x * (TM * TM);
  gemm_pipeline(it, it, [&](int, const double(&acc)[8][4][2], const FragCoord &fc) {
    store_block(dst, true, 1.0, nullptr, acc, fc);
  });
}
//   W_ij = -Winv_ii T_ij   -> W (normal) and WT (transposed)
__global__ void __launch_bounds__(GEMM_THREADS, 1) trtri_w_kernel(CholGemmParams p) {
  const int jb = blockIdx.x, i = jb + p.j;
  LinearIt it{p.Winv + (size_t)i * (TM * TM), p.TT + (size_t)blockIdx.x * (TM * TM), KT_PER_BLOCK};
  double *dW = p.W + ((size_t)i * p.ktiles + (size_t)jb * KT_PER_BLOCK) * TILE_ELEMS;
  double *dWT = p.WT + ((size_t)jb * p.ktiles + (size_t)i * KT_PER_BLOCK) * TILE_ELEMS;
  gemm_pipeline(it, it, [&](int, const double(&acc)[8][4][2], const FragCoord &fc) {
    store_block(dW, false, -1.0, nullptr, acc, fc);
    store_block(dWT, true, -1.0, nullptr, acc, fc);
  });
}

// Generic C = A B^T on P-layout operands (test hook boss_dbg_gemm_nt): grid (N/128, M/128)
struct DbgGemmParams {
  const double *A, *B;
  double *C;
  int ktilesAB;  // K/16
  int ktilesC;   // N/16
  int lower_only;  // skip tiles above the block diagonal (symmetric products: boss_gp_cov)
};
__global__ void __launch_bounds__(GEMM_THREADS, 1) dbg_gemm_kernel(DbgGemmParams p) {
  const int cb = blockIdx.x, rb = blockIdx.y;
  if (p.lower_only && cb > rb) return;
  LinearIt it{p.A + (size_t)rb * p.ktilesAB * TILE_ELEMS, p.B + (size_t)cb * p.ktilesAB * TILE_ELEMS, p.ktilesAB};
  double *dst = p.C + ((size_t)rb * p.ktilesC + (size_t)cb * KT_PER_BLOCK) * TILE_ELEMS;
  gemm_pipeline(it, it, [&](int, const double(&acc)[8][4][2], const FragCoord &fc) {
    store_block(dst, false, 1.0, nullptr, acc, fc);
  });
}

// ---------------------------------------------------------------------------------------------
// POTRF of one 128x128 diagonal block per matrix + its triangular inverse (shared memory)
// ---------------------------------------------------------------------------------------------
constexpr int POTRF_LD = 129;
constexpr int POTRF_SMEM_BYTES = (128 * POTRF_LD + 128 + 3 * 32 * 33) * 8;

// decode offset e (0..16383) inside a 128x128 P-layout block -> (r, c)
__host__ __device__ __forceinline__ void block_decode(int e, int &r, int &c) {
  const int kt = e >> 11, micro = (e >> 6) & 31, within = e & 63;
  r = ((micro >> 1) << 3) + (within >> 3);
  c = (kt << 4) + ((micro & 1) << 3) + ((within >> 1) & 3) + ((within & 1) << 2);
}

constexpr int POTRF_PB = 32;  // panel width of the in-CTA blocked factorisation

// One CTA per matrix.  Shared-memory 128x128 block: lower triangle = L, strict upper triangle is
// reused for W^T (W = L^-1) in the inverse phase, diagonal of W in wd[].
//   factorisation: 4 panels of 32 columns.  Inside a panel every thread owns one row in registers; per
//                  column ONE barrier: every thread recomputes the pivot's dot product from the
//                  published pivot row (shared broadcast reads shared with its own dot product), so the
//                  pivot -> rest serialisation disappears; 1/sqrt via rsqrt.  The rank-32 trailing update
//                  handles 4 columns per pass (4 independent FMA chains, no store->load stalls).
//   inverse:       32x32 diagonal blocks by register forward substitution (one warp each), off-diagonal
//                  blocks by block distance with all 256 threads (short independent dot products).
__global__ void __launch_bounds__(256, 1) potrf_diag_kernel(PotrfParams p) {
  extern __shared__ __align__(16) double sm[];
  double *A = sm;                       // [128][129]
  double *wd = sm + 128 * POTRF_LD;     // [128] log(L_kk) scratch, then diagonal of W
  double *tbuf = wd + 128;              // [3][32][33] T blocks of the inverse recurrence
  const int s = blockIdx.x, tid = threadIdx.x;
  double *blk = p.L + (size_t)s * p.L_stride + ((size_t)p.j * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS;

  for (int e = tid; e < TM * TM; e += 256) {
    int r, c;
    block_decode(e, r, c);
    A[r * POTRF_LD + c] = blk[e];
  }
  __syncthreads();

  bool bad = false;
  for (int pnl = 0; pnl < TM / POTRF_PB; ++pnl) {
    const int c0 = pnl * POTRF_PB;
    // ---- (a) panel factorisation, rows c0..127, columns c0..c0+31 ----
    double r[POTRF_PB];
    const bool own = (tid < 128) && (tid >= c0);
#pragma unroll
    for (int k = 0; k < POTRF_PB; ++k) r[k] = own ? A[tid * POTRF_LD + c0 + k] : 0.0;
#pragma unroll
    for (int k = 0; k < POTRF_PB; ++k) {
      const int gk = c0 + k;
      const double *pr = A + gk * POTRF_LD + c0;  // pivot row: entries m < k are final
      double p0 = 0.0, p1 = 0.0, s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int m = 0; m < k; ++m) {
        const double v = pr[m];
        if (m & 1) {
          p1 = fma(v, v, p1);
          s1 = fma(r[m], v, s1);
        } else {
          p0 = fma(v, v, p0);
          s0 = fma(r[m], v, s0);
        }
      }
      double dkk = pr[k] - (p0 + p1);
      if (!(dkk > 0.0)) {  // LAPACK dpotrf info > 0 (also catches NaN); every thread sees the same value
        bad = true;
        dkk = 1.0;
      }
      const double rs = rsqrt(dkk);
      if (own && tid >= gk) {
        r[k] = (tid == gk) ? dkk * rs : (r[k] - (s0 + s1)) * rs;
        if (tid < c0 + POTRF_PB) A[tid * POTRF_LD + gk] = r[k];  // future pivot rows publish column k
      }
      __syncthreads();
    }
    if (own && tid >= c0 + POTRF_PB) {
#pragma unroll
      for (int k = 0; k < POTRF_PB; ++k) A[tid * POTRF_LD + c0 + k] = r[k];
    }
    __syncthreads();
    // ---- (b) rank-32 trailing update of the lower triangle right of the panel ----
    {
      const int i = tid & 127, jh = tid >> 7;
      if (i >= c0 + POTRF_PB) {
        double li[POTRF_PB];
#pragma unroll
        for (int m = 0; m < POTRF_PB; ++m) li[m] = A[i * POTRF_LD + c0 + m];
        for (int jj = c0 + POTRF_PB + jh; jj <= i; jj += 8) {
          const int j0 = jj, j1 = min(jj + 2, 127), j2 = min(jj + 4, 127), j3 = min(jj + 6, 127);
          const double *l0 = A + j0 * POTRF_LD + c0, *l1 = A + j1 * POTRF_LD + c0, *l2 = A + j2 * POTRF_LD + c0,
                       *l3 = A + j3 * POTRF_LD + c0;
          double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
          for (int m = 0; m < POTRF_PB; ++m) {
            a0 = fma(li[m], l0[m], a0);
            a1 = fma(li[m], l1[m], a1);
            a2 = fma(li[m], l2[m], a2);
            a3 = fma(li[m], l3[m], a3);
          }
          double *ai = A + i * POTRF_LD;
          const double o0 = ai[j0], o1 = ai[j1], o2 = ai[j2], o3 = ai[j3];
          ai[j0] = o0 - a0;
          if (jj + 2 <= i) ai[jj + 2] = o1 - a1;
          if (jj + 4 <= i) ai[jj + 4] = o2 - a2;
          if (jj + 6 <= i) ai[jj + 6] = o3 - a3;
        }
      }
    }
    __syncthreads();
  }

  // ---- log-determinant contribution (fixed summation order) ----
  if (tid < 128) wd[tid] = log(A[tid * POTRF_LD + tid]);
  __syncthreads();
  if (tid == 0) {
    if (bad && p.status[s] == 0) p.status[s] = 1;
    double ld = 0.0;
    for (int k = 0; k < 128; ++k) ld += wd[k];
    p.logdet_blk[(size_t)s * p.nblk + p.j] = ld;
  }
  // ---- write L_jj (upper part zeroed) ----
  for (int e = tid; e < TM * TM; e += 256) {
    int r2, c2;
    block_decode(e, r2, c2);
    blk[e] = (c2 <= r2) ? A[r2 * POTRF_LD + c2] : 0.0;
  }
  __syncthreads();

  // ---- inverse, step 1: D_b = inv(L_bb) for the four 32x32 diagonal blocks (warp b, lane = column) ----
  if (tid < 128) {
    const int b0 = (tid >> 5) << 5, c = tid & 31;
    double w[POTRF_PB];
#pragma unroll
    for (int i = 0; i < POTRF_PB; ++i) {
      const double *Li = A + (b0 + i) * POTRF_LD + b0;
      double s0 = (i == c) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
      for (int m = 0; m < i; ++m) {
        if (m & 1)
          s1 = fma(-Li[m], w[m], s1);
        else
          s0 = fma(-Li[m], w[m], s0);
      }
      w[i] = (s0 + s1) / Li[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < POTRF_PB; ++i) {
      if (i > c) A[(b0 + c) * POTRF_LD + b0 + i] = w[i];   // D_b[i][c] stored transposed
      if (i == c) wd[b0 + c] = w[i];
    }
  }
  __syncthreads();
  // ---- inverse, step 2: off-diagonal 32x32 blocks by block distance ----
  {
    const int lane = tid & 31;
    for (int delta = 1; delta < 4; ++delta) {
      const int npair = 4 - delta;
      // T_ij = sum_{q = 32j}^{32i-1} L[32i + r][q] * W[q][32j + c]
      for (int e = tid; e < npair * 1024; e += 256) {
        const int bp = e >> 10, c = (e & 1023) >> 5, rr = lane;
        const int jb = bp, ib = bp + delta;
        const double *lrow = A + (32 * ib + rr) * POTRF_LD;
        const double *wrow = A + (32 * jb + c) * POTRF_LD;
        const int cc = 32 * jb + c;
        const double wcc = wd[cc];
        double t0 = 0.0, t1 = 0.0;
        for (int q = 32 * jb; q < 32 * ib; q += 2) {
          const double w0 = (q > cc) ? wrow[q] : (q == cc ? wcc : 0.0);
          const double w1 = (q + 1 > cc) ? wrow[q + 1] : (q + 1 == cc ? wcc : 0.0);
          t0 = fma(lrow[q], w0, t0);
          t1 = fma(lrow[q + 1], w1, t1);
        }
        tbuf[(bp * 32 + rr) * 33 + c] = t0 + t1;
      }
      __syncthreads();
      // W_ij[r][c] = - sum_{m <= r} D_i[r][m] T[m][c]   -> stored transposed at A[32j + c][32i + r]
      for (int e = tid; e < npair * 1024; e += 256) {
        const int bp = e >> 10, c = (e & 1023) >> 5, rr = lane;
        const int jb = bp, ib = bp + delta;
        const double *tb = tbuf + bp * 32 * 33 + c;
        const double drr = wd[32 * ib + rr];
        double t0 = 0.0, t1 = 0.0;
        for (int m = 0; m < 32; m += 2) {
          const double d0 = (m < rr) ? A[(32 * ib + m) * POTRF_LD + 32 * ib + rr] : (m == rr ? drr : 0.0);
          const double d1 = (m + 1 < rr) ? A[(32 * ib + m + 1) * POTRF_LD + 32 * ib + rr] : (m + 1 == rr ? drr : 0.0);
          t0 = fma(d0, tb[m * 33], t0);
          t1 = fma(d1, tb[(m + 1) * 33], t1);
        }
        A[(32 * jb + c) * POTRF_LD + 32 * ib + rr] = -(t0 + t1);
      }
      __syncthreads();
    }
  }

  double *wi = p.Winv + (size_t)s * p.Winv_stride + (size_t)p.j * (TM * TM);
  double *wfull = p.W ? p.W + ((size_t)p.j * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS : nullptr;
  double *wtfull = p.WT ? p.WT + ((size_t)p.j * p.ktiles + (size_t)p.j * KT_PER_BLOCK) * TILE_ELEMS : nullptr;
  for (int e = tid; e < TM * TM; e += 256) {
    int r2, c2;
    block_decode(e, r2, c2);
    const double v = (c2 < r2) ? A[c2 * POTRF_LD + r2] : (c2 == r2 ? wd[r2] : 0.0);
    wi[e] = v;
    if (wfull) wfull[e] = v;
    if (wtfull) wtfull[e] = (r2 < c2) ? A[r2 * POTRF_LD + c2] : (c2 == r2 ? wd[r2] : 0.0);
  }
}

// ---------------------------------------------------------------------------------------------
// y = M x for a P-layout n_pad x n_pad matrix (one warp per 8-row micro-row).  Used for
// w = W delta and alpha = W^T w in the posterior fit.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) matvec_p_kernel(const double *__restrict__ Mx, const double *__restrict__ x,
                                                       double *__restrict__ y, int ktiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int mrow = blockIdx.x * 8 + warp;      // global micro-row
  const int rb = mrow >> 4, mr = mrow & 15;
  const int c4 = lane & 3;
  const double *base = Mx + (size_t)rb * ktiles * TILE_ELEMS + (mr * 2) * 64 + lane * 2;
  double acc = 0.0;
  for (int kt = 0; kt < ktiles; ++kt) {
#pragma unroll
    for (int mc = 0; mc < 2; ++mc) {
      const double2 v = *reinterpret_cast<const double2 *>(base + (size_t)kt * TILE_ELEMS + mc * 64);
      const int k = kt * 16 + mc * 8 + c4;
      acc = fma(v.x, x[k], acc);
      acc = fma(v.y, x[k + 4], acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (c4 == 0) y[mrow * 8 + (lane >> 2)] = acc;
}

// ---------------------------------------------------------------------------------------------
// Blocked forward substitution + log-likelihood assembly, one CTA per sample:
//   w_i = Winv_ii (delta_i - sum_{k<i} L_ik w_k) ;  ll = -(n log 2pi + 2 sum log L_kk + |w|^2)/2
// (AbstractGPs logpdf(::FiniteGP, y), reference call site src/models/gaussian_process.jl:278-279)
// ---------------------------------------------------------------------------------------------
struct FwdParams {
  const double *L;
  size_t L_stride;
  const double *Winv;
  size_t Winv_stride;
  int nblk, ktiles, n;
  const double *ymm;   // Y - m(X): shared (ldy = 0) or per sample
  long long ldy;
  const double *logdet_blk;
  const int *status;
  double *loglik;      // S
  double *w_out;       // optional S x n_pad (unused for loglik)
};

__global__ void __launch_bounds__(256) fwd_solve_loglik_kernel(FwdParams p) {
  extern __shared__ __align__(16) double sm[];
  double *w = sm;                        // [n_pad]
  double *t = sm + p.nblk * 128;         // [128]
  __shared__ double red[8];
  const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c4 = lane & 3, r8 = lane >> 2;
  const double *Lm = p.L + (size_t)s * p.L_stride;
  const double *Wi = p.Winv + (size_t)s * p.Winv_stride;
  const double *ymm = p.ymm + (size_t)s * p.ldy;

  for (int i = 0; i < p.nblk; ++i) {
    for (int mr = warp; mr < 16; mr += 8) {
      const double *base = Lm + (size_t)i * p.ktiles * TILE_ELEMS + (mr * 2) * 64 + lane * 2;
      double acc = 0.0;
      for (int kt = 0; kt < i * KT_PER_BLOCK; ++kt) {
#pragma unroll
        for (int mc = 0; mc < 2; ++mc) {
          const double2 v = *reinterpret_cast<const double2 *>(base + (size_t)kt * TILE_ELEMS + mc * 64);
          const int k = kt * 16 + mc * 8 + c4;
          acc = fma(v.x, w[k], acc);
          acc = fma(v.y, w[k + 4], acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (c4 == 0) {
        const int row = i * 128 + mr * 8 + r8;
        const double dl = (row < p.n) ? ymm[row] : 0.0;
        t[mr * 8 + r8] = dl - acc;
      }
    }
    __syncthreads();
    for (int mr = warp; mr < 16; mr += 8) {
      const double *base = Wi + (size_t)i * (TM * TM) + (mr * 2) * 64 + lane * 2;
      double acc = 0.0;
#pragma unroll
      for (int kt = 0; kt < KT_PER_BLOCK; ++kt) {
#pragma unroll
        for (int mc = 0; mc < 2; ++mc) {
          const double2 v = *reinterpret_cast<const double2 *>(base + kt * TILE_ELEMS + mc * 64);
          const int k = kt * 16 + mc * 8 + c4;
          acc = fma(v.x, t[k], acc);
          acc = fma(v.y, t[k + 4], acc);
        }
      }
      acc += __shfl_xor_sync(0xffffffffu, acc, 1);
      acc += __shfl_xor_sync(0xffffffffu, acc, 2);
      if (c4 == 0) w[i * 128 + mr * 8 + r8] = acc;
    }
    __syncthreads();
  }
  // |w|^2 : fixed-order reduction (thread-strided partials -> warp shuffle tree -> 8 warp partials in order)
  double part = 0.0;
  for (int k = tid; k < p.nblk * 128; k += 256) part = fma(w[k], w[k], part);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) red[warp] = part;
  if (p.w_out)
    for (int k = tid; k < p.nblk * 128; k += 256) p.w_out[(size_t)s * p.nblk * 128 + k] = w[k];
  __syncthreads();
  if (tid == 0) {
    double mahal = 0.0;
    for (int q = 0; q < 8; ++q) mahal += red[q];
    double ld = 0.0;
    for (int b = 0; b < p.nblk; ++b) ld += p.logdet_blk[(size_t)s * p.nblk + b];
    double ll = -((double)p.n * 1.8378770664093453 + 2.0 * ld + mahal) * 0.5;
    const int st = p.status[s];
    if (st == 1) ll = -INFINITY;
    if (st < 0) ll = NAN;
    p.loglik[s] = ll;
  }
}

}  // namespace boss
