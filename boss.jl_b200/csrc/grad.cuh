// grad.cuh -- analytic x-gradients of the acquisition for the multi-start optimiser
// (reference: OptimizationAM pushes ForwardDiff.Dual candidates through the whole posterior stack,
//  src/acquisition_maximizers/optimization.jl:36,100-113; here the same derivative in closed form).
//
//   mu(x)    = m(x) + sum_k alpha_k k(x, x_k)         d mu  = sum_k alpha_k dk_k
//   var(x)   = a^2 - |W k*(x)|^2                       d var = -2 sum_k u_k dk_k ,  u = W^T (W k*) = K^-1 k*
//   dk_k/dx_j = a^2 kappa'(r)/r * (x~_j - x~_kj) / l_j
//
// Pipeline per chunk and slice: xcov -> score_trmm (also stores V^T) -> wtv_kernel (U^T = (W^T V)^T, the
// second triangular product, same DMMA mainloop) -> grad_kernel (kernel derivatives, FP64 pipe) ->
// acq_grad_kernel (chain rule through EI x PoF).
#pragma once
#include "score.cuh"

namespace boss {

// U[k][c] = sum_{r >= k} WT[k][r] V[r][c] ;  A = WT row block i (tiles 8i .. ktiles-1), B = V^T block of the CTA
struct WtvParams {
  const double *WT;
  const double *VT;   // chunk scratch: rows = candidates, cols = V row index
  double *UT;         // chunk scratch: rows = candidates, cols = training index
  int nblk, ktiles;
};
struct WtvIt {
  const double *wt;      // W^T (upper triangular): row block i uses k-tiles 8i .. ktiles-1
  const double *vt;
  int j, i, kt, nblk, ktiles, y, ns2;
  __device__ __forceinline__ bool valid() const { return i < nblk; }
  __device__ __forceinline__ const double *A() const { return wt + ((size_t)i * ktiles + kt) * TILE_ELEMS; }
  __device__ __forceinline__ const double *B() const { return vt + (size_t)kt * TILE_ELEMS; }
  __device__ __forceinline__ bool tile_end() const { return kt == ktiles - 1; }
  __device__ __forceinline__ int tile() const { return i; }
  // the first 8 k-tiles of row block i are the diagonal block W_ii^T: upper triangular (gemm_core.cuh, has_tri_stages)
  static constexpr int kTriMode = 2;
  __device__ __forceinline__ bool tri_diag() const { return kt < (i + 1) * KT_PER_BLOCK; }
  __device__ __forceinline__ int tri_g() const { return kt - i * KT_PER_BLOCK; }
  __device__ __forceinline__ void next() {
    if (kt == ktiles - 1) {
      ++j;
      i = zigzag_row(j, y, ns2);
      kt = i * KT_PER_BLOCK;
    } else {
      ++kt;
    }
  }
};
// grid = (candidate blocks, ns row-block splits), same zig-zag deal as score_trmm_kernel
__global__ void __launch_bounds__(GEMM_THREADS, 1) wtv_kernel(WtvParams p) {
  const int cb = blockIdx.x, y = blockIdx.y, ns2 = 2 * gridDim.y;
  const int i0 = zigzag_row(0, y, ns2);
  WtvIt it{p.WT, p.VT + (size_t)cb * p.ktiles * TILE_ELEMS, 0, i0, i0 * KT_PER_BLOCK, p.nblk, p.ktiles, y, ns2};
  double *ut = p.UT + (size_t)cb * p.ktiles * TILE_ELEMS;
  gemm_pipeline(it, it, [&](int tile, const double(&acc)[8][4][2], const FragCoord &fc) {
    store_block(ut + (size_t)tile * KT_PER_BLOCK * TILE_ELEMS, true, 1.0, nullptr, acc, fc);
  });
}

// d mu / dx and d var / dx for one slice and one chunk.


// Fixed-order sum of the per-chunk partials, then 1/l_j, the -2 of d var and the discrete mask:
//   dmu[j][c] = invl_j * sum_p gm_part[p][j][c] ;  dvar[j][c] = -2 invl_j * sum_p gv_part[p][j][c]
struct GradReduceParams {
  const double *gm_part, *gv_part;
  int P, d, chunk_ld, count;
  const double *invl;
  unsigned long long disc_bits;
  double *dmu, *dvar;  // [d][chunk_ld] each (this slice's section)
};
__global__ void __launch_bounds__(256) grad_reduce_kernel(GradReduceParams p) {
  const int c = blockIdx.x * 256 + threadIdx.x, j = blockIdx.y;
  if (c >= p.count) return;
  const size_t o0 = (size_t)j * p.chunk_ld + c, st = (size_t)p.d * p.chunk_ld;
  const double am = ordered_sum(p.gm_part + o0, st, p.P), av = ordered_sum(p.gv_part + o0, st, p.P);
  const bool disc = (p.disc_bits >> j) & 1ull;
  const double il = p.invl[j];
  p.dmu[(size_t)j * p.chunk_ld + c] = disc ? 0.0 : am * il;
  p.dvar[(size_t)j * p.chunk_ld + c] = disc ? 0.0 : -2.0 * av * il;
}

// Chain rule through EI x PoF.  dmu / dvar: [(sample*y_dim + slice)][d][chunk_ld].
struct AcqGradParams {
  AcqParams a;
  const double *dmu, *dvar;
  const double *prior_mean_grad;  // y_dim x d x M (column-major: index ((m)*d + j)*y_dim + i) or null
  double *grad;                   // d x M, indexed by (m - out_off)
};

__global__ void __launch_bounds__(128) acq_grad_kernel(AcqGradParams q) {
  const AcqParams &p = q.a;
  const int c = blockIdx.x * 128 + threadIdx.x;
  const long long m = p.m0 + c;
  if (c >= p.chunk || m >= p.M) return;
  const int d = p.d;
  double gacc[32];
  for (int j = 0; j < d; ++j) gacc[j] = 0.0;
  bool failed = false;
  for (int s = 0; s < p.n_samples; ++s) {
    // pass 1: scalars
    double cdf_i[MAX_YDIM], wmu_i[MAX_YDIM], wvar_i[MAX_YDIM];
    bool live_i[MAX_YDIM];
    double mu_f = 0.0, s2 = 0.0, pof = 1.0;
    for (int i = 0; i < p.y_dim; ++i) {
      const size_t row = (size_t)(s * p.y_dim + i) * p.chunk_ld + c;
      double mu = p.mu[row];
      if (p.prior_mean) mu = p.prior_mean[(size_t)(m - p.in_off) * p.y_dim + i] + mu;
      double var = p.a2[s * p.y_dim + i] - p.sumsq[row] + VAR_JITTER;
      const bool unclipped = var >= 0.0;
      if (!clip_var(var)) failed = true;
      live_i[i] = unclipped;   // the clipped branch of _clip_var returns a constant zero -> zero derivative
      mu_f = fma(p.coefs[i], mu, mu_f);
      s2 = fma(p.coefs[i] * p.coefs[i], var, s2);
      cdf_i[i] = 1.0;
      wmu_i[i] = wvar_i[i] = 0.0;
      if (p.has_ymax && !(p.y_max[i] == INFINITY)) {
        const double sd = sqrt(var);
        double z = (p.y_max[i] - mu) / sd;
        if (sd == 0.0 && p.y_max[i] == mu) z = INFINITY;
        cdf_i[i] = norm_cdf(z);
        if (sd > 0.0) {   // d cdf = pdf(z) * ( -dmu/sd - (ymax-mu) dvar / (2 sd var) )
          const double pz = norm_pdf(z);
          wmu_i[i] = -pz / sd;
          wvar_i[i] = -pz * (p.y_max[i] - mu) / (2.0 * sd * var);
        }
        pof *= cdf_i[i];
      }
    }
    double ei = 0.0, cphi = 0.0, cpdf_over_2sf = 0.0;   // d ei = cphi * d mu_f + cpdf_over_2sf * d s2
    if (p.has_best) {
      const double sf = sqrt(s2), diff = mu_f - p.best;
      if (diff == 0.0 && sf == 0.0) {
        ei = 0.0;
      } else {
        const double z = diff / sf;
        const double cz = norm_cdf(z), pz = norm_pdf(z);
        ei = diff * cz + sf * pz;
        if (sf > 0.0) {
          cphi = cz;
          cpdf_over_2sf = pz / (2.0 * sf);
        } else {
          cphi = diff > 0.0 ? 1.0 : 0.0;
        }
      }
    }
    // pass 2: per input dimension
    for (int j = 0; j < d; ++j) {
      double dmu_f = 0.0, ds2 = 0.0, dpof = 0.0;
      for (int i = 0; i < p.y_dim; ++i) {
        const size_t row = ((size_t)(s * p.y_dim + i) * d + j) * p.chunk_ld + c;
        double dm = q.dmu[row];
        if (q.prior_mean_grad) dm += q.prior_mean_grad[((size_t)(m - p.in_off) * d + j) * p.y_dim + i];
        const double dv = live_i[i] ? q.dvar[row] : 0.0;
        dmu_f = fma(p.coefs[i], dm, dmu_f);
        ds2 = fma(p.coefs[i] * p.coefs[i], dv, ds2);
        if (p.has_ymax && !(p.y_max[i] == INFINITY)) {
          double others = 1.0;
          for (int b = 0; b < p.y_dim; ++b)
            if (b != i) others *= cdf_i[b];
          dpof = fma(wmu_i[i] * dm + wvar_i[i] * dv, others, dpof);
        }
      }
      const double dei = cphi * dmu_f + cpdf_over_2sf * ds2;
      double g;
      if (p.has_best)
        g = p.has_ymax ? dei * pof + ei * dpof : dei;
      else
        g = p.has_ymax ? dpof : 0.0;
      gacc[j] += g;
    }
  }
  bool zero = failed;
  if (p.lb) {
    for (int i = 0; i < d; ++i) {
      const double x = p.Xs[(size_t)(m - p.in_off) * d + i];
      if (x < p.lb[i] || x > p.ub[i]) zero = true;
    }
  }
  if (p.cons_mask && p.cons_mask[m - p.in_off] == 0) zero = true;
  for (int j = 0; j < d; ++j) q.grad[(size_t)(m - p.out_off) * d + j] = zero ? 0.0 : gacc[j] / (double)p.n_samples;
}

}  // namespace boss
