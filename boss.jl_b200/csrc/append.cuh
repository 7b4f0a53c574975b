// append.cuh -- incremental factor cache: add one training point to a fitted GP in O(n^2).
//
// BOSS.jl refactors K from scratch whenever the data set grows by a point: every speculative evaluation of
// SequentialBatchAM (src/acquisition_maximizers/batch.jl:26-38 -> augment_dataset!, model_posterior) and every
// BO iteration (src/bo.jl:40-45) pay a full Cholesky.  With hyper-parameters unchanged the factor only gains a row:
//
//   k = k(X, x+)            l = L^-1 k = W k           lambda = sqrt(a^2 + s^2 - l.l)
//   L+ = [L 0; l^T lambda]  W+ = [W 0; -(W^T l)^T/lambda  1/lambda]
//   w+ = (delta+ - l.w)/lambda  (w = L^-1 delta)        alpha+ = W+^T [w; w+]
//   loglik+ = loglik - (log 2pi)/2 - log(lambda) - w+^2 / 2
//
// Row n of the 128-padded P-layout matrices (identity padding so far) is overwritten in place.
#pragma once
#include "cholesky.cuh"

namespace boss {


// One CTA: lambda, w+ and status from l = W k (fixed-order reductions).
//   out[0] = lambda, out[1] = w+, out[2] = l.l ; status = 1 when a^2 + s^2 - l.l <= 0 (not positive definite)
__global__ void __launch_bounds__(256) append_scalars_kernel(const double *l, const double *w, int n, double kss,
                                                             double delta_new, double *out, int *status) {
  __shared__ double r1[256], r2[256];
  const int tid = threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  for (int i = tid; i < n; i += 256) {
    s1 = fma(l[i], l[i], s1);
    s2 = fma(l[i], w[i], s2);
  }
  r1[tid] = s1;
  r2[tid] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      r1[tid] += r1[tid + o];
      r2[tid] += r2[tid + o];
    }
    __syncthreads();
  }
  if (tid == 0) {
    const double piv = kss - r1[0];
    if (!(piv > 0.0)) {
      *status = 1;
      out[0] = out[1] = 0.0;
    } else {
      const double lam = sqrt(piv);
      out[0] = lam;
      out[1] = (delta_new - r2[0]) / lam;
    }
    out[2] = r1[0];
  }
}

// Write row n of L and W, column n of W^T, w[n] and delta[n].  u = W^T l.
__global__ void append_scatter_kernel(double *L, double *W, double *WT, int n, int ktiles, const double *l,
                                      const double *u, const double *sc, const int *status, double *wvec, double *ymm,
                                      double delta_new) {
  if (*status != 0) return;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const double lam = sc[0], il = 1.0 / lam;
  if (i < n) {
    const double wv = -u[i] * il;
    L[p_index(n, i, ktiles)] = l[i];
    W[p_index(n, i, ktiles)] = wv;
    WT[p_index(i, n, ktiles)] = wv;
  } else {
    L[p_index(n, n, ktiles)] = lam;
    W[p_index(n, n, ktiles)] = il;
    WT[p_index(n, n, ktiles)] = il;
    wvec[n] = sc[1];
    ymm[n] = delta_new;
  }
}

// Grow a P-layout matrix from nblk_old to nblk_new row/column blocks: copy the old tiles, identity on the new
// diagonal, zero elsewhere.  One thread per destination element.
__global__ void repack_grow_kernel(const double *src, int kt_old, int npad_old, double *dst, int kt_new, int npad_new) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t total = (size_t)npad_new * npad_new;
  if (e >= total) return;
  // decode destination (r, c) from its P-layout offset
  const size_t tile = e / TILE_ELEMS;
  const int within = (int)(e % TILE_ELEMS);
  const int rb = (int)(tile / kt_new), kt = (int)(tile % kt_new);
  const int micro = within >> 6, w = within & 63;
  const int r = rb * 128 + ((micro >> 1) << 3) + (w >> 3);
  const int c = kt * 16 + ((micro & 1) << 3) + ((w >> 1) & 3) + ((w & 1) << 2);
  double v;
  if (r < npad_old && c < npad_old)
    v = src[p_index(r, c, kt_old)];
  else
    v = (r == c) ? 1.0 : 0.0;
  dst[e] = v;
}

}  // namespace boss
