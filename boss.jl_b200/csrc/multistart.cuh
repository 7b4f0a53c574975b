// multistart.cuh -- per-start pieces of the device-resident lock-step multi-start optimiser
// (projected L-BFGS with backtracking, maximisation).  One thread per start; every value + gradient evaluation of
// ALL starts is one pass of the scoring pipeline (score.cuh / grad.cuh) on device-resident points.
//
// Reference: OptimizationAM (src/acquisition_maximizers/optimization.jl:55-118) runs `multistart` independent
// local solves through optimize_multistart (src/utils/optim_multistart.jl:10-90), each evaluating acq(x) one
// point at a time; here the starts advance together so the GPU sees batches.
#pragma once
#include "common.cuh"

namespace boss {

constexpr int MS_MAXD = 32;
constexpr int MS_MAXH = 16;

struct MsState {
  int d, H;
  long long M;
  double *X, *Xt, *Xn;       // [M][d] current / trial / accepted
  double *f, *ft, *fn;       // [M]
  double *g, *gt, *gn;       // [M][d]
  double *dirn, *t;          // [M][d], [M]
  int *done;                 // [M] accepted a step in this iteration (frozen starts are born done)
  int *frozen;               // [M] no acceptable step exists any more: converged, excluded from further work
  int *idx;                  // [M] compacted list of the starts still backtracking
  double *Xc, *fc, *gc;      // compact trial points / values / gradients of those starts
  double *Sh, *Yh;           // [H][M][d] curvature pairs (ring)
  double lb[MS_MAXD], ub[MS_MAXD];
  double step0;
  unsigned long long *moved_bits;   // max |Xn - X| as ordered bit pattern
  int *counters;             // [0] = compact list length, [1] = any valid curvature pair, [2] = starts rejected this trial
};

__device__ __forceinline__ double ms_clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void ms_init_kernel(MsState st, const double *starts) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= st.M) return;
  for (int j = 0; j < st.d; ++j) st.X[m * st.d + j] = ms_clip(starts[m * st.d + j], st.lb[j], st.ub[j]);
}

// non-finite acquisition values (SafeFunction's -Inf, NaN) count as -Inf
__global__ void ms_sanitize_kernel(double *f, long long M) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m < M && !isfinite(f[m])) f[m] = -INFINITY;
}

// two-loop recursion (ascent direction) ; hist_len pairs, oldest at ring slot hist_start
__global__ void ms_direction_kernel(MsState st, int hist_len, int hist_start) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= st.M) return;
  const int d = st.d;
  if (st.frozen[m]) {
    for (int j = 0; j < d; ++j) {
      st.Xn[m * d + j] = st.X[m * d + j];
      st.gn[m * d + j] = st.g[m * d + j];
    }
    st.fn[m] = st.f[m];
    st.done[m] = 1;
    return;
  }
  double q[MS_MAXD], gl[MS_MAXD], al[MS_MAXH], rh[MS_MAXH];
  double gnorm2 = 0.0;
  for (int j = 0; j < d; ++j) {
    gl[j] = st.g[m * d + j];
    q[j] = gl[j];
    gnorm2 = fma(gl[j], gl[j], gnorm2);
  }
  const double sd = st.step0 / fmax(sqrt(gnorm2), 1e-300);
  for (int k = hist_len - 1; k >= 0; --k) {        // newest -> oldest
    const size_t o = ((size_t)((hist_start + k) % st.H) * st.M + m) * d;
    double sy = 0.0, sq = 0.0;
    for (int j = 0; j < d; ++j) {
      sy = fma(st.Sh[o + j], st.Yh[o + j], sy);
      sq = fma(st.Sh[o + j], q[j], sq);
    }
    rh[k] = 1.0 / fmax(sy, 1e-300);
    al[k] = rh[k] * sq;
    for (int j = 0; j < d; ++j) q[j] = fma(-al[k], st.Yh[o + j], q[j]);
  }
  if (hist_len > 0) {
    const size_t o = ((size_t)((hist_start + hist_len - 1) % st.H) * st.M + m) * d;
    double sy = 0.0, yy = 0.0;
    for (int j = 0; j < d; ++j) {
      sy = fma(st.Sh[o + j], st.Yh[o + j], sy);
      yy = fma(st.Yh[o + j], st.Yh[o + j], yy);
    }
    const double gamma = sy / fmax(yy, 1e-300);
    const double sc = (isfinite(gamma) && gamma > 0.0) ? gamma : 1.0;
    for (int j = 0; j < d; ++j) q[j] *= sc;
  } else {
    for (int j = 0; j < d; ++j) q[j] *= sd;
  }
  for (int k = 0; k < hist_len; ++k) {             // oldest -> newest
    const size_t o = ((size_t)((hist_start + k) % st.H) * st.M + m) * d;
    double yq = 0.0;
    for (int j = 0; j < d; ++j) yq = fma(st.Yh[o + j], q[j], yq);
    const double b = rh[k] * yq;
    for (int j = 0; j < d; ++j) q[j] = fma(st.Sh[o + j], al[k] - b, q[j]);
  }
  double dg = 0.0;
  for (int j = 0; j < d; ++j) dg = fma(q[j], gl[j], dg);
  const bool bad = !(dg > 0.0);                    // not an ascent direction -> steepest ascent
  for (int j = 0; j < d; ++j) {
    st.dirn[m * d + j] = bad ? gl[j] * sd : q[j];
    st.Xn[m * d + j] = st.X[m * d + j];
    st.gn[m * d + j] = gl[j];
  }
  st.fn[m] = st.f[m];
  st.t[m] = 1.0;
  st.done[m] = 0;
}

// compact the starts that still need a trial point: idx[slot] = m, Xc[slot] = clip(X + t dirn)
__global__ void ms_compact_kernel(MsState st) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= st.M || st.done[m]) return;
  const int slot = atomicAdd(&st.counters[0], 1);
  st.idx[slot] = (int)m;
  const double t = st.t[m];
  for (int j = 0; j < st.d; ++j)
    st.Xc[(size_t)slot * st.d + j] = ms_clip(fma(st.dirn[m * st.d + j], t, st.X[m * st.d + j]), st.lb[j], st.ub[j]);
}

// Armijo acceptance of the compacted trial points; rejected starts halve their step
__global__ void ms_accept_kernel(MsState st, int count) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= count) return;
  const long long m = st.idx[slot];
  const int d = st.d;
  double ft = st.fc[slot];
  if (!isfinite(ft)) ft = -INFINITY;
  double slope = 0.0;
  for (int j = 0; j < d; ++j) slope = fma(st.g[m * d + j], st.Xc[(size_t)slot * d + j] - st.X[m * d + j], slope);
  if (ft >= st.f[m] + 1e-4 * slope) {
    for (int j = 0; j < d; ++j) {
      st.Xn[m * d + j] = st.Xc[(size_t)slot * d + j];
      st.gn[m * d + j] = st.gc[(size_t)slot * d + j];
    }
    st.fn[m] = ft;
    st.done[m] = 1;
  } else {
    st.t[m] *= 0.5;
    atomicAdd(&st.counters[2], 1);
  }
}

// Remaining backtracking steps of the stragglers in ONE batch: trial points for t, t/2, ..., t/2^(MS_FAN-1)
// (the sequential loop would evaluate them one after the other and stop at the first that passes).
constexpr int MS_FAN = 10;
__global__ void ms_fan_kernel(MsState st, int count) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= count * MS_FAN) return;
  const int slot = e / MS_FAN, k = e % MS_FAN;
  const long long m = st.idx[slot];
  const double t = ldexp(st.t[m], -k);
  for (int j = 0; j < st.d; ++j)
    st.Xc[(size_t)e * st.d + j] = ms_clip(fma(st.dirn[m * st.d + j], t, st.X[m * st.d + j]), st.lb[j], st.ub[j]);
}
__global__ void ms_accept_fan_kernel(MsState st, int count) {
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= count) return;
  const long long m = st.idx[slot];
  const int d = st.d;
  for (int k = 0; k < MS_FAN; ++k) {      // largest step first == the order of the sequential backtracking
    const size_t e = (size_t)slot * MS_FAN + k;
    double ft = st.fc[e];
    if (!isfinite(ft)) ft = -INFINITY;
    double slope = 0.0;
    for (int j = 0; j < d; ++j) slope = fma(st.g[m * d + j], st.Xc[e * d + j] - st.X[m * d + j], slope);
    if (ft >= st.f[m] + 1e-4 * slope) {
      for (int j = 0; j < d; ++j) {
        st.Xn[m * d + j] = st.Xc[e * d + j];
        st.gn[m * d + j] = st.gc[e * d + j];
      }
      st.fn[m] = ft;
      st.done[m] = 1;
      return;
    }
  }
}

// starts that exhausted the backtracking budget cannot improve any more
__global__ void ms_freeze_kernel(MsState st) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m < st.M && !st.done[m]) st.frozen[m] = 1;
}

// curvature pair into ring slot `slot` (zero pair when s.y is not safely positive), step size statistics, X <- Xn
__global__ void ms_update_kernel(MsState st, int slot) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= st.M) return;
  const int d = st.d;
  const size_t o = ((size_t)slot * st.M + m) * d;
  double sy = 0.0, mv = 0.0;
  for (int j = 0; j < d; ++j) {
    const double s = st.Xn[m * d + j] - st.X[m * d + j], y = -(st.gn[m * d + j] - st.g[m * d + j]);
    sy = fma(s, y, sy);
    mv = fmax(mv, fabs(s));
  }
  const bool valid = sy > 1e-16;
  for (int j = 0; j < d; ++j) {
    const double s = st.Xn[m * d + j] - st.X[m * d + j], y = -(st.gn[m * d + j] - st.g[m * d + j]);
    st.Sh[o + j] = valid ? s : 0.0;
    st.Yh[o + j] = valid ? y : 0.0;
    st.X[m * d + j] = st.Xn[m * d + j];
    st.g[m * d + j] = st.gn[m * d + j];
  }
  st.f[m] = st.fn[m];
  if (valid) st.counters[1] = 1;
  atomicMax(st.moved_bits, (unsigned long long)__double_as_longlong(mv));
}

// affine prior mean m_i(x) = c_i + b_i . x evaluated on the device (constant means and linear parametric models
// of a Semiparametric surrogate; general closures stay on the host): aff[i*(d+1)] = c_i, aff[i*(d+1)+1+j] = b_ij
__global__ void ms_affine_mean_kernel(const double *X, long long M, int d, int y_dim, const double *aff, double *pm,
                                      double *pmg) {
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= M) return;
  for (int i = 0; i < y_dim; ++i) {
    double v = aff[i * (d + 1)];
    for (int j = 0; j < d; ++j) {
      v = fma(aff[i * (d + 1) + 1 + j], X[m * d + j], v);
      pmg[((size_t)m * d + j) * y_dim + i] = aff[i * (d + 1) + 1 + j];
    }
    pm[(size_t)m * y_dim + i] = v;
  }
}

// final discrete rounding (optimization.jl:116-117)
__global__ void ms_round_kernel(double *X, long long M, int d, unsigned long long disc_bits) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= M * d) return;
  if ((disc_bits >> (e % d)) & 1ull) X[e] = rint(X[e]);
}

}  // namespace boss
