// kernel_fn.cuh -- covariance function evaluation shared by the kernel-matrix builder (a1) and the
// cross-covariance generator (a3).  Mirrors KernelFunctions' kappa for SqExponential / Matern32 /
// Matern52 composed with ARDTransform(1/l) and ScaledKernel(a^2)  (reference call site:
// src/models/gaussian_process.jl:243), optional DiscreteKernel rounding (src/models/utils/kernels.jl:56-64).
#pragma once
#include "common.cuh"

namespace boss {

constexpr double MIN_PARAM_VALUE = 1e-8;  // src/models/gaussian_process.jl:5
constexpr double MAX_NEG_VAR = 1e-8;      // src/models/gaussian_process.jl:13
constexpr double VAR_JITTER = 1e-18;      // AbstractGPs default FiniteGP observation noise

#ifdef __CUDACC__
template <int KID>
__device__ __forceinline__ double kappa(double d2) {
  if (KID == 0) {
    return exp(-0.5 * d2);
  } else if (KID == 1) {
    const double t = 1.7320508075688772 * sqrt(d2);
    return (1.0 + t) * exp(-t);
  } else {
    const double t = 2.23606797749979 * sqrt(d2);
    return (1.0 + t + (5.0 / 3.0) * d2) * exp(-t);
  }
}

// (d kappa / d r) / r as a function of d2 = r^2 : finite at r = 0 for all three kernels.
template <int KID>
__device__ __forceinline__ double kappa_dr_over_r(double d2) {
  if (KID == 0) {
    return -exp(-0.5 * d2);
  } else if (KID == 1) {
    return -3.0 * exp(-1.7320508075688772 * sqrt(d2));
  } else {
    const double t = 2.23606797749979 * sqrt(d2);
    return -(5.0 / 3.0) * (1.0 + t) * exp(-t);
  }
}

// Load point `p` (d raw coordinates), round flagged dims (ties-to-even like Julia's round), scale by 1/l.
template <int DP>
__device__ __forceinline__ void load_scaled_point(double (&x)[DP], const double *p, int d, const double *invl,
                                                  unsigned long long disc_bits, bool valid) {
#pragma unroll
  for (int i = 0; i < DP; ++i) {
    double v = 0.0;
    if (valid && i < d) {
      v = p[i];
      if ((disc_bits >> i) & 1ull) v = rint(v);
      v *= invl[i];
    }
    x[i] = v;
  }
}
#endif

}  // namespace boss
