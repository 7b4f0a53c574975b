// boss_b200.cu -- host runtime + C ABI of libboss_b200.so (see include/boss_b200.h).
//
// One process drives one B200.  All numerics run in the hand-written sm_100a kernels of
// cholesky.cuh / score.cuh on the library's stream; this file owns device memory (grow-only
// workspaces sized for 180 GB HBM3e), the fitted-GP handles (the factor cache) and the chunked
// candidate pipeline.  There is no CPU fallback: every entry point fails with BOSS_ERR_STATE /
// BOSS_ERR_CUDA if the device is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/boss_b200.h"
#include "cholesky.cuh"
#include "score.cuh"
#include "grad.cuh"
#include "small.cuh"
#include "append.cuh"
#include "multistart.cuh"

using namespace boss;

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
namespace {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <class T>
  T *as() const {
    return reinterpret_cast<T *>(p);
  }
};

constexpr int N_TIMERS = 4;
constexpr int LL_GROUPS = 16;  // max concurrent sub-batches of a log-likelihood window (separate streams)
constexpr int EV_POOL = 512;

struct Ctx {
  int device = -1;
  cudaStream_t stream = nullptr;
  std::mutex mu;
  std::string err;
  int64_t launches = 0;
  bool timing = false;
  // workspaces
  DevBuf ks, muv, sumsq, xs_stage, pm_stage, cm_stage, acq_stage, mu_stage, var_stage, st_stage, grad_stage;
  DevBuf blk_val, blk_idx, small, chol_L, chol_Winv, chol_misc, tt, vt, ut, dmu, dvar, pmg_stage;
  DevBuf part_mu, part_ss, part_gm, part_gv, cov_p, cov_stage, chol_W, chol_WT, ll_vec, ll_part, ms_buf;   // split-invariant partial sums (score.cuh / grad.cuh)
  // event pool for per-kernel-class timing
  cudaEvent_t ev_a[EV_POOL], ev_b[EV_POOL];
  int ev_class[EV_POOL];
  int ev_used = 0;
  bool ev_ready = false;
  double last_ms[N_TIMERS] = {0, 0, 0, 0};
  int last_cnt[N_TIMERS] = {0, 0, 0, 0};
  cudaEvent_t call_a = nullptr, call_b = nullptr;
  // handle-buffer pool: a BO loop refits GPs of the same size every iteration; cudaMalloc / cudaFree of the three
  // n_pad^2 factors cost more than the factorisation itself.  Freed buffers are kept (up to POOL_CAP bytes) by size.
  std::multimap<size_t, void *> pool;
  std::map<void *, size_t> pool_sizes;
  size_t pool_bytes = 0;
  cudaStream_t ll_stream[LL_GROUPS] = {};
  cudaEvent_t ll_fork = nullptr, ll_join[LL_GROUPS] = {};
};

Ctx g;

constexpr size_t POOL_CAP = (size_t)16 << 30;

cudaError_t pool_alloc(double **out, size_t bytes) {
  auto it = g.pool.find(bytes);
  if (it != g.pool.end()) {
    *out = reinterpret_cast<double *>(it->second);
    g.pool_bytes -= bytes;
    g.pool.erase(it);
    return cudaSuccess;
  }
  void *p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess && !g.pool.empty()) {   // out of memory: drop the pooled buffers and retry
    for (auto &kv : g.pool) {
      cudaFree(kv.second);
      g.pool_sizes.erase(kv.second);
    }
    g.pool.clear();
    g.pool_bytes = 0;
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
  }
  if (e == cudaSuccess) {
    g.pool_sizes[p] = bytes;
    *out = reinterpret_cast<double *>(p);
  }
  return e;
}
void pool_free(void *p) {
  if (!p) return;
  auto it = g.pool_sizes.find(p);
  if (it == g.pool_sizes.end() || g.pool_bytes + it->second > POOL_CAP || g.device < 0) {
    if (it != g.pool_sizes.end()) g.pool_sizes.erase(it);
    cudaFree(p);
    return;
  }
  g.pool.emplace(it->second, p);
  g.pool_bytes += it->second;
}
void pool_release_all() {
  for (auto &kv : g.pool) cudaFree(kv.second);
  g.pool.clear();
  g.pool_sizes.clear();
  g.pool_bytes = 0;
}

int fail(int code, const std::string &msg) {
  g.err = msg;
  return code;
}

#define CUDA_TRY(expr)                                                                                 \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess)                                                                             \
      return fail(BOSS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                                     std::to_string(__LINE__) + ")");                                  \
  } while (0)

// timing helpers: bracket one launch with a pooled event pair
struct Timed {
  int slot = -1;
  Timed(int cls) {
    if (g.timing && g.ev_used < EV_POOL) {
      slot = g.ev_used++;
      g.ev_class[slot] = cls;
      cudaEventRecord(g.ev_a[slot], g.stream);
    }
  }
  ~Timed() {
    if (slot >= 0) cudaEventRecord(g.ev_b[slot], g.stream);
  }
};

void timing_begin() {
  g.ev_used = 0;
  if (g.timing) cudaEventRecord(g.call_a, g.stream);
}
void timing_end() {  // call after the stream has been synchronised
  for (int c = 0; c < N_TIMERS; ++c) {
    g.last_ms[c] = 0;
    g.last_cnt[c] = 0;
  }
  if (!g.timing) return;
  cudaEventRecord(g.call_b, g.stream);
  cudaEventSynchronize(g.call_b);
  float ms = 0;
  cudaEventElapsedTime(&ms, g.call_a, g.call_b);
  g.last_ms[3] = ms;
  g.last_cnt[3] = 1;
  for (int i = 0; i < g.ev_used; ++i) {
    cudaEventElapsedTime(&ms, g.ev_a[i], g.ev_b[i]);
    g.last_ms[g.ev_class[i]] += ms;
    g.last_cnt[g.ev_class[i]] += 1;
  }
}

int pick_dp(int d) {
  if (d <= 2) return 2;
  if (d <= 4) return 4;
  if (d <= 6) return 6;
  if (d <= 8) return 8;
  if (d <= 12) return 12;
  if (d <= 16) return 16;
  if (d <= 32) return 32;
  return -1;
}

unsigned long long mask_bits(const uint8_t *mask, int d) {
  unsigned long long b = 0;
  if (mask)
    for (int i = 0; i < d; ++i)
      if (mask[i]) b |= (1ull << i);
  return b;
}

struct boss_gp_view {   // the fields append_kvec needs (boss_gp is defined after the anonymous namespace)
  int d, n, n_pad;
  const double *invl;
  unsigned long long disc;
  double a2;
  double *Xt;
};

// ---- kernel dispatch on (kernel_id, DP) ----
template <int KID, int DP>
void launch_build_k_t(const BuildKParams &p, dim3 grid) {
  build_k_kernel<KID, DP><<<grid, 256, 0, g.stream>>>(p);
}
template <int KID, int DP>
void launch_loglik_small_t(const SmallLoglikParams &p, int nblocks) {
  loglik_small_kernel<KID, DP><<<nblocks, SMALL_WARPS * 32, 0, g.stream>>>(p);
}
template <int KID, int DP>
void launch_append_kvec_t(const double *xnew, const boss_gp_view &v, double *kvec) {
  append_kvec_kernel<KID, DP><<<(v.n_pad + 255) / 256, 256, 0, g.stream>>>(xnew, v.d, v.n, v.n_pad, v.invl, v.disc, v.a2,
                                                                        v.Xt, kvec);
}
template <int KID, int DP>
void launch_llgrad_tile_t(const LlGradParams &p, dim3 grid) {
  loglik_grad_tile_kernel<KID, DP><<<grid, 256, 0, g.stream>>>(p);
}
template <int KID, int DP>
void launch_xcov_t(const XcovParams &p, dim3 grid) {
  xcov_kernel<KID, DP><<<grid, 256, 0, g.stream>>>(p);
}
template <int KID, int DP>
void launch_cov_finish_t(const CovFinishParams &p, dim3 grid) {
  cov_finish_kernel<KID, DP><<<grid, 256, 0, g.stream>>>(p);
}
template <int KID, int DP>
void launch_grad_t(const GradParams &p, dim3 grid) {
  grad_kernel<KID, DP><<<grid, 256, 0, g.stream>>>(p);
}
#define DISPATCH_KID_DP(FN, kid, dp, ...)                  \
  do {                                                     \
    switch ((kid) * 100 + (dp)) {                          \
      case 2: FN<0, 2>(__VA_ARGS__); break;                \
      case 4: FN<0, 4>(__VA_ARGS__); break;                \
      case 6: FN<0, 6>(__VA_ARGS__); break;                \
      case 8: FN<0, 8>(__VA_ARGS__); break;                \
      case 12: FN<0, 12>(__VA_ARGS__); break;              \
      case 16: FN<0, 16>(__VA_ARGS__); break;              \
      case 32: FN<0, 32>(__VA_ARGS__); break;              \
      case 102: FN<1, 2>(__VA_ARGS__); break;              \
      case 104: FN<1, 4>(__VA_ARGS__); break;              \
      case 106: FN<1, 6>(__VA_ARGS__); break;              \
      case 108: FN<1, 8>(__VA_ARGS__); break;              \
      case 112: FN<1, 12>(__VA_ARGS__); break;             \
      case 116: FN<1, 16>(__VA_ARGS__); break;             \
      case 132: FN<1, 32>(__VA_ARGS__); break;             \
      case 202: FN<2, 2>(__VA_ARGS__); break;              \
      case 204: FN<2, 4>(__VA_ARGS__); break;              \
      case 206: FN<2, 6>(__VA_ARGS__); break;              \
      case 208: FN<2, 8>(__VA_ARGS__); break;              \
      case 212: FN<2, 12>(__VA_ARGS__); break;             \
      case 216: FN<2, 16>(__VA_ARGS__); break;             \
      case 232: FN<2, 32>(__VA_ARGS__); break;             \
      default: break;                                      \
    }                                                      \
  } while (0)

// Number of zig-zag row-block splits per candidate block for the triangular products (score_trmm / wtv):
// minimise waves x (largest per-CTA share of the nblk(nblk+1)/2 block-steps + ~1 block-step of pipeline fill).
int pick_row_splits(int ncb, int nblk) {
  if (ncb >= 4 * 148 || nblk < 2) return 1;
  double best_t = 1e300;
  int best = 1;
  for (int ns = 1; ns <= nblk; ++ns) {
    int maxw = 0;
    for (int y = 0; y < ns; ++y) {
      int w = 0;
      for (int j = 0;; ++j) {
        const int i = (j >> 1) * 2 * ns + ((j & 1) ? 2 * ns - 1 - y : y);
        if (i >= nblk) break;
        w += i + 1;
      }
      maxw = std::max(maxw, w);
    }
    const long ctas = (long)ncb * ns, waves = (ctas + 147) / 148;
    const double t = (double)waves * (maxw + 1.0);
    if (t < best_t * (1.0 - 1e-9)) {
      best_t = t;
      best = ns;
    }
  }
  return best;
}

bool g_attr_done = false;
int set_kernel_attrs() {
  if (g_attr_done) return 0;
  CUDA_TRY(cudaFuncSetAttribute(chol_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(trtri_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(kinv_wtw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(chol_update_rl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(trtri_acc_rl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(trtri_row_rl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(score_trmm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(wtv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
  CUDA_TRY(cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PT_SMEM_BYTES));
  g_attr_done = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// batched blocked Cholesky driver (shared by fit and loglik)
// ---------------------------------------------------------------------------------------------
// On return L holds the factors, Winv the inverted diagonal blocks, logdet_blk / status are filled.
// W / WT non-null (S must be 1): also form the full triangular inverse.
// fwd non-null (left-looking only): the forward substitution w = L^-1 delta rides along (r_j in the diagonal-tile
// SYRK, w_j and |w_j|^2 in potrf_tile_kernel), so the log-likelihood needs no pass over L afterwards.
struct FwdInline {
  const double *ymm;   // delta of the first matrix; matrix s reads ymm + s * ldy
  long long ldy;
  int n, n_pad;
  double *r, *w;       // [S][n_pad] scratch / result
  double *ssq;         // [S][nblk]
};
int run_cholesky(double *L, size_t L_stride, double *Winv, size_t Winv_stride, int nblk, int ktiles, int S,
                 double *logdet_blk, int *status, double *W, double *WT, double *TT, size_t W_stride = 0,
                 size_t TT_stride = 0, bool single_fit = false, const FwdInline *fwd = nullptr) {
  CholGemmParams gp{};
  gp.L = L;
  gp.L_stride = L_stride;
  gp.Winv = Winv;
  gp.Winv_stride = Winv_stride;
  gp.nblk = nblk;
  gp.ktiles = ktiles;
  gp.W = W;
  gp.WT = WT;
  gp.TT = TT;
  gp.W_stride = W_stride;
  gp.TT_stride = TT_stride;
  PotrfParams pp{};
  pp.L = L;
  pp.L_stride = L_stride;
  pp.Winv = Winv;
  pp.Winv_stride = Winv_stride;
  pp.nblk = nblk;
  pp.ktiles = ktiles;
  pp.logdet_blk = logdet_blk;
  pp.status = status;
  pp.W = W;
  pp.WT = WT;
  pp.W_stride = W_stride;
  if (fwd) {
    gp.fwd_w = fwd->w;
    gp.fwd_r = fwd->r;
    gp.n_pad = fwd->n_pad;
    pp.fwd_ymm = fwd->ymm;
    pp.fwd_ldy = fwd->ldy;
    pp.fwd_n = fwd->n;
    pp.n_pad = fwd->n_pad;
    pp.fwd_r = fwd->r;
    pp.fwd_w = fwd->w;
    pp.fwd_ssq = fwd->ssq;
  }
  // Left-looking keeps accumulators in registers over the whole k-range (best when S x nblk CTAs fill the GPU);
  // a small batch (a single posterior fit) cannot fill 148 SMs that way -> right-looking steps of K = 128 tiles.
  // Only the posterior fit takes this path: batched log-likelihoods always run left-looking, so a sample's value
  // never depends on how the batch was split (over sub-batches, streams or GPUs).
  const bool right_looking = single_fit && (long long)S * nblk <= 148;
  for (int j = 0; j < nblk; ++j) {
    gp.j = j;
    pp.j = j;
    // left-looking, column j >= 1: SYRK of the diagonal tile -> potrf -> fused update + panel solve of the rows below
    const bool fused_panel = j > 0 && !right_looking;
    if (fused_panel) {
      Timed t(2);
      chol_update_kernel<<<dim3(1, S), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(gp);
      ++g.launches;
    }
    potrf_tile_kernel<<<S, 256, PT_SMEM_BYTES, g.stream>>>(pp);
    ++g.launches;
    if (j < nblk - 1) {
      Timed t(2);
      if (fused_panel)
        chol_panel_kernel<<<dim3(nblk - j - 1, S), GEMM_THREADS, PANEL_SMEM_BYTES, g.stream>>>(gp);
      else
        chol_trsm_kernel<<<dim3(nblk - j - 1, S), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(gp);
      ++g.launches;
      if (right_looking) {
        const int m = nblk - j - 1;
        chol_update_rl_kernel<<<dim3(m * (m + 1) / 2, S), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(gp);
        ++g.launches;
      }
    }
  }
  if (W && right_looking) {
    for (int k = 0; k + 1 < nblk; ++k) {
      Timed t(2);
      gp.j = k;
      trtri_acc_rl_kernel<<<dim3((nblk - k - 1) * (k + 1), S), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(gp);
      gp.j = k + 1;
      trtri_row_rl_kernel<<<dim3(k + 1, S), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(gp);
      g.launches += 2;
    }
  } else if (W) {
    for (int delta = 1; delta < nblk; ++delta) {
      gp.j = delta;
      Timed t(2);
      trtri_fused_kernel<<<dim3(nblk - delta, S), GEMM_THREADS, PANEL_SMEM_BYTES, g.stream>>>(gp);
      ++g.launches;
    }
  }
  CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct boss_gp {
  int n = 0, d = 0, n_pad = 0, nblk = 0, ktiles = 0, kernel_id = 0, dp = 0;
  unsigned long long disc = 0;
  double amp = 0, noise = 0, a2 = 0;
  double loglik = 0;
  double *W = nullptr, *WT = nullptr, *L = nullptr, *alpha = nullptr, *Xt = nullptr, *invl = nullptr;
  double *wvec = nullptr, *ymm = nullptr;   // w = L^-1 delta and delta = y - m(X), kept for boss_gp_append
  void free_dev() {
    for (double **q : {&W, &WT, &L, &alpha, &Xt, &invl, &wvec, &ymm}) {
      if (*q) pool_free(*q);
      *q = nullptr;
    }
  }
};

extern "C" {

int boss_version(void) { return 100; }
int boss_device(void) { return g.device; }
void *boss_stream(void) { return (void *)g.stream; }
const char *boss_last_error(void) { return g.err.c_str(); }
int64_t boss_launch_count(void) { return g.launches; }
double boss_last_kernel_ms(int which) { return (which >= 0 && which < N_TIMERS) ? g.last_ms[which] : 0.0; }
int boss_last_kernel_count(int which) { return (which >= 0 && which < N_TIMERS) ? g.last_cnt[which] : 0; }
void boss_set_timing(int on) { g.timing = on != 0; }

int boss_init(int device) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.device == device && g.stream) return 0;
  if (g.device >= 0 && g.device != device) return fail(BOSS_ERR_STATE, "boss_init: already initialised on another device");
  int count = 0;
  CUDA_TRY(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) return fail(BOSS_ERR_ARG, "boss_init: no such CUDA device");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(BOSS_ERR_STATE, std::string("boss_init: built for sm_100a, device is ") + prop.name);
  CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  for (int i = 0; i < EV_POOL; ++i) {
    CUDA_TRY(cudaEventCreate(&g.ev_a[i]));
    CUDA_TRY(cudaEventCreate(&g.ev_b[i]));
  }
  CUDA_TRY(cudaEventCreate(&g.call_a));
  CUDA_TRY(cudaEventCreate(&g.call_b));
  CUDA_TRY(cudaEventCreateWithFlags(&g.ll_fork, cudaEventDisableTiming));
  for (int i = 0; i < LL_GROUPS; ++i) {
    CUDA_TRY(cudaStreamCreateWithFlags(&g.ll_stream[i], cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&g.ll_join[i], cudaEventDisableTiming));
  }
  g.ev_ready = true;
  g.device = device;
  return set_kernel_attrs();
}

void boss_shutdown(void) {
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.device < 0) return;
  cudaSetDevice(g.device);
  cudaStreamSynchronize(g.stream);
  for (DevBuf *b : {&g.ks, &g.muv, &g.sumsq, &g.xs_stage, &g.pm_stage, &g.cm_stage, &g.acq_stage, &g.mu_stage,
                    &g.var_stage, &g.st_stage, &g.grad_stage, &g.blk_val, &g.blk_idx, &g.small, &g.chol_L,
                    &g.chol_Winv, &g.chol_misc, &g.tt, &g.vt, &g.ut, &g.dmu, &g.dvar, &g.pmg_stage, &g.part_mu,
                    &g.part_ss, &g.part_gm, &g.part_gv, &g.cov_p, &g.cov_stage, &g.chol_W, &g.chol_WT, &g.ll_vec,
                    &g.ll_part, &g.ms_buf})
    b->release();
  pool_release_all();
  if (g.ev_ready) {
    for (int i = 0; i < EV_POOL; ++i) {
      cudaEventDestroy(g.ev_a[i]);
      cudaEventDestroy(g.ev_b[i]);
    }
    cudaEventDestroy(g.call_a);
    cudaEventDestroy(g.call_b);
    cudaEventDestroy(g.ll_fork);
    for (int i = 0; i < LL_GROUPS; ++i) {
      cudaStreamDestroy(g.ll_stream[i]);
      cudaEventDestroy(g.ll_join[i]);
    }
    g.ev_ready = false;
  }
  cudaStreamDestroy(g.stream);
  g.stream = nullptr;
  g.device = -1;
}

#define REQUIRE_INIT()                                                                  \
  if (g.device < 0) return fail(BOSS_ERR_STATE, "boss_init() has not been called");    \
  CUDA_TRY(cudaSetDevice(g.device))

// ---------------------------------------------------------------------------------------------
// fit
// ---------------------------------------------------------------------------------------------
int boss_gp_fit(const double *X, int d, int n, const double *y_minus_mean, const double *lengthscales,
                double amplitude, double noise_std, int kernel_id, const uint8_t *discrete_mask, boss_gp **out,
                double *loglik_out) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!out) return fail(BOSS_ERR_ARG, "boss_gp_fit: out is NULL");
  *out = nullptr;
  if (!X || !y_minus_mean || !lengthscales || d < 1 || n < 1) return fail(BOSS_ERR_ARG, "boss_gp_fit: bad arguments");
  if (kernel_id < 0 || kernel_id > 2) return fail(BOSS_ERR_ARG, "boss_gp_fit: unknown kernel_id");
  const int dp = pick_dp(d);
  if (dp < 0) return fail(BOSS_ERR_ARG, "boss_gp_fit: x_dim > 32 is not supported");
  // reference asserts (src/models/gaussian_process.jl:227-229)
  for (int i = 0; i < d; ++i)
    if (!(lengthscales[i] >= 0)) return fail(BOSS_ERR_ARG, "boss_gp_fit: negative lengthscale");
  if (!(amplitude >= 0) || !(noise_std >= 0)) return fail(BOSS_ERR_ARG, "boss_gp_fit: negative amplitude / noise_std");

  boss_gp *h = new boss_gp();
  h->n = n;
  h->d = d;
  h->n_pad = round_up(n, TM);
  h->nblk = h->n_pad / TM;
  h->ktiles = h->n_pad / TK;
  h->kernel_id = kernel_id;
  h->dp = dp;
  h->disc = mask_bits(discrete_mask, d);
  h->amp = amplitude + MIN_PARAM_VALUE;
  h->noise = noise_std + MIN_PARAM_VALUE;
  h->a2 = h->amp * h->amp;
  const size_t npad = h->n_pad, mat = npad * npad;
  const int nblk = h->nblk;

  auto bail = [&](int code) {
    h->free_dev();
    delete h;
    return code;
  };
#define FIT_TRY(expr)                                                                                         \
  do {                                                                                                        \
    cudaError_t _e = (expr);                                                                                  \
    if (_e != cudaSuccess)                                                                                    \
      return bail(fail(BOSS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (line " +       \
                                          std::to_string(__LINE__) + ")"));                                   \
  } while (0)

  FIT_TRY(pool_alloc(&h->L, mat * 8));
  FIT_TRY(pool_alloc(&h->W, mat * 8));
  FIT_TRY(pool_alloc(&h->WT, mat * 8));
  FIT_TRY(pool_alloc(&h->alpha, npad * 8));
  FIT_TRY(pool_alloc(&h->Xt, npad * dp * 8));
  FIT_TRY(pool_alloc(&h->invl, dp * 8));
  FIT_TRY(pool_alloc(&h->wvec, npad * 8));
  FIT_TRY(pool_alloc(&h->ymm, npad * 8));
  FIT_TRY(cudaMemsetAsync(h->L, 0, mat * 8, g.stream));
  FIT_TRY(cudaMemsetAsync(h->W, 0, mat * 8, g.stream));
  FIT_TRY(cudaMemsetAsync(h->WT, 0, mat * 8, g.stream));

  // misc device scratch: X | ymm_pad | ls | amp | noise | invl(host computed) | logdet_blk | w | status
  const size_t off_X = 0, off_y = off_X + (size_t)n * d, off_ls = off_y + npad, off_amp = off_ls + d,
               off_noise = off_amp + 1, off_ld = off_noise + 1, off_w = off_ld + nblk, off_st = off_w + npad,
               total = off_st + 2;
  FIT_TRY(g.chol_misc.ensure(total * 8));
  double *misc = g.chol_misc.as<double>();
  FIT_TRY(cudaMemsetAsync(misc, 0, total * 8, g.stream));
  FIT_TRY(cudaMemcpyAsync(misc + off_X, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, g.stream));
  FIT_TRY(cudaMemcpyAsync(misc + off_y, y_minus_mean, (size_t)n * 8, cudaMemcpyHostToDevice, g.stream));
  FIT_TRY(cudaMemcpyAsync(misc + off_ls, lengthscales, (size_t)d * 8, cudaMemcpyHostToDevice, g.stream));
  FIT_TRY(cudaMemcpyAsync(misc + off_amp, &amplitude, 8, cudaMemcpyHostToDevice, g.stream));
  FIT_TRY(cudaMemcpyAsync(misc + off_noise, &noise_std, 8, cudaMemcpyHostToDevice, g.stream));
  std::vector<double> invl(dp, 0.0);
  for (int i = 0; i < d; ++i) invl[i] = 1.0 / (lengthscales[i] + MIN_PARAM_VALUE);
  FIT_TRY(cudaMemcpyAsync(h->invl, invl.data(), dp * 8, cudaMemcpyHostToDevice, g.stream));
  int *status = reinterpret_cast<int *>(misc + off_st);

  FIT_TRY(g.chol_Winv.ensure((size_t)nblk * TM * TM * 8));

  timing_begin();
  BuildKParams bk{};
  bk.X = misc + off_X;
  bk.d = d;
  bk.n = n;
  bk.nblk = nblk;
  bk.ktiles = h->ktiles;
  bk.ls = misc + off_ls;
  bk.amp = misc + off_amp;
  bk.noise = misc + off_noise;
  bk.disc_bits = h->disc;
  bk.K = h->L;
  bk.K_stride = mat;
  bk.status = status;
  {
    Timed t(1);
    DISPATCH_KID_DP(launch_build_k_t, kernel_id, dp, bk, dim3(nblk * (nblk + 1) / 2, 1));
    ++g.launches;
  }
  int rc = run_cholesky(h->L, mat, g.chol_Winv.as<double>(), (size_t)nblk * TM * TM, nblk, h->ktiles, 1,
                        misc + off_ld, status, h->W, h->WT, nullptr, 0, 0, true);
  if (rc) return bail(rc);
  // w = W delta ; alpha = W^T w
  matvec_p_kernel<<<h->n_pad / 64, 256, 0, g.stream>>>(h->W, misc + off_y, misc + off_w, h->ktiles);
  matvec_p_kernel<<<h->n_pad / 64, 256, 0, g.stream>>>(h->WT, misc + off_w, h->alpha, h->ktiles);
  {
    const int tot = h->n_pad * dp;
    scale_train_kernel<<<(tot + 255) / 256, 256, 0, g.stream>>>(misc + off_X, d, n, h->n_pad, dp, h->invl, h->disc, h->Xt);
  }
  g.launches += 3;
  FIT_TRY(cudaGetLastError());
  FIT_TRY(cudaMemcpyAsync(h->wvec, misc + off_w, npad * 8, cudaMemcpyDeviceToDevice, g.stream));
  FIT_TRY(cudaMemcpyAsync(h->ymm, misc + off_y, npad * 8, cudaMemcpyDeviceToDevice, g.stream));
  std::vector<double> host(nblk + npad + 2);
  FIT_TRY(cudaMemcpyAsync(host.data(), misc + off_ld, (nblk + npad + 2) * 8, cudaMemcpyDeviceToHost, g.stream));
  FIT_TRY(cudaStreamSynchronize(g.stream));
  timing_end();
  int st;
  std::memcpy(&st, &host[nblk + npad], sizeof(int));
  if (st != 0) {
    bail(0);
    if (loglik_out) *loglik_out = -std::numeric_limits<double>::infinity();
    g.err = "boss_gp_fit: kernel matrix is not positive definite";
    return BOSS_NOT_POSDEF;
  }
  double ld = 0.0, mahal = 0.0;
  for (int b = 0; b < nblk; ++b) ld += host[b];
  for (size_t k = 0; k < npad; ++k) mahal = std::fma(host[nblk + k], host[nblk + k], mahal);
  h->loglik = -((double)n * 1.8378770664093453 + 2.0 * ld + mahal) * 0.5;
  if (loglik_out) *loglik_out = h->loglik;
  *out = h;
  return 0;
#undef FIT_TRY
}

void boss_gp_free(boss_gp *gp) {
  if (!gp) return;
  std::lock_guard<std::mutex> lk(g.mu);
  if (g.device >= 0) cudaSetDevice(g.device);
  gp->free_dev();
  delete gp;
}
// Grow a handle's capacity by one 128-block (P-layout stride changes -> repack on the device).
static int grow_handle(boss_gp *h) {
  const int npad_new = h->n_pad + TM, kt_new = npad_new / TK;
  const size_t mat_new = (size_t)npad_new * npad_new;
  double *nL = nullptr, *nW = nullptr, *nWT = nullptr, *nal = nullptr, *nXt = nullptr, *nw = nullptr, *ny = nullptr;
  auto cleanup = [&]() {
    for (double *q : {nL, nW, nWT, nal, nXt, nw, ny})
      if (q) pool_free(q);
  };
#define GROW_TRY(expr)                                                                                       \
  do {                                                                                                       \
    cudaError_t _e = (expr);                                                                                 \
    if (_e != cudaSuccess) {                                                                                 \
      cleanup();                                                                                             \
      return fail(BOSS_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));                        \
    }                                                                                                        \
  } while (0)
  GROW_TRY(pool_alloc(&nL, mat_new * 8));
  GROW_TRY(pool_alloc(&nW, mat_new * 8));
  GROW_TRY(pool_alloc(&nWT, mat_new * 8));
  GROW_TRY(pool_alloc(&nal, (size_t)npad_new * 8));
  GROW_TRY(pool_alloc(&nXt, (size_t)npad_new * h->dp * 8));
  GROW_TRY(pool_alloc(&nw, (size_t)npad_new * 8));
  GROW_TRY(pool_alloc(&ny, (size_t)npad_new * 8));
  const unsigned nb = (unsigned)((mat_new + 255) / 256);
  repack_grow_kernel<<<nb, 256, 0, g.stream>>>(h->L, h->ktiles, h->n_pad, nL, kt_new, npad_new);
  repack_grow_kernel<<<nb, 256, 0, g.stream>>>(h->W, h->ktiles, h->n_pad, nW, kt_new, npad_new);
  repack_grow_kernel<<<nb, 256, 0, g.stream>>>(h->WT, h->ktiles, h->n_pad, nWT, kt_new, npad_new);
  g.launches += 3;
  struct {
    double *dst, *src;
    size_t n_new, n_old;
  } vecs[4] = {{nal, h->alpha, (size_t)npad_new, (size_t)h->n_pad},
               {nXt, h->Xt, (size_t)npad_new * h->dp, (size_t)h->n_pad * h->dp},
               {nw, h->wvec, (size_t)npad_new, (size_t)h->n_pad},
               {ny, h->ymm, (size_t)npad_new, (size_t)h->n_pad}};
  for (auto &v : vecs) {
    GROW_TRY(cudaMemsetAsync(v.dst, 0, v.n_new * 8, g.stream));
    GROW_TRY(cudaMemcpyAsync(v.dst, v.src, v.n_old * 8, cudaMemcpyDeviceToDevice, g.stream));
  }
  GROW_TRY(cudaGetLastError());
  GROW_TRY(cudaStreamSynchronize(g.stream));
#undef GROW_TRY
  for (double *q : {h->L, h->W, h->WT, h->alpha, h->Xt, h->wvec, h->ymm}) pool_free(q);
  h->L = nL; h->W = nW; h->WT = nWT; h->alpha = nal; h->Xt = nXt; h->wvec = nw; h->ymm = ny;
  h->n_pad = npad_new;
  h->nblk = npad_new / TM;
  h->ktiles = kt_new;
  return 0;
}

int boss_gp_append(boss_gp *h, const double *x_new, double y_minus_mean_new, double *loglik_out) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!h || !x_new) return fail(BOSS_ERR_ARG, "boss_gp_append: bad arguments");
  if (h->n == h->n_pad) {
    int rc = grow_handle(h);
    if (rc) return rc;
  }
  const int n = h->n, npad = h->n_pad, d = h->d;
  // workspace: xnew[32] | kvec[npad] | l[npad] | u[npad] | sc[4] | status
  const size_t o_k = 32, o_l = o_k + npad, o_u = o_l + npad, o_sc = o_u + npad, o_st = o_sc + 4, total = o_st + 1;
  CUDA_TRY(g.chol_misc.ensure(total * 8));
  double *ws = g.chol_misc.as<double>();
  int *status = reinterpret_cast<int *>(ws + o_st);
  CUDA_TRY(cudaMemsetAsync(status, 0, 8, g.stream));
  CUDA_TRY(cudaMemcpyAsync(ws, x_new, (size_t)d * 8, cudaMemcpyHostToDevice, g.stream));
  boss_gp_view v{d, n, npad, h->invl, h->disc, h->a2, h->Xt};
  DISPATCH_KID_DP(launch_append_kvec_t, h->kernel_id, h->dp, ws, v, ws + o_k);
  matvec_p_kernel<<<npad / 64, 256, 0, g.stream>>>(h->W, ws + o_k, ws + o_l, h->ktiles);
  append_scalars_kernel<<<1, 256, 0, g.stream>>>(ws + o_l, h->wvec, n, h->a2 + h->noise * h->noise, y_minus_mean_new,
                                                 ws + o_sc, status);
  matvec_p_kernel<<<npad / 64, 256, 0, g.stream>>>(h->WT, ws + o_l, ws + o_u, h->ktiles);
  append_scatter_kernel<<<(n + 256) / 256, 256, 0, g.stream>>>(h->L, h->W, h->WT, n, h->ktiles, ws + o_l, ws + o_u,
                                                              ws + o_sc, status, h->wvec, h->ymm, y_minus_mean_new);
  g.launches += 5;
  CUDA_TRY(cudaGetLastError());
  double hs[5];
  CUDA_TRY(cudaMemcpyAsync(hs, ws + o_sc, 40, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  int st;
  std::memcpy(&st, &hs[4], sizeof(int));
  if (st != 0) {
    if (loglik_out) *loglik_out = -std::numeric_limits<double>::infinity();
    g.err = "boss_gp_append: the extended kernel matrix is not positive definite (handle left unchanged)";
    return BOSS_NOT_POSDEF;
  }
  h->n = n + 1;
  matvec_p_kernel<<<npad / 64, 256, 0, g.stream>>>(h->WT, h->wvec, h->alpha, h->ktiles);   // alpha = W^T w
  ++g.launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  h->loglik += -0.5 * 1.8378770664093453 - std::log(hs[0]) - 0.5 * hs[1] * hs[1];
  if (loglik_out) *loglik_out = h->loglik;
  return 0;
}

int boss_gp_n(const boss_gp *gp) { return gp ? gp->n : -1; }
int boss_gp_d(const boss_gp *gp) { return gp ? gp->d : -1; }

// ---------------------------------------------------------------------------------------------
// scoring core (shared by predict / ei_score, host- and device-pointer variants)
// ---------------------------------------------------------------------------------------------
namespace {

struct ScoreArgs {
  const boss_gp *const *slices;
  int y_dim, n_samples;
  const double *Xs;  // host or device
  long long M;
  const double *prior_mean;  // y_dim x M, host or device
  const double *coefs, *best, *y_max, *lb, *ub;  // host
  const uint8_t *cons_mask;                      // host or device
  double *acq, *grad;                            // host or device
  const double *prior_mean_grad;                 // y_dim x d x M (host or device) or null
  double *mu_out, *var_out;                      // predict mode (single slice), host or device
  int32_t *status_out;
  double *best_val;  // host
  int64_t *best_idx; // host
  bool dev;          // arrays are device pointers
  bool want_argmax;
  int any_fail;      // out
  const CandGen *gen = nullptr;   // candidates generated on the device (Xs == NULL); other arrays are host pointers
};

int score_core(ScoreArgs &a) {
  const int nsl = a.y_dim * a.n_samples;
  if (nsl < 1 || a.y_dim > MAX_YDIM) return fail(BOSS_ERR_ARG, "score: y_dim must be in 1..16");
  const boss_gp *g0 = a.slices[0];
  for (int q = 0; q < nsl; ++q) {
    if (!a.slices[q]) return fail(BOSS_ERR_ARG, "score: NULL slice handle");
    if (a.slices[q]->d != g0->d) return fail(BOSS_ERR_ARG, "score: slices disagree on x_dim");
  }
  const int d = g0->d;
  int max_npad = 0;
  for (int q = 0; q < nsl; ++q) max_npad = std::max(max_npad, a.slices[q]->n_pad);
  a.any_fail = 0;
  if (a.M <= 0) {
    if (a.best_idx) *a.best_idx = -1;
    return 0;
  }
  // chunk size: K*^T scratch budget 4 GiB, at most 4 waves of 148 CTAs
  long long ncb_max = (4ll << 30) / (1024ll * max_npad);
  ncb_max = std::min<long long>(ncb_max, 592);
  if (ncb_max >= 148) ncb_max = ncb_max / 148 * 148;
  if (ncb_max < 1) return fail(BOSS_ERR_ARG, "score: n too large for the scratch budget");
  if (a.grad) ncb_max = std::max<long long>(1, std::min<long long>(ncb_max, ncb_max >= 296 ? 296 : ncb_max));
  const long long need_cb = (a.M + 127) / 128;
  const int ncb_cap = (int)std::min<long long>(ncb_max, need_cb);
  const int CH = ncb_cap * 128;

  CUDA_TRY(g.ks.ensure((size_t)CH * max_npad * 8));
  CUDA_TRY(g.muv.ensure((size_t)nsl * CH * 8));
  CUDA_TRY(g.sumsq.ensure((size_t)nsl * CH * 8));
  const int max_nblk = max_npad / TM;
  CUDA_TRY(g.part_mu.ensure((size_t)2 * max_nblk * CH * 8));
  CUDA_TRY(g.part_ss.ensure((size_t)2 * max_nblk * CH * 8));
  if (a.grad) {
    if (d > 32) return fail(BOSS_ERR_ARG, "score: gradients need x_dim <= 32");
    CUDA_TRY(g.vt.ensure((size_t)CH * max_npad * 8));
    CUDA_TRY(g.ut.ensure((size_t)CH * max_npad * 8));
    CUDA_TRY(g.dmu.ensure((size_t)nsl * d * CH * 8));
    CUDA_TRY(g.dvar.ensure((size_t)nsl * d * CH * 8));
    CUDA_TRY(g.part_gm.ensure((size_t)2 * max_nblk * d * CH * 8));
    CUDA_TRY(g.part_gv.ensure((size_t)2 * max_nblk * d * CH * 8));
  }
  const int nblk_acq_max = (CH + 255) / 256;
  CUDA_TRY(g.blk_val.ensure((size_t)nblk_acq_max * 8));
  CUDA_TRY(g.blk_idx.ensure((size_t)nblk_acq_max * 8));
  // small: a2[nsl] | lb[d] | ub[d] | best(1) | bidx(1 as int64) | any_fail(int)
  const size_t sm_n = (size_t)nsl + 2 * d + 4;
  CUDA_TRY(g.small.ensure(sm_n * 8));
  double *small = g.small.as<double>();
  std::vector<double> hs(sm_n, 0.0);
  for (int q = 0; q < nsl; ++q) hs[q] = a.slices[q]->a2;
  if (a.lb && a.ub) {
    for (int i = 0; i < d; ++i) {
      hs[nsl + i] = a.lb[i];
      hs[nsl + d + i] = a.ub[i];
    }
  }
  long long neg1 = -1;
  std::memcpy(&hs[nsl + 2 * d + 1], &neg1, 8);
  CUDA_TRY(cudaMemcpyAsync(small, hs.data(), sm_n * 8, cudaMemcpyHostToDevice, g.stream));
  double *d_best = small + nsl + 2 * d;
  long long *d_bidx = reinterpret_cast<long long *>(small + nsl + 2 * d + 1);
  int *d_anyfail = reinterpret_cast<int *>(small + nsl + 2 * d + 2);

  if (!a.dev) {
    CUDA_TRY(g.xs_stage.ensure((size_t)CH * d * 8));
    if (a.prior_mean) CUDA_TRY(g.pm_stage.ensure((size_t)CH * a.y_dim * 8));
    if (a.cons_mask) CUDA_TRY(g.cm_stage.ensure((size_t)CH));
    if (a.acq) CUDA_TRY(g.acq_stage.ensure((size_t)CH * 8));
    if (a.mu_out) CUDA_TRY(g.mu_stage.ensure((size_t)CH * 8));
    if (a.var_out) CUDA_TRY(g.var_stage.ensure((size_t)CH * 8));
    if (a.status_out) CUDA_TRY(g.st_stage.ensure((size_t)CH * 4));
    if (a.grad) CUDA_TRY(g.grad_stage.ensure((size_t)CH * d * 8));
    if (a.grad && a.prior_mean_grad) CUDA_TRY(g.pmg_stage.ensure((size_t)CH * d * a.y_dim * 8));
  }

  timing_begin();
  for (long long m0 = 0; m0 < a.M; m0 += CH) {
    const int ch = (int)std::min<long long>(CH, a.M - m0);
    const int ncb = (ch + 127) / 128;
    const double *xs_dev;
    const double *pm_dev = nullptr;
    const unsigned char *cm_dev = nullptr;
    long long in_off, out_off;
    double *acq_dev = nullptr, *mu_dev = nullptr, *var_dev = nullptr, *grad_dev = nullptr;
    const double *pmg_dev = nullptr;
    int *st_dev = nullptr;
    if (a.dev) {
      xs_dev = a.Xs;
      pm_dev = a.prior_mean;
      cm_dev = a.cons_mask;
      in_off = 0;
      out_off = 0;
      acq_dev = a.acq;
      mu_dev = a.mu_out;
      var_dev = a.var_out;
      st_dev = a.status_out;
      grad_dev = a.grad;
      pmg_dev = a.prior_mean_grad;
    } else {
      if (a.gen) {
        gen_candidates_kernel<<<(ch * d + 255) / 256, 256, 0, g.stream>>>(*a.gen, m0, ch, g.xs_stage.as<double>());
        ++g.launches;
      } else {
        CUDA_TRY(cudaMemcpyAsync(g.xs_stage.p, a.Xs + (size_t)m0 * d, (size_t)ch * d * 8, cudaMemcpyHostToDevice, g.stream));
      }
      xs_dev = g.xs_stage.as<double>();
      if (a.prior_mean) {
        CUDA_TRY(cudaMemcpyAsync(g.pm_stage.p, a.prior_mean + (size_t)m0 * a.y_dim, (size_t)ch * a.y_dim * 8,
                                 cudaMemcpyHostToDevice, g.stream));
        pm_dev = g.pm_stage.as<double>();
      }
      if (a.cons_mask) {
        CUDA_TRY(cudaMemcpyAsync(g.cm_stage.p, a.cons_mask + m0, (size_t)ch, cudaMemcpyHostToDevice, g.stream));
        cm_dev = g.cm_stage.as<unsigned char>();
      }
      in_off = m0;
      out_off = m0;
      if (a.acq) acq_dev = g.acq_stage.as<double>();
      if (a.mu_out) mu_dev = g.mu_stage.as<double>();
      if (a.var_out) var_dev = g.var_stage.as<double>();
      if (a.status_out) st_dev = g.st_stage.as<int>();
      if (a.grad) grad_dev = g.grad_stage.as<double>();
      if (a.grad && a.prior_mean_grad) {
        CUDA_TRY(cudaMemcpyAsync(g.pmg_stage.p, a.prior_mean_grad + (size_t)m0 * d * a.y_dim,
                                 (size_t)ch * d * a.y_dim * 8, cudaMemcpyHostToDevice, g.stream));
        pmg_dev = g.pmg_stage.as<double>();
      }
    }
    for (int q = 0; q < nsl; ++q) {
      const boss_gp *h = a.slices[q];
      // Small candidate batches (multi-start optimiser iterations, single-point calls) cannot fill 148 SMs
      // with one CTA per 128 candidates: deal the training chunks / W row blocks of a candidate block to
      // several CTAs.  Partial sums are kept per chunk / per row block, so results do not depend on the split.
      const int want = (16 * 148 + ncb - 1) / ncb;   // >= 16 CTAs per SM in flight over the launch (4 waves of 4)
      const int nks = std::max(1, std::min(h->nblk, want));            // xcov / grad: splits over training chunks
      const int nsp = pick_row_splits(ncb, h->nblk);                   // trmm / wtv: zig-zag row-block splits
      const int cnt = ncb * 128;
      XcovParams xp{};
      xp.Xs = xs_dev;
      xp.M = a.M;
      xp.m0 = m0;
      xp.in_off = in_off;
      xp.d = d;
      xp.n = h->n;
      xp.n_pad = h->n_pad;
      xp.ktiles = h->ktiles;
      xp.Xt = h->Xt;
      xp.invl = h->invl;
      xp.disc_bits = h->disc;
      xp.alpha = h->alpha;
      xp.a2 = h->a2;
      xp.Ks = g.ks.as<double>();
      xp.mu_part = g.part_mu.as<double>();
      xp.ld = CH;
      {
        Timed t(1);
        DISPATCH_KID_DP(launch_xcov_t, h->kernel_id, h->dp, xp, dim3(ncb, nks));
      }
      reduce_rows_kernel<<<(cnt + 255) / 256, 256, 0, g.stream>>>(g.part_mu.as<double>(), 2 * h->nblk, (size_t)CH,
                                                                 g.muv.as<double>() + (size_t)q * CH, cnt);
      ScoreParams sp{};
      sp.W = h->W;
      sp.Ks = g.ks.as<double>();
      sp.nblk = h->nblk;
      sp.ktiles = h->ktiles;
      sp.ss_part = g.part_ss.as<double>();
      sp.ld = CH;
      sp.VT = a.grad ? g.vt.as<double>() : nullptr;
      {
        Timed t(0);
        score_trmm_kernel<<<dim3(ncb, nsp), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(sp);
      }
      reduce_rows_kernel<<<(cnt + 255) / 256, 256, 0, g.stream>>>(g.part_ss.as<double>(), 2 * h->nblk, (size_t)CH,
                                                                 g.sumsq.as<double>() + (size_t)q * CH, cnt);
      g.launches += 4;
      if (a.grad) {
        WtvParams wp{h->WT, g.vt.as<double>(), g.ut.as<double>(), h->nblk, h->ktiles};
        {
          Timed t(0);
          wtv_kernel<<<dim3(ncb, nsp), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(wp);
        }
        GradParams gq{};
        gq.Xs = xs_dev;
        gq.M = a.M;
        gq.m0 = m0;
        gq.in_off = in_off;
        gq.d = d;
        gq.n = h->n;
        gq.n_pad = h->n_pad;
        gq.ktiles = h->ktiles;
        gq.chunk_ld = CH;
        gq.Xt = h->Xt;
        gq.invl = h->invl;
        gq.alpha = h->alpha;
        gq.disc_bits = h->disc;
        gq.a2 = h->a2;
        gq.UT = g.ut.as<double>();
        gq.gm_part = g.part_gm.as<double>();
        gq.gv_part = g.part_gv.as<double>();
        {
          Timed t(1);
          DISPATCH_KID_DP(launch_grad_t, h->kernel_id, h->dp, gq, dim3(ncb, nks));
        }
        GradReduceParams gr{g.part_gm.as<double>(), g.part_gv.as<double>(), 2 * h->nblk, d, CH, cnt, h->invl, h->disc,
                            g.dmu.as<double>() + (size_t)q * d * CH, g.dvar.as<double>() + (size_t)q * d * CH};
        grad_reduce_kernel<<<dim3((cnt + 255) / 256, d), 256, 0, g.stream>>>(gr);
        g.launches += 3;
      }
    }
    AcqParams ap{};
    ap.y_dim = a.y_dim;
    ap.n_samples = a.n_samples;
    ap.d = d;
    ap.M = a.M;
    ap.m0 = m0;
    ap.in_off = in_off;
    ap.out_off = out_off;
    ap.chunk = ch;
    ap.chunk_ld = CH;
    ap.mu = g.muv.as<double>();
    ap.sumsq = g.sumsq.as<double>();
    ap.a2 = small;
    ap.prior_mean = pm_dev;
    for (int i = 0; i < a.y_dim; ++i) {
      ap.coefs[i] = a.coefs ? a.coefs[i] : (i == 0 ? 1.0 : 0.0);
      ap.y_max[i] = a.y_max ? a.y_max[i] : INFINITY;
    }
    ap.has_best = a.best != nullptr;
    ap.best = a.best ? *a.best : 0.0;
    ap.has_ymax = a.y_max != nullptr;
    ap.Xs = xs_dev;
    ap.lb = (a.lb && a.ub) ? small + nsl : nullptr;
    ap.ub = (a.lb && a.ub) ? small + nsl + d : nullptr;
    ap.cons_mask = cm_dev;
    ap.acq = acq_dev;
    ap.mu_out = mu_dev;
    ap.var_out = var_dev;
    ap.status_out = st_dev;
    ap.any_fail = d_anyfail;
    ap.blk_val = a.want_argmax ? g.blk_val.as<double>() : nullptr;
    ap.blk_idx = a.want_argmax ? g.blk_idx.as<long long>() : nullptr;
    const int nb = (ch + 255) / 256;
    acq_kernel<<<nb, 256, 0, g.stream>>>(ap);
    ++g.launches;
    if (a.want_argmax) {
      argmax_final_kernel<<<1, 256, 0, g.stream>>>(g.blk_val.as<double>(), g.blk_idx.as<long long>(), nb, d_best, d_bidx);
      ++g.launches;
    }
    if (a.grad) {
      AcqGradParams gp2{ap, g.dmu.as<double>(), g.dvar.as<double>(), pmg_dev, grad_dev};
      acq_grad_kernel<<<(ch + 127) / 128, 128, 0, g.stream>>>(gp2);
      ++g.launches;
      if (!a.dev)
        CUDA_TRY(cudaMemcpyAsync(a.grad + (size_t)m0 * d, grad_dev, (size_t)ch * d * 8, cudaMemcpyDeviceToHost, g.stream));
    }
    if (!a.dev) {
      if (a.acq) CUDA_TRY(cudaMemcpyAsync(a.acq + m0, acq_dev, (size_t)ch * 8, cudaMemcpyDeviceToHost, g.stream));
      if (a.mu_out) CUDA_TRY(cudaMemcpyAsync(a.mu_out + m0, mu_dev, (size_t)ch * 8, cudaMemcpyDeviceToHost, g.stream));
      if (a.var_out) CUDA_TRY(cudaMemcpyAsync(a.var_out + m0, var_dev, (size_t)ch * 8, cudaMemcpyDeviceToHost, g.stream));
      if (a.status_out)
        CUDA_TRY(cudaMemcpyAsync(a.status_out + m0, st_dev, (size_t)ch * 4, cudaMemcpyDeviceToHost, g.stream));
    }
  }
  CUDA_TRY(cudaGetLastError());
  double hres[3];
  CUDA_TRY(cudaMemcpyAsync(hres, d_best, 24, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  timing_end();
  long long bidx;
  std::memcpy(&bidx, &hres[1], 8);
  int af;
  std::memcpy(&af, &hres[2], 4);
  a.any_fail = af;
  if (a.want_argmax) {
    if (a.best_val) *a.best_val = hres[0];
    if (a.best_idx) *a.best_idx = bidx;
  }
  return 0;
}

}  // namespace

int boss_gp_predict(const boss_gp *gp, const double *Xs, int64_t M, const double *prior_mean_s, double *mu,
                    double *var, int32_t *status) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!gp || (!Xs && M > 0)) return fail(BOSS_ERR_ARG, "boss_gp_predict: bad arguments");
  const boss_gp *sl[1] = {gp};
  ScoreArgs a{};
  a.slices = sl;
  a.y_dim = 1;
  a.n_samples = 1;
  a.Xs = Xs;
  a.M = M;
  a.prior_mean = prior_mean_s;
  a.mu_out = mu;
  a.var_out = var;
  a.status_out = status;
  a.dev = false;
  a.want_argmax = false;
  int rc = score_core(a);
  if (rc) return rc;
  return a.any_fail ? BOSS_NEG_VARIANCE : 0;
}

int boss_gp_predict_dev(const boss_gp *gp, const double *Xs_dev, int64_t M, const double *prior_mean_s_dev,
                        double *mu_dev, double *var_dev, int32_t *status_dev) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!gp || (!Xs_dev && M > 0)) return fail(BOSS_ERR_ARG, "boss_gp_predict_dev: bad arguments");
  const boss_gp *sl[1] = {gp};
  ScoreArgs a{};
  a.slices = sl;
  a.y_dim = 1;
  a.n_samples = 1;
  a.Xs = Xs_dev;
  a.M = M;
  a.prior_mean = prior_mean_s_dev;
  a.mu_out = mu_dev;
  a.var_out = var_dev;
  a.status_out = status_dev;
  a.dev = true;
  a.want_argmax = false;
  int rc = score_core(a);
  if (rc) return rc;
  return a.any_fail ? BOSS_NEG_VARIANCE : 0;
}

static int ei_score_impl(const boss_gp *const *slices, int y_dim, int n_samples, const double *Xs, int64_t M,
                         const double *prior_mean_s, const double *fit_coefs, const double *best,
                         const double *y_max, const double *lb, const double *ub, const uint8_t *cons_mask,
                         double *acq, double *grad, double *best_val, int64_t *best_idx, bool dev,
                         const double *prior_mean_grad = nullptr) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!slices || y_dim < 1 || n_samples < 1 || (!Xs && M > 0) || !fit_coefs)
    return fail(BOSS_ERR_ARG, "boss_ei_score: bad arguments");
  if ((lb == nullptr) != (ub == nullptr)) return fail(BOSS_ERR_ARG, "boss_ei_score: lb and ub must be given together");
  ScoreArgs a{};
  a.slices = slices;
  a.y_dim = y_dim;
  a.n_samples = n_samples;
  a.Xs = Xs;
  a.M = M;
  a.prior_mean = prior_mean_s;
  a.coefs = fit_coefs;
  a.best = best;
  a.y_max = y_max;
  a.lb = lb;
  a.ub = ub;
  a.cons_mask = cons_mask;
  a.acq = acq;
  a.grad = grad;
  a.prior_mean_grad = prior_mean_grad;
  a.best_val = best_val;
  a.best_idx = best_idx;
  a.dev = dev;
  a.want_argmax = (best_val != nullptr) || (best_idx != nullptr);
  return score_core(a);
}

int boss_ei_score(const boss_gp *const *slices, int y_dim, int n_samples, const double *Xs, int64_t M,
                  const double *prior_mean_s, const double *fit_coefs, const double *best, const double *y_max,
                  const double *lb, const double *ub, const uint8_t *cons_mask, double *acq, double *grad,
                  double *best_val, int64_t *best_idx) {
  return ei_score_impl(slices, y_dim, n_samples, Xs, M, prior_mean_s, fit_coefs, best, y_max, lb, ub, cons_mask, acq,
                       grad, best_val, best_idx, false);
}

int boss_ei_score_dev(const boss_gp *const *slices, int y_dim, int n_samples, const double *Xs_dev, int64_t M,
                      const double *prior_mean_s_dev, const double *fit_coefs, const double *best,
                      const double *y_max, const double *lb, const double *ub, const uint8_t *cons_mask_dev,
                      double *acq_dev, double *grad_dev, double *best_val, int64_t *best_idx, void *stream) {
  (void)stream;  // work is ordered on the library stream and synchronised before return
  return ei_score_impl(slices, y_dim, n_samples, Xs_dev, M, prior_mean_s_dev, fit_coefs, best, y_max, lb, ub,
                       cons_mask_dev, acq_dev, grad_dev, best_val, best_idx, true);
}

static int ei_score_generated(const CandGen &gen, int64_t M, const boss_gp *const *slices, int y_dim, int n_samples,
                              const double *prior_mean_s, const double *fit_coefs, const double *best, const double *y_max,
                              const double *lb, const double *ub, const uint8_t *cons_mask, double *acq, double *best_val,
                              int64_t *best_idx, double *best_x) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!slices || y_dim < 1 || n_samples < 1 || !fit_coefs || !slices[0]) return fail(BOSS_ERR_ARG, "boss_ei_score_*: bad arguments");
  if ((lb == nullptr) != (ub == nullptr)) return fail(BOSS_ERR_ARG, "boss_ei_score_*: lb and ub must be given together");
  if (gen.d != slices[0]->d) return fail(BOSS_ERR_ARG, "boss_ei_score_*: x_dim mismatch");
  ScoreArgs a{};
  a.slices = slices;
  a.y_dim = y_dim;
  a.n_samples = n_samples;
  a.Xs = nullptr;
  a.M = M;
  a.prior_mean = prior_mean_s;
  a.coefs = fit_coefs;
  a.best = best;
  a.y_max = y_max;
  a.lb = lb;
  a.ub = ub;
  a.cons_mask = cons_mask;
  a.acq = acq;
  int64_t bi = -1;
  double bv = 0.0;
  a.best_val = &bv;
  a.best_idx = &bi;
  a.dev = false;
  a.want_argmax = true;
  a.gen = &gen;
  int rc = score_core(a);
  if (rc) return rc;
  if (best_val) *best_val = bv;
  if (best_idx) *best_idx = bi < 0 ? bi : gen.first + bi;                  // global index
  if (best_x && bi >= 0)
    for (int j = 0; j < gen.d; ++j) best_x[j] = cand_coord(gen, gen.first + bi, j);   // same arithmetic as the device generator
  return 0;
}

int boss_ei_score_grid(const boss_gp *const *slices, int y_dim, int n_samples, int d, const double *grid_lo,
                       const double *grid_step, const int64_t *grid_count, int64_t first, int64_t M,
                       const double *prior_mean_s, const double *fit_coefs, const double *best, const double *y_max,
                       const double *lb, const double *ub, const uint8_t *cons_mask, double *acq, double *best_val,
                       int64_t *best_idx, double *best_x) {
  if (d < 1 || d > 32 || !grid_lo || !grid_step || !grid_count || first < 0)
    return fail(BOSS_ERR_ARG, "boss_ei_score_grid: bad grid");
  CandGen gen{};
  gen.mode = 1;
  gen.d = d;
  gen.first = first;
  int64_t total = 1;
  for (int j = 0; j < d; ++j) {
    if (grid_count[j] < 1) return fail(BOSS_ERR_ARG, "boss_ei_score_grid: grid_count must be >= 1");
    gen.lo[j] = grid_lo[j];
    gen.step[j] = grid_step[j];
    gen.count[j] = grid_count[j];
    if (total > ((int64_t)1 << 40)) return fail(BOSS_ERR_ARG, "boss_ei_score_grid: grid too large");
    total *= grid_count[j];
  }
  if (M < 0) M = total - first;
  if (first + M > total) return fail(BOSS_ERR_ARG, "boss_ei_score_grid: shard exceeds the grid");
  return ei_score_generated(gen, M, slices, y_dim, n_samples, prior_mean_s, fit_coefs, best, y_max, lb, ub, cons_mask, acq,
                            best_val, best_idx, best_x);
}

int boss_ei_score_uniform(const boss_gp *const *slices, int y_dim, int n_samples, int d, uint64_t seed, int64_t first,
                          int64_t M, const double *box_lb, const double *box_ub, const double *prior_mean_s,
                          const double *fit_coefs, const double *best, const double *y_max, const uint8_t *cons_mask,
                          double *acq, double *best_val, int64_t *best_idx, double *best_x) {
  if (d < 1 || d > 32 || !box_lb || !box_ub || M < 0 || first < 0)
    return fail(BOSS_ERR_ARG, "boss_ei_score_uniform: bad arguments");
  CandGen gen{};
  gen.mode = 2;
  gen.d = d;
  gen.seed = seed;
  gen.first = first;
  for (int j = 0; j < d; ++j) {
    gen.lo[j] = box_lb[j];
    gen.step[j] = box_ub[j] - box_lb[j];
  }
  return ei_score_generated(gen, M, slices, y_dim, n_samples, prior_mean_s, fit_coefs, best, y_max, box_lb, box_ub,
                            cons_mask, acq, best_val, best_idx, best_x);
}

// Device-resident lock-step multi-start maximisation of the acquisition (multistart.cuh).
int boss_ei_maximize_multistart(const boss_gp *const *slices, int y_dim, int n_samples, const double *starts, int64_t M,
                                int iters, int history, const double *prior_mean_affine, const double *fit_coefs,
                                const double *best, const double *y_max, const double *lb, const double *ub,
                                const uint8_t *discrete_mask, double *x_out, double *f_out, double *best_x,
                                double *best_val, int64_t *best_idx, int *evals_out) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!slices || !slices[0] || y_dim < 1 || n_samples < 1 || !starts || M < 1 || !fit_coefs || !lb || !ub)
    return fail(BOSS_ERR_ARG, "boss_ei_maximize_multistart: bad arguments (the box lb/ub is required)");
  const int d = slices[0]->d;
  if (d > MS_MAXD) return fail(BOSS_ERR_ARG, "boss_ei_maximize_multistart: x_dim > 32");
  const int H = std::max(1, std::min(history, MS_MAXH - 1));
  const int RC = H + 1;   // ring capacity: H live pairs + one free slot for the pair being formed
  if (iters < 0) iters = 0;
  const size_t Md = (size_t)M * d;
  // carve one workspace: X Xt Xn | g gt gn | dirn | Sh Yh | f ft fn t | done | moved, counters
  const size_t FANCAP = 4;   // compact trial buffers hold up to 4 M points (fan of MS_FAN steps for <= 0.4 M stragglers)
  const size_t n_aff = prior_mean_affine ? (size_t)y_dim * (d + 1) + FANCAP * ((size_t)M * y_dim + Md * y_dim) : 0;
  const size_t n_dbl = 3 * Md + 3 * Md + Md + 2 * (size_t)RC * Md + 4 * (size_t)M + n_aff + FANCAP * (2 * Md + M);
  CUDA_TRY(g.ms_buf.ensure(n_dbl * 8 + (size_t)M * 12 + 64));
  double *base = g.ms_buf.as<double>();
  MsState st{};
  st.d = d;
  st.H = RC;
  st.M = M;
  st.X = base;
  st.Xt = st.X + Md;
  st.Xn = st.Xt + Md;
  st.g = st.Xn + Md;
  st.gt = st.g + Md;
  st.gn = st.gt + Md;
  st.dirn = st.gn + Md;
  st.Sh = st.dirn + Md;
  st.Yh = st.Sh + (size_t)RC * Md;
  st.f = st.Yh + (size_t)RC * Md;
  st.ft = st.f + M;
  st.fn = st.ft + M;
  st.t = st.fn + M;
  double *aff = st.t + M, *pm = aff + (size_t)y_dim * (d + 1), *pmg = pm + FANCAP * (size_t)M * y_dim;
  if (prior_mean_affine)
    CUDA_TRY(cudaMemcpyAsync(aff, prior_mean_affine, (size_t)y_dim * (d + 1) * 8, cudaMemcpyHostToDevice, g.stream));
  st.Xc = st.t + M + n_aff;
  st.gc = st.Xc + FANCAP * Md;
  st.fc = st.gc + FANCAP * Md;
  st.moved_bits = reinterpret_cast<unsigned long long *>(st.fc + FANCAP * M);
  st.counters = reinterpret_cast<int *>(st.moved_bits + 1);
  st.done = st.counters + 4;
  st.frozen = st.done + M;
  st.idx = st.frozen + M;
  CUDA_TRY(cudaMemsetAsync(st.frozen, 0, (size_t)M * 4, g.stream));
  double span = 0.0;
  for (int j = 0; j < d; ++j) {
    st.lb[j] = lb[j];
    st.ub[j] = ub[j];
    span = std::max(span, ub[j] - lb[j]);
  }
  st.step0 = 0.1 * span;
  const unsigned nbm = (unsigned)((M + 127) / 128);

  long long evaluated = 0;
  auto eval = [&](const double *Xp, long long Mp, double *fp, double *gp, bool want_best, double *bv, int64_t *bi) -> int {
    ScoreArgs a{};
    a.slices = slices;
    a.y_dim = y_dim;
    a.n_samples = n_samples;
    a.Xs = Xp;
    a.M = Mp;
    evaluated += Mp;
    if (prior_mean_affine) {
      ms_affine_mean_kernel<<<(unsigned)((Mp + 127) / 128), 128, 0, g.stream>>>(Xp, Mp, d, y_dim, aff, pm, pmg);
      ++g.launches;
      a.prior_mean = pm;
      a.prior_mean_grad = gp ? pmg : nullptr;
    }
    a.coefs = fit_coefs;
    a.best = best;
    a.y_max = y_max;
    a.lb = lb;
    a.ub = ub;
    a.acq = fp;
    a.grad = gp;
    a.best_val = bv;
    a.best_idx = bi;
    a.dev = true;
    a.want_argmax = want_best;
    return score_core(a);
  };

  CUDA_TRY(cudaMemcpyAsync(st.Xt, starts, Md * 8, cudaMemcpyHostToDevice, g.stream));
  ms_init_kernel<<<nbm, 128, 0, g.stream>>>(st, st.Xt);
  int rc = eval(st.X, M, st.f, st.g, false, nullptr, nullptr);
  if (rc) return rc;
  ms_sanitize_kernel<<<nbm, 128, 0, g.stream>>>(st.f, M);
  int hist_len = 0, hist_start = 0;
  for (int it = 0; it < iters; ++it) {
    ms_direction_kernel<<<nbm, 128, 0, g.stream>>>(st, hist_len, hist_start);
    for (int trial = 0; trial < 12; ++trial) {
      // only the starts that have not yet accepted a step are evaluated again (compacted batch)
      CUDA_TRY(cudaMemsetAsync(st.counters, 0, 16, g.stream));
      ms_compact_kernel<<<nbm, 128, 0, g.stream>>>(st);
      int count = 0;
      CUDA_TRY(cudaMemcpyAsync(&count, st.counters, 4, cudaMemcpyDeviceToHost, g.stream));
      CUDA_TRY(cudaStreamSynchronize(g.stream));
      ++g.launches;
      if (count == 0) break;
      if (trial >= 2 && (size_t)count * MS_FAN <= FANCAP * (size_t)M) {
        // few stragglers left: all remaining step sizes in one batch instead of up to 10 tiny sequential ones
        ms_fan_kernel<<<(count * MS_FAN + 127) / 128, 128, 0, g.stream>>>(st, count);
        rc = eval(st.Xc, (long long)count * MS_FAN, st.fc, st.gc, false, nullptr, nullptr);
        if (rc) return rc;
        ms_accept_fan_kernel<<<(count + 127) / 128, 128, 0, g.stream>>>(st, count);
        g.launches += 2;
        break;
      }
      rc = eval(st.Xc, count, st.fc, st.gc, false, nullptr, nullptr);
      if (rc) return rc;
      ms_accept_kernel<<<(count + 127) / 128, 128, 0, g.stream>>>(st, count);
      ++g.launches;
    }
    ms_freeze_kernel<<<nbm, 128, 0, g.stream>>>(st);
    CUDA_TRY(cudaMemsetAsync(st.moved_bits, 0, 8 + 16, g.stream));
    ms_update_kernel<<<nbm, 128, 0, g.stream>>>(st, (hist_start + hist_len) % RC);   // always a free slot
    g.launches += 3;
    struct {
      unsigned long long moved;
      int cnt[4];
    } hs;
    CUDA_TRY(cudaMemcpyAsync(&hs, st.moved_bits, 24, cudaMemcpyDeviceToHost, g.stream));
    CUDA_TRY(cudaStreamSynchronize(g.stream));
    if (hs.cnt[1]) {   // at least one start produced a valid pair: commit the slot
      if (hist_len == H)
        hist_start = (hist_start + 1) % RC;   // drop the oldest pair
      else
        ++hist_len;
    }
    double moved;
    std::memcpy(&moved, &hs.moved, 8);
    if (moved < 1e-10) break;
  }
  // final rounding of discrete dimensions and re-evaluation (optimization.jl:116-117), argmax over the starts
  const unsigned long long disc = mask_bits(discrete_mask, d);
  if (disc) ms_round_kernel<<<(unsigned)((Md + 255) / 256), 256, 0, g.stream>>>(st.X, M, d, disc);
  double bv = 0.0;
  int64_t bi = -1;
  rc = eval(st.X, M, st.f, nullptr, true, &bv, &bi);
  if (rc) return rc;
  if (x_out) CUDA_TRY(cudaMemcpyAsync(x_out, st.X, Md * 8, cudaMemcpyDeviceToHost, g.stream));
  if (f_out) CUDA_TRY(cudaMemcpyAsync(f_out, st.f, (size_t)M * 8, cudaMemcpyDeviceToHost, g.stream));
  if (best_x && bi >= 0) CUDA_TRY(cudaMemcpyAsync(best_x, st.X + (size_t)bi * d, (size_t)d * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  if (best_val) *best_val = bv;
  if (best_idx) *best_idx = bi;
  if (evals_out) *evals_out = (int)((evaluated + M - 1) / M);   // work in units of full-batch evaluations
  return 0;
}

int boss_ei_value_grad(const boss_gp *const *slices, int y_dim, int n_samples, const double *Xs, int64_t M,
                       const double *prior_mean_s, const double *prior_mean_grad_s, const double *fit_coefs,
                       const double *best, const double *y_max, const double *lb, const double *ub,
                       const uint8_t *cons_mask, double *acq, double *grad) {
  if (!grad) return fail(BOSS_ERR_ARG, "boss_ei_value_grad: grad is NULL");
  return ei_score_impl(slices, y_dim, n_samples, Xs, M, prior_mean_s, fit_coefs, best, y_max, lb, ub, cons_mask, acq,
                       grad, nullptr, nullptr, false, prior_mean_grad_s);
}

int boss_ei_value_grad_dev(const boss_gp *const *slices, int y_dim, int n_samples, const double *Xs_dev, int64_t M,
                           const double *prior_mean_s_dev, const double *prior_mean_grad_s_dev,
                           const double *fit_coefs, const double *best, const double *y_max, const double *lb,
                           const double *ub, const uint8_t *cons_mask_dev, double *acq_dev, double *grad_dev) {
  if (!grad_dev) return fail(BOSS_ERR_ARG, "boss_ei_value_grad_dev: grad is NULL");
  return ei_score_impl(slices, y_dim, n_samples, Xs_dev, M, prior_mean_s_dev, fit_coefs, best, y_max, lb, ub,
                       cons_mask_dev, acq_dev, grad_dev, nullptr, nullptr, true, prior_mean_grad_s_dev);
}

int boss_gp_cov(const boss_gp *gp, const double *Xs, int64_t M, const double *prior_mean_s, double *mu, double *cov) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!gp || !Xs || !cov || M < 1) return fail(BOSS_ERR_ARG, "boss_gp_cov: bad arguments");
  if (M > 8192) return fail(BOSS_ERR_ARG, "boss_gp_cov: M > 8192 (the full covariance is meant for small batches)");
  const boss_gp *h = gp;
  const int d = h->d, ncb = (int)((M + 127) / 128), Mp = ncb * 128, cnt = Mp;
  const size_t blk_elems = (size_t)Mp * h->n_pad;
  CUDA_TRY(g.ks.ensure(blk_elems * 8));
  CUDA_TRY(g.vt.ensure(blk_elems * 8));
  CUDA_TRY(g.part_mu.ensure((size_t)2 * h->nblk * Mp * 8));
  CUDA_TRY(g.part_ss.ensure((size_t)2 * h->nblk * Mp * 8));
  CUDA_TRY(g.muv.ensure((size_t)Mp * 8));
  CUDA_TRY(g.cov_p.ensure((size_t)Mp * Mp * 8));
  CUDA_TRY(g.cov_stage.ensure((size_t)M * M * 8));
  CUDA_TRY(g.xs_stage.ensure((size_t)Mp * d * 8));
  CUDA_TRY(g.mu_stage.ensure((size_t)Mp * 8));
  CUDA_TRY(g.small.ensure(64));
  if (prior_mean_s) CUDA_TRY(g.pm_stage.ensure((size_t)M * 8));
  int *d_fail = g.small.as<int>();
  CUDA_TRY(cudaMemsetAsync(d_fail, 0, 4, g.stream));
  CUDA_TRY(cudaMemcpyAsync(g.xs_stage.p, Xs, (size_t)M * d * 8, cudaMemcpyHostToDevice, g.stream));
  if (prior_mean_s) CUDA_TRY(cudaMemcpyAsync(g.pm_stage.p, prior_mean_s, (size_t)M * 8, cudaMemcpyHostToDevice, g.stream));
  const int want = (2 * 148 + ncb - 1) / ncb;
  const int nks = std::max(1, std::min(h->nblk, want));
  const int nsp = pick_row_splits(ncb, h->nblk);
  XcovParams xp{};
  xp.Xs = g.xs_stage.as<double>();
  xp.M = M;
  xp.d = d;
  xp.n = h->n;
  xp.n_pad = h->n_pad;
  xp.ktiles = h->ktiles;
  xp.Xt = h->Xt;
  xp.invl = h->invl;
  xp.disc_bits = h->disc;
  xp.alpha = h->alpha;
  xp.a2 = h->a2;
  xp.Ks = g.ks.as<double>();
  xp.mu_part = g.part_mu.as<double>();
  xp.ld = Mp;
  DISPATCH_KID_DP(launch_xcov_t, h->kernel_id, h->dp, xp, dim3(ncb, nks));
  reduce_rows_kernel<<<(cnt + 255) / 256, 256, 0, g.stream>>>(g.part_mu.as<double>(), 2 * h->nblk, (size_t)Mp,
                                                             g.muv.as<double>(), cnt);
  ScoreParams sp{};
  sp.W = h->W;
  sp.Ks = g.ks.as<double>();
  sp.nblk = h->nblk;
  sp.ktiles = h->ktiles;
  sp.ss_part = g.part_ss.as<double>();
  sp.ld = Mp;
  sp.VT = g.vt.as<double>();
  score_trmm_kernel<<<dim3(ncb, nsp), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(sp);
  GemmNtParams gm{g.vt.as<double>(), g.vt.as<double>(), g.cov_p.as<double>(), h->ktiles, Mp / TK, 1};
  gemm_nt_kernel<<<dim3(ncb, ncb), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(gm);
  CovFinishParams cf{};
  cf.Xs = g.xs_stage.as<double>();
  cf.M = (int)M;
  cf.d = d;
  cf.ktilesC = Mp / TK;
  cf.invl = h->invl;
  cf.disc_bits = h->disc;
  cf.a2 = h->a2;
  cf.C = g.cov_p.as<double>();
  cf.mu = g.muv.as<double>();
  cf.prior_mean = prior_mean_s ? g.pm_stage.as<double>() : nullptr;
  cf.mu_out = mu ? g.mu_stage.as<double>() : nullptr;
  cf.cov = g.cov_stage.as<double>();
  cf.any_fail = d_fail;
  DISPATCH_KID_DP(launch_cov_finish_t, h->kernel_id, h->dp, cf, dim3((unsigned)((M + 15) / 16), (unsigned)((M + 15) / 16)));
  g.launches += 5;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(cov, g.cov_stage.p, (size_t)M * M * 8, cudaMemcpyDeviceToHost, g.stream));
  if (mu) CUDA_TRY(cudaMemcpyAsync(mu, g.mu_stage.p, (size_t)M * 8, cudaMemcpyDeviceToHost, g.stream));
  int hfail = 0;
  CUDA_TRY(cudaMemcpyAsync(&hfail, d_fail, 4, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  return hfail ? BOSS_NEG_VARIANCE : 0;
}

// ---------------------------------------------------------------------------------------------
// batched log marginal likelihood
// ---------------------------------------------------------------------------------------------
static int loglik_impl(const double *X, int d, int n, const double *Ymm, int64_t ldy, const double *ls,
                       const double *amp, const double *noise, int kernel_id, const uint8_t *discrete_mask, int64_t S,
                       double *loglik, bool dev, double *grad = nullptr, boss_gp **fit_out = nullptr) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!X || !Ymm || !ls || !amp || !noise || !loglik || d < 1 || n < 1 || S < 0)
    return fail(BOSS_ERR_ARG, "boss_gp_loglik_batch: bad arguments");
  if (kernel_id < 0 || kernel_id > 2) return fail(BOSS_ERR_ARG, "boss_gp_loglik_batch: unknown kernel_id");
  const int dp = pick_dp(d);
  if (dp < 0) return fail(BOSS_ERR_ARG, "boss_gp_loglik_batch: x_dim > 32 is not supported");
  if (S == 0) return 0;
  const int n_pad = round_up(n, TM), nblk = n_pad / TM, ktiles = n_pad / TK;
  const size_t mat = (size_t)n_pad * n_pad;
  const unsigned long long disc = mask_bits(discrete_mask, d);

  const bool need_w = grad || fit_out;        // gradient and batched-fit modes also form W = L^-1, W^T and alpha
  const bool small = n <= SMALL_N && !need_w; // warp-register path (small.cuh): no workspace at all
  // sub-batch so that L + Winv (+ W, W^T and the trtri scratch) stay within a 32 GiB workspace
  const size_t per = (need_w ? 3 : 1) * mat * 8 + (size_t)(need_w ? 2 * nblk : nblk) * TM * TM * 8;
  long long Sb = std::max<long long>(1, std::min<long long>(S, (32ll << 30) / (long long)per));
  Sb = std::min<long long>(Sb, 32768);
  if (small) Sb = 1;
  if (!small) {
    CUDA_TRY(g.chol_L.ensure((size_t)Sb * mat * 8));
    CUDA_TRY(g.chol_Winv.ensure((size_t)Sb * nblk * TM * TM * 8));
  }
  const int ntiles = nblk * (nblk + 1) / 2;
  const size_t tt_stride = (size_t)std::max(1, nblk - 1) * TM * TM;
  if (need_w) {
    CUDA_TRY(g.chol_W.ensure((size_t)Sb * mat * 8));
    CUDA_TRY(g.chol_WT.ensure((size_t)Sb * mat * 8));
    CUDA_TRY(g.ll_vec.ensure((size_t)Sb * n_pad * 3 * 8));                 // delta_pad | w | alpha
    if (grad) CUDA_TRY(g.ll_part.ensure((size_t)Sb * ntiles * (dp + 2) * 8));
  }

  // device copies of the inputs when called with host pointers
  const double *dX = X, *dY = Ymm, *dls = ls, *damp = amp, *dnoise = noise;
  double *dll = loglik;
  const size_t ycount = ldy ? (size_t)S * ldy : (size_t)n;
  // misc: [X | Y | ls | amp | noise | ll] (host variant) then logdet_blk[Sb*nblk] | ssq_blk[Sb*nblk] | r, w [Sb*n_pad] | status[Sb]
  size_t off = 0, oX = 0, oY = 0, ols = 0, oamp = 0, onoise = 0, oll = 0;
  if (!dev) {
    oX = off; off += (size_t)n * d;
    oY = off; off += ycount;
    ols = off; off += (size_t)S * d;
    oamp = off; off += S;
    onoise = off; off += S;
    oll = off; off += S;
  }
  size_t ogr = 0;
  if (!dev && grad) {
    ogr = off; off += (size_t)S * (d + 2);
  }
  const size_t old_ = off;
  off += (size_t)Sb * nblk;
  const size_t ossq = off;                      // |w_j|^2 per block, then r and w of the in-line forward substitution
  off += (size_t)Sb * nblk;
  const size_t orv = off;
  off += small ? 0 : (size_t)Sb * n_pad;
  const size_t owv = off;
  off += small ? 0 : (size_t)Sb * n_pad;
  const size_t ost = off;
  off += (size_t)(Sb + 1) / 2 + 1;
  CUDA_TRY(g.chol_misc.ensure(off * 8));
  double *misc = g.chol_misc.as<double>();
  if (!dev) {
    CUDA_TRY(cudaMemcpyAsync(misc + oX, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(misc + oY, Ymm, ycount * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(misc + ols, ls, (size_t)S * d * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(misc + oamp, amp, (size_t)S * 8, cudaMemcpyHostToDevice, g.stream));
    CUDA_TRY(cudaMemcpyAsync(misc + onoise, noise, (size_t)S * 8, cudaMemcpyHostToDevice, g.stream));
    dX = misc + oX; dY = misc + oY; dls = misc + ols; damp = misc + oamp; dnoise = misc + onoise; dll = misc + oll;
  }
  double *dgrad = (grad && !dev) ? misc + ogr : grad;
  double *logdet_blk = misc + old_;
  int *status = reinterpret_cast<int *>(misc + ost);

  timing_begin();
  if (small) {
    SmallLoglikParams sp{};
    sp.X = dX;
    sp.d = d;
    sp.n = n;
    sp.ymm = dY;
    sp.ldy = ldy;
    sp.ls = dls;
    sp.amp = damp;
    sp.noise = dnoise;
    sp.disc_bits = disc;
    sp.S = S;
    sp.loglik = dll;
    Timed t(2);
    DISPATCH_KID_DP(launch_loglik_small_t, kernel_id, dp, sp, (int)((S + SMALL_WARPS - 1) / SMALL_WARPS));
    ++g.launches;
  }
  // A window of up to Sb matrices is live at a time.  Inside a window the matrices are processed as up to
  // LL_GROUPS independent groups on separate streams: one group's latency-bound steps (diagonal-block
  // factorisation, forward solve) and partially filled last waves overlap with another group's GEMM launches.
  cudaStream_t main_stream = g.stream;
  const size_t winv_stride = (size_t)nblk * TM * TM;
  for (long long s0 = 0; s0 < S && !small; s0 += Sb) {
    const int sb = (int)std::min<long long>(Sb, S - s0);
    int want_groups = 4;
    if (const char *e = getenv("BOSS_LL_GROUPS")) want_groups = std::max(1, std::min(LL_GROUPS, atoi(e)));
    const int ngroups = (g.timing || sb < 2 * want_groups) ? 1 : want_groups;   // per-kernel-class timing needs one stream
    const int gsz = (sb + ngroups - 1) / ngroups;
    if (ngroups > 1) CUDA_TRY(cudaEventRecord(g.ll_fork, main_stream));
    int rc_all = 0;
    for (int gi = 0; gi < ngroups; ++gi) {
      const int wo = gi * gsz;                           // offset of the group inside the window
      const int gs = std::min(gsz, sb - wo);
      if (gs <= 0) break;
      const long long so = s0 + wo;                      // offset of the group inside the batch
      if (ngroups > 1) {
        g.stream = g.ll_stream[gi];
        CUDA_TRY(cudaStreamWaitEvent(g.stream, g.ll_fork, 0));
      }
      double *Lg = g.chol_L.as<double>() + (size_t)wo * mat;
      double *Wig = g.chol_Winv.as<double>() + (size_t)wo * winv_stride;
      double *ldg = logdet_blk + (size_t)wo * nblk;
      int *stg = status + wo;
      double *Wg = need_w ? g.chol_W.as<double>() + (size_t)wo * mat : nullptr;
      double *WTg = need_w ? g.chol_WT.as<double>() + (size_t)wo * mat : nullptr;
      double *TTg = nullptr;   // the fused triangular-inverse step keeps T on chip
      cudaMemsetAsync(stg, 0, (size_t)gs * 4, g.stream);
      BuildKParams bk{};
      bk.X = dX;
      bk.d = d;
      bk.n = n;
      bk.nblk = nblk;
      bk.ktiles = ktiles;
      bk.ls = dls + (size_t)so * d;
      bk.amp = damp + so;
      bk.noise = dnoise + so;
      bk.disc_bits = disc;
      bk.K = Lg;
      bk.K_stride = mat;
      bk.status = stg;
      {
        Timed t(1);
        DISPATCH_KID_DP(launch_build_k_t, kernel_id, dp, bk, dim3(nblk * (nblk + 1) / 2, gs));
        ++g.launches;
      }
      if (need_w) {   // zero initial state of W and W^T (strictly-upper resp. strictly-lower blocks are never written)
        cudaMemsetAsync(Wg, 0, (size_t)gs * mat * 8, g.stream);
        cudaMemsetAsync(WTg, 0, (size_t)gs * mat * 8, g.stream);
      }
      FwdInline fw{dY + (ldy ? (size_t)so * ldy : 0), ldy, n, n_pad, misc + orv + (size_t)wo * n_pad,
                   misc + owv + (size_t)wo * n_pad, misc + ossq + (size_t)wo * nblk};
      int rc = run_cholesky(Lg, mat, Wig, winv_stride, nblk, ktiles, gs, ldg, stg, Wg, WTg, TTg, mat, tt_stride, false, &fw);
      if (rc) rc_all = rc;
      loglik_finish_kernel<<<(gs + 127) / 128, 128, 0, g.stream>>>(ldg, fw.ssq, stg, nblk, n, gs, dll + so);
      ++g.launches;
      double *vec = g.ll_vec.as<double>();
      double *dpad = vec + (size_t)wo * n_pad, *wv = vec + ((size_t)Sb + wo) * n_pad, *al = vec + ((size_t)2 * Sb + wo) * n_pad;
      if (need_w) {   // alpha = W^T (W delta)
        pad_delta_kernel<<<dim3((n_pad + 255) / 256, gs), 256, 0, g.stream>>>(dY + (ldy ? (size_t)so * ldy : 0), ldy, n, n_pad, dpad);
        matvec_p_kernel<<<dim3(n_pad / 64, gs), 256, 0, g.stream>>>(Wg, dpad, wv, ktiles, mat, n_pad, n_pad);
        matvec_p_kernel<<<dim3(n_pad / 64, gs), 256, 0, g.stream>>>(WTg, wv, al, ktiles, mat, n_pad, n_pad);
        g.launches += 3;
      }
      if (grad) {
        // K^-1 = W^T W over the factor's storage;  tile partial sums;  final scaling
        double *partg = g.ll_part.as<double>() + (size_t)wo * ntiles * (dp + 2);
        KinvParams kp{WTg, mat, Lg, mat, nblk, ktiles};
        {
          Timed t(2);
          kinv_wtw_kernel<<<dim3(ntiles, gs), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(kp);
        }
        LlGradParams lg{};
        lg.X = dX;
        lg.d = d;
        lg.n = n;
        lg.nblk = nblk;
        lg.ktiles = ktiles;
        lg.ls = dls + (size_t)so * d;
        lg.amp = damp + so;
        lg.noise = dnoise + so;
        lg.disc_bits = disc;
        lg.Kinv = Lg;
        lg.K_stride = mat;
        lg.alpha = al;
        lg.part = partg;
        {
          Timed t(1);
          DISPATCH_KID_DP(launch_llgrad_tile_t, kernel_id, dp, lg, dim3(ntiles, gs));
        }
        loglik_grad_final_kernel<<<(gs + 127) / 128, 128, 0, g.stream>>>(partg, ntiles, dp, d, lg.ls, lg.amp, lg.noise, stg,
                                                                        dgrad + (size_t)so * (d + 2), gs);
        g.launches += 3;
      }
      if (ngroups > 1) {
        cudaEventRecord(g.ll_join[gi], g.stream);
        g.stream = main_stream;
        cudaStreamWaitEvent(main_stream, g.ll_join[gi], 0);
      }
    }
    g.stream = main_stream;
    if (rc_all) return rc_all;
    if (fit_out) {
      // batched posterior fit: detach every factorisation of the window into its own handle
      std::vector<int> hst(sb);
      std::vector<double> hll(sb);
      CUDA_TRY(cudaMemcpyAsync(hst.data(), status, (size_t)sb * 4, cudaMemcpyDeviceToHost, g.stream));
      CUDA_TRY(cudaMemcpyAsync(hll.data(), dll + s0, (size_t)sb * 8, cudaMemcpyDeviceToHost, g.stream));
      CUDA_TRY(cudaStreamSynchronize(g.stream));
      double *vec = g.ll_vec.as<double>();
      for (int q = 0; q < sb; ++q) {
        const long long sidx = s0 + q;
        fit_out[sidx] = nullptr;
        if (hst[q] != 0) continue;
        boss_gp *h = new boss_gp();
        h->n = n; h->d = d; h->n_pad = n_pad; h->nblk = nblk; h->ktiles = ktiles; h->kernel_id = kernel_id; h->dp = dp;
        h->disc = disc;
        h->amp = amp[sidx] + MIN_PARAM_VALUE;
        h->noise = noise[sidx] + MIN_PARAM_VALUE;
        h->a2 = h->amp * h->amp;
        h->loglik = hll[q];
        cudaError_t e = cudaSuccess;
        for (auto pr : {std::make_pair(&h->L, mat * 8), std::make_pair(&h->W, mat * 8), std::make_pair(&h->WT, mat * 8),
                        std::make_pair(&h->alpha, (size_t)n_pad * 8), std::make_pair(&h->Xt, (size_t)n_pad * dp * 8),
                        std::make_pair(&h->invl, (size_t)dp * 8), std::make_pair(&h->wvec, (size_t)n_pad * 8),
                        std::make_pair(&h->ymm, (size_t)n_pad * 8)})
          if (e == cudaSuccess) e = pool_alloc(pr.first, pr.second);
        if (e != cudaSuccess) {
          h->free_dev();
          delete h;
          return fail(BOSS_ERR_CUDA, std::string("boss_gp_fit_batch: ") + cudaGetErrorString(e));
        }
        std::vector<double> invl(dp, 0.0);
        for (int i = 0; i < d; ++i) invl[i] = 1.0 / (ls[(size_t)sidx * d + i] + MIN_PARAM_VALUE);
        cudaMemcpyAsync(h->invl, invl.data(), (size_t)dp * 8, cudaMemcpyHostToDevice, g.stream);
        cudaStreamSynchronize(g.stream);   // invl is a stack temporary
        cudaMemcpyAsync(h->L, g.chol_L.as<double>() + (size_t)q * mat, mat * 8, cudaMemcpyDeviceToDevice, g.stream);
        cudaMemcpyAsync(h->W, g.chol_W.as<double>() + (size_t)q * mat, mat * 8, cudaMemcpyDeviceToDevice, g.stream);
        cudaMemcpyAsync(h->WT, g.chol_WT.as<double>() + (size_t)q * mat, mat * 8, cudaMemcpyDeviceToDevice, g.stream);
        cudaMemcpyAsync(h->ymm, vec + (size_t)q * n_pad, (size_t)n_pad * 8, cudaMemcpyDeviceToDevice, g.stream);
        cudaMemcpyAsync(h->wvec, vec + ((size_t)Sb + q) * n_pad, (size_t)n_pad * 8, cudaMemcpyDeviceToDevice, g.stream);
        cudaMemcpyAsync(h->alpha, vec + ((size_t)2 * Sb + q) * n_pad, (size_t)n_pad * 8, cudaMemcpyDeviceToDevice, g.stream);
        scale_train_kernel<<<(n_pad * dp + 255) / 256, 256, 0, g.stream>>>(dX, d, n, n_pad, dp, h->invl, disc, h->Xt);
        ++g.launches;
        fit_out[sidx] = h;
      }
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaStreamSynchronize(g.stream));
    }
  }
  CUDA_TRY(cudaGetLastError());
  if (!dev) CUDA_TRY(cudaMemcpyAsync(loglik, dll, (size_t)S * 8, cudaMemcpyDeviceToHost, g.stream));
  if (!dev && grad) CUDA_TRY(cudaMemcpyAsync(grad, dgrad, (size_t)S * (d + 2) * 8, cudaMemcpyDeviceToHost, g.stream));
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  timing_end();
  if (!dev) {
    int any = 0;
    for (int64_t s = 0; s < S; ++s) {
      if (std::isnan(loglik[s])) return fail(BOSS_ERR_ARG, "boss_gp_loglik_batch: negative hyper-parameter");
      if (std::isinf(loglik[s])) any = 1;
    }
    return any ? BOSS_NOT_POSDEF : 0;
  }
  return 0;
}

int boss_gp_loglik_batch(const double *X, int d, int n, const double *Y_minus_mean, int64_t ldy,
                         const double *lengthscales, const double *amplitude, const double *noise_std, int kernel_id,
                         const uint8_t *discrete_mask, int64_t S, double *loglik) {
  return loglik_impl(X, d, n, Y_minus_mean, ldy, lengthscales, amplitude, noise_std, kernel_id, discrete_mask, S,
                     loglik, false);
}
int boss_gp_loglik_batch_dev(const double *X_dev, int d, int n, const double *Y_minus_mean_dev, int64_t ldy,
                             const double *lengthscales_dev, const double *amplitude_dev, const double *noise_std_dev,
                             int kernel_id, const uint8_t *discrete_mask, int64_t S, double *loglik_dev, void *stream) {
  (void)stream;
  return loglik_impl(X_dev, d, n, Y_minus_mean_dev, ldy, lengthscales_dev, amplitude_dev, noise_std_dev, kernel_id,
                     discrete_mask, S, loglik_dev, true);
}

int boss_gp_fit_batch(const double *X, int d, int n, const double *Y_minus_mean, int64_t ldy, const double *lengthscales,
                      const double *amplitude, const double *noise_std, int kernel_id, const uint8_t *discrete_mask,
                      int64_t S, boss_gp **out, double *loglik_out) {
  if (!out) return fail(BOSS_ERR_ARG, "boss_gp_fit_batch: out is NULL");
  for (int64_t s = 0; s < S; ++s) out[s] = nullptr;
  std::vector<double> ll((size_t)std::max<int64_t>(S, 1));
  int rc = loglik_impl(X, d, n, Y_minus_mean, ldy, lengthscales, amplitude, noise_std, kernel_id, discrete_mask, S, ll.data(),
                       false, nullptr, out);
  if (loglik_out)
    for (int64_t s = 0; s < S; ++s) loglik_out[s] = ll[s];
  if (rc < 0)
    for (int64_t s = 0; s < S; ++s)
      if (out[s]) {
        boss_gp_free(out[s]);
        out[s] = nullptr;
      }
  return rc;
}

int boss_gp_loglik_grad_batch(const double *X, int d, int n, const double *Y_minus_mean, int64_t ldy,
                              const double *lengthscales, const double *amplitude, const double *noise_std, int kernel_id,
                              const uint8_t *discrete_mask, int64_t S, double *loglik, double *grad) {
  if (!grad) return fail(BOSS_ERR_ARG, "boss_gp_loglik_grad_batch: grad is NULL");
  return loglik_impl(X, d, n, Y_minus_mean, ldy, lengthscales, amplitude, noise_std, kernel_id, discrete_mask, S, loglik,
                     false, grad);
}
int boss_gp_loglik_grad_batch_dev(const double *X_dev, int d, int n, const double *Y_minus_mean_dev, int64_t ldy,
                                  const double *lengthscales_dev, const double *amplitude_dev, const double *noise_std_dev,
                                  int kernel_id, const uint8_t *discrete_mask, int64_t S, double *loglik_dev,
                                  double *grad_dev) {
  if (!grad_dev) return fail(BOSS_ERR_ARG, "boss_gp_loglik_grad_batch_dev: grad is NULL");
  return loglik_impl(X_dev, d, n, Y_minus_mean_dev, ldy, lengthscales_dev, amplitude_dev, noise_std_dev, kernel_id,
                     discrete_mask, S, loglik_dev, true, grad_dev);
}

// ---------------------------------------------------------------------------------------------
// debug hooks
// ---------------------------------------------------------------------------------------------
static void pack_host(const double *dense, int R, int C, int Rp, int Cp, std::vector<double> &out) {
  out.assign((size_t)Rp * Cp, 0.0);
  const int kt = Cp / TK;
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < C; ++c) out[p_index(r, c, kt)] = dense[(size_t)r * C + c];
}

int boss_dbg_gemm_nt(const double *A, const double *B, int M, int N, int K, double *C) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  const int Mp = round_up(M, TM), Np = round_up(N, TM), Kp = round_up(K, TK);
  std::vector<double> pa, pb;
  pack_host(A, M, K, Mp, Kp, pa);
  pack_host(B, N, K, Np, Kp, pb);
  double *dA, *dB, *dC;
  CUDA_TRY(cudaMalloc(&dA, pa.size() * 8));
  CUDA_TRY(cudaMalloc(&dB, pb.size() * 8));
  CUDA_TRY(cudaMalloc(&dC, (size_t)Mp * Np * 8));
  CUDA_TRY(cudaMemcpy(dA, pa.data(), pa.size() * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemcpy(dB, pb.data(), pb.size() * 8, cudaMemcpyHostToDevice));
  GemmNtParams p{dA, dB, dC, Kp / TK, Np / TK, 0};
  gemm_nt_kernel<<<dim3(Np / TM, Mp / TM), GEMM_THREADS, GEMM_SMEM_BYTES, g.stream>>>(p);
  ++g.launches;
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  std::vector<double> pc((size_t)Mp * Np);
  CUDA_TRY(cudaMemcpy(pc.data(), dC, pc.size() * 8, cudaMemcpyDeviceToHost));
  for (int r = 0; r < M; ++r)
    for (int c = 0; c < N; ++c) C[(size_t)r * N + c] = pc[p_index(r, c, Np / TK)];
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dC);
  return 0;
}

__global__ void dbg_kernel_fn_kernel(int which, const double *t, double *out, int n) {
  __shared__ double tab[EXPTAB_N];
  exptab_init(tab);
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = t[i];
    double v;
    switch (which) {
      case 0: v = fast_exp_neg(x, tab); break;
      case 1: v = fast_sqrt(x); break;
      case 2: v = kappa_fast<0>(x, tab); break;
      case 3: v = kappa_fast<1>(x, tab); break;
      case 4: v = kappa_fast<2>(x, tab); break;
      case 5: v = kappa_dr_over_r_fast<0>(x, tab); break;
      case 6: v = kappa_dr_over_r_fast<1>(x, tab); break;
      default: v = kappa_dr_over_r_fast<2>(x, tab); break;
    }
    out[i] = v;
  }
}
int boss_dbg_kernel_fn(int which, const double *t, int n, double *out) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (which < 0 || which > 7 || n < 0) return fail(BOSS_ERR_ARG, "boss_dbg_kernel_fn: bad argument");
  double *dt, *dout;
  CUDA_TRY(cudaMalloc(&dt, (size_t)n * 8 + 8));
  CUDA_TRY(cudaMalloc(&dout, (size_t)n * 8 + 8));
  CUDA_TRY(cudaMemcpy(dt, t, (size_t)n * 8, cudaMemcpyHostToDevice));
  dbg_kernel_fn_kernel<<<148, 256, 0, g.stream>>>(which, dt, dout, n);
  CUDA_TRY(cudaStreamSynchronize(g.stream));
  CUDA_TRY(cudaMemcpy(out, dout, (size_t)n * 8, cudaMemcpyDeviceToHost));
  cudaFree(dt);
  cudaFree(dout);
  return 0;
}

int boss_dbg_factors(const boss_gp *gp, double *L, double *W, double *alpha) {
  std::lock_guard<std::mutex> lk(g.mu);
  REQUIRE_INIT();
  if (!gp) return fail(BOSS_ERR_ARG, "boss_dbg_factors: NULL handle");
  const size_t mat = (size_t)gp->n_pad * gp->n_pad;
  std::vector<double> buf(mat);
  const int n = gp->n;
  if (L) {
    CUDA_TRY(cudaMemcpy(buf.data(), gp->L, mat * 8, cudaMemcpyDeviceToHost));
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) L[(size_t)r * n + c] = (c <= r) ? buf[p_index(r, c, gp->ktiles)] : 0.0;
  }
  if (W) {
    CUDA_TRY(cudaMemcpy(buf.data(), gp->W, mat * 8, cudaMemcpyDeviceToHost));
    for (int r = 0; r < n; ++r)
      for (int c = 0; c < n; ++c) W[(size_t)r * n + c] = buf[p_index(r, c, gp->ktiles)];
  }
  if (alpha) CUDA_TRY(cudaMemcpy(alpha, gp->alpha, (size_t)n * 8, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
