// Shared host/device helpers of the lasso_b200 extension (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/lasso_b200.h"

namespace lasso {

// ---- error plumbing -------------------------------------------------------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define LASSO_CUDA_TRY(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::lasso::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                         __FILE__, __LINE__);                                         \
      return LASSO_B200_ERR_CUDA;                                                     \
    }                                                                                 \
  } while (0)

#define LASSO_CHECK_LAUNCH()                                                          \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::lasso::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),  \
                         __FILE__, __LINE__);                                         \
      return LASSO_B200_ERR_CUDA;                                                     \
    }                                                                                 \
  } while (0)

// ---- per-iteration control block shared by all FISTA kernels ---------------
// hist[i] = sum |z_i - z_{i+1}| of iteration i (double, device).  A kernel of
// iteration i exits at once when an earlier iteration already met the stop
// test, so the stream of launches needs no host synchronisation
// (reference: one blocking sync per iteration at ista.py:93).
struct StepCtl {
  double* hist;     // [maxiter] device
  double tol_abs;   // < 0 : stop test disabled
  int iter;         // index of this iteration
};

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float soft_threshold(float v, float lam) {
  // ATen softshrink: v > lam ? v - lam : (v < -lam ? v + lam : 0)   (ista.py:90)
  // Branch-free and bit-identical: v + lam == -(|v| - lam) exactly for v < -lam, the
  // result is +0 inside the dead zone and for NaN (both comparisons false), like ATen's.
  const float mag = fmaxf(__fsub_rn(fabsf(v), lam), 0.0f);
  return mag > 0.0f ? copysignf(mag, v) : 0.0f;
}

__device__ __forceinline__ float momentum_point(float zc, float zp, float beta) {
  // y = z_next + beta * (z_next - z), three roundings, no FMA (ista.py:100)
  return __fadd_rn(zc, __fmul_rn(beta, __fsub_rn(zc, zp)));
}

__device__ __forceinline__ float ista_update(float y, float g, float lr, float lam) {
  // softshrink(y - lr * g, alpha * lr), two roundings before the shrink (ista.py:90)
  return soft_threshold(__fsub_rn(y, __fmul_rn(lr, g)), lam);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// true when an earlier iteration already satisfied the stop test
__device__ __forceinline__ bool already_stopped(const StepCtl& c, int lookback) {
  if (c.tol_abs < 0.0) return false;
  for (int b = 1; b <= lookback; ++b) {
    int j = c.iter - b;
    if (j >= 0 && c.hist[j] <= c.tol_abs) return true;
  }
  return false;
}

// ---- launchers implemented in the .cu files ---------------------------------
struct FistaArgs {
  const float* x;
  const float* w;
  float* z_a;  // ping
  float* z_b;  // pong
  int64_t n;
  int d, k;
  float lr, lam;
  int maxiter, fast;
  double tol_abs;
  double* hist;  // [maxiter] device, zero-initialised
  int zero_start = 0;  // z_a holds the all-zero start (kernels that rescale the codes may skip it)
  int record = 1;      // 0: nobody reads hist (no stop test, no history asked for); kernels may skip the sums
};

// scale factors of the resident kernel (fista_res.cu): all powers of two
struct ResScalars {
  float sw;    // W' = sw W; row r of x is scaled by its own sx_r, its codes by sx_r / sw
  float isw;   // 1 / sw
  float lr;    // lr / sw^2
  float lam;   // lam / sw   (times sx_r per row)
  int bad;     // dictionary not finite / scaled step not representable
};

int fista_ffma_run(const FistaArgs& a, float* z_out, cudaStream_t st);
struct ConvShape {
  int64_t n_img;
  int cin, h, w, kh, kw;
  int stride, pad;
  // code grid of conv2d(x, W, stride, padding): (h + 2 pad - kh) / stride + 1 (exact division, else
  // conv_transpose2d of the codes would not reproduce x's size)
  __host__ __device__ int oh() const { return (h + 2 * pad - kh) / stride + 1; }
  __host__ __device__ int ow() const { return (w + 2 * pad - kw) / stride + 1; }
};
bool fista_blk_supported(int64_t n, int d, int k);
bool conv2d_blk_supported(const ConvShape& c, int k);
int fista_blk_run(const FistaArgs& a, int* fell_back, cudaStream_t st, const ConvShape* conv = nullptr);
bool fista_gram_supported(int64_t n, int d, int k);
int fista_gram_run(const FistaArgs& a, int* fell_back, cudaStream_t st);
bool fista_res_supported(int64_t n, int d, int k);
int fista_res_prepare(const float* w, int d, int k, float lr, float lam, int iters, int fast,
                      cudaStream_t st);
int64_t fista_res_wave_rows(int64_t n);
int64_t fista_res_tile_rows(int64_t n);
int fista_res_launch(const float* x, const float* z0, float* z_out, int64_t n, int d, int k, int iters,
                     double* hist, int hist_mode, int64_t tile_rows, cudaStream_t st);
int fista_res_finish(int* fell_back, cudaStream_t st);
int fista_res_run(const float* x, const float* w, const float* z0, float* z_out, int64_t n, int d,
                  int k, float lr, float lam, int iters, int fast, double* hist, int hist_mode,
                  int* fell_back, cudaStream_t st);
bool fista_tc_supported(int64_t n, int d, int k);
int fista_tc_run(const FistaArgs& a, float* z_out, cudaStream_t st);

int lipschitz_run(const float* w, int d, int k, int iters, double* l_dev, double* scratch,
                  cudaStream_t st);
int lambda_max_run(double* scratch, int m, int iters, double* l_dev, cudaStream_t st);
size_t lambda_max_scratch_doubles(int m);
size_t conv_lipschitz_scratch_bytes(int cin, int kh, int kw, int h, int w);
int conv_lipschitz_run(const float* weight, int f, int cin, int kh, int kw, int h, int w, int stride,
                       int pad, int iters, double* l_out, void* scratch, cudaStream_t st);
int matmul_run(const float* x, const float* t, int64_t n, int d, int k, float* z, cudaStream_t st);
size_t ridge_scratch_bytes(int d, int k);
int ridge_init_run(const float* x, const float* w, int64_t n, int d, int k, double alpha, float* z_out,
                   void* scratch, int* not_pd, cudaStream_t st);
int loss_terms_run(const float* x, const float* z, const float* w, int64_t n, int d, int k,
                   double* out, cudaStream_t st);
int gradient_run(const float* x, const float* point, const float* w, int64_t n, int d, int k,
                 float* grad, double* f_terms, cudaStream_t st);
int trial_run(const float* x, const float* point, const float* grad, const float* w, int64_t n,
              int d, int k, float step, float lam, float* cand, double* sums4, cudaStream_t st);
int momentum_run(const float* z_next, const float* z, float beta, float* y, int64_t count,
                 double* delta, cudaStream_t st);
bool gram_tc_supported(int64_t n, int d, int k);
int gram_tc_run(const float* z, const float* x, int64_t n, int d, int k, double* gzz, double* gzx,
                void* scratch, cudaStream_t st);
int gram_run(const float* z, const float* x, int64_t n, int d, int k, double* gzz,
             double* gzx, cudaStream_t st);
int zero_columns_run(float* z, int64_t n, int k, const int* mask, cudaStream_t st);
bool dict_update_blocked_supported(int d, int k);
int dict_update_blocked_run(float* dict, double* gzz, double* gzx, int d, int k, double eps, const float* redraw,
                            int* zeroed, int positive, cudaStream_t st);
int dict_update_run(float* dict, double* gzz, double* gzx, int d, int k, double eps,
                    const float* redraw, int* zeroed, int positive, cudaStream_t st);

}  // namespace lasso
