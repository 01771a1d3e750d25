/*
 * lasso_b200.h -- C ABI of the B200-native ISTA/FISTA sparse-encode engine.
 *
 * The reference (rfeinman/pytorch-lasso) has no FFI: its "operator API" is the
 * Python functions lasso.linear.sparse_encode / dict_learning / solvers.ista.
 * Each entry point below replaces the arithmetic of one reference call site
 * (file:line relative to the reference checkout) and is what a ctypes binding
 * on the reference side would call.
 *
 * Conventions
 *   - all matrices are row-major, contiguous, float32:
 *       x[n,d]  weight[d,k] (atoms are COLUMNS)  z[n,k]
 *   - "device" entry points take device pointers valid on the current CUDA
 *     device and a cudaStream_t passed as void*; they never synchronise unless
 *     stated.  "_host" entry points take host pointers, do the H2D / D2H copies
 *     themselves and return after the result is in the host buffer.
 *   - the caller owns every buffer passed in; the library owns only its private
 *     workspace (freed by lasso_b200_release_workspace or at process exit).
 *   - the workspace is per device and shared by all callers: the entry points
 *     that use it (the solves, lipschitz) take a per-device lock while they
 *     enqueue, and a call on another stream first waits (on the device) for the
 *     previous call's work, so calls from several threads / streams of one
 *     device are serialised, never interleaved.  The _host entry point holds
 *     the lock until its result is in the host buffer.
 *   - return value: 0 on success, negative lasso_b200_status on failure;
 *     lasso_b200_last_error() gives a thread-local message.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns LASSO_B200_ERR_CUDA.
 */
#ifndef LASSO_B200_H_
#define LASSO_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum lasso_b200_status {
  LASSO_B200_OK = 0,
  LASSO_B200_ERR_INVALID = -1,     /* bad argument (shape, null pointer, flag)   */
  LASSO_B200_ERR_CUDA = -2,        /* CUDA runtime / driver failure               */
  LASSO_B200_ERR_UNSUPPORTED = -3, /* shape or option not supported by this path  */
  LASSO_B200_ERR_NOMEM = -4        /* workspace allocation failed                 */
} lasso_b200_status;

/* which inner-loop kernel lasso_b200_fista_f32 uses */
typedef enum lasso_b200_path {
  LASSO_B200_PATH_AUTO = 0,    /* resident, else k-blocked / Gram-form tcgen05 kernel when the shape fits, else FFMA */
  LASSO_B200_PATH_FFMA = 1,    /* CUDA-core fp32 FFMA kernel: any n, d, k                    */
  LASSO_B200_PATH_TCGEN05 = 2, /* streaming tcgen05 kernel, one launch per iteration (bf16x3) */
  LASSO_B200_PATH_RESIDENT = 3, /* resident tcgen05 kernel (any d <= 64, k <= 256): all iterations
                                  of a 128-row tile on chip in ONE launch (fp16x2 operand split of the rescaled
                                  problem).  Synchronises the stream once per solve; falls
                                  back to LASSO_B200_PATH_TCGEN05 by itself if an iterate
                                  leaves the fp16 operand range.                               */
  LASSO_B200_PATH_BLOCKED = 4, /* k-blocked streaming tcgen05 kernel for dictionaries that do not
                                  fit one SM (d <= 128, k <= 1024, multiples of 4): one launch per
                                  iteration, codes and dictionary slices stream through a TMA
                                  ring.  Synchronises the stream once per solve; falls back to
                                  LASSO_B200_PATH_FFMA by itself like the resident path.        */
  LASSO_B200_PATH_GRAM = 5     /* Gram-form tcgen05 kernel for many features and few atoms (d > 128,
                                  k <= 320, k a multiple of 4 -- the reference notebook's d = 289,
                                  k = 300): W^T W and x W once per solve, then one GEMM per
                                  iteration with y resident in TMEM.  Same synchronisation and
                                  fallback behaviour as LASSO_B200_PATH_BLOCKED.               */
} lasso_b200_path;

/* ABI version: major*1000 + minor */
int32_t lasso_b200_version(void);

/* message of the last failure on this thread ("" if none) */
const char* lasso_b200_last_error(void);

/* path that lasso_b200_fista_f32 would take for (n,d,k) with LASSO_B200_PATH_AUTO */
int32_t lasso_b200_select_path(int64_t n, int32_t d, int32_t k);

/* number of kernels this library has launched in this process (bench: gpu_launches) */
int64_t lasso_b200_launch_count(void);

/* number of resident solves that had to be redone by the streaming kernel (fp16 range) */
int64_t lasso_b200_resident_fallbacks(void);

/*
 * ISTA / FISTA solve -- replaces the loop of lasso/linear/solvers/ista.py:57-104
 * (gradient step ista.py:71-73,90; stop test ista.py:64,93-95; momentum
 * ista.py:77-78,98-101) for backtrack=False.
 *
 *   x, weight      device, read-only
 *   z0             device [n,k] or NULL for the all-zero start (sparse_encode.py:23)
 *   z_out          device [n,k]; receives the returned code.  May alias z0.
 *   alpha, lr      python floats of the reference (double); the library rounds
 *                  lr and alpha*lr to float32 exactly as torch does
 *   maxiter        >= 0;  0 copies z0 (or zeros) to z_out
 *   fast           non-zero = FISTA momentum
 *   tol_abs        absolute threshold of the batch-global stop test, i.e. the
 *                  reference's z0.numel()*tol.  Negative disables the test.  The
 *                  streaming paths evaluate it on the device (later launches turn
 *                  into no-ops); the resident path records every iteration's sum
 *                  and replays a shorter run when the test fired before maxiter.
 *   iters_done     optional HOST pointer: number of iterations executed (forces
 *                  one stream synchronisation when non-NULL)
 *   delta_hist     optional DEVICE pointer to maxiter doubles: sum|z_i - z_{i+1}|
 *                  of every executed iteration (0 for skipped ones)
 *   path           lasso_b200_path
 */
int32_t lasso_b200_fista_f32(const float* x, const float* weight, const float* z0,
                             float* z_out, int64_t n, int32_t d, int32_t k,
                             double alpha, double lr, int32_t maxiter, int32_t fast,
                             double tol_abs, int32_t* iters_done, double* delta_hist,
                             int32_t path, void* stream);

/*
 * Same solve on HOST buffers (pinned or pageable): copies x, weight (and z0)
 * to the current device, runs lasso_b200_fista_f32 on `stream`, copies the code
 * back and synchronises the stream.  delta_hist, if given, is a HOST pointer here;
 * z_out may alias z0.  (sparse_encode.py:38-73 called with CPU tensors; what
 * dict_evaluate, dict_learning.py:16-20, does with held-out host data.)
 */
int32_t lasso_b200_fista_f32_host(const float* x, const float* weight, const float* z0,
                                  float* z_out, int64_t n, int32_t d, int32_t k,
                                  double alpha, double lr, int32_t maxiter, int32_t fast,
                                  double tol_abs, int32_t* iters_done, double* delta_hist,
                                  int32_t path, void* stream);

/*
 * Convolutional ISTA / FISTA -- replaces the loop of lasso/conv2d/ista.py:7-49 as
 * im2col -> linear: rows are the oh x ow patches of every (zero-padded) image
 * (oh = (h + 2*padding - kh) / stride + 1, ow likewise; both divisions must be exact),
 * features the cin*kh*kw patch entries, atoms the filters.
 *   conv_transpose2d(z, W) = fold(Z weight_lin^T)   (ista.py:18)
 *   conv2d(r, W)           = unfold(r) weight_lin    (ista.py:19)
 * so every iteration is the first half of the k-blocked tcgen05 kernel (R = Y weight_lin^T),
 * the residual in image space (overlap-add of the patches, minus x, re-unfolded) and its
 * second half (gradient, soft threshold, momentum).
 *   x            device [n_img, cin, h, w]
 *   weight_lin   device [cin*kh*kw, k]: weight_lin[c*kh*kw + a*kw + b, f] = W[f, c, a, b]
 *   z0, z_out    device [n_img*oh*ow, k] (codes in patch-major "NHWC" order; z0 may be NULL)
 * Limits: cin*kh*kw <= 128 and k <= 1024 (multiples of 4), one image's patch matrix <= 200 KB.
 * Other arguments as lasso_b200_fista_f32.  Synchronises the stream once.
 */
int32_t lasso_b200_conv2d_fista_f32(const float* x, const float* weight_lin, const float* z0,
                                    float* z_out, int64_t n_img, int32_t cin, int32_t h, int32_t w,
                                    int32_t kh, int32_t kw, int32_t stride, int32_t padding,
                                    int32_t k, double alpha, double lr, int32_t maxiter,
                                    int32_t fast, double tol_abs, int32_t* iters_done,
                                    double* delta_hist, void* stream);

/*
 * Lipschitz constant L = lambda_max(W^T W) -- replaces _lipschitz_constant,
 * ista.py:8-14 (Gram + D2H + ARPACK eigsh) by an on-device float64 power
 * iteration on the smaller Gram.  Synchronises; result in *l_out (host).
 */
int32_t lasso_b200_lipschitz_f32(const float* weight, int32_t d, int32_t k,
                                 int32_t iters, double* l_out, void* stream);

/*
 * Exact Lipschitz constant of the convolutional dictionary: lambda_max of conv2d^T conv2d on images
 * of size [cin, h, w] -- replaces lip_constant, lasso/conv2d/lip_const.py:8-31 (ARPACK eigsh on a host
 * LinearOperator with one conv2d + conv_transpose2d + D2H/H2D per step).  The operator is formed
 * densely in image space (cin*h*w <= 4096) from the tap Gram of the filters and goes through the same
 * lambda_max kernels as lasso_b200_lipschitz_f32.  Any kernel size / stride / padding (the reference's
 * fast bound, lip_const.py:96-135, takes odd kernels and stride 1 only).
 *   weight   device [filters, cin, kh, kw] float32      iters   cap on the power steps (2000 is plenty)
 *   l_out    HOST double.  Synchronises the stream.
 */
int32_t lasso_b200_conv2d_lipschitz_f32(const float* weight, int32_t filters, int32_t cin, int32_t kh,
                                        int32_t kw, int32_t h, int32_t w, int32_t stride,
                                        int32_t padding, int32_t iters, double* l_out, void* stream);

/*
 * Ridge warm start  z_out = ((W^T W + alpha I)^-1 W^T x^T)^T  -- initialize_code(mode='ridge'),
 * sparse_encode.py:28-29 -> ridge, utils.py:28-40 (k x k Gram, float32 Cholesky, n right-hand sides).
 * Computed as ONE [n,d] x [d,k] product z_out = x T with T = (W W^T + alpha I)^-1 W = W (W^T W +
 * alpha I)^-1 taken from the smaller of the two systems (float64 Gram, one-CTA blocked Cholesky --
 * float64 for min(d,k) <= 64, float32 beyond, like the reference's own factorisation -- and one warp per
 * right-hand side); min(d,k) <= 320.
 *   x, weight   device, read-only          z_out   device [n,k]
 *   not_positive_definite   HOST int32: set to 1 when the regularised Gram has a non-positive
 *                           pivot -- the caller raises the reference's RuntimeError (utils.py:35-38)
 * Synchronises the stream (the reference reads `info` back at the same point).
 */
int32_t lasso_b200_ridge_init_f32(const float* x, const float* weight, int64_t n, int32_t d, int32_t k,
                                  double alpha, float* z_out, int32_t* not_positive_definite,
                                  void* stream);

/*
 * z_out[n,k] = x[n,d] t[d,k], float32 (FFMA, fp32 accumulate): initialize_code(mode='transpose'),
 * sparse_encode.py:30-31 (`torch.matmul(x, weight)`), and the n-sized step of the ridge start when its
 * small system is solved by the caller.  All pointers device; asynchronous.
 */
int32_t lasso_b200_matmul_f32(const float* x, const float* t, int64_t n, int32_t d, int32_t k,
                              float* z_out, void* stream);

/*
 * Loss terms of lasso_loss, dict_learning.py:10-13:
 *   out[0] = sum (x - z weight^T)^2      out[1] = sum |z|
 * out: DEVICE pointer to 2 doubles (overwritten).  The caller forms
 * (0.5*out[0] + alpha*out[1]) / n  (after an all-reduce when sharded).
 */
int32_t lasso_b200_loss_terms_f32(const float* x, const float* z, const float* weight,
                                  int64_t n, int32_t d, int32_t k, double* out,
                                  void* stream);

/*
 * Sufficient statistics of the dictionary update (M-step):
 *   gram_zz[k,k] = z^T z     gram_zx[k,d] = z^T x      (float64, overwritten)
 * Replaces the data passes of update_dict (dict_learning.py:82-101) and
 * update_dict_ridge (dict_learning.py:117-118); these are the buffers that are
 * all-reduced across GPUs.
 */
int32_t lasso_b200_gram_f32(const float* z, const float* x, int64_t n, int32_t d,
                            int32_t k, double* gram_zz, double* gram_zx, void* stream);

/*
 * Gauss-Seidel atom sweep of update_dict (dict_learning.py:83-101) in Gram
 * space, one CTA:  u_j = B[j] - D A[:,j] + A[j,j] d_j ; d_j = u_j/|u_j|.
 *   dict      device [d,k] float32, updated in place
 *   gram_zz   device [k,k] float64 (rows/cols of re-drawn atoms are zeroed)
 *   gram_zx   device [k,d] float64
 *   redraw    device [d,k] float32 N(0,1) draws that replace degenerate atoms
 *             (|u_j| < eps, dict_learning.py:91-98), or NULL: the atom is then
 *             left for the caller to re-draw (its statistics are zeroed either way)
 *   zeroed    device [k] int32: 1 for degenerate atoms -- the caller must zero
 *             the code column z[:,j] (dict_learning.py:98)
 *   positive  non-zero: every atom (and every re-drawn atom) is clamped at zero
 *             before it is normalised (update_dict(positive=True), dict_learning.py:87-88, 94-95)
 */
int32_t lasso_b200_dict_update_gram_f32(float* dict, double* gram_zz, double* gram_zx,
                                        int32_t d, int32_t k, double eps,
                                        const float* redraw, int32_t* zeroed,
                                        int32_t positive, void* stream);

/*
 * z[:, j] = 0 for every atom j with mask[j] != 0 -- the in-place clearing of the codes of a
 * re-drawn atom (dict_learning.py:98, `Z[:, k] = 0`) for all flagged atoms in one masked pass,
 * so the host does not have to read the `zeroed` mask of lasso_b200_dict_update_gram_f32 back.
 *   z      device [n,k] float32, updated in place      mask   device [k] int32
 */
int32_t lasso_b200_zero_columns_f32(float* z, int64_t n, int32_t k, const int32_t* mask,
                                    void* stream);

/*
 * Building blocks of the slow path of ista(): backtrack=True (Beck-Teboulle line search with
 * batch-global F and Q, ista.py:17-54) and verbose=True (ista.py:80-81).  The host keeps the
 * reference's per-trial decision (one scalar read-back per trial, like `if F_next <= Q_next`).
 *
 *   gradient:  grad[n,k] = (point weight^T - x) weight ;  f_sum[0] = sum (point weight^T - x)^2
 *              (ista.py:22-24: fval_0 = 0.5 f_sum, fgrad_0 = grad)
 *   trial:     cand = softshrink(point - step grad, alpha step)          (ista.py:40)
 *              sums[0] = sum (cand weight^T - x)^2   sums[1] = sum |cand|
 *              sums[2] = sum (cand - point) grad     sums[3] = sum (cand - point)^2   (ista.py:26-35)
 *   momentum:  y = z_next + beta (z_next - z) (y may be NULL), delta[0] = sum |z - z_next|
 *              (ista.py:93, 100)
 * f_sum / sums / delta are DEVICE pointers to doubles and are overwritten.
 */
int32_t lasso_b200_gradient_f32(const float* x, const float* point, const float* weight,
                                int64_t n, int32_t d, int32_t k, float* grad, double* f_sum,
                                void* stream);
int32_t lasso_b200_linesearch_trial_f32(const float* x, const float* point, const float* grad,
                                        const float* weight, int64_t n, int32_t d, int32_t k,
                                        double step, double alpha, float* cand, double* sums,
                                        void* stream);
int32_t lasso_b200_momentum_f32(const float* z_next, const float* z, double beta, float* y,
                                int64_t count, double* delta, void* stream);

/* free the per-device private workspace (buffers are re-grown on demand) */
int32_t lasso_b200_release_workspace(void);

#ifdef __cplusplus
}
#endif
#endif /* LASSO_B200_H_ */
