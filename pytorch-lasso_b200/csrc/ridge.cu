// Ridge warm start  z0 = ((W^T W + alpha I)^-1 W^T x^T)^T  -- sparse_encode.py:28-29, utils.py:28-40
// (the start the notebook uses with ISTA, examples/dict_learning_omniglot.ipynb:638,1027).
//
// The reference forms the k x k Gram, factors it in float32 (Cholesky), multiplies W^T x^T (an
// [k, n] matrix) and solves n right-hand sides.  Here the n-sized work is ONE GEMM:
//
//     z0 = x T,   T = W (W^T W + alpha I)^-1 = (W W^T + alpha I)^-1 W      [d, k]
//
// and T comes from the SMALLER of the two systems (push-through identity): m = min(d, k).
//   K-r1  small_gram_kernel (aux_kernels.cu, shared with the Lipschitz constant): S = W W^T or W^T W, float64
//   K-r2  ridge_factor_kernel: S + alpha I = L L^T, blocked right-looking Cholesky, float64, one CTA
//         (panel in shared memory, trailing matrix in L2); a non-positive pivot raises the flag the
//         host turns into the reference's RuntimeError (utils.py:35-38)
//   K-r3  ridge_solve_kernel: one warp per right-hand side (a column of W / a row of W): forward and
//         backward substitution with the dot products split over the lanes -> T, rounded to float32
//   K-r4  ridge_apply_kernel: z0 = x T, float32 FFMA GEMM (128 x 64 tiles, 8 x 4 per thread)
#include <algorithm>

#include "common.cuh"

namespace lasso {

// defined in aux_kernels.cu
void small_gram_launch(const float* w, int d, int k, int m, int len, int row_gram, double* gram, cudaStream_t st);

namespace {

constexpr int kNB = 32;   // Cholesky block size

// gram (m x m, row-major, float64) + alpha I -> L (lower triangle of a, type T).  flag[0] = 1 on a
// pivot <= 0.  T = double up to m = 64 (cheap there); float beyond: this GPU retires ~3 float64 FMAs
// per clock and SM (1.36 ms for the 8 M of a 289 x 289 factorisation, measured), and float32 is the
// precision the reference factors in (utils.py:34, torch.linalg.cholesky_ex on a float32 Gram).
// Per block column of 32: the diagonal block is factored and inverted by one warp (a row per lane, in
// registers), the panel below becomes A21 L11^-T as a small product, the trailing matrix takes the
// rank-32 update tile by tile.
constexpr int kPS = kNB + 1;   // shared-memory row pitch
template <typename T>
__global__ void __launch_bounds__(1024) ridge_factor_kernel(const double* __restrict__ gram, T* __restrict__ a,
                                                            int m, double alpha, int* __restrict__ flag) {
  extern __shared__ __align__(16) unsigned char sh_raw[];
  T* diag = reinterpret_cast<T*>(sh_raw);   // [32][33]  L11
  T* inv = diag + kNB * kPS;                // [32][33]  L11^-1
  T* panel = inv + kNB * kPS;               // [rows below][33]  A21, then L21
  const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31;
  for (int e = tid; e < m * m; e += nt) {
    const int r = e / m, c = e % m;
    a[e] = (T)(gram[e] + (r == c ? alpha : 0.0));
  }
  __syncthreads();
  for (int j0 = 0; j0 < m; j0 += kNB) {
    const int nb = min(kNB, m - j0), below = m - j0 - nb;
    // (1) diagonal block (padded with the identity) -> one warp: Cholesky, then the inverse
    if (tid < 32) {
      T row[kNB];
#pragma unroll
      for (int c = 0; c < kNB; ++c)
        row[c] = (lane < nb && c < nb) ? a[(int64_t)(j0 + lane) * m + j0 + c] : (T)(lane == c ? 1 : 0);
#pragma unroll
      for (int c = 0; c < kNB; ++c) {
        T piv = __shfl_sync(0xffffffffu, row[c], c);
        if (!(piv > (T)0)) {
          if (lane == 0 && c < nb) flag[0] = 1;
          piv = (T)1;
        }
        const T l = sqrt(piv), il = (T)1 / l;
        if (lane == c) row[c] = l;
        else if (lane > c) row[c] *= il;
#pragma unroll
        for (int cc = c + 1; cc < kNB; ++cc) {
          const T lc = __shfl_sync(0xffffffffu, row[c], cc);   // L[cc][c]
          if (lane >= cc) row[cc] -= row[c] * lc;
        }
      }
#pragma unroll
      for (int c = 0; c < kNB; ++c) diag[lane * kPS + c] = c <= lane ? row[c] : (T)0;
      __syncwarp();
      // column `lane` of X = L11^-1 by forward substitution (all lanes read the same L entries: broadcasts)
      T x[kNB];
#pragma unroll
      for (int i = 0; i < kNB; ++i) {
        T s = (T)(i == lane ? 1 : 0);
#pragma unroll
        for (int t = 0; t < i; ++t) s -= diag[i * kPS + t] * x[t];
        x[i] = s / diag[i * kPS + i];
      }
#pragma unroll
      for (int i = 0; i < kNB; ++i) inv[i * kPS + lane] = x[i];
    }
    // meanwhile everybody stages the panel A21 (columns beyond the block: zero)
    for (int e = tid; e < below * kNB; e += nt) {
      const int r = e >> 5, c = e & 31;
      panel[r * kPS + c] = c < nb ? a[(int64_t)(j0 + nb + r) * m + j0 + c] : (T)0;
    }
    __syncthreads();
    for (int e = tid; e < nb * nb; e += nt) {
      const int r = e / nb, c = e % nb;
      if (c <= r) a[(int64_t)(j0 + r) * m + j0 + c] = diag[r * kPS + c];
    }
    // (2) L21 = A21 L11^-T:  L21[r][c] = sum_{t <= c} A21[r][t] X[c][t]; a warp owns a row
    for (int r = tid >> 5; r < below; r += nt >> 5) {
      T s = (T)0;
      for (int t = 0; t <= lane; ++t) s = fma(panel[r * kPS + t], inv[lane * kPS + t], s);
      __syncwarp();
      panel[r * kPS + lane] = s;
      if (lane < nb) a[(int64_t)(j0 + nb + r) * m + j0 + lane] = s;
    }
    __syncthreads();
    // (3) trailing matrix (lower triangle): A22 -= L21 L21^T in 32 x 32 tiles, thread (ty, tx) per entry
    const int ntile = (below + kNB - 1) / kNB, ty = tid >> 5, tx = lane;
    for (int ti = 0; ti < ntile; ++ti)
      for (int tj = 0; tj <= ti; ++tj) {
        const int r = ti * kNB + ty, c = tj * kNB + tx;
        if (r < below && c <= r) {
          T s = (T)0;
#pragma unroll 8
          for (int t = 0; t < kNB; ++t) s = fma(panel[r * kPS + t], panel[c * kPS + t], s);
          a[(int64_t)(j0 + nb + r) * m + j0 + nb + c] -= s;
        }
      }
    __syncthreads();
  }
}

template <typename T>
__device__ __forceinline__ T warp_sum_t(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp per right-hand side, the factor in shared memory (packed lower triangle).  d <= k (by_col = 1):
// rhs c = column c of W (d entries), solution -> T[:, c].  d > k: rhs c = row c of W, solution -> T[c, :].
template <typename T>
__global__ void __launch_bounds__(1024) ridge_solve_kernel(const T* __restrict__ l, int m,
                                                           const float* __restrict__ w, int d, int k, int by_col,
                                                           float* __restrict__ t_out) {
  extern __shared__ __align__(16) unsigned char ysh_raw[];
  T* lp = reinterpret_cast<T*>(ysh_raw);                 // packed: L[i][j] at i (i + 1) / 2 + j
  T* idiag = lp + (size_t)m * (m + 1) / 2;               // 1 / L[i][i]
  T* ys = idiag + m;                                     // [warps][m]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int i = warp; i < m; i += nwarp) {
    for (int j = lane; j <= i; j += 32) lp[(size_t)i * (i + 1) / 2 + j] = l[(int64_t)i * m + j];
    if (lane == 0) idiag[i] = (T)1 / l[(int64_t)i * m + i];
  }
  __syncthreads();
  const int nrhs = by_col ? k : d;
  T* y = ys + (size_t)warp * m;
  for (int c = blockIdx.x * nwarp + warp; c < nrhs; c += gridDim.x * nwarp) {
    for (int i = lane; i < m; i += 32) y[i] = by_col ? (T)w[(int64_t)i * k + c] : (T)w[(int64_t)c * k + i];
    __syncwarp();
    for (int i = 0; i < m; ++i) {          // L y' = b
      const T* row = lp + (size_t)i * (i + 1) / 2;
      T s = (T)0;
      for (int j = lane; j < i; j += 32) s = fma(row[j], y[j], s);
      s = warp_sum_t(s);
      if (lane == 0) y[i] = (y[i] - s) * idiag[i];
      __syncwarp();
    }
    for (int i = m - 1; i >= 0; --i) {     // L^T t = y'
      T s = (T)0;
      for (int j = i + 1 + lane; j < m; j += 32) s = fma(lp[(size_t)j * (j + 1) / 2 + i], y[j], s);
      s = warp_sum_t(s);
      if (lane == 0) y[i] = (y[i] - s) * idiag[i];
      __syncwarp();
    }
    for (int i = lane; i < m; i += 32) {
      if (by_col) t_out[(int64_t)i * k + c] = (float)y[i];
      else t_out[(int64_t)c * k + i] = (float)y[i];
    }
    __syncwarp();
  }
}

// z0[n, k] = x[n, d] T[d, k], float32, 128 x 64 tiles, 256 threads, 8 x 4 outputs per thread.  The next 16-deep
// tile travels global -> registers while the current one is multiplied out of shared memory, and the products read
// shared memory as 128-bit vectors (3 loads per 32 FMAs).  The first version loaded and multiplied in turn, scalar:
// 18 TFLOP/s at n = 10000, d = 289, k = 300.
constexpr int kRM = 128, kRN = 64, kRK = 16;
__global__ void __launch_bounds__(256) ridge_apply_kernel(const float* __restrict__ x, const float* __restrict__ t,
                                                          float* __restrict__ z, int64_t n, int d, int k) {
  __shared__ __align__(16) float xs[kRK][kRM + 4];   // transposed x tile
  __shared__ __align__(16) float ts[kRK][kRN];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // 16 x 16 threads
  const int64_t row0 = (int64_t)blockIdx.y * kRM;
  const int col0 = blockIdx.x * kRN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float px[8], pt[4];
  auto fetch = [&](int c0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + 256 * q, r = e / kRK, c = e % kRK;
      const int64_t gr = row0 + r;
      px[q] = (gr < n && c0 + c < d) ? __ldg(x + gr * d + c0 + c) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 256 * q, r = e / kRN, c = e % kRN;
      pt[q] = (c0 + r < d && col0 + c < k) ? __ldg(t + (int64_t)(c0 + r) * k + col0 + c) : 0.f;
    }
  };
  fetch(0);
  for (int c0 = 0; c0 < d; c0 += kRK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = tid + 256 * q;
      xs[e % kRK][e / kRK] = px[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + 256 * q;
      ts[e / kRN][e % kRN] = pt[q];
    }
    __syncthreads();
    if (c0 + kRK < d) fetch(c0 + kRK);
#pragma unroll
    for (int c = 0; c < kRK; ++c) {
      const float4 a0 = *reinterpret_cast<const float4*>(&xs[c][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&xs[c][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&ts[c][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gr = row0 + ty * 8 + i;
    if (gr >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gc = col0 + tx * 4 + j;
      if (gc < k) z[gr * k + gc] = acc[i][j];
    }
  }
}

}  // namespace

// z[n, k] = x[n, d] t[d, k] (float32 FFMA): the n-sized step of the ridge start, also the 'transpose' start
int matmul_run(const float* x, const float* t, int64_t n, int d, int k, float* z, cudaStream_t st) {
  if (n == 0) return LASSO_B200_OK;
  dim3 grid((k + kRN - 1) / kRN, (unsigned)((n + kRM - 1) / kRM));
  ridge_apply_kernel<<<grid, 256, 0, st>>>(x, t, z, n, d, k);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

size_t ridge_scratch_bytes(int d, int k) {
  const size_t m = (size_t)std::min(d, k);
  return sizeof(double) * m * m * 2 + sizeof(float) * (size_t)d * k + 64;
}

template <typename T>
static int ridge_factor_and_solve(const double* gram, void* work, int m, double alpha, int* flag, const float* w,
                                  int d, int k, int by_col, float* t, cudaStream_t st) {
  T* a = reinterpret_cast<T*>(work);
  const size_t fsmem = sizeof(T) * ((size_t)2 * kNB * kPS + (size_t)std::max(m - kNB, 1) * kPS);
  // solve: packed factor + 1 / diagonal + one vector per warp; as many warps as fit (the right-hand sides
  // are independent, every CTA stages its own copy of the factor)
  const size_t packed = sizeof(T) * ((size_t)m * (m + 1) / 2 + m);
  int warps = 32;
  while (warps > 1 && packed + sizeof(T) * (size_t)warps * m > 220 * 1024) warps >>= 1;
  const size_t ssmem = packed + sizeof(T) * (size_t)warps * m;
  if (fsmem > 220 * 1024 || ssmem > 220 * 1024) {
    set_error("ridge init: min(d, k) = %d is beyond the shared-memory factorisation (<= 320)", m);
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  if (fsmem > 48 * 1024)
    LASSO_CUDA_TRY(cudaFuncSetAttribute(ridge_factor_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
  ridge_factor_kernel<T><<<1, 1024, fsmem, st>>>(gram, a, m, alpha, flag);
  const int nrhs = by_col ? k : d;
  if (ssmem > 48 * 1024)
    LASSO_CUDA_TRY(cudaFuncSetAttribute(ridge_solve_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssmem));
  ridge_solve_kernel<T><<<(nrhs + warps - 1) / warps, warps * 32, ssmem, st>>>(a, m, w, d, k, by_col, t);
  return LASSO_B200_OK;
}

// scratch: ridge_scratch_bytes(d, k) bytes of device memory (8-byte aligned).  *not_pd = 1 when the
// regularised Gram is not positive definite (utils.py:35-38).  Synchronises.
int ridge_init_run(const float* x, const float* w, int64_t n, int d, int k, double alpha, float* z_out,
                   void* scratch, int* not_pd, cudaStream_t st) {
  const int by_col = d <= k ? 1 : 0;       // which Gram: W W^T (d x d) or W^T W (k x k)
  const int m = by_col ? d : k, len = by_col ? k : d;
  double* gram = reinterpret_cast<double*>(scratch);
  void* work = gram + (size_t)m * m;
  float* t = reinterpret_cast<float*>(gram + 2 * (size_t)m * m);
  int* flag = reinterpret_cast<int*>(t + (size_t)d * k);
  LASSO_CUDA_TRY(cudaMemsetAsync(flag, 0, sizeof(int), st));
  small_gram_launch(w, d, k, m, len, by_col, gram, st);
  int rc = m <= 64 ? ridge_factor_and_solve<double>(gram, work, m, alpha, flag, w, d, k, by_col, t, st)
                   : ridge_factor_and_solve<float>(gram, work, m, alpha, flag, w, d, k, by_col, t, st);
  if (rc) return rc;
  if (n > 0) {
    dim3 grid((k + kRN - 1) / kRN, (unsigned)((n + kRM - 1) / kRM));
    ridge_apply_kernel<<<grid, 256, 0, st>>>(x, t, z_out, n, d, k);
  }
  LASSO_CHECK_LAUNCH();
  count_launch(n > 0 ? 3 : 2);
  int host_flag = 0;
  LASSO_CUDA_TRY(cudaMemcpyAsync(&host_flag, flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  LASSO_CUDA_TRY(cudaStreamSynchronize(st));
  *not_pd = host_flag;
  return LASSO_B200_OK;
}

}  // namespace lasso
