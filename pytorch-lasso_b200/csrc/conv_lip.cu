// Exact Lipschitz constant of the convolutional dictionary: lambda_max(A^T A) for A = conv2d(., W)
// on images of a given size -- the quantity the reference's lip_constant computes with ARPACK on a
// host LinearOperator (lasso/conv2d/lip_const.py:8-31: one conv2d + conv_transpose2d + D2H + H2D per
// Lanczos step).
//
// A^T A and A A^T share their non-zero spectrum and the image is the smaller side (cin*h*w against
// filters*oh*ow), so the operator is formed DENSELY in image space and handed to the same
// lambda_max machinery as the dictionary's Lipschitz constant (aux_kernels.cu: power iteration on the
// 16th power of the matrix + one float64 Rayleigh quotient).  A plain power iteration on the operator
// does not do: the top of a convolution's spectrum is a cluster (relative gaps ~1e-3), 4000 steps
// still sat 2e-4 below the eigenvalue.
//
//   R[t, t']   = sum_f W[f, t] W[f, t']                          t = (c, a, b): tap Gram, (cin kh kw)^2
//   G[r, r']   = sum over the filter positions (p, q) that cover both pixels r = (c, i, j), r' = (c', i', j')
//                of R[(c, i + pad - p s, j + pad - q s), (c', i' + pad - p s, j' + pad - q s)]
#include <algorithm>

#include "common.cuh"

namespace lasso {

void small_gram_launch(const float* w, int d, int k, int m, int len, int row_gram, double* gram, cudaStream_t st);

namespace {

struct ConvGeom {
  int cin, kh, kw, h, w, stride, pad, oh, ow;
};

__global__ void __launch_bounds__(256) conv_dense_operator_kernel(const double* __restrict__ tap_gram, ConvGeom g,
                                                                  double* __restrict__ out) {
  const int npix = g.cin * g.h * g.w, taps = g.cin * g.kh * g.kw;
  const int r = blockIdx.x;
  const int c = r / (g.h * g.w), i = (r / g.w) % g.h, j = r % g.w;
  for (int s = threadIdx.x; s < npix; s += blockDim.x) {
    const int c2 = s / (g.h * g.w), i2 = (s / g.w) % g.h, j2 = s % g.w;
    double acc = 0.0;
    if (abs(i - i2) < g.kh && abs(j - j2) < g.kw) {
      for (int p = 0; p < g.oh; ++p) {
        const int a = i + g.pad - p * g.stride, a2 = i2 + g.pad - p * g.stride;
        if (a < 0 || a >= g.kh || a2 < 0 || a2 >= g.kh) continue;
        for (int q = 0; q < g.ow; ++q) {
          const int b = j + g.pad - q * g.stride, b2 = j2 + g.pad - q * g.stride;
          if (b < 0 || b >= g.kw || b2 < 0 || b2 >= g.kw) continue;
          acc += tap_gram[(int64_t)((c * g.kh + a) * g.kw + b) * taps + (c2 * g.kh + a2) * g.kw + b2];
        }
      }
    }
    out[(int64_t)r * npix + s] = acc;
  }
}

}  // namespace

size_t conv_lipschitz_scratch_bytes(int cin, int kh, int kw, int h, int w) {
  const size_t taps = (size_t)cin * kh * kw;
  return sizeof(double) * (lambda_max_scratch_doubles(cin * h * w) + taps * taps);
}

// lambda_max of conv2d^T conv2d on [cin, h, w] images; weight [filters, cin, kh, kw]; scratch:
// conv_lipschitz_scratch_bytes(...) of device memory.  Synchronises.
int conv_lipschitz_run(const float* weight, int f, int cin, int kh, int kw, int h, int w, int stride, int pad,
                       int iters, double* l_out, void* scratch_mem, cudaStream_t st) {
  ConvGeom g{cin, kh, kw, h, w, stride, pad, (h + 2 * pad - kh) / stride + 1, (w + 2 * pad - kw) / stride + 1};
  const int npix = cin * h * w, taps = cin * kh * kw;
  if (npix > 4096 || scratch_mem == nullptr) {
    set_error("conv lipschitz: %d x %d x %d image: the dense operator is built for cin*h*w <= 4096", cin, h, w);
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  double* scratch = reinterpret_cast<double*>(scratch_mem);
  double* tap_gram = scratch + lambda_max_scratch_doubles(npix);
  // weight as a [filters, taps] matrix: tap Gram = its k x k Gram (row_gram = 0)
  small_gram_launch(weight, f, taps, taps, f, 0, tap_gram, st);
  conv_dense_operator_kernel<<<npix, 256, 0, st>>>(tap_gram, g, scratch);
  LASSO_CHECK_LAUNCH();
  count_launch();
  double* l_dev = scratch + (size_t)npix * npix + 2 * (size_t)npix;
  int rc = lambda_max_run(scratch, npix, iters, l_dev, st);
  if (rc == LASSO_B200_OK) {
    cudaError_t e = cudaMemcpyAsync(l_out, l_dev, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error("conv lipschitz: %s", cudaGetErrorString(e));
      rc = LASSO_B200_ERR_CUDA;
    }
  }
  return rc;
}

}  // namespace lasso
