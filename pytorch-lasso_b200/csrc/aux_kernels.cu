// Small kernels around the FISTA loop: Lipschitz constant (power iteration),
// M-step sufficient statistics (Gram matrices) and the Gram-space atom sweep.
#include <cstdlib>

#include "common.cuh"

namespace lasso {
namespace {

// ---------------------------------------------------------------------------
// K2: L = lambda_max(W^T W) = lambda_max(W W^T)        (ista.py:8-14)
// ---------------------------------------------------------------------------
// small symmetric Gram in float64: m = min(d,k); M = W W^T (d<=k) or W^T W.
__global__ void small_gram_kernel(const float* __restrict__ w, int d, int k, int m, int len,
                                  int row_gram, double* __restrict__ gram) {
  __shared__ double ta[16][17], tb[16][17];
  const int a = blockIdx.y * 16 + threadIdx.y;
  const int b = blockIdx.x * 16 + threadIdx.x;
  double acc = 0.0;
  for (int c0 = 0; c0 < len; c0 += 16) {
    // element (row r of the Gram, contraction index c)
    const int ca = c0 + threadIdx.x;
    const int ra = blockIdx.y * 16 + threadIdx.y;
    const int rb = blockIdx.x * 16 + threadIdx.y;
    double va = 0.0, vb = 0.0;
    if (ca < len) {
      if (ra < m) va = row_gram ? (double)w[(int64_t)ra * k + ca] : (double)w[(int64_t)ca * k + ra];
      if (rb < m) vb = row_gram ? (double)w[(int64_t)rb * k + ca] : (double)w[(int64_t)ca * k + rb];
    }
    ta[threadIdx.y][threadIdx.x] = va;
    tb[threadIdx.y][threadIdx.x] = vb;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 16; ++c) acc += ta[threadIdx.y][c] * tb[threadIdx.x][c];
    __syncthreads();
  }
  if (a < m && b < m) gram[(int64_t)a * m + b] = acc;
}

// one CTA: normalised power iteration on the m x m Gram, Rayleigh quotient out.
__global__ void __launch_bounds__(1024) power_iter_kernel(const double* __restrict__ gram, int m,
                                                          int iters, double* __restrict__ vec,
                                                          double* __restrict__ out) {
  __shared__ double red[32];
  __shared__ double s_norm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  double* v = vec;       // [m]
  double* u = vec + m;   // [m]
  for (int a = tid; a < m; a += blockDim.x) {
    // deterministic, non-symmetric start vector
    unsigned h = (unsigned)a * 2654435761u;
    v[a] = 1.0 + 0.25 * (double)((h >> 8) & 0xffff) / 65536.0;
  }
  __syncthreads();
  double lambda = 0.0, prev = -1.0;
  int stable = 0;
  for (int it = 0; it < iters; ++it) {
    for (int a = warp; a < m; a += nwarps) {
      double s = 0.0;
      for (int c = lane; c < m; c += 32) s += gram[(int64_t)a * m + c] * v[c];
      s = warp_sum(s);
      if (lane == 0) u[a] = s;
    }
    __syncthreads();
    double part = 0.0, dotp = 0.0;
    for (int a = tid; a < m; a += blockDim.x) {
      part += u[a] * u[a];
      dotp += u[a] * v[a];
    }
    part = warp_sum(part);
    dotp = warp_sum(dotp);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int wi = 0; wi < nwarps; ++wi) s += red[wi];
      s_norm = s;
    }
    __syncthreads();
    const double nn = s_norm;
    if (lane == 0) red[warp] = dotp;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int wi = 0; wi < nwarps; ++wi) s += red[wi];
      red[0] = s;
    }
    __syncthreads();
    // v was unit-norm (after the first step), so v^T M v = <u, v> is the Rayleigh quotient
    const double vv_dot = red[0];
    const double nrm = sqrt(nn);
    __syncthreads();
    if (nrm == 0.0) {
      lambda = 0.0;
      break;
    }
    for (int a = tid; a < m; a += blockDim.x) v[a] = u[a] / nrm;
    __syncthreads();
    if (it > 0) {
      lambda = vv_dot;
      // the Rayleigh quotient converges quadratically in the vector error: once it moves by
      // less than 1e-13 relative for 4 steps in a row it is good to ~1e-12
      if (fabs(lambda - prev) <= 1e-13 * fabs(lambda)) {
        if (++stable >= 4) break;
      } else {
        stable = 0;
      }
      prev = lambda;
    }
  }
  if (tid == 0) out[0] = lambda;
}

// m <= 64 (the usual case: d = 64): the Gram lives in shared memory, one thread per row, two
// warps; an iteration is a 64-term dot product per thread and one block barrier.
__global__ void __launch_bounds__(64) power_iter_small_kernel(const double* __restrict__ gram,
                                                              int m, int iters,
                                                              double* __restrict__ out) {
  __shared__ double M[64 * 64];   // M[c * 64 + a] = gram[a][c] (symmetric): conflict-free reads
  __shared__ double v[2][64];
  __shared__ double part[2][2][2];
  const int a = threadIdx.x, lane = a & 31, warp = a >> 5;
  for (int e = a; e < 64 * 64; e += 64) {
    const int c = e >> 6, r = e & 63;
    M[e] = (r < m && c < m) ? gram[(int64_t)r * m + c] : 0.0;
  }
  {
    unsigned h = (unsigned)a * 2654435761u;
    v[0][a] = a < m ? 1.0 + 0.25 * (double)((h >> 8) & 0xffff) / 65536.0 : 0.0;
  }
  __syncthreads();
  double lambda = 0.0, prev = -1.0;
  int stable = 0, cur = 0;
  for (int it = 0; it < iters; ++it) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 4
    for (int c = 0; c < 64; c += 4) {
      s0 = fma(M[(c + 0) * 64 + a], v[cur][c + 0], s0);
      s1 = fma(M[(c + 1) * 64 + a], v[cur][c + 1], s1);
      s2 = fma(M[(c + 2) * 64 + a], v[cur][c + 2], s2);
      s3 = fma(M[(c + 3) * 64 + a], v[cur][c + 3], s3);
    }
    const double u = (s0 + s1) + (s2 + s3);
    const double uu = warp_sum(u * u), uv = warp_sum(u * v[cur][a]);
    if (lane == 0) {
      part[it & 1][warp][0] = uu;
      part[it & 1][warp][1] = uv;
    }
    __syncthreads();
    const double nn = part[it & 1][0][0] + part[it & 1][1][0];
    const double rayleigh = part[it & 1][0][1] + part[it & 1][1][1];   // v unit-norm for it > 0
    const double nrm = sqrt(nn);
    if (nrm == 0.0) {
      lambda = 0.0;
      break;
    }
    v[cur ^ 1][a] = u / nrm;
    cur ^= 1;
    __syncthreads();
    if (it > 0) {
      lambda = rayleigh;
      if (fabs(lambda - prev) <= 1e-13 * fabs(lambda)) {
        if (++stable >= 4) break;
      } else {
        stable = 0;
      }
      prev = lambda;
    }
  }
  if (a == 0) out[0] = lambda;
}

// ---------------------------------------------------------------------------
// K3: gram_zz = Z^T Z (k x k), gram_zx = Z^T X (k x d), float64 outputs
// ---------------------------------------------------------------------------
constexpr int kGT = 64;      // output tile
constexpr int kGR = 32;      // rows staged per step
constexpr int kGSlab = 512;  // rows per CTA

__global__ void __launch_bounds__(256) gram_kernel(const float* __restrict__ z,
                                                   const float* __restrict__ x, int64_t n, int d,
                                                   int k, double* __restrict__ gzz,
                                                   double* __restrict__ gzx) {
  __shared__ __align__(16) float As[kGR][kGT + 4];
  __shared__ __align__(16) float Bs[kGR][kGT + 4];
  const int nbj = (k + d + kGT - 1) / kGT;
  const int bi = blockIdx.y / nbj, bj = blockIdx.y % nbj;
  const int a0 = bi * kGT;   // atom block
  const int c0 = bj * kGT;   // column block of [Z | X]
  // Z^T Z is symmetric: among the full 64-atom blocks, those below the diagonal are mirrored from
  // the ones above
  const int nzb = k / kGT;
  if (bj < bi && bi < nzb) return;
  const bool mirror = bi < bj && bj < nzb;
  const int tid = threadIdx.x, ta = tid >> 4, tb = tid & 15;
  const int64_t r_begin = (int64_t)blockIdx.x * kGSlab;
  const int64_t r_end = min(n, r_begin + kGSlab);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t r0 = r_begin; r0 < r_end; r0 += kGR) {
    __syncthreads();
    for (int e = tid; e < kGR * kGT; e += 256) {
      const int rr = e / kGT, cc = e % kGT;
      const int64_t gr = r0 + rr;
      float va = 0.f, vb = 0.f;
      if (gr < r_end) {
        if (a0 + cc < k) va = z[gr * k + a0 + cc];
        const int c = c0 + cc;
        if (c < k) vb = z[gr * k + c];
        else if (c < k + d) vb = x[gr * d + (c - k)];
      }
      As[rr][cc] = va;
      Bs[rr][cc] = vb;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < kGR; ++rr) {
      const float4 a = *reinterpret_cast<const float4*>(&As[rr][ta * 4]);
      // lasso codes are mostly zeros: skip the row when the warp's two atom quads are empty
      if (__all_sync(0xffffffffu, a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 0.f)) continue;
      const float4 b = *reinterpret_cast<const float4*>(&Bs[rr][tb * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int a = a0 + ta * 4 + i, c = c0 + tb * 4 + j;
      if (a >= k || acc[i][j] == 0.f) continue;
      if (c < k) {
        atomicAdd(&gzz[(int64_t)a * k + c], (double)acc[i][j]);
        if (mirror) atomicAdd(&gzz[(int64_t)c * k + a], (double)acc[i][j]);
      } else if (c < k + d) {
        atomicAdd(&gzx[(int64_t)a * d + (c - k)], (double)acc[i][j]);
      }
    }
}

// ---------------------------------------------------------------------------
// K4: Gauss-Seidel atom sweep in Gram space, one CTA   (dict_learning.py:83-101)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) dict_sweep_kernel(float* dict, double* gzz, double* gzx,
                                                          int d, int k, double eps,
                                                          const float* __restrict__ redraw,
                                                          int* __restrict__ zeroed,
                                                          double* __restrict__ u) {
  __shared__ double red[32];
  __shared__ double s_val;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int j = 0; j < k; ++j) {
    const double ajj = gzz[(int64_t)j * k + j];
    // u_i = B[j,i] - sum_l D[i,l] A[l,j] + A[j,j] D[i,j]      (A symmetric: A[l,j] = A[j,l])
    for (int i = warp; i < d; i += nwarps) {
      double s = 0.0;
      for (int l = lane; l < k; l += 32)
        s += (double)dict[(int64_t)i * k + l] * gzz[(int64_t)j * k + l];
      s = warp_sum(s);
      if (lane == 0) u[i] = gzx[(int64_t)j * d + i] - s + ajj * (double)dict[(int64_t)i * k + j];
    }
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < d; i += blockDim.x) part += u[i] * u[i];
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int wi = 0; wi < nwarps; ++wi) s += red[wi];
      s_val = sqrt(s);
    }
    __syncthreads();
    double nrm = s_val;
    const bool degenerate = nrm < eps;
    if (degenerate) {
      // the atom's codes are dropped (dict_learning.py:92-98): its row/column of the
      // statistics vanish, so the replacement never influences the later atoms
      for (int l = tid; l < k; l += blockDim.x) {
        gzz[(int64_t)j * k + l] = 0.0;
        gzz[(int64_t)l * k + j] = 0.0;
      }
      for (int i = tid; i < d; i += blockDim.x) gzx[(int64_t)j * d + i] = 0.0;
      if (tid == 0) zeroed[j] = 1;
      nrm = 0.0;
      if (redraw != nullptr) {
        __syncthreads();
        part = 0.0;
        for (int i = tid; i < d; i += blockDim.x) {
          const double r = (double)redraw[(int64_t)i * k + j];
          u[i] = r;
          part += r * r;
        }
        part = warp_sum(part);
        if (lane == 0) red[warp] = part;
        __syncthreads();
        if (tid == 0) {
          double s = 0.0;
          for (int wi = 0; wi < nwarps; ++wi) s += red[wi];
          s_val = sqrt(s);
        }
        __syncthreads();
        nrm = s_val;
      }
    } else if (tid == 0) {
      zeroed[j] = 0;
    }
    if (nrm > 0.0) {
      for (int i = tid; i < d; i += blockDim.x) dict[(int64_t)i * k + j] = (float)(u[i] / nrm);
    }
    __syncthreads();
  }
}


// Fast variant for dictionaries that fit shared memory (d * k floats <= 150 KB, k <= 256).
//  * the dictionary lives in shared memory for the whole sweep; the rows of A = Z^T Z, B = Z^T X
//    that step j needs are prefetched three steps ahead with cp.async (no global load on the
//    critical path of the 256-step chain);
//  * the diagonal term cancels analytically: u_i = B[j,i] - sum_{l != j} D[i,l] A[j,l], so the
//    two large contributions A[j,j] D[i,j] never meet and the inner products can run in float32
//    (as the reference's own update does, dict_learning.py:85-86) -- the float64 pipe of this GPU
//    issues only a few operations per clock and was the bottleneck of the float64 version;
//    u, its norm and the normalisation stay in float64;
//  * rows/columns of dropped atoms are masked on read (the prefetched copies may predate their
//    clearing in global memory).
constexpr int kSweepMaxPerLane = 8;    // k <= 256
constexpr int kSweepDepth = 4;         // row buffers
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__global__ void __launch_bounds__(1024) dict_sweep_smem_kernel(float* dict, double* gzz, double* gzx, int d,
                                                               int k, double eps,
                                                               const float* __restrict__ redraw,
                                                               int* __restrict__ zeroed) {
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  double* arow = reinterpret_cast<double*>(sweep_smem);                               // [depth][k]
  double* brow = arow + kSweepDepth * k;                                              // [depth][d]
  double* us = brow + kSweepDepth * d;                                                // [d]
  float* ds = reinterpret_cast<float*>(us + d);                                       // [d][k]
  float* af = ds + (size_t)d * k;                                                     // [k]
  unsigned char* dead = reinterpret_cast<unsigned char*>(af + k);                     // [k]
  __shared__ double s_inv, s_nrm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int a_chunks = k / 2, b_chunks = d / 2;   // 16-byte pieces of a row of A / B (k, d even)
  auto prefetch = [&](int j) {
    if (j < k) {
      const int slot = j % kSweepDepth;
      if (tid < a_chunks) cp_async16(arow + slot * k + tid * 2, gzz + (int64_t)j * k + tid * 2);
      else if (tid < a_chunks + b_chunks)
        cp_async16(brow + slot * d + (tid - a_chunks) * 2, gzx + (int64_t)j * d + (tid - a_chunks) * 2);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int e = tid; e < d * k; e += blockDim.x) ds[e] = dict[e];
  for (int l = tid; l < k; l += blockDim.x) dead[l] = 0;
  for (int j = 0; j < kSweepDepth - 1; ++j) prefetch(j);
  for (int j = 0; j < k; ++j) {
    prefetch(j + kSweepDepth - 1);                                   // into the slot step j - 1 released
    asm volatile("cp.async.wait_group %0;" ::"n"(kSweepDepth - 1) : "memory");   // row j has landed
    __syncthreads();
    const double* aj = arow + (j % kSweepDepth) * k;
    const double* bj = brow + (j % kSweepDepth) * d;
    for (int l = tid; l < k; l += blockDim.x) af[l] = (l == j || dead[l]) ? 0.f : (float)aj[l];
    __syncthreads();
    // u_i = B[j,i] - sum_{l != j} D[i,l] A[j,l]; two rows per warp in flight
    for (int i0 = warp; i0 < d; i0 += 2 * nwarps) {
      const int i1 = i0 + nwarps;
      const bool two = i1 < d;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int m = 0; m < kSweepMaxPerLane; ++m) {
        const int l = lane + 32 * m;
        if (l < k) {
          const float a = af[l];
          s0 = fmaf(ds[i0 * k + l], a, s0);
          if (two) s1 = fmaf(ds[i1 * k + l], a, s1);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (lane == 0) {
        us[i0] = bj[i0] - (double)s0;
        if (two) us[i1] = bj[i1] - (double)s1;
      }
    }
    __syncthreads();
    // one warp forms the norm (float64 is scarce on this GPU: 32 warps doing it redundantly cost
    // more than a barrier).  1 / |u| comes from a float rsqrt refined by three Newton steps in
    // float64 (full double accuracy); the library sqrt and divide are long software sequences
    if (warp == 0) {
      double part = 0.0;
      for (int i = lane; i < d; i += 32) part += us[i] * us[i];
      const double ss = warp_sum(part);
      double inv = 0.0;
      if (ss > 0.0 && ss < 1e300) {
        if (ss < 1e-30 || ss > 1e30) {
          inv = 1.0 / sqrt(ss);                             // out of float range: slow exact path
        } else {
          inv = (double)rsqrtf((float)ss);
#pragma unroll
          for (int t = 0; t < 3; ++t) inv = inv * (1.5 - 0.5 * ss * inv * inv);
        }
      }
      if (lane == 0) {
        s_inv = inv;
        s_nrm = ss * inv;      // |u| (0 for an all-zero u)
      }
    }
    __syncthreads();
    double inv = s_inv, nrm = s_nrm, part;
    const bool degenerate = nrm < eps;
    if (tid == 0) zeroed[j] = degenerate ? 1 : 0;
    if (degenerate) {
      // the atom's codes are dropped (dict_learning.py:92-98): its row/column of the statistics
      // vanish, so the replacement never influences the later atoms
      for (int l = tid; l < k; l += blockDim.x) {
        gzz[(int64_t)j * k + l] = 0.0;
        gzz[(int64_t)l * k + j] = 0.0;
      }
      for (int i = tid; i < d; i += blockDim.x) gzx[(int64_t)j * d + i] = 0.0;
      if (tid == 0) dead[j] = 1;
      nrm = 0.0;
      if (redraw != nullptr) {
        __syncthreads();
        for (int i = tid; i < d; i += blockDim.x) us[i] = (double)redraw[(int64_t)i * k + j];
        __syncthreads();
        part = 0.0;
        for (int i = lane; i < d; i += 32) part += us[i] * us[i];
        nrm = sqrt(warp_sum(part));
        inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
      }
    }
    if (nrm > 0.0) {
      for (int i = tid; i < d; i += blockDim.x) {
        const float v = (float)(us[i] * inv);
        ds[i * k + j] = v;
        dict[(int64_t)i * k + j] = v;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int lipschitz_run(const float* w, int d, int k, int iters, double* l_dev, double* scratch,
                  cudaStream_t st) {
  // scratch: [m*m] Gram + [2*m] vectors
  const int row_gram = d <= k ? 1 : 0;
  const int m = row_gram ? d : k;
  const int len = row_gram ? k : d;
  dim3 grid((m + 15) / 16, (m + 15) / 16), block(16, 16);
  small_gram_kernel<<<grid, block, 0, st>>>(w, d, k, m, len, row_gram, scratch);
  LASSO_CHECK_LAUNCH();
  count_launch();
  if (m <= 64) power_iter_small_kernel<<<1, 64, 0, st>>>(scratch, m, iters, l_dev);
  else power_iter_kernel<<<1, 1024, 0, st>>>(scratch, m, iters, scratch + (size_t)m * m, l_dev);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

int gram_run(const float* z, const float* x, int64_t n, int d, int k, double* gzz, double* gzx,
             cudaStream_t st) {
  LASSO_CUDA_TRY(cudaMemsetAsync(gzz, 0, sizeof(double) * (size_t)k * k, st));
  LASSO_CUDA_TRY(cudaMemsetAsync(gzx, 0, sizeof(double) * (size_t)k * d, st));
  if (n == 0) return LASSO_B200_OK;
  const int nbi = (k + kGT - 1) / kGT, nbj = (k + d + kGT - 1) / kGT;
  dim3 grid((unsigned)((n + kGSlab - 1) / kGSlab), (unsigned)(nbi * nbj));
  gram_kernel<<<grid, 256, 0, st>>>(z, x, n, d, k, gzz, gzx);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

int dict_update_run(float* dict, double* gzz, double* gzx, int d, int k, double eps,
                    const float* redraw, int* zeroed, cudaStream_t st) {
  const size_t smem = sizeof(double) * ((size_t)kSweepDepth * (k + d) + d) +
                      sizeof(float) * ((size_t)d * k + k) + (size_t)k;
  if (smem <= 200 * 1024 && k <= 32 * kSweepMaxPerLane && (k % 2) == 0 && (d % 2) == 0 && k / 2 + d / 2 <= 1024) {
    static bool attr_set = false;
    if (!attr_set) {
      LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)dict_sweep_smem_kernel,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    int threads = 1024;
    if (const char* t = getenv("LASSO_B200_SWEEP_THREADS")) threads = atoi(t);
    if (threads < 256 || threads > 1024 || (threads % 32) != 0 || k / 2 + d / 2 > threads) threads = 1024;
    dict_sweep_smem_kernel<<<1, threads, smem, st>>>(dict, gzz, gzx, d, k, eps, redraw, zeroed);
    LASSO_CHECK_LAUNCH();
    count_launch();
    return LASSO_B200_OK;
  }
  double* u = nullptr;
  LASSO_CUDA_TRY(cudaMallocAsync(&u, sizeof(double) * (size_t)d, st));
  dict_sweep_kernel<<<1, 1024, 0, st>>>(dict, gzz, gzx, d, k, eps, redraw, zeroed, u);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(u, st);
  if (e != cudaSuccess) {
    set_error("dict_sweep_kernel launch failed: %s", cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  count_launch();
  return LASSO_B200_OK;
}

}  // namespace lasso
