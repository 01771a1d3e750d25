// Small kernels around the FISTA loop: Lipschitz constant (power iteration),
// M-step sufficient statistics (Gram matrices) and the Gram-space atom sweep.
#include <cstdlib>

#include <cooperative_groups.h>

#include <algorithm>

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

// The leading eigenvector is found in float32 -- this GPU issues only a few float64 operations per clock, and
// the Rayleigh quotient is second order in the vector error, so a float32-converged vector (error ~1e-6) gives
// the eigenvalue to ~1e-11 -- and not on G but on G^4096 (twelve trace-normalised squarings, gram_square_kernel):
// same eigenvectors, the eigenvalue ratio that sets the convergence rate raised to the 4096th power.  A learned
// dictionary's two largest eigenvalues are often within 0.05 % of each other: on G^16 (what this used until the
// notebook workload was profiled) that ran all 2000 iterations, 3.4 ms per EM step at m = 289, each iteration
// streaming the m x m matrix through a single SM; on G^4096 the same pair converges in ~10, and the iteration
// count is capped at 256: with relative gap g the quotient is off by g exp(-2 * 4096 * 256 g) <= 1.8e-7 at worst.  Rounding in the
// squarings perturbs the vector by ~eps / gap, the Rayleigh quotient by at most min(gap, eps^2 / gap) <= eps.
// ONE float64 matrix-vector product with the original Gram and a Rayleigh quotient finish.
constexpr int kSqTile = 32;       // output tile of a block of 16 x 16 threads (2 x 2 outputs each)
constexpr int kSquarings = 12;
constexpr int kPowerIterCap = 256;   // on G^4096: the Rayleigh quotient is then within 1 / (2 * 4096 * 256 * e) = 1.8e-7 for ANY gap
constexpr int kTraceSlots = 32;
// C = (A / tr_in) (A / tr_in) for a symmetric m x m matrix; tr_out += trace(C).  a_dbl != nullptr: first
// squaring, the input is the float64 Gram itself.  C = A A^T for symmetric A, so both operand tiles are rows of A
// (coalesced), only blocks on or above the diagonal are computed and the others are mirrored on the store.
__global__ void __launch_bounds__(256) gram_square_kernel(const float* __restrict__ a_f, const double* __restrict__ a_dbl,
                                                          int m, const float* __restrict__ tr_in, float* __restrict__ c,
                                                          float* __restrict__ tr_out) {
  if (blockIdx.x < blockIdx.y) return;
  __shared__ float ta[kSqTile][17], tb[kSqTile][17];
  const float scale = 1.0f / fmaxf(tr_in[0], 1e-30f);
  const int i0 = blockIdx.y * kSqTile, j0 = blockIdx.x * kSqTile;
  const int tx = threadIdx.x, ty = threadIdx.y;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  for (int k0 = 0; k0 < m; k0 += 16) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int r = ty + 16 * e, kk = k0 + tx;
      float va = 0.f, vb = 0.f;
      if (kk < m) {
        if (i0 + r < m) va = a_dbl ? (float)a_dbl[(int64_t)(i0 + r) * m + kk] : a_f[(int64_t)(i0 + r) * m + kk];
        if (j0 + r < m) vb = a_dbl ? (float)a_dbl[(int64_t)(j0 + r) * m + kk] : a_f[(int64_t)(j0 + r) * m + kk];
      }
      ta[r][tx] = va * scale;
      tb[r][tx] = vb * scale;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float a0 = ta[ty][kk], a1 = ta[ty + 16][kk], b0 = tb[tx][kk], b1 = tb[tx + 16][kk];
      acc[0][0] = fmaf(a0, b0, acc[0][0]);
      acc[0][1] = fmaf(a0, b1, acc[0][1]);
      acc[1][0] = fmaf(a1, b0, acc[1][0]);
      acc[1][1] = fmaf(a1, b1, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < 2; ++e)
#pragma unroll
    for (int f = 0; f < 2; ++f) {
      const int i = i0 + ty + 16 * e, j = j0 + tx + 16 * f;
      if (i < m && j < m) {
        c[(int64_t)i * m + j] = acc[e][f];
        if (blockIdx.x != blockIdx.y) c[(int64_t)j * m + i] = acc[e][f];
        if (i == j) atomicAdd(tr_out, acc[e][f]);
      }
    }
}
// tr[0] = trace of the float64 Gram (as float), the other slots = 0
__global__ void gram_trace_kernel(const double* __restrict__ gram, int m, float* __restrict__ tr) {
  double s = 0.0;
  for (int a = threadIdx.x; a < m; a += blockDim.x) s += gram[(int64_t)a * m + a];
  s = warp_sum(s);
  __shared__ double red[32];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) t += red[wi];
    tr[0] = (float)t;
    for (int i = 1; i < kTraceSlots; ++i) tr[i] = 0.f;
  }
}
__global__ void __launch_bounds__(1024) power_iter_kernel(const double* __restrict__ gram, int m,
                                                          int iters, const float* __restrict__ gram_f,
                                                          double* __restrict__ out) {
  extern __shared__ float pi_smem[];   // v [m], u [m]
  float* v = pi_smem;
  float* u = pi_smem + m;
  __shared__ float redf[32], redm[32];
  __shared__ double redd[32], rede[32];
  __shared__ float s_nn, s_dv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int a = tid; a < m; a += blockDim.x) {
    // deterministic, non-symmetric start vector
    unsigned h = (unsigned)a * 2654435761u;
    v[a] = 1.0f + 0.25f * (float)((h >> 8) & 0xffff) / 65536.0f;
  }
  __syncthreads();
  int calm = 0;
  for (int it = 0; it < iters; ++it) {
    // four rows x four 32-column steps per warp at a time: 16 independent loads per lane in flight, so one L2
    // latency covers 2 KB per warp (a dependent load per step made an iteration cost 18 us at m = 289)
    for (int a = warp; a < m; a += 4 * nwarps) {
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c0 = lane; c0 < m; c0 += 128) {
        float g[4][4], vc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int cc = c0 + 32 * j;
          vc[j] = cc < m ? v[cc] : 0.f;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int ar = a + r * nwarps;
            g[r][j] = (cc < m && ar < m) ? gram_f[(int64_t)ar * m + cc] : 0.f;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r] = fmaf(g[r][j], vc[j], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float sr = acc[r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sr += __shfl_xor_sync(0xffffffffu, sr, o);
        if (lane == 0 && a + r * nwarps < m) u[a + r * nwarps] = sr;
      }
    }
    __syncthreads();
    float part = 0.f;
    for (int a = tid; a < m; a += blockDim.x) part = fmaf(u[a], u[a], part);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) redf[warp] = part;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int wi = 0; wi < nwarps; ++wi) s += redf[wi];
      s_nn = s;
    }
    __syncthreads();
    const float nn = s_nn;
    if (!(nn > 0.f)) break;   // zero matrix (or NaN): the float64 pass below reports it
    const float inv = rsqrtf(nn);
    float dmax = 0.f;
    for (int a = tid; a < m; a += blockDim.x) {
      const float vn = u[a] * inv;
      dmax = fmaxf(dmax, fabsf(vn - v[a]));
      v[a] = vn;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
    if (lane == 0) redm[warp] = dmax;
    __syncthreads();
    if (tid == 0) {
      float s = 0.f;
      for (int wi = 0; wi < nwarps; ++wi) s = fmaxf(s, redm[wi]);
      s_dv = s;
    }
    __syncthreads();
    // components of a unit vector are O(m^-1/2): stop when none moves by more than float32 noise
    if (s_dv <= 4e-7f) {
      if (++calm >= 3) break;
    } else {
      calm = 0;
    }
  }
  // float64 Rayleigh quotient of the float32-converged vector:  lambda = v^T G v / v^T v
  double num = 0.0, den = 0.0;
  for (int a = warp; a < m; a += nwarps) {
    double s = 0.0;
    for (int c = lane; c < m; c += 32) s += gram[(int64_t)a * m + c] * (double)v[c];
    s = warp_sum(s);
    if (lane == 0) {
      num += s * (double)v[a];
      den += (double)v[a] * (double)v[a];
    }
  }
  if (lane == 0) {
    redd[warp] = num;
    rede[warp] = den;
  }
  __syncthreads();
  if (tid == 0) {
    double sn = 0.0, sd = 0.0;
    for (int wi = 0; wi < nwarps; ++wi) {
      sn += redd[wi];
      sd += rede[wi];
    }
    out[0] = sd > 0.0 ? sn / sd : 0.0;
  }
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
                                                          double* __restrict__ u, int positive) {
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
      if (lane == 0) {
        const double v = gzx[(int64_t)j * d + i] - s + ajj * (double)dict[(int64_t)i * k + j];
        u[i] = positive ? fmax(v, 0.0) : v;     // positive: clamp before the norm (dict_learning.py:87-88)
      }
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
          double r = (double)redraw[(int64_t)i * k + j];
          if (positive) r = fmax(r, 0.0);           // dict_learning.py:94-95
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
                                                               int* __restrict__ zeroed, int positive) {
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
        const double v0 = bj[i0] - (double)s0, v1 = two ? bj[i1] - (double)s1 : 0.0;
        us[i0] = positive ? fmax(v0, 0.0) : v0;   // positive: clamp before the norm (dict_learning.py:87-88)
        if (two) us[i1] = positive ? fmax(v1, 0.0) : v1;
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
        for (int i = tid; i < d; i += blockDim.x) {
          const double r = (double)redraw[(int64_t)i * k + j];
          us[i] = positive ? fmax(r, 0.0) : r;      // dict_learning.py:94-95
        }
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

// Cluster variant for dictionaries that do not fit one SM's shared memory (e.g. the reference
// notebook's d = 289, k = 300).  The rows i of u_i = B[j,i] - sum_{l != j} D[i,l] A[j,l] are
// independent, so a cluster of kSweepCluster CTAs splits them: each CTA keeps its slice of the
// dictionary rows in shared memory for the whole sweep and only the squared norm of u crosses CTAs
// (every CTA stores its partial into every peer's shared memory, then one cluster barrier).  The
// partials are summed in rank order by every CTA, so all of them see the same |u| bit for bit.
// Same arithmetic as dict_sweep_smem_kernel (float32 inner products without the diagonal term,
// float64 u / norm); rows of A, B arrive by 8-byte cp.async, so d and k may be odd.
constexpr int kSweepCluster = 8;
constexpr int kSweepClusterThreads = 512;
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc)
               : "memory");
}
__global__ void __cluster_dims__(kSweepCluster, 1, 1) __launch_bounds__(kSweepClusterThreads)
    dict_sweep_cluster_kernel(float* dict, double* gzz, double* gzx, int d, int k, int dl, double eps,
                              int* __restrict__ zeroed, int positive) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int i_lo = min(d, rank * dl), nloc = min(d, i_lo + dl) - i_lo;    // this CTA's rows of D / u
  extern __shared__ __align__(16) unsigned char sweep_smem[];
  double* arow = reinterpret_cast<double*>(sweep_smem);                   // [depth][k]
  double* brow = arow + kSweepDepth * k;                                  // [depth][dl]
  double* us = brow + kSweepDepth * dl;                                   // [dl]
  double* parts = us + dl;                                                // [2][cluster] squared-norm partials
  float* ds = reinterpret_cast<float*>(parts + 2 * kSweepCluster);        // [dl][k]
  float* af = ds + (size_t)dl * k;                                        // [k]
  unsigned char* dead = reinterpret_cast<unsigned char*>(af + k);         // [k]
  __shared__ double s_inv, s_nrm;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  auto prefetch = [&](int j) {
    if (j < k) {
      const int slot = j % kSweepDepth;
      for (int t = tid; t < k + nloc; t += blockDim.x) {
        if (t < k) cp_async8(arow + slot * k + t, gzz + (int64_t)j * k + t);
        else cp_async8(brow + slot * dl + (t - k), gzx + (int64_t)j * d + i_lo + (t - k));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int e = tid; e < nloc * k; e += blockDim.x) ds[e] = dict[(int64_t)i_lo * k + e];
  for (int l = tid; l < k; l += blockDim.x) dead[l] = 0;
  for (int j = 0; j < kSweepDepth - 1; ++j) prefetch(j);
  for (int j = 0; j < k; ++j) {
    prefetch(j + kSweepDepth - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kSweepDepth - 1) : "memory");
    __syncthreads();
    const double* aj = arow + (j % kSweepDepth) * k;
    const double* bj = brow + (j % kSweepDepth) * dl;
    for (int l = tid; l < k; l += blockDim.x) af[l] = (l == j || dead[l]) ? 0.f : (float)aj[l];
    __syncthreads();
    for (int i0 = warp; i0 < nloc; i0 += 2 * nwarps) {                    // two rows per warp in flight
      const int i1 = i0 + nwarps;
      const bool two = i1 < nloc;
      const float* r0 = ds + (size_t)i0 * k;
      const float* r1 = ds + (size_t)(two ? i1 : i0) * k;
      float s0 = 0.f, s1 = 0.f;
      for (int l = lane; l < k; l += 32) {
        const float a = af[l];
        s0 = fmaf(r0[l], a, s0);
        s1 = fmaf(r1[l], a, s1);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      }
      if (lane == 0) {
        const double v0 = bj[i0] - (double)s0, v1 = two ? bj[i1] - (double)s1 : 0.0;
        us[i0] = positive ? fmax(v0, 0.0) : v0;   // positive: clamp before the norm (dict_learning.py:87-88)
        if (two) us[i1] = positive ? fmax(v1, 0.0) : v1;
      }
    }
    __syncthreads();
    if (warp == 0) {                                                      // this CTA's share of |u|^2 to every peer
      double part = 0.0;
      for (int i = lane; i < nloc; i += 32) part += us[i] * us[i];
      part = warp_sum(part);
      if (lane < kSweepCluster) cluster.map_shared_rank(parts, lane)[(j & 1) * kSweepCluster + rank] = part;
    }
    cluster.sync();
    if (warp == 0) {
      double ss = 0.0;
#pragma unroll
      for (int c = 0; c < kSweepCluster; ++c) ss += parts[(j & 1) * kSweepCluster + c];
      double inv = 0.0;
      if (ss > 0.0 && ss < 1e300) {
        if (ss < 1e-30 || ss > 1e30) {
          inv = 1.0 / sqrt(ss);
        } else {
          inv = (double)rsqrtf((float)ss);
#pragma unroll
          for (int t = 0; t < 3; ++t) inv = inv * (1.5 - 0.5 * ss * inv * inv);
        }
      }
      if (lane == 0) {
        s_inv = inv;
        s_nrm = ss * inv;
      }
    }
    __syncthreads();
    const double inv = s_inv, nrm = s_nrm;
    if (nrm < eps) {
      // degenerate atom (dict_learning.py:92-98): statistics row/column dropped, atom left for the
      // caller to redraw; the copies already prefetched are masked through dead[]
      if (rank == 0) {
        for (int l = tid; l < k; l += blockDim.x) {
          gzz[(int64_t)j * k + l] = 0.0;
          gzz[(int64_t)l * k + j] = 0.0;
        }
        for (int i = tid; i < d; i += blockDim.x) gzx[(int64_t)j * d + i] = 0.0;
        if (tid == 0) zeroed[j] = 1;
      }
      if (tid == 0) dead[j] = 1;
    } else {
      if (rank == 0 && tid == 0) zeroed[j] = 0;
      for (int i = tid; i < nloc; i += blockDim.x) {
        const float v = (float)(us[i] * inv);
        ds[(size_t)i * k + j] = v;
        dict[(int64_t)(i_lo + i) * k + j] = v;
      }
    }
  }
  cluster.sync();   // no CTA leaves while a peer may still store into its shared memory
}

// z[:, j] = 0 for every column whose mask entry is non-zero (dict_learning.py:98: the codes of a
// re-drawn atom are cleared); masked so that the host never has to read the mask back
__global__ void zero_columns_kernel(float* __restrict__ z, int64_t n, int k, const int* __restrict__ mask) {
  const int64_t total = n * (int64_t)k, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
    if (mask[(int)(i % k)]) z[i] = 0.f;
}

}  // namespace

void small_gram_launch(const float* w, int d, int k, int m, int len, int row_gram, double* gram, cudaStream_t st) {
  dim3 grid((m + 15) / 16, (m + 15) / 16), block(16, 16);
  small_gram_kernel<<<grid, block, 0, st>>>(w, d, k, m, len, row_gram, gram);
  count_launch();
}

int zero_columns_run(float* z, int64_t n, int k, const int* mask, cudaStream_t st) {
  if (n == 0) return LASSO_B200_OK;
  const int64_t total = n * (int64_t)k;
  const int blocks = (int)std::min<int64_t>((total + 1023) / 1024, 148 * 8);
  zero_columns_kernel<<<blocks, 256, 0, st>>>(z, n, k, mask);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

// lambda_max of a symmetric PSD m x m matrix already in scratch[0 .. m*m) (float64); scratch layout as in
// lipschitz_run.  Shared by the dictionary's Lipschitz constant and the convolutional one (conv_lip.cu).
int lambda_max_run(double* scratch, int m, int iters, double* l_dev, cudaStream_t st) {
  {
    // (one path for every size: the float64 iteration that small dictionaries used to take was slower --
    // 0.155 vs 0.111 ms at d = 64 -- because this GPU retires only ~3 float64 FMAs per clock and SM)
    // float region behind the doubles: two m x m ping-pong matrices + the traces
    float* f0 = reinterpret_cast<float*>(scratch + (size_t)m * m + 2 * (size_t)m + 8);
    float* f1 = f0 + (size_t)m * m;
    float* tr = f1 + (size_t)m * m;
    gram_trace_kernel<<<1, 256, 0, st>>>(scratch, m, tr);
    dim3 sgrid((m + kSqTile - 1) / kSqTile, (m + kSqTile - 1) / kSqTile), sblock(16, 16);
    const float* src = nullptr;
    float* dst = f0;
    for (int sq = 0; sq < kSquarings; ++sq) {                     // G^2, G^4, ... G^4096
      gram_square_kernel<<<sgrid, sblock, 0, st>>>(src, sq == 0 ? scratch : nullptr, m, tr + sq, dst, tr + sq + 1);
      src = dst;
      dst = dst == f0 ? f1 : f0;
    }
    LASSO_CHECK_LAUNCH();
    count_launch(1 + kSquarings);
    power_iter_kernel<<<1, 1024, 2 * (size_t)m * sizeof(float), st>>>(scratch, m, std::min(iters, kPowerIterCap), src, l_dev);
  }
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

// doubles needed by lambda_max_run for an m x m matrix (incl. the result slot behind the Gram)
size_t lambda_max_scratch_doubles(int m) {
  return (size_t)m * m + 2 * (size_t)m + 8 + ((2 * (size_t)m * m + kTraceSlots) * sizeof(float) + 7) / 8;
}

int lipschitz_run(const float* w, int d, int k, int iters, double* l_dev, double* scratch,
                  cudaStream_t st) {
  // scratch (doubles): [m*m] Gram | [2*m] spare | [8] result | then floats: 2 x [m*m] powers of the Gram, the traces
  const int row_gram = d <= k ? 1 : 0;
  const int m = row_gram ? d : k;
  const int len = row_gram ? k : d;
  dim3 grid((m + 15) / 16, (m + 15) / 16), block(16, 16);
  small_gram_kernel<<<grid, block, 0, st>>>(w, d, k, m, len, row_gram, scratch);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return lambda_max_run(scratch, m, iters, l_dev, st);
}

int gram_run(const float* z, const float* x, int64_t n, int d, int k, double* gzz, double* gzx,
             cudaStream_t st) {
  LASSO_CUDA_TRY(cudaMemsetAsync(gzz, 0, sizeof(double) * (size_t)k * k, st));
  LASSO_CUDA_TRY(cudaMemsetAsync(gzx, 0, sizeof(double) * (size_t)k * d, st));
  if (n == 0) return LASSO_B200_OK;
  // k <= 256, d <= 128 (BASELINE configs 2, 4): fp16-split tcgen05 kernel (gram_tc.cu); everything else and
  // small batches: the FFMA kernel below.  LASSO_B200_GRAM=ffma forces the latter (tests compare the two).
  static void* tc_scratch[64] = {nullptr};
  const char* force = getenv("LASSO_B200_GRAM");
  if (gram_tc_supported(n, d, k) && !(force && force[0] == 'f')) {
    int dev = 0;
    LASSO_CUDA_TRY(cudaGetDevice(&dev));
    if (!tc_scratch[dev]) LASSO_CUDA_TRY(cudaMalloc(&tc_scratch[dev], 256));
    return gram_tc_run(z, x, n, d, k, gzz, gzx, tc_scratch[dev], st);
  }
  const int nbi = (k + kGT - 1) / kGT, nbj = (k + d + kGT - 1) / kGT;
  dim3 grid((unsigned)((n + kGSlab - 1) / kGSlab), (unsigned)(nbi * nbj));
  gram_kernel<<<grid, 256, 0, st>>>(z, x, n, d, k, gzz, gzx);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

int dict_update_run(float* dict, double* gzz, double* gzx, int d, int k, double eps,
                    const float* redraw, int* zeroed, int positive, cudaStream_t st) {
  // default: the blocked sweep (sweep_blk.cu); LASSO_B200_SWEEP=legacy keeps the atom-by-atom kernels below
  // (the tests compare the two)
  {
    const char* mode = getenv("LASSO_B200_SWEEP");
    if (dict_update_blocked_supported(d, k) && !(mode && mode[0] == 'l'))
      return dict_update_blocked_run(dict, gzz, gzx, d, k, eps, redraw, zeroed, positive, st);
  }
  const size_t smem = sizeof(double) * ((size_t)kSweepDepth * (k + d) + d) +
                      sizeof(float) * ((size_t)d * k + k) + (size_t)k;
  if (smem <= 200 * 1024 && k <= 32 * kSweepMaxPerLane && (k % 2) == 0 && (d % 2) == 0 && k / 2 + d / 2 <= 1024) {
    // per device (context) attribute: set on every launch, one process may drive several GPUs
    LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)dict_sweep_smem_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int threads = 1024;
    if (const char* t = getenv("LASSO_B200_SWEEP_THREADS")) threads = atoi(t);
    if (threads < 256 || threads > 1024 || (threads % 32) != 0 || k / 2 + d / 2 > threads) threads = 1024;
    dict_sweep_smem_kernel<<<1, threads, smem, st>>>(dict, gzz, gzx, d, k, eps, redraw, zeroed, positive);
    LASSO_CHECK_LAUNCH();
    count_launch();
    return LASSO_B200_OK;
  }
  // larger dictionaries: rows split over a cluster of CTAs
  const int dl = (d + kSweepCluster - 1) / kSweepCluster;
  const size_t csmem = sizeof(double) * ((size_t)kSweepDepth * (k + dl) + dl + 2 * kSweepCluster) +
                       sizeof(float) * ((size_t)dl * k + k) + (size_t)k;
  if (redraw == nullptr && csmem <= 200 * 1024 && getenv("LASSO_B200_SWEEP_NO_CLUSTER") == nullptr) {
    LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)dict_sweep_cluster_kernel,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dict_sweep_cluster_kernel<<<kSweepCluster, kSweepClusterThreads, csmem, st>>>(dict, gzz, gzx, d, k, dl, eps,
                                                                                  zeroed, positive);
    LASSO_CHECK_LAUNCH();
    count_launch();
    return LASSO_B200_OK;
  }
  double* u = nullptr;
  LASSO_CUDA_TRY(cudaMallocAsync(&u, sizeof(double) * (size_t)d, st));
  dict_sweep_kernel<<<1, 1024, 0, st>>>(dict, gzz, gzx, d, k, eps, redraw, zeroed, u, positive);
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
