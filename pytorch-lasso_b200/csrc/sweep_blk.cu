// K4 (blocked): the Gauss-Seidel atom sweep of update_dict (dict_learning.py:82-101) in Gram space,
//     u_j = B_j - sum_{l != j} d_l A_jl   (d_l already updated for l < j),   d_j <- u_j / |u_j|,
// restructured so that the serial chain per atom is a short correction and a norm.
//
// The sweep kernels of aux_kernels.cu walk the atoms one by one and pay a d x k matrix-vector product plus
// 3-4 CTA (or cluster) barriers per atom: 2.0 us per atom at d = 64, k = 256 and 2.7 us at d = 289, k = 300 --
// 0.5 / 0.8 ms per EM step, the largest item of the M-step.  Here the atoms are taken in blocks of 64:
//   phase A (all SMs that the rows fill, one launch per block): the part of u_j that does not depend on the
//           block's own updates, U0_j = B_j - sum_{l != j} d_l^{start} A_jl for the 64 atoms at once -- a
//           [d x k] x [k x 64] product (float32 FMAs on float32 copies of A's rows, as the other sweep kernels
//           do; the subtraction from B_j in float64)
//   phase B (one CTA, thread = row of the dictionary): atom by atom
//           u_j = U0_j - sum_{l in block, l < j} (d_l^{new} - d_l^{start}) A_jl
//           with the differences of the block's earlier atoms in a shared-memory column the thread owns, then the
//           norm (float32 partial sums, one CTA barrier), the normalisation and the store.  A degenerate atom (|u| < eps,
//           dict_learning.py:91-98) drops out of the statistics: its difference is -d_l^{start}, later blocks see
//           its column of A as zero through the `dead` flags, and its row / column of the statistics are zeroed
//           like the other kernels do; the replacement comes from `redraw` when the host supplies one.
// Mathematically the same sweep (the terms are only grouped differently); any d <= 1024, any k.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace lasso {
namespace {

constexpr int kSbBlock = 64;     // atoms per block

// U0[jj][i] for the atoms j0 .. j0 + nb of one block; CTA = 32 rows x 64 atoms, 32 x 32 threads, 2 atoms each.
// The contraction runs in chunks of 96 columns so that a chunk's 9 loads per thread are in flight together (with
// 32-column chunks a launch was a chain of k / 32 global round trips: 27 us at k = 256).
constexpr int kSbChunk = 96;
__global__ void __launch_bounds__(1024) sweep_block_gemm_kernel(const float* __restrict__ dict, const double* __restrict__ gzz,
                                                                const double* __restrict__ gzx, int d, int k, int j0,
                                                                int nb, const int* __restrict__ dead,
                                                                float* __restrict__ u0) {
  // row pitch kSbChunk + 4: 16-byte aligned rows, and the 128-bit loads of 8 consecutive rows hit all banks
  __shared__ __align__(16) float td[32][kSbChunk + 4], ta[kSbBlock][kSbChunk + 4];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.x * 32;
  float acc0 = 0.f, acc1 = 0.f;
  for (int l0 = 0; l0 < k; l0 += kSbChunk) {
    float dv[kSbChunk / 32], av[2][kSbChunk / 32];
#pragma unroll
    for (int c = 0; c < kSbChunk / 32; ++c) {
      const int l = l0 + 32 * c + tx;
      const bool alive = l < k && dead[l] == 0;
      dv[c] = (l < k && i0 + ty < d) ? dict[(int64_t)(i0 + ty) * k + l] : 0.f;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = ty + 32 * e, j = j0 + jj;
        av[e][c] = (jj < nb && alive && l != j) ? (float)gzz[(int64_t)j * k + l] : 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < kSbChunk / 32; ++c) {
      td[ty][32 * c + tx] = dv[c];
      ta[ty][32 * c + tx] = av[0][c];
      ta[ty + 32][32 * c + tx] = av[1][c];
    }
    __syncthreads();
    // thread (ty, tx) -> row i0 + tx, atoms ty and ty + 32; 128-bit shared-memory loads: one conflict-free load of the
    // row's four values and two broadcast loads per 8 FMAs (scalar loads: 3 per 2 FMAs, and the kernel was bound by them)
#pragma unroll 8
    for (int l4 = 0; l4 < kSbChunk / 4; ++l4) {
      const float4 v = *reinterpret_cast<const float4*>(&td[tx][4 * l4]);
      const float4 a0 = *reinterpret_cast<const float4*>(&ta[ty][4 * l4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&ta[ty + 32][4 * l4]);
      acc0 = fmaf(v.x, a0.x, acc0); acc1 = fmaf(v.x, a1.x, acc1);
      acc0 = fmaf(v.y, a0.y, acc0); acc1 = fmaf(v.y, a1.y, acc1);
      acc0 = fmaf(v.z, a0.z, acc0); acc1 = fmaf(v.z, a1.z, acc1);
      acc0 = fmaf(v.w, a0.w, acc0); acc1 = fmaf(v.w, a1.w, acc1);
    }
    __syncthreads();
  }
  const int i = i0 + tx;
  if (i < d) {
    if (ty < nb) u0[(int64_t)ty * d + i] = (float)(gzx[(int64_t)(j0 + ty) * d + i] - (double)acc0);
    if (ty + 32 < nb) u0[(int64_t)(ty + 32) * d + i] = (float)(gzx[(int64_t)(j0 + ty + 32) * d + i] - (double)acc1);
  }
}

// the serial part of one block; one CTA, thread i = row i of the dictionary (blockDim.x >= d, multiple of 32).
// kB atoms per block: 64 for d <= 512, 32 beyond (the differences of the block live in shared memory, kB x d floats).
// Measured per atom at d = 64 (clock64, one thread): ~1700 cycles for ~350 instructions -- with one or two warps
// per scheduler nothing hides the issue-to-issue latency of a dependent instruction stream, so the step costs what
// its instruction count costs (corrections 840, shuffles 180, barrier 110, sums + root 270, normalise + stores 250).
// A version with the differences in registers (predicated sweep to write one of 64) measured the same; unrolling
// the loop over the atoms made it wait for instruction fetches instead.
template <int kB, int kMaxThreads>
__global__ void __launch_bounds__(kMaxThreads) sweep_block_seq_kernel(float* __restrict__ dict, double* __restrict__ gzz,
                                                               double* __restrict__ gzx, int d, int k, int j0, int nb,
                                                               double eps, const float* __restrict__ redraw,
                                                               int* __restrict__ zeroed, int* __restrict__ dead,
                                                               const float* __restrict__ u0, int positive) {
  __shared__ __align__(16) float ajj[kB][kB + 4];     // A[j][l] for j, l in the block (0 on the diagonal)
  extern __shared__ float delta_s[];                  // [kB][dp]: d_l^{new} - d_l^{start}, column i owned by thread i
  __shared__ double red[32];
  __shared__ float redf[2][32];
  __shared__ double s_inv, s_nrm;
  const int i = threadIdx.x, lane = i & 31, warp = i >> 5, nwarps = blockDim.x >> 5;
  const bool row = i < d;
  const int dp = d | 1;
  for (int e = i; e < kB * kB; e += blockDim.x) {
    const int jj = e / kB, ll = e % kB;
    ajj[jj][ll] = (jj < nb && ll < nb && jj != ll) ? (float)gzz[(int64_t)(j0 + jj) * k + j0 + ll] : 0.f;
  }
  __syncthreads();
  // 1 / sqrt(sum over the CTA of v) in float64 for sums outside the comfortable float range
  auto inv_norm = [&](double part, double& inv, double& nrm) {
    part = warp_sum(part);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (warp == 0) {
      double ss = lane < nwarps ? red[lane] : 0.0;
      ss = warp_sum(ss);
      double r = 0.0;
      if (ss > 0.0 && ss < 1e300) r = 1.0 / sqrt(ss);
      if (lane == 0) {
        s_inv = r;
        s_nrm = ss * r;        // |u| (0 for an all-zero u)
      }
    }
    __syncthreads();
    inv = s_inv;
    nrm = s_nrm;
  };
  // U0 and the old value of the atom four steps ahead are fetched while this atom's norm is formed (every store to
  // the dictionary drops the line from L1, so each of these loads is an L2 round trip).  The loop is NOT unrolled
  // over the atoms: 64 copies of the body, each executed once, wait for their instruction fetches.
  auto fetch_u = [&](int a) { return (row && a < nb) ? u0[(int64_t)a * d + i] : 0.f; };
  auto fetch_d = [&](int a) { return (row && a < nb) ? dict[(int64_t)i * k + j0 + a] : 0.f; };
  float u1 = fetch_u(0), u2 = fetch_u(1), u3 = fetch_u(2), u4 = fetch_u(3);
  float d1 = fetch_d(0), d2 = fetch_d(1), d3 = fetch_d(2), d4 = fetch_d(3);
  const float eps_f = (float)eps;
  const int col = row ? i : 0;
#pragma unroll 1
  for (int jj = 0; jj < nb; ++jj) {
    const int j = j0 + jj;
    const float d_old = d1;
    float u = u1;
    u1 = u2; u2 = u3; u3 = u4; u4 = fetch_u(jj + 4);
    d1 = d2; d2 = d3; d3 = d4; d4 = fetch_d(jj + 4);
    {
      // corrections of the block's earlier atoms, eight at a time: the loads of a group are issued together
      float c0 = 0.f, c1 = 0.f;
      for (int l0 = 0; l0 < jj; l0 += 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(&ajj[jj][l0]);
        const float4 a1 = *reinterpret_cast<const float4*>(&ajj[jj][l0 + 4]);
        float dv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dv[e] = (row && l0 + e < jj) ? delta_s[(l0 + e) * dp + col] : 0.f;   // (idle threads read nothing)
        c0 = fmaf(dv[0], a0.x, c0); c1 = fmaf(dv[1], a0.y, c1);
        c0 = fmaf(dv[2], a0.z, c0); c1 = fmaf(dv[3], a0.w, c1);
        c0 = fmaf(dv[4], a1.x, c0); c1 = fmaf(dv[5], a1.y, c1);
        c0 = fmaf(dv[6], a1.z, c0); c1 = fmaf(dv[7], a1.w, c1);
      }
      u = row ? u - (c0 + c1) : 0.f;
    }
    if (positive) u = fmaxf(u, 0.f);            // clamp before the norm (dict_learning.py:87-88)
    // |u|: float32 partial sums (what the reference's own norm is made of), ONE barrier -- every thread adds the
    // per-warp sums itself, in the same order; sums outside the comfortable float range take the float64 path
    float inv_f, nrm_f;
    {
      float part = u * u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == 0) redf[jj & 1][warp] = part;
      __syncthreads();
      float ss = 0.f;
      for (int w = 0; w < nwarps; ++w) ss += redf[jj & 1][w];
      if (ss > 1e-30f && ss < 1e30f) {
        float r = rsqrtf(ss);
        r = r * (1.5f - 0.5f * ss * r * r);
        inv_f = r;
        nrm_f = ss * r;
      } else {
        double inv, nrm;
        inv_norm((double)u * (double)u, inv, nrm);
        inv_f = (float)fmin(inv, 3.0e38);
        nrm_f = (float)fmin(nrm, 3.0e38);
        if (nrm < eps) nrm_f = 0.f;
        else if (nrm_f < eps_f) nrm_f = eps_f;
      }
    }
    const bool degenerate = nrm_f < eps_f;
    float d_new = d_old, dl = 0.f;
    if (degenerate) {
      // the atom's codes are dropped (dict_learning.py:92-98): it vanishes from the statistics
      if (i == 0) {
        zeroed[j] = 1;
        dead[j] = 1;
      }
      for (int l = i; l < k; l += blockDim.x) {
        gzz[(int64_t)j * k + l] = 0.0;
        gzz[(int64_t)l * k + j] = 0.0;
      }
      for (int c = i; c < d; c += blockDim.x) gzx[(int64_t)j * d + c] = 0.0;
      // later atoms of THIS block had -d_j^{start} A_lj in their U0: the difference -d_j^{start} takes it out again
      dl = -d_old;
      if (redraw != nullptr) {
        float r = row ? redraw[(int64_t)i * k + j] : 0.f;
        if (positive) r = fmaxf(r, 0.f);        // dict_learning.py:94-95
        double rinv, rnrm;
        inv_norm((double)r * (double)r, rinv, rnrm);
        if (rnrm > 0.0) d_new = (float)((double)r * rinv);
      }
    } else {
      if (i == 0) zeroed[j] = 0;
      d_new = u * inv_f;
      dl = d_new - d_old;
    }
    if (row && d_new != d_old) dict[(int64_t)i * k + j] = d_new;
    if (row) delta_s[jj * dp + i] = dl;
  }
}

struct SweepBlkState {
  float* u0 = nullptr;
  size_t u0_cap = 0;
  int* dead = nullptr;
  int dead_cap = 0;
};
SweepBlkState g_sb[64];

}  // namespace

bool dict_update_blocked_supported(int d, int k) { return d >= 1 && d <= 1024 && k >= 1; }

int dict_update_blocked_run(float* dict, double* gzz, double* gzx, int d, int k, double eps, const float* redraw,
                            int* zeroed, int positive, cudaStream_t st) {
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  SweepBlkState& S = g_sb[dev];
  const size_t need = sizeof(float) * (size_t)kSbBlock * d;
  if (need > S.u0_cap) {
    if (S.u0) LASSO_CUDA_TRY(cudaFree(S.u0));
    S.u0 = nullptr;
    S.u0_cap = 0;
    LASSO_CUDA_TRY(cudaMalloc(&S.u0, need));
    S.u0_cap = need;
  }
  if (k > S.dead_cap) {
    if (S.dead) LASSO_CUDA_TRY(cudaFree(S.dead));
    S.dead = nullptr;
    S.dead_cap = 0;
    LASSO_CUDA_TRY(cudaMalloc(&S.dead, sizeof(int) * (size_t)k));
    S.dead_cap = k;
  }
  LASSO_CUDA_TRY(cudaMemsetAsync(S.dead, 0, sizeof(int) * (size_t)k, st));
  const int threads = std::max(128, (d + 31) / 32 * 32);      // (>= 128: the block's A tile is loaded by all of them)
  const dim3 ggrid((unsigned)((d + 31) / 32)), gblock(32, 32);
  const int bsz = d <= 512 ? kSbBlock : 32;
  const size_t seq_smem = sizeof(float) * (size_t)bsz * (d | 1);          // <= 131 KB: the block's differences
  if (d <= 512)
    LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)sweep_block_seq_kernel<kSbBlock, 512>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem));
  else
    LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)sweep_block_seq_kernel<32, 1024>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seq_smem));
  for (int j0 = 0; j0 < k; j0 += bsz) {
    const int nb = std::min(bsz, k - j0);
    sweep_block_gemm_kernel<<<ggrid, gblock, 0, st>>>(dict, gzz, gzx, d, k, j0, nb, S.dead, S.u0);
    if (d <= 512)
      sweep_block_seq_kernel<kSbBlock, 512><<<1, threads, seq_smem, st>>>(dict, gzz, gzx, d, k, j0, nb, eps, redraw, zeroed,
                                                                   S.dead, S.u0, positive);
    else
      sweep_block_seq_kernel<32, 1024><<<1, threads, seq_smem, st>>>(dict, gzz, gzx, d, k, j0, nb, eps, redraw, zeroed, S.dead,
                                                              S.u0, positive);
    LASSO_CHECK_LAUNCH();
    count_launch(2);
  }
  return LASSO_B200_OK;
}

}  // namespace lasso
