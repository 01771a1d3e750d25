// K3 on the tensor cores: the M-step statistics  A = Z^T Z (k x k),  B = Z^T X (k x d)  of
// update_dict / update_dict_ridge (dict_learning.py:82-101, 117-118) as fp16-split tcgen05 GEMMs whose
// contraction runs over the ROWS of the batch.
//
//   output tile   128 atoms (M) x 128 columns of [Z | X] (N); a CTA owns one tile for one slab of rows.
//                 k, d <= 512: the Z tiles on or above the diagonal and all X tiles (5 tile types at k = 256,
//                 d = 64; 15 at the notebook's k = 300, d = 289); an off-diagonal Z tile is mirrored into the
//                 lower triangle when it is flushed
//   per 64 rows   A = Z^T block: thread = (atom, k-step of 16 rows) reads its 16 values (coalesced across
//                 the atoms of a warp), splits them into fp16 pieces h + l and stores them to a TMEM stage
//                 (the layout of the resident kernel's piece slots);
//                 B = [Z | X] block: pieces to a 128-byte-swizzled shared-memory image, MN-major (columns
//                 contiguous), two 64-column slabs -- the layout of the resident kernel's dictionary image;
//                 MMAs: lead product h h' into accumulator L, cross products h l' + l h' into accumulator C
//   accumulate    the tensor core truncates (round toward zero) inside every accumulate: over a run of
//                 same-sign terms that is a bias that grows with the run (measured on dense codes: 1.1e-6
//                 of the diagonal after 16 accumulations, 5.8e-7 after 8, 3.6e-7 after 4).  So L / C only
//                 ever see the 4 accumulations of one 64-row block; then the compute warps fold them into a
//                 second TMEM accumulator with ordinary round-to-nearest adds while the next block's
//                 operands are already converted, and only that accumulator is flushed (float64 atomics)
//                 at the end of the slab.  Against the float64 Gram: 2e-7 (sparse codes) .. 4e-7 (dense).
//   range         one power-of-two scale for Z and one for X (max |z'|, max |x'| in [256, 512)) from a
//                 max-abs pre-pass; the statistics are unscaled exactly when they are flushed.
// TMEM columns: A stages 2 x 64 | L 128 | C 128 | level-2 accumulator 128 = 512.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace lasso {
namespace {

using namespace sm100;

constexpr int kGtRows = 64;                 // rows per block (4 k-steps)
constexpr int kGtThreads = 544;             // 16 compute warps + the MMA issuer (last warp)
constexpr uint32_t kGtSlab = kGtRows * 128;             // [64 rows][128 B] = 64 columns of one piece
constexpr uint32_t kGtPiece = 2 * kGtSlab;              // 128 columns
constexpr uint32_t kGtStage = 2 * kGtPiece;             // h, l: 32 KB
constexpr uint32_t kGtSmem = 2 * kGtStage;              // two stages
constexpr uint32_t kGtColA = 0;       // A stages: 2 x [h 32 cols | l 32 cols]
constexpr uint32_t kGtColL = 128;     // lead accumulator
constexpr uint32_t kGtColC = 256;     // cross accumulator
constexpr uint32_t kGtColS = 384;     // level-2 accumulator (round-to-nearest adds)

struct GramTcParams {
  const float* z;
  const float* x;
  int64_t n;
  int d, k;
  int64_t slab_rows;        // rows per CTA slab (multiple of 64)
  int ntypes;               // tile types per slab
  const float* scales;      // [0] = sz, [1] = sx (powers of two), [2] = 1/sz^2, [3] = 1/(sz sx)
  double* gzz;
  double* gzx;
};

#define GT_WAIT(bar, parity)                                                              \
  do {                                                                                    \
    const uint32_t _addr = smem_u32(bar), _par = (parity) & 1u;                           \
    uint32_t _ok, _n = 0;                                                                 \
    for (;;) {                                                                            \
      asm volatile(                                                                       \
          "{\n\t.reg .pred P;\n\t"                                                       \
          "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"                  \
          "selp.b32 %0, 1, 0, P;\n\t}\n"                                                   \
          : "=r"(_ok)                                                                     \
          : "r"(_addr), "r"(_par), "r"(20000u)                                            \
          : "memory");                                                                    \
      if (_ok) break;                                                                     \
      if (++_n > (1u << 17)) __trap();   /* a protocol bug must not hang the GPU */       \
    }                                                                                     \
  } while (0)

__device__ __forceinline__ void gsplit2(float2 v, uint32_t& wh, uint32_t& wl) {
  const float2 t = make_float2(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                               __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
  const float2 r = __ffma2_rn(make_float2(-1.f, -1.f), t, v);
  const __half2 h = __floats2half2_rn(t.x, t.y);
  const __half2 l = __floats2half2_rn(r.x, r.y);
  wh = *reinterpret_cast<const uint32_t*>(&h);
  wl = *reinterpret_cast<const uint32_t*>(&l);
}

// tile type -> (atom offset m0 of the M tile, column offset n0 of the N tile, N tile of Z or of X).
// With MT = ceil(k / 128) M tiles: the Z tiles on or above the diagonal first, (0,0) (0,1) .. (0,MT-1) (1,1) ..,
// then the MT x ceil(d / 128) X tiles.  k <= 256, d <= 128: (0,Z0) (0,Z1) (1,Z1) (0,X) (1,X).
__host__ __device__ inline int gram_ntypes(int d, int k) {
  const int mt = (k + 127) / 128, xt = (d + 127) / 128;
  return mt * (mt + 1) / 2 + mt * xt;
}
__device__ __forceinline__ void gram_tile(int type, int d, int k, int& m0, int& n0, int& is_x) {
  const int mt = (k + 127) / 128, xt = (d + 127) / 128;
  const int nz = mt * (mt + 1) / 2;
  if (type < nz) {
    int row = 0, left = type;
    while (left >= mt - row) {
      left -= mt - row;
      ++row;
    }
    m0 = row * 128; n0 = (row + left) * 128; is_x = 0;
  } else {
    const int e = type - nz;
    m0 = (e / xt) * 128; n0 = (e % xt) * 128; is_x = 1;
  }
}

__global__ void __launch_bounds__(kGtThreads, 1) gram_tc_kernel(GramTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_accfull, bar_accfree;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int type = blockIdx.x % p.ntypes;
  const int64_t slab = blockIdx.x / p.ntypes;
  const int64_t r_begin = slab * p.slab_rows, r_end = min(p.n, r_begin + p.slab_rows);
  const int nblocks = r_begin < r_end ? (int)((r_end - r_begin + kGtRows - 1) / kGtRows) : 0;
  int m0, n0, is_x;
  gram_tile(type, p.d, p.k, m0, n0, is_x);

  if (tid == 0) {
    mbar_init(&bar_full[0], 512);
    mbar_init(&bar_full[1], 512);
    mbar_init(&bar_empty[0], 1);
    mbar_init(&bar_empty[1], 1);
    mbar_init(&bar_accfull, 1);
    mbar_init(&bar_accfree, 512);
    fence_mbar_init();
  }
  if (warp == 16) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 16) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc(kFmtF16, 128, 128, 0, 1);   // A from TMEM, B MN-major, N = 128
    for (int b = 0; b < nblocks; ++b) {
      const uint32_t s = (uint32_t)b & 1u;
      GT_WAIT(&bar_full[s], ((uint32_t)b >> 1) & 1u);
      // every block is its own level-1 run: L / C are overwritten, so the previous run must have been folded
      if (b > 0) GT_WAIT(&bar_accfree, ((uint32_t)b - 1u) & 1u);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t desc = make_smem_desc_sw128(smem_u32(smem + s * kGtStage), kGtSlab, 1024);
        const uint32_t d_lo = (uint32_t)desc, d_hi = (uint32_t)(desc >> 32);
        const uint32_t t_a = tbase + kGtColA + s * 64;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t bh = ((uint64_t)d_hi << 32) | (d_lo + ks * 128);
          const uint64_t bl = ((uint64_t)d_hi << 32) | (d_lo + (kGtPiece >> 4) + ks * 128);
          const uint32_t ah = t_a + ks * 8, al = ah + 32;
          const uint32_t acc_on = ks == 0 ? 0u : 1u;
          mma_ts<false>(tbase + kGtColC, ah, bl, idesc, acc_on);
          mma_ts<false>(tbase + kGtColC, al, bh, idesc, 1);
          mma_ts<false>(tbase + kGtColL, ah, bh, idesc, acc_on);
        }
        mma_commit(&bar_empty[s]);
        mma_commit(&bar_accfull);
      }
      __syncwarp();
    }
  } else {
    // ===================== compute warps: operand conversion, level-2 accumulation, flush =====================
    const int quad = warp & 3, wg = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float sz = p.scales[0], sx = p.scales[1];
    const float sb = is_x ? sx : sz;
    const int atom = m0 + quad * 32 + lane;            // this thread's row of the output tile
    const bool atom_ok = atom < p.k;
    const int ncols = is_x ? p.d : p.k;                // valid columns of the N side
    const float* bsrc = is_x ? p.x : p.z;
    const int bpitch = is_x ? p.d : p.k;
    // The operands of block b + 1 are fetched into registers right after block b has been converted: the loads
    // are in flight while this thread waits for the tensor pipe (stage hand-over, level-1 fold), otherwise every
    // block pays a full HBM latency with nothing else to do.
    float2 va[8], vb[8];
    auto fetch = [&](int b) {
      const int64_t r0 = r_begin + (int64_t)b * kGtRows, rr0 = r0 + wg * 16;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        va[j] = make_float2(0.f, 0.f);
        if (atom_ok) {
          if (rr0 + 2 * j < r_end) va[j].x = __ldg(p.z + (rr0 + 2 * j) * p.k + atom);
          if (rr0 + 2 * j + 1 < r_end) va[j].y = __ldg(p.z + (rr0 + 2 * j + 1) * p.k + atom);
        }
      }
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int idx = (tid & 511) + t * 512;       // 64 rows x 64 column pairs
        const int i = idx >> 6, j = (idx & 63) * 2;
        const int64_t r = r0 + i;
        vb[t] = make_float2(0.f, 0.f);
        if (r < r_end) {
          const int c = n0 + j;
          if (c < ncols) vb[t].x = __ldg(bsrc + r * bpitch + c);
          if (c + 1 < ncols) vb[t].y = __ldg(bsrc + r * bpitch + c + 1);
        }
      }
    };
    // operands in registers (va, vb) -> TMEM stage / shared-memory image of block b, then signal the MMA warp
    auto convert = [&](int b) {
      const uint32_t s = (uint32_t)b & 1u;
      if (b >= 2) GT_WAIT(&bar_empty[s], (((uint32_t)b >> 1) - 1u) & 1u);
      tc_fence_after();
      {   // A: 16 rows of this thread's atom (k-step wg) -> fp16 pieces -> TMEM stage
        uint32_t wh[8], wl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) gsplit2(make_float2(va[j].x * sz, va[j].y * sz), wh[j], wl[j]);
        const uint32_t t_a = tbase + lane_base + kGtColA + s * 64 + wg * 8;
        tmem_st8(t_a, wh);
        tmem_st8(t_a + 32, wl);
      }
      {   // B: [64 rows][128 columns] -> pieces -> swizzled shared-memory image (pairs of columns)
        uint8_t* img = smem + s * kGtStage;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const int idx = (tid & 511) + t * 512;
          const int i = idx >> 6, j = (idx & 63) * 2;
          uint32_t wh, wl;
          gsplit2(make_float2(vb[t].x * sb, vb[t].y * sb), wh, wl);
          const uint32_t off = (uint32_t)(j >> 6) * kGtSlab + sw128_offset((uint32_t)i, (uint32_t)(j & 63) * 2u);
          *reinterpret_cast<uint32_t*>(img + off) = wh;
          *reinterpret_cast<uint32_t*>(img + kGtPiece + off) = wl;
        }
      }
      if (b + 1 < nblocks) fetch(b + 1);
      tmem_wait_st();
      fence_proxy_async_smem();      // the MMA reads the image through the async proxy
      tc_fence_before();
      mbar_arrive(&bar_full[s]);
    };
    if (nblocks > 0) {
      fetch(0);
      convert(0);
    }
    for (int b = 0; b < nblocks; ++b) {
      // block b + 1 is converted while the tensor pipe works on block b ...
      if (b + 1 < nblocks) convert(b + 1);
      // ... then block b's level-1 run (4 accumulations) is folded into the level-2 accumulator, round to nearest
      GT_WAIT(&bar_accfull, (uint32_t)b & 1u);
      tc_fence_after();
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        const uint32_t col = wg * 32 + hcol * 16;
        uint32_t lv[16], cv[16], sv[16];
        tmem_ld16(tbase + lane_base + kGtColL + col, lv);
        tmem_ld16(tbase + lane_base + kGtColC + col, cv);
        if (b > 0) tmem_ld16(tbase + lane_base + kGtColS + col, sv);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float run = __uint_as_float(lv[j]) + __uint_as_float(cv[j]);
          sv[j] = __float_as_uint(b > 0 ? __uint_as_float(sv[j]) + run : run);
        }
        tmem_st16(tbase + lane_base + kGtColS + col, sv);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_accfree);
    }
    // ---- flush the tile: float64 atomics, unscaled ----
    // (tcgen05.ld is warp-collective: every lane takes part, only the atomics are per valid atom)
    if (nblocks > 0) {
      const float unscale = is_x ? p.scales[3] : p.scales[2];
#pragma unroll
      for (int hcol = 0; hcol < 2; ++hcol) {
        const uint32_t col = wg * 32 + hcol * 16;
        uint32_t sv[16];
        tmem_ld16(tbase + lane_base + kGtColS + col, sv);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = n0 + (int)col + j;
          const float v = __uint_as_float(sv[j]);
          if (!atom_ok || c >= ncols || v == 0.f) continue;
          const double dv = (double)v * (double)unscale;
          if (is_x) {
            atomicAdd(&p.gzx[(int64_t)atom * p.d + c], dv);
          } else if (m0 != n0) {                 // off-diagonal tile: the transposed tile comes for free
            atomicAdd(&p.gzz[(int64_t)atom * p.k + c], dv);
            atomicAdd(&p.gzz[(int64_t)c * p.k + atom], dv);
          } else if (c >= atom) {                // diagonal tile: upper triangle, mirrored: exactly symmetric
            atomicAdd(&p.gzz[(int64_t)atom * p.k + c], dv);
            if (c != atom) atomicAdd(&p.gzz[(int64_t)c * p.k + atom], dv);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tbase, 512);
}

// scales[0..3] from the maximum magnitudes of Z and X (powers of two: max |z'|, max |x'| in [256, 512))
__global__ void gram_maxabs_kernel(const float* __restrict__ z, int64_t nz, const float* __restrict__ x, int64_t nx,
                                   unsigned* __restrict__ maxbits) {
  unsigned mz = 0, mx = 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i = i0; i < nz; i += stride) mz = max(mz, __float_as_uint(z[i]) & 0x7FFFFFFFu);
  for (int64_t i = i0; i < nx; i += stride) mx = max(mx, __float_as_uint(x[i]) & 0x7FFFFFFFu);
  mz = __reduce_max_sync(0xffffffffu, mz);
  mx = __reduce_max_sync(0xffffffffu, mx);
  if ((threadIdx.x & 31) == 0) {
    if (mz) atomicMax(&maxbits[0], mz);
    if (mx) atomicMax(&maxbits[1], mx);
  }
}
__global__ void gram_scales_kernel(const unsigned* __restrict__ maxbits, float* __restrict__ scales) {
  float s[2];
  for (int i = 0; i < 2; ++i) {
    const float m = __uint_as_float(maxbits[i]);
    int e = 0;
    if (m > 0.f && m < 3.0e38f) e = 8 - ilogbf(m);        // max -> [256, 512)
    e = max(-60, min(60, e));
    s[i] = ldexpf(1.f, e);
  }
  scales[0] = s[0];
  scales[1] = s[1];
  scales[2] = 1.f / (s[0] * s[0]);
  scales[3] = 1.f / (s[0] * s[1]);
  scales[4] = (__uint_as_float(maxbits[0]) < 3.0e38f && __uint_as_float(maxbits[1]) < 3.0e38f) ? 0.f : 1.f;
}

}  // namespace

bool gram_tc_supported(int64_t n, int d, int k) {
  // (up to 4 x 4 tiles of 128: 26 tile types, still >= 5 slabs of rows on 148 SMs)
  return n >= 4096 && k >= 1 && k <= 512 && d >= 1 && d <= 512;
}

// scratch: 64 bytes of device memory.  gzz / gzx must be zero on entry.
int gram_tc_run(const float* z, const float* x, int64_t n, int d, int k, double* gzz, double* gzx, void* scratch,
                cudaStream_t st) {
  int dev = 0, sms = 148;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  LASSO_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  unsigned* maxbits = reinterpret_cast<unsigned*>(scratch);
  float* scales = reinterpret_cast<float*>(maxbits + 2);
  LASSO_CUDA_TRY(cudaMemsetAsync(maxbits, 0, 2 * sizeof(unsigned), st));
  gram_maxabs_kernel<<<sms * 4, 512, 0, st>>>(z, n * k, x, n * d, maxbits);
  gram_scales_kernel<<<1, 1, 0, st>>>(maxbits, scales);
  GramTcParams p{};
  p.z = z;
  p.x = x;
  p.n = n;
  p.d = d;
  p.k = k;
  p.ntypes = gram_ntypes(d, k);
  int64_t slabs = std::max<int64_t>(1, sms / p.ntypes);
  int64_t rows = (n + slabs - 1) / slabs;
  rows = ((rows + kGtRows - 1) / kGtRows) * kGtRows;
  slabs = (n + rows - 1) / rows;
  p.slab_rows = rows;
  p.scales = scales;
  p.gzz = gzz;
  p.gzx = gzx;
  const unsigned grid = (unsigned)(slabs * p.ntypes);
  LASSO_CUDA_TRY(cudaFuncSetAttribute(gram_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGtSmem));
  gram_tc_kernel<<<grid, kGtThreads, kGtSmem, st>>>(p);
  LASSO_CHECK_LAUNCH();
  count_launch(3);
  return LASSO_B200_OK;
}

}  // namespace lasso
