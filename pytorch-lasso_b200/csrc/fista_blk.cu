// k-blocked streaming tcgen05 FISTA step for sm_100a: dictionaries too large to keep a tile's
// state on one SM (BASELINE config 3: d = 128, k = 1024).
//
// Shapes: d <= 128, k <= 1024, d % 4 == 0, k % 4 == 0 (TMA row pitch).  One persistent launch per
// iteration; a CTA walks over 128-row tiles, and over each tile twice in 64-atom chunks:
//
//   pass 1   for every chunk q: TMA-staged z_cur / z_prev chunk -> y = z_cur + beta (z_cur - z_prev)
//            (ista.py:100) -> fp16 pieces -> TMEM slot -> GEMM1 slice  R += Y_q W_q^T   (N = 128)
//   phase B  r = R - x (x from global / L2, rescaled) -> fp16 pieces -> TMEM
//   pass 2   for every chunk q: GEMM2 chunk  G = r W_q  (N = 64), then z+ = softshrink(y - lr g,
//            lam) (ista.py:90) with y recomputed from the re-staged z chunks, stop-test sum
//            |z_cur - z+| (ista.py:93), z+ written over z_prev in HBM
//
// The dictionary slices (two fp16 piece images per 64 atoms, 32 KB) stream from L2 through the
// same two-stage TMA ring as the code chunks; one shared-memory image serves GEMM1 (K-major view)
// and GEMM2 (MN-major view), as in the other tcgen05 kernels.  Operands are the fp16x2 split of
// the per-row rescaled problem (see fista_res.cu): codes live in HBM in scaled units for the whole
// solve and are converted back by blk_unscale_kernel, which also raises the hand-over flag when an
// iterate left the fp16 range.
//
// HBM traffic per iteration: n (d + 5 k) floats (both code buffers are read in both passes).
// TMEM columns: R_big 128 | R_small 128 | r pieces 128 | G 64 | piece slot 64 = 512.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace lasso {
namespace {

using namespace sm100;

constexpr int kTileM = 128;
constexpr int kQ = 64;                  // atoms per chunk
constexpr int kDPB = 128;               // padded d
constexpr int kKMaxB = 1024;
constexpr int kThreadsB = 576;          // warps 0..15 compute, 16 MMA issuer, 17 TMA producer
constexpr uint32_t kBoxBytes = kTileM * 128;            // [128 rows][32 atoms fp32] = 16 KB
constexpr uint32_t kZChunkBytes = 2 * kBoxBytes;        // 64 atoms
constexpr uint32_t kWPieceBytes = kDPB * 128;           // [128 features][64 atoms fp16] = 16 KB
constexpr uint32_t kWSliceBytes = 2 * kWPieceBytes;     // h, l
constexpr uint32_t kStageBytes = 2 * kZChunkBytes + kWSliceBytes;   // z_cur, z_prev, W slice: 96 KB
constexpr int kStagesB = 2;
constexpr uint32_t kSmemBytesB = kStagesB * kStageBytes;            // 192 KB

constexpr uint32_t kColR = 0;       // R_big   = sum of the leading products h h' (128 features)
constexpr uint32_t kColRs = 128;    // R_small = sum of the cross products h l' + l h'.  Two accumulators:
                                    // the tensor core truncates on every accumulate and with k = 1024
                                    // (192 accumulations) a single one costs a factor 2-3 in accuracy
constexpr uint32_t kColRp = 256;    // r pieces: h 64 cols | l 64 cols
constexpr uint32_t kColG = 384;     // G chunk, 64 atoms
constexpr uint32_t kColS = 448;     // piece slot: [h 32 cols | l 32 cols]
constexpr uint32_t kTmemColsB = 512;

struct BlkScalars {
  float sw, isw, lr, lam;   // W' = sw W;  lr / sw^2;  lam / sw (times the row's x scale)
  int bad;
};

struct BlkParams {
  const uint8_t* w_image;   // [nq][2][kWPieceBytes], scaled fp16 pieces
  const float* x;           // [n][d]
  const float* row_scale;   // [n] power-of-two scale of every row of x
  const float* z_cur;       // scaled codes z_i
  float* z_io;              // z_{i-1} on entry, z_{i+1} on exit (scaled)
  int64_t n;
  int d, k;
  float beta;
  int use_prev;
  int trows;                // rows per tile: 128 (64 for experiments: half of the MMA rows idle)
  int l2_hints;             // TMA cache hints (LASSO_B200_BLK_L2=1 enables)
  int mode;                 // 0: whole iteration.  Convolutional lasso (patches of an image overlap, so
                            // the residual is formed in image space by conv_resid_kernel between the
                            // two halves): 1 = pass 1 only, R -> r_buf;  2 = r <- r_buf, pass 2 only
  float* r_buf;             // [n][d] (modes 1, 2)
  const BlkScalars* scal;
  StepCtl ctl;
  volatile int* dbg;
};

__device__ __noinline__ void blk_wait_slow_path(uint64_t& t0, volatile int* dbg, int line, int iter, uint32_t parity) {
  const uint64_t now = global_timer_ns();
  if (t0 == 0) {
    t0 = now;
    return;
  }
  if (now - t0 < 4000000000ull) return;
  if (dbg) {
    dbg[1] = line; dbg[2] = blockIdx.x; dbg[3] = threadIdx.x; dbg[4] = iter; dbg[5] = (int)parity;
    __threadfence_system();
    dbg[0] = 1;
    __threadfence_system();
  }
  __trap();
}
#define BLK_WAIT(bar, parity)                                                             \
  do {                                                                                    \
    const uint32_t _addr = smem_u32(bar), _par = (parity) & 1u;                           \
    uint32_t _ok, _n = 0;                                                                 \
    uint64_t _t0 = 0;                                                                     \
    for (;;) {                                                                            \
      asm volatile(                                                                       \
          "{\n\t.reg .pred P;\n\t"                                                       \
          "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"                  \
          "selp.b32 %0, 1, 0, P;\n\t}\n"                                                   \
          : "=r"(_ok)                                                                     \
          : "r"(_addr), "r"(_par), "r"(20000u)                                            \
          : "memory");                                                                    \
      if (_ok) break;                                                                     \
      if ((++_n & 1023u) == 0) blk_wait_slow_path(_t0, p.dbg, __LINE__, p.ctl.iter, _par); \
    }                                                                                     \
  } while (0)

__device__ __forceinline__ float2 bsub2(float2 a, float2 b) {
  return __ffma2_rn(make_float2(-1.f, -1.f), b, a);
}
// fp32 pair -> packed fp16 pieces (see fista_res.cu)
__device__ __forceinline__ void bsplit2(float2 v, uint32_t& wh, uint32_t& wl) {
  const float2 t = make_float2(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                               __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
  const float2 r = bsub2(v, t);
  const __half2 h = __floats2half2_rn(t.x, t.y);
  const __half2 l = __floats2half2_rn(r.x, r.y);
  wh = *reinterpret_cast<const uint32_t*>(&h);
  wl = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(kThreadsB, 1)
fista_blk_kernel(const __grid_constant__ CUtensorMap tm_cur, const __grid_constant__ CUtensorMap tm_prev, BlkParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[kStagesB], bar_empty[kStagesB], bar_aready, bar_sfree;
  __shared__ uint64_t bar_rfull, bar_rready, bar_gfull, bar_gfree;
  __shared__ uint32_t tmem_base_s;

  // an earlier iteration already met the stop test -> this launch is a no-op (ista.py:93-95)
  if (p.ctl.tol_abs >= 0.0 && p.ctl.iter >= 1 && p.ctl.hist[p.ctl.iter - 1] <= p.ctl.tol_abs) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nq = (p.k + kQ - 1) / kQ;
  const int dsteps = (p.d + 15) >> 4;
  const int trows = p.trows;
  const int ntiles = (int)((p.n + trows - 1) / trows);
  const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (tid == 0) {
    for (int s = 0; s < kStagesB; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 513);   // 512 compute threads (done reading z) + 1 MMA commit (done reading W)
    }
    mbar_init(&bar_aready, 512);
    mbar_init(&bar_sfree, 1);
    mbar_init(&bar_rfull, 1);
    mbar_init(&bar_rready, 512);
    mbar_init(&bar_gfull, 1);
    mbar_init(&bar_gfree, 512);
    fence_mbar_init();
  }
  if (warp == 16) {
    tmem_alloc(&tmem_base_s, kTmemColsB);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 17) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      prefetch_tmap(&tm_cur);
      prefetch_tmap(&tm_prev);
    }
    __syncwarp();
    uint32_t cc = 0;   // chunk loads issued (2 nq per tile)
    for (int tile = 0; tile < my_tiles; ++tile) {
      const int row0 = ((int)blockIdx.x + tile * (int)gridDim.x) * trows;
      const uint32_t box_bytes = (uint32_t)trows * 128u;
      for (int pass = 0; pass < 2; ++pass) {
        if ((pass == 0 && p.mode == 2) || (pass == 1 && p.mode == 1)) continue;
        for (int q = 0; q < nq; ++q, ++cc) {
          const uint32_t s = cc & 1u, ph = (cc >> 1) & 1u;
          BLK_WAIT(&bar_empty[s], ph ^ 1u);
          if (elect_one()) {
            uint8_t* st = smem + s * kStageBytes;
            mbar_expect_tx(&bar_full[s], 4u * box_bytes + kWSliceBytes);
            // L2 priorities: z_cur of pass 1 is read again ~40 us later in pass 2 -> keep it; the
            // rest is touched for the last time in this launch -> first to go
            const uint64_t pol_cur = (pass == 0 && p.l2_hints) ? kEvictLast : (p.l2_hints ? kEvictFirst : kEvictNormal);
            const uint64_t pol_prev = p.l2_hints ? kEvictFirst : kEvictNormal;
            tma_load_2d_hint(st, &tm_cur, q * kQ, row0, &bar_full[s], pol_cur);
            tma_load_2d_hint(st + kBoxBytes, &tm_cur, q * kQ + 32, row0, &bar_full[s], pol_cur);
            tma_load_2d_hint(st + kZChunkBytes, &tm_prev, q * kQ, row0, &bar_full[s], pol_prev);
            tma_load_2d_hint(st + kZChunkBytes + kBoxBytes, &tm_prev, q * kQ + 32, row0, &bar_full[s], pol_prev);
            bulk_load(st + 2 * kZChunkBytes, p.w_image + (size_t)q * kWSliceBytes, 16384, &bar_full[s]);
            bulk_load(st + 2 * kZChunkBytes + 16384, p.w_image + (size_t)q * kWSliceBytes + 16384, 16384, &bar_full[s]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 16) {
    // ===================== MMA issuer =====================
    const uint32_t idesc1 = make_idesc(kFmtF16, 128, kDPB, 0, 0);   // B K-major  (GEMM1, N = 128 features)
    const uint32_t idesc2 = make_idesc(kFmtF16, 128, kQ, 0, 1);     // B MN-major (GEMM2, N = 64 atoms)
    constexpr uint32_t kPiece16 = kWPieceBytes >> 4;
    uint32_t cc = 0, pc = 0, gc = 0, ti = 0;   // chunk loads / piece chunks / G chunks / tiles seen
    for (int tile = 0; tile < my_tiles; ++tile, ++ti) {
      // ---- pass 1: GEMM1 slices ----
      for (int q = 0; q < (p.mode == 2 ? 0 : nq); ++q, ++cc, ++pc) {
        const uint32_t s = cc & 1u;
        BLK_WAIT(&bar_full[s], (cc >> 1) & 1u);       // dictionary slice landed
        BLK_WAIT(&bar_aready, pc & 1u);               // pieces of y_q staged
        tc_fence_after();
        if (elect_one()) {
          const uint64_t desc = make_smem_desc_sw128(smem_u32(smem + s * kStageBytes + 2 * kZChunkBytes), 0, 1024);
          const uint32_t d_lo = (uint32_t)desc, d_hi = (uint32_t)(desc >> 32);
          const uint32_t t_slot = tbase + kColS;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t qh = ((uint64_t)d_hi << 32) | (d_lo + ks * 2);
            const uint64_t ql = ((uint64_t)d_hi << 32) | (d_lo + ks * 2 + kPiece16);
            const uint32_t ah = t_slot + ks * 8, al = ah + 32;
            const uint32_t acc_on = (q > 0 || ks > 0) ? 1u : 0u;
            mma_ts<false>(tbase + kColRs, ah, ql, idesc1, acc_on);
            mma_ts<false>(tbase + kColRs, al, qh, idesc1, 1);
            mma_ts<false>(tbase + kColR, ah, qh, idesc1, acc_on);
          }
          mma_commit(&bar_sfree);
          mma_commit(&bar_empty[s]);
          if (q == nq - 1) mma_commit(&bar_rfull);
        }
        __syncwarp();
      }
      // ---- pass 2: GEMM2 chunks ----
      BLK_WAIT(&bar_rready, ti & 1u);   // (mode 1: R has been read, the next tile may overwrite it)
      tc_fence_after();
      for (int q = 0; q < (p.mode == 1 ? 0 : nq); ++q, ++cc, ++gc) {
        const uint32_t s = cc & 1u;
        BLK_WAIT(&bar_full[s], (cc >> 1) & 1u);
        if (gc > 0) BLK_WAIT(&bar_gfree, (gc - 1) & 1u);   // the single G buffer has been drained
        tc_fence_after();
        if (elect_one()) {
          const uint64_t desc = make_smem_desc_sw128(smem_u32(smem + s * kStageBytes + 2 * kZChunkBytes), kWPieceBytes, 1024);
          const uint32_t d_lo = (uint32_t)desc, d_hi = (uint32_t)(desc >> 32);
          uint32_t acc_on = 0;
          // small products first (l h', h l'), leading product last
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              if (ks < dsteps) {
                const uint64_t bd = ((uint64_t)d_hi << 32) | (d_lo + pb[t] * kPiece16 + ks * 128);
                mma_ts<false>(tbase + kColG, tbase + kColRp + pa[t] * 64 + ks * 8, bd, idesc2, acc_on);
                acc_on = 1;
              }
            }
          }
          mma_commit(&bar_gfull);
          mma_commit(&bar_empty[s]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== compute warps =====================
    const int quad = warp & 3;
    const int wg = warp >> 2;                  // atoms [16 wg, +16) of a chunk / features [32 wg, +32)
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const BlkScalars sc = *p.scal;
    const float2 nlr2 = make_float2(-sc.lr, -sc.lr);
    const float2 beta2 = make_float2(p.beta, p.beta);
    const uint32_t zbox = (uint32_t)(wg >> 1) * kBoxBytes;   // which 32-atom box of the chunk
    const uint32_t zbyte = (uint32_t)(wg & 1) * 64;          // byte offset of this thread's 16 atoms in the box row
    uint32_t cc = 0, pc = 0, gc = 0, ti = 0;
    double dsum = 0.0;
    for (int tile = 0; tile < my_tiles; ++tile, ++ti) {
      const int64_t grow = (int64_t)((int)blockIdx.x + tile * (int)gridDim.x) * trows + row;
      const bool row_ok = row < trows && grow < p.n;
      const float sxr = row_ok ? __ldg(p.row_scale + grow) : 1.f;
      const float lam = sc.lam * sxr;
      const float uz_row = sc.sw / sxr;
      // ---------------- pass 1: y chunks -> pieces ----------------
      for (int q = 0; q < (p.mode == 2 ? 0 : nq); ++q, ++cc, ++pc) {
        const uint32_t s = cc & 1u;
        BLK_WAIT(&bar_full[s], (cc >> 1) & 1u);
        const uint8_t* zc_s = smem + s * kStageBytes + zbox;
        const uint8_t* zp_s = zc_s + kZChunkBytes;
        uint32_t wh[8], wl[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw128_offset(row, zbyte + j * 16);
          const float4 zc = *reinterpret_cast<const float4*>(zc_s + off);
          float2 ya = make_float2(zc.x, zc.y), yb = make_float2(zc.z, zc.w);
          if (p.use_prev) {
            const float4 zp = *reinterpret_cast<const float4*>(zp_s + off);
            ya = __ffma2_rn(beta2, bsub2(ya, make_float2(zp.x, zp.y)), ya);
            yb = __ffma2_rn(beta2, bsub2(yb, make_float2(zp.z, zp.w)), yb);
          }
          bsplit2(ya, wh[2 * j], wl[2 * j]);
          bsplit2(yb, wh[2 * j + 1], wl[2 * j + 1]);
        }
        mbar_arrive(&bar_empty[s]);                       // done reading the stage
        if (pc >= 1) BLK_WAIT(&bar_sfree, (pc - 1u) & 1u);   // slot consumed by the GEMM1 slice before
        tc_fence_after();
        const uint32_t t_slot = tbase + lane_base + kColS + wg * 8;
        tmem_st8(t_slot, wh);
        tmem_st8(t_slot + 32, wl);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_aready);
      }
      // ---------------- phase B: r = (R_big + R_small) - x -> pieces (32 features per thread) ----------------
      if (p.mode == 2) {
        // the residual comes from conv_resid_kernel (image space), already in scaled units
        float4 rv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = wg * 32 + 4 * j;
          rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && col < p.d) rv[j] = __ldg(reinterpret_cast<const float4*>(p.r_buf + grow * p.d + col));
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t wh[8], wl[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 rr = rv[hf * 4 + j];
            bsplit2(make_float2(rr.x, rr.y), wh[2 * j], wl[2 * j]);
            bsplit2(make_float2(rr.z, rr.w), wh[2 * j + 1], wl[2 * j + 1]);
          }
          const uint32_t t_r = tbase + lane_base + kColRp + wg * 16 + hf * 8;
          tmem_st8(t_r, wh);
          tmem_st8(t_r + 64, wl);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_rready);
      } else {
        float4 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = wg * 32 + 4 * j;
          xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (p.mode == 0 && row_ok && col < p.d) {
            xv[j] = __ldg(reinterpret_cast<const float4*>(p.x + grow * p.d + col));
            xv[j].x *= sxr; xv[j].y *= sxr; xv[j].z *= sxr; xv[j].w *= sxr;
          }
        }
        BLK_WAIT(&bar_rfull, ti & 1u);
        tc_fence_after();
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t rb[16], rs[16];
          tmem_ld16(tbase + lane_base + kColR + wg * 32 + hf * 16, rb);
          tmem_ld16(tbase + lane_base + kColRs + wg * 32 + hf * 16, rs);
          tmem_wait_ld();
          uint32_t wh[8], wl[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 xx = xv[hf * 4 + j];
            const float2 ra = bsub2(__fadd2_rn(make_float2(__uint_as_float(rb[4 * j + 0]), __uint_as_float(rb[4 * j + 1])),
                                               make_float2(__uint_as_float(rs[4 * j + 0]), __uint_as_float(rs[4 * j + 1]))),
                                    make_float2(xx.x, xx.y));
            const float2 rc = bsub2(__fadd2_rn(make_float2(__uint_as_float(rb[4 * j + 2]), __uint_as_float(rb[4 * j + 3])),
                                               make_float2(__uint_as_float(rs[4 * j + 2]), __uint_as_float(rs[4 * j + 3]))),
                                    make_float2(xx.z, xx.w));
            if (p.mode == 1) {
              // R = Y W^T (scaled units) goes to HBM: the overlap-add happens in image space
              const int col = wg * 32 + hf * 16 + 4 * j;
              if (row_ok && col < p.d)
                *reinterpret_cast<float4*>(p.r_buf + grow * p.d + col) = make_float4(ra.x, ra.y, rc.x, rc.y);
            } else {
              bsplit2(ra, wh[2 * j], wl[2 * j]);
              bsplit2(rc, wh[2 * j + 1], wl[2 * j + 1]);
            }
          }
          if (p.mode == 0) {
            const uint32_t t_r = tbase + lane_base + kColRp + wg * 16 + hf * 8;
            tmem_st8(t_r, wh);
            tmem_st8(t_r + 64, wl);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_rready);
      }
      // ---------------- pass 2: fused update, z+ over z_prev in HBM ----------------
      float part = 0.f;
      for (int q = 0; q < (p.mode == 1 ? 0 : nq); ++q, ++cc, ++gc) {
        const uint32_t s = cc & 1u;
        BLK_WAIT(&bar_full[s], (cc >> 1) & 1u);
        const uint8_t* zc_s = smem + s * kStageBytes + zbox;
        const uint8_t* zp_s = zc_s + kZChunkBytes;
        float2 zc2[8], y2[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw128_offset(row, zbyte + j * 16);
          const float4 zc = *reinterpret_cast<const float4*>(zc_s + off);
          zc2[2 * j] = make_float2(zc.x, zc.y);
          zc2[2 * j + 1] = make_float2(zc.z, zc.w);
          y2[2 * j] = zc2[2 * j];
          y2[2 * j + 1] = zc2[2 * j + 1];
          if (p.use_prev) {
            const float4 zp = *reinterpret_cast<const float4*>(zp_s + off);
            y2[2 * j] = __ffma2_rn(beta2, bsub2(zc2[2 * j], make_float2(zp.x, zp.y)), zc2[2 * j]);
            y2[2 * j + 1] = __ffma2_rn(beta2, bsub2(zc2[2 * j + 1], make_float2(zp.z, zp.w)), zc2[2 * j + 1]);
          }
        }
        mbar_arrive(&bar_empty[s]);                       // the stage may be refilled
        BLK_WAIT(&bar_gfull, gc & 1u);
        tc_fence_after();
        uint32_t g[16];
        tmem_ld16(tbase + lane_base + kColG + wg * 16, g);
        tmem_wait_ld();
        tc_fence_before();
        mbar_arrive(&bar_gfree);
        float* outp = p.z_io + grow * p.k + q * kQ + wg * 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 zo[2];
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const float2 gg = make_float2(__uint_as_float(g[4 * j + 2 * h2]), __uint_as_float(g[4 * j + 2 * h2 + 1]));
            const float2 v = __ffma2_rn(nlr2, gg, y2[2 * j + h2]);
            const float2 c = make_float2(fminf(fmaxf(v.x, -lam), lam), fminf(fmaxf(v.y, -lam), lam));
            const float2 zn = bsub2(v, c);
            const float2 dl = bsub2(zn, zc2[2 * j + h2]);
            part += fabsf(dl.x) + fabsf(dl.y);
            zo[h2] = zn;
          }
          const int col = q * kQ + wg * 16 + 4 * j;
          if (row_ok && col < p.k) *reinterpret_cast<float4*>(outp + 4 * j) = make_float4(zo[0].x, zo[0].y, zo[1].x, zo[1].y);
        }
      }
      if (row_ok) dsum += (double)(part * uz_row);
    }
    if (p.ctl.hist != nullptr && p.mode != 1) {
      dsum = warp_sum(dsum);
      if (lane == 0 && dsum != 0.0) atomicAdd(&p.ctl.hist[p.ctl.iter], dsum);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tbase, kTmemColsB);
}

// ---- set-up / tear-down kernels ------------------------------------------------------------
__global__ void blk_setup_kernel(const float* __restrict__ w, int nw, float lr, float lam, BlkScalars* __restrict__ sc,
                                 int* __restrict__ flag) {
  __shared__ unsigned s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  unsigned mw = 0;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) mw = max(mw, __float_as_uint(w[i]) & 0x7FFFFFFFu);
  mw = __reduce_max_sync(0xffffffffu, mw);
  if ((threadIdx.x & 31) == 0 && mw) atomicMax(&s_max, mw);
  __syncthreads();
  if (threadIdx.x != 0) return;
  const float aw = __uint_as_float(s_max);
  int bad = 0, ew = 0;
  if (!(aw < 3.0e38f)) bad = 1;
  if (aw > 0.f && !bad) ew = 3 - ilogbf(aw);    // max |W'| in [8, 16)
  ew = max(-40, min(40, ew));
  sc->sw = ldexpf(1.f, ew);
  sc->isw = ldexpf(1.f, -ew);
  sc->lr = ldexpf(lr, -2 * ew);
  sc->lam = ldexpf(lam, -ew);
  if (!(sc->lr > 0.f) || !(sc->lr < 3.0e38f) || !(sc->lam < 3.0e38f) || (lam > 0.f && !(sc->lam > 0.f))) bad = 1;
  sc->bad = bad;
  *flag = bad;
}

// dictionary [d][k] fp32 -> per 64-atom chunk two scaled fp16 piece images [128 features][128 B]
__global__ void blk_prep_w_kernel(const float* __restrict__ w, int d, int k, int nq, const BlkScalars* __restrict__ sc,
                                  uint8_t* __restrict__ image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (feature i, padded atom j)
  if (idx >= kDPB * nq * kQ) return;
  const int i = idx / (nq * kQ), j = idx % (nq * kQ);
  const float v = (i < d && j < k) ? w[(int64_t)i * k + j] * sc->sw : 0.f;
  const __half h = __float2half_rn(__uint_as_float(__float_as_uint(v) & 0xFFFFE000u));
  const __half l = __float2half_rn(v - __half2float(h));
  const size_t off = (size_t)(j / kQ) * kWSliceBytes + sw128_offset(i, (j % kQ) * 2);
  *reinterpret_cast<__half*>(image + off) = h;
  *reinterpret_cast<__half*>(image + kWPieceBytes + off) = l;
}

// per-row power-of-two scale of x (max |x'_r| in [64, 128)); one warp per row.  Scales the start
// codes (z_a) into the kernel's units in the same pass.
__global__ void blk_rowscale_kernel(const float* __restrict__ x, int64_t n, int d, int k, const BlkScalars* __restrict__ sc,
                                    float* __restrict__ row_scale, float* __restrict__ z_a, int scale_z) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  float m = 0.f;
  for (int c = lane; c < d; c += 32) m = fmaxf(m, fabsf(x[r * d + c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sxr = 1.f;
  if (m > 0.f && m < 3.0e38f) {
    const int be = 260 - (int)((__float_as_uint(m) >> 23) & 0xFFu);
    sxr = __uint_as_float((uint32_t)min(max(be, 1), 254) << 23);
  }
  if (lane == 0) row_scale[r] = sxr;
  if (scale_z) {
    const float szr = sxr * sc->isw;
    for (int c = lane; c < k; c += 32) z_a[r * k + c] *= szr;
  }
}

// codes back to the caller's units (one warp per row); raises the flag when an operand left the
// fp16 range
__global__ void blk_unscale_kernel(float* __restrict__ z, int64_t n, int k, const float* __restrict__ row_scale,
                                   const BlkScalars* __restrict__ sc, float limit, int* __restrict__ flag) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  bool bad = false;
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < n; r += warps) {
    const float uz = sc->sw / row_scale[r];
    float* zr = z + r * k;
    for (int c = lane; c < k; c += 32) {
      const float v = zr[c];
      bad |= !(fabsf(v) < limit);
      zr[c] = v * uz;
    }
  }
  if (bad) atomicExch(flag, 1);
}


// ---- convolutional lasso (lasso/conv2d/ista.py:7-49) as im2col -> linear ---------------------
// rows = the oh x ow patches of every image, features = cin x kh x kw, atoms = filters:
//   conv_transpose2d(z, W) = fold(Z W_lin^T),  conv2d(r, W) = unfold(r) W_lin  (stride 1, no padding)
// so an iteration is pass 1 (R = Y W_lin^T, mode 1), the residual in IMAGE space
// r = unfold(fold(R) - x) -- patches overlap, rows are coupled inside an image -- and pass 2 (mode 2).

// per-image power-of-two scale (all patch rows of an image share it): one CTA per image
__global__ void __launch_bounds__(256) conv_scale_kernel(const float* __restrict__ x, int img_elems, int P, int k,
                                                         const BlkScalars* __restrict__ sc,
                                                         float* __restrict__ row_scale, float* __restrict__ z_a,
                                                         int scale_z) {
  __shared__ float red[8];
  __shared__ float s_sx;
  const int64_t img = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float m = 0.f;
  for (int i = tid; i < img_elems; i += blockDim.x) m = fmaxf(m, fabsf(x[img * img_elems + i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) {
    float mm = 0.f;
    for (int i = 0; i < 8; ++i) mm = fmaxf(mm, red[i]);
    float sxr = 1.f;
    if (mm > 0.f && mm < 3.0e38f) {
      const int be = 260 - (int)((__float_as_uint(mm) >> 23) & 0xFFu);
      sxr = __uint_as_float((uint32_t)min(max(be, 1), 254) << 23);
    }
    s_sx = sxr;
  }
  __syncthreads();
  const float sxr = s_sx;
  for (int pch = tid; pch < P; pch += blockDim.x) row_scale[img * P + pch] = sxr;
  if (scale_z) {
    const float szr = sxr * sc->isw;
    float* zi = z_a + img * (int64_t)P * k;
    for (int64_t i = tid; i < (int64_t)P * k; i += blockDim.x) zi[i] *= szr;
  }
}

// r = unfold(fold(R) - sx x), in place on the [P][d] block of one image; one CTA per image.
// Patch (pi, pj) covers the pixels (stride pi + a - pad, stride pj + b - pad): conv_transpose2d crops the
// `pad` border of the overlap-add and conv2d pads the residual with zeros, so border entries of a patch
// that fall outside the image contribute nothing and read back as zero (conv2d/ista.py:18-19).
// Shared-memory rows of R are d + 1 floats apart: in the overlap-add neighbouring threads read
// neighbouring patches (a stride of d floats would put a whole warp on one bank).  blockDim.x is a
// multiple of d, so a thread keeps its feature (c, a, b) and walks the patches without divisions.
__host__ __device__ inline int conv_row_ld(int d) { return d | 1; }
template <bool kUnitStride>
__global__ void __launch_bounds__(1024) conv_resid_kernel(float* __restrict__ rbuf, const float* __restrict__ x,
                                                          const float* __restrict__ row_scale, ConvShape cs,
                                                          StepCtl ctl) {
  extern __shared__ __align__(16) float conv_smem[];
  if (ctl.tol_abs >= 0.0 && ctl.iter >= 1 && ctl.hist[ctl.iter - 1] <= ctl.tol_abs) return;
  const int cin = cs.cin, H = cs.h, W = cs.w, kh = cs.kh, kw = cs.kw, st = cs.stride, pad = cs.pad;
  const int oh = cs.oh(), ow = cs.ow(), P = oh * ow, kk = kh * kw, d = cin * kk, ld = conv_row_ld(d);
  float* Rs = conv_smem;              // [P][ld]
  float* img = conv_smem + P * ld;    // [cin][H][W]
  const int64_t i_img = blockIdx.x;
  float* rb = rbuf + i_img * (int64_t)P * d;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int f = tid % d, grp = tid / d, ngrp = nthr / d;     // this thread's feature, its first patch
  {
    // d % 4 == 0: float4 loads (eight in flight per thread), scalar stores into the padded rows
    const int q4 = d >> 2, total = P * q4;
    for (int e0 = 0; e0 < total; e0 += 8 * nthr) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * nthr + tid;
        if (e < total) v[u] = __ldcs(reinterpret_cast<const float4*>(rb) + e);   // (rewritten below: no .nc load)
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int e = e0 + u * nthr + tid;
        if (e < total) {
          float* dst = Rs + (e / q4) * ld + (e % q4) * 4;
          dst[0] = v[u].x; dst[1] = v[u].y; dst[2] = v[u].z; dst[3] = v[u].w;
        }
      }
    }
  }
  const float sx = row_scale[i_img * P];
  __syncthreads();
  for (int pix = tid; pix < cin * H * W; pix += nthr) {
    const int c = pix / (H * W), i = (pix / W) % H, j = pix % W;
    float acc = 0.f;
    if (kUnitStride) {
      // patches (i + pad - a, j + pad - b) inside the code grid: contiguous ranges of a and b, no tests inside
      const int a_lo = max(0, i + pad - (oh - 1)), a_hi = min(kh - 1, i + pad);
      const int b_lo = max(0, j + pad - (ow - 1)), b_hi = min(kw - 1, j + pad);
      for (int a = a_lo; a <= a_hi; ++a) {
        const float* rrow = Rs + ((i + pad - a) * ow + (j + pad)) * ld + c * kk + a * kw;
        for (int b = b_lo; b <= b_hi; ++b) acc += rrow[b - b * ld];
      }
    } else {
      for (int a = 0; a < kh; ++a) {
        const int ti = i + pad - a;
        if (ti < 0 || (ti % st) != 0 || ti / st >= oh) continue;
        const float* rrow = Rs + (ti / st) * ow * ld + c * kk + a * kw;
        for (int b = 0; b < kw; ++b) {
          const int tj = j + pad - b;
          if (tj < 0 || (tj % st) != 0 || tj / st >= ow) continue;
          acc += rrow[(tj / st) * ld + b];
        }
      }
    }
    img[pix] = acc - sx * __ldg(x + i_img * (int64_t)(cin * H * W) + pix);
  }
  __syncthreads();
  {
    const int c = f / kk, a = (f / kw) % kh, b = f % kw;
    const float* ibase = img + c * H * W;
    int pi = grp / ow, pj = grp % ow;
    const int dpi = ngrp / ow, dpj = ngrp % ow;
    for (int pch = grp; pch < P; pch += ngrp) {
      const int i = pi * st + a - pad, j = pj * st + b - pad;
      rb[(int64_t)pch * d + f] = (i >= 0 && i < H && j >= 0 && j < W) ? ibase[i * W + j] : 0.f;
      pi += dpi;
      pj += dpj;
      if (pj >= ow) {
        pj -= ow;
        ++pi;
      }
    }
  }
}

typedef CUresult (*EncodeTiledFnB)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int blk_make_map(CUtensorMap* map, const float* base, int64_t rows, int cols, int tile_rows) {
  static EncodeTiledFnB fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnB)ptr;
  }
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return LASSO_B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)tile_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%d)", (int)r, (long long)rows, cols);
    return LASSO_B200_ERR_CUDA;
  }
  return LASSO_B200_OK;
}

struct BlkState {
  uint8_t* w_image = nullptr;
  BlkScalars* scal = nullptr;
  int* flag = nullptr;
  float* row_scale = nullptr;
  int64_t row_cap = 0;
  float* r_buf = nullptr;     // [rows][d] residual of the convolutional path
  size_t r_cap = 0;
  bool conv_attr_set = false;
  int num_sms = 0;
  bool attr_set = false;
  int* dbg_host = nullptr;
  int* dbg_dev = nullptr;
};
BlkState g_blk[64];

}  // namespace

bool fista_blk_supported(int64_t n, int d, int k) {
  return n >= 1 && d >= 4 && d <= kDPB && k >= 4 && k <= kKMaxB && (d % 4) == 0 && (k % 4) == 0 &&
         n < ((int64_t)1 << 31) - kTileM;
}

bool conv2d_blk_supported(const ConvShape& c, int k) {
  if (c.n_img < 1 || c.cin < 1 || c.kh < 1 || c.kw < 1 || c.stride < 1 || c.pad < 0 || c.h + 2 * c.pad < c.kh ||
      c.w + 2 * c.pad < c.kw || (c.h + 2 * c.pad - c.kh) % c.stride != 0 || (c.w + 2 * c.pad - c.kw) % c.stride != 0)
    return false;
  const int64_t P = (int64_t)c.oh() * c.ow();
  const int d = c.cin * c.kh * c.kw;
  return fista_blk_supported(c.n_img * P, d, k) &&
         (size_t)(P * (d | 1) + (int64_t)c.cin * c.h * c.w) * sizeof(float) <= 200 * 1024;   // conv_row_ld(d)
}

// Same contract as fista_tc_run: z_i lives in (i even ? z_a : z_b).  *fell_back = 1 when an
// iterate left the fp16 operand range (the buffers are then unspecified).  Synchronises the stream.
// conv != nullptr: a.x is the image batch [n_img][cin][h][w], rows are its patches (a.n = n_img * P,
// a.d = cin * kh * kw) and every iteration is pass 1, conv_resid_kernel, pass 2.
int fista_blk_run(const FistaArgs& a, int* fell_back, cudaStream_t st, const ConvShape* conv) {
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  BlkState& S = g_blk[dev];
  if (!S.w_image) {
    LASSO_CUDA_TRY(cudaMalloc(&S.w_image, (size_t)(kKMaxB / kQ) * kWSliceBytes));
    LASSO_CUDA_TRY(cudaMalloc(&S.scal, sizeof(BlkScalars)));
    LASSO_CUDA_TRY(cudaMalloc(&S.flag, sizeof(int)));
    LASSO_CUDA_TRY(cudaDeviceGetAttribute(&S.num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (a.n > S.row_cap) {
    if (S.row_scale) LASSO_CUDA_TRY(cudaFree(S.row_scale));
    S.row_scale = nullptr;
    S.row_cap = 0;
    LASSO_CUDA_TRY(cudaMalloc(&S.row_scale, sizeof(float) * (size_t)a.n));
    S.row_cap = a.n;
  }
  if (!S.dbg_host && getenv("LASSO_B200_DEBUG")) {
    LASSO_CUDA_TRY(cudaHostAlloc((void**)&S.dbg_host, 4096, cudaHostAllocMapped));
    memset(S.dbg_host, 0, 4096);
    LASSO_CUDA_TRY(cudaHostGetDevicePointer((void**)&S.dbg_dev, S.dbg_host, 0));
  }
  if (!S.attr_set) {
    LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)fista_blk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kSmemBytesB));
    S.attr_set = true;
  }
  const int nq = (a.k + kQ - 1) / kQ;
  blk_setup_kernel<<<1, 256, 0, st>>>(a.w, a.d * a.k, a.lr, a.lam, S.scal, S.flag);
  LASSO_CHECK_LAUNCH();
  blk_prep_w_kernel<<<(kDPB * nq * kQ + 255) / 256, 256, 0, st>>>(a.w, a.d, a.k, nq, S.scal, S.w_image);
  LASSO_CHECK_LAUNCH();
  int conv_P = 0;
  size_t conv_smem = 0;
  if (conv) {
    conv_P = conv->oh() * conv->ow();
    conv_smem = ((size_t)conv_P * conv_row_ld(a.d) + (size_t)conv->cin * conv->h * conv->w) * sizeof(float);
    const size_t need = sizeof(float) * (size_t)a.n * a.d;
    if (need > S.r_cap) {
      if (S.r_buf) LASSO_CUDA_TRY(cudaFree(S.r_buf));
      S.r_buf = nullptr;
      S.r_cap = 0;
      LASSO_CUDA_TRY(cudaMalloc(&S.r_buf, need));
      S.r_cap = need;
    }
    if (!S.conv_attr_set) {
      LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)conv_resid_kernel<true>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)conv_resid_kernel<false>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      S.conv_attr_set = true;
    }
    conv_scale_kernel<<<(unsigned)conv->n_img, 256, 0, st>>>(a.x, conv->cin * conv->h * conv->w, conv_P, a.k, S.scal,
                                                             S.row_scale, a.z_a, a.zero_start ? 0 : 1);
  } else {
    blk_rowscale_kernel<<<(unsigned)((a.n + 7) / 8), 256, 0, st>>>(a.x, a.n, a.d, a.k, S.scal, S.row_scale, a.z_a,
                                                                   a.zero_start ? 0 : 1);
  }
  LASSO_CHECK_LAUNCH();
  count_launch(3);

  CUtensorMap tm_a, tm_b;
  int rc;
  // LASSO_B200_BLK_ROWS=64: half-height tiles (measured slower at C3, 1.47 vs 1.11 ms: the
  // dictionary slices are re-streamed from L2 per tile, and that traffic doubles)
  int trows = kTileM;
  if (const char* t = getenv("LASSO_B200_BLK_ROWS")) trows = atoi(t) == 64 ? 64 : kTileM;
  // LASSO_B200_BLK_L2=1: evict-last on pass 1's z_cur, evict-first elsewhere.  Measured at C3: no
  // gain (1.070 vs 1.056 ms) -- 148 tiles x 1 MB in flight exceed the L2 either way
  int l2_hints = 0;
  if (const char* t = getenv("LASSO_B200_BLK_L2")) l2_hints = atoi(t) != 0;
  if ((rc = blk_make_map(&tm_a, a.z_a, a.n, a.k, trows))) return rc;
  if ((rc = blk_make_map(&tm_b, a.z_b, a.n, a.k, trows))) return rc;
  const int64_t ntiles = (a.n + trows - 1) / trows;
  const unsigned grid = (unsigned)(ntiles < S.num_sms ? ntiles : S.num_sms);
  double t = 1.0;
  for (int it = 0; it < a.maxiter; ++it) {
    BlkParams p{};
    p.w_image = S.w_image;
    p.x = a.x;
    p.row_scale = S.row_scale;
    p.z_cur = (it & 1) ? a.z_b : a.z_a;
    p.z_io = (it & 1) ? a.z_a : a.z_b;
    p.n = a.n;
    p.d = a.d;
    p.k = a.k;
    double beta = 0.0;
    if (a.fast && it > 0) {
      const double t_next = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;
      beta = (t - 1.0) / t_next;
      t = t_next;
    }
    p.beta = (float)beta;
    p.use_prev = it > 0 ? 1 : 0;
    p.trows = trows;
    p.l2_hints = l2_hints;
    p.scal = S.scal;
    p.ctl.hist = a.hist;
    p.ctl.tol_abs = a.tol_abs;
    p.ctl.iter = it;
    p.dbg = S.dbg_dev;
    p.r_buf = S.r_buf;
    for (int half = 0; half < (conv ? 2 : 1); ++half) {
      p.mode = conv ? half + 1 : 0;
      if (it & 1) fista_blk_kernel<<<grid, kThreadsB, kSmemBytesB, st>>>(tm_b, tm_a, p);
      else fista_blk_kernel<<<grid, kThreadsB, kSmemBytesB, st>>>(tm_a, tm_b, p);
      LASSO_CHECK_LAUNCH();
      count_launch();
      if (conv && half == 0) {
        if (conv->stride == 1)
          conv_resid_kernel<true><<<(unsigned)conv->n_img, (1024 / a.d) * a.d, conv_smem, st>>>(S.r_buf, a.x, S.row_scale,
                                                                                               *conv, p.ctl);
        else
          conv_resid_kernel<false><<<(unsigned)conv->n_img, (1024 / a.d) * a.d, conv_smem, st>>>(S.r_buf, a.x, S.row_scale,
                                                                                                *conv, p.ctl);
        LASSO_CHECK_LAUNCH();
        count_launch();
      }
    }
  }
  // back to the caller's units: the buffer that holds z_maxiter, and the other one too when the stop
  // test may select it
  float limit = 32768.0f;
  if (const char* lim = getenv("LASSO_B200_RES_LIMIT")) limit = (float)atof(lim);   // tests: force the fallback
  const int blocks = (int)std::min<int64_t>((a.n + 7) / 8, (int64_t)S.num_sms * 16);
  float* z_final = (a.maxiter & 1) ? a.z_b : a.z_a;
  float* z_other = (a.maxiter & 1) ? a.z_a : a.z_b;
  blk_unscale_kernel<<<blocks, 256, 0, st>>>(z_final, a.n, a.k, S.row_scale, S.scal, limit, S.flag);
  LASSO_CHECK_LAUNCH();
  count_launch();
  if (a.maxiter > 0 && a.tol_abs >= 0.0) {
    blk_unscale_kernel<<<blocks, 256, 0, st>>>(z_other, a.n, a.k, S.row_scale, S.scal, limit, S.flag);
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  int flag = 0;
  cudaError_t e = cudaMemcpyAsync(&flag, S.flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (S.dbg_host && S.dbg_host[0]) {
    set_error("blocked tcgen05 kernel barrier timeout: line %d block %d thread %d iter %d parity %d (%s)",
              S.dbg_host[1], S.dbg_host[2], S.dbg_host[3], S.dbg_host[4], S.dbg_host[5], cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  if (e != cudaSuccess) {
    set_error("blocked tcgen05 kernel failed: %s", cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  *fell_back = flag;
  return LASSO_B200_OK;
}

}  // namespace lasso
