// tcgen05 / TMEM FISTA step for sm_100a: one persistent launch per iteration.
//
// Shapes: d <= 64, k <= 256, d % 4 == 0, k % 4 == 0 (TMA row pitch); other shapes take
// the FFMA path.  Per 128-row tile (one CTA per SM, grid-stride over tiles):
//
//   phase A  for each 32-atom chunk: TMA-staged z_cur / z_prev -> y = z_cur + beta (z_cur -
//            z_prev) (exact fp32, ista.py:100) -> y kept in TMEM (fp32 master) and split
//            into three bf16 pieces (y = p1 + p2 + p3 exactly) stored to TMEM as the A operand.
//            The lagged stop-test sum |z_prev - z_cur| (ista.py:93) is accumulated here.
//   GEMM1    R = Y W^T as six bf16 tcgen05.mma products per k-step, A from TMEM, B = the
//            dictionary pieces resident in shared memory (K-major view).  The leading product
//            p1*q1 accumulates in its own TMEM accumulator, the five small ones in a second
//            one: the tensor core truncates (RZ) on every accumulate, and that bias scales
//            with the accumulator magnitude (probe E4, tools/tc_probe.cu).
//   phase B  r = (R_big + R_small) - x, split into bf16 pieces -> TMEM (A of GEMM2)
//   GEMM2    G = r W per 64-atom chunk, B = the SAME shared-memory image read through an
//            MN-major descriptor; small products first, leading product last.
//   phase C  z_next = softshrink(y - lr g, alpha lr) (ista.py:90) from TMEM (g, y), stored
//            over z_prev in HBM.
//
// HBM traffic per iteration: n (d + 3k) floats.  TMEM columns: y 256 | piece stages / r
// pieces 96 | R_big / G buffer 0: 64 | R_small / G buffer 1: 64  = 480 of 512.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace lasso {
namespace {

using namespace sm100;

constexpr int kTileM = 128;
constexpr int kChunk = 32;       // atoms per phase-A chunk (one 128-byte fp32 row)
constexpr int kQ = 64;           // atoms per GEMM2 chunk (one 128-byte bf16 row)
constexpr int kDP = 64;          // padded d
constexpr int kKP = 256;         // padded k
constexpr int kStages = 4;       // two per compute group (fixed ownership, see phase A)
constexpr int kThreads = 320;    // warp 0: TMA, warp 1: MMA, warps 2..9: compute

constexpr uint32_t kSlabBytes = kDP * 128;               // [64 rows][128 B] = 64 atoms of one piece
constexpr uint32_t kPieceBytes = (kKP / 64) * kSlabBytes;  // 32 KB
constexpr uint32_t kWBytes = 3 * kPieceBytes;            // 96 KB
constexpr uint32_t kBoxBytes = kTileM * 128;             // [128 rows][128 B] = 16 KB
constexpr uint32_t kStageBytes = 2 * kBoxBytes;          // z_cur + z_prev chunk
constexpr uint32_t kSmemW = 0;
constexpr uint32_t kSmemStage = kSmemW + kWBytes;
constexpr uint32_t kSmemBytes = kSmemStage + kStages * kStageBytes;  // 229376

// TMEM column map
constexpr uint32_t kColY = 0;
constexpr uint32_t kColStage = 256;   // 2 stages x 48 columns; r pieces alias [256, 352)
constexpr uint32_t kColAcc0 = 352;    // R_big   / G buffer 0
constexpr uint32_t kColAcc1 = 416;    // R_small / G buffer 1
constexpr uint32_t kTmemCols = 512;

struct TcParams {
  const uint8_t* w_image;  // global, kWBytes
  const float* x;          // [n][d]
  float* z_io;             // z_prev on entry, z_next on exit
  int64_t n;
  int d, k;
  float lr, lam, beta;
  int use_prev;
  int cur_is_a;            // which tensor map holds z_cur
  StepCtl ctl;
  volatile int* dbg;       // host-mapped debug record or nullptr
};

// A barrier that does not complete is a protocol bug: record where (host-mapped debug
// record, if armed with LASSO_B200_DEBUG=1) and trap instead of hanging the GPU.
#define TC_WAIT(bar, parity)                                              \
  do {                                                                    \
    if (p.dbg) p.dbg[64 + blockIdx.x * 16 + (threadIdx.x >> 5)] = __LINE__ * 16 + (int)(parity); \
    if (!mbar_wait((bar), (parity))) {                                    \
      if (p.dbg) {                                                        \
        p.dbg[64 + blockIdx.x * 16 + (threadIdx.x >> 5)] = -(__LINE__ * 16 + (int)(parity)); \
        p.dbg[1] = __LINE__; p.dbg[2] = blockIdx.x; p.dbg[3] = threadIdx.x; \
        p.dbg[4] = p.ctl.iter; p.dbg[5] = (int)(parity);                  \
        __threadfence_system();                                           \
        p.dbg[0] = 1;                                                     \
        __threadfence_system();                                           \
      }                                                                   \
      __trap();                                                           \
    }                                                                     \
    if (p.dbg) p.dbg[64 + blockIdx.x * 16 + (threadIdx.x >> 5)] = 0;      \
  } while (0)

// exact three-way bf16 split of an fp32 value: v = p1 + p2 + p3 (upper 16 bits each)
__device__ __forceinline__ void split3(float v, uint32_t& p1, uint32_t& p2, uint32_t& p3) {
  p1 = __float_as_uint(v) & 0xFFFF0000u;
  const float r1 = __fsub_rn(v, __uint_as_float(p1));
  p2 = __float_as_uint(r1) & 0xFFFF0000u;
  const float r2 = __fsub_rn(r1, __uint_as_float(p2));
  p3 = __float_as_uint(r2);
}
// two bf16 (upper halves of a, b) -> one 32-bit word, element a in the low half
__device__ __forceinline__ uint32_t pack_hi(uint32_t a, uint32_t b) {
  return __byte_perm(a, b, 0x7632);
}

__global__ void __launch_bounds__(kThreads, 1)
fista_tc_kernel(const __grid_constant__ CUtensorMap tm_za, const __grid_constant__ CUtensorMap tm_zb,
                TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_full[kStages], bar_empty[kStages];
  __shared__ uint64_t bar_aready[2], bar_sfree[2], bar_rfull, bar_rready, bar_gfull[2], bar_gfree[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ double red[8];

  // stop test of two iterations ago already satisfied -> this launch is a no-op.
  // (hist[iter-1] is produced by THIS launch, see phase A.)
  if (p.ctl.tol_abs >= 0.0 && p.ctl.iter >= 2 && p.ctl.hist[p.ctl.iter - 2] <= p.ctl.tol_abs) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (p.n + kTileM - 1) / kTileM;
  const int nc = (p.k + kChunk - 1) / kChunk;   // phase-A chunks
  const int nq = (p.k + kQ - 1) / kQ;           // GEMM2 chunks
  const int dsteps = (p.d + 15) / 16;           // k-steps of GEMM2
  const CUtensorMap* tm_cur = p.cur_is_a ? &tm_za : &tm_zb;
  const CUtensorMap* tm_prev = p.cur_is_a ? &tm_zb : &tm_za;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 128);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_aready[b], 128);
      mbar_init(&bar_sfree[b], 1);
      mbar_init(&bar_gfull[b], 1);
      mbar_init(&bar_gfree[b], 128);
    }
    mbar_init(&bar_rfull, 1);
    mbar_init(&bar_rready, 256);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      prefetch_tmap(&tm_za);
      prefetch_tmap(&tm_zb);
      mbar_expect_tx(&bar_w, kWBytes);
      for (uint32_t off = 0; off < kWBytes; off += 16384)
        bulk_load(smem + kSmemW + off, p.w_image + off, 16384, &bar_w);
      // Chunk c belongs to compute group c & 1, and each group owns its own two-stage ring
      // (stages g and g + 2).  One consumer group per barrier keeps every waiter within one
      // phase of the barrier, which is all a parity wait can disambiguate.
      uint32_t m[2] = {0, 0};
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = (int)(tile * kTileM);
        for (int c = 0; c < nc; ++c) {
          const int g = c & 1;
          const uint32_t s = g + 2 * (m[g] & 1), ph = (m[g] >> 1) & 1;
          ++m[g];
          TC_WAIT(&bar_empty[s], ph ^ 1);
          mbar_expect_tx(&bar_full[s], kStageBytes);
          uint8_t* dst = smem + kSmemStage + s * kStageBytes;
          tma_load_2d(dst, tm_cur, c * kChunk, row0, &bar_full[s]);
          tma_load_2d(dst + kBoxBytes, tm_prev, c * kChunk, row0, &bar_full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(kFmtBF16, 128, kDP, 0, 0);   // B K-major  (GEMM1)
      const uint32_t idesc2 = make_idesc(kFmtBF16, 128, kQ, 0, 1);    // B MN-major (GEMM2)
      const uint32_t w_addr = smem_u32(smem + kSmemW);
      TC_WAIT(&bar_w, 0);
      uint32_t a_cnt[2] = {0, 0}, g_cnt[2] = {0, 0}, ti = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        // accumulators alias the G buffers of the previous tile: wait until both were drained
        TC_WAIT(&bar_gfree[0], (g_cnt[0] & 1) ^ 1);
        TC_WAIT(&bar_gfree[1], (g_cnt[1] & 1) ^ 1);
        tc_fence_after();
        // ---- GEMM1: R[128 x 64] = Y[128 x k] * W^T ----
        for (int c = 0; c < nc; ++c) {
          const int b = c & 1;
          TC_WAIT(&bar_aready[b], a_cnt[b] & 1);
          ++a_cnt[b];
          tc_fence_after();
          const uint32_t t_stage = tbase + kColStage + b * 48;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t koff = (uint32_t)(c >> 1) * kSlabBytes + (uint32_t)((c & 1) * 32 + ks * 16) * 2;
            const uint32_t acc_on = (c > 0 || ks > 0) ? 1u : 0u;
            auto bdesc = [&](int piece) {
              return make_smem_desc_sw128(w_addr + piece * kPieceBytes + koff, 0, 1024);
            };
            auto aaddr = [&](int piece) { return t_stage + piece * 16 + ks * 8; };
            // small products (p1 q2, p2 q1, p1 q3, p2 q2, p3 q1) -> R_small
            mma_ts<false>(tbase + kColAcc1, aaddr(0), bdesc(2), idesc1, acc_on);
            mma_ts<false>(tbase + kColAcc1, aaddr(1), bdesc(1), idesc1, 1);
            mma_ts<false>(tbase + kColAcc1, aaddr(2), bdesc(0), idesc1, 1);
            mma_ts<false>(tbase + kColAcc1, aaddr(0), bdesc(1), idesc1, 1);
            mma_ts<false>(tbase + kColAcc1, aaddr(1), bdesc(0), idesc1, 1);
            // leading product -> R_big
            mma_ts<false>(tbase + kColAcc0, aaddr(0), bdesc(0), idesc1, acc_on);
          }
          mma_commit(&bar_sfree[b]);
        }
        mma_commit(&bar_rfull);
        // ---- GEMM2: G[128 x 64q] = r[128 x d] * W[:, chunk] ----
        TC_WAIT(&bar_rready, ti & 1);
        tc_fence_after();
        for (int q = 0; q < nq; ++q) {
          const int b = q & 1;
          TC_WAIT(&bar_gfree[b], (g_cnt[b] & 1) ^ 1);
          tc_fence_after();
          const uint32_t t_acc = tbase + (b ? kColAcc1 : kColAcc0);
          auto bdesc = [&](int piece, int ks) {
            return make_smem_desc_sw128(w_addr + piece * kPieceBytes + q * kSlabBytes + ks * 2048,
                                        kSlabBytes, 1024);
          };
          auto aaddr = [&](int piece, int ks) { return tbase + kColStage + piece * 32 + ks * 8; };
          uint32_t acc_on = 0;
          // small products first (their truncation error scales with a small accumulator)
          const int pa[5] = {2, 1, 0, 1, 0}, pb[5] = {0, 1, 2, 0, 1};
#pragma unroll
          for (int t = 0; t < 5; ++t)
            for (int ks = 0; ks < dsteps; ++ks) {
              mma_ts<false>(t_acc, aaddr(pa[t], ks), bdesc(pb[t], ks), idesc2, acc_on);
              acc_on = 1;
            }
          for (int ks = 0; ks < dsteps; ++ks) mma_ts<false>(t_acc, aaddr(0, ks), bdesc(0, ks), idesc2, 1);
          mma_commit(&bar_gfull[b]);
          ++g_cnt[b];
        }
      }
    }
  } else {
    // ===================== compute warps =====================
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int grp = (warp - 2) >> 2;           // 0 / 1: which chunks / G buffer
    const int row = quad * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    double dsum = 0.0;
    uint32_t a_cnt = 0, g_cnt = 0, ti = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const int64_t grow = tile * kTileM + row;
      // ---------------- phase A ----------------
      for (int c = grp; c < nc; c += 2) {
        // a_cnt = chunks this group consumed so far = index into its private stage ring
        const uint32_t s = grp + 2 * (a_cnt & 1), ph = (a_cnt >> 1) & 1;
        TC_WAIT(&bar_full[s], ph);
        const uint8_t* zc_s = smem + kSmemStage + s * kStageBytes;
        const uint8_t* zp_s = zc_s + kBoxBytes;
        float y[kChunk];
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t off = sw128_offset(row, j * 16);
          const float4 zc = *reinterpret_cast<const float4*>(zc_s + off);
          if (p.use_prev) {
            const float4 zp = *reinterpret_cast<const float4*>(zp_s + off);
            y[4 * j + 0] = momentum_point(zc.x, zp.x, p.beta);
            y[4 * j + 1] = momentum_point(zc.y, zp.y, p.beta);
            y[4 * j + 2] = momentum_point(zc.z, zp.z, p.beta);
            y[4 * j + 3] = momentum_point(zc.w, zp.w, p.beta);
            part += fabsf(__fsub_rn(zp.x, zc.x)) + fabsf(__fsub_rn(zp.y, zc.y)) +
                    fabsf(__fsub_rn(zp.z, zc.z)) + fabsf(__fsub_rn(zp.w, zc.w));
          } else {
            y[4 * j + 0] = zc.x; y[4 * j + 1] = zc.y; y[4 * j + 2] = zc.z; y[4 * j + 3] = zc.w;
          }
        }
        mbar_arrive(&bar_empty[s]);   // stage consumed (values are in registers)
        dsum += (double)part;
        uint32_t yb[kChunk], w1[16], w2[16], w3[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          uint32_t a1, a2, a3, b1, b2, b3;
          split3(y[2 * j], a1, a2, a3);
          split3(y[2 * j + 1], b1, b2, b3);
          w1[j] = pack_hi(a1, b1);
          w2[j] = pack_hi(a2, b2);
          w3[j] = pack_hi(a3, b3);
          yb[2 * j] = __float_as_uint(y[2 * j]);
          yb[2 * j + 1] = __float_as_uint(y[2 * j + 1]);
        }
        // the piece stage is free once the MMAs of its previous chunk completed
        TC_WAIT(&bar_sfree[grp], (a_cnt & 1) ^ 1);
        ++a_cnt;
        tc_fence_after();
        tmem_st32(tbase + lane_base + kColY + c * kChunk, yb);
        const uint32_t t_stage = tbase + lane_base + kColStage + grp * 48;
        tmem_st16(t_stage, w1);
        tmem_st16(t_stage + 16, w2);
        tmem_st16(t_stage + 32, w3);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_aready[grp]);
      }
      // ---------------- phase B: r = R - x, pieces of r ----------------
      {
        // this thread's 32 features of x, straight from global (issued before the wait on
        // GEMM1 so the latency overlaps the tail of the MMAs)
        float4 xv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int col = grp * 32 + 4 * j;
          xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (grow < p.n && col < p.d)
            xv[j] = __ldg(reinterpret_cast<const float4*>(p.x + grow * p.d + col));
        }
        TC_WAIT(&bar_rfull, ti & 1);
        tc_fence_after();
        uint32_t rb[32], rs[32];
        tmem_ld32(tbase + lane_base + kColAcc0 + grp * 32, rb);
        tmem_ld32(tbase + lane_base + kColAcc1 + grp * 32, rs);
        tmem_wait_ld();
        uint32_t w1[16], w2[16], w3[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xr[4] = {xv[j].x, xv[j].y, xv[j].z, xv[j].w};
          float r[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            r[e] = __fsub_rn(__fadd_rn(__uint_as_float(rb[4 * j + e]), __uint_as_float(rs[4 * j + e])),
                             xr[e]);
          uint32_t a1, a2, a3, b1, b2, b3;
          split3(r[0], a1, a2, a3);
          split3(r[1], b1, b2, b3);
          w1[2 * j] = pack_hi(a1, b1); w2[2 * j] = pack_hi(a2, b2); w3[2 * j] = pack_hi(a3, b3);
          split3(r[2], a1, a2, a3);
          split3(r[3], b1, b2, b3);
          w1[2 * j + 1] = pack_hi(a1, b1); w2[2 * j + 1] = pack_hi(a2, b2); w3[2 * j + 1] = pack_hi(a3, b3);
        }
        const uint32_t t_r = tbase + lane_base + kColStage + grp * 16;
        tmem_st16(t_r, w1);
        tmem_st16(t_r + 32, w2);
        tmem_st16(t_r + 64, w3);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_rready);
      }
      // ---------------- phase C: fused update ----------------
      for (int q = grp; q < nq; q += 2) {
        TC_WAIT(&bar_gfull[grp], g_cnt & 1);
        ++g_cnt;
        tc_fence_after();
        const uint32_t t_g = tbase + lane_base + (grp ? kColAcc1 : kColAcc0);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t g[32], yv[32];
          tmem_ld32(t_g + h * 32, g);
          tmem_ld32(tbase + lane_base + kColY + q * kQ + h * 32, yv);
          tmem_wait_ld();
          const int col0 = q * kQ + h * 32;
          if (grow < p.n) {
            float* dst = p.z_io + grow * p.k + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + 4 * j < p.k) {
                float4 o;
                o.x = ista_update(__uint_as_float(yv[4 * j + 0]), __uint_as_float(g[4 * j + 0]), p.lr, p.lam);
                o.y = ista_update(__uint_as_float(yv[4 * j + 1]), __uint_as_float(g[4 * j + 1]), p.lr, p.lam);
                o.z = ista_update(__uint_as_float(yv[4 * j + 2]), __uint_as_float(g[4 * j + 2]), p.lr, p.lam);
                o.w = ista_update(__uint_as_float(yv[4 * j + 3]), __uint_as_float(g[4 * j + 3]), p.lr, p.lam);
                *reinterpret_cast<float4*>(dst + 4 * j) = o;
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bar_gfree[grp]);
      }
      // y master / r pieces are rewritten by the next tile: all compute warps must be done
      asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    // lagged stop-test sum of the previous iteration
    dsum = warp_sum(dsum);
    if (lane == 0) red[warp - 2] = dsum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 2 && lane == 0 && p.use_prev && p.ctl.hist != nullptr && p.ctl.iter >= 1) {
      double s = 0.0;
      for (int i = 0; i < 8; ++i) s += red[i];
      atomicAdd(&p.ctl.hist[p.ctl.iter - 1], s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, kTmemCols);
}

// dictionary [d][k] fp32 -> three bf16 piece images, each [k/64 slabs][64 rows][128 B] with
// the 128-byte swizzle; zero padded to d = 64, k = 256.
__global__ void prep_w_image_kernel(const float* __restrict__ w, int d, int k,
                                    uint8_t* __restrict__ image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per (i, j)
  if (idx >= kDP * kKP) return;
  const int i = idx / kKP, j = idx % kKP;
  const float v = (i < d && j < k) ? w[(int64_t)i * k + j] : 0.f;
  uint32_t p1, p2, p3;
  split3(v, p1, p2, p3);
  const uint32_t off = (uint32_t)(j / 64) * kSlabBytes + sw128_offset(i, (j % 64) * 2);
  *reinterpret_cast<uint16_t*>(image + 0 * kPieceBytes + off) = (uint16_t)(p1 >> 16);
  *reinterpret_cast<uint16_t*>(image + 1 * kPieceBytes + off) = (uint16_t)(p2 >> 16);
  *reinterpret_cast<uint16_t*>(image + 2 * kPieceBytes + off) = (uint16_t)(p3 >> 16);
}

// ---- host side ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// 2-D fp32 tensor [rows][cols] row-major, box = [128 rows][32 cols], 128-byte swizzle
int make_map(CUtensorMap* map, const float* base, int64_t rows, int cols) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return LASSO_B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)kTileM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%d)", (int)r,
              (long long)rows, cols);
    return LASSO_B200_ERR_CUDA;
  }
  return LASSO_B200_OK;
}

struct TcState {
  uint8_t* w_image = nullptr;
  int num_sms = 0;
  bool attr_set = false;
  int* dbg_host = nullptr;   // LASSO_B200_DEBUG=1: mapped record of the first barrier timeout
  int* dbg_dev = nullptr;
};
TcState g_tc[64];

}  // namespace

bool fista_tc_supported(int64_t n, int d, int k) {
  return n >= 1 && d >= 4 && d <= kDP && k >= 4 && k <= kKP && (d % 4) == 0 && (k % 4) == 0 &&
         n < (int64_t)1 << 31;
}

int fista_tc_run(const FistaArgs& a, float* /*z_out*/, cudaStream_t st) {
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  TcState& S = g_tc[dev];
  if (!S.w_image) {
    LASSO_CUDA_TRY(cudaMalloc(&S.w_image, kWBytes));
    LASSO_CUDA_TRY(cudaDeviceGetAttribute(&S.num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!S.dbg_host && getenv("LASSO_B200_DEBUG")) {
    LASSO_CUDA_TRY(cudaHostAlloc((void**)&S.dbg_host, 65536, cudaHostAllocMapped));
    memset(S.dbg_host, 0, 65536);
    LASSO_CUDA_TRY(cudaHostGetDevicePointer((void**)&S.dbg_dev, S.dbg_host, 0));
  }
  if (!S.attr_set) {
    LASSO_CUDA_TRY(cudaFuncSetAttribute(fista_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)kSmemBytes));
    S.attr_set = true;
  }
  prep_w_image_kernel<<<(kDP * kKP + 255) / 256, 256, 0, st>>>(a.w, a.d, a.k, S.w_image);
  LASSO_CHECK_LAUNCH();
  count_launch();

  CUtensorMap tm_za, tm_zb;
  int rc;
  if ((rc = make_map(&tm_za, a.z_a, a.n, a.k))) return rc;
  if ((rc = make_map(&tm_zb, a.z_b, a.n, a.k))) return rc;

  const int64_t ntiles = (a.n + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(ntiles < S.num_sms ? ntiles : S.num_sms);
  double t = 1.0;
  for (int it = 0; it < a.maxiter; ++it) {
    TcParams p{};
    p.w_image = S.w_image;
    p.x = a.x;
    p.z_io = (it & 1) ? a.z_a : a.z_b;
    p.cur_is_a = (it & 1) ? 0 : 1;
    p.n = a.n;
    p.d = a.d;
    p.k = a.k;
    p.lr = a.lr;
    p.lam = a.lam;
    double beta = 0.0;
    if (a.fast && it > 0) {
      const double t_next = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;
      beta = (t - 1.0) / t_next;
      t = t_next;
    }
    p.beta = (float)beta;
    // z_prev is always read from iteration 1 on: it feeds the lagged stop-test sum even
    // for plain ISTA (beta = 0 leaves y = z_cur + 0 * (z_cur - z_prev) = z_cur exactly)
    p.use_prev = it > 0 ? 1 : 0;
    p.ctl.hist = a.hist;
    p.ctl.tol_abs = a.tol_abs;
    p.ctl.iter = it;
    p.dbg = S.dbg_dev;
    fista_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(tm_za, tm_zb, p);
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  if (S.dbg_host) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (S.dbg_host[0]) {
      for (int b = 0; b < 148; ++b) {
        bool any = false;
        for (int wi = 0; wi < 10; ++wi) any |= S.dbg_host[64 + b * 16 + wi] != 0;
        if (!any) continue;
        fprintf(stderr, "[lasso_b200 dbg] block %3d:", b);
        for (int wi = 0; wi < 10; ++wi) {
          const int v = S.dbg_host[64 + b * 16 + wi];
          const int a = v < 0 ? -v : v;
          fprintf(stderr, " w%d=%s%d/%d", wi, v < 0 ? "T" : "", a / 16, a % 16);
        }
        fprintf(stderr, "\n");
      }
      set_error("tcgen05 kernel barrier timeout: line %d block %d thread %d iter %d parity %d (%s)",
                S.dbg_host[1], S.dbg_host[2], S.dbg_host[3], S.dbg_host[4], S.dbg_host[5],
                cudaGetErrorString(e));
      return LASSO_B200_ERR_CUDA;
    }
    if (e != cudaSuccess) {
      set_error("tcgen05 kernel failed: %s", cudaGetErrorString(e));
      return LASSO_B200_ERR_CUDA;
    }
  }
  return LASSO_B200_OK;
}

}  // namespace lasso
