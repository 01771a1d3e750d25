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
constexpr int kThreads = 576;    // warp 0: TMA, warp 1: MMA, warps 2..17: compute (4 groups)

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
  int tile_rows;           // rows per tile (<= 128): chosen so that the last wave of tiles is full
  float lr, lam, beta;
  int use_prev;
  int cur_is_a;            // which tensor map holds z_cur
  // v1 schedule: CTAs of class (blockIdx & 1) own contiguous row ranges cut into `waves` tiles of
  // cls_rows[class] rows; class 1 starts stagger_cycles late and gets shorter tiles, so that half
  // of the SMs stream codes (HBM-bound phase A) while the other half is in the MMA-bound phases.
  int cls_rows[2];
  int64_t cls_base[2], cls_end[2];
  int waves;
  int stagger_cycles;
  // L2 eviction priorities (TMA cache hints) of the two code buffers, see fista_tc_run
  uint64_t pol_cur, pol_prev;
  StepCtl ctl;
  volatile int* dbg;       // host-mapped debug record or nullptr
  unsigned long long* trace;  // LASSO_B200_TRACE: per-warp (clock << 8 | event) log of block 0
};

// A barrier that does not complete is a protocol bug: after ~2 s record where (host-mapped
// debug record, armed with LASSO_B200_DEBUG=1) and trap instead of hanging the GPU.
__device__ __noinline__ void tc_wait_slow_path(uint64_t& t0, volatile int* dbg, int line, int iter,
                                               uint32_t parity) {
  const uint64_t now = global_timer_ns();
  if (t0 == 0) {
    t0 = now;
    return;
  }
  if (now - t0 < 2000000000ull) return;
  if (dbg) {
    dbg[1] = line; dbg[2] = blockIdx.x; dbg[3] = threadIdx.x; dbg[4] = iter; dbg[5] = (int)parity;
    __threadfence_system();
    dbg[0] = 1;
    __threadfence_system();
  }
  __trap();
}
// The polling loop is spelled out in the macro (not an inline function) so that profiler
// samples of a wait land on the line of the wait, i.e. tell WHICH barrier a warp sat on.
#define TC_WAIT(bar, parity)                                                              \
  do {                                                                                    \
    const uint32_t _addr = smem_u32(bar), _par = (parity);                                \
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
      if ((++_n & 1023u) == 0) tc_wait_slow_path(_t0, p.dbg, __LINE__, p.ctl.iter, _par); \
    }                                                                                     \
  } while (0)

// timeline instrumentation (block 0, lane 0 of every warp), enabled with LASSO_B200_TRACE=<file>
#define TRACE(id)                                                                         \
  do {                                                                                    \
    if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && tr_n < 510)   \
      p.trace[(threadIdx.x >> 5) * 512 + tr_n++] = ((unsigned long long)clock64() << 8) | (id); \
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

// named barrier of one pair of compute groups (8 warps); ids 2 and 3
// (0 = __syncthreads, 1 = all compute warps)
__device__ __forceinline__ void pair_sync(int grp) {
  asm volatile("bar.sync %0, 256;" ::"r"(2 + grp) : "memory");
}

// ---- packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): one issue slot per pair.
// a - b is formed as fma(-1, b, a): a single rounding of the exact difference, i.e. the
// same bits as __fsub_rn, so the reference's rounding sequence is preserved.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  return __ffma2_rn(make_float2(-1.f, -1.f), b, a);
}
// pair version of split3: packed words of the three bf16 pieces of (v.x, v.y)
__device__ __forceinline__ void split3_pair(float2 v, uint32_t& w1, uint32_t& w2, uint32_t& w3) {
  const float2 t1 = make_float2(__uint_as_float(__float_as_uint(v.x) & 0xFFFF0000u),
                                __uint_as_float(__float_as_uint(v.y) & 0xFFFF0000u));
  const float2 r1 = sub2(v, t1);
  const float2 t2 = make_float2(__uint_as_float(__float_as_uint(r1.x) & 0xFFFF0000u),
                                __uint_as_float(__float_as_uint(r1.y) & 0xFFFF0000u));
  const float2 r2 = sub2(r1, t2);
  w1 = pack_hi(__float_as_uint(t1.x), __float_as_uint(t1.y));
  w2 = pack_hi(__float_as_uint(t2.x), __float_as_uint(t2.y));
  w3 = pack_hi(__float_as_uint(r2.x), __float_as_uint(r2.y));
}
// softshrink(y - lr*g, lam) on a pair: v - clamp(v, -lam, lam) is bit-identical to ATen's
// three-way select for finite v (v - lam, v + lam or +0 with one rounding each).
__device__ __forceinline__ float2 ista_update_pair(float2 y, float2 g, float2 lr2, float lam) {
  const float2 v = sub2(y, __fmul2_rn(lr2, g));
  const float2 c = make_float2(fminf(fmaxf(v.x, -lam), lam), fminf(fmaxf(v.y, -lam), lam));
  return sub2(v, c);
}

template <int kDSteps>
__global__ void __launch_bounds__(kThreads, 1)
fista_tc_kernel(const __grid_constant__ CUtensorMap tm_za, const __grid_constant__ CUtensorMap tm_zb,
                const __grid_constant__ CUtensorMap tm_za1, const __grid_constant__ CUtensorMap tm_zb1,
                TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_full[kStages], bar_empty[kStages];
  __shared__ uint64_t bar_aready[2], bar_sfree[2], bar_rfull, bar_rready, bar_gfull[2], bar_gfree[2];
  __shared__ uint64_t bar_tdone[2];   // pair p has finished phase C of the current tile
  __shared__ uint32_t tmem_base_s;
  __shared__ double red[16];

  // stop test of two iterations ago already satisfied -> this launch is a no-op.
  // (hist[iter-1] is produced by THIS launch, see phase A.)
  if (p.ctl.tol_abs >= 0.0 && p.ctl.iter >= 2 && p.ctl.hist[p.ctl.iter - 2] <= p.ctl.tol_abs) return;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int tr_n = 0;
  // this CTA's rows: `my_tiles` tiles of `trows` rows starting at cta_row0 (see TcParams)
  // (tiles of a class are interleaved over its CTAs: tile j of this CTA is class tile
  //  (blockIdx >> 1) + j * cls_ctas; all row indices fit 32 bits, see fista_tc_supported)
  const int cls = blockIdx.x & 1;
  const int trows = p.cls_rows[cls];
  const int cls_ctas = ((int)gridDim.x + 1 - cls) >> 1;
  const int row_limit = (int)p.cls_end[cls];
  const int first_row = (int)p.cls_base[cls] + (int)(blockIdx.x >> 1) * trows;
  const int row_step = cls_ctas * trows;
  const int my_tiles = first_row < row_limit ? (row_limit - first_row + row_step - 1) / row_step : 0;
  const int nc = (p.k + kChunk - 1) / kChunk;   // phase-A chunks
  const int nq = (p.k + kQ - 1) / kQ;           // GEMM2 chunks
  constexpr int dsteps = kDSteps;               // k-steps of GEMM2 = ceil(d / 16), compile time so
                                                // that every MMA descriptor is base + constant
  const CUtensorMap* tm_cur = cls ? (p.cur_is_a ? &tm_za1 : &tm_zb1) : (p.cur_is_a ? &tm_za : &tm_zb);
  const CUtensorMap* tm_prev = cls ? (p.cur_is_a ? &tm_zb1 : &tm_za1) : (p.cur_is_a ? &tm_zb : &tm_za);

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bar_full[s], 1);
      mbar_init(&bar_empty[s], 256);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bar_aready[b], 256);
      mbar_init(&bar_sfree[b], 1);
      mbar_init(&bar_gfull[b], 1);
      mbar_init(&bar_gfree[b], 256);
    }
    mbar_init(&bar_rfull, 1);
    mbar_init(&bar_rready, 512);
    mbar_init(&bar_tdone[0], 256);
    mbar_init(&bar_tdone[1], 256);
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
    if (elect_one()) {
      prefetch_tmap(tm_cur);
      prefetch_tmap(tm_prev);
      mbar_expect_tx(&bar_w, kWBytes);
      for (uint32_t off = 0; off < kWBytes; off += 16384)
        bulk_load(smem + kSmemW + off, p.w_image + off, 16384, &bar_w);
    }
    __syncwarp();
    if (cls == 1 && p.stagger_cycles > 0) {
      const long long t0 = clock64();
      while (clock64() - t0 < p.stagger_cycles) __nanosleep(200);
    }
    // Chunk c belongs to compute group c & 1, and each group owns its own two-stage ring
    // (stages g and g + 2).  One consumer group per barrier keeps every waiter within one
    // phase of the barrier, which is all a parity wait can disambiguate.
    uint32_t m0 = 0, m1 = 0;
    for (int tile = 0; tile < my_tiles; ++tile) {
      const int row0 = first_row + tile * row_step;
      for (int c = 0; c < nc; ++c) {
        const int g = c & 1;
        const uint32_t mg = g ? m1 : m0;
        const uint32_t s = g + 2 * (mg & 1), ph = (mg >> 1) & 1;
        if (g) ++m1; else ++m0;
        TC_WAIT(&bar_empty[s], ph ^ 1);
        TRACE(1);
        if (elect_one()) {
          mbar_expect_tx(&bar_full[s], 2u * (uint32_t)trows * 128u);
          uint8_t* dst = smem + kSmemStage + s * kStageBytes;
          tma_load_2d_hint(dst, tm_cur, c * kChunk, row0, &bar_full[s], p.pol_cur);
          tma_load_2d_hint(dst + kBoxBytes, tm_prev, c * kChunk, row0, &bar_full[s], p.pol_prev);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs this control flow (every value is warp-uniform); only the
    // tcgen05.mma / commit instructions sit under elect_one(), so the descriptors live in
    // uniform registers and one MMA costs a handful of issue slots.
    const uint32_t idesc1 = make_idesc(kFmtBF16, 128, kDP, 0, 0);   // B K-major  (GEMM1)
    const uint32_t idesc2 = make_idesc(kFmtBF16, 128, kQ, 0, 1);    // B MN-major (GEMM2)
    const uint32_t w_addr = smem_u32(smem + kSmemW);
    // descriptor of piece 0 / offset 0 in both views; per MMA only the low word moves
    const uint64_t desc1 = make_smem_desc_sw128(w_addr, 0, 1024);
    const uint64_t desc2 = make_smem_desc_sw128(w_addr, kSlabBytes, 1024);
    const uint32_t d1_lo = (uint32_t)desc1, d1_hi = (uint32_t)(desc1 >> 32);
    const uint32_t d2_lo = (uint32_t)desc2, d2_hi = (uint32_t)(desc2 >> 32);
    constexpr uint32_t kPiece16 = kPieceBytes >> 4;   // descriptor address units (16 B)
    auto make64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    TC_WAIT(&bar_w, 0);
    uint32_t a_cnt0 = 0, a_cnt1 = 0, g_cnt0 = 0, g_cnt1 = 0, ti = 0;
    for (int tile = 0; tile < my_tiles; ++tile, ++ti) {
      // accumulators alias the G buffers of the previous tile: wait until both were drained
      TC_WAIT(&bar_gfree[0], (g_cnt0 & 1) ^ 1);
      TC_WAIT(&bar_gfree[1], (g_cnt1 & 1) ^ 1);
      TRACE(10);
      tc_fence_after();
      // ---- GEMM1: R[128 x 64] = Y[128 x k] * W^T ----
      for (int c = 0; c < nc; ++c) {
        const int b = c & 1;
        if (b == 0) {
          TC_WAIT(&bar_aready[0], a_cnt0 & 1);
          ++a_cnt0;
        } else {
          TC_WAIT(&bar_aready[1], a_cnt1 & 1);
          ++a_cnt1;
        }
        TRACE(11);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_stage = tbase + kColStage + b * 48;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            // 64-atom slab (c >> 1), 16 atoms per k-step, 2 bytes each -> units of 16 B
            const uint32_t koff = (uint32_t)(c >> 1) * (kSlabBytes >> 4) + (uint32_t)((c & 1) * 4 + ks * 2);
            const uint32_t acc_on = (c > 0 || ks > 0) ? 1u : 0u;
            const uint64_t q1 = make64(d1_lo + koff, d1_hi);
            const uint64_t q2 = make64(d1_lo + koff + kPiece16, d1_hi);
            const uint64_t q3 = make64(d1_lo + koff + 2 * kPiece16, d1_hi);
            const uint32_t p1 = t_stage + ks * 8, p2 = p1 + 16, p3 = p1 + 32;
            // small products (p1 q3, p2 q2, p3 q1, p1 q2, p2 q1) -> R_small
            mma_ts<false>(tbase + kColAcc1, p1, q3, idesc1, acc_on);
            mma_ts<false>(tbase + kColAcc1, p2, q2, idesc1, 1);
            mma_ts<false>(tbase + kColAcc1, p3, q1, idesc1, 1);
            mma_ts<false>(tbase + kColAcc1, p1, q2, idesc1, 1);
            mma_ts<false>(tbase + kColAcc1, p2, q1, idesc1, 1);
            // leading product -> R_big
            mma_ts<false>(tbase + kColAcc0, p1, q1, idesc1, acc_on);
          }
          // one commit per chunk: the tile's last chunk signals "GEMM1 complete" instead of
          // "piece stage free" (two back-to-back commits stall the issuing thread until the
          // first one retires)
          if (c == nc - 1) mma_commit(&bar_rfull);
          else mma_commit(&bar_sfree[b]);
        }
        __syncwarp();
        TRACE(12);
      }
      // ---- GEMM2: G[128 x 64q] = r[128 x d] * W[:, chunk] ----
      TC_WAIT(&bar_rready, ti & 1);
      TRACE(14);
      tc_fence_after();
      for (int q = 0; q < nq; ++q) {
        const int b = q & 1;
        if (b == 0) {
          TC_WAIT(&bar_gfree[0], (g_cnt0 & 1) ^ 1);
          ++g_cnt0;
        } else {
          TC_WAIT(&bar_gfree[1], (g_cnt1 & 1) ^ 1);
          ++g_cnt1;
        }
        TRACE(15);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_acc = tbase + (b ? kColAcc1 : kColAcc0);
          const uint32_t t_r = tbase + kColStage;
          const uint32_t qoff = (uint32_t)q * (kSlabBytes >> 4);
          uint32_t acc_on = 0;
          // small products first (their truncation error scales with a small accumulator):
          // (r3 q1) (r2 q2) (r1 q3) (r2 q1) (r1 q2), leading (r1 q1) last
#pragma unroll
          for (int t = 0; t < 6; ++t) {
            constexpr int pa[6] = {2, 1, 0, 1, 0, 0}, pb[6] = {0, 1, 2, 0, 1, 0};
#pragma unroll
            for (int ks = 0; ks < dsteps; ++ks) {
              // 16 k-rows of 128 B per k-step = 2048 B = 128 units
              const uint64_t bd = make64(d2_lo + qoff + pb[t] * kPiece16 + ks * 128, d2_hi);
              mma_ts<false>(t_acc, t_r + pa[t] * 32 + ks * 8, bd, idesc2, acc_on);
              acc_on = 1;
            }
          }
          mma_commit(&bar_gfull[b]);
        }
        __syncwarp();
        TRACE(16);
      }
    }
  } else {
    // ===================== compute warps =====================
    // 16 warps = 4 groups of 4 (a group covers the 128 TMEM lanes).  Groups {0,2} own the
    // even phase-A chunks / G buffer 0, groups {1,3} the odd ones / G buffer 1 ("pair" = wg & 1);
    // inside a pair the two groups split the columns of every chunk in halves (wg >> 1).
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int wg = (warp - 2) >> 2;            // 0..3
    const int grp = wg & 1;                    // pair: stage ring / piece stage / G buffer
    const int half = wg >> 1;                  // which half of the pair's columns
    const int row = quad * 32 + lane;          // row inside the tile == TMEM lane
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    double dsum = 0.0;
    const float2 lr2 = make_float2(p.lr, p.lr);
    const float2 beta2 = make_float2(p.beta, p.beta);
    const bool has_out = nq > grp;   // this pair owns at least one GEMM2 chunk per tile
    const bool store_leader = (wg == grp) && (quad == 2) && (lane == 0);   // warp 2 / warp 6
    uint32_t a_cnt = 0, g_cnt = 0, ti = 0, sf_base = 0;
    const uint32_t sf_per_tile = (uint32_t)((nc - grp + 1) / 2) - ((((nc - 1) & 1) == grp) ? 1u : 0u);
    for (int tile = 0; tile < my_tiles; ++tile, ++ti, sf_base += sf_per_tile) {
      const int grow32 = first_row + tile * row_step + row;
      const int64_t grow = grow32;
      const bool row_ok = row < trows && grow32 < row_limit;   // lanes beyond the tile carry stale data
      uint32_t keep_stage = grp;   // set in phase A whenever has_out
      // pull this thread's 64-byte slice of x towards L2 now; phase B reads it ~10k cycles later
      if (row_ok && wg * 16 < p.d)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.x + grow * p.d + wg * 16));
      // ---------------- phase A ----------------
      for (int c = grp; c < nc; c += 2) {
        // a_cnt = chunks this pair consumed so far = index into its private stage ring
        const uint32_t s = grp + 2 * (a_cnt & 1), ph = (a_cnt >> 1) & 1;
        TC_WAIT(&bar_full[s], ph);
        TRACE(20);
        const uint8_t* zc_s = smem + kSmemStage + s * kStageBytes;
        const uint8_t* zp_s = zc_s + kBoxBytes;
        // y = z_cur + beta (z_cur - z_prev) on fp32 pairs; |z_cur - z_prev| feeds the lagged
        // stop-test sum; then the exact three-way bf16 split of every y (16 atoms per thread)
        uint32_t yb[16], w1[8], w2[8], w3[8];
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = sw128_offset(row, (half * 4 + j) * 16);
          const float4 zc = *reinterpret_cast<const float4*>(zc_s + off);
          float2 ya = make_float2(zc.x, zc.y), yc = make_float2(zc.z, zc.w);
          if (p.use_prev) {
            const float4 zp = *reinterpret_cast<const float4*>(zp_s + off);
            const float2 da = sub2(ya, make_float2(zp.x, zp.y));
            const float2 dc = sub2(yc, make_float2(zp.z, zp.w));
            part += (fabsf(da.x) + fabsf(da.y)) + (fabsf(dc.x) + fabsf(dc.y));
            ya = __fadd2_rn(ya, __fmul2_rn(beta2, da));
            yc = __fadd2_rn(yc, __fmul2_rn(beta2, dc));
          }
          yb[4 * j + 0] = __float_as_uint(ya.x);
          yb[4 * j + 1] = __float_as_uint(ya.y);
          yb[4 * j + 2] = __float_as_uint(yc.x);
          yb[4 * j + 3] = __float_as_uint(yc.y);
          split3_pair(ya, w1[2 * j], w2[2 * j], w3[2 * j]);
          split3_pair(yc, w1[2 * j + 1], w2[2 * j + 1], w3[2 * j + 1]);
        }
        // stage consumed (values are in registers).  The last stage of the tile is withheld:
        // phase C stages its output there and releases it afterwards.
        if (has_out && c + 2 >= nc) keep_stage = s;
        else mbar_arrive(&bar_empty[s]);
        if (row_ok) dsum += (double)part;
        TRACE(21);
        // The piece stage is free once the MMAs of its previous chunk completed (the first
        // chunk of a tile needs no wait: everybody saw GEMM1 of the previous tile complete).
        // Phases of bar_sfree are counted in sf_base: one per chunk of this pair except the
        // tile's very last chunk, which commits to bar_rfull instead.
        if (c >= 2) TC_WAIT(&bar_sfree[grp], (sf_base + (uint32_t)(c >> 1) - 1u) & 1u);
        // first TMEM store of a tile: the OTHER pair must have finished phase C of the previous
        // tile (its reads of the y master, and -- having seen every G chunk -- all GEMM2 MMAs
        // that read the r pieces aliased with the piece stages)
        else if (ti > 0) TC_WAIT(&bar_tdone[grp ^ 1], (ti - 1) & 1);
        TRACE(22);
        ++a_cnt;
        tc_fence_after();
        tmem_st16(tbase + lane_base + kColY + c * kChunk + half * 16, yb);
        const uint32_t t_stage = tbase + lane_base + kColStage + grp * 48 + half * 8;
        tmem_st8(t_stage, w1);
        tmem_st8(t_stage + 16, w2);
        tmem_st8(t_stage + 32, w3);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_aready[grp]);
        TRACE(23);
      }
      // ---------------- phase B: r = R - x, pieces of r (16 features per thread) ----------------
      {
        // this thread's features of x, straight from global (prefetched into L2 at tile start;
        // issued before the wait on GEMM1 so the latency overlaps the tail of the MMAs)
        float4 xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int col = wg * 16 + 4 * j;
          xv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && col < p.d)
            xv[j] = __ldg(reinterpret_cast<const float4*>(p.x + grow * p.d + col));
        }
        TC_WAIT(&bar_rfull, ti & 1);
        TRACE(30);
        tc_fence_after();
        uint32_t rb[16], rs[16];
        tmem_ld16(tbase + lane_base + kColAcc0 + wg * 16, rb);
        tmem_ld16(tbase + lane_base + kColAcc1 + wg * 16, rs);
        tmem_wait_ld();
        uint32_t w1[8], w2[8], w3[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          // r = (R_big + R_small) - x, two roundings
          const float2 ra = sub2(__fadd2_rn(make_float2(__uint_as_float(rb[4 * j + 0]), __uint_as_float(rb[4 * j + 1])),
                                            make_float2(__uint_as_float(rs[4 * j + 0]), __uint_as_float(rs[4 * j + 1]))),
                                 make_float2(xv[j].x, xv[j].y));
          const float2 rc = sub2(__fadd2_rn(make_float2(__uint_as_float(rb[4 * j + 2]), __uint_as_float(rb[4 * j + 3])),
                                            make_float2(__uint_as_float(rs[4 * j + 2]), __uint_as_float(rs[4 * j + 3]))),
                                 make_float2(xv[j].z, xv[j].w));
          split3_pair(ra, w1[2 * j], w2[2 * j], w3[2 * j]);
          split3_pair(rc, w1[2 * j + 1], w2[2 * j + 1], w3[2 * j + 1]);
        }
        const uint32_t t_r = tbase + lane_base + kColStage + wg * 8;
        tmem_st8(t_r, w1);
        tmem_st8(t_r + 32, w2);
        tmem_st8(t_r + 64, w3);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_rready);
        TRACE(31);
      }
      // ---------------- phase C: fused update ----------------
      // z_next goes to HBM as coalesced TMA stores: the pair stages each 64-atom chunk (two
      // [128 x 32] boxes, one per group, 128-byte swizzle, conflict-free 16-byte writes) in
      // the input stage it consumed last in phase A and withheld from the producer.
      uint8_t* out_s = smem + kSmemStage + keep_stage * kStageBytes;
      for (int q = grp; q < nq; q += 2) {
        TC_WAIT(&bar_gfull[grp], g_cnt & 1);
        TRACE(40);
        ++g_cnt;
        tc_fence_after();
        uint32_t g[32], yv[32];
        tmem_ld32(tbase + lane_base + (grp ? kColAcc1 : kColAcc0) + half * 32, g);
        tmem_ld32(tbase + lane_base + kColY + q * kQ + half * 32, yv);
        // the previous chunk's TMA stores must have finished READING the staging buffer
        if (store_leader) tma_store_wait_read<0>();
        pair_sync(grp);
        tmem_wait_ld();
        // accumulator read: hand the G buffer back before the stores
        tc_fence_before();
        mbar_arrive(&bar_gfree[grp]);
        uint8_t* box = out_s + half * kBoxBytes;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 oa = ista_update_pair(
              make_float2(__uint_as_float(yv[4 * j + 0]), __uint_as_float(yv[4 * j + 1])),
              make_float2(__uint_as_float(g[4 * j + 0]), __uint_as_float(g[4 * j + 1])), lr2, p.lam);
          const float2 oc = ista_update_pair(
              make_float2(__uint_as_float(yv[4 * j + 2]), __uint_as_float(yv[4 * j + 3])),
              make_float2(__uint_as_float(g[4 * j + 2]), __uint_as_float(g[4 * j + 3])), lr2, p.lam);
          *reinterpret_cast<float4*>(box + sw128_offset(row, j * 16)) =
              make_float4(oa.x, oa.y, oc.x, oc.y);
        }
        fence_proxy_async_smem();   // generic-proxy writes -> visible to the TMA engine
        pair_sync(grp);
        if (store_leader) {
          const int row0 = first_row + tile * row_step;
          tma_store_2d_hint(tm_prev, out_s, q * kQ, row0, p.pol_prev);               // rows / columns
          tma_store_2d_hint(tm_prev, out_s + kBoxBytes, q * kQ + 32, row0, p.pol_prev);  // beyond n, k are clipped
          tma_store_commit();
        }
        TRACE(41);
      }
      if (has_out) {
        // release the withheld stage to the producer once the last stores have read it
        if (store_leader) tma_store_wait_read<0>();
        pair_sync(grp);
        mbar_arrive(&bar_empty[keep_stage]);
      }
      // y master / r pieces are rewritten by the next tile, but only its first TMEM store has to
      // wait for the other pair (see phase A): the pair that finishes first already loads and
      // splits its first chunk of the next tile meanwhile
      tc_fence_before();
      mbar_arrive(&bar_tdone[grp]);
      TRACE(50);
    }
    if (store_leader) tma_store_wait_all<0>();
    // lagged stop-test sum of the previous iteration
    dsum = warp_sum(dsum);
    if (lane == 0) red[warp - 2] = dsum;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    if (warp == 2 && lane == 0 && p.use_prev && p.ctl.hist != nullptr && p.ctl.iter >= 1) {
      double s = 0.0;
      for (int i = 0; i < 16; ++i) s += red[i];
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
int make_map(CUtensorMap* map, const float* base, int64_t rows, int cols, int tile_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return LASSO_B200_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {32, (cuuint32_t)tile_rows};
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
  unsigned long long* trace = nullptr;   // LASSO_B200_TRACE=<file>
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
  const char* trace_path = getenv("LASSO_B200_TRACE");
  if (trace_path && !S.trace) LASSO_CUDA_TRY(cudaMalloc(&S.trace, 32 * 512 * 8));
  if (S.trace) LASSO_CUDA_TRY(cudaMemsetAsync(S.trace, 0, 32 * 512 * 8, st));
  if (!S.attr_set) {
    const void* kernels[4] = {(const void*)fista_tc_kernel<1>, (const void*)fista_tc_kernel<2>,
                              (const void*)fista_tc_kernel<3>, (const void*)fista_tc_kernel<4>};
    for (const void* kfn : kernels)
      LASSO_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    S.attr_set = true;
  }
  prep_w_image_kernel<<<(kDP * kKP + 255) / 256, 256, 0, st>>>(a.w, a.d, a.k, S.w_image);
  LASSO_CHECK_LAUNCH();
  count_launch();

  // Tile height: the smallest multiple of 8 rows for which the tiles still fit the same number
  // of waves as 128-row tiles would need -- the last wave is then (almost) full instead of
  // leaving SMs idle (n = 65536 on 148 SMs: 592 slots, 111 -> 112 rows, 586 tiles).
  const int64_t slots = (int64_t)S.num_sms;
  const int64_t waves = ((a.n + kTileM - 1) / kTileM + slots - 1) / slots;
  int64_t tile_rows = (a.n + waves * slots - 1) / (waves * slots);
  tile_rows = ((tile_rows + 7) / 8) * 8;
  if (tile_rows > kTileM) tile_rows = kTileM;
  const int64_t ntiles = (a.n + tile_rows - 1) / tile_rows;
  const unsigned grid = (unsigned)(ntiles < S.num_sms ? ntiles : S.num_sms);

  // v1 schedule (see TcParams): two classes of CTAs.  With >= 2 waves the odd CTAs start
  // `stagger` cycles late and get shorter tiles (measured on B200: all SMs in phase A at once
  // saturate HBM while it idles during the MMA-bound phases; de-phasing half of the SMs cut
  // the step from 55.6 to 52.5 us even with the delay simply added on top).
  int64_t h[2] = {tile_rows, tile_rows};
  int stagger = 0;
  const char* st_env = getenv("LASSO_B200_STAGGER");
  if (waves >= 2 && grid == (unsigned)S.num_sms) {
    stagger = st_env ? atoi(st_env) : 12000;   // swept on B200 at C2: 0 -> 56.3 us, 8k -> 54.0, 12k -> 51.6, 16k -> 56.9
    // one row costs ~80 cycles of a tile's HBM share: take stagger / (waves * 80) rows off the
    // late class per tile and give them to the early class
    const double h_avg = (double)a.n / (double)(waves * grid);
    int64_t dh = (int64_t)(stagger / (double)(waves * 80) + 0.5);
    int64_t h0 = ((int64_t)(h_avg + dh / 2.0 + 0.999) + 7) / 8 * 8;
    if (h0 > kTileM) h0 = kTileM;
    const int64_t n_even = ((int64_t)grid + 1) / 2, n_odd = (int64_t)grid / 2;
    int64_t rest = a.n - n_even * waves * h0;
    int64_t h1 = rest > 0 ? ((rest + n_odd * waves - 1) / (n_odd * waves) + 7) / 8 * 8 : 8;
    if (h1 > kTileM) {   // cannot happen for a consistent h_avg, but never drop rows
      h0 = tile_rows;
      h1 = tile_rows;
      stagger = 0;
    }
    h[0] = h0;
    h[1] = h1;
  }
  const int64_t n_even = ((int64_t)grid + 1) / 2;
  const int64_t region_a = n_even * waves * h[0];
  CUtensorMap tm_za, tm_zb, tm_za1, tm_zb1;
  int rc;
  if ((rc = make_map(&tm_za, a.z_a, a.n, a.k, (int)h[0]))) return rc;
  if ((rc = make_map(&tm_zb, a.z_b, a.n, a.k, (int)h[0]))) return rc;
  if ((rc = make_map(&tm_za1, a.z_a, a.n, a.k, (int)h[1]))) return rc;
  if ((rc = make_map(&tm_zb1, a.z_b, a.n, a.k, (int)h[1]))) return rc;

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
    p.tile_rows = (int)tile_rows;
    p.stagger_cycles = stagger;
    // L2 residency: the two code buffers and x (151 MB at C2) cycle through a 126 MB L2, which
    // a plain LRU turns into 100 % misses.  Buffer A is therefore always accessed with
    // evict-last priority and buffer B with evict-first: A stays resident (reads and in-place
    // writes hit L2), only B streams through HBM.
    {
      const char* l2_env = getenv("LASSO_B200_L2PIN");
      const bool pin = !(l2_env && l2_env[0] == '0');
      const uint64_t pol_a = pin ? kEvictLast : kEvictNormal, pol_b = pin ? kEvictFirst : kEvictNormal;
      p.pol_cur = p.cur_is_a ? pol_a : pol_b;
      p.pol_prev = p.cur_is_a ? pol_b : pol_a;
    }
    p.cls_rows[0] = (int)h[0];
    p.cls_rows[1] = (int)h[1];
    p.cls_base[0] = 0;
    p.cls_end[0] = region_a < a.n ? region_a : a.n;
    p.cls_base[1] = region_a;
    p.cls_end[1] = a.n;
    p.waves = (int)waves;
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
    p.trace = (it == a.maxiter - 1) ? S.trace : nullptr;
    const int dsteps = (a.d + 15) / 16;
#define LASSO_TC_LAUNCH(DS)                                                                \
  do {                                                                                     \
    fista_tc_kernel<DS><<<grid, kThreads, kSmemBytes, st>>>(tm_za, tm_zb, tm_za1, tm_zb1, p);     \
  } while (0)
    switch (dsteps) {
      case 1: LASSO_TC_LAUNCH(1); break;
      case 2: LASSO_TC_LAUNCH(2); break;
      case 3: LASSO_TC_LAUNCH(3); break;
      default: LASSO_TC_LAUNCH(4); break;
    }
#undef LASSO_TC_LAUNCH
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  if (S.trace && trace_path) {
    static unsigned long long host_trace[32 * 512];
    LASSO_CUDA_TRY(cudaMemcpyAsync(host_trace, S.trace, sizeof(host_trace), cudaMemcpyDeviceToHost, st));
    LASSO_CUDA_TRY(cudaStreamSynchronize(st));
    if (FILE* f = fopen(trace_path, "w")) {
      for (int w = 0; w < 32; ++w)
        for (int i = 0; i < 512 && host_trace[w * 512 + i]; ++i)
          fprintf(f, "%d %llu %llu\n", w, host_trace[w * 512 + i] >> 8, host_trace[w * 512 + i] & 255);
      fclose(f);
    }
  }
  if (S.dbg_host) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (S.dbg_host[0]) {
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
