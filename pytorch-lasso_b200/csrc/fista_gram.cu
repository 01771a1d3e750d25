// K1g: Gram-form FISTA on tcgen05 for dictionaries whose feature count does not fit the two-GEMM kernels
// (128 < d, k <= 320: the shape of the reference's notebook, d = 289, k = 300).
//
// The reference's gradient (ista.py:71-73) is  (y W^T - x) W.  With  G_w = W^T W  (k x k) and  b = x W  (n x k)
// taken ONCE per solve it is  y G_w - b : one GEMM per iteration with K = k instead of two with K = d and K = k,
// no residual tile, no x tile.  What has to live on chip per 128-row tile is then only y:
//
//   TMEM      columns [0,160) h pieces, [160,320) l pieces of y' (fp16 x 2, 16 atoms per 8 columns, the layout
//             of the resident kernel's piece slots), [320,384) lead accumulator, [384,448) cross accumulator
//   smem      two stages of one 64-atom output slab of G' = sg G_w: [h | l][kpad rows][128 B], 128-byte swizzle,
//             MN-major (output atoms contiguous) -- at most 80 KB per stage, fetched by ONE cp.async.bulk from
//             the image in L2; the issuer prefetches slab c + 1 while slab c is multiplied
//   per slab  3 MMAs per 16 input atoms (h h' -> L, h l' + l h' -> C), N = 64; the compute warps drain L + C,
//             rebuild y' = h + l from their own piece columns, take  v = y' - (lr / sg)(L + C - b'),
//             soft-threshold and store z+ in the caller's units
//   scaling   row r works in units s_r = sx_r / sw (x' = sx_r x in [64, 128), W' = sw W in [8, 16), as in the
//             resident kernel); sg puts max |G'| in [256, 512); all powers of two, so every rescaling is exact.
//             An iterate beyond the fp16 operand range raises a flag and the batch is solved again by the FFMA
//             kernel (cabi.cu).
//   threads   8 compute warps with 168 registers each (lane = row of the tile; warps 0-3 / 4-7 take the even / odd
//             32-atom units) + one issuer warp (slab copies, MMAs)
// ALL iterations of a tile run in one launch: tiles are independent, and every thread reads back exactly the
// elements it stored one iteration earlier.  The batch-global stop test is taken afterwards from the recorded sums
// (the record of iteration i is summed while iteration i + 1 loads z_i and z_{i+1}); cabi.cu replays an early stop
// with that iteration count.  Codes stay in caller units in HBM, so there is no set-up or unscale pass over them.
// Per iteration a tile reads z_i, z_{i-1}, b and writes z_{i+1}: 4 x n k floats, plus the 2 x 2 x kpad x k bytes
// of G' per tile from L2.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace lasso {
using namespace sm100;
void small_gram_launch(const float* w, int d, int k, int m, int len, int row_gram, double* gram, cudaStream_t st);
namespace {

constexpr int kGfThreads = 288;            // 8 compute warps + the MMA / copy issuer (last warp)
constexpr int kGfIssuer = 8;
constexpr int kGfTileM = 128;
constexpr int kGfQ = 64;                   // output atoms per slab
constexpr int kGfKMax = 320;
constexpr uint32_t kGfColPH = 0;           // h pieces of y'
constexpr uint32_t kGfColPL = 160;         // l pieces
constexpr uint32_t kGfColL = 320;          // lead accumulator
constexpr uint32_t kGfColC = 384;          // cross accumulator
constexpr uint32_t kGfStage = 2u * kGfKMax * 128u;      // 80 KB
constexpr uint32_t kGfSmemImage = 2u * kGfStage;
constexpr uint32_t kGfSmem = kGfSmemImage + 8u * 8192u;      // + two 4 KB scratches per compute warp (transposition, b prefetch)

struct GfScalars {
  float isw;    // 1 / sw
  float sg;     // G' = sg G_w
  float lrs;    // lr / sg
  int bad;      // dictionary / Gram not finite, or the scaled step not representable
};

struct GfParams {
  const uint8_t* image;      // [nq][h | l][kpad][128 B]
  const float* b;            // [n][k]  x W, caller units
  const float* row_scale;    // [n]  s_r
  float* z_a;                // z_i lives in (i even ? z_a : z_b); z_{i+1} is written over z_{i-1}
  float* z_b;
  const float* beta;         // [iters] momentum coefficient of iteration i (0 for i = 0)
  double* hist;              // [iters] sum |z_i - z_{i+1}| (float64, zero-initialised) or nullptr
  int iters;
  int64_t n;
  int k, nks, nq, tile_rows;
  uint32_t piece_bytes;      // kpad * 128
  float lam, limit;
  const GfScalars* scal;
  int* flag;
  long long* trace;          // LASSO_B200_GRAM_TRACE: clock64 time line of CTA 0 (issuer: [0,64), compute warp 0: [64,128))
};

#define GF_WAIT(bar, parity)                                                              \
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
      if (++_n > (1u << 18)) __trap();   /* a protocol bug must not hang the GPU */       \
    }                                                                                     \
  } while (0)

// 16-byte asynchronous copy global -> shared (no staging registers); src_bytes = 0 fills with zeros
__device__ __forceinline__ void gf_cp_async16(void* smem_dst, const void* gsrc, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void gf_cp_async_wait() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void gf_split2(float a, float b, uint32_t& wh, uint32_t& wl) {
  const float ta = __uint_as_float(__float_as_uint(a) & 0xFFFFE000u);
  const float tb = __uint_as_float(__float_as_uint(b) & 0xFFFFE000u);
  const __half2 h = __floats2half2_rn(ta, tb);
  const __half2 l = __floats2half2_rn(__fsub_rn(a, ta), __fsub_rn(b, tb));
  wh = *reinterpret_cast<const uint32_t*>(&h);
  wl = *reinterpret_cast<const uint32_t*>(&l);
}

#ifdef LASSO_GRAM_TRACE
#define GF_MARK(slot) do { if (p.trace && blockIdx.x == 0 && lane == 0 && (warp == kGfIssuer || warp == 0)) p.trace[(warp == kGfIssuer ? 0 : 64) + ((slot) & 63)] = clock64(); } while (0)
#else
#define GF_MARK(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kGfThreads, 1) fista_gram_kernel(GfParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_full[2], bar_empty[2], bar_pready, bar_gfull, bar_gfree;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  GF_MARK(0);
  const int64_t ntiles = (p.n + p.tile_rows - 1) / p.tile_rows;
  const int my_tiles = blockIdx.x < ntiles ? (int)((ntiles - 1 - blockIdx.x) / gridDim.x) + 1 : 0;

  if (tid == 0) {
    mbar_init(&bar_full[0], 1);
    mbar_init(&bar_full[1], 1);
    mbar_init(&bar_empty[0], 1);
    mbar_init(&bar_empty[1], 1);
    mbar_init(&bar_pready, 256);
    mbar_init(&bar_gfull, 1);
    mbar_init(&bar_gfree, 256);
    fence_mbar_init();
  }
  if (warp == kGfIssuer) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  GF_MARK(1);
  const uint32_t slab_bytes = 2u * p.piece_bytes;
  const int nchunks = my_tiles * p.iters * p.nq;

  if (warp == kGfIssuer) {
    // ===================== slab copies + MMA issue =====================
    const uint32_t idesc = make_idesc(kFmtF16, 128, kGfQ, 0, 1);   // A from TMEM, B MN-major
    const bool leader = elect_one();
    auto load = [&](int c) {
      const uint32_t s = (uint32_t)c & 1u;
      mbar_expect_tx(&bar_full[s], slab_bytes);
      bulk_load(smem + s * kGfStage, p.image + (size_t)(c % p.nq) * slab_bytes, slab_bytes, &bar_full[s]);
    };
    if (leader) {
      if (nchunks > 0) load(0);
      if (nchunks > 1) load(1);
    }
    __syncwarp();
    int c = 0;
    for (int u = 0; u < my_tiles * p.iters; ++u) {            // (tile, iteration) pairs
      GF_WAIT(&bar_pready, (uint32_t)u & 1u);
      GF_MARK(2);
      for (int q = 0; q < p.nq; ++q, ++c) {
        const uint32_t s = (uint32_t)c & 1u;
        GF_WAIT(&bar_full[s], ((uint32_t)c >> 1) & 1u);
        GF_MARK(4 + 4 * q);
        if (c > 0) GF_WAIT(&bar_gfree, ((uint32_t)c - 1u) & 1u);
        GF_MARK(5 + 4 * q);
        tc_fence_after();
        if (leader) {
          const uint64_t desc = make_smem_desc_sw128(smem_u32(smem + s * kGfStage), p.piece_bytes, 1024);
          const uint32_t d_lo = (uint32_t)desc, d_hi = (uint32_t)(desc >> 32);
          const uint32_t l_off = p.piece_bytes >> 4;
#pragma unroll 4
          for (int ks = 0; ks < p.nks; ++ks) {
            const uint64_t bh = ((uint64_t)d_hi << 32) | (d_lo + (uint32_t)ks * 128u);
            const uint64_t bl = ((uint64_t)d_hi << 32) | (d_lo + l_off + (uint32_t)ks * 128u);
            const uint32_t ah = tbase + kGfColPH + (uint32_t)ks * 8u, al = tbase + kGfColPL + (uint32_t)ks * 8u;
            const uint32_t acc_on = ks == 0 ? 0u : 1u;
            mma_ts<false>(tbase + kGfColC, ah, bl, idesc, acc_on);
            mma_ts<false>(tbase + kGfColC, al, bh, idesc, 1);
            mma_ts<false>(tbase + kGfColL, ah, bh, idesc, acc_on);
          }
          mma_commit(&bar_empty[s]);
          mma_commit(&bar_gfull);
        }
        __syncwarp();
        GF_MARK(6 + 4 * q);
        // the stage of slab c + 2 is the one slab c occupies: refill it as soon as slab c has been multiplied
        // (slab c + 1 is already in flight or resident, so the tensor pipe does not wait for this copy)
        if (c + 2 < nchunks) {
          GF_WAIT(&bar_empty[s], ((uint32_t)c >> 1) & 1u);
          if (leader) load(c + 2);
          __syncwarp();
          GF_MARK(7 + 4 * q);
        }
      }
    }
  } else {
    // ===================== compute warps =====================
    // Global memory is touched in a row-major "wide" layout (lane = (row lane/8 + 4 i, 16-byte chunk lane%8):
    // a warp instruction covers 4 rows x 128 contiguous bytes) and TMEM wants lane = row, so every operand goes
    // through a 4 KB per-warp shared-memory scratch (128-byte swizzle: conflict-free both ways).  Thread-per-row
    // global accesses would put 32 cache lines under every instruction (measured: 37 us per tile-iteration).
    // 8 compute warps with up to 168 registers each rather than 16 with 96: the loads of the next unit of work
    // are always in flight while the current one is converted, and that needs 64 registers per operand pair.
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const float sg = p.scal->sg, lrs = p.scal->lrs;
    uint8_t* scr = smem + kGfSmemImage + warp * 8192;
    uint8_t* scr_b = scr + 4096;                               // b of the next slab lands here (cp.async)
    const int wc = lane & 7;                                   // this lane's 16-byte chunk in the wide layout
    auto scr_at = [&](int r, int c) { return reinterpret_cast<float4*>(scr + r * 128 + ((c ^ (r & 7)) << 4)); };
    auto scrb_at = [&](int r, int c) { return reinterpret_cast<float4*>(scr_b + r * 128 + ((c ^ (r & 7)) << 4)); };
    const int nkp = (p.nks + 1) >> 1;                          // units of 32 atoms (two k-steps)
    bool bad = false;
    int c = 0;
    for (int t = 0; t < my_tiles; ++t) {
      const int64_t tile = blockIdx.x + (int64_t)t * gridDim.x;
      const int64_t row0 = tile * p.tile_rows;
      const int row = quad * 32 + lane;
      const int rows_here = (int)min((int64_t)p.tile_rows, p.n - row0);
      const bool valid = row < rows_here;
      const float s_r = valid ? __ldg(p.row_scale + row0 + row) : 1.f;
      const float inv_s = 1.f / s_r;
      const float lam_r = p.lam * s_r, bs = s_r * sg;
      // wide layout: rows wr0 + 4 i of the tile; invalid rows read row 0 (their values are masked or never stored)
      const int64_t tile_off = row0 * p.k;
      const int wr0 = quad * 32 + (lane >> 3);
      auto wok_f = [&](int i) { return wr0 + 4 * i < rows_here; };
      auto woff_f = [&](int i) { return (wok_f(i) ? wr0 + 4 * i : 0) * p.k + 4 * wc; };
      // All iterations of this tile, then the next tile: tiles are independent (the stop test is the only
      // batch-global quantity; it is taken from the recorded sums afterwards, cabi.cu), and every thread reads back
      // exactly the elements it stored one iteration earlier (same warp, same wide layout), so the iterations need
      // no synchronisation beyond program order.  z is therefore read with plain loads (not the .nc path).
      const float* b_t = p.b + tile_off;
      auto issue_b = [&](int q) {
        const int a0 = q * kGfQ + half * 32;
        const bool col_ok = a0 + 4 * wc < p.k;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          gf_cp_async16(scrb_at((lane >> 3) + 4 * i, wc), b_t + woff_f(i) + (col_ok ? a0 : -4 * wc),
                        (wok_f(i) && col_ok) ? 16u : 0u);
      };
      for (int it = 0; it <= p.iters; ++it) {
      const float* zc_t = ((it & 1) ? p.z_b : p.z_a) + tile_off;
      float* zn_t = ((it & 1) ? p.z_a : p.z_b) + tile_off;
      const float* zp_t = it > 0 ? zn_t : zc_t;                // z_{i-1}; z_{i+1} goes over it
      const bool need_hist = p.hist != nullptr && it > 0;      // this pass sums |z_{i-1} - z_i| of the previous iteration
      const bool last_pass = it == p.iters;                    // only that sum is left to take
      if (last_pass && !need_hist) break;
      const float beta = last_pass ? 0.f : __ldg(p.beta + it);
      if (!last_pass) issue_b(0);
      if (!last_pass) GF_MARK(1);
      float msum = 0.f;
      // ---- y' = s_r (z_i + beta (z_i - z_{i-1})) -> fp16 pieces in TMEM (16 atoms per 8 columns) ----
      // (the previous iteration's last slab has been drained by every thread, so its MMAs are complete)
      // Loads are unconditional (addresses clamped into the tile, values masked afterwards): a branch around a
      // load and its consumer serialises the round trips (measured: 28 k cycles for this phase).
      {
        float4 zc[8], zp[8];
        auto issue = [&](int kp) {
          const int col = kp * 32 + 4 * wc < p.k ? kp * 32 : -4 * wc;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            zc[i] = *reinterpret_cast<const float4*>(zc_t + woff_f(i) + col);
            zp[i] = *reinterpret_cast<const float4*>(zp_t + woff_f(i) + col);
          }
        };
        if (half < nkp) issue(half);
        for (int kp = half; kp < nkp; kp += 2) {
          const bool col_ok = kp * 32 + 4 * wc < p.k;          // k % 4 == 0: a float4 is inside or outside
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 y = make_float4(momentum_point(zc[i].x, zp[i].x, beta), momentum_point(zc[i].y, zp[i].y, beta),
                                         momentum_point(zc[i].z, zp[i].z, beta), momentum_point(zc[i].w, zp[i].w, beta));
            const bool ok = wok_f(i) && col_ok;
            *scr_at((lane >> 3) + 4 * i, wc) = ok ? y : make_float4(0.f, 0.f, 0.f, 0.f);
            if (need_hist && ok)        // stop-test record of the PREVIOUS iteration: sum |z_{i-1} - z_i| (ista.py:93)
              msum += fabsf(__fsub_rn(zp[i].x, zc[i].x)) + fabsf(__fsub_rn(zp[i].y, zc[i].y)) +
                      fabsf(__fsub_rn(zp[i].z, zc[i].z)) + fabsf(__fsub_rn(zp[i].w, zc[i].w));
          }
          if (kp + 2 < nkp) issue(kp + 2);
          __syncwarp();
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const int ks = 2 * kp + sub;
            if (ks < p.nks && !last_pass) {                    // warp-uniform
              uint32_t wh[8], wl[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 v = *scr_at(lane, 4 * sub + j);
                const float y0 = v.x * s_r, y1 = v.y * s_r, y2 = v.z * s_r, y3 = v.w * s_r;
                bad |= !(fabsf(y0) < p.limit) | !(fabsf(y1) < p.limit) | !(fabsf(y2) < p.limit) | !(fabsf(y3) < p.limit);
                gf_split2(y0, y1, wh[2 * j], wl[2 * j]);
                gf_split2(y2, y3, wh[2 * j + 1], wl[2 * j + 1]);
              }
              tmem_st8(tbase + lane_base + kGfColPH + (uint32_t)ks * 8u, wh);
              tmem_st8(tbase + lane_base + kGfColPL + (uint32_t)ks * 8u, wl);
            }
          }
          __syncwarp();
        }
      }
      if (need_hist) {
        const double w = warp_sum((double)msum);
        if (lane == 0 && w != 0.0) atomicAdd(p.hist + (it - 1), w);
      }
      if (last_pass) break;
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&bar_pready);
      GF_MARK(2);
      // ---- per output slab: drain L + C, step, threshold, store ----
      // b of slab q + 1 is fetched right after slab q has been drained
      // b of a slab travels global -> shared memory by cp.async (no staging registers), one slab ahead; the first
      // slab's copy was issued before the extrapolation above
      for (int q = 0; q < p.nq; ++q, ++c) {
        const int a0 = q * kGfQ + half * 32;
        const int ks0 = q * 4 + half * 2;
        const bool in_k = a0 < p.k;                            // warp-uniform
        const bool col_ok = a0 + 4 * wc < p.k;
        float gs[32];
        gf_cp_async_wait();
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *scrb_at(lane, j);
          gs[4 * j] = v.x * bs; gs[4 * j + 1] = v.y * bs; gs[4 * j + 2] = v.z * bs; gs[4 * j + 3] = v.w * bs;
        }
        __syncwarp();
        if (q + 1 < p.nq) issue_b(q + 1);
        GF_MARK(4 + 4 * q);
        GF_WAIT(&bar_gfull, (uint32_t)c & 1u);
        GF_MARK(5 + 4 * q);
        tc_fence_after();
#pragma unroll
        for (int qt = 0; qt < 4; ++qt) {            // quarters: 16 registers of accumulators in flight at a time
          uint32_t lv[8], cv[8];
          tmem_ld8(tbase + lane_base + kGfColL + (uint32_t)(half * 32 + qt * 8), lv);
          tmem_ld8(tbase + lane_base + kGfColC + (uint32_t)(half * 32 + qt * 8), cv);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i)
            gs[qt * 8 + i] = __fsub_rn(__fadd_rn(__uint_as_float(lv[i]), __uint_as_float(cv[i])), gs[qt * 8 + i]);
        }
        tc_fence_before();
        mbar_arrive(&bar_gfree);
        GF_MARK(6 + 4 * q);
        if (!in_k) continue;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          const int ks = ks0 + sub;
          if (ks < p.nks) {                                    // warp-uniform
            uint32_t ph[8], pl[8];
            tmem_ld8(tbase + lane_base + kGfColPH + (uint32_t)ks * 8u, ph);
            tmem_ld8(tbase + lane_base + kGfColPL + (uint32_t)ks * 8u, pl);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float zn[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * j + e;
                const __half2 hh = *reinterpret_cast<const __half2*>(&ph[i >> 1]);
                const __half2 ll = *reinterpret_cast<const __half2*>(&pl[i >> 1]);
                const float yy = (i & 1) ? __high2float(hh) + __high2float(ll) : __low2float(hh) + __low2float(ll);
                zn[e] = ista_update(yy, gs[16 * sub + i], lrs, lam_r) * inv_s;
              }
              *scr_at(lane, 4 * sub + j) = make_float4(zn[0], zn[1], zn[2], zn[3]);
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 v = *scr_at((lane >> 3) + 4 * i, wc);
          if (wok_f(i) && col_ok) *reinterpret_cast<float4*>(zn_t + woff_f(i) + a0) = v;
        }
        __syncwarp();
        GF_MARK(7 + 4 * q);
      }
      }   // iterations
    }
    if (bad) atomicOr(p.flag, 1);
  }
  GF_MARK(3);
  tc_fence_before();
  __syncthreads();
  if (warp == kGfIssuer) tmem_dealloc(tbase, 512);
}

// ---- once per solve ---------------------------------------------------------------------------------------
// max |W| and max |G_w| as float bit patterns (non-negative floats order like unsigned integers; a NaN lands
// above the infinity pattern and is caught by the range test in gf_setup_kernel)
__global__ void gf_max_kernel(const float* __restrict__ w, int wcount, const double* __restrict__ gram, int gcount,
                              unsigned* __restrict__ maxbits) {
  unsigned mw = 0, mg = 0;
  const int stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
  for (int i = i0; i < wcount; i += stride) mw = max(mw, __float_as_uint(w[i]) & 0x7FFFFFFFu);
  for (int i = i0; i < gcount; i += stride) mg = max(mg, __float_as_uint((float)gram[i]) & 0x7FFFFFFFu);
  mw = __reduce_max_sync(0xffffffffu, mw);
  mg = __reduce_max_sync(0xffffffffu, mg);
  if ((threadIdx.x & 31) == 0) {
    if (mw) atomicMax(&maxbits[0], mw);
    if (mg) atomicMax(&maxbits[1], mg);
  }
}

// sw from max |W|, sg from max |G_w|, the scaled step
__global__ void gf_setup_kernel(const unsigned* __restrict__ maxbits, float lr, GfScalars* __restrict__ sc,
                                int* __restrict__ flag) {
  const float mw = __uint_as_float(maxbits[0]), mg = __uint_as_float(maxbits[1]);
  int bad = !(mw < 3.0e38f) || !(mg < 3.0e38f);
  int ew = 0, eg = 0;
  if (!bad && mw > 0.f) ew = 3 - ilogbf(mw);       // max |W'| in [8, 16)
  if (!bad && mg > 0.f) eg = 8 - ilogbf(mg);       // max |G'| in [256, 512)
  ew = max(-60, min(60, ew));
  eg = max(-60, min(60, eg));
  const float sw = ldexpf(1.f, ew), sg = ldexpf(1.f, eg);
  const float lrs = lr / sg;
  if (!(lrs > 1e-30f) || !(lrs < 1e30f)) bad = 1;
  sc->isw = 1.f / sw;
  sc->sg = sg;
  sc->lrs = lrs;
  sc->bad = bad;
  *flag = bad;
}

// G_w (float64, k x k) -> swizzled fp16 h / l image of G' = sg G_w, one slab per 64 output atoms
__global__ void gf_image_kernel(const double* __restrict__ gram, int k, int kpad, int nq, const GfScalars* __restrict__ sc,
                                uint8_t* __restrict__ image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (slab, input atom, pair of output atoms)
  if (idx >= nq * kpad * 32) return;
  const int jp = idx & 31, i = (idx >> 5) % kpad, q = idx / (32 * kpad);
  const int col = q * kGfQ + 2 * jp;
  const float sg = sc->sg;
  float v0 = 0.f, v1 = 0.f;
  if (i < k) {
    if (col < k) v0 = (float)(gram[(size_t)i * k + col] * (double)sg);
    if (col + 1 < k) v1 = (float)(gram[(size_t)i * k + col + 1] * (double)sg);
  }
  uint32_t wh, wl;
  gf_split2(v0, v1, wh, wl);
  const uint32_t piece = (uint32_t)kpad * 128u;
  uint8_t* slab = image + (size_t)q * 2u * piece;
  const uint32_t off = sw128_offset((uint32_t)i, (uint32_t)jp * 4u);
  *reinterpret_cast<uint32_t*>(slab + off) = wh;
  *reinterpret_cast<uint32_t*>(slab + piece + off) = wl;
}

// s_r = sx_r / sw with sx_r the power of two that puts max |x_r| in [64, 128); one warp per row
__global__ void gf_rowscale_kernel(const float* __restrict__ x, int64_t n, int d, const GfScalars* __restrict__ sc,
                                   float* __restrict__ row_scale) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n) return;
  const int lane = threadIdx.x & 31;
  float mm = 0.f;
  for (int c = lane; c < d; c += 32) mm = fmaxf(mm, fabsf(__ldg(x + r * d + c)));
  for (int o = 16; o > 0; o >>= 1) mm = fmaxf(mm, __shfl_xor_sync(0xffffffffu, mm, o));
  if (lane == 0) {
    float sxr = 1.f;
    if (mm > 0.f && mm < 3.0e38f) {
      const int be = 260 - (int)((__float_as_uint(mm) >> 23) & 0xFFu);
      sxr = __uint_as_float((uint32_t)min(max(be, 1), 254) << 23);
    }
    row_scale[r] = sxr * sc->isw;
  }
}

// momentum coefficients: python floats of ista.py:77-78, 98-101 (t_0 = 1; beta_i = (t_i - 1) / t_{i+1}), applied to
// the extrapolation that opens iteration i
__global__ void gf_beta_kernel(float* __restrict__ beta, int iters, int fast) {
  double t = 1.0;
  for (int i = 0; i < iters; ++i) {
    double b = 0.0;
    if (fast && i > 0) {
      const double t_next = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;
      b = (t - 1.0) / t_next;
      t = t_next;
    }
    beta[i] = (float)b;
  }
}

struct GfState {
  uint8_t* image = nullptr;
  double* gram = nullptr;
  GfScalars* scal = nullptr;
  unsigned* maxbits = nullptr;
  int* flag = nullptr;
  float* row_scale = nullptr;
  int64_t row_cap = 0;
  float* b = nullptr;
  size_t b_cap = 0;
  float* beta = nullptr;
  int beta_cap = 0;
  int beta_fast = -1;        // which table `beta` holds (-1: none); every table is a prefix of a longer one
  int num_sms = 0;
  long long* trace_host = nullptr;
  long long* trace_dev = nullptr;
};
GfState g_gf[64];

}  // namespace

bool fista_gram_supported(int64_t n, int d, int k) {
  // worth it only where the two-GEMM tcgen05 kernels do not reach (d > 128) and the pieces of y fit TMEM
  return n >= 1 && d > 128 && d <= 4096 && k >= 4 && k <= kGfKMax && (k % 4) == 0 &&
         n < ((int64_t)1 << 31) - kGfTileM;
}

// z_a holds z_0; runs ALL a.maxiter iterations in one launch (z_i ends up in (i even ? z_a : z_b), caller units
// throughout) and records hist[i] = sum |z_i - z_{i+1}| when a.record; the batch-global stop test is the caller's
// (cabi.cu replays the run with the iteration count the records give, like the resident path).
// *fell_back = 1 when an iterate left the fp16 operand range (the buffers are then unspecified).
// Synchronises the stream.
int fista_gram_run(const FistaArgs& a, int* fell_back, cudaStream_t st) {
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  GfState& S = g_gf[dev];
  const int nq = (a.k + kGfQ - 1) / kGfQ, nks = (a.k + 15) / 16, kpad = nks * 16;
  if (!S.image) {
    LASSO_CUDA_TRY(cudaMalloc(&S.image, (size_t)(kGfKMax / kGfQ) * kGfStage));
    LASSO_CUDA_TRY(cudaMalloc(&S.gram, sizeof(double) * (size_t)kGfKMax * kGfKMax));
    LASSO_CUDA_TRY(cudaMalloc(&S.scal, sizeof(GfScalars)));
    LASSO_CUDA_TRY(cudaMalloc(&S.maxbits, 2 * sizeof(unsigned)));
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
  const size_t b_bytes = sizeof(float) * (size_t)a.n * (size_t)a.k;
  if (b_bytes > S.b_cap) {
    if (S.b) LASSO_CUDA_TRY(cudaFree(S.b));
    S.b = nullptr;
    S.b_cap = 0;
    LASSO_CUDA_TRY(cudaMalloc(&S.b, b_bytes));
    S.b_cap = b_bytes;
  }
  LASSO_CUDA_TRY(cudaFuncSetAttribute((const void*)fista_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)kGfSmem));
  // once per solve: G_w = W^T W (float64), its scaled fp16 image, the row scales, b = x W
  small_gram_launch(a.w, a.d, a.k, a.k, a.d, 0, S.gram, st);
  LASSO_CHECK_LAUNCH();
  LASSO_CUDA_TRY(cudaMemsetAsync(S.maxbits, 0, 2 * sizeof(unsigned), st));
  gf_max_kernel<<<64, 256, 0, st>>>(a.w, a.d * a.k, S.gram, a.k * a.k, S.maxbits);
  LASSO_CHECK_LAUNCH();
  gf_setup_kernel<<<1, 1, 0, st>>>(S.maxbits, a.lr, S.scal, S.flag);
  LASSO_CHECK_LAUNCH();
  gf_image_kernel<<<(nq * kpad * 32 + 255) / 256, 256, 0, st>>>(S.gram, a.k, kpad, nq, S.scal, S.image);
  LASSO_CHECK_LAUNCH();
  gf_rowscale_kernel<<<(unsigned)((a.n + 7) / 8), 256, 0, st>>>(a.x, a.n, a.d, S.scal, S.row_scale);
  LASSO_CHECK_LAUNCH();
  count_launch(4);
  int rc;
  if ((rc = matmul_run(a.x, a.w, a.n, a.d, a.k, S.b, st))) return rc;

  if (!S.trace_host && getenv("LASSO_B200_GRAM_TRACE")) {
    LASSO_CUDA_TRY(cudaHostAlloc((void**)&S.trace_host, 128 * sizeof(long long), cudaHostAllocMapped));
    LASSO_CUDA_TRY(cudaHostGetDevicePointer((void**)&S.trace_dev, S.trace_host, 0));
  }
  float limit = 32768.0f;
  if (const char* lim = getenv("LASSO_B200_RES_LIMIT")) limit = (float)atof(lim);   // tests: force the fallback
  // a batch of less than one wave is cut into as many tiles as there are SMs (fewer rows per M = 128 MMA: the
  // tensor pipe is not the bound, the per-row loads / stores and the serial slab chain are)
  int tile_rows = kGfTileM;
  // (no rounding of the tile height: rounded up to a multiple of 8, n = 10000 gave 139 tiles of 72 rows and 9 idle SMs)
  if (a.n < (int64_t)kGfTileM * S.num_sms) tile_rows = (int)std::max<int64_t>(8, (a.n + S.num_sms - 1) / S.num_sms);
  if (const char* tr = getenv("LASSO_B200_GRAM_ROWS")) tile_rows = std::min(kGfTileM, std::max(8, atoi(tr)));
  const int64_t ntiles = (a.n + tile_rows - 1) / tile_rows;
  const unsigned grid = (unsigned)(ntiles < S.num_sms ? ntiles : S.num_sms);
  if (a.maxiter > S.beta_cap) {
    if (S.beta) LASSO_CUDA_TRY(cudaFree(S.beta));
    S.beta = nullptr;
    S.beta_cap = 0;
    const int cap = std::max(a.maxiter, 1024);
    LASSO_CUDA_TRY(cudaMalloc(&S.beta, sizeof(float) * (size_t)cap));
    S.beta_cap = cap;
    S.beta_fast = -1;
  }
  if (S.beta_fast != (a.fast ? 1 : 0)) {      // once per device: a serial chain of float64 roots (0.27 us per entry)
    gf_beta_kernel<<<1, 1, 0, st>>>(S.beta, S.beta_cap, a.fast ? 1 : 0);
    LASSO_CHECK_LAUNCH();
    S.beta_fast = a.fast ? 1 : 0;
  }
  GfParams p{};
  p.image = S.image;
  p.b = S.b;
  p.row_scale = S.row_scale;
  p.z_a = a.z_a;
  p.z_b = a.z_b;
  p.beta = S.beta;
  p.hist = (a.hist != nullptr && a.record) ? a.hist : nullptr;
  p.iters = a.maxiter;
  p.n = a.n;
  p.k = a.k;
  p.nks = nks;
  p.nq = nq;
  p.tile_rows = tile_rows;
  p.piece_bytes = (uint32_t)kpad * 128u;
  p.lam = a.lam;
  p.limit = limit;
  p.scal = S.scal;
  p.flag = S.flag;
  p.trace = S.trace_dev;
  fista_gram_kernel<<<grid, kGfThreads, kGfSmem, st>>>(p);
  LASSO_CHECK_LAUNCH();
  count_launch(2);
  int flag = 0;
  cudaError_t e = cudaMemcpyAsync(&flag, S.flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error("Gram-form tcgen05 kernel failed: %s", cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  if (S.trace_host) {     // time line of the last launch, CTA 0, cycles since kernel entry
    const long long t0 = S.trace_host[0];
    const char* names[4] = {"full", "gfree", "issued", "refill"};
    fprintf(stderr, "gram trace (last iteration of CTA 0): issuer pready %lld end %lld | compute start %lld pready %lld end %lld\n",
            S.trace_host[2] - t0, S.trace_host[3] - t0, S.trace_host[64 + 1] - t0, S.trace_host[64 + 2] - t0,
            S.trace_host[64 + 3] - t0);
    for (int c = 0; c < nq; ++c) {
      fprintf(stderr, "  slab %d issuer:", c);
      for (int e = 0; e < 4; ++e) fprintf(stderr, " %s %lld", names[e], S.trace_host[4 + 4 * c + e] - t0);
      fprintf(stderr, " | compute: ready %lld gfull %lld drained %lld stored %lld\n", S.trace_host[64 + 4 + 4 * c] - t0,
              S.trace_host[64 + 5 + 4 * c] - t0, S.trace_host[64 + 6 + 4 * c] - t0, S.trace_host[64 + 7 + 4 * c] - t0);
    }
  }
  *fell_back = flag;
  return LASSO_B200_OK;
}

}  // namespace lasso
