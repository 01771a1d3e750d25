// Resident tcgen05 FISTA for sm_100a: ALL iterations of a 128-row tile run on chip.
//
// The streaming kernel (fista_tc.cu) moves n (d + 3k) floats through HBM per iteration and is
// bound by that traffic.  Rows are independent lasso problems (ista.py:71-73, 90), so a tile
// can run every iteration without leaving the SM: per call HBM sees x once, z0 once and the
// final codes once.  What bounds the kernel is then the tensor pipe and the CUDA-core
// epilogues, which is why the operands are split differently here:
//
//   operands  fp32 -> two fp16 pieces  v = h + l  (h = rn16(v), l = rn16(v - h); |v - h - l| <=
//             2^-22 |v|), three tcgen05.mma kind::f16 products per k-step (h h', h l', l h')
//             instead of the six of the bf16x3 split.  fp16 has a narrow exponent range, so
//             every row's problem is rescaled by powers of two first (max|x_r| -> [64,128),
//             max|W| -> [8,16), codes by their ratio): exact, the iterates are bit-for-bit the
//             scaled iterates of the unscaled problem, and the low pieces stay normal numbers.
//             An iterate that would overflow fp16 anyway ends as inf / NaN, raises a flag at
//             the final store and the caller re-runs the batch with the streaming bf16x3 kernel.
//   state     z_i  : shared memory, 128 x 256 fp32, XOR-swizzled rows (128 KB)
//             y_i  : TMEM, fp32 (the momentum point, ista.py:100): 192 columns + 16 registers per thread
//             x    : shared memory 128 x 64 fp32 (32 KB);  dictionary pieces: 64 KB
//   per iteration (one tile; 16 compute warps in two sets of 8, 1 MMA warp):
//     B   r = R - x -> fp16 pieces, written over R (in place)   | GEMM1  R = Y W^T  (TS MMAs,
//     C   z+ = softshrink(y - lr g, lam) (ista.py:90), delta,   |         B = resident W, K-major)
//         y+ = z+ + beta (z+ - z) (ista.py:100), in place,      | GEMM2  G = r W per 64 atoms
//         pieces of y+ -> slot of the chunk's set               |         (same image, MN-major)
//   TMEM columns: y 192 | S 64 | Q 64 | R / r pieces 64 | G0 64 | G1 64 = 512 (see the kernel).
//   The stop test (ista.py:93) cannot be taken mid-run (it is a batch-global sum): every
//   iteration's sum goes to hist[] and the caller replays a shorter run when it fired early.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sm100_ptx.cuh"

namespace lasso {
namespace {

using namespace sm100;

constexpr int kTileM = 128;
constexpr int kDP = 64;            // padded d
constexpr int kKP = 256;           // padded k
constexpr int kThreadsR = 544;     // warps 0..15: compute, warp 16: MMA issuer (the SM's warp arbiter favours
                                   // high warp ids: as warp 0 the issuer starved behind the epilogue math)
// (registers: the SM hands them out per 4 warps, so 17 warps are budgeted as 20 and 96 per thread is
// the ceiling -- a 112-register build fails to launch; setmaxnreg would let the compute warps grow, but
// ptxas 12.9 refuses to allocate this kernel once the instruction is present)
constexpr uint32_t kSlabBytes = kDP * 128;                 // [64 features][128 B] = 64 atoms of one piece
constexpr uint32_t kPieceBytes = (kKP / 64) * kSlabBytes;  // 32 KB
constexpr uint32_t kWBytes = 2 * kPieceBytes;              // 64 KB: h image, l image
constexpr uint32_t kSmemW = 0;
constexpr uint32_t kSmemZ = kSmemW + kWBytes;              // [128][256] fp32, 1 KB rows
constexpr uint32_t kSmemX = kSmemZ + kTileM * kKP * 4;     // [128][64] fp32, 256 B rows
constexpr uint32_t kSmemBytesR = kSmemX + kTileM * kDP * 4;   // 229376

constexpr uint32_t kColY = 0;         // y, fp32: chunks 0, 1 and half of chunks 2, 3 (the rest: registers)
constexpr uint32_t kColS = 192;       // piece slot S: [h 32 cols][l 32 cols] = 64 atoms of y (even chunks)
constexpr uint32_t kColQ = 256;       // piece slot Q (odd chunks)
constexpr uint32_t kColAccR = 320;    // R = Y W^T (GEMM1); phase B overwrites it with the pieces of r, [h 8 | l 8] per k-step
constexpr uint32_t kColAccG = 384;    // G = r W, two buffers of one 64-atom chunk (GEMM2)
constexpr uint32_t kTmemCols = 512;

constexpr float kPieceLimit = 32768.0f;   // |operand| beyond this: fall back (fp16 max 65504)

struct ResParams {
  const uint8_t* w_image;   // kWBytes, scaled fp16 pieces
  const float* x;           // [n][d]
  const float* z0;          // [n][k] or nullptr
  float* z_out;             // [n][k]
  int64_t n;
  int d, k;
  int trows;                // rows per tile (<= 128)
  int ntiles;
  int iters;
  const float* beta;        // [iters] momentum coefficient applied at the END of iteration i
  const ResScalars* scal;
  double* hist;             // [iters] sum |z_i - z_{i+1}| or nullptr
  double* part;             // [iters][gridDim.x] per-CTA partial records (added into hist by res_hist_reduce_kernel)
  int* flag;                // set to 1 when an operand left the fp16 range
  volatile int* dbg;        // host-mapped debug record or nullptr
  unsigned long long* trace;   // LASSO_B200_TRACE: per-warp (clock << 8 | event) log of block 0
  float limit;              // |operand| at which the solve is handed to the streaming kernel
  bool vec_x, vec_z0, vec_z; // rows of x / z0 / z_out are 16-byte aligned (float4 access)
};

#ifndef LASSO_RES_HINT
#define LASSO_RES_HINT 20000   // suspend-time hint of the waits, ns (0 / 1000 / 20000 measure the same)
#endif
// Slow path of a barrier wait, out of line: the first test failed.  A protocol bug must not hang
// the GPU: after 4 s the wait records where it stood (LASSO_B200_DEBUG) and traps.
__device__ __noinline__ void res_wait_spin(uint32_t addr, uint32_t par, volatile int* dbg, int line) {
  uint64_t t0 = 0;
  uint32_t n = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(addr), "r"(par), "r"((unsigned)LASSO_RES_HINT)
        : "memory");
    if (ok) return;
    if ((++n & 255u) != 0) continue;
    const uint64_t now = global_timer_ns();
    if (t0 == 0) {
      t0 = now;
      continue;
    }
    if (now - t0 < 4000000000ull) continue;
    if (dbg) {
      dbg[1] = line; dbg[2] = blockIdx.x; dbg[3] = threadIdx.x; dbg[4] = -1; dbg[5] = (int)par;
      __threadfence_system();
      dbg[0] = 1;
      __threadfence_system();
    }
    __trap();
  }
}
#define RES_WAIT(bar, parity) RES_WAIT_A(smem_u32(bar), parity)
// the same on a precomputed shared-memory address
#define RES_WAIT_A(addr, parity)                                                          \
  do {                                                                                    \
    const uint32_t _addr = (addr), _par = (parity) & 1u;                                  \
    uint32_t _ok;                                                                         \
    asm volatile(                                                                         \
        "{\n\t.reg .pred P;\n\t"                                                         \
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"                        \
        "selp.b32 %0, 1, 0, P;\n\t}\n"                                                     \
        : "=r"(_ok)                                                                       \
        : "r"(_addr), "r"(_par)                                                           \
        : "memory");                                                                      \
    if (!_ok) res_wait_spin(_addr, _par, p.dbg, __LINE__);                                \
  } while (0)

// timeline instrumentation (block 0, lane 0 of every warp, iterations 3 and 4 of its first
// tile): compiled in with -DLASSO_RES_TRACE (LASSO_B200_BUILD_TRACE=1 python build_ext.py),
// enabled at run time with LASSO_B200_TRACE=<file>
#ifdef LASSO_RES_TRACE
#define RTRACE(id)                                                                        \
  do {                                                                                    \
    if (p.trace != nullptr && tr_on && (threadIdx.x & 31) == 0 && tr_n < 126)             \
      p.trace[(threadIdx.x >> 5) * 128 + tr_n++] = ((unsigned long long)clock64() << 8) | (id); \
  } while (0)
#else
#define RTRACE(id) do { (void)tr_on; (void)tr_n; } while (0)
#endif

__device__ __forceinline__ float2 rsub2(float2 a, float2 b) {   // a - b, one rounding each
  return __ffma2_rn(make_float2(-1.f, -1.f), b, a);
}
// fp32 pair -> packed fp16 pieces (element .x in the low half = lower k index).
// h = v truncated to its top 11 significant bits (a mask: integer pipe, and float(h) is known
// without converting back), l = rn16(v - h); |v - h - l| <= 2^-22 |v|.
__device__ __forceinline__ void split2_pair(float2 v, uint32_t& wh, uint32_t& wl) {
  const float2 t = make_float2(__uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u),
                               __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
  const float2 r = rsub2(v, t);                   // exact
  const __half2 h = __floats2half2_rn(t.x, t.y);  // exact unless sub-normal in fp16
  const __half2 l = __floats2half2_rn(r.x, r.y);
  wh = *reinterpret_cast<const uint32_t*>(&h);
  wl = *reinterpret_cast<const uint32_t*>(&l);
}
// byte offset of (row, 16-byte chunk c) in the swizzled z tile (1 KB rows) / x tile (256 B rows)
__device__ __forceinline__ uint32_t z_off(uint32_t row, uint32_t chunk) {
  return row * 1024u + ((chunk ^ (row & 7u)) << 4);
}
__device__ __forceinline__ uint32_t x_off(uint32_t row, uint32_t chunk) {
  return row * 256u + ((chunk ^ (row & 7u)) << 4);
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
// 4 consecutive floats of a row of `cols` entries starting at column c (zero beyond the row);
// vec = rows are 16-byte aligned (cols % 4 == 0 and an aligned base), else element by element
__device__ __forceinline__ float4 load4(const float* __restrict__ rowp, int c, int cols, bool vec) {
  if (vec) return __ldg(reinterpret_cast<const float4*>(rowp + c));
  float4 v;
  v.x = c + 0 < cols ? __ldg(rowp + c + 0) : 0.f;
  v.y = c + 1 < cols ? __ldg(rowp + c + 1) : 0.f;
  v.z = c + 2 < cols ? __ldg(rowp + c + 2) : 0.f;
  v.w = c + 3 < cols ? __ldg(rowp + c + 3) : 0.f;
  return v;
}
__device__ __forceinline__ void store4(float* __restrict__ rowp, int c, int cols, bool vec, float4 v) {
  if (vec) {
    *reinterpret_cast<float4*>(rowp + c) = v;
    return;
  }
  if (c + 0 < cols) rowp[c + 0] = v.x;
  if (c + 1 < cols) rowp[c + 1] = v.y;
  if (c + 2 < cols) rowp[c + 2] = v.z;
  if (c + 3 < cols) rowp[c + 3] = v.w;
}

// Tile load / store live in their own (non-inlined) functions: they run once per tile and solve,
// and kept inline they cost the iteration loop registers (6 % at C2).  They take what they need BY VALUE: a
// reference to the kernel's parameter struct forces the whole struct into local memory, and every p.field in the
// iteration loop then becomes a local-memory load instead of a constant-bank operand.
struct TileIo {
  const float* x;
  const float* z0;
  float* z_out;
  int d, k;
  int vec_x, vec_z0, vec_z;
  float limit;
};
//
// Row r is scaled by sx_r = 2^(6 - exponent(max |x_r|)) (max |x'_r| in [64, 128)); its codes then
// live in units of sx_r / sw.  A row's result depends on that row alone, so any row split of a
// batch gives the same bits.
__device__ __noinline__ void res_load_tile(const TileIo p, uint8_t* xs, uint8_t* zs, float* row_sx,
                                           int64_t row0, int valid, float isw, int ct) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = ct + i * 512, r = idx >> 4, c4 = idx & 15;   // 16 consecutive lanes share a row
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < valid && c4 * 4 < p.d) v = load4(p.x + (row0 + r) * p.d, c4 * 4, p.d, p.vec_x);
    float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    // a non-finite row keeps scale 1, turns into inf / NaN codes and is caught at the store
    float sxr = 1.f;
    if (m > 0.f && m < 3.0e38f) {
      // 2^(6 - e), e = unbiased exponent of m: biased exponent 127 + 6 - (E - 127)
      const int be = 260 - (int)((__float_as_uint(m) >> 23) & 0xFFu);
      sxr = __uint_as_float((uint32_t)min(max(be, 1), 254) << 23);
    }
    v.x *= sxr; v.y *= sxr; v.z *= sxr; v.w *= sxr;
    *reinterpret_cast<float4*>(xs + x_off(r, c4)) = v;
    if (c4 == 0) row_sx[r] = sxr;
  }
  compute_sync();
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int idx = ct + i * 512, r = idx >> 6, c4 = idx & 63;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (p.z0 != nullptr && r < valid && c4 * 4 < p.k) {
      v = load4(p.z0 + (row0 + r) * p.k, c4 * 4, p.k, p.vec_z0);
      const float szr = row_sx[r] * isw;
      v.x *= szr; v.y *= szr; v.z *= szr; v.w *= szr;
    }
    *reinterpret_cast<float4*>(zs + z_off(r, c4)) = v;
  }
  compute_sync();
}

// returns true when a code is inf / NaN / beyond the limit: an operand left the fp16 range
__device__ __noinline__ bool res_store_tile(const TileIo p, const uint8_t* zs, const float* row_sx,
                                            int64_t row0, int valid, float sw, int ct) {
  bool bad = false;
#pragma unroll 4
  for (int i = 0; i < 16; ++i) {
    const int idx = ct + i * 512, r = idx >> 6, c4 = idx & 63;
    if (r < valid && c4 * 4 < p.k) {
      float4 v = *reinterpret_cast<const float4*>(zs + z_off(r, c4));
      bad |= !(fabsf(v.x) < p.limit) || !(fabsf(v.y) < p.limit) || !(fabsf(v.z) < p.limit) || !(fabsf(v.w) < p.limit);
      const float uzr = sw / row_sx[r];
      v.x *= uzr; v.y *= uzr; v.z *= uzr; v.w *= uzr;
      store4(p.z_out + (row0 + r) * p.k, c4 * 4, p.k, p.vec_z, v);
    }
  }
  return bad;
}

// Schedule of one iteration (NQ = 4 chunks of 64 atoms).  The 16 compute warps form two sets of
// 8 (4 TMEM lane quadrants x 2 column halves): set A runs the even chunks, set B the odd ones, a
// thread works on 32 atoms of one row in two sub-steps of 16.  Every SM sub-partition hosts two warps
// of each set, and the sets are out of phase by one GEMM2 chunk, so one set's TMEM / barrier
// latencies hide behind the other set's arithmetic (all 16 warps in lock step left the ALUs idle
// for ~40 % of a chunk):
//
//   MMA warp   GEMM2 q0 | q1 | q2 | q3 | G1' q0 | G1' q1 | G1' q2 | G1' q3 ........ (B) GEMM2 q0 | q1
//   set A          C(0)           |      C(2)           | idle            |  B  |
//   set B               C(1)           |      C(3)           | idle       |  B  |
//
// C(q): z+ = softshrink(y - lr g, lam), stop-test record, y+ = z+ + beta (z+ - z) in place, then
// the fp16 pieces of y+ go to a piece slot, where slice q of the NEXT iteration's GEMM1 (G1')
// picks them up.  The tensor pipe is the busier side (96 MMAs ~ 4.1 k cycles per iteration), so
// its queue must never wait on the epilogue:
//   * G has two buffers, one per set: GEMM2 q+2 starts as soon as C(q) has the accumulator in registers;
//   * the pieces have two slots, one per set (S: even chunks, Q: odd chunks);
//   * the pieces of r overwrite the accumulator R they were computed from (phase B, in place; the
//     MMAs execute in issue order, so G1' q0 of the next iteration rewrites R after GEMM2 q3 has read r);
//   * the last 64 columns this needs come from y: a quarter of y lives in registers (16 per thread:
//     the second sub-step of chunks 2 and 3).
// TMEM columns: y 192 | S 64 | Q 64 | R / r pieces 64 | G0 64 | G1 64 = 512.
// Barriers (each completes once per iteration of a tile; phase = global iteration counter):
//   aready[q]  256 arrivals   pieces of chunk q stored             set q & 1 -> MMA
//   sfree[j]   commit         G1' slice j done: its slot is free for chunk j + 2   MMA -> set j
//   rfull      commit         GEMM1 complete                       MMA -> compute
//   rready     512 arrivals   pieces of r stored                   compute -> MMA
//   gfull[q]   commit         GEMM2 chunk q complete               MMA -> set q & 1
//   gfree[j]   256 arrivals   C(j) has G buffer j in registers     set j -> MMA (GEMM2 j + 2)
// kHist: 0 no stop-test record, 1 record sum |z+ - z| (ista.py:93), 2 record "some z+ != z" (all a
// threshold of exactly 0 needs; cheaper than the sum).  Records go to per-CTA slots part[it][cta] and are
// added into hist[it] in a fixed order by res_hist_reduce_kernel after the launch.
template <int NQ, int kHist>
__global__ void __launch_bounds__(kThreadsR, 1) fista_res_kernel(ResParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar_w, bar_rfull, bar_rready, bar_aready[4], bar_sfree[2], bar_gfull[4], bar_gfree[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float row_sx[kTileM];   // per-row power-of-two scale of x (see the file header)
  __shared__ float hist_s[2][16];    // per-warp stop-test records of the last two iterations

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int my_tiles = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int iters = p.iters;

  if (tid == 0) {
    mbar_init(&bar_w, 1);
    mbar_init(&bar_rfull, 1);
    mbar_init(&bar_rready, 512);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mbar_init(&bar_aready[q], 256);
      mbar_init(&bar_gfull[q], 1);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      mbar_init(&bar_sfree[j], 1);
      mbar_init(&bar_gfree[j], 256);
    }
    fence_mbar_init();
  }
  if (warp == 16) {
    tmem_alloc(&tmem_base_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;

  if (warp == 16) {
    // ===================== MMA issuer (whole warp runs the control flow) =====================
    if (elect_one()) {
      mbar_expect_tx(&bar_w, kWBytes);
      for (uint32_t off = 0; off < kWBytes; off += 16384)
        bulk_load(smem + kSmemW + off, p.w_image + off, 16384, &bar_w);
    }
    __syncwarp();
    const int dsteps = (p.d + 15) >> 4;
    const uint32_t idesc1 = make_idesc(kFmtF16, 128, kDP, 0, 0);   // B K-major  (GEMM1, N = 64 features)
    const uint32_t idesc2 = make_idesc(kFmtF16, 128, 64, 0, 1);    // B MN-major (GEMM2, N = 64 atoms)
    const uint32_t w_addr = smem_u32(smem + kSmemW);
    const uint64_t desc1 = make_smem_desc_sw128(w_addr, 0, 1024);
    const uint64_t desc2 = make_smem_desc_sw128(w_addr, kSlabBytes, 1024);
    const uint32_t d1_lo = (uint32_t)desc1, d1_hi = (uint32_t)(desc1 >> 32);
    const uint32_t d2_lo = (uint32_t)desc2, d2_hi = (uint32_t)(desc2 >> 32);
    constexpr uint32_t kPiece16 = kPieceBytes >> 4;
    // descriptor = (hi, base + offset); the add is opaque to the compiler on purpose: left to itself it
    // hoists ~100 loop-invariant descriptor words out of the iteration loop, spills them, and reloads
    // them from local memory (~200 cycles) in front of every batch of MMAs, which the tensor pipe,
    // whose queue is short, then waits for
    auto make64 = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; };
    auto off = [](uint32_t base, uint32_t o) {
      uint32_t r;
      asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(base), "r"(o));
      return r;
    };
    int tr_n = 0;
    bool tr_on = false;
    // slice q of GEMM1 of pass `tg`: R (+)= Y[:, 64 q ...] W^T, three products per k-step
    // into ONE accumulator (small ones first).  Slice 0 overwrites R, i.e. the pieces of r: it is
    // issued after the last GEMM2 chunk, and the tensor pipe executes in issue order.
    auto gemm1_slice = [&](int q, uint32_t tg) {
      RES_WAIT(&bar_aready[q], tg);
      RTRACE(11);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t t_slot = tbase + ((q & 1) ? kColQ : kColS);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t koff = (uint32_t)q * (kSlabBytes >> 4) + (uint32_t)(ks * 2);
          const uint64_t qh = make64(off(d1_lo, koff), d1_hi);
          const uint64_t ql = make64(off(d1_lo, koff + kPiece16), d1_hi);
          const uint32_t ah = t_slot + ks * 8, al = ah + 32;
          mma_ts<false>(tbase + kColAccR, ah, ql, idesc1, (q > 0 || ks > 0) ? 1u : 0u);
          mma_ts<false>(tbase + kColAccR, al, qh, idesc1, 1);
          mma_ts<false>(tbase + kColAccR, ah, qh, idesc1, 1);
        }
        if (q + 2 < NQ) mma_commit(&bar_sfree[q]);   // the slot may take the pieces of chunk q + 2
        if (q == NQ - 1) mma_commit(&bar_rfull);
      }
      __syncwarp();
      RTRACE(12);
    };
    RES_WAIT(&bar_w, 0);
    uint32_t gi = 0;                   // global iteration counter (over tiles)
    for (int tile = 0; tile < my_tiles; ++tile) {
      // GEMM1 of the tile's first iteration, from the pieces of y_0
#pragma unroll
      for (int q = 0; q < NQ; ++q) gemm1_slice(q, gi);
      for (int it = 0; it < iters; ++it, ++gi) {
        tr_on = blockIdx.x == 0 && tile == 0 && (it == 3 || it == 4);
        const bool more = it + 1 < iters;
        RTRACE(18);
        RES_WAIT(&bar_rready, gi);
        RTRACE(14);
        tc_fence_after();
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          // ---- GEMM2 chunk q: G[q & 1] = r W[:, 64 q ...]; chunk q - 2 must have left the buffer ----
          if (q >= 2) {
            RES_WAIT(&bar_gfree[q - 2], gi);
            tc_fence_after();
          }
          RTRACE(15);
          if (elect_one()) {
            const uint32_t t_acc = tbase + kColAccG + (uint32_t)(q & 1) * 64u;
            const uint32_t t_r = tbase + kColAccR;   // the pieces of r live where R was
            const uint32_t qoff = (uint32_t)q * (kSlabBytes >> 4);
            // small products first (l h', h l'), leading product last
            constexpr int pa[3] = {1, 0, 0}, pb[3] = {0, 1, 0};
            if (dsteps == 4) {
              // d > 48: straight-line issue (the predicated form below costs the issuing thread ~15 cycles
              // more per MMA, which the tensor pipe then waits for)
#pragma unroll
              for (int t = 0; t < 3; ++t) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                  const uint64_t bd = make64(off(d2_lo, qoff + pb[t] * kPiece16 + ks * 128), d2_hi);
                  mma_ts<false>(t_acc, t_r + pa[t] * 8 + ks * 16, bd, idesc2, (t > 0 || ks > 0) ? 1u : 0u);
                }
              }
            } else {
              uint32_t acc_on = 0;
#pragma unroll
              for (int t = 0; t < 3; ++t) {
#pragma unroll
                for (int ks = 0; ks < 3; ++ks) {
                  if (ks < dsteps) {
                    const uint64_t bd = make64(off(d2_lo, qoff + pb[t] * kPiece16 + ks * 128), d2_hi);
                    mma_ts<false>(t_acc, t_r + pa[t] * 8 + ks * 16, bd, idesc2, acc_on);
                    acc_on = 1;
                  }
                }
              }
            }
            mma_commit(&bar_gfull[q]);
          }
          __syncwarp();
          RTRACE(16);
        }
        if (kHist != 0 && it > 0 && lane == 0) {
          // stop-test record of the previous iteration: the compute warps left their partial sums (mode 2:
          // "moved" flags) in shared memory before they arrived on bar_rready.  They go to THIS CTA's slot of
          // the per-iteration record -- 148 CTAs x 200 iterations of atomics / stores on the same few
          // addresses cost 6 % of the kernel -- while the tensor pipe works through the GEMM2 chunks just
          // queued; res_hist_reduce_kernel adds the slots up in a fixed order afterwards.
          // (float adds in four chains: this GPU's float64 adds have a long latency and this thread is the one
          // that feeds the tensor pipe; the 16 partials are float32 anyway)
          float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int wi = 0; wi < 16; ++wi) s4[wi & 3] += hist_s[(it - 1) & 1][wi];
          const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
          // (a reduction without return value: fire and forget, the issuing thread does not wait for the slot)
          if (s != 0.f) atomicAdd(&p.part[(size_t)(it - 1) * gridDim.x + blockIdx.x], (double)s);
        }
        // ---- GEMM1 of the next iteration, slice by slice as the epilogue delivers the pieces ----
        if (more) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) gemm1_slice(q, gi + 1);
        }
      }
    }
  } else if (warp < 16) {
    // ===================== compute warps =====================
    const int quad = warp & 3;                 // TMEM lane quadrant of this warp
    const int wg = warp >> 2;                  // phase B: features [16 wg, 16 wg + 16)
    const int set = wg >> 1;                   // phase C: set 0 runs chunks 0 and 2, set 1 chunks 1 and 3
    const int half = wg & 1;                   //          atoms [32 half, 32 half + 32) of the chunk
    const int row = quad * 32 + lane;          // row inside the tile == TMEM lane
    const int ct = tid;                        // 0..511
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint8_t* zs = smem + kSmemZ;
    uint8_t* xs = smem + kSmemX;
    const ResScalars sc = *p.scal;
    const float2 nlr2 = make_float2(-sc.lr, -sc.lr);
    int tr_n = 0;
    bool tr_on = false;
    bool bad = false;
    uint32_t gi = 0;
    uint32_t yk[16];                           // y of the second sub-step of this set's second chunk (NQ = 4)
#pragma unroll
    for (int j = 0; j < 16; ++j) yk[j] = 0u;
    const uint32_t t_slot = tbase + lane_base + (set ? kColQ : kColS) + half * 16;   // + 8 s (h), + 32 (l)
    const uint32_t t_g = tbase + lane_base + kColAccG + set * 64 + half * 32;        // + 16 s

    // y of (chunk q, sub-step s) lives in registers?  Only with four chunks: 192 columns hold chunks
    // 0, 1 and the first sub-steps of chunks 2, 3.
    auto y_in_regs = [](int qi, int s) { return NQ == 4 && qi == 1 && s == 1; };
    // TMEM column of y of (qi-th chunk of this set, sub-step s)
    auto y_col = [&](int qi, int s) -> uint32_t {
      const int q = set + 2 * qi;
      if (NQ == 4 && qi == 1) return kColY + 128 + set * 32 + half * 16;   // s == 0 only
      return kColY + q * 64 + half * 32 + s * 16;
    };
    // 16 values -> fp16 pieces -> this thread's 8 + 8 words of its set's slot [h 32 cols][l 32 cols]
    auto split_pieces = [&](const uint32_t (&yv)[16], uint32_t (&wh)[8], uint32_t (&wl)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        split2_pair(make_float2(__uint_as_float(yv[2 * j]), __uint_as_float(yv[2 * j + 1])), wh[j], wl[j]);
    };
    auto store_pieces = [&](const uint32_t (&yv)[16], int s) {
      uint32_t wh[8], wl[8];
      split_pieces(yv, wh, wl);
      tmem_st8(t_slot + s * 8, wh);
      tmem_st8(t_slot + s * 8 + 32, wl);
    };
#ifdef LASSO_RES_HOLD
    uint32_t hold_h[8], hold_l[8];
#endif

    for (int tile = 0; tile < my_tiles; ++tile) {
      const int64_t row0 = (int64_t)(blockIdx.x + (int64_t)tile * gridDim.x) * p.trows;
      int valid = p.trows;
      if (row0 + valid > p.n) valid = (int)(p.n - row0);
      // ---------------- load the tile: x and z0, rescaled per row ----------------
      res_load_tile(TileIo{p.x, p.z0, p.z_out, p.d, p.k, p.vec_x, p.vec_z0, p.vec_z, p.limit}, xs, zs, row_sx, row0, valid,
                    sc.isw, ct);
      const float lam = sc.lam * row_sx[row];           // lam sx_r / sw
      const float uz_row = sc.sw / row_sx[row];         // code units -> caller units (exact)
      // ---------------- y_0 = z_0 (ista.py:76) and its pieces ----------------
#pragma unroll
      for (int qi = 0; qi < 2; ++qi) {
        const int q = set + 2 * qi;
        if (q < NQ) {
          // the slot of the set's second chunk is free once slice q - 2 of this pass has read it
          if (qi == 1) RES_WAIT(&bar_sfree[set], gi);
          tc_fence_after();
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            uint32_t yv[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4 z4 = *reinterpret_cast<const float4*>(zs + z_off(row, q * 16 + half * 8 + s * 4 + j));
              yv[4 * j + 0] = __float_as_uint(z4.x);
              yv[4 * j + 1] = __float_as_uint(z4.y);
              yv[4 * j + 2] = __float_as_uint(z4.z);
              yv[4 * j + 3] = __float_as_uint(z4.w);
            }
            if (y_in_regs(qi, s)) {
#pragma unroll
              for (int j = 0; j < 16; ++j) yk[j] = yv[j];
            } else {
              tmem_st16(tbase + lane_base + y_col(qi, s), yv);
            }
            store_pieces(yv, s);
          }
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bar_aready[q]);
        }
      }

      float beta_next = __ldg(p.beta);         // fetched one iteration ahead: a global load costs ~600 cycles
      for (int it = 0; it < iters; ++it) {
        tr_on = blockIdx.x == 0 && tile == 0 && (it == 3 || it == 4);
        const bool more = it + 1 < iters;      // pieces are only needed if another GEMM1 follows
        const float2 beta2 = make_float2(beta_next, beta_next);
        if (more) beta_next = __ldg(p.beta + it + 1);
        // ---------------- phase B: r = R - x -> pieces, in place (16 features per thread) ----------------
        {
          float4 xv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) xv[j] = *reinterpret_cast<const float4*>(xs + x_off(row, wg * 4 + j));
          RES_WAIT(&bar_rfull, gi);
          RTRACE(30);
          tc_fence_after();
          uint32_t rr[16];
          tmem_ld16(tbase + lane_base + kColAccR + wg * 16, rr);
          tmem_wait_ld();
          uint32_t wh[8], wl[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 ra = rsub2(make_float2(__uint_as_float(rr[4 * j + 0]), __uint_as_float(rr[4 * j + 1])),
                                    make_float2(xv[j].x, xv[j].y));
            const float2 rc = rsub2(make_float2(__uint_as_float(rr[4 * j + 2]), __uint_as_float(rr[4 * j + 3])),
                                    make_float2(xv[j].z, xv[j].w));
            split2_pair(ra, wh[2 * j], wl[2 * j]);
            split2_pair(rc, wh[2 * j + 1], wl[2 * j + 1]);
          }
          // the pieces go where R was, k-step by k-step: the 16 features this thread read become [h 8 cols |
          // l 8 cols] in the very columns it read them from, so no warp ever overwrites what another still
          // has to load (the MMA takes the A operand of every k-step from its own TMEM address anyway)
          const uint32_t t_r = tbase + lane_base + kColAccR + wg * 16;
          tmem_st8(t_r, wh);
          tmem_st8(t_r + 8, wl);
          tmem_wait_st();
          tc_fence_before();
          mbar_arrive(&bar_rready);
          RTRACE(31);
        }
        // ---------------- phase C (+ pieces of the next iteration) ----------------
        float part = 0.f, part_b = 0.f;   // two chains: 64 dependent adds per iteration otherwise
        bool moved_p = false;      // mode 2: a predicate register, not data registers (the kernel sits at its register cap)
#pragma unroll
        for (int qi = 0; qi < 2; ++qi) {
          const int q = set + 2 * qi;
          if (q < NQ) {
#ifdef LASSO_RES_EARLY_DRAIN
            // both halves of this thread's part of G leave the buffer at once, so that GEMM2 q + 2 can
            // start ~600 cycles earlier (costs 16 live registers during the first sub-step)
            uint32_t g2[2][16];
            RES_WAIT(&bar_gfull[q], gi);
            RTRACE(40);
            tc_fence_after();
            tmem_ld16(t_g, g2[0]);
            tmem_ld16(t_g + 16, g2[1]);
            tmem_wait_ld();
            if (q + 2 < NQ) {
              tc_fence_before();
              mbar_arrive(&bar_gfree[set]);
            }
#endif
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              // y does not depend on the MMAs: its load is in flight while this thread waits for G
              uint32_t yv[16];
              if (y_in_regs(qi, s)) {
#pragma unroll
                for (int j = 0; j < 16; ++j) yv[j] = yk[j];
              } else {
                tmem_ld16(tbase + lane_base + y_col(qi, s), yv);
              }
#ifdef LASSO_RES_EARLY_DRAIN
              uint32_t (&g)[16] = g2[s];
              if (!y_in_regs(qi, s)) tmem_wait_ld();
#else
              uint32_t g[16];
              if (s == 0) {
                RES_WAIT(&bar_gfull[q], gi);
                RTRACE(40);
                tc_fence_after();
              }
              tmem_ld16(t_g + s * 16, g);
              tmem_wait_ld();
              if (s == 1 && q + 2 < NQ) {
                tc_fence_before();
                mbar_arrive(&bar_gfree[set]);   // accumulator is in registers: hand the buffer to chunk q + 2
              }
#endif
              RTRACE(42);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                uint8_t* zp = zs + z_off(row, q * 16 + half * 8 + s * 4 + j);
                float4 z4 = *reinterpret_cast<const float4*>(zp);
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                  const float2 yy = make_float2(__uint_as_float(yv[4 * j + 2 * h2]), __uint_as_float(yv[4 * j + 2 * h2 + 1]));
                  const float2 gg = make_float2(__uint_as_float(g[4 * j + 2 * h2]), __uint_as_float(g[4 * j + 2 * h2 + 1]));
                  const float2 zz = h2 ? make_float2(z4.z, z4.w) : make_float2(z4.x, z4.y);
                  // softshrink(y - lr g, lam) (ista.py:90): v - clamp(v, +-lam) is bit-identical to the
                  // three-way select; the step itself is one fused multiply-add
                  const float2 v = __ffma2_rn(nlr2, gg, yy);
                  const float2 c = make_float2(fminf(fmaxf(v.x, -lam), lam), fminf(fmaxf(v.y, -lam), lam));
                  const float2 zn = rsub2(v, c);
                  const float2 dl = rsub2(zn, zz);                       // z+ - z
                  if (kHist == 1) {                                      // stop-test sum (ista.py:93)
                    if (h2) part_b += fabsf(dl.x) + fabsf(dl.y);
                    else part += fabsf(dl.x) + fabsf(dl.y);
                  }
                  // mode 2 only asks whether anything moved: one packed multiply-add per pair on the FMA pipe
                  // (sum of squares; a non-zero dl is a difference of O(1) float32 numbers in scaled units,
                  // >= 1e-16, so its square cannot underflow) instead of OR-ing bit patterns on the ALU pipe
                  if (kHist == 2) moved_p = moved_p || (dl.x != 0.f) || (dl.y != 0.f);
                  const float2 yn = __ffma2_rn(beta2, dl, zn);           // ista.py:100
                  if (h2) { z4.z = zn.x; z4.w = zn.y; }
                  else { z4.x = zn.x; z4.y = zn.y; }
                  yv[4 * j + 2 * h2] = __float_as_uint(yn.x);
                  yv[4 * j + 2 * h2 + 1] = __float_as_uint(yn.y);
                }
                *reinterpret_cast<float4*>(zp) = z4;
              }
              if (y_in_regs(qi, s)) {
#pragma unroll
                for (int j = 0; j < 16; ++j) yk[j] = yv[j];
              } else {
                tmem_st16(tbase + lane_base + y_col(qi, s), yv);
              }
              RTRACE(41);
              if (more) {
                // The set's first chunk finds its slot free (GEMM1 of the previous pass completed:
                // everybody saw bar_rfull).  The second one has to wait for slice q - 2 of this pass,
                // which the tensor pipe reaches only after the last GEMM2 chunk: its first sub-step
                // keeps its pieces in registers (LASSO_RES_HOLD) or re-reads y+ from TMEM afterwards
                // instead of stalling in the middle of the chunk.
                if (qi == 0) {
                  store_pieces(yv, s);
                } else if (s == 0) {
#ifdef LASSO_RES_HOLD
                  split_pieces(yv, hold_h, hold_l);
#endif
                } else {
                  RES_WAIT(&bar_sfree[set], gi + 1u);
                  tc_fence_after();
                  RTRACE(22);
                  store_pieces(yv, 1);
#ifdef LASSO_RES_HOLD
                  tmem_st8(t_slot, hold_h);
                  tmem_st8(t_slot + 32, hold_l);
#else
                  tmem_wait_st();          // y+ of sub-step 0 is in TMEM (never the register-resident part)
                  tmem_ld16(tbase + lane_base + y_col(1, 0), yv);
                  tmem_wait_ld();
                  store_pieces(yv, 0);
#endif
                }
              }
            }
            tmem_wait_st();            // pieces and y
            if (more) {
              tc_fence_before();
              mbar_arrive(&bar_aready[q]);
              RTRACE(23);
            }
          }
        }
        if (kHist != 0) {
          float s;
          if (kHist == 1) {
            s = (part + part_b) * uz_row;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          } else {
            s = __any_sync(0xffffffffu, moved_p) ? 1.f : 0.f;
          }
          // (lane and warp index are re-read here: kept across the iteration they cost two of the 96 registers and
          // the record variants spilled loop-carried values into the iteration's critical path)
          unsigned tid_now;
          asm volatile("mov.u32 %0, %%tid.x;" : "=r"(tid_now));
          if ((tid_now & 31u) == 0u) {
            // the MMA warp adds the 16 partial records up after the next bar_rready; nobody comes after a
            // tile's last iteration, so there the warps add to the CTA's slot themselves
            if (more) hist_s[it & 1][tid_now >> 5] = s;
            else if (s != 0.f) atomicAdd(&p.part[(size_t)it * gridDim.x + blockIdx.x], (double)s);
          }
        }
        ++gi;
      }
      // ---------------- store the codes ----------------
      compute_sync();
      bad |= res_store_tile(TileIo{p.x, p.z0, p.z_out, p.d, p.k, p.vec_x, p.vec_z0, p.vec_z, p.limit}, zs, row_sx, row0,
                            valid, sc.sw, ct);
      compute_sync();
    }
    if (bad) atomicExch(p.flag, 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tbase, kTmemCols);
}

// ---- set-up kernels ---------------------------------------------------------------------
// momentum table: python floats of ista.py:77-78, 98-101 (t0 = 1; beta_i = (t_i - 1) / t_{i+1}).  A serial chain of
// float64 square roots and divisions (software sequences on this GPU: 55 us for 200 entries, 2 % of a config-2 solve
// when it ran inside res_setup_kernel), but it depends on nothing but `fast` and every table is a prefix of a longer
// one: computed once per device up to the buffer's capacity and kept.
__global__ void res_beta_kernel(float* __restrict__ beta, int count, int fast) {
  double t = 1.0;
  for (int i = 0; i < count; ++i) {
    const double t_next = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;
    beta[i] = fast ? (float)((t - 1.0) / t_next) : 0.f;
    t = t_next;
  }
}

// scale of the dictionary (power of two, max |W'| in [8, 16)), scaled step / threshold, flag reset.  One block.
__global__ void res_setup_kernel(const float* __restrict__ w, int nw, float lr, float lam,
                                 ResScalars* __restrict__ sc, int* __restrict__ flag) {
  __shared__ unsigned s_max;
  if (threadIdx.x == 0) s_max = 0;
  __syncthreads();
  unsigned mw = 0;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) mw = max(mw, __float_as_uint(w[i]) & 0x7FFFFFFFu);
  mw = __reduce_max_sync(0xffffffffu, mw);
  if ((threadIdx.x & 31) == 0 && mw) atomicMax(&s_max, mw);
  __syncthreads();
  if (threadIdx.x != 0) return;
  const float aw = __uint_as_float(s_max);   // NaN / inf have the largest bit patterns
  int bad = 0, ew = 0;
  if (!(aw < 3.0e38f)) bad = 1;
  if (aw > 0.f && !bad) ew = 3 - ilogbf(aw);
  ew = max(-40, min(40, ew));
  sc->sw = ldexpf(1.f, ew);
  sc->isw = ldexpf(1.f, -ew);
  sc->lr = ldexpf(lr, -2 * ew);
  sc->lam = ldexpf(lam, -ew);
  if (!(sc->lr > 0.f) || !(sc->lr < 3.0e38f) || !(sc->lam < 3.0e38f) || (lam > 0.f && !(sc->lam > 0.f))) bad = 1;
  sc->bad = bad;
  *flag = bad;
}

// dictionary [d][k] fp32 -> two scaled fp16 piece images, each [k/64 slabs][64 rows][128 B] with
// the 128-byte swizzle; zero padded to d = 64, k = 256
__global__ void res_prep_w_kernel(const float* __restrict__ w, int d, int k, const ResScalars* __restrict__ sc,
                                  uint8_t* __restrict__ image) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= kDP * kKP) return;
  const int i = idx / kKP, j = idx % kKP;
  const float v = (i < d && j < k) ? w[(int64_t)i * k + j] * sc->sw : 0.f;
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn(v - __half2float(h));
  const uint32_t off = (uint32_t)(j / 64) * kSlabBytes + sw128_offset(i, (j % 64) * 2);
  *reinterpret_cast<__half*>(image + off) = h;
  *reinterpret_cast<__half*>(image + kPieceBytes + off) = l;
}

// hist[it] += sum over the CTAs' slots (fixed order: the sums are reproducible run to run); slots cleared
// for the next launch.  One warp per iteration.
__global__ void res_hist_reduce_kernel(double* __restrict__ part, int slots, int iters, double* __restrict__ hist) {
  const int it = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (it >= iters) return;
  double* row = part + (size_t)it * slots;
  double s = 0.0;
  for (int c = lane; c < slots; c += 32) {
    s += row[c];
    row[c] = 0.0;
  }
  s = warp_sum(s);
  if (lane == 0 && s != 0.0) hist[it] += s;
}

struct ResState {
  double* part = nullptr;
  size_t part_cap = 0;
  uint8_t* w_image = nullptr;
  ResScalars* scal = nullptr;
  int* flag = nullptr;
  float* beta = nullptr;
  int beta_cap = 0;
  int beta_fast = -1;       // which table S.beta holds (-1: none)
  int num_sms = 0;
  bool attr_set = false;
  int* dbg_host = nullptr;
  int* dbg_dev = nullptr;
  unsigned long long* trace = nullptr;
};
ResState g_res[64];

}  // namespace

bool fista_res_supported(int64_t n, int d, int k) {
  // any d <= 64, k <= 256: the dictionary image is zero padded and unaligned rows of x / z are
  // read and written element by element (only at tile load / store, once per solve)
  return n >= 1 && d >= 1 && d <= kDP && k >= 1 && k <= kKP && n < (int64_t)1 << 31;
}

// ---- host side: prepare (dictionary image, scalars, momentum table) / launch (a row range) /
// finish (read the hand-over flag; the only synchronisation) -------------------------------------
int fista_res_prepare(const float* w, int d, int k, float lr, float lam, int iters, int fast, cudaStream_t st) {
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  ResState& S = g_res[dev];
  if (!S.w_image) {
    LASSO_CUDA_TRY(cudaMalloc(&S.w_image, kWBytes));
    LASSO_CUDA_TRY(cudaMalloc(&S.scal, sizeof(ResScalars)));
    LASSO_CUDA_TRY(cudaMalloc(&S.flag, sizeof(int)));
    LASSO_CUDA_TRY(cudaDeviceGetAttribute(&S.num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (iters > S.beta_cap) {
    if (S.beta) LASSO_CUDA_TRY(cudaFree(S.beta));
    S.beta = nullptr;
    S.beta_cap = 0;
    const int cap = iters < 1024 ? 1024 : iters;
    LASSO_CUDA_TRY(cudaMalloc(&S.beta, sizeof(float) * (size_t)cap));
    S.beta_cap = cap;
    S.beta_fast = -1;
  }
  if (S.beta_fast != (fast ? 1 : 0)) {        // (calls on one device are serialised and stream-ordered by the lease)
    res_beta_kernel<<<1, 1, 0, st>>>(S.beta, S.beta_cap, fast ? 1 : 0);
    LASSO_CHECK_LAUNCH();
    count_launch();
    S.beta_fast = fast ? 1 : 0;
  }
  if (!S.dbg_host && getenv("LASSO_B200_DEBUG")) {
    LASSO_CUDA_TRY(cudaHostAlloc((void**)&S.dbg_host, 4096, cudaHostAllocMapped));
    memset(S.dbg_host, 0, 4096);
    LASSO_CUDA_TRY(cudaHostGetDevicePointer((void**)&S.dbg_dev, S.dbg_host, 0));
  }
  const char* trace_path = getenv("LASSO_B200_TRACE");
  if (trace_path && !S.trace) LASSO_CUDA_TRY(cudaMalloc(&S.trace, 32 * 128 * 8));
  if (S.trace) LASSO_CUDA_TRY(cudaMemsetAsync(S.trace, 0, 32 * 128 * 8, st));
  if (!S.attr_set) {
    const void* kernels[12] = {
        (const void*)fista_res_kernel<1, 0>, (const void*)fista_res_kernel<2, 0>, (const void*)fista_res_kernel<3, 0>,
        (const void*)fista_res_kernel<4, 0>, (const void*)fista_res_kernel<1, 1>, (const void*)fista_res_kernel<2, 1>,
        (const void*)fista_res_kernel<3, 1>, (const void*)fista_res_kernel<4, 1>, (const void*)fista_res_kernel<1, 2>,
        (const void*)fista_res_kernel<2, 2>, (const void*)fista_res_kernel<3, 2>, (const void*)fista_res_kernel<4, 2>};
    for (const void* kfn : kernels)
      LASSO_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytesR));
    S.attr_set = true;
  }
  res_setup_kernel<<<1, 256, 0, st>>>(w, d * k, lr, lam, S.scal, S.flag);
  LASSO_CHECK_LAUNCH();
  res_prep_w_kernel<<<(kDP * kKP + 255) / 256, 256, 0, st>>>(w, d, k, S.scal, S.w_image);
  LASSO_CHECK_LAUNCH();
  count_launch(2);
  return LASSO_B200_OK;
}

// rows per tile: smallest multiple of 8 that keeps the number of waves of 128-row tiles
static int64_t res_tile_rows(int64_t n, int num_sms) {
  const int64_t slots = num_sms;
  const int64_t waves = ((n + kTileM - 1) / kTileM + slots - 1) / slots;
  int64_t trows = (n + waves * slots - 1) / (waves * slots);
  trows = ((trows + 7) / 8) * 8;
  return trows > kTileM ? kTileM : trows;
}

// rows one wave of tiles covers (one tile per SM) when a batch of n rows is cut for the host
// pipeline of lasso_b200_fista_f32_host
int64_t fista_res_wave_rows(int64_t n) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return res_tile_rows(n, sms) * sms;
}

int64_t fista_res_tile_rows(int64_t n) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return res_tile_rows(n, sms);
}

// iterations on rows [0, n) of x / z0 / z_out (pointers to the first of these rows); hist is shared
// by all launches of a solve.  tile_rows = 0 picks the tile height from n.  No synchronisation.
int fista_res_launch(const float* x, const float* z0, float* z_out, int64_t n, int d, int k, int iters,
                     double* hist, int hist_mode, int64_t tile_rows, cudaStream_t st) {
  if (hist == nullptr) hist_mode = 0;
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  ResState& S = g_res[dev];
  const int64_t trows = tile_rows > 0 ? tile_rows : res_tile_rows(n, S.num_sms);
  const int64_t ntiles = (n + trows - 1) / trows;
  const unsigned grid = (unsigned)(ntiles < S.num_sms ? ntiles : S.num_sms);
  const char* trace_path = getenv("LASSO_B200_TRACE");

  ResParams p{};
  p.w_image = S.w_image;
  p.x = x;
  p.z0 = z0;
  p.z_out = z_out;
  p.n = n;
  p.d = d;
  p.k = k;
  p.trows = (int)trows;
  p.ntiles = (int)ntiles;
  p.iters = iters;
  p.beta = S.beta;
  p.scal = S.scal;
  p.hist = hist;
  p.part = nullptr;
  if (hist_mode != 0) {
    // per-CTA slots of the stop-test record, [iters][grid] doubles, all zero between launches
    const size_t need = (size_t)iters * S.num_sms;
    if (need > S.part_cap) {
      if (S.part) LASSO_CUDA_TRY(cudaFree(S.part));
      S.part = nullptr;
      S.part_cap = 0;
      LASSO_CUDA_TRY(cudaMalloc(&S.part, sizeof(double) * need));
      LASSO_CUDA_TRY(cudaMemsetAsync(S.part, 0, sizeof(double) * need, st));
      S.part_cap = need;
    }
    p.part = S.part;
  }
  p.flag = S.flag;
  p.dbg = S.dbg_dev;
  p.trace = trace_path ? S.trace : nullptr;
  p.vec_x = (d % 4) == 0 && ((uintptr_t)x % 16) == 0;
  p.vec_z0 = (k % 4) == 0 && ((uintptr_t)z0 % 16) == 0;
  p.vec_z = (k % 4) == 0 && ((uintptr_t)z_out % 16) == 0;
  p.limit = kPieceLimit;
  if (const char* lim = getenv("LASSO_B200_RES_LIMIT")) p.limit = (float)atof(lim);   // tests: force the fallback
#define LASSO_RES_LAUNCH(NQ_)                                                               \
  do {                                                                                      \
    if (hist_mode == 0) fista_res_kernel<NQ_, 0><<<grid, kThreadsR, kSmemBytesR, st>>>(p);  \
    else if (hist_mode == 1) fista_res_kernel<NQ_, 1><<<grid, kThreadsR, kSmemBytesR, st>>>(p); \
    else fista_res_kernel<NQ_, 2><<<grid, kThreadsR, kSmemBytesR, st>>>(p);                 \
  } while (0)
  switch ((k + 63) / 64) {   // 64-atom chunks
    case 1: LASSO_RES_LAUNCH(1); break;
    case 2: LASSO_RES_LAUNCH(2); break;
    case 3: LASSO_RES_LAUNCH(3); break;
    default: LASSO_RES_LAUNCH(4); break;
  }
#undef LASSO_RES_LAUNCH
  LASSO_CHECK_LAUNCH();
  count_launch();
  if (hist_mode != 0) {
    res_hist_reduce_kernel<<<(iters + 7) / 8, 256, 0, st>>>(S.part, (int)grid, iters, hist);
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  return LASSO_B200_OK;
}

// *fell_back = 1 when an operand left the fp16 range in any launch since the last prepare (the
// codes are then unspecified and the caller must use another path).  Synchronises the stream.
int fista_res_finish(int* fell_back, cudaStream_t st) {
  int dev = 0;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  ResState& S = g_res[dev];
  const char* trace_path = getenv("LASSO_B200_TRACE");
  int flag = 0;
  cudaError_t e = cudaMemcpyAsync(&flag, S.flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (S.dbg_host && S.dbg_host[0]) {
    set_error("resident kernel barrier timeout: line %d block %d thread %d parity %d (%s)", S.dbg_host[1],
              S.dbg_host[2], S.dbg_host[3], S.dbg_host[5], cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  if (e != cudaSuccess) {
    set_error("resident kernel failed: %s", cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  if (S.trace && trace_path) {
    static unsigned long long host_trace[32 * 128];
    LASSO_CUDA_TRY(cudaMemcpy(host_trace, S.trace, sizeof(host_trace), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen(trace_path, "w")) {
      for (int w = 0; w < 32; ++w)
        for (int i = 0; i < 128 && host_trace[w * 128 + i]; ++i)
          fprintf(f, "%d %llu %llu\n", w, host_trace[w * 128 + i] >> 8, host_trace[w * 128 + i] & 255);
      fclose(f);
    }
  }
  *fell_back = flag;
  return LASSO_B200_OK;
}

// Runs `iters` iterations from z0 (nullptr = zeros) into z_out.  hist_mode: 0 none, 1 hist[it] =
// sum |z_it - z_it+1|, 2 hist[it] > 0 iff z_it+1 != z_it (enough for a threshold of exactly 0).
int fista_res_run(const float* x, const float* w, const float* z0, float* z_out, int64_t n, int d, int k,
                  float lr, float lam, int iters, int fast, double* hist, int hist_mode, int* fell_back,
                  cudaStream_t st) {
  int rc = fista_res_prepare(w, d, k, lr, lam, iters, fast, st);
  if (rc) return rc;
  if ((rc = fista_res_launch(x, z0, z_out, n, d, k, iters, hist, hist_mode, 0, st))) return rc;
  return fista_res_finish(fell_back, st);
}

}  // namespace lasso
