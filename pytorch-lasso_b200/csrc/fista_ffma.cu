// CUDA-core (FFMA) FISTA step: one launch per iteration, any n, d, k.
//
// This is the exact-fp32 path of the engine and the fallback for shapes the
// tcgen05 kernel does not take.  Per 64-row tile (TM = 64/32/16 by d):
//
//   phase 1  R = Y W^T - X          Y = z_cur + beta (z_cur - z_prev) recomputed
//                                    from the two code buffers (ista.py:72, 100)
//   phase 2  G = R W                 (ista.py:73)
//   epilogue z_next = S(Y - lr G)    written over z_prev; sum|z_cur - z_next|
//                                    goes to hist[iter]           (ista.py:90, 93)
//
// HBM traffic per iteration: n (d + 3k) floats (Y is never stored).  The
// dictionary is re-read per tile from L2.  Both GEMM phases run a 4x4
// register-blocked FFMA tile from padded shared memory.
#include <cstdlib>

#include "common.cuh"

namespace lasso {
namespace {

constexpr int kThreads = 256;
constexpr int kBlk = 64;   // output block width (d-block in phase 1, atom chunk in phase 2)
constexpr int kCk = 32;    // contraction chunk staged in shared memory
constexpr int kPad = 4;    // row padding (floats) -> conflict-free 128-bit LDS

__device__ __forceinline__ float4 ld4(const float* base, int64_t row, int col, int64_t nrows,
                                      int ncols, int ld, bool vec_ok) {
  // guarded, zero-padded load of base[row, col..col+3]
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= nrows || col >= ncols) return v;
  const float* p = base + row * (int64_t)ld + col;
  if (vec_ok && col + 3 < ncols) return *reinterpret_cast<const float4*>(p);
  v.x = p[0];
  if (col + 1 < ncols) v.y = p[1];
  if (col + 2 < ncols) v.z = p[2];
  if (col + 3 < ncols) v.w = p[3];
  return v;
}

__device__ __forceinline__ void st4(float* base, int64_t row, int col, int64_t nrows, int ncols,
                                    int ld, bool vec_ok, float4 v) {
  if (row >= nrows || col >= ncols) return;
  float* p = base + row * (int64_t)ld + col;
  if (vec_ok && col + 3 < ncols) {
    *reinterpret_cast<float4*>(p) = v;
    return;
  }
  p[0] = v.x;
  if (col + 1 < ncols) p[1] = v.y;
  if (col + 2 < ncols) p[2] = v.z;
  if (col + 3 < ncols) p[3] = v.w;
}

struct StepParams {
  const float* x;
  const float* w;
  const float* z_cur;
  float* z_io;  // holds z_prev on entry, z_next on exit
  int64_t n;
  int d, k;
  float lr, lam, beta;
  int use_prev;  // 0: Y = z_cur (first iteration or plain ISTA)
  int mode;      // 0: FISTA step, 1: loss terms (sum r^2, sum |z|) into out2
  double* out2;
  const float* aux;  // MODE 3: gradient at z_cur
  float* out;        // MODE 3: candidate code
  StepCtl ctl;
};

// MODE 0 = iteration step, MODE 1 = loss terms only,
// MODE 2 = gradient: z_io <- (z_cur W^T - x) W, out2[0] += sum r^2            (ista.py:22-24)
// MODE 3 = line-search trial: cand = softshrink(z_cur - lr * aux, lam) -> out,  (ista.py:40)
//          out2 += { sum (cand W^T - x)^2, sum |cand|, sum dz * aux, sum dz^2 } (ista.py:26-35)
// BLK = width of an output block (d-block in phase 1, atom chunk in phase 2); BLK * 4 threads, each a
// (TM / 16) x 4 register tile.  BLK = 128 (512 threads) halves the number of block steps of a tile (tuning
// variant, see use_wide_blocks).
template <int TM, int MODE, int BLK>
__global__ void __launch_bounds__(BLK * 4, BLK == 128 ? 1 : (TM == 64 ? 2 : 3)) fista_ffma_kernel(StepParams p) {
  constexpr int NT = BLK * 4;     // threads
  constexpr int CT = BLK / 4;     // threads across an output block (4 columns each)
  constexpr int RPT = TM / 16;  // rows per thread
  constexpr int A_IT = (TM * (kCk / 4) + NT - 1) / NT;   // float4 per thread of a Y chunk
  constexpr int W_IT = BLK * (kCk / 4) / NT;                  // float4 per thread of a W chunk
  extern __shared__ __align__(16) float smem[];
  const int d_pad = (p.d + 3) & ~3;
  const int r_ld = d_pad + kPad;
  float* Rs = smem;                                   // [TM][r_ld]
  float* As = Rs + TM * r_ld;                         // [TM][kCk + kPad]   (Y chunk)
  float* Ws = As + TM * (kCk + kPad);                 // max([BLK][kCk+kPad], [kCk][BLK+kPad])
  __shared__ double red[NT / 32][4];

  if (MODE == 0 && p.ctl.tol_abs >= 0.0 && p.ctl.iter >= 1 &&
      p.ctl.hist[p.ctl.iter - 1] <= p.ctl.tol_abs)
    return;  // an earlier iteration met the stop test (ista.py:93-95)

  const int tid = threadIdx.x;
  const int tx = tid % CT, ty = tid / CT;
  const bool kvec = (p.k & 3) == 0;
  const bool dvec = (p.d & 3) == 0;
  const int64_t ntiles = (p.n + TM - 1) / TM;
  double acc_a = 0.0, acc_b = 0.0;  // MODE 0: delta ; MODE 1/2/3: sum r^2, sum |z|
  double acc_c = 0.0, acc_d = 0.0;  // MODE 3: sum dz*g, sum dz^2

  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t row0 = tile * TM;

    // ---------------- phase 1: R = Y W^T - X ------------------------------
    // The Y and W chunks of step j0 + kCk are fetched into registers (fetch1) before the products of
    // step j0 are formed, and written to shared memory (commit1) after them, so the global / L2 latency
    // of the staging hides behind the FFMA block of the same CTA.
    float4 ra[A_IT], rp[A_IT], rg[A_IT], rw[W_IT];
    auto fetch1 = [&](int db, int j0) {
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        const int e = tid + it * NT;
        if (e < TM * (kCk / 4)) {
          const int r = e / (kCk / 4), c4 = (e % (kCk / 4)) * 4;
          ra[it] = ld4(p.z_cur, row0 + r, j0 + c4, p.n, p.k, p.k, kvec);
          if (MODE == 0 && p.use_prev) rp[it] = ld4(p.z_io, row0 + r, j0 + c4, p.n, p.k, p.k, kvec);
          if (MODE == 3) rg[it] = ld4(p.aux, row0 + r, j0 + c4, p.n, p.k, p.k, kvec);
        }
      }
#pragma unroll
      for (int it = 0; it < W_IT; ++it) {
        const int e = tid + it * NT;
        const int i = e / (kCk / 4), c4 = (e % (kCk / 4)) * 4;
        rw[it] = ld4(p.w, db + i, j0 + c4, p.d, p.k, p.k, kvec);
      }
    };
    auto commit1 = [&](int db, int j0) {
#pragma unroll
      for (int it = 0; it < A_IT; ++it) {
        const int e = tid + it * NT;
        if (e < TM * (kCk / 4)) {
          const int r = e / (kCk / 4), c4 = (e % (kCk / 4)) * 4;
          float4 zc = ra[it];
          if (MODE == 0 && p.use_prev) {
            const float4 zp = rp[it];
            zc.x = momentum_point(zc.x, zp.x, p.beta);
            zc.y = momentum_point(zc.y, zp.y, p.beta);
            zc.z = momentum_point(zc.z, zp.z, p.beta);
            zc.w = momentum_point(zc.w, zp.w, p.beta);
          }
          if (MODE == 3) {
            const float4 g = rg[it];
            float4 cd;
            cd.x = ista_update(zc.x, g.x, p.lr, p.lam);
            cd.y = ista_update(zc.y, g.y, p.lr, p.lam);
            cd.z = ista_update(zc.z, g.z, p.lr, p.lam);
            cd.w = ista_update(zc.w, g.w, p.lr, p.lam);
            if (db == 0) {
              st4(p.out, row0 + r, j0 + c4, p.n, p.k, p.k, kvec, cd);
              const float dx = __fsub_rn(cd.x, zc.x), dy = __fsub_rn(cd.y, zc.y);
              const float dz = __fsub_rn(cd.z, zc.z), dw = __fsub_rn(cd.w, zc.w);
              acc_c += (double)dx * g.x + (double)dy * g.y + (double)dz * g.z + (double)dw * g.w;
              acc_d += (double)dx * dx + (double)dy * dy + (double)dz * dz + (double)dw * dw;
            }
            zc = cd;
          }
          if ((MODE == 1 || MODE == 3) && db == 0)
            acc_b += (double)(fabsf(zc.x) + fabsf(zc.y)) + (double)(fabsf(zc.z) + fabsf(zc.w));
          *reinterpret_cast<float4*>(&As[r * (kCk + kPad) + c4]) = zc;
        }
      }
#pragma unroll
      for (int it = 0; it < W_IT; ++it) {
        const int e = tid + it * NT;
        const int i = e / (kCk / 4), c4 = (e % (kCk / 4)) * 4;
        *reinterpret_cast<float4*>(&Ws[i * (kCk + kPad) + c4]) = rw[it];
      }
    };
    for (int db = 0; db < p.d; db += BLK) {
      float acc[RPT][4];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

      fetch1(db, 0);
      for (int j0 = 0; j0 < p.k; j0 += kCk) {
        __syncthreads();
        commit1(db, j0);
        __syncthreads();
        if (j0 + kCk < p.k) fetch1(db, j0 + kCk);
#pragma unroll
        for (int jj = 0; jj < kCk; jj += 4) {
          float4 a[RPT], b[4];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
            a[r] = *reinterpret_cast<const float4*>(&As[(ty * RPT + r) * (kCk + kPad) + jj]);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            b[c] = *reinterpret_cast<const float4*>(&Ws[(tx + CT * c) * (kCk + kPad) + jj]);
#pragma unroll
          for (int r = 0; r < RPT; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              acc[r][c] = fmaf(a[r].x, b[c].x, acc[r][c]);
              acc[r][c] = fmaf(a[r].y, b[c].y, acc[r][c]);
              acc[r][c] = fmaf(a[r].z, b[c].z, acc[r][c]);
              acc[r][c] = fmaf(a[r].w, b[c].w, acc[r][c]);
            }
        }
      }
      // residual block -> shared (zero outside the problem)
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int lr_ = ty * RPT + r, i = db + tx + CT * c;
          const int64_t gr = row0 + lr_;
          float v = 0.f;
          if (gr < p.n && i < p.d) v = __fsub_rn(acc[r][c], p.x[gr * p.d + i]);
          if (i < d_pad) Rs[lr_ * r_ld + i] = v;
          if (MODE != 0) acc_a += (double)v * (double)v;
        }
    }
    if (MODE == 1 || MODE == 3) continue;

    // ---------------- phase 2: G = R W, fused update ----------------------
    auto fetch2 = [&](int j0, int i0) {   // W chunk [kCk][BLK] (output-contiguous), one step ahead
#pragma unroll
      for (int it = 0; it < W_IT; ++it) {
        const int e = tid + it * NT;
        const int ii = e / (BLK / 4), c4 = (e % (BLK / 4)) * 4;
        rw[it] = ld4(p.w, i0 + ii, j0 + c4, p.d, p.k, p.k, kvec);
      }
    };
    auto commit2 = [&]() {
#pragma unroll
      for (int it = 0; it < W_IT; ++it) {
        const int e = tid + it * NT;
        const int ii = e / (BLK / 4), c4 = (e % (BLK / 4)) * 4;
        *reinterpret_cast<float4*>(&Ws[ii * (BLK + kPad) + c4]) = rw[it];
      }
    };
    for (int j0 = 0; j0 < p.k; j0 += BLK) {
      float acc[RPT][4];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

      fetch2(j0, 0);
      for (int i0 = 0; i0 < d_pad; i0 += kCk) {
        __syncthreads();
        commit2();
        __syncthreads();
        if (i0 + kCk < d_pad) fetch2(j0, i0 + kCk);
        const int imax = min(kCk, d_pad - i0);
        for (int ii = 0; ii < imax; ii += 4) {
          float4 a[RPT], b[4];
#pragma unroll
          for (int r = 0; r < RPT; ++r)
            a[r] = *reinterpret_cast<const float4*>(&Rs[(ty * RPT + r) * r_ld + i0 + ii]);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            b[q] = *reinterpret_cast<const float4*>(&Ws[(ii + q) * (BLK + kPad) + tx * 4]);
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            acc[r][0] = fmaf(a[r].x, b[0].x, acc[r][0]);
            acc[r][1] = fmaf(a[r].x, b[0].y, acc[r][1]);
            acc[r][2] = fmaf(a[r].x, b[0].z, acc[r][2]);
            acc[r][3] = fmaf(a[r].x, b[0].w, acc[r][3]);
            acc[r][0] = fmaf(a[r].y, b[1].x, acc[r][0]);
            acc[r][1] = fmaf(a[r].y, b[1].y, acc[r][1]);
            acc[r][2] = fmaf(a[r].y, b[1].z, acc[r][2]);
            acc[r][3] = fmaf(a[r].y, b[1].w, acc[r][3]);
            acc[r][0] = fmaf(a[r].z, b[2].x, acc[r][0]);
            acc[r][1] = fmaf(a[r].z, b[2].y, acc[r][1]);
            acc[r][2] = fmaf(a[r].z, b[2].z, acc[r][2]);
            acc[r][3] = fmaf(a[r].z, b[2].w, acc[r][3]);
            acc[r][0] = fmaf(a[r].w, b[3].x, acc[r][0]);
            acc[r][1] = fmaf(a[r].w, b[3].y, acc[r][1]);
            acc[r][2] = fmaf(a[r].w, b[3].z, acc[r][2]);
            acc[r][3] = fmaf(a[r].w, b[3].w, acc[r][3]);
          }
        }
      }
      // epilogue: y recomputed from the code buffers, shrink, store over z_prev
#pragma unroll
      for (int r = 0; r < RPT; ++r) {
        const int64_t gr = row0 + ty * RPT + r;
        const int col = j0 + tx * 4;
        if (gr >= p.n || col >= p.k) continue;
        if (MODE == 2) {   // gradient only
          st4(p.z_io, gr, col, p.n, p.k, p.k, kvec,
              make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]));
          continue;
        }
        float4 zc = ld4(p.z_cur, gr, col, p.n, p.k, p.k, kvec);
        float4 y = zc;
        if (p.use_prev) {
          float4 zp = ld4(p.z_io, gr, col, p.n, p.k, p.k, kvec);
          y.x = momentum_point(zc.x, zp.x, p.beta);
          y.y = momentum_point(zc.y, zp.y, p.beta);
          y.z = momentum_point(zc.z, zp.z, p.beta);
          y.w = momentum_point(zc.w, zp.w, p.beta);
        }
        float4 zn;
        zn.x = ista_update(y.x, acc[r][0], p.lr, p.lam);
        zn.y = ista_update(y.y, acc[r][1], p.lr, p.lam);
        zn.z = ista_update(y.z, acc[r][2], p.lr, p.lam);
        zn.w = ista_update(y.w, acc[r][3], p.lr, p.lam);
        st4(p.z_io, gr, col, p.n, p.k, p.k, kvec, zn);
        float dsum = fabsf(__fsub_rn(zc.x, zn.x));
        if (col + 1 < p.k) dsum += fabsf(__fsub_rn(zc.y, zn.y));
        if (col + 2 < p.k) dsum += fabsf(__fsub_rn(zc.z, zn.z));
        if (col + 3 < p.k) dsum += fabsf(__fsub_rn(zc.w, zn.w));
        acc_a += (double)dsum;
      }
    }
    (void)dvec;
  }

  // block reduction -> one double atomic per CTA and output
  acc_a = warp_sum(acc_a);
  acc_b = warp_sum(acc_b);
  acc_c = warp_sum(acc_c);
  acc_d = warp_sum(acc_d);
  if ((tid & 31) == 0) {
    red[tid >> 5][0] = acc_a;
    red[tid >> 5][1] = acc_b;
    red[tid >> 5][2] = acc_c;
    red[tid >> 5][3] = acc_d;
  }
  __syncthreads();
  if (tid == 0) {
    double sa = 0.0, sb = 0.0, sc = 0.0, sd = 0.0;
    for (int wi = 0; wi < NT / 32; ++wi) {
      sa += red[wi][0];
      sb += red[wi][1];
      sc += red[wi][2];
      sd += red[wi][3];
    }
    if (MODE == 0) {
      if (p.ctl.hist) atomicAdd(&p.ctl.hist[p.ctl.iter], sa);
    } else {
      atomicAdd(&p.out2[0], sa);
      if (MODE != 2) atomicAdd(&p.out2[1], sb);
      if (MODE == 3) {
        atomicAdd(&p.out2[2], sc);
        atomicAdd(&p.out2[3], sd);
      }
    }
  }
}

// y = z_next + beta (z_next - z) (ista.py:100) and delta = sum |z - z_next| (ista.py:93)
__global__ void momentum_kernel(const float* __restrict__ z_next, const float* __restrict__ z,
                                float beta, float* __restrict__ y, int64_t count,
                                double* __restrict__ delta) {
  __shared__ double red[8];
  double acc = 0.0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float zn = z_next[i], zo = z[i];
    if (y) y[i] = momentum_point(zn, zo, beta);
    acc += (double)fabsf(__fsub_rn(zo, zn));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) s += red[wi];
    atomicAdd(delta, s);
  }
}

size_t smem_bytes(int tm, int d, int blk = kBlk) {
  const int d_pad = (d + 3) & ~3;
  size_t ws = (size_t)max(blk * (kCk + kPad), kCk * (blk + kPad));
  return sizeof(float) * ((size_t)tm * (d_pad + kPad) + (size_t)tm * (kCk + kPad) + ws);
}

int pick_tm(int d) {
  const size_t budget = 200 * 1024;
  if (const char* e = getenv("LASSO_B200_FFMA_TM")) {   // tuning override
    const int tm = atoi(e);
    if ((tm == 64 || tm == 32 || tm == 16) && smem_bytes(tm, d) <= budget) return tm;
  }
  // 64-row tiles while two CTAs fit one SM (tools/ffma_tm.py: with the register-staged prefetch they beat
  // 32-row tiles by 4-20 % at d = 200 and 289)
  if (smem_bytes(64, d) <= 96 * 1024) return 64;
  if (smem_bytes(32, d) <= budget) return 32;
  if (smem_bytes(16, d) <= budget) return 16;
  return 0;
}

int sm_count() {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms < 1)
      num_sms = 148;
  }
  return num_sms;
}

template <int TM, int MODE, int BLK>
int launch(const StepParams& p, cudaStream_t st) {
  const int num_sms = sm_count();
  const size_t smem = smem_bytes(TM, p.d, BLK);
  // the attribute is per device (context): set it on every launch that needs more than the default
  // 48 KB rather than caching a process-wide flag (one process may drive several GPUs)
  if (smem > 48 * 1024)
    LASSO_CUDA_TRY(cudaFuncSetAttribute(fista_ffma_kernel<TM, MODE, BLK>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  LASSO_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(
      &per_sm, fista_ffma_kernel<TM, MODE, BLK>, BLK * 4, smem));
  if (per_sm < 1) per_sm = 1;
  const int64_t ntiles = (p.n + TM - 1) / TM;
  int64_t grid = (int64_t)num_sms * per_sm;
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  fista_ffma_kernel<TM, MODE, BLK><<<(unsigned)grid, BLK * 4, smem, st>>>(p);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

template <int MODE>
int dispatch(const StepParams& p, cudaStream_t st) {
  switch (pick_tm(p.d)) {
    case 64: return launch<64, MODE, kBlk>(p, st);
    case 32: return launch<32, MODE, kBlk>(p, st);
    case 16: return launch<16, MODE, kBlk>(p, st);
    default:
      set_error("FFMA path: d=%d needs more shared memory than one SM has", p.d);
      return LASSO_B200_ERR_UNSUPPORTED;
  }
}

}  // namespace

// Runs all iterations.  z_0 is in a.z_a; iteration i reads (i even ? z_a : z_b)
// and writes the other buffer.  z_out is unused here (cabi.cu selects the
// result buffer); kept in the signature for symmetry with the tcgen05 runner.
int fista_ffma_run(const FistaArgs& a, float* /*z_out*/, cudaStream_t st) {
  double t = 1.0;
  for (int it = 0; it < a.maxiter; ++it) {
    StepParams p{};
    p.x = a.x;
    p.w = a.w;
    p.z_cur = (it & 1) ? a.z_b : a.z_a;
    p.z_io = (it & 1) ? a.z_a : a.z_b;
    p.n = a.n;
    p.d = a.d;
    p.k = a.k;
    p.lr = a.lr;
    p.lam = a.lam;
    // momentum coefficient that produced y_it, i.e. the one computed at the end
    // of iteration it-1 (ista.py:99-100); doubles on the host like the reference
    double beta = 0.0;
    if (a.fast && it > 0) {
      // t currently holds t_{it-1}; advance
      const double t_next = (1.0 + sqrt(1.0 + 4.0 * t * t)) / 2.0;
      beta = (t - 1.0) / t_next;
      t = t_next;
    }
    p.beta = (float)beta;
    p.use_prev = (a.fast && it > 0) ? 1 : 0;
    p.mode = 0;
    p.out2 = nullptr;
    p.ctl.hist = a.hist;
    p.ctl.tol_abs = a.tol_abs;
    p.ctl.iter = it;
    int rc = dispatch<0>(p, st);
    if (rc != LASSO_B200_OK) return rc;
  }
  return LASSO_B200_OK;
}

int loss_terms_run(const float* x, const float* z, const float* w, int64_t n, int d, int k,
                   double* out, cudaStream_t st) {
  LASSO_CUDA_TRY(cudaMemsetAsync(out, 0, 2 * sizeof(double), st));
  StepParams p{};
  p.x = x;
  p.w = w;
  p.z_cur = z;
  p.z_io = nullptr;
  p.n = n;
  p.d = d;
  p.k = k;
  p.mode = 1;
  p.out2 = out;
  p.ctl.hist = nullptr;
  p.ctl.tol_abs = -1.0;
  p.ctl.iter = 0;
  return dispatch<1>(p, st);
}

int gradient_run(const float* x, const float* point, const float* w, int64_t n, int d, int k,
                 float* grad, double* f_terms, cudaStream_t st) {
  LASSO_CUDA_TRY(cudaMemsetAsync(f_terms, 0, sizeof(double), st));
  StepParams p{};
  p.x = x;
  p.w = w;
  p.z_cur = point;
  p.z_io = grad;
  p.n = n;
  p.d = d;
  p.k = k;
  p.mode = 2;
  p.out2 = f_terms;
  p.ctl.tol_abs = -1.0;
  return dispatch<2>(p, st);
}

int trial_run(const float* x, const float* point, const float* grad, const float* w, int64_t n,
              int d, int k, float step, float lam, float* cand, double* sums4, cudaStream_t st) {
  LASSO_CUDA_TRY(cudaMemsetAsync(sums4, 0, 4 * sizeof(double), st));
  StepParams p{};
  p.x = x;
  p.w = w;
  p.z_cur = point;
  p.aux = grad;
  p.out = cand;
  p.n = n;
  p.d = d;
  p.k = k;
  p.lr = step;
  p.lam = lam;
  p.mode = 3;
  p.out2 = sums4;
  p.ctl.tol_abs = -1.0;
  return dispatch<3>(p, st);
}

int momentum_run(const float* z_next, const float* z, float beta, float* y, int64_t count,
                 double* delta, cudaStream_t st) {
  LASSO_CUDA_TRY(cudaMemsetAsync(delta, 0, sizeof(double), st));
  if (count == 0) return LASSO_B200_OK;
  int blocks = (int)((count + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  momentum_kernel<<<blocks, 256, 0, st>>>(z_next, z, beta, y, count, delta);
  LASSO_CHECK_LAUNCH();
  count_launch();
  return LASSO_B200_OK;
}

}  // namespace lasso
