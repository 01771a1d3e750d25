// tcgen05 rate probe (B200): how many cycles one tcgen05.mma of the FISTA kernel's shapes
// really costs, and how fast compute warps can read / write TMEM next to it.  The resident
// FISTA kernel (csrc/fista_res.cu) was budgeted from these numbers.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/mma_rate tools/mma_rate.cu
//
//   T1  one thread issues `reps` MMAs back to back (M=128, K=16, 16-bit operands), A from
//       TMEM (TS) or shared memory (SS), N = 64 / 128 / 256  -> cycles per MMA
//   T2  W warps stream tcgen05.ld / tcgen05.st (32x32b.x32)        -> cycles per 4 KB
//   T3  T1 and T2 at the same time                                  -> interference
//   T4  fp16 operands (a_format = b_format = F16): exactness of a K=64 product
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>

#include "../pytorch-lasso_b200/csrc/sm100_ptx.cuh"

using namespace sm100;

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

struct RateArgs {
  long long* out;   // [0] mma cycles, [1] ld/st cycles (max over warps), [2] status
  int n;            // MMA N
  int a_in_tmem;
  int mma_reps;     // 0: no MMA stream
  int ldst_warps;   // 0: no TMEM traffic from compute warps (<= 16)
  int ldst_reps;
  int do_store;     // 1: tcgen05.st instead of ld
  int b_mn_major;
  int m;            // MMA M (128, or 64: half the TMEM lanes)
};

// warp 0: MMA issuer; warps 1..16: TMEM readers / writers (lane quadrant = warp % 4)
__global__ void __launch_bounds__(544) rate_kernel(RateArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // zero operands (timing does not depend on the values)
  for (int i = tid; i < (64 * 1024) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  long long my_cycles = 0;
  if (warp == 0) {
    if (p.mma_reps > 0) {
      const uint32_t idesc = make_idesc(kFmtBF16, p.m, p.n, 0, p.b_mn_major);
      const uint32_t sa = smem_u32(smem), sb = smem_u32(smem) + 16384;
      const uint64_t da = make_smem_desc_sw128(sa, 0, 1024);
      const uint64_t db = p.b_mn_major ? make_smem_desc_sw128(sb, 8192, 1024) : make_smem_desc_sw128(sb, 0, 1024);
      __syncwarp();
      const long long t0 = clock64();
      if (elect_one()) {
        for (int r = 0; r < p.mma_reps; ++r) {
          // alternate two accumulators and four operand slots like the real kernel
          const uint32_t t_d = tbase + (r & 1) * 256;
          if (p.a_in_tmem) mma_ts<false>(t_d, tbase + 480 + (r & 3) * 8, db + (r & 3) * 2, idesc, 1);
          else mma_ss<false>(t_d, da + (r & 3) * 2, db + (r & 3) * 2, idesc, 1);
        }
        mma_commit(&bar);
      }
      __syncwarp();
      const bool ok = mbar_wait(&bar, 0);
      my_cycles = clock64() - t0;
      if (lane == 0) {
        p.out[0] = my_cycles;
        if (!ok) p.out[2] = 1;
      }
    }
  } else if (warp - 1 < p.ldst_warps) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = j;
    uint32_t sink = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int r = 0; r < p.ldst_reps; ++r) {
      const uint32_t col = 64 + ((warp >> 2) & 3) * 32 + (r & 1) * 128;   // away from the accumulators' first columns
      if (p.do_store) {
        tmem_st32(tbase + lane_base + col, v);
        if ((r & 3) == 3) tmem_wait_st();
      } else {
        tmem_ld32(tbase + lane_base + col, v);
        if ((r & 3) == 3) {
          tmem_wait_ld();
          sink += v[r & 31];
        }
      }
    }
    if (p.do_store) tmem_wait_st(); else tmem_wait_ld();
    my_cycles = clock64() - t0;
    if (lane == 0) {
      atomicMax((unsigned long long*)&p.out[1], (unsigned long long)my_cycles);
      if (sink == 0xdeadbeef) p.out[3] = sink;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static void rate(const char* name, int n, int a_tmem, int mma_reps, int warps, int ldst_reps, int store,
                 int b_mn, int grid = 1, int m = 128) {
  long long* d_out;
  CK(cudaMalloc(&d_out, 64));
  CK(cudaMemset(d_out, 0, 64));
  RateArgs p{d_out, n, a_tmem, mma_reps, warps, ldst_reps, store, b_mn, m};
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  rate_kernel<<<grid, 544, 65536>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-44s KERNEL ERROR: %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  long long h[4];
  CK(cudaMemcpy(h, d_out, 32, cudaMemcpyDeviceToHost));
  printf("%-44s grid=%3d status=%lld", name, grid, h[2]);
  if (mma_reps) printf("  mma: %7lld cyc / %d = %6.1f cyc/MMA", h[0], mma_reps, (double)h[0] / mma_reps);
  if (warps) printf("  tmem %s: %7lld cyc / %d = %6.1f cyc per x32 (per warp, %d warps)", store ? "st" : "ld", h[1],
                    ldst_reps, (double)h[1] / ldst_reps, warps);
  printf("\n");
  cudaFree(d_out);
}

// ---- T5: tcgen05.ld shapes: cycles per warp-instruction with W warps streaming loads into two
// alternating register sets (no WAW serialisation inside a warp)
#include "tmem_shapes.inc"
template <int kShape>
__global__ void __launch_bounds__(512) shape_kernel(long long* out, int warps, int reps) {
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  if (warp < warps) {
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t a[64], b[64];
    uint32_t sink = 0;
    __syncwarp();
    const long long t0 = clock64();
    for (int r = 0; r < reps; r += 2) {
      const uint32_t t = tbase + lane_base + ((warp >> 2) & 3) * 64;
      if (kShape == 0) { ld_32x32b_x32(t, a); ld_32x32b_x32(t + 32, b); }
      if (kShape == 1) { ld_16x64b_x32(t, a); ld_16x64b_x32(t + 32, b); }
      if (kShape == 2) { ld_16x128b_x16(t, a); ld_16x128b_x16(t + 32, b); }
      if (kShape == 3) { ld_16x256b_x8(t, a); ld_16x256b_x8(t + 32, b); }
      if (kShape == 4) { ld_32x32b_x64(t, a); ld_32x32b_x64(t + 64, b); }
      if (kShape == 5) { ld_16x256b_x16(t, a); ld_16x256b_x16(t + 64, b); }
      if (kShape == 6) { ld_32x32b_x16(t, a); ld_32x32b_x16(t + 16, b); }
      if (kShape == 7) { ld_32x32b_x8(t, a); ld_32x32b_x8(t + 8, b); }
      tmem_wait_ld();
      sink += a[r & 7] + b[r & 7];
    }
    const long long dt = clock64() - t0;
    if (lane == 0) {
      atomicMax((unsigned long long*)&out[0], (unsigned long long)dt);
      if (sink == 0xdeadbeef) out[1] = sink;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

template <int kShape>
static void shape(const char* name, int bytes_per_instr) {
  for (int w : {1, 4, 16}) {
    long long* d_out;
    CK(cudaMalloc(&d_out, 64));
    CK(cudaMemset(d_out, 0, 64));
    const int reps = 512;
    shape_kernel<kShape><<<1, 512>>>(d_out, w, reps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("T5 %-18s KERNEL ERROR: %s\n", name, cudaGetErrorString(e));
      exit(3);
    }
    long long h[2];
    CK(cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost));
    printf("T5 ld %-16s warps=%2d  %7.1f cyc per instr per warp  -> %6.1f B/cyc/SM\n", name, w,
           (double)h[0] / reps, (double)w * bytes_per_instr * reps / (double)h[0]);
    cudaFree(d_out);
  }
}

// ---- T4: fp16 numerics, TS mode, B K-major, N = 64, K = 64 -------------------------------
__global__ void __launch_bounds__(128) fp16_kernel(const __half* a, const __half* b, float* d, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < 64 * 64; e += 128) {
    const int r = e / 64, kk = e % 64;
    *(__half*)(smem + sw128_offset(r, kk * 2)) = b[e];
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    tmem_alloc(&tmem_base_s, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_s;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = ((const uint32_t*)a)[(size_t)row * 32 + c0 + j];
    tmem_st8(tbase + 256 + ((uint32_t)(warp * 32) << 16) + c0, v);
  }
  tmem_wait_st();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(kFmtF16, 128, 64, 0, 0);
    for (int ks = 0; ks < 4; ++ks)
      mma_ts<false>(tbase, tbase + 256 + ks * 8, make_smem_desc_sw128(smem_u32(smem) + ks * 32, 0, 1024), idesc,
                    ks > 0);
    mma_commit(&bar);
  }
  __syncwarp();
  const bool ok = mbar_wait(&bar, 0);
  tc_fence_after();
  if (!ok && lane == 0) atomicExch(status, 1);
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t v[8];
    tmem_ld8(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 8; ++j) d[(size_t)row * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

static void fp16_test() {
  std::vector<__half> a(128 * 64), b(64 * 64);
  std::vector<float> af(a.size()), bf(b.size());
  for (size_t i = 0; i < a.size(); ++i) {
    // mix of normal and sub-normal magnitudes
    float v = ((float)rand() / RAND_MAX * 2.f - 1.f) * ((i % 7 == 0) ? 1e-5f : 1.f);
    a[i] = __float2half(v);
    af[i] = __half2float(a[i]);
  }
  for (size_t i = 0; i < b.size(); ++i) {
    float v = ((float)rand() / RAND_MAX * 2.f - 1.f) * ((i % 5 == 0) ? 3e-5f : 0.25f);
    b[i] = __float2half(v);
    bf[i] = __half2float(b[i]);
  }
  __half *da, *db;
  float* dd;
  int* ds;
  CK(cudaMalloc(&da, a.size() * 2));
  CK(cudaMalloc(&db, b.size() * 2));
  CK(cudaMalloc(&dd, 128 * 64 * 4));
  CK(cudaMalloc(&ds, 4));
  CK(cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(ds, 0, 4));
  CK(cudaFuncSetAttribute(fp16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  fp16_kernel<<<1, 128, 16384>>>(da, db, dd, ds);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("T4 fp16 KERNEL ERROR: %s\n", cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> d(128 * 64);
  int status;
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&status, ds, 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < 64; ++c) {
      double s = 0;
      for (int kk = 0; kk < 64; ++kk) s += (double)af[r * 64 + kk] * bf[c * 64 + kk];
      maxerr = fmax(maxerr, fabs(d[r * 64 + c] - s));
      maxref = fmax(maxref, fabs(s));
    }
  printf("T4 fp16 TS K=64 (with sub-normal operands)   status=%d  |D|max=%.4f  maxerr vs exact=%.3e\n", status,
         maxref, maxerr);
}

int main() {
  srand(7);
  for (int grid : {1, 148}) {
    rate("T1 TS N=64 K-major", 64, 1, 512, 0, 0, 0, 0, grid);
    rate("T1 TS N=64 MN-major", 64, 1, 512, 0, 0, 0, 1, grid);
    rate("T1 TS N=128", 128, 1, 512, 0, 0, 0, 0, grid);
    rate("T1 TS N=256", 256, 1, 512, 0, 0, 0, 0, grid);
    rate("T1 SS N=64", 64, 0, 512, 0, 0, 0, 0, grid);
    rate("T1 SS N=128", 128, 0, 512, 0, 0, 0, 0, grid);
    rate("T1 SS N=256", 256, 0, 512, 0, 0, 0, 0, grid);
  }
  // M = 64: does a half-height MMA cost half?  (it would decide whether two interleaved 64-row tiles per CTA pay)
  rate("T1 TS N=64 M=64", 64, 1, 512, 0, 0, 0, 1, 1, 64);
  rate("T1 TS N=128 M=64", 128, 1, 512, 0, 0, 0, 1, 1, 64);
  rate("T1 SS N=64 M=64", 64, 0, 512, 0, 0, 0, 1, 1, 64);
  for (int w : {1, 4, 8, 16}) rate("T2 tmem ld x32", 64, 1, 0, w, 256, 0, 0);
  for (int w : {1, 4, 8, 16}) rate("T2 tmem st x32", 64, 1, 0, w, 256, 1, 0);
  rate("T3 TS N=64 + 16 warps ld", 64, 1, 512, 16, 256, 0, 0);
  rate("T3 TS N=64 + 8 warps ld", 64, 1, 512, 8, 256, 0, 0);
  rate("T3 TS N=64 + 16 warps st", 64, 1, 512, 16, 256, 1, 0);
  rate("T3 SS N=64 + 16 warps ld", 64, 0, 512, 16, 256, 0, 0);
  fp16_test();
  shape<0>("32x32b.x32", 4096);
  shape<1>("16x64b.x32", 4096);
  shape<2>("16x128b.x16", 4096);
  shape<3>("16x256b.x8", 4096);
  shape<4>("32x32b.x64", 8192);
  shape<5>("16x256b.x16", 8192);
  shape<6>("32x32b.x16", 2048);
  shape<7>("32x32b.x8", 1024);
  printf("rate probe done\n");
  return 0;
}
