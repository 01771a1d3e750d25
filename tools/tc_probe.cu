// tcgen05 probe: validates, on a real B200, every hardware convention the FISTA
// tensor-core kernel relies on, before that kernel is written around them.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tc_probe tools/tc_probe.cu
//
// Experiments (one CTA each; all waits are bounded so a wrong guess cannot hang the GPU):
//   E1  bf16  SS   A[128xK] K-major SW128, B[NxK] K-major SW128
//   E2  bf16  SS   B given N-contiguous ([K][N] in memory) -> MN-major SW128 descriptor
//   E3  bf16  TS   A from TMEM (two bf16 per 32-bit column, written with tcgen05.st 32x32b)
//   E4  bf16  TS   long accumulation (same tiles re-issued 64x): rounding of the fp32 accumulate
//   E5  tf32  SS   operand conversion fp32->tf32 (truncate vs round) with full-mantissa inputs
//   E6  tf32  TS   A fp32 from TMEM
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>

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

struct ProbeArgs {
  const void* a;   // A row-major [128][K] (bf16 or fp32)
  const void* b;   // B: K-major: [N][K]; MN-major: [K][N]
  float* d;        // D [128][N]
  int* status;     // 0 ok, 1 barrier timeout
  int n;           // N (multiple of 16, <= 256)
  int kblocks;     // number of 128-byte K blocks (bf16: 64 elems, tf32: 32 elems)
  int b_mn_major;  // 1: B is N-contiguous
  int a_in_tmem;   // 1: TS form
  int repeats;     // re-issue the whole K loop this many times (accumulating)
};

// Layout of dynamic smem: A tile blocks [kblocks][128 rows][128 B], then B tile blocks.
//   K-major B: [kblocks][N rows][128 B]
//   MN-major B (N-contiguous): [N/atomN blocks][K rows][128 B], atomN = 128 B of N-elements
template <bool kTf32>
__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  constexpr int ES = kTf32 ? 4 : 2;          // element size
  constexpr int KB = 128 / ES;               // elements per 128-byte block
  constexpr int UK = 32 / ES;                // K per MMA (32 bytes)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = p.kblocks * KB;
  uint8_t* sa = smem;
  uint8_t* sb = smem + (size_t)p.kblocks * 128 * 128;

  // ---- fill shared-memory images (generic proxy) ----
  if (!p.a_in_tmem) {
    for (int e = tid; e < 128 * K; e += 128) {
      const int r = e / K, kk = e % K;
      const int blk = kk / KB, kin = kk % KB;
      const uint32_t off = blk * (128 * 128) + sw128_offset(r, kin * ES);
      if (kTf32) *(float*)(sa + off) = ((const float*)p.a)[e];
      else *(__nv_bfloat16*)(sa + off) = ((const __nv_bfloat16*)p.a)[e];
    }
  }
  if (!p.b_mn_major) {
    for (int e = tid; e < p.n * K; e += 128) {
      const int r = e / K, kk = e % K;
      const int blk = kk / KB, kin = kk % KB;
      const uint32_t off = blk * (p.n * 128) + sw128_offset(r, kin * ES);
      if (kTf32) *(float*)(sb + off) = ((const float*)p.b)[e];
      else *(__nv_bfloat16*)(sb + off) = ((const __nv_bfloat16*)p.b)[e];
    }
  } else {
    // memory [K][N]; image: N-blocks of KB elements, each [K rows][128 B]
    for (int e = tid; e < K * p.n; e += 128) {
      const int kk = e / p.n, nn = e % p.n;
      const int nblk = nn / KB, nin = nn % KB;
      const uint32_t off = nblk * (K * 128) + sw128_offset(kk, nin * ES);
      if (kTf32) *(float*)(sb + off) = ((const float*)p.b)[e];
      else *(__nv_bfloat16*)(sb + off) = ((const __nv_bfloat16*)p.b)[e];
    }
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
  const uint32_t t_d = tbase;            // accumulator: columns [0, N)
  const uint32_t t_a = tbase + 256;      // A operand:   columns [256, ...)

  if (p.a_in_tmem) {
    // thread <-> row; 32-bit column c holds K elements (bf16: 2c, 2c+1; tf32: c)
    const int row = warp * 32 + lane;
    const int cols = K * ES / 4;
    for (int c0 = 0; c0 < cols; c0 += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = ((const uint32_t*)p.a)[(size_t)row * cols + c0 + j];
      tmem_st8(t_a + ((uint32_t)(warp * 32) << 16) + c0, v);
    }
    tmem_wait_st();
    tc_fence_before();
  }
  __syncthreads();

  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc(kTf32 ? kFmtTF32 : kFmtBF16, 128, p.n, 0, p.b_mn_major);
    uint32_t acc = 0;
    for (int rep = 0; rep < p.repeats; ++rep) {
      for (int kk = 0; kk < K; kk += UK) {
        const int blk = kk / KB, kin = kk % KB;
        uint64_t db;
        if (!p.b_mn_major) {
          // K-major: 8-row groups 1024 B apart; step inside the 128 B row by 32 B per MMA
          db = make_smem_desc_sw128(smem_u32(sb) + blk * (p.n * 128) + kin * ES, 0, 1024);
        } else {
          // MN-major: rows are K; 8-row groups 1024 B apart (SBO); N-blocks K*128 B apart (LBO)
          db = make_smem_desc_sw128(smem_u32(sb) + kk * 128, (uint32_t)K * 128, 1024);
        }
        if (p.a_in_tmem) {
          mma_ts<kTf32>(t_d, t_a + kk * ES / 4, db, idesc, acc);
        } else {
          const uint64_t da =
              make_smem_desc_sw128(smem_u32(sa) + blk * (128 * 128) + kin * ES, 0, 1024);
          mma_ss<kTf32>(t_d, da, db, idesc, acc);
        }
        acc = 1;
      }
    }
    mma_commit(&bar);
  }
  __syncwarp();
  const bool ok = mbar_wait(&bar, 0);
  tc_fence_after();
  if (!ok) {
    if (lane == 0) atomicExch(p.status, 1);
  } else {
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < p.n; c0 += 8) {
      uint32_t v[8];
      tmem_ld8(t_d + ((uint32_t)(warp * 32) << 16) + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 8; ++j) p.d[(size_t)row * p.n + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 512);
}

// ------------------------------------------------------------------ host side
static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }
static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float tf32_rn(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u += 0x00000FFFu + ((u >> 13) & 1u);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}
static float frand() { return (float)rand() / RAND_MAX * 2.f - 1.f; }

template <bool kTf32>
static void run(const char* name, int n, int kblocks, int b_mn, int a_tmem, int repeats,
                bool positive = false) {
  constexpr int ES = kTf32 ? 4 : 2;
  const int K = kblocks * (128 / ES);
  std::vector<float> a(128 * K), b((size_t)n * K);
  for (auto& v : a) v = positive ? 0.5f + 0.5f * fabsf(frand()) : frand();
  for (auto& v : b) v = positive ? 0.5f + 0.5f * fabsf(frand()) : frand();
  if (!kTf32) {
    for (auto& v : a) v = bf16_round(v);
    for (auto& v : b) v = bf16_round(v);
  }
  // b is logically B[n][k]; device memory order depends on b_mn
  std::vector<uint8_t> ha(128 * (size_t)K * ES), hb((size_t)n * K * ES);
  for (int i = 0; i < 128 * K; ++i) {
    if (kTf32) ((float*)ha.data())[i] = a[i];
    else ((__nv_bfloat16*)ha.data())[i] = __float2bfloat16(a[i]);
  }
  for (int nn = 0; nn < n; ++nn)
    for (int kk = 0; kk < K; ++kk) {
      const size_t idx = b_mn ? (size_t)kk * n + nn : (size_t)nn * K + kk;
      if (kTf32) ((float*)hb.data())[idx] = b[(size_t)nn * K + kk];
      else ((__nv_bfloat16*)hb.data())[idx] = __float2bfloat16(b[(size_t)nn * K + kk]);
    }
  void *da, *db;
  float* dd;
  int* ds;
  CK(cudaMalloc(&da, ha.size()));
  CK(cudaMalloc(&db, hb.size()));
  CK(cudaMalloc(&dd, 128 * n * sizeof(float)));
  CK(cudaMalloc(&ds, sizeof(int)));
  CK(cudaMemcpy(da, ha.data(), ha.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dd, 0xFF, 128 * n * sizeof(float)));
  CK(cudaMemset(ds, 0, sizeof(int)));
  ProbeArgs p{da, db, dd, ds, n, kblocks, b_mn, a_tmem, repeats};
  const size_t smem = (size_t)kblocks * 128 * 128 + (size_t)n * K * ES + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<kTf32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                          (int)smem));
  probe_kernel<kTf32><<<1, 128, smem>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%-28s KERNEL ERROR: %s\n", name, cudaGetErrorString(e));
    exit(3);
  }
  std::vector<float> d(128 * n);
  int status = 0;
  CK(cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&status, ds, 4, cudaMemcpyDeviceToHost));
  // references in double: exact inputs, truncated-to-tf32 inputs, rn-to-tf32 inputs
  double err_exact = 0, err_trunc = 0, err_rn = 0, ref_max = 0, bias = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < n; ++c) {
      double s0 = 0, s1 = 0, s2 = 0;
      for (int kk = 0; kk < K; ++kk) {
        const float av = a[(size_t)r * K + kk], bv = b[(size_t)c * K + kk];
        s0 += (double)av * bv;
        if (kTf32) {
          s1 += (double)tf32_trunc(av) * tf32_trunc(bv);
          s2 += (double)tf32_rn(av) * tf32_rn(bv);
        }
      }
      s0 *= repeats; s1 *= repeats; s2 *= repeats;
      const double got = d[(size_t)r * n + c];
      err_exact = fmax(err_exact, fabs(got - s0));
      err_trunc = fmax(err_trunc, fabs(got - s1));
      err_rn = fmax(err_rn, fabs(got - s2));
      ref_max = fmax(ref_max, fabs(s0));
      bias += (got - s0) / (fabs(s0) > 1e-30 ? fabs(s0) : 1.0);
    }
  bias /= (128.0 * n);
  printf("%-28s status=%d  N=%3d K=%4d rep=%2d  |D|max=%9.3f  maxerr exact=%.3e", name, status, n,
         K, repeats, ref_max, err_exact);
  if (kTf32) printf("  trunc=%.3e  rn=%.3e", err_trunc, err_rn);
  printf("  mean signed rel err=%.3e\n", bias);
  cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(ds);
}

int main() {
  srand(1234);
  run<false>("E1 bf16 SS K-major", 64, 1, 0, 0, 1);
  run<false>("E1b bf16 SS K-major N=256 K=256", 256, 4, 0, 0, 1);
  run<false>("E2 bf16 SS B MN-major", 64, 1, 1, 0, 1);
  run<false>("E2b bf16 SS B MN-major N=256", 256, 1, 1, 0, 1);
  run<false>("E2c bf16 SS B MN N=128 K=128", 128, 2, 1, 0, 1);
  run<false>("E3 bf16 TS A in TMEM", 64, 1, 0, 1, 1);
  run<false>("E3b bf16 TS K=256", 64, 4, 0, 1, 1);
  run<false>("E3c bf16 TS + B MN-major", 64, 1, 1, 1, 1);
  run<false>("E4 bf16 TS accumulate x64", 64, 4, 0, 1, 64, true);
  run<false>("E4b bf16 SS accumulate x64", 64, 4, 0, 0, 64, true);
  run<true>("E5 tf32 SS K-major", 64, 1, 0, 0, 1);
  run<true>("E5b tf32 SS K=128", 64, 4, 0, 0, 1);
  run<true>("E6 tf32 TS A in TMEM", 64, 2, 0, 1, 1);
  printf("probe done\n");
  return 0;
}
