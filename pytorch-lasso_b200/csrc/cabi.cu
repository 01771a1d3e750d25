// extern "C" entry points declared in include/lasso_b200.h: argument checks,
// private workspace, buffer rotation of the FISTA loop, result selection.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace lasso {

std::atomic<long long> g_launches{0};
std::atomic<long long> g_res_fallbacks{0};   // resident solves redone by the streaming kernel
static thread_local char t_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

namespace {

// ---- grow-only device workspace, one per device ----------------------------
struct Buffer {
  void* ptr = nullptr;
  size_t bytes = 0;
};
struct Workspace {
  Buffer code;     // second code buffer [n,k]
  Buffer hist;     // per-iteration delta sums (double)
  Buffer ctl;      // small scalars (int/double)
  Buffer scratch;  // Lipschitz Gram etc.
  Buffer hx, hw, hz;  // device copies for the _host entry points
  // The workspace (and the per-device dictionary images / flags of the kernels) is shared by every
  // caller of a device.  `mutex` serialises the host side of the entry points; `last_use` is recorded
  // on the stream of the call that used it last, and a call on ANOTHER stream waits for it first, so
  // solves issued from different streams or threads of one device run one after the other instead of
  // overwriting each other's buffers.
  std::mutex mutex;
  cudaEvent_t last_use = nullptr;
  cudaStream_t last_stream = nullptr;
  bool used = false;
};
constexpr int kMaxDevices = 64;
Workspace g_ws[kMaxDevices];
// copy streams / events of the host entry point's pipeline
struct HostPipe {
  cudaStream_t s_in = nullptr, s_out = nullptr;
  std::vector<cudaEvent_t> events;
};
HostPipe g_pipe[kMaxDevices];

int ensure(Buffer& b, size_t bytes) {
  if (bytes <= b.bytes) return LASSO_B200_OK;
  if (b.ptr) {
    cudaError_t e = cudaFree(b.ptr);
    b.ptr = nullptr;
    b.bytes = 0;
    if (e != cudaSuccess) {
      set_error("cudaFree failed: %s", cudaGetErrorString(e));
      return LASSO_B200_ERR_CUDA;
    }
  }
  cudaError_t e = cudaMalloc(&b.ptr, bytes);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return e == cudaErrorMemoryAllocation ? LASSO_B200_ERR_NOMEM : LASSO_B200_ERR_CUDA;
  }
  b.bytes = bytes;
  return LASSO_B200_OK;
}

int current_workspace(Workspace** out) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("no usable CUDA device: %s", cudaGetErrorString(e));
    return LASSO_B200_ERR_CUDA;
  }
  if (dev < 0 || dev >= kMaxDevices) {
    set_error("device ordinal %d out of range", dev);
    return LASSO_B200_ERR_INVALID;
  }
  *out = &g_ws[dev];
  return LASSO_B200_OK;
}

// ctl[0] = number of iterations the reference loop would have executed: index of the first
// iteration (before the last) whose delta met the stop test, plus one; else maxiter.
__global__ void find_stop_kernel(const double* __restrict__ hist, int maxiter, double tol_abs,
                                 int* __restrict__ ctl) {
  __shared__ int best;
  if (threadIdx.x == 0) best = maxiter;
  __syncthreads();
  if (tol_abs >= 0.0) {
    int mine = maxiter;
    for (int i = threadIdx.x; i < maxiter - 1; i += blockDim.x)
      if (hist[i] <= tol_abs) {
        mine = i + 1;
        break;
      }
    if (mine < maxiter) atomicMin(&best, mine);
  }
  __syncthreads();
  if (threadIdx.x == 0) ctl[0] = best;
}

// copy the buffer that holds z_done into z_out when it is not already there
__global__ void select_result_kernel(const float* __restrict__ z_a, const float* __restrict__ z_b,
                                     float* __restrict__ z_out, int64_t count,
                                     const int* __restrict__ ctl) {
  const int done = ctl[0];
  const float* src = (done & 1) ? z_b : z_a;  // z_i lives in (i even ? z_a : z_b)
  if (src == z_out) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    z_out[i] = src[i];
}

// Scope of one entry point's use of the device workspace: locks the device's mutex, makes `st` wait
// for the previous user if that was another stream, and records the hand-over event on exit.
class Lease {
 public:
  Lease() = default;
  Lease(const Lease&) = delete;
  Lease& operator=(const Lease&) = delete;
  int acquire(cudaStream_t st) {
    int rc = current_workspace(&ws_);
    if (rc) return rc;
    ws_->mutex.lock();
    locked_ = true;
    st_ = st;
    if (!ws_->last_use) {
      cudaError_t e = cudaEventCreateWithFlags(&ws_->last_use, cudaEventDisableTiming);
      if (e != cudaSuccess) {
        set_error("cudaEventCreate failed: %s", cudaGetErrorString(e));
        return LASSO_B200_ERR_CUDA;
      }
    }
    if (ws_->used && ws_->last_stream != st) {
      cudaError_t e = cudaStreamWaitEvent(st, ws_->last_use, 0);
      if (e != cudaSuccess) {
        set_error("cudaStreamWaitEvent failed: %s", cudaGetErrorString(e));
        return LASSO_B200_ERR_CUDA;
      }
    }
    return LASSO_B200_OK;
  }
  Workspace* operator->() { return ws_; }
  Workspace* get() { return ws_; }
  ~Lease() {
    if (!locked_) return;
    if (ws_->last_use && cudaEventRecord(ws_->last_use, st_) == cudaSuccess) {
      ws_->last_stream = st_;
      ws_->used = true;
    }
    ws_->mutex.unlock();
  }

 private:
  Workspace* ws_ = nullptr;
  cudaStream_t st_ = nullptr;
  bool locked_ = false;
};

// Drives the resident kernel under the batch-global stop test (ista.py:64, 93-95).  The kernel cannot stop
// in place (the sum spans all tiles), so every run records its per-iteration sums and the decision is taken
// afterwards: `run(iters, &fell_back)` executes `iters` iterations from the start codes into z_out and leaves
// hist[0 .. iters) on the device.
//   * no early stop: one run of maxiter iterations, one read-back of the record;
//   * the test fired at iteration `done` < the run's length: ONE replay with exactly `done` iterations
//     (deterministic kernel => the codes of stopping in place);
//   * a real tolerance and a long loop (tol_abs > 0, maxiter >= 256): probe runs of maxiter / 8 (>= 64), x4, ...
//     iterations first, so that a solve that stops after 50 of 1000 iterations costs 64 + 50 iterations, not
//     1000 + 50.  The default tol with the default maxiter = 10, and tol = 0, never probe.
// Returns the number of iterations of the run whose codes are in z_out (*done_out) or sets *fell_back.
template <typename Run>
int resident_stop_driver(Run&& run, int maxiter, double tol_abs, const double* hist, cudaStream_t st, int* done_out,
                         int* fell_back) {
  int rc, run_iters = maxiter;
  if (tol_abs > 0.0 && maxiter >= 256) run_iters = std::max(64, maxiter / 8);
  std::vector<double> h;
  for (;;) {
    if ((rc = run(run_iters, fell_back))) return rc;
    *done_out = run_iters;
    if (*fell_back || tol_abs < 0.0) return LASSO_B200_OK;
    h.resize((size_t)run_iters);
    LASSO_CUDA_TRY(cudaMemcpyAsync(h.data(), hist, sizeof(double) * h.size(), cudaMemcpyDeviceToHost, st));
    LASSO_CUDA_TRY(cudaStreamSynchronize(st));
    int stop = 0;
    for (int i = 0; i < run_iters; ++i)
      if (h[(size_t)i] <= tol_abs) {
        stop = i + 1;
        break;
      }
    if (stop == run_iters) return LASSO_B200_OK;        // stopped on the run's last iteration: these are the codes
    if (stop) {                                          // stopped earlier: replay exactly that many iterations
      if ((rc = run(stop, fell_back))) return rc;
      *done_out = stop;
      return LASSO_B200_OK;
    }
    if (run_iters == maxiter) return LASSO_B200_OK;     // ran to the end
    run_iters = (int)std::min<int64_t>((int64_t)run_iters * 4, maxiter);
  }
}

int check_problem(const void* x, const void* w, const void* z_out, int64_t n, int d, int k) {
  if (n < 0 || d <= 0 || k <= 0) {
    set_error("invalid shape n=%lld d=%d k=%d", (long long)n, d, k);
    return LASSO_B200_ERR_INVALID;
  }
  if (n > 0 && (!x || !z_out)) {
    set_error("null x / z_out pointer");
    return LASSO_B200_ERR_INVALID;
  }
  if (!w) {
    set_error("null weight pointer");
    return LASSO_B200_ERR_INVALID;
  }
  return LASSO_B200_OK;
}

}  // namespace
}  // namespace lasso

using namespace lasso;

extern "C" {

int32_t lasso_b200_version(void) { return 1000; }

const char* lasso_b200_last_error(void) { return t_error; }

int64_t lasso_b200_launch_count(void) { return (int64_t)g_launches.load(); }

int64_t lasso_b200_resident_fallbacks(void) { return (int64_t)g_res_fallbacks.load(); }

int32_t lasso_b200_select_path(int64_t n, int32_t d, int32_t k) {
  if (fista_res_supported(n, d, k)) return LASSO_B200_PATH_RESIDENT;
  if (fista_tc_supported(n, d, k)) return LASSO_B200_PATH_TCGEN05;
  if (fista_blk_supported(n, d, k)) return LASSO_B200_PATH_BLOCKED;
  return fista_gram_supported(n, d, k) ? LASSO_B200_PATH_GRAM : LASSO_B200_PATH_FFMA;
}

}  // extern "C" (reopened below)

namespace lasso {
namespace {
// body of lasso_b200_fista_f32; the caller holds the device's workspace lease
int fista_device_impl(Workspace* ws, const float* x, const float* weight, const float* z0, float* z_out,
                      int64_t n, int32_t d, int32_t k, double alpha, double lr, int32_t maxiter,
                      int32_t fast, double tol_abs, int32_t* iters_done, double* delta_hist,
                      int32_t path, cudaStream_t st);
}  // namespace
}  // namespace lasso

extern "C" {

int32_t lasso_b200_fista_f32(const float* x, const float* weight, const float* z0, float* z_out,
                             int64_t n, int32_t d, int32_t k, double alpha, double lr,
                             int32_t maxiter, int32_t fast, double tol_abs, int32_t* iters_done,
                             double* delta_hist, int32_t path, void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, weight, z_out, n, d, k);
  if (rc) return rc;
  if (maxiter < 0 || !(lr > 0.0) || !std::isfinite(lr) || !std::isfinite(alpha)) {
    set_error("invalid maxiter=%d / lr=%g / alpha=%g", maxiter, lr, alpha);
    return LASSO_B200_ERR_INVALID;
  }
  if (n == 0) {
    if (iters_done) *iters_done = 0;
    return LASSO_B200_OK;
  }
  if (path == LASSO_B200_PATH_AUTO) path = lasso_b200_select_path(n, d, k);
  if (path != LASSO_B200_PATH_FFMA && path != LASSO_B200_PATH_TCGEN05 &&
      path != LASSO_B200_PATH_RESIDENT && path != LASSO_B200_PATH_BLOCKED && path != LASSO_B200_PATH_GRAM) {
    set_error("unknown path %d", path);
    return LASSO_B200_ERR_INVALID;
  }
  if ((path == LASSO_B200_PATH_TCGEN05 && !fista_tc_supported(n, d, k)) ||
      (path == LASSO_B200_PATH_RESIDENT && !fista_res_supported(n, d, k)) ||
      (path == LASSO_B200_PATH_BLOCKED && !fista_blk_supported(n, d, k)) ||
      (path == LASSO_B200_PATH_GRAM && !fista_gram_supported(n, d, k))) {
    set_error("tcgen05 paths do not take n=%lld d=%d k=%d", (long long)n, d, k);
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t code_bytes = sizeof(float) * (size_t)n * (size_t)k;

  if (maxiter == 0) {
    if (z0 == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(z_out, 0, code_bytes, st));
    else if (z0 != z_out)
      LASSO_CUDA_TRY(cudaMemcpyAsync(z_out, z0, code_bytes, cudaMemcpyDeviceToDevice, st));
    if (iters_done) *iters_done = 0;
    return LASSO_B200_OK;
  }

  Lease ws;
  if ((rc = ws.acquire(st))) return rc;
  return fista_device_impl(ws.get(), x, weight, z0, z_out, n, d, k, alpha, lr, maxiter, fast, tol_abs,
                           iters_done, delta_hist, path, st);
}

}  // extern "C" (reopened below)

namespace lasso {
namespace {
int fista_device_impl(Workspace* ws, const float* x, const float* weight, const float* z0, float* z_out,
                      int64_t n, int32_t d, int32_t k, double alpha, double lr, int32_t maxiter,
                      int32_t fast, double tol_abs, int32_t* iters_done, double* delta_hist,
                      int32_t path, cudaStream_t st) {
  int rc;
  const size_t code_bytes = sizeof(float) * (size_t)n * (size_t)k;
  if ((rc = ensure(ws->code, code_bytes))) return rc;
  if ((rc = ensure(ws->hist, sizeof(double) * (size_t)maxiter))) return rc;
  if ((rc = ensure(ws->ctl, 256))) return rc;

  double* hist = (double*)ws->hist.ptr;
  const float lr_f = (float)lr;               // torch rounds the python scalar to the tensor dtype
  const float lam_f = (float)(alpha * lr);    // softshrink lambda = alpha*lr, product in double (ista.py:90)

  if (path == LASSO_B200_PATH_RESIDENT) {
    // All iterations on chip (fista_res.cu).  The stop test is batch-global, so it is taken
    // afterwards from the recorded sums: when it fired before maxiter the run is replayed with
    // exactly that many iterations (deterministic kernel => same codes as stopping in place).
    // If an iterate leaves the fp16 operand range the batch is solved again by the streaming
    // bf16x3 kernel below.  Both need z0 intact, so an aliased z0 is stashed first.
    const float* z_start = z0;
    if (z0 != nullptr && z0 == z_out) {
      LASSO_CUDA_TRY(cudaMemcpyAsync(ws->code.ptr, z0, code_bytes, cudaMemcpyDeviceToDevice, st));
      z_start = (const float*)ws->code.ptr;
    }
    const bool need_hist = tol_abs >= 0.0 || delta_hist != nullptr;
    // a threshold of exactly 0 only asks whether anything moved; the sums are not needed then
    const int hist_mode = !need_hist ? 0 : ((delta_hist == nullptr && tol_abs == 0.0) ? 2 : 1);
    int run_iters = maxiter, fell_back = 0;
    auto run = [&](int iters, int* fb) -> int {
      if (need_hist) LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));
      return fista_res_run(x, weight, z_start, z_out, n, d, k, lr_f, lam_f, iters, fast ? 1 : 0,
                           need_hist ? hist : nullptr, hist_mode, fb, st);
    };
    if ((rc = resident_stop_driver(run, maxiter, tol_abs, hist, st, &run_iters, &fell_back))) return rc;
    if (!fell_back) {
      if (delta_hist)
        LASSO_CUDA_TRY(cudaMemcpyAsync(delta_hist, hist, sizeof(double) * (size_t)maxiter,
                                       cudaMemcpyDeviceToDevice, st));
      if (iters_done) *iters_done = run_iters;
      return LASSO_B200_OK;
    }
    g_res_fallbacks.fetch_add(1, std::memory_order_relaxed);
    z0 = z_start;
    path = fista_tc_supported(n, d, k) ? LASSO_B200_PATH_TCGEN05 : LASSO_B200_PATH_FFMA;
  }

  if (path == LASSO_B200_PATH_GRAM) {
    // Gram-form kernel (fista_gram.cu): all iterations of a tile in one launch, so -- like the resident path -- the
    // batch-global stop test is taken afterwards from the recorded sums and a run that should have stopped early
    // is replayed with exactly that many iterations.  z_i lives in (i even ? A : B); A, B are chosen per run so
    // that the last iterate lands in z_out.  Start codes that alias z_out are stashed first.
    float* wsbuf = (float*)ws->code.ptr;
    float* stash = nullptr;
    const float* z_start = z0;
    if (z0 != nullptr && z0 == z_out) {
      LASSO_CUDA_TRY(cudaMallocAsync((void**)&stash, code_bytes, st));
      LASSO_CUDA_TRY(cudaMemcpyAsync(stash, z0, code_bytes, cudaMemcpyDeviceToDevice, st));
      z_start = stash;
    }
    const bool need_hist = tol_abs >= 0.0 || delta_hist != nullptr;
    int run_iters = maxiter, fell_back = 0;
    auto run = [&](int iters, int* fb) -> int {
      FistaArgs a{};
      a.x = x;
      a.w = weight;
      a.z_a = (iters & 1) ? wsbuf : z_out;
      a.z_b = (iters & 1) ? z_out : wsbuf;
      a.n = n;
      a.d = d;
      a.k = k;
      a.lr = lr_f;
      a.lam = lam_f;
      a.maxiter = iters;
      a.fast = fast ? 1 : 0;
      a.tol_abs = tol_abs;
      a.hist = hist;
      a.record = need_hist ? 1 : 0;
      if (z_start == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(a.z_a, 0, code_bytes, st));
      else LASSO_CUDA_TRY(cudaMemcpyAsync(a.z_a, z_start, code_bytes, cudaMemcpyDeviceToDevice, st));
      if (need_hist) LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));
      return fista_gram_run(a, fb, st);
    };
    rc = resident_stop_driver(run, maxiter, tol_abs, hist, st, &run_iters, &fell_back);
    if (!rc && !fell_back) {
      if (stash) cudaFreeAsync(stash, st);
      if (delta_hist)
        LASSO_CUDA_TRY(cudaMemcpyAsync(delta_hist, hist, sizeof(double) * (size_t)maxiter,
                                       cudaMemcpyDeviceToDevice, st));
      if (iters_done) *iters_done = run_iters;
      return LASSO_B200_OK;
    }
    if (rc) {
      if (stash) cudaFreeAsync(stash, st);
      return rc;
    }
    // an iterate left the fp16 operand range: the exact-fp32 FFMA kernel solves the batch from the start codes
    g_res_fallbacks.fetch_add(1, std::memory_order_relaxed);
    if (stash) {
      LASSO_CUDA_TRY(cudaMemcpyAsync(z_out, stash, code_bytes, cudaMemcpyDeviceToDevice, st));
      cudaFreeAsync(stash, st);     // (stream-ordered: the copy above is enqueued first)
    }
    path = LASSO_B200_PATH_FFMA;
  }

  // z_i lives in (i even ? z_a : z_b); put z_maxiter into z_out without a copy
  float* wsbuf = (float*)ws->code.ptr;
  float* z_a = (maxiter & 1) ? wsbuf : z_out;
  float* z_b = (maxiter & 1) ? z_out : wsbuf;
  if (z0 == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(z_a, 0, code_bytes, st));
  else if (z0 != z_a)
    LASSO_CUDA_TRY(cudaMemcpyAsync(z_a, z0, code_bytes, cudaMemcpyDeviceToDevice, st));
  LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));

  FistaArgs a{};
  a.x = x;
  a.w = weight;
  a.z_a = z_a;
  a.z_b = z_b;
  a.n = n;
  a.d = d;
  a.k = k;
  a.lr = lr_f;
  a.lam = lam_f;
  a.maxiter = maxiter;
  a.fast = fast ? 1 : 0;
  a.tol_abs = tol_abs;
  a.hist = hist;
  a.zero_start = z0 == nullptr ? 1 : 0;
  a.record = (tol_abs >= 0.0 || delta_hist != nullptr) ? 1 : 0;
  if (path == LASSO_B200_PATH_BLOCKED) {
    // like the resident path: an iterate beyond the fp16 operand range hands the batch to the FFMA
    // kernel, which needs the start codes again (they may alias z_out)
    float* stash = nullptr;
    if (z0 != nullptr && z0 == z_out) {
      LASSO_CUDA_TRY(cudaMallocAsync((void**)&stash, code_bytes, st));
      LASSO_CUDA_TRY(cudaMemcpyAsync(stash, z_a, code_bytes, cudaMemcpyDeviceToDevice, st));
    }
    int fell_back = 0;
    rc = fista_blk_run(a, &fell_back, st);
    if (!rc && fell_back) {
      g_res_fallbacks.fetch_add(1, std::memory_order_relaxed);
      const float* src = stash ? stash : z0;
      if (src == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(z_a, 0, code_bytes, st));
      else if (src != z_a) LASSO_CUDA_TRY(cudaMemcpyAsync(z_a, src, code_bytes, cudaMemcpyDeviceToDevice, st));
      LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));
      rc = fista_ffma_run(a, z_out, st);
    }
    if (stash) cudaFreeAsync(stash, st);
  } else {
    rc = (path == LASSO_B200_PATH_TCGEN05) ? fista_tc_run(a, z_out, st) : fista_ffma_run(a, z_out, st);
  }
  if (rc) return rc;

  int* ctl = (int*)ws->ctl.ptr;
  if (tol_abs >= 0.0 || iters_done) {
    find_stop_kernel<<<1, 256, 0, st>>>(hist, maxiter, tol_abs, ctl);
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  if (tol_abs >= 0.0) {
    int blocks = (int)std::min<int64_t>(((int64_t)n * k + 1023) / 1024, 148 * 8);
    select_result_kernel<<<blocks, 256, 0, st>>>(z_a, z_b, z_out, (int64_t)n * k, ctl);
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  if (delta_hist)
    LASSO_CUDA_TRY(cudaMemcpyAsync(delta_hist, hist, sizeof(double) * (size_t)maxiter,
                                   cudaMemcpyDeviceToDevice, st));
  if (iters_done) {
    int done = 0;
    LASSO_CUDA_TRY(cudaMemcpyAsync(&done, ctl, sizeof(int), cudaMemcpyDeviceToHost, st));
    LASSO_CUDA_TRY(cudaStreamSynchronize(st));
    *iters_done = done;
  }
  return LASSO_B200_OK;
}

}  // namespace
}  // namespace lasso

extern "C" {

int32_t lasso_b200_conv2d_fista_f32(const float* x, const float* weight_lin, const float* z0, float* z_out,
                                    int64_t n_img, int32_t cin, int32_t h, int32_t w, int32_t kh, int32_t kw,
                                    int32_t stride, int32_t padding, int32_t k, double alpha, double lr,
                                    int32_t maxiter, int32_t fast, double tol_abs, int32_t* iters_done,
                                    double* delta_hist, void* stream) {
  t_error[0] = 0;
  if (n_img < 0 || cin <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || padding < 0 || h + 2 * padding < kh ||
      w + 2 * padding < kw || k <= 0 || !weight_lin || (n_img > 0 && (!x || !z_out))) {
    set_error("invalid argument to conv2d_fista");
    return LASSO_B200_ERR_INVALID;
  }
  if (maxiter < 0 || !(lr > 0.0) || !std::isfinite(lr) || !std::isfinite(alpha)) {
    set_error("invalid maxiter=%d / lr=%g / alpha=%g", maxiter, lr, alpha);
    return LASSO_B200_ERR_INVALID;
  }
  if (n_img == 0) {
    if (iters_done) *iters_done = 0;
    return LASSO_B200_OK;
  }
  ConvShape shape{n_img, cin, h, w, kh, kw, stride, padding};
  if ((h + 2 * padding - kh) % stride != 0 || (w + 2 * padding - kw) % stride != 0) {
    // conv_transpose2d of the codes would be smaller than x (the reference fails on `x_hat - x`)
    set_error("conv2d_fista: (h + 2*padding - kh) and (w + 2*padding - kw) must be multiples of stride=%d", stride);
    return LASSO_B200_ERR_INVALID;
  }
  const int64_t P = (int64_t)shape.oh() * shape.ow(), n = n_img * P;
  const int d = cin * kh * kw;
  if (!conv2d_blk_supported(shape, k)) {
    set_error("conv2d path needs cin*kh*kw <= 128 (multiple of 4), filters <= 1024 (multiple of 4) and one image's "
              "patch matrix within 200 KB; got cin=%d %dx%d kernel, %d filters, %dx%d images, stride %d, padding %d",
              cin, kh, kw, k, h, w, stride, padding);
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t code_bytes = sizeof(float) * (size_t)n * (size_t)k;
  if (maxiter == 0) {
    if (z0 == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(z_out, 0, code_bytes, st));
    else if (z0 != z_out) LASSO_CUDA_TRY(cudaMemcpyAsync(z_out, z0, code_bytes, cudaMemcpyDeviceToDevice, st));
    if (iters_done) *iters_done = 0;
    return LASSO_B200_OK;
  }
  Lease ws;
  int rc = ws.acquire(st);
  if (rc) return rc;
  if ((rc = ensure(ws->code, code_bytes))) return rc;
  if ((rc = ensure(ws->hist, sizeof(double) * (size_t)maxiter))) return rc;
  if ((rc = ensure(ws->ctl, 256))) return rc;
  double* hist = (double*)ws->hist.ptr;
  float* wsbuf = (float*)ws->code.ptr;
  float* z_a = (maxiter & 1) ? wsbuf : z_out;   // z_i lives in (i even ? z_a : z_b)
  float* z_b = (maxiter & 1) ? z_out : wsbuf;
  if (z0 == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(z_a, 0, code_bytes, st));
  else if (z0 != z_a) LASSO_CUDA_TRY(cudaMemcpyAsync(z_a, z0, code_bytes, cudaMemcpyDeviceToDevice, st));
  LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));
  FistaArgs a{};
  a.x = x;
  a.w = weight_lin;
  a.z_a = z_a;
  a.z_b = z_b;
  a.n = n;
  a.d = d;
  a.k = k;
  a.lr = (float)lr;
  a.lam = (float)(alpha * lr);
  a.maxiter = maxiter;
  a.fast = fast ? 1 : 0;
  a.tol_abs = tol_abs;
  a.hist = hist;
  a.zero_start = z0 == nullptr ? 1 : 0;
  int fell_back = 0;
  if ((rc = fista_blk_run(a, &fell_back, st, &shape))) return rc;
  if (fell_back) {
    set_error("conv2d_fista: an iterate left the fp16 operand range of the tensor-core kernel (non-finite input, or "
              "codes more than 2^9 times larger than the per-image scaling allows); there is no CUDA-core conv path");
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  int* ctl = (int*)ws->ctl.ptr;
  find_stop_kernel<<<1, 256, 0, st>>>(hist, maxiter, tol_abs, ctl);
  LASSO_CHECK_LAUNCH();
  count_launch();
  if (tol_abs >= 0.0) {
    int blocks = (int)std::min<int64_t>(((int64_t)n * k + 1023) / 1024, 148 * 8);
    select_result_kernel<<<blocks, 256, 0, st>>>(z_a, z_b, z_out, (int64_t)n * k, ctl);
    LASSO_CHECK_LAUNCH();
    count_launch();
  }
  if (delta_hist)
    LASSO_CUDA_TRY(cudaMemcpyAsync(delta_hist, hist, sizeof(double) * (size_t)maxiter, cudaMemcpyDeviceToDevice, st));
  if (iters_done) {
    int done = 0;
    LASSO_CUDA_TRY(cudaMemcpyAsync(&done, ctl, sizeof(int), cudaMemcpyDeviceToHost, st));
    LASSO_CUDA_TRY(cudaStreamSynchronize(st));
    *iters_done = done;
  }
  return LASSO_B200_OK;
}

}  // extern "C" (reopened below)

namespace lasso {
namespace {

// Resident solve of a host-resident batch: the batch is cut into waves (one tile per SM) that
// flow through a three-stage pipeline -- H2D of x (and z0) on one copy stream, the solve on the
// compute stream, D2H of the codes on a second copy stream -- through a ring of kPipeSlots wave
// buffers.  All but the first upload and the last download hide behind the kernels, and the
// device footprint is kPipeSlots waves whatever n is (host batches larger than HBM stream through).
// hist (device, [iters]) accumulates the stop-test record of all waves.  *fell_back = 1: an iterate
// left the fp16 range somewhere; the caller must redo the batch on another path.
// true when the pointer is page-locked host memory the device can address (cudaHostAlloc / cudaHostRegister)
bool host_pinned(const void* ptr) {
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess) {
    (void)cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost && attr.devicePointer == ptr;
}

constexpr int kPipeSlots = 3;
// `st` is the caller's stream (the solves run on it); the caller holds the workspace lease.
int host_pipeline(Workspace* ws, const float* x, const float* z0, float* z_out, const float* dw, int64_t n,
                  int d, int k, float lr_f, float lam_f, int iters, int fast, double* hist, int hist_mode,
                  int* fell_back, cudaStream_t st) {
  int dev = 0, rc;
  LASSO_CUDA_TRY(cudaGetDevice(&dev));
  HostPipe& hp = g_pipe[dev];
  if (!hp.s_in) {
    LASSO_CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_in, cudaStreamNonBlocking));
    LASSO_CUDA_TRY(cudaStreamCreateWithFlags(&hp.s_out, cudaStreamNonBlocking));
  }
  // Waves of full-height tiles, the remainder in the LAST wave: a wave's duration does not depend on its tile height
  // (a tile-iteration is a serial chain, not tensor-pipe bound), so the balanced split only made the last download --
  // the part of the pipeline nothing overlaps -- as large as all the others.  LASSO_B200_PIPE=even: the balanced split.
  int64_t wave = fista_res_wave_rows(n), trows = fista_res_tile_rows(n);
  {
    const char* pm = getenv("LASSO_B200_PIPE");
    if (!(pm && pm[0] == 'e')) {
      wave = (wave / trows) * 128;       // one 128-row tile per SM
      trows = 0;                         // every launch picks its tile height from its own rows
    }
  }
  const int64_t nchunks = (n + wave - 1) / wave;
  const int slots = (int)std::min<int64_t>(kPipeSlots, nchunks);
  if ((rc = ensure(ws->hx, sizeof(float) * (size_t)slots * wave * d))) return rc;
  if ((rc = ensure(ws->hz, sizeof(float) * (size_t)slots * wave * k))) return rc;
  float* dx = (float*)ws->hx.ptr;
  float* dz = (float*)ws->hz.ptr;
  while ((int)hp.events.size() < 3 * kPipeSlots + 1) {
    cudaEvent_t ev;
    LASSO_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    hp.events.push_back(ev);
  }
  // events: [3 s] chunk in slot s uploaded, [3 s + 1] solved, [3 s + 2] downloaded; [last] start
  cudaEvent_t ev_start = hp.events[3 * kPipeSlots];
  if ((rc = fista_res_prepare(dw, d, k, lr_f, lam_f, iters, fast, st))) return rc;
  if (hist) LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)iters, st));
  // the copy streams must not run ahead of work of an earlier call still queued on `st`
  LASSO_CUDA_TRY(cudaEventRecord(ev_start, st));
  LASSO_CUDA_TRY(cudaStreamWaitEvent(hp.s_in, ev_start, 0));
  LASSO_CUDA_TRY(cudaStreamWaitEvent(hp.s_out, ev_start, 0));
  auto upload = [&](int64_t c) -> int {
    const int s = (int)(c % slots);
    const int64_t r0 = c * wave, rows = std::min<int64_t>(wave, n - r0);
    // the slot's previous tenant (chunk c - slots) must have been downloaded
    if (c >= slots) LASSO_CUDA_TRY(cudaStreamWaitEvent(hp.s_in, hp.events[3 * s + 2], 0));
    LASSO_CUDA_TRY(cudaMemcpyAsync(dx + (int64_t)s * wave * d, x + r0 * d, sizeof(float) * (size_t)rows * d,
                                   cudaMemcpyHostToDevice, hp.s_in));
    if (z0)
      LASSO_CUDA_TRY(cudaMemcpyAsync(dz + (int64_t)s * wave * k, z0 + r0 * k, sizeof(float) * (size_t)rows * k,
                                     cudaMemcpyHostToDevice, hp.s_in));
    LASSO_CUDA_TRY(cudaEventRecord(hp.events[3 * s], hp.s_in));
    return LASSO_B200_OK;
  };
  for (int64_t c = 0; c < slots; ++c)
    if ((rc = upload(c))) return rc;
  for (int64_t c = 0; c < nchunks; ++c) {
    const int s = (int)(c % slots);
    const int64_t r0 = c * wave, rows = std::min<int64_t>(wave, n - r0);
    float* zc = dz + (int64_t)s * wave * k;
    LASSO_CUDA_TRY(cudaStreamWaitEvent(st, hp.events[3 * s], 0));
    if ((rc = fista_res_launch(dx + (int64_t)s * wave * d, z0 ? zc : nullptr, zc, rows, d, k, iters, hist,
                               hist_mode, trows, st)))
      return rc;
    LASSO_CUDA_TRY(cudaEventRecord(hp.events[3 * s + 1], st));
    LASSO_CUDA_TRY(cudaStreamWaitEvent(hp.s_out, hp.events[3 * s + 1], 0));
    LASSO_CUDA_TRY(cudaMemcpyAsync(z_out + r0 * k, zc, sizeof(float) * (size_t)rows * k,
                                   cudaMemcpyDeviceToHost, hp.s_out));
    LASSO_CUDA_TRY(cudaEventRecord(hp.events[3 * s + 2], hp.s_out));
    if (c + slots < nchunks && (rc = upload(c + slots))) return rc;
  }
  if ((rc = fista_res_finish(fell_back, st))) return rc;
  LASSO_CUDA_TRY(cudaStreamSynchronize(hp.s_out));
  return LASSO_B200_OK;
}

}  // namespace
}  // namespace lasso

extern "C" {

int32_t lasso_b200_fista_f32_host(const float* x, const float* weight, const float* z0,
                                  float* z_out, int64_t n, int32_t d, int32_t k, double alpha,
                                  double lr, int32_t maxiter, int32_t fast, double tol_abs,
                                  int32_t* iters_done, double* delta_hist, int32_t path,
                                  void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, weight, z_out, n, d, k);
  if (rc) return rc;
  if (maxiter < 0) {
    set_error("invalid maxiter=%d", maxiter);
    return LASSO_B200_ERR_INVALID;
  }
  if (n == 0) {
    if (iters_done) *iters_done = 0;
    return LASSO_B200_OK;
  }
  const size_t xb = sizeof(float) * (size_t)n * d, wb = sizeof(float) * (size_t)d * k;
  const size_t zb = sizeof(float) * (size_t)n * k;
  cudaStream_t st = (cudaStream_t)stream;
  int done = maxiter;
  // the lease is held for the whole call (the entry point returns only after the codes are in the
  // host buffer): two host threads on one device take turns
  Lease ws;
  if ((rc = ws.acquire(st))) return rc;
  if ((rc = ensure(ws->hw, wb))) return rc;
  if ((rc = ensure(ws->hist, sizeof(double) * (size_t)(maxiter + 1)))) return rc;
  float* dw = (float*)ws->hw.ptr;
  double* hist = (double*)ws->hist.ptr;

  const int resolved = path == LASSO_B200_PATH_AUTO ? lasso_b200_select_path(n, d, k) : path;
  if (resolved == LASSO_B200_PATH_RESIDENT && fista_res_supported(n, d, k) && maxiter > 0 &&
      std::isfinite(lr) && lr > 0.0 && std::isfinite(alpha)) {
    const bool need_hist = tol_abs >= 0.0 || delta_hist != nullptr;
    // a threshold of exactly 0 only asks whether anything moved; the sums are not needed then
    const int hist_mode = !need_hist ? 0 : ((delta_hist == nullptr && tol_abs == 0.0) ? 2 : 1);
    const float lr_f = (float)lr, lam_f = (float)(alpha * lr);
    LASSO_CUDA_TRY(cudaMemcpyAsync(dw, weight, wb, cudaMemcpyHostToDevice, st));
    // z0 may alias z_out: a second pass (early stop) or the streaming fallback needs the start codes
    // again after pass 1 has overwritten them
    std::vector<float> z0_keep;
    const float* z_start = z0;
    if (z0 != nullptr && z0 == z_out) {
      z0_keep.assign(z0, z0 + (size_t)n * k);
      z_start = z0_keep.data();
    }
    // Pinned (page-locked) host buffers are addressable from the device (unified addressing), so the
    // resident kernel could read its x tiles and write its code tiles straight through PCIe with no
    // staging at all.  Measured at config 2 it is SLOWER than the staged three-stream pipeline (3.80 vs
    // 3.32 ms per 200-iteration solve): all SMs reach their tile stores together and then stall on the
    // bus, while the copy engines of the pipeline run beside the kernels.  Kept as an opt-in
    // (LASSO_B200_ZERO_COPY=1) for hosts where staging memory is the constraint.
    const bool zero_copy = getenv("LASSO_B200_ZERO_COPY") != nullptr && host_pinned(x) && host_pinned(z_out) &&
                           (z_start == nullptr || host_pinned(z_start));
    int fell_back = 0, run_iters = maxiter;
    auto run = [&](int iters, int* fb) -> int {
      int rc2;
      if (zero_copy) {
        if ((rc2 = fista_res_prepare(dw, d, k, lr_f, lam_f, iters, fast ? 1 : 0, st))) return rc2;
        if (need_hist) LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));
        if ((rc2 = fista_res_launch(x, z_start, z_out, n, d, k, iters, need_hist ? hist : nullptr, hist_mode, 0, st)))
          return rc2;
        return fista_res_finish(fb, st);
      }
      if (need_hist) LASSO_CUDA_TRY(cudaMemsetAsync(hist, 0, sizeof(double) * (size_t)maxiter, st));
      return host_pipeline(ws.get(), x, z_start, z_out, dw, n, d, k, lr_f, lam_f, iters, fast ? 1 : 0,
                           need_hist ? hist : nullptr, hist_mode, fb, st);
    };
    if ((rc = resident_stop_driver(run, maxiter, tol_abs, hist, st, &run_iters, &fell_back))) return rc;
    if (!fell_back) {
      if (delta_hist) {
        std::vector<double> h((size_t)maxiter, 0.0);
        LASSO_CUDA_TRY(cudaMemcpyAsync(h.data(), hist, sizeof(double) * (size_t)run_iters, cudaMemcpyDeviceToHost, st));
        LASSO_CUDA_TRY(cudaStreamSynchronize(st));
        memcpy(delta_hist, h.data(), sizeof(double) * (size_t)maxiter);
      }
      if (iters_done) *iters_done = run_iters;
      return LASSO_B200_OK;
    }
    g_res_fallbacks.fetch_add(1, std::memory_order_relaxed);
    path = fista_tc_supported(n, d, k) ? LASSO_B200_PATH_TCGEN05 : LASSO_B200_PATH_FFMA;
    if (z_start != z0) {
      // restore the start codes the first pass overwrote
      memcpy(z_out, z0_keep.data(), zb);
    }
  }

  // plain copy - solve - copy (streaming kernels; whole batch on the device)
  if ((rc = ensure(ws->hx, xb))) return rc;
  if ((rc = ensure(ws->hz, zb + sizeof(double) * (size_t)(maxiter + 1)))) return rc;
  float* dx = (float*)ws->hx.ptr;
  float* dz = (float*)ws->hz.ptr;
  double* dh = delta_hist ? (double*)((char*)ws->hz.ptr + ((zb + 7) & ~(size_t)7)) : nullptr;
  LASSO_CUDA_TRY(cudaMemcpyAsync(dx, x, xb, cudaMemcpyHostToDevice, st));
  LASSO_CUDA_TRY(cudaMemcpyAsync(dw, weight, wb, cudaMemcpyHostToDevice, st));
  if (z0) LASSO_CUDA_TRY(cudaMemcpyAsync(dz, z0, zb, cudaMemcpyHostToDevice, st));
  if (path == LASSO_B200_PATH_AUTO) path = lasso_b200_select_path(n, d, k);
  if (maxiter == 0) {
    if (z0 == nullptr) LASSO_CUDA_TRY(cudaMemsetAsync(dz, 0, zb, st));
    done = 0;
  } else {
    if (path != LASSO_B200_PATH_FFMA && path != LASSO_B200_PATH_TCGEN05 && path != LASSO_B200_PATH_RESIDENT &&
        path != LASSO_B200_PATH_BLOCKED && path != LASSO_B200_PATH_GRAM) {
      set_error("unknown path %d", path);
      return LASSO_B200_ERR_INVALID;
    }
    if (!(lr > 0.0) || !std::isfinite(lr) || !std::isfinite(alpha)) {
      set_error("invalid lr=%g / alpha=%g", lr, alpha);
      return LASSO_B200_ERR_INVALID;
    }
    if ((path == LASSO_B200_PATH_TCGEN05 && !fista_tc_supported(n, d, k)) ||
        (path == LASSO_B200_PATH_RESIDENT && !fista_res_supported(n, d, k)) ||
        (path == LASSO_B200_PATH_BLOCKED && !fista_blk_supported(n, d, k)) ||
        (path == LASSO_B200_PATH_GRAM && !fista_gram_supported(n, d, k))) {
      set_error("tcgen05 paths do not take n=%lld d=%d k=%d", (long long)n, d, k);
      return LASSO_B200_ERR_UNSUPPORTED;
    }
    rc = fista_device_impl(ws.get(), dx, dw, z0 ? dz : nullptr, dz, n, d, k, alpha, lr, maxiter, fast, tol_abs,
                           &done, dh, path, st);
    if (rc) return rc;
  }
  LASSO_CUDA_TRY(cudaMemcpyAsync(z_out, dz, zb, cudaMemcpyDeviceToHost, st));
  if (delta_hist && maxiter > 0)
    LASSO_CUDA_TRY(cudaMemcpyAsync(delta_hist, dh, sizeof(double) * (size_t)maxiter,
                                   cudaMemcpyDeviceToHost, st));
  LASSO_CUDA_TRY(cudaStreamSynchronize(st));
  if (iters_done) *iters_done = done;
  return LASSO_B200_OK;
}

int32_t lasso_b200_lipschitz_f32(const float* weight, int32_t d, int32_t k, int32_t iters,
                                 double* l_out, void* stream) {
  t_error[0] = 0;
  if (!weight || !l_out || d <= 0 || k <= 0 || iters <= 0) {
    set_error("invalid argument to lipschitz");
    return LASSO_B200_ERR_INVALID;
  }
  const int m = d <= k ? d : k;
  if (m > 4096) {
    set_error("lipschitz: min(d,k)=%d exceeds 4096", m);
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Lease ws;
  int rc = ws.acquire(st);
  if (rc) return rc;
  if ((rc = ensure(ws->scratch, sizeof(double) * lambda_max_scratch_doubles(m)))) return rc;
  double* scratch = (double*)ws->scratch.ptr;
  double* l_dev = scratch + (size_t)m * m + 2 * (size_t)m;
  if ((rc = lipschitz_run(weight, d, k, iters, l_dev, scratch, st))) return rc;
  LASSO_CUDA_TRY(cudaMemcpyAsync(l_out, l_dev, sizeof(double), cudaMemcpyDeviceToHost, st));
  LASSO_CUDA_TRY(cudaStreamSynchronize(st));
  return LASSO_B200_OK;
}

int32_t lasso_b200_conv2d_lipschitz_f32(const float* weight, int32_t filters, int32_t cin, int32_t kh,
                                        int32_t kw, int32_t h, int32_t w, int32_t stride, int32_t padding,
                                        int32_t iters, double* l_out, void* stream) {
  t_error[0] = 0;
  if (!weight || !l_out || filters <= 0 || cin <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || padding < 0 ||
      h + 2 * padding < kh || w + 2 * padding < kw || iters <= 0) {
    set_error("invalid argument to conv2d_lipschitz");
    return LASSO_B200_ERR_INVALID;
  }
  if ((int64_t)cin * h * w > 4096) {
    set_error("conv2d_lipschitz: cin*h*w = %lld exceeds 4096 (dense operator)", (long long)cin * h * w);
    return LASSO_B200_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Lease ws;
  int rc = ws.acquire(st);
  if (rc) return rc;
  if ((rc = ensure(ws->scratch, conv_lipschitz_scratch_bytes(cin, kh, kw, h, w)))) return rc;
  return conv_lipschitz_run(weight, filters, cin, kh, kw, h, w, stride, padding, iters, l_out,
                            ws->scratch.ptr, st);
}

int32_t lasso_b200_ridge_init_f32(const float* x, const float* weight, int64_t n, int32_t d, int32_t k,
                                  double alpha, float* z_out, int32_t* not_positive_definite, void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, weight, z_out, n, d, k);
  if (rc) return rc;
  if (!not_positive_definite || !std::isfinite(alpha)) {
    set_error("invalid argument to ridge_init");
    return LASSO_B200_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Lease ws;
  if ((rc = ws.acquire(st))) return rc;
  if ((rc = ensure(ws->scratch, ridge_scratch_bytes(d, k)))) return rc;
  int flag = 0;
  rc = ridge_init_run(x, weight, n, d, k, alpha, z_out, ws->scratch.ptr, &flag, st);
  *not_positive_definite = flag;
  return rc;
}

int32_t lasso_b200_matmul_f32(const float* x, const float* t, int64_t n, int32_t d, int32_t k, float* z_out,
                              void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, t, z_out, n, d, k);
  if (rc) return rc;
  return matmul_run(x, t, n, d, k, z_out, (cudaStream_t)stream);
}

int32_t lasso_b200_loss_terms_f32(const float* x, const float* z, const float* weight, int64_t n,
                                  int32_t d, int32_t k, double* out, void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, weight, z, n, d, k);
  if (rc) return rc;
  if (!out) {
    set_error("null out pointer");
    return LASSO_B200_ERR_INVALID;
  }
  if (n == 0) {
    LASSO_CUDA_TRY(cudaMemsetAsync(out, 0, 2 * sizeof(double), (cudaStream_t)stream));
    return LASSO_B200_OK;
  }
  return loss_terms_run(x, z, weight, n, d, k, out, (cudaStream_t)stream);
}

int32_t lasso_b200_gram_f32(const float* z, const float* x, int64_t n, int32_t d, int32_t k,
                            double* gram_zz, double* gram_zx, void* stream) {
  t_error[0] = 0;
  if (n < 0 || d <= 0 || k <= 0 || !gram_zz || !gram_zx || (n > 0 && (!z || !x))) {
    set_error("invalid argument to gram");
    return LASSO_B200_ERR_INVALID;
  }
  // (the tensor-core kernel keeps its scales in per-device scratch: serialised like every other user of the workspace)
  Lease ws;
  int rc = ws.acquire((cudaStream_t)stream);
  if (rc) return rc;
  return gram_run(z, x, n, d, k, gram_zz, gram_zx, (cudaStream_t)stream);
}

int32_t lasso_b200_dict_update_gram_f32(float* dict, double* gram_zz, double* gram_zx, int32_t d,
                                        int32_t k, double eps, const float* redraw,
                                        int32_t* zeroed, int32_t positive, void* stream) {
  t_error[0] = 0;
  if (!dict || !gram_zz || !gram_zx || !zeroed || d <= 0 || k <= 0) {
    set_error("invalid argument to dict_update_gram");
    return LASSO_B200_ERR_INVALID;
  }
  // (the blocked sweep keeps U0 and the dead-atom flags in per-device scratch)
  Lease ws;
  int rc = ws.acquire((cudaStream_t)stream);
  if (rc) return rc;
  return dict_update_run(dict, gram_zz, gram_zx, d, k, eps, redraw, zeroed, positive ? 1 : 0, (cudaStream_t)stream);
}

int32_t lasso_b200_zero_columns_f32(float* z, int64_t n, int32_t k, const int32_t* mask, void* stream) {
  t_error[0] = 0;
  if (n < 0 || k <= 0 || !mask || (n > 0 && !z)) {
    set_error("invalid argument to zero_columns");
    return LASSO_B200_ERR_INVALID;
  }
  return zero_columns_run(z, n, k, mask, (cudaStream_t)stream);
}

int32_t lasso_b200_gradient_f32(const float* x, const float* point, const float* weight, int64_t n,
                                int32_t d, int32_t k, float* grad, double* f_sum, void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, weight, grad, n, d, k);
  if (rc) return rc;
  if (!f_sum || (n > 0 && !point)) {
    set_error("null pointer passed to gradient");
    return LASSO_B200_ERR_INVALID;
  }
  if (n == 0) {
    LASSO_CUDA_TRY(cudaMemsetAsync(f_sum, 0, sizeof(double), (cudaStream_t)stream));
    return LASSO_B200_OK;
  }
  return gradient_run(x, point, weight, n, d, k, grad, f_sum, (cudaStream_t)stream);
}

int32_t lasso_b200_linesearch_trial_f32(const float* x, const float* point, const float* grad,
                                        const float* weight, int64_t n, int32_t d, int32_t k,
                                        double step, double alpha, float* cand, double* sums,
                                        void* stream) {
  t_error[0] = 0;
  int rc = check_problem(x, weight, cand, n, d, k);
  if (rc) return rc;
  if (!sums || (n > 0 && (!point || !grad)) || !(step > 0.0)) {
    set_error("invalid argument to linesearch_trial");
    return LASSO_B200_ERR_INVALID;
  }
  if (n == 0) {
    LASSO_CUDA_TRY(cudaMemsetAsync(sums, 0, 4 * sizeof(double), (cudaStream_t)stream));
    return LASSO_B200_OK;
  }
  return trial_run(x, point, grad, weight, n, d, k, (float)step, (float)(alpha * step), cand, sums,
                   (cudaStream_t)stream);
}

int32_t lasso_b200_momentum_f32(const float* z_next, const float* z, double beta, float* y,
                                int64_t count, double* delta, void* stream) {
  t_error[0] = 0;
  if (count < 0 || !delta || (count > 0 && (!z_next || !z))) {
    set_error("invalid argument to momentum");
    return LASSO_B200_ERR_INVALID;
  }
  return momentum_run(z_next, z, (float)beta, y, count, delta, (cudaStream_t)stream);
}

int32_t lasso_b200_release_workspace(void) {
  Workspace* ws = nullptr;
  int rc = current_workspace(&ws);
  if (rc) return rc;
  std::lock_guard<std::mutex> lock(ws->mutex);
  if (ws->used) cudaEventSynchronize(ws->last_use);   // nobody may still be running on the buffers
  Buffer* all[] = {&ws->code, &ws->hist, &ws->ctl, &ws->scratch, &ws->hx, &ws->hw, &ws->hz};
  for (Buffer* b : all) {
    if (b->ptr) cudaFree(b->ptr);
    b->ptr = nullptr;
    b->bytes = 0;
  }
  return LASSO_B200_OK;
}

}  // extern "C"
