// C ABI of the B200-native LoMPC hot path (see include/lompc_b200.h).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>

#include "lompc_common.cuh"
#include "lompc_solve.cuh"

namespace {

thread_local char g_cuda_err[256] = "";
std::atomic<int64_t> g_launches{0};

int cuda_fail(cudaError_t e, const char* what) {
  snprintf(g_cuda_err, sizeof(g_cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return LOMPC_ERR_CUDA;
}
#define CK(call)                                   \
  do {                                             \
    cudaError_t e__ = (call);                      \
    if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
  } while (0)

// settings.py:7-9
constexpr double kMinMaxBatSoc = 0.75;
constexpr double kMaxMaxBatSoc = 0.9;
constexpr double kMaxBatChargeRate = 0.25;

}  // namespace

struct lompc_handle {
  lompc::Consts cs;
  int device;
  int max_iter;
  double tol;
  double delta;
  // grow-only device workspace for the _host entry points
  void* ws;
  size_t ws_bytes;
};

namespace {

int ensure_ws(lompc_handle* h, size_t bytes) {
  if (h->ws_bytes >= bytes) return LOMPC_OK;
  if (h->ws) CK(cudaFree(h->ws));
  h->ws = nullptr;
  h->ws_bytes = 0;
  size_t want = bytes + bytes / 4;
  CK(cudaMalloc(&h->ws, want));
  h->ws_bytes = want;
  return LOMPC_OK;
}

template <int NSEG>
int launch_solve(const lompc_handle* h, const lompc::SolveArgs& a, cudaStream_t stream) {
  const int N = h->cs.N;
  // Threads per block: as many as fit the 227 KB of shared memory, capped at 128;
  // small batches use small blocks so that more SMs take part.
  int T = 128;
  while (T > 32 && lompc::SmemLayout<NSEG>::bytes(N, T) > 200 * 1024) T >>= 1;
  while (T > 32 && (a.B + T - 1) / T < 148) T >>= 1;
  const size_t smem = lompc::SmemLayout<NSEG>::bytes(N, T);
  if (smem > 227 * 1024) return LOMPC_ERR_ARG;
  static thread_local size_t configured[2] = {0, 0};
  const int slot = NSEG > 1 ? 1 : 0;
  if (smem > configured[slot]) {
    CK(cudaFuncSetAttribute(lompc::lompc_solve_kernel<NSEG>,
                            cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured[slot] = 227 * 1024;
  }
  const int64_t blocks = (a.B + T - 1) / T;
  lompc::lompc_solve_kernel<NSEG><<<(unsigned)blocks, T, smem, stream>>>(h->cs, a);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  CK(cudaGetLastError());
  return LOMPC_OK;
}

}  // namespace

extern "C" {

const char* lompc_version(void) { return "lompc_b200 0.1 (sm_100a)"; }

const char* lompc_strerror(int code) {
  switch (code) {
    case LOMPC_OK: return "ok";
    case LOMPC_ERR_CONSTS: return "invalid LoMPC constants (lompc.py:36-38)";
    case LOMPC_ERR_ARG: return "invalid argument";
    case LOMPC_ERR_CUDA: return "CUDA error";
    case LOMPC_ERR_GAMMA: return "gamma > y_max (lompc.py:87)";
    case LOMPC_ERR_NEGATIVE: return "negative value for a nonneg parameter (lompc.py:78-82)";
    case LOMPC_ERR_NOT_CONVERGED: return "solver did not converge";
    case LOMPC_ERR_NO_DEVICE: return "no CUDA device (this library has no CPU fallback)";
    default: return "unknown error";
  }
}

const char* lompc_last_cuda_error(void) { return g_cuda_err; }

int lompc_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int64_t lompc_launch_count(void) { return g_launches.load(); }

int lompc_create(int N, double delta, double theta, double y_max, double w_max, int ev_type,
                 int device, lompc_t** out) {
  if (!out || N < 1 || N > 4096) return LOMPC_ERR_ARG;
  *out = nullptr;
  // lompc.py:36-38
  if (!(y_max >= kMinMaxBatSoc && y_max <= kMaxMaxBatSoc)) return LOMPC_ERR_CONSTS;
  if (!(w_max >= 0.0 && w_max <= kMaxBatChargeRate)) return LOMPC_ERR_CONSTS;
  if (ev_type != LOMPC_EV_SMALL && ev_type != LOMPC_EV_LARGE) return LOMPC_ERR_CONSTS;
  if (!(w_max > 0.0) || !(theta > 0.0) || !(delta > 0.0)) return LOMPC_ERR_ARG;
  if (lompc_device_count() <= device || device < 0) return LOMPC_ERR_NO_DEVICE;

  lompc_handle* h = new (std::nothrow) lompc_handle();
  if (!h) return LOMPC_ERR_ARG;
  lompc::Consts& cs = h->cs;
  memset(&cs, 0, sizeof(cs));
  cs.N = N;
  cs.large = ev_type == LOMPC_EV_LARGE;
  cs.theta = theta;
  cs.w_max = w_max;
  cs.y_max = y_max;
  cs.c = 2.0 * delta * theta * theta;       // lompc.py:71
  cs.q_scale = 3.0 * theta / (4.0 * w_max);  // lompc.py:67
  cs.theta2 = theta * theta;
  if (!cs.large) {
    cs.d_base = 2.0 * theta * theta / 0.81;  // theta^2 * sum_squares(w / 0.9), lompc.py:107
    cs.nseg = 1;
    cs.brk[0] = 0.0;
    cs.brk[1] = w_max;
    cs.slope[0] = 0.0;
  } else {
    // (theta*w_max)^2 * sum max(0, x-.125, 1.5x-.375, 2x-.75), x = w/w_max  (lompc.py:109-116)
    cs.d_base = 0.0;
    cs.nseg = 4;
    const double brk_rel[5] = {0.0, 0.125, 0.5, 0.75, 1.0};
    const double slope_rel[4] = {0.0, 1.0, 1.5, 2.0};
    const double per_w = (theta * w_max) * (theta * w_max) / w_max;
    for (int i = 0; i <= 4; ++i) cs.brk[i] = brk_rel[i] * w_max;
    cs.brk[4] = w_max;
    for (int j = 0; j < 4; ++j) cs.slope[j] = per_w * slope_rel[j];
  }
  h->device = device;
  h->max_iter = 200;
  h->tol = 1e-11;
  h->delta = delta;
  h->ws = nullptr;
  h->ws_bytes = 0;
  *out = h;
  return LOMPC_OK;
}

int lompc_destroy(lompc_t* h) {
  if (!h) return LOMPC_OK;
  if (h->ws) {
    cudaSetDevice(h->device);
    cudaFree(h->ws);
  }
  delete h;
  return LOMPC_OK;
}

double lompc_sc_modulus(const lompc_t* h) { return h ? h->cs.c : NAN; }

int lompc_set_options(lompc_t* h, int max_iter, double tol) {
  if (!h || max_iter < 1 || !(tol > 0.0)) return LOMPC_ERR_ARG;
  h->max_iter = max_iter;
  h->tol = tol;
  return LOMPC_OK;
}

int lompc_solve_batch_dev(lompc_t* h, int64_t B, const double* lmbd, int64_t lmbd_stride,
                          const double* lmbd_r, int64_t lmbd_r_stride, const double* gamma,
                          double* w_out, double* cost_out, int32_t* status, int32_t* iters,
                          double* kkt_res, void* stream) {
  if (!h || B < 0 || !lmbd || !lmbd_r || !gamma || !w_out || !cost_out) return LOMPC_ERR_ARG;
  if (lmbd_stride != 0 && lmbd_stride < 3 * (int64_t)h->cs.N) return LOMPC_ERR_ARG;
  if (lmbd_r_stride != 0 && lmbd_r_stride != 1) return LOMPC_ERR_ARG;
  if (B == 0) return LOMPC_OK;
  CK(cudaSetDevice(h->device));
  lompc::SolveArgs a;
  a.B = B;
  a.lmbd = lmbd;
  a.lmbd_stride = lmbd_stride;
  a.lmbd_r = lmbd_r;
  a.lmbd_r_stride = lmbd_r_stride;
  a.gamma = gamma;
  a.w_out = w_out;
  a.cost_out = cost_out;
  a.status = status;
  a.iters = iters;
  a.kkt_res = kkt_res;
  a.max_iter = h->max_iter;
  a.tol = h->tol;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return h->cs.large ? launch_solve<4>(h, a, s) : launch_solve<1>(h, a, s);
}

int lompc_solve_batch_host(lompc_t* h, int64_t B, const double* lmbd, int64_t lmbd_stride,
                           const double* lmbd_r, int64_t lmbd_r_stride, const double* gamma,
                           double* w_out, double* cost_out, int32_t* status, int32_t* iters,
                           double* kkt_res) {
  if (!h || B < 0 || !lmbd || !lmbd_r || !gamma || !w_out || !cost_out) return LOMPC_ERR_ARG;
  if (B == 0) return LOMPC_OK;
  const int N = h->cs.N;
  if (lmbd_stride != 0 && lmbd_stride != 3 * (int64_t)N) return LOMPC_ERR_ARG;
  if (lmbd_r_stride != 0 && lmbd_r_stride != 1) return LOMPC_ERR_ARG;
  CK(cudaSetDevice(h->device));
  const size_t n_lm = (lmbd_stride ? (size_t)B : 1) * 3 * N;
  const size_t n_lr = lmbd_r_stride ? (size_t)B : 1;
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t o_lm = 0;
  const size_t o_lr = o_lm + al(n_lm * 8);
  const size_t o_ga = o_lr + al(n_lr * 8);
  const size_t o_w = o_ga + al((size_t)B * 8);
  const size_t o_c = o_w + al((size_t)B * N * 8);
  const size_t o_k = o_c + al((size_t)B * 8);
  const size_t o_st = o_k + al((size_t)B * 8);
  const size_t o_it = o_st + al((size_t)B * 4);
  const size_t total = o_it + al((size_t)B * 4);
  int rc = ensure_ws(h, total);
  if (rc) return rc;
  char* ws = static_cast<char*>(h->ws);
  cudaStream_t s = 0;
  CK(cudaMemcpyAsync(ws + o_lm, lmbd, n_lm * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ws + o_lr, lmbd_r, n_lr * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(ws + o_ga, gamma, (size_t)B * 8, cudaMemcpyHostToDevice, s));
  rc = lompc_solve_batch_dev(h, B, (const double*)(ws + o_lm), lmbd_stride,
                             (const double*)(ws + o_lr), lmbd_r_stride, (const double*)(ws + o_ga),
                             (double*)(ws + o_w), (double*)(ws + o_c), (int32_t*)(ws + o_st),
                             (int32_t*)(ws + o_it), (double*)(ws + o_k), s);
  if (rc) return rc;
  CK(cudaMemcpyAsync(w_out, ws + o_w, (size_t)B * N * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(cost_out, ws + o_c, (size_t)B * 8, cudaMemcpyDeviceToHost, s));
  // status is always fetched: it carries the reference's error conventions
  int32_t* st_host = status;
  int32_t* st_tmp = nullptr;
  if (!st_host) {
    st_tmp = new (std::nothrow) int32_t[B];
    if (!st_tmp) return LOMPC_ERR_ARG;
    st_host = st_tmp;
  }
  cudaError_t e = cudaMemcpyAsync(st_host, ws + o_st, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && iters)
    e = cudaMemcpyAsync(iters, ws + o_it, (size_t)B * 4, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess && kkt_res)
    e = cudaMemcpyAsync(kkt_res, ws + o_k, (size_t)B * 8, cudaMemcpyDeviceToHost, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    delete[] st_tmp;
    return cuda_fail(e, "copy-out");
  }
  rc = LOMPC_OK;
  for (int64_t i = 0; i < B; ++i) {
    if (st_host[i] == LOMPC_ST_BAD_GAMMA) { rc = LOMPC_ERR_GAMMA; break; }
    if (st_host[i] == LOMPC_ST_NEGATIVE) { rc = LOMPC_ERR_NEGATIVE; break; }
    if (st_host[i] == LOMPC_ST_MAXITER) rc = LOMPC_ERR_NOT_CONVERGED;
  }
  delete[] st_tmp;
  return rc;
}

}  // extern "C"
