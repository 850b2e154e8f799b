// Launcher of the warp-cooperative K1 (lompc_solve_warp.cuh) and the C ABI of the solve set
// (include/lompc_b200.h: lompc_set_*): several LoMPC objects, one launch, one copy each way.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "lompc_common.cuh"
#include "lompc_solve_warp.cuh"

#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return lompc_detail::cuda_fail(e__, #call); \
  } while (0)

namespace {

template <int N, int SPL>
int launch_warp_geom(lompc::WarpArgs& wa, cudaStream_t s) {
  constexpr int QPW = 32 / (N / SPL);
  int warps = 0;
  for (int i = 0; i < wa.nsegs; ++i) {
    wa.seg[i].warp_begin = warps;
    warps += (int)((wa.seg[i].a.B + QPW - 1) / QPW);
  }
  wa.total_warps = warps;
  if (warps == 0) return LOMPC_OK;
  lompc::lompc_solve_warp_kernel<N, SPL><<<warps, 32, 0, s>>>(wa);  // one warp per CTA (see the kernel)
  lompc_detail::count_launch();
  CK(cudaGetLastError());
  return LOMPC_OK;
}

__global__ void status_summary_kernel(const int32_t* __restrict__ st, int64_t n, const unsigned long long* epoch_src,
                                      unsigned long long* summary) {
  int worst = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    worst = max(worst, st[i]);
  worst = __reduce_max_sync(0xffffffffu, worst);
  if ((threadIdx.x & 31) == 0 && (worst != 0 || (blockIdx.x == 0 && threadIdx.x == 0))) {
    const unsigned long long ep = epoch_src ? *epoch_src : 0ull;
    atomicMax(summary, ep * 4ull + (unsigned long long)worst);
  }
}

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

namespace lompc_detail {

// 3 stages per lane: N / 3 = 4, 8, 16, 32 lanes per QP.  (6 stages per lane were measured slower at every
// horizon - the serial part of a lane grows faster than the scans shrink - and are not compiled.)
bool warp_kernel_supports(int N, int spl) { return spl == 3 && (N == 12 || N == 24 || N == 48 || N == 96); }

int launch_k1_warp(int device, int N, int spl, lompc::WarpArgs& wa, cudaStream_t s) {
  (void)device;
  if (spl == 3) {
    switch (N) {
      case 12: return launch_warp_geom<12, 3>(wa, s);
      case 24: return launch_warp_geom<24, 3>(wa, s);
      case 48: return launch_warp_geom<48, 3>(wa, s);
      case 96: return launch_warp_geom<96, 3>(wa, s);
    }
  }
  return LOMPC_ERR_ARG;
}

}  // namespace lompc_detail

// ------------------------------------------------------------------------------------------------
// Solve set
// ------------------------------------------------------------------------------------------------
struct lompc_set {
  int n = 0;
  int N = 0;
  int device = 0;
  lompc_t* hs[lompc::kMaxWarpSegs] = {};
  int64_t B[lompc::kMaxWarpSegs] = {};
  int64_t total = 0;
  // byte offsets inside the in / out / info blocks
  size_t o_lm[lompc::kMaxWarpSegs], o_lr[lompc::kMaxWarpSegs], o_ga[lompc::kMaxWarpSegs];
  size_t o_w[lompc::kMaxWarpSegs], o_c[lompc::kMaxWarpSegs];
  size_t o_st[lompc::kMaxWarpSegs], o_it[lompc::kMaxWarpSegs], o_kk[lompc::kMaxWarpSegs];
  size_t in_bytes = 0, out_bytes = 0, info_bytes = 0;
  char *h_in = nullptr, *h_out = nullptr, *h_info = nullptr;  // pinned
  char *d_in = nullptr, *d_out = nullptr, *d_info = nullptr;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t graph[2] = {nullptr, nullptr};  // [want_info]
  bool no_graph = false;
  bool mapped = true;  // the kernel reads / writes the pinned host blocks directly (no copy engine); LOMPC_SET_MAPPED=0: staged
  unsigned long long epoch = 0;
  bool pending = false;
  bool pending_mapped = false;  // the call in flight wrote its per-QP status straight into the pinned info block
  // solver options / kernel choice of the handles when graph[] was captured (they are kernel arguments)
  double sig_tol[2][lompc::kMaxWarpSegs] = {};
  int sig_iter[2][lompc::kMaxWarpSegs] = {}, sig_var[2][lompc::kMaxWarpSegs] = {};
};

namespace {

// Enqueues the kernels of one call on `s`: device blocks -> device blocks, or (host_blocks) the pinned host blocks
// themselves - they are mapped into the device's address space, so the kernel's loads ARE the host->device transfer
// and its stores the transfer back (per-QP status into the pinned info block, reduced by the host).
int set_launch(lompc_set* S, int want_info, cudaStream_t s, bool host_blocks = false, char* in_at = nullptr,
               char* out_at = nullptr) {
  const int N = S->N;
  // in_at / out_at: caller-owned device blocks with the set's layout (lompc_set_solve_dev_at)
  char* const in = in_at ? in_at : (host_blocks ? S->h_in : S->d_in);
  char* const outb = out_at ? out_at : (host_blocks ? S->h_out : S->d_out);
  char* const info_b = host_blocks ? S->h_info : S->d_info;
  const unsigned long long* epoch_src = reinterpret_cast<const unsigned long long*>(in_at ? in_at : S->d_in);
  unsigned long long* summary = host_blocks ? nullptr : reinterpret_cast<unsigned long long*>(outb);
  auto fill = [&](int i, lompc::SolveArgs& a, bool info) {
    const lompc_detail::HandleView v = lompc_detail::handle_view(S->hs[i]);
    memset(&a, 0, sizeof(a));
    a.B = S->B[i];
    a.lmbd = reinterpret_cast<const double*>(in + S->o_lm[i]);
    a.lmbd_stride = 3 * (int64_t)N;
    a.lmbd_r = reinterpret_cast<const double*>(in + S->o_lr[i]);
    a.lmbd_r_stride = 1;
    a.gamma = reinterpret_cast<const double*>(in + S->o_ga[i]);
    a.w_out = reinterpret_cast<double*>(outb + S->o_w[i]);
    a.cost_out = reinterpret_cast<double*>(outb + S->o_c[i]);
    if (info || host_blocks) a.status = reinterpret_cast<int32_t*>(info_b + S->o_st[i]);
    if (info) {
      a.iters = reinterpret_cast<int32_t*>(info_b + S->o_it[i]);
      a.kkt_res = reinterpret_cast<double*>(info_b + S->o_kk[i]);
    }
    a.max_iter = v.max_iter;
    a.tol = v.tol;
  };
  // The warp-cooperative kernel serves every segment in ONE launch; batches large enough to fill the GPU with
  // one QP per thread (or a horizon it is not compiled for) take one thread-kernel launch per segment.
  bool forced_thread = false;
  const int spl = 3;
  for (int i = 0; i < S->n; ++i) {
    const int var = lompc_detail::handle_view(S->hs[i]).variant;
    if (var >= 1 && var <= 7) forced_thread = true;
  }
  // one warp-kernel launch for all segments while that beats one thread-kernel launch per segment
  const int64_t warp_limit = (N == 12 || N == 24) ? 16384 : INT64_MAX;
  const bool warp_ok = lompc_detail::warp_kernel_supports(N, spl) && !forced_thread && S->total <= warp_limit;
  if (host_blocks && !warp_ok) return LOMPC_ERR_ARG;  // (the caller falls back to the staged round trip)
  if (warp_ok) {
    lompc::WarpArgs wa;
    memset(&wa, 0, sizeof(wa));
    wa.nsegs = S->n;
    wa.epoch_src = epoch_src;
    wa.summary = summary;
    for (int i = 0; i < S->n; ++i) {
      wa.seg[i].cs = *lompc_detail::handle_view(S->hs[i]).cs;
      fill(i, wa.seg[i].a, want_info != 0);
    }
    return lompc_detail::launch_k1_warp(S->device, N, spl, wa, s);
  }
  for (int i = 0; i < S->n; ++i) {
    if (S->B[i] == 0) continue;
    lompc::SolveArgs a;
    fill(i, a, true);
    int rc = lompc_detail::launch_k1(S->hs[i], a, s);
    if (rc) return rc;
  }
  // worst status of all segments (the status arrays of the info block are contiguous per segment)
  for (int i = 0; i < S->n; ++i) {
    if (S->B[i] == 0) continue;
    const int64_t n = S->B[i];
    const int blocks = (int)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    status_summary_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const int32_t*>(S->d_info + S->o_st[i]), n,
                                                 epoch_src, summary);
    lompc_detail::count_launch();
  }
  CK(cudaGetLastError());
  return LOMPC_OK;
}

bool set_uses_host_blocks(const lompc_set* S) {
  if (!S->mapped || S->total > 8192 || !lompc_detail::warp_kernel_supports(S->N, 3)) return false;
  for (int i = 0; i < S->n; ++i) {
    const int var = lompc_detail::handle_view(S->hs[i]).variant;
    if (var >= 1 && var <= 7) return false;
  }
  return true;
}

int set_enqueue_round_trip(lompc_set* S, int want_info, cudaStream_t s) {
  if (set_uses_host_blocks(S)) return set_launch(S, want_info, s, true);
  CK(cudaMemcpyAsync(S->d_in, S->h_in, S->in_bytes, cudaMemcpyHostToDevice, s));
  int rc = set_launch(S, want_info, s);
  if (rc) return rc;
  CK(cudaMemcpyAsync(S->h_out, S->d_out, S->out_bytes, cudaMemcpyDeviceToHost, s));
  if (want_info) CK(cudaMemcpyAsync(S->h_info, S->d_info, S->info_bytes, cudaMemcpyDeviceToHost, s));
  return LOMPC_OK;
}

}  // namespace

extern "C" {

int lompc_set_create(lompc_t* const* handles, int n_handles, const int64_t* batch_sizes, lompc_set_t** out) {
  if (!out || !handles || !batch_sizes || n_handles < 1 || n_handles > lompc::kMaxWarpSegs) return LOMPC_ERR_ARG;
  *out = nullptr;
  for (int i = 0; i < n_handles; ++i)
    if (!handles[i] || batch_sizes[i] < 0) return LOMPC_ERR_ARG;
  const lompc_detail::HandleView v0 = lompc_detail::handle_view(handles[0]);
  for (int i = 1; i < n_handles; ++i) {
    const lompc_detail::HandleView v = lompc_detail::handle_view(handles[i]);
    if (v.cs->N != v0.cs->N || v.device != v0.device) return LOMPC_ERR_ARG;
  }
  lompc_set* S = new (std::nothrow) lompc_set();
  if (!S) return LOMPC_ERR_ARG;
  S->n = n_handles;
  S->N = v0.cs->N;
  S->device = v0.device;
  const size_t N = (size_t)S->N;
  size_t in = 256, outb = 256, info = 0;
  for (int i = 0; i < n_handles; ++i) {
    S->hs[i] = handles[i];
    const size_t B = (size_t)batch_sizes[i];
    S->B[i] = batch_sizes[i];
    S->total += batch_sizes[i];
    S->o_lm[i] = in;   in += al256(B * 3 * N * 8);
    S->o_lr[i] = in;   in += al256(B * 8);
    S->o_ga[i] = in;   in += al256(B * 8);
    S->o_w[i] = outb;  outb += al256(B * N * 8);
    S->o_c[i] = outb;  outb += al256(B * 8);
    S->o_st[i] = info; info += al256(B * 4);
    S->o_it[i] = info; info += al256(B * 4);
    S->o_kk[i] = info; info += al256(B * 8);
  }
  S->in_bytes = in;
  S->out_bytes = outb;
  S->info_bytes = info > 0 ? info : 256;
  const char* ng = getenv("LOMPC_SET_NO_GRAPH");
  S->no_graph = ng && ng[0] == '1';
  const char* mp = getenv("LOMPC_SET_MAPPED");
  S->mapped = !(mp && mp[0] == '0');
#define CKS(call)                                       \
  do {                                                  \
    cudaError_t e__ = (call);                           \
    if (e__ != cudaSuccess) {                           \
      lompc_set_destroy(S);                             \
      return lompc_detail::cuda_fail(e__, #call);       \
    }                                                   \
  } while (0)
  CKS(cudaSetDevice(S->device));
  CKS(cudaMallocHost(&S->h_in, S->in_bytes));
  CKS(cudaMallocHost(&S->h_out, S->out_bytes));
  CKS(cudaMallocHost(&S->h_info, S->info_bytes));
  CKS(cudaMalloc(&S->d_in, S->in_bytes));
  CKS(cudaMalloc(&S->d_out, S->out_bytes));
  CKS(cudaMalloc(&S->d_info, S->info_bytes));
  CKS(cudaStreamCreateWithFlags(&S->stream, cudaStreamNonBlocking));
  memset(S->h_in, 0, S->in_bytes);
  memset(S->h_out, 0, S->out_bytes);
  memset(S->h_info, 0, S->info_bytes);
  CKS(cudaMemset(S->d_in, 0, S->in_bytes));
  CKS(cudaMemset(S->d_out, 0, S->out_bytes));
  CKS(cudaMemset(S->d_info, 0, S->info_bytes));
#undef CKS
  *out = S;
  return LOMPC_OK;
}

int lompc_set_destroy(lompc_set_t* S) {
  if (!S) return LOMPC_OK;
  cudaSetDevice(S->device);
  if (S->stream) cudaStreamSynchronize(S->stream);
  for (auto& g : S->graph)
    if (g) cudaGraphExecDestroy(g);
  if (S->stream) cudaStreamDestroy(S->stream);
  if (S->h_in) cudaFreeHost(S->h_in);
  if (S->h_out) cudaFreeHost(S->h_out);
  if (S->h_info) cudaFreeHost(S->h_info);
  if (S->d_in) cudaFree(S->d_in);
  if (S->d_out) cudaFree(S->d_out);
  if (S->d_info) cudaFree(S->d_info);
  delete S;
  return LOMPC_OK;
}

int lompc_set_buffers(lompc_set_t* S, int which, int i, double** lmbd, double** lmbd_r, double** gamma, double** w,
                      double** cost) {
  if (!S || i < 0 || i >= S->n || (which != 0 && which != 1)) return LOMPC_ERR_ARG;
  char* in = which ? S->d_in : S->h_in;
  char* o = which ? S->d_out : S->h_out;
  if (lmbd) *lmbd = reinterpret_cast<double*>(in + S->o_lm[i]);
  if (lmbd_r) *lmbd_r = reinterpret_cast<double*>(in + S->o_lr[i]);
  if (gamma) *gamma = reinterpret_cast<double*>(in + S->o_ga[i]);
  if (w) *w = reinterpret_cast<double*>(o + S->o_w[i]);
  if (cost) *cost = reinterpret_cast<double*>(o + S->o_c[i]);
  return LOMPC_OK;
}

int lompc_set_info_buffers(lompc_set_t* S, int i, int32_t** status, int32_t** iters, double** kkt_res) {
  if (!S || i < 0 || i >= S->n) return LOMPC_ERR_ARG;
  if (status) *status = reinterpret_cast<int32_t*>(S->h_info + S->o_st[i]);
  if (iters) *iters = reinterpret_cast<int32_t*>(S->h_info + S->o_it[i]);
  if (kkt_res) *kkt_res = reinterpret_cast<double*>(S->h_info + S->o_kk[i]);
  return LOMPC_OK;
}

int64_t lompc_set_bytes(const lompc_set_t* S, int which) {
  if (!S) return 0;
  return which == 0 ? (int64_t)S->in_bytes : (int64_t)S->out_bytes;
}

int lompc_set_solve_host_async(lompc_set_t* S, int want_info) {
  if (!S) return LOMPC_ERR_ARG;
  if (S->pending) return LOMPC_ERR_ARG;  // one call in flight per set
  if (S->total == 0) return LOMPC_OK;
  CK(cudaSetDevice(S->device));
  want_info = want_info ? 1 : 0;
  ++S->epoch;
  *reinterpret_cast<unsigned long long*>(S->h_in) = S->epoch;  // travels with the in block
  if (S->no_graph) {
    int rc = set_enqueue_round_trip(S, want_info, S->stream);
    if (rc) return rc;
  } else {
    for (int i = 0; i < S->n && S->graph[want_info]; ++i) {
      const lompc_detail::HandleView v = lompc_detail::handle_view(S->hs[i]);
      if (v.tol != S->sig_tol[want_info][i] || v.max_iter != S->sig_iter[want_info][i] ||
          v.variant != S->sig_var[want_info][i]) {
        cudaGraphExecDestroy(S->graph[want_info]);  // lompc_set_options / _kernel_variant since the capture
        S->graph[want_info] = nullptr;
      }
    }
    if (!S->graph[want_info]) {
      for (int i = 0; i < S->n; ++i) {
        const lompc_detail::HandleView v = lompc_detail::handle_view(S->hs[i]);
        S->sig_tol[want_info][i] = v.tol;
        S->sig_iter[want_info][i] = v.max_iter;
        S->sig_var[want_info][i] = v.variant;
      }
      // capture the round trip once (the epoch is data, not a kernel argument, so the graph never changes)
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(S->stream, cudaStreamCaptureModeThreadLocal));
      int rc = set_enqueue_round_trip(S, want_info, S->stream);
      cudaError_t e = cudaStreamEndCapture(S->stream, &g);
      if (rc) {
        if (g) cudaGraphDestroy(g);
        return rc;
      }
      if (e != cudaSuccess) return lompc_detail::cuda_fail(e, "cudaStreamEndCapture");
      e = cudaGraphInstantiate(&S->graph[want_info], g, 0);
      cudaGraphDestroy(g);
      if (e != cudaSuccess) return lompc_detail::cuda_fail(e, "cudaGraphInstantiate");
    }
    CK(cudaGraphLaunch(S->graph[want_info], S->stream));
  }
  S->pending = true;
  S->pending_mapped = set_uses_host_blocks(S);
  return LOMPC_OK;
}

int lompc_set_wait(lompc_set_t* S) {
  if (!S) return LOMPC_ERR_ARG;
  if (!S->pending) return LOMPC_OK;
  S->pending = false;
  CK(cudaSetDevice(S->device));
  CK(cudaStreamSynchronize(S->stream));
  if (S->pending_mapped) {  // the kernel wrote the per-QP status straight into the pinned info block
    int worst = 0;
    for (int i = 0; i < S->n; ++i) {
      const int32_t* st = reinterpret_cast<const int32_t*>(S->h_info + S->o_st[i]);
      for (int64_t b = 0; b < S->B[i]; ++b) worst = st[b] > worst ? st[b] : worst;
    }
    switch (worst) {
      case LOMPC_ST_OK: return LOMPC_OK;
      case LOMPC_ST_MAXITER: return LOMPC_ERR_NOT_CONVERGED;
      case LOMPC_ST_BAD_GAMMA: return LOMPC_ERR_GAMMA;
      default: return LOMPC_ERR_NEGATIVE;
    }
  }
  const unsigned long long sum = *reinterpret_cast<const unsigned long long*>(S->h_out);
  if ((sum >> 2) != S->epoch) return LOMPC_ERR_NOT_CONVERGED;  // the launch did not report: treat as a failure
  switch ((int)(sum & 3ull)) {
    case LOMPC_ST_OK: return LOMPC_OK;
    case LOMPC_ST_MAXITER: return LOMPC_ERR_NOT_CONVERGED;
    case LOMPC_ST_BAD_GAMMA: return LOMPC_ERR_GAMMA;
    default: return LOMPC_ERR_NEGATIVE;
  }
}

int lompc_set_solve_host(lompc_set_t* S, int want_info) {
  int rc = lompc_set_solve_host_async(S, want_info);
  if (rc) return rc;
  return lompc_set_wait(S);
}

int lompc_set_solve_dev(lompc_set_t* S, int want_info, void* stream) {
  if (!S) return LOMPC_ERR_ARG;
  if (S->total == 0) return LOMPC_OK;
  CK(cudaSetDevice(S->device));
  return set_launch(S, want_info ? 1 : 0, static_cast<cudaStream_t>(stream));
}

int lompc_set_solve_dev_at(lompc_set_t* S, void* in_block, void* out_block, void* stream) {
  if (!S || !in_block || !out_block) return LOMPC_ERR_ARG;
  if (S->total == 0) return LOMPC_OK;
  CK(cudaSetDevice(S->device));
  return set_launch(S, 0, static_cast<cudaStream_t>(stream), false, static_cast<char*>(in_block),
                    static_cast<char*>(out_block));
}

int lompc_set_offsets(const lompc_set_t* S, int i, int64_t* offsets) {
  if (!S || i < 0 || i >= S->n || !offsets) return LOMPC_ERR_ARG;
  offsets[0] = (int64_t)S->o_lm[i];
  offsets[1] = (int64_t)S->o_lr[i];
  offsets[2] = (int64_t)S->o_ga[i];
  offsets[3] = (int64_t)S->o_w[i];
  offsets[4] = (int64_t)S->o_c[i];
  return LOMPC_OK;
}

int lompc_set_copy(lompc_set_t* S, int which, void* stream) {
  if (!S || (which != 0 && which != 1)) return LOMPC_ERR_ARG;
  CK(cudaSetDevice(S->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (which == 0) {
    ++S->epoch;
    *reinterpret_cast<unsigned long long*>(S->h_in) = S->epoch;
    CK(cudaMemcpyAsync(S->d_in, S->h_in, S->in_bytes, cudaMemcpyHostToDevice, s));
  } else {
    CK(cudaMemcpyAsync(S->h_out, S->d_out, S->out_bytes, cudaMemcpyDeviceToHost, s));
  }
  return LOMPC_OK;
}

}  // extern "C"
