/*
 * lompc_b200.h -- C ABI of the B200-native lower-level MPC (LoMPC) hot path.
 *
 * The reference (AkshayThiru/incentive-design-mpc) has no FFI of its own: the
 * boundary of this path is its Python class API.  Each
 * entry point below names the reference method it replaces (file:line relative
 * to the reference tree).  The Python mirror of those classes, in
 * incentive-design-mpc_b200/chargingstation/, binds this header with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions: plain C, fp64, row-major, B = batch.  Functions return 0 on
 * success or a negative LOMPC_ERR_* code; they never throw.  "_dev" entry
 * points take DEVICE pointers and are asynchronous on `stream` (a
 * cudaStream_t passed as void*); "_host" entry points take HOST pointers and
 * do the host<->device copies themselves and synchronise before returning.
 * A handle is bound to one (EV type, horizon N, device); it is not
 * thread-safe (neither are the reference's objects).
 */
#ifndef LOMPC_B200_H
#define LOMPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOMPC_OK 0
#define LOMPC_ERR_CONSTS (-1)      /* lompc.py:36-38 asserts  -> AssertionError */
#define LOMPC_ERR_ARG (-2)         /* bad pointer / size / shape */
#define LOMPC_ERR_CUDA (-3)        /* CUDA runtime failure (see lompc_last_cuda_error) */
#define LOMPC_ERR_GAMMA (-4)       /* gamma > y_max, lompc.py:87 -> AssertionError */
#define LOMPC_ERR_NEGATIVE (-5)    /* nonneg cv.Parameter got < 0, lompc.py:78-82 -> ValueError */
#define LOMPC_ERR_NOT_CONVERGED (-6)
#define LOMPC_ERR_NO_DEVICE (-7)   /* no CUDA device: there is NO CPU fallback */

#define LOMPC_EV_SMALL 0
#define LOMPC_EV_LARGE 1

/* per-QP status written by the solve kernels */
#define LOMPC_ST_OK 0
#define LOMPC_ST_MAXITER 1
#define LOMPC_ST_BAD_GAMMA 2
#define LOMPC_ST_NEGATIVE 3

typedef struct lompc_handle lompc_t;

/* Library / build information. */
const char* lompc_version(void);
const char* lompc_strerror(int code);
const char* lompc_last_cuda_error(void);
int lompc_device_count(void);

/* Replaces LoMPC.__init__ / _set_constants (lompc.py:30-71): validates the
 * constants exactly like lompc.py:36-38 (LOMPC_ERR_CONSTS), derives q_scale
 * (lompc.py:67) and the strong-convexity modulus m (lompc.py:71).           */
int lompc_create(int N, double delta, double theta, double y_max, double w_max,
                 int ev_type, int device, lompc_t** out);
int lompc_destroy(lompc_t* h);

/* LoMPC.get_sc_modulus (lompc.py:158). */
double lompc_sc_modulus(const lompc_t* h);

/* Solver knobs (defaults: max_iter 200, tol 1e-11 relative KKT residual). */
int lompc_set_options(lompc_t* h, int max_iter, double tol);

/* Replaces LoMPC.solve_lompc (lompc.py:137-156), batched over B independent
 * QPs.  lmbd: B rows of 3N prices (row stride lmbd_stride doubles; 0 =
 * broadcast ONE price vector to the whole batch, the _get_w_err case
 * price_solver.py:203-204).  lmbd_r: stride 0 = one scalar for all.
 * gamma[B].  Outputs: w_out[B,N], cost_out[B] (the full objective including
 * theta*w_max*sum(lmbd2), as self.cost.value lompc.py:155); optional (may be
 * NULL) status[B], iters[B], kkt_res[B] (relative KKT residual).             */
int lompc_solve_batch_dev(lompc_t* h, int64_t B, const double* lmbd, int64_t lmbd_stride,
                          const double* lmbd_r, int64_t lmbd_r_stride, const double* gamma,
                          double* w_out, double* cost_out, int32_t* status, int32_t* iters,
                          double* kkt_res, void* stream);

/* Same call with HOST buffers: copies in, solves, copies out, synchronises.
 * Returns LOMPC_ERR_GAMMA / LOMPC_ERR_NEGATIVE / LOMPC_ERR_NOT_CONVERGED if
 * any QP reported that status (outputs are still written).                  */
int lompc_solve_batch_host(lompc_t* h, int64_t B, const double* lmbd, int64_t lmbd_stride,
                           const double* lmbd_r, int64_t lmbd_r_stride, const double* gamma,
                           double* w_out, double* cost_out, int32_t* status, int32_t* iters,
                           double* kkt_res);

/* Measures the device's FP64 FMA peak (TFLOP/s, FMA = 2 flops) with a
 * register-resident DFMA chain kernel: the roofline denominator of this
 * FP64-bound path (MEASURED_PEAKS.json has no FP64 figure).                 */
int lompc_measure_fp64_peak(int device, int iters, double* tflops_out, double* ms_out);

/* Number of kernels this library has launched since load (bench.py's
 * gpu_launches claim is read from here).                                    */
int64_t lompc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* LOMPC_B200_H */
