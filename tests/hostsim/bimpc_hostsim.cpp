// Host compilation of the BiMPC kernel body (csrc/bimpc_solve.cuh with BIMPC_HOSTSIM): one
// "thread" per station, barriers are no-ops.  TEST INFRASTRUCTURE ONLY - lets the CPU test
// suite check the kernel's algorithm against the dense oracle without a GPU; the product
// library never contains this code path.
#define BIMPC_HOSTSIM 1
#include <stdlib.h>
#include "bimpc_solve.cuh"

extern "C" int bimpc_hostsim_solve(int N, int P, double delta, double c_g, double u_g_max, double u_b_max,
                                   double x_max, int cost_type, double theta_s, double theta_l,
                                   double w_max_s, double w_max_l, int S, const double* omega,
                                   const double* Mp_s, const double* Mp_l, const double* beta_s,
                                   const double* beta_l, const double* gamma_sm, const double* gamma_lm,
                                   const double* x0, const double* demand, double* w_hat_s, double* w_hat_l,
                                   double* u_g, int32_t* status, int32_t* iters, double* objective,
                                   double tol, int max_iter) {
  if (N > bimpc::kMaxN) return -2;
  bimpc::BiConsts c{N, P, delta, c_g, u_g_max, u_b_max, x_max, cost_type, theta_s, theta_l, w_max_s, w_max_l};
  bimpc::BiArgs a{S, omega, Mp_s, Mp_l, beta_s, beta_l, gamma_sm, gamma_lm, x0, demand,
                  w_hat_s, w_hat_l, u_g, status, iters, objective, tol, max_iter, nullptr, nullptr};
  const size_t n = bimpc::scratch_doubles(N, P, 1);
  double* sm = (double*)malloc(n * sizeof(double));
  double* li = (double*)malloc(bimpc::global_scratch_doubles(N, P) * sizeof(double));
  if (!sm || !li) return -1;
  for (int s = 0; s < S; ++s) bimpc::solve_station(c, a, s, sm, li, 0, 1);
  free(sm);
  free(li);
  return 0;
}
