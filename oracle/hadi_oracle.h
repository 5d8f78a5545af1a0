/* TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's Heston ADI hot path.
 *
 * This is the parity oracle of the hadi project.  It is NOT part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * Every function cites the reference file:line whose arithmetic (operation order included) it
 * restates.  Parity status: PINNED — checked bit-for-bit against the reference's own sources
 * compiled here (oracle/_ref, see oracle/ref_driver.cpp) by tests/test_oracle_vs_ref.py and against
 * the committed fixtures in tests/golden/ (generated from oracle/_ref by oracle/make_golden.py).
 * The one unpinned boundary is the summation order inside KokkosBlas gemm/gemv (un-vendored
 * KokkosKernels, version unpinned in the reference's CMakeLists.txt:8-13): ascending k is used.
 */
#ifndef HADI_ORACLE_H
#define HADI_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  double S0, V0, r_d, r_f;             /* spot, initial variance, domestic / foreign rate */
  double kappa, eta, sigma, rho;        /* Heston parameters */
} ho_model;

typedef struct {
  int m1, m2;                           /* grid intervals in s and v (m1+1, m2+1 nodes) */
  double theta;                         /* ADI theta */
  int style;                            /* 0 European, 1 American (Ikonen-Toivanen) */
  int payoff_put;                       /* 0: max(s-K,0)  1: max(K-s,0) (reference call BCs) */
  int scheme;                           /* 0 Douglas (device path), 1 Craig-Sneyd (host path), 2 Modified Craig-Sneyd
                                           as the reference ships it (src/solver.hpp:917-1075, pinned to oracle/_ref),
                                           3 Hundsdorfer-Verwer (extension, not in the reference: parity unpinned) */
  int nd;                               /* number of dividends */
  const double *div_dates, *div_amounts, *div_pcts;
  /* opt-in extensions beyond the reference's device path (PARITY UNPINNED: the
   * reference has no such code path, the restatement below is the definition both sides are tested against) */
  int bc;                               /* 0 reference call boundary vectors; 1 put-correct set: b1 = b2 = 0 and
                                           Dirichlet U(s_0, v, tau) = K exp(-r_d tau) after every step */
  int div_all;                          /* 0 device schedule (one dividend per step at most, quirk Q7); 1 every
                                           dividend whose date falls in the step, in order (src/solver.hpp:363) */
} ho_numerics;

/* grids: src/grid.cpp:16-96, src/grid_pod.hpp:25-87 */
void ho_grid_s(int m1, double S, double S0, double K, double c, double *s, double *ds);
void ho_grid_v(int m2, double V, double V0, double d, double *v, double *dv);
int ho_find_index(const double *x, int n, double x0);

/* one solve; U_out[P], lambda_out[P] optional.  Returns 0, or -1 if S0 is not a grid node. */
int ho_solve(const ho_model *mdl, const ho_numerics *num, double K, int N, double dt,
             double V0_for_grid, double *price, double *U_out, double *lambda_out);

/* batches: src/jacobian_computation.cpp:204-448 and variants; multi-maturity
 * src/heston_calibration.cpp:2174-2424, 2936-3243.  J is [n][5] in (kappa, eta, sigma, rho, v0). */
int ho_price_batch(const ho_model *mdl, const ho_numerics *num, int n, const double *strikes,
                   const int *Ns, const double *dts, double *prices);
int ho_jacobian_batch(const ho_model *mdl, const ho_numerics *num, int n, const double *strikes,
                      const int *Ns, const double *dts, double eps, double *J, double *base);

/* LM pieces: src/jacobian_computation.cpp:20-195 */
void ho_solve5(const double *A, const double *b, double *x);
void ho_lm_update(int n, const double *J, const double *r, double lambda, double *delta);

typedef struct {
  int max_iter;
  double tol, delta_tol, lambda0, eps;
} ho_lm_opts;
typedef struct {
  double params[5]; /* kappa, eta, sigma, rho, v0 */
  double final_error, lambda, delta_norm;
  int iterations, converged, pde_solves;
} ho_lm_result;
/* LM loop: src/heston_calibration.cpp:2692-2831 */
int ho_calibrate(const ho_model *mdl0, const ho_numerics *num, int n, const double *strikes,
                 const int *Ns, const double *dts, const double *market, const ho_lm_opts *opt,
                 ho_lm_result *res);

/* src/bs.hpp:44-55 */
double ho_bs_call(double S, double K, double r, double vol, double T);

#ifdef __cplusplus
}
#endif
#endif
