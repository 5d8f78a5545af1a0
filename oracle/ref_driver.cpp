// TEST INFRASTRUCTURE — not part of the product; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load the library built from this file.
//
// C-ABI driver around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference/src against oracle/kokkos_shim).  Every number this library returns is
// produced by the reference's own functions:
//   compute_base_prices{,_american,_dividends,_american_dividends}   src/jacobian_computation.cpp:368,629,922,1232
//   compute_jacobian{,_american,_dividends,_american_dividends}      src/jacobian_computation.cpp:204,457,726,1031
//   compute_{jacobian,base_prices}_multi_maturity                     src/heston_calibration.cpp:2174,2339
//   compute_{jacobian,base_prices}_multi_maturity_american_dividends  src/heston_calibration.cpp:2936,3140
//   compute_parameter_update_on_device / solve_5x5_device            src/jacobian_computation.cpp:107,20
//   CS_scheme_shuffled / DO_scheme_shuffle (host-driven)              src/solver.hpp:781,98
//   Grid::Grid                                                        src/grid.cpp:16
//   BlackScholes::call_price                                          src/bs.hpp:44
// The set-up code below (allocation of solver/grid/workspace views, payoff fill) restates the
// reference's own test drivers, e.g. src/device_solver.cpp:1330-1420 and
// src/heston_calibration.cpp:2563-2660; it contains no solver arithmetic.
#include <Kokkos_Core.hpp>

#include <chrono>
#include <cstdio>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <iostream>
#include <sstream>
#include <cstring>
#include <vector>

#include "bs.hpp"
#include "grid.hpp"
#include "grid_pod.hpp"
#include "jacobian_computation.hpp"
#include "solver.hpp"

using Device = Kokkos::DefaultExecutionSpace;

// Must match src/heston_calibration.cpp:2165-2171 (defined in the .cpp, not in a header).
struct CalibrationPoint {
  double strike;
  double maturity;
  int time_steps;
  double delta_t;
  int global_index;
};

// Defined (not declared in any header) in src/heston_calibration.cpp.
void compute_jacobian_multi_maturity(
    const double S_0, const double V_0, const double r_d, const double r_f, const double rho,
    const double sigma, const double kappa, const double eta, const int m1, const int m2,
    const int total_size, const double theta,
    const Kokkos::View<CalibrationPoint*>& d_calibration_points, const int total_calibration_size,
    const Kokkos::View<Device_A0_heston<Kokkos::DefaultExecutionSpace>*>& A0_solvers,
    const Kokkos::View<Device_A1_heston<Kokkos::DefaultExecutionSpace>*>& A1_solvers,
    const Kokkos::View<Device_A2_shuffled_heston<Kokkos::DefaultExecutionSpace>*>& A2_solvers,
    const Kokkos::View<Device_BoundaryConditions<Kokkos::DefaultExecutionSpace>*>& bounds_d,
    const Kokkos::View<GridViews*>& deviceGrids, const Kokkos::View<double**>& U_0,
    DO_Workspace<Kokkos::DefaultExecutionSpace>& workspace, Kokkos::View<double**>& J,
    Kokkos::View<double*>& base_prices,
    const Kokkos::TeamPolicy<Kokkos::DefaultExecutionSpace>& policy, const double eps);
void compute_base_prices_multi_maturity(
    const double S_0, const double V_0, const double r_d, const double r_f, const double rho,
    const double sigma, const double kappa, const double eta, const int m1, const int m2,
    const int total_size, const double theta,
    const Kokkos::View<CalibrationPoint*>& d_calibration_points, const int total_calibration_size,
    const Kokkos::View<Device_A0_heston<Kokkos::DefaultExecutionSpace>*>& A0_solvers,
    const Kokkos::View<Device_A1_heston<Kokkos::DefaultExecutionSpace>*>& A1_solvers,
    const Kokkos::View<Device_A2_shuffled_heston<Kokkos::DefaultExecutionSpace>*>& A2_solvers,
    const Kokkos::View<Device_BoundaryConditions<Kokkos::DefaultExecutionSpace>*>& bounds_d,
    const Kokkos::View<GridViews*>& deviceGrids,
    DO_Workspace<Kokkos::DefaultExecutionSpace>& workspace, Kokkos::View<double*>& base_prices,
    const Kokkos::TeamPolicy<Kokkos::DefaultExecutionSpace>& policy);
void compute_jacobian_multi_maturity_american_dividends(
    const double S_0, const double V_0, const double r_d, const double r_f, const double rho,
    const double sigma, const double kappa, const double eta, const int m1, const int m2,
    const int total_size, const double theta,
    const Kokkos::View<CalibrationPoint*>& d_calibration_points, const int total_calibration_size,
    const Kokkos::View<Device_A0_heston<Kokkos::DefaultExecutionSpace>*>& A0_solvers,
    const Kokkos::View<Device_A1_heston<Kokkos::DefaultExecutionSpace>*>& A1_solvers,
    const Kokkos::View<Device_A2_shuffled_heston<Kokkos::DefaultExecutionSpace>*>& A2_solvers,
    const Kokkos::View<Device_BoundaryConditions<Kokkos::DefaultExecutionSpace>*>& bounds_d,
    const Kokkos::View<GridViews*>& deviceGrids, const Kokkos::View<double**>& U_0,
    DO_Workspace<Kokkos::DefaultExecutionSpace>& workspace, const int num_dividends,
    const Kokkos::View<double*>& dividend_dates, const Kokkos::View<double*>& dividend_amounts,
    const Kokkos::View<double*>& dividend_percentages, Kokkos::View<double**>& J,
    Kokkos::View<double*>& base_prices,
    const Kokkos::TeamPolicy<Kokkos::DefaultExecutionSpace>& policy, const double eps);
void compute_base_prices_multi_maturity_american_dividends(
    const double S_0, const double V_0, const double r_d, const double r_f, const double rho,
    const double sigma, const double kappa, const double eta, const int m1, const int m2,
    const int total_size, const double theta,
    const Kokkos::View<CalibrationPoint*>& d_calibration_points, const int total_calibration_size,
    const Kokkos::View<Device_A0_heston<Kokkos::DefaultExecutionSpace>*>& A0_solvers,
    const Kokkos::View<Device_A1_heston<Kokkos::DefaultExecutionSpace>*>& A1_solvers,
    const Kokkos::View<Device_A2_shuffled_heston<Kokkos::DefaultExecutionSpace>*>& A2_solvers,
    const Kokkos::View<Device_BoundaryConditions<Kokkos::DefaultExecutionSpace>*>& bounds_d,
    const Kokkos::View<GridViews*>& deviceGrids, const Kokkos::View<double**>& U_0,
    DO_Workspace<Kokkos::DefaultExecutionSpace>& workspace, const int num_dividends,
    const Kokkos::View<double*>& dividend_dates, const Kokkos::View<double*>& dividend_amounts,
    const Kokkos::View<double*>& dividend_percentages, Kokkos::View<double*>& base_prices,
    const Kokkos::TeamPolicy<Kokkos::DefaultExecutionSpace>& policy);

// Shipped LM drivers (print to stdout, write CSVs into the cwd).
void test_calibration_european_multi_maturity();
void test_calibration_european();
void test_calibration_american_divident_multi_maturity();

namespace {

struct Problem {
  int n, m1, m2, P;
  Kokkos::View<Device_A0_heston<Device>*> A0;
  Kokkos::View<Device_A1_heston<Device>*> A1;
  Kokkos::View<Device_A2_shuffled_heston<Device>*> A2;
  Kokkos::View<Device_BoundaryConditions<Device>*> bounds;
  std::vector<GridViews> hostGrids;
  Kokkos::View<GridViews*> deviceGrids;
  Kokkos::View<double**> U_0;
  Kokkos::View<CalibrationPoint*> points;
  DO_Workspace<Device>* ws = nullptr;
  ~Problem() { delete ws; }
};

// Mirrors src/heston_calibration.cpp:2563-2660 / src/device_solver.cpp:1330-1412.
void build_problem(Problem& p, int n, const double* strikes, const double* maturities,
                   const int* Ns, const double* dts, double S_0, double V_0_grid, double r_d,
                   double r_f, int m1, int m2, int payoff_put) {
  p.n = n;
  p.m1 = m1;
  p.m2 = m2;
  p.P = (m1 + 1) * (m2 + 1);
  p.A0 = Kokkos::View<Device_A0_heston<Device>*>("A0_solvers", n);
  p.A1 = Kokkos::View<Device_A1_heston<Device>*>("A1_solvers", n);
  p.A2 = Kokkos::View<Device_A2_shuffled_heston<Device>*>("A2_solvers", n);
  p.bounds = Kokkos::View<Device_BoundaryConditions<Device>*>("bounds_d", n);
  p.points = Kokkos::View<CalibrationPoint*>("points", n);
  for (int i = 0; i < n; ++i) {
    p.A0(i) = Device_A0_heston<Device>(m1, m2);
    p.A1(i) = Device_A1_heston<Device>(m1, m2);
    p.A2(i) = Device_A2_shuffled_heston<Device>(m1, m2);
    p.bounds(i) = Device_BoundaryConditions<Device>(m1, m2, r_d, r_f, Ns[i], dts[i]);
    p.points(i) = CalibrationPoint{strikes[i], maturities ? maturities[i] : Ns[i] * dts[i], Ns[i],
                                   dts[i], i};
  }
  buildMultipleGridViews(p.hostGrids, n, m1, m2);
  p.deviceGrids = Kokkos::View<GridViews*>("deviceGrids", n);
  p.U_0 = Kokkos::View<double**>("U_0", n, p.P);
  for (int i = 0; i < n; ++i) {
    const double K = strikes[i];
    Grid g(m1, 8 * K, S_0, K, K / 5, m2, 5.0, V_0_grid, 5.0 / 500);
    for (int j = 0; j <= m1; j++) p.hostGrids[i].device_Vec_s(j) = g.Vec_s[j];
    for (int j = 0; j <= m2; j++) p.hostGrids[i].device_Vec_v(j) = g.Vec_v[j];
    for (int j = 0; j < m1; j++) p.hostGrids[i].device_Delta_s(j) = g.Delta_s[j];
    for (int j = 0; j < m2; j++) p.hostGrids[i].device_Delta_v(j) = g.Delta_v[j];
    p.deviceGrids(i) = p.hostGrids[i];
    for (int j = 0; j <= m2; j++)
      for (int k = 0; k <= m1; k++) {
        const double s = g.Vec_s[k];
        p.U_0(i, k + j * (m1 + 1)) = payoff_put ? std::max(K - s, 0.0) : std::max(s - K, 0.0);
      }
  }
  p.ws = new DO_Workspace<Device>(n, p.P);
  Kokkos::deep_copy(p.ws->U, p.U_0);
}

}  // namespace

namespace {
double g_last_compute_s = 0.0;   // wall time of the reference entry point alone in the last hadi_ref_solve_batch
}

extern "C" {

// Seconds the reference's own entry point (compute_base_prices* / compute_jacobian*) took in the last
// hadi_ref_solve_batch call: the problem set-up of this driver (grids, U_0, workspace) is outside it.
double hadi_ref_last_compute_seconds() { return g_last_compute_s; }
// Host threads the league loop of the Kokkos stand-in spreads over (1 in the serial build).
int hadi_ref_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Overrides OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1 to its children); returns the count in effect.
int hadi_ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

// Reference grid (src/grid.cpp:16-96).  Outputs: s[m1+1], ds[m1], v[m2+1], dv[m2].
int hadi_ref_grid(int m1, double S, double S_0, double K, double c, int m2, double V, double V_0,
                  double d, double* s, double* ds, double* v, double* dv) {
  Grid g(m1, S, S_0, K, c, m2, V, V_0, d);
  for (int i = 0; i <= m1; i++) s[i] = g.Vec_s[i];
  for (int i = 0; i < m1; i++) ds[i] = g.Delta_s[i];
  for (int i = 0; i <= m2; i++) v[i] = g.Vec_v[i];
  for (int i = 0; i < m2; i++) dv[i] = g.Delta_v[i];
  return 0;
}

// Batched prices / Jacobian through the reference entry points.
//   style: 0 European, 1 American;  nd: number of dividends (0 = none)
//   multi: 0 -> single-maturity functions (all N/dt must be equal), 1 -> multi-maturity functions
//          (only European without dividends and American with dividends exist in the reference)
//   jac:   0 -> compute_base_prices*, 1 -> compute_jacobian*
//   out_prices[n]; out_J[n*5] (jac only, may be NULL); out_U[n*P], out_lambda[n*P] may be NULL.
int hadi_ref_solve_batch(int n, const double* strikes, const double* maturities, const int* Ns,
                         const double* dts, double S_0, double V_0, double r_d, double r_f,
                         double rho, double sigma, double kappa, double eta, int m1, int m2,
                         double theta, int style, int payoff_put, int nd, const double* div_dates,
                         const double* div_amounts, const double* div_pcts, int multi, int jac,
                         double eps, double V_0_grid, double* out_prices, double* out_J,
                         double* out_U, double* out_lambda) {
  Problem p;
  build_problem(p, n, strikes, maturities, Ns, dts, S_0, V_0_grid, r_d, r_f, m1, m2, payoff_put);
  const int P = p.P;
  Kokkos::View<double**> J("J", n, 5);
  Kokkos::View<double*> base("base", n);
  const int ndv = nd > 0 ? nd : 1;
  Kokkos::View<double*> dd("dd", ndv), da("da", ndv), dp("dp", ndv);
  for (int k = 0; k < nd; ++k) {
    dd(k) = div_dates[k];
    da(k) = div_amounts[k];
    dp(k) = div_pcts[k];
  }
  const int N = Ns[0];
  const double dt = dts[0];
  const double T = maturities ? maturities[0] : N * dt;
  if (!multi) {
    for (int i = 1; i < n; ++i)
      if (Ns[i] != N || dts[i] != dt) return -2;
  }
  Kokkos::TeamPolicy<Device> policy(n, Kokkos::AUTO);
  const bool div = nd > 0;
  const auto t_compute0 = std::chrono::steady_clock::now();
  if (multi) {
    if (style == 0 && !div) {
      if (jac)
        compute_jacobian_multi_maturity(S_0, V_0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, theta,
                                        p.points, n, p.A0, p.A1, p.A2, p.bounds, p.deviceGrids,
                                        p.U_0, *p.ws, J, base, policy, eps);
      else
        compute_base_prices_multi_maturity(S_0, V_0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P,
                                           theta, p.points, n, p.A0, p.A1, p.A2, p.bounds,
                                           p.deviceGrids, *p.ws, base, policy);
    } else if (style == 1) {
      if (jac)
        compute_jacobian_multi_maturity_american_dividends(
            S_0, V_0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, theta, p.points, n, p.A0, p.A1,
            p.A2, p.bounds, p.deviceGrids, p.U_0, *p.ws, nd, dd, da, dp, J, base, policy, eps);
      else
        compute_base_prices_multi_maturity_american_dividends(
            S_0, V_0, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, theta, p.points, n, p.A0, p.A1,
            p.A2, p.bounds, p.deviceGrids, p.U_0, *p.ws, nd, dd, da, dp, base, policy);
    } else {
      return -3;
    }
  } else if (style == 0 && !div) {
    if (jac)
      compute_jacobian(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, N, theta, dt, n,
                       p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, p.U_0, *p.ws, J, base, eps);
    else
      compute_base_prices(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, N, theta, dt, n,
                          p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, *p.ws, base);
  } else if (style == 1 && !div) {
    if (jac)
      compute_jacobian_american(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, N, theta,
                                dt, n, p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, p.U_0, *p.ws, J,
                                base, eps);
    else
      compute_base_prices_american(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, N,
                                   theta, dt, n, p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, p.U_0,
                                   *p.ws, base);
  } else if (style == 0 && div) {
    if (jac)
      compute_jacobian_dividends(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, N, theta,
                                 dt, n, p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, p.U_0, *p.ws, nd,
                                 dd, da, dp, J, base, eps);
    else
      compute_base_prices_dividends(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P, N,
                                    theta, dt, n, p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, p.U_0,
                                    *p.ws, nd, dd, da, dp, base);
  } else {
    if (jac)
      compute_jacobian_american_dividends(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2, P,
                                          N, theta, dt, n, p.A0, p.A1, p.A2, p.bounds,
                                          p.deviceGrids, p.U_0, *p.ws, nd, dd, da, dp, J, base, eps);
    else
      compute_base_prices_american_dividends(S_0, V_0, T, r_d, r_f, rho, sigma, kappa, eta, m1, m2,
                                             P, N, theta, dt, n, p.A0, p.A1, p.A2, p.bounds,
                                             p.deviceGrids, p.U_0, *p.ws, nd, dd, da, dp, base);
  }
  g_last_compute_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_compute0).count();
  for (int i = 0; i < n; ++i) out_prices[i] = base(i);
  if (jac && out_J)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < 5; ++k) out_J[i * 5 + k] = J(i, k);
  if (out_U)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < P; ++k) out_U[(size_t)i * P + k] = p.ws->U(i, k);
  if (out_lambda)
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < P; ++k) out_lambda[(size_t)i * P + k] = p.ws->lambda_bar(i, k);
  return 0;
}

// Operator dump for one option after a European base-price call (debug aid for the
// restatement): A1 explicit diagonals [m2+1][m1+1] (lower/upper padded with a trailing 0),
// A2 explicit diagonals of column 0 [5][m2+1] (order lower2, lower, main, upper, upper2; padded),
// b, b1, b2 [P].
int hadi_ref_dump_operators(double K, int N, double dt, double S_0, double V_0, double r_d,
                            double r_f, double rho, double sigma, double kappa, double eta, int m1,
                            int m2, double theta, double* a1_lower, double* a1_main,
                            double* a1_upper, double* a2_diags, double* b, double* b1, double* b2,
                            double* a0_values) {
  Problem p;
  double mat = N * dt;
  build_problem(p, 1, &K, &mat, &N, &dt, S_0, V_0, r_d, r_f, m1, m2, 0);
  Kokkos::View<double*> base("base", 1);
  compute_base_prices(S_0, V_0, mat, r_d, r_f, rho, sigma, kappa, eta, m1, m2, p.P, N, theta, dt, 1,
                      p.A0, p.A1, p.A2, p.bounds, p.deviceGrids, *p.ws, base);
  auto& A1 = p.A1(0);
  auto& A2 = p.A2(0);
  for (int j = 0; j <= m2; ++j)
    for (int i = 0; i <= m1; ++i) {
      a1_main[j * (m1 + 1) + i] = A1.main_diags(j, i);
      a1_lower[j * (m1 + 1) + i] = i < m1 ? A1.lower_diags(j, i) : 0.0;
      a1_upper[j * (m1 + 1) + i] = i < m1 ? A1.upper_diags(j, i) : 0.0;
    }
  for (int j = 0; j <= m2; ++j) {
    a2_diags[0 * (m2 + 1) + j] = j < m2 - 1 ? A2.lower2_diags(0, j) : 0.0;
    a2_diags[1 * (m2 + 1) + j] = j < m2 ? A2.lower_diags(0, j) : 0.0;
    a2_diags[2 * (m2 + 1) + j] = A2.main_diags(0, j);
    a2_diags[3 * (m2 + 1) + j] = j < m2 ? A2.upper_diags(0, j) : 0.0;
    a2_diags[4 * (m2 + 1) + j] = j < m2 - 1 ? A2.upper2_diags(0, j) : 0.0;
  }
  auto& bd = p.bounds(0);
  for (int k = 0; k < p.P; ++k) {
    b[k] = bd.b_(k);
    b1[k] = bd.b1_(k);
    b2[k] = bd.b2_(k);
  }
  if (a0_values)
    for (int j = 0; j < m2 - 1; ++j)
      for (int k = 0; k < (m1 - 1) * 9; ++k) a0_values[j * (m1 - 1) * 9 + k] = p.A0(0).values(j, k);
  return 0;
}

// LM normal-equation update (src/jacobian_computation.cpp:107-195).
int hadi_ref_lm_update(int n, const double* J, const double* r, double lambda, double* delta) {
  Kokkos::View<double**> Jv("J", n, 5);
  Kokkos::View<double*> rv("r", n), dv("delta", 5);
  for (int i = 0; i < n; ++i) {
    rv(i) = r[i];
    for (int k = 0; k < 5; ++k) Jv(i, k) = J[i * 5 + k];
  }
  compute_parameter_update_on_device(Jv, rv, lambda, dv);
  for (int k = 0; k < 5; ++k) delta[k] = dv(k);
  return 0;
}

int hadi_ref_solve5(const double* A, const double* b, double* x) {
  Kokkos::View<double**> Av("A", 5, 5);
  Kokkos::View<double*> bv("b", 5), xv("x", 5);
  for (int i = 0; i < 5; ++i) {
    bv(i) = b[i];
    for (int k = 0; k < 5; ++k) Av(i, k) = A[i * 5 + k];
  }
  solve_5x5_device(Av, bv, xv);
  for (int k = 0; k < 5; ++k) x[k] = xv(k);
  return 0;
}

double hadi_ref_bs_call(double S, double K, double r, double vol, double T) {
  return BlackScholes::call_price(1, S, K, r, vol, T);
}

// Implied-volatility helpers and the synthetic market generators (src/bs.hpp:58-192).  The reference prints
// from inside them; std::cout is silenced for the duration of the call.
namespace {
struct QuietCout {
  std::streambuf* old;
  std::ostringstream sink;
  QuietCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~QuietCout() { std::cout.rdbuf(old); }
};
}  // namespace
double hadi_ref_bs_vega(double S, double K, double r, double vol, double T) {
  return BlackScholes::call_vega(1, S, K, r, vol, T);
}
double hadi_ref_reverse_bs(double S, double K, double r, double T, double v0, double target, double eps) {
  QuietCout q;
  return BlackScholes::reverse_BS(1, S, K, r, T, v0, target, eps);
}
double hadi_ref_reverse_bs_dic(double S, double K, double r, double T, double target, double eps, double a,
                               double b) {
  QuietCout q;
  return BlackScholes::reverse_BS_dic(1, S, K, r, T, target, eps, a, b);
}
// nd = 0: generate_market_data, else generate_market_data_with_dividends; prices[n] for strikes[n] at maturity T
int hadi_ref_market(double S0, double T, double r_d, int n, const double* strikes, int nd, const double* dates,
                    const double* amounts, const double* pcts, double* prices) {
  QuietCout q;
  std::vector<double> k(strikes, strikes + n);
  Kokkos::View<double*> mp("market", n);
  auto h = Kokkos::create_mirror_view(mp);
  if (nd == 0) {
    BlackScholes::generate_market_data(S0, T, r_d, k, h);
  } else {
    std::vector<double> dd(dates, dates + nd), da(amounts, amounts + nd), dp(pcts, pcts + nd);
    BlackScholes::generate_market_data_with_dividends(S0, T, r_d, k, dd, da, dp, h);
  }
  for (int i = 0; i < n; ++i) prices[i] = h(i);
  return 0;
}

// Host-driven schemes of src/solver.hpp on the reference's host matrix classes
// (src/hes_mat_fac.*, src/hes_A2_mat.*, src/BoundaryConditions.*), set up as in
// src/solver.cpp test drivers.  scheme: 0 = DO_scheme_shuffle, 1 = CS_scheme_shuffled.
int hadi_ref_host_scheme(int scheme, double K, double S_0, double V_0, double T, double r_d,
                         double r_f, double rho, double sigma, double kappa, double eta, int m1,
                         int m2, int N, double theta, double* out_price, double* out_U) {
  const int m = (m1 + 1) * (m2 + 1);
  const double delta_t = T / N;
  Grid grid(m1, 8 * K, S_0, K, K / 5, m2, 5.0, V_0, 5.0 / 500);
  heston_A0Storage_gpu A0(m1, m2);
  heston_A1Storage_gpu A1(m1, m2);
  heston_A2_shuffled A2_shuf(m1, m2);
  A0.build_matrix(grid, rho, sigma);
  A1.build_matrix(grid, rho, sigma, r_d, r_f);
  A2_shuf.build_matrix(grid, rho, sigma, r_d, kappa, eta);
  A1.build_implicit(theta, delta_t);
  A2_shuf.build_implicit(theta, delta_t);
  BoundaryConditions bounds(m1, m2, r_d, r_f, N, delta_t);
  bounds.initialize(Kokkos::View<double*>(grid.Vec_s.data(), m1 + 1));
  Kokkos::View<double*> U_0("U_0", m), U("U", m);
  for (int j = 0; j <= m2; j++)
    for (int i = 0; i <= m1; i++) U_0(i + j * (m1 + 1)) = std::max(grid.Vec_s[i] - K, 0.0);
  if (scheme == 0)
    DO_scheme_shuffle<Kokkos::View<double*>>(m, m1, m2, N, U_0, delta_t, theta, A0, A1, A2_shuf,
                                              bounds, r_f, U);
  else if (scheme == 2)   // Modified Craig-Sneyd as the reference ships it (src/solver.hpp:917-1075)
    MCS_scheme_shuffled<Kokkos::View<double*>>(m, m1, m2, N, U_0, delta_t, theta, A0, A1, A2_shuf,
                                                bounds, r_f, U);
  else
    CS_scheme_shuffled<Kokkos::View<double*>>(m, m1, m2, N, U_0, delta_t, theta, A0, A1, A2_shuf,
                                               bounds, r_f, U);
  int index_s = -1, index_v = -1;
  for (int i = 0; i <= m1; i++)
    if (std::abs(grid.Vec_s[i] - S_0) < 1e-10) {
      index_s = i;
      break;
    }
  for (int j = 0; j <= m2; j++)
    if (std::abs(grid.Vec_v[j] - V_0) < 1e-10) {
      index_v = j;
      break;
    }
  if (index_s < 0 || index_v < 0) return -1;
  *out_price = U(index_s + index_v * (m1 + 1));
  if (out_U)
    for (int k = 0; k < m; ++k) out_U[k] = U(k);
  return 0;
}

// Run one of the reference's shipped LM drivers (prints to stdout, writes CSVs into cwd).
int hadi_ref_run_shipped(int which) {
  if (which == 0)
    test_calibration_european_multi_maturity();
  else if (which == 1)
    test_calibration_european();
  else if (which == 2)
    test_calibration_american_divident_multi_maturity();
  else
    return -1;
  return 0;
}

}  // extern "C"
