// hadi — the callers either side of the calibration driver: synthetic market
// generation, dividend-adjusted spot, implied-volatility inversion and the calibration report writer.
// Host-only C++ behind the C ABI of include/hadi.h; restates
//   BlackScholes::call_vega / reverse_BS_dic / reverse_BS            src/bs.hpp:124-192
//   BlackScholes::generate_market_data{,_with_dividends}             src/bs.hpp:58-112
//   the implied-vol post-processing and CSV export of the LM drivers src/heston_calibration.cpp:436-511,
//                                                                    1498-1584, 2853-2923
// The reference streams doubles with operator<< at the default precision; the writers below do the same,
// so that for equal numbers the files are equal byte for byte (the Time= field aside).
#include <chrono>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <string>
#include <vector>

#include "../../include/hadi.h"

namespace {

// src/bs.hpp:37-39 with j = 1
double d_1(double S, double K, double r, double v, double T) {
  return (std::log(S / K) + (r + std::pow(-1.0, 0) * 0.5 * v * v) * T) / (v * std::sqrt(T));
}

const int kNewtonCap = 100000;  // the reference's Newton loop is unbounded; a run this long falls back to bisection

}  // namespace

extern "C" {

// src/bs.hpp:124-127 with CP = 1
double hadi_bs_vega(double S, double K, double r, double vol, double T) {
  const double d = d_1(S, K, r, vol, T);
  return 1 * S * std::exp(-d * d / 2.0) * std::sqrt(T / (2.0 * M_PI));
}

// src/bs.hpp:131-160 (the warning print is dropped; the returned midpoint is the same)
double hadi_bs_implied_vol_bisect(double S, double K, double r, double T, double target, double epsilon,
                                  double a, double b) {
  const int MAX_ITER = 1000;
  int iter = 0;
  double x = (b + a) / 2;
  double C = hadi_bs_call(S, K, r, x, T);
  while (std::abs(C - target) > epsilon && iter < MAX_ITER) {
    C = hadi_bs_call(S, K, r, x, T);
    if (C > target)
      b = x;
    else
      a = x;
    x = (b + a) / 2;
    iter++;
  }
  return x;
}

// src/bs.hpp:164-192: Newton from v0; when vega vanishes, bisection on [0.001, 1]
double hadi_bs_implied_vol(double S, double K, double r, double T, double v0, double target, double epsilon) {
  double x = v0;
  double C = hadi_bs_call(S, K, r, x, T);
  bool fail = false;
  int guard = 0;
  while (std::abs(C - target) > epsilon) {
    C = hadi_bs_call(S, K, r, x, T);
    const double V = hadi_bs_vega(S, K, r, x, T);
    if (std::abs(V) < 1e-10 || ++guard > kNewtonCap) {
      fail = true;
      break;
    }
    x -= (C - target) / V;
  }
  if (fail) x = hadi_bs_implied_vol_bisect(S, K, r, T, target, epsilon, 0.001, 1.0);
  return x;
}

// src/bs.hpp:91-102 (same loop at src/heston_calibration.cpp:1500-1511, 2068-2079)
double hadi_dividend_adjusted_spot(double S0, double T, double r_d, int num_dividends, const double* dates,
                                   const double* amounts, const double* percentages) {
  double S_adjusted = S0;
  for (int i = 0; i < num_dividends; ++i) {
    if (dates[i] < T) {
      S_adjusted -= amounts[i] * std::exp(-r_d * dates[i]);
      S_adjusted -= (S0 * percentages[i]) * std::exp(-r_d * dates[i]);
    }
  }
  return S_adjusted;
}

// generate_market_data (src/bs.hpp:58-77) / generate_market_data_with_dividends (:79-112), one price per
// calibration point (its own strike and maturity, as the multi-maturity drivers call them per maturity:
// src/heston_calibration.cpp:2553-2561); results land at global_index.
int hadi_market_prices(double S0, double r_d, double market_vol, int n, const hadi_point* points,
                       int num_dividends, const double* dates, const double* amounts, const double* percentages,
                       double* prices) {
  if (n < 0 || (n > 0 && (!points || !prices))) return HADI_ERR_ARG;
  if (num_dividends < 0 || (num_dividends > 0 && (!dates || !amounts || !percentages))) return HADI_ERR_ARG;
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return HADI_ERR_ARG;
    const double T = points[k].maturity;
    const double S = num_dividends > 0 ? hadi_dividend_adjusted_spot(S0, T, r_d, num_dividends, dates, amounts, percentages)
                                       : S0;
    prices[gi] = hadi_bs_call(S, points[k].strike, r_d, market_vol, T);
  }
  return HADI_OK;
}

// src/heston_calibration.cpp:441-460 / 2890-2910: implied vols of market and fitted prices (Newton from 0.5)
// and their absolute difference; `spot` is S_0 (plain and multi-maturity drivers) or the dividend-adjusted
// spot (single-maturity dividend drivers, :1525).  Arrays are indexed by global_index; outputs may be NULL.
int hadi_implied_vols(double spot, double r_d, int n, const hadi_point* points, const double* market,
                      const double* fitted, double epsilon, double* market_iv, double* fitted_iv, double* iv_diff) {
  if (n < 0 || (n > 0 && (!points || !market || !fitted))) return HADI_ERR_ARG;
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return HADI_ERR_ARG;
    const double K = points[k].strike, T = points[k].maturity;
    const double miv = hadi_bs_implied_vol(spot, K, r_d, T, 0.5, market[gi], epsilon);
    const double fiv = hadi_bs_implied_vol(spot, K, r_d, T, 0.5, fitted[gi], epsilon);
    if (market_iv) market_iv[gi] = miv;
    if (fitted_iv) fitted_iv[gi] = fiv;
    if (iv_diff) iv_diff[gi] = std::abs(miv - fiv);
  }
  return HADI_OK;
}

// The calibration report the reference's LM drivers export for plotting.
//   format 0 (single maturity, src/heston_calibration.cpp:467-508): "# <n> options, Time=..." and rows
//            Strike,MarketPrice,FittedPrice,IVDifference
//   format 1 (multi maturity, :2857-2921): "# Calibration with <M> maturities, <S> strikes per maturity, ..."
//            and rows Maturity,Strike,MarketPrice,FittedPrice,MarketIV,FittedIV,IVDifference in
//            maturity-major order (index m*S + s)
// market / fitted are indexed by global_index; `spot` is the spot the implied vols are inverted at.
int hadi_write_calibration_csv(const char* path, int format, double spot, double r_d, int num_maturities,
                               int num_strikes, const hadi_point* points, const double* market,
                               const double* fitted, const hadi_model* initial, const hadi_lm_result* result,
                               double total_time_s, double iv_epsilon) {
  if (!path || !points || !market || !fitted || !initial || !result) return HADI_ERR_ARG;
  if (format != 0 && format != 1) return HADI_ERR_ARG;
  if (num_maturities < 1 || num_strikes < 0 || (format == 0 && num_maturities != 1)) return HADI_ERR_ARG;
  const int n = num_maturities * num_strikes;
  std::vector<double> miv((size_t)n + 1), fiv((size_t)n + 1), div((size_t)n + 1);
  // position of every global index in `points`
  std::vector<int> where((size_t)n + 1, -1);
  for (int k = 0; k < n; ++k) {
    const int gi = points[k].global_index;
    if (gi < 0 || gi >= n) return HADI_ERR_ARG;
    where[gi] = k;
  }
  for (int gi = 0; gi < n; ++gi)
    if (where[gi] < 0) return HADI_ERR_ARG;
  int rc = hadi_implied_vols(spot, r_d, n, points, market, fitted, iv_epsilon, miv.data(), fiv.data(), div.data());
  if (rc != HADI_OK) return rc;
  std::ofstream out(path);
  if (!out.is_open()) return HADI_ERR_ARG;
  if (format == 0) {
    out << "# " << num_strikes << " options, Time=" << total_time_s << " s, FinalError=" << result->final_error
        << ", iterationCount=" << result->iterations << ", TotalPdeSolves=" << result->pde_solves
        << ", init_kappa=" << initial->kappa << ", init_eta=" << initial->eta << ", init_sigma=" << initial->sigma
        << ", init_rho=" << initial->rho << ", init_v0=" << initial->V0 << ", kappa=" << result->params[0]
        << ", eta=" << result->params[1] << ", sigma=" << result->params[2] << ", rho=" << result->params[3]
        << ", v0=" << result->params[4] << "\n";
    out << "Strike,MarketPrice,FittedPrice,IVDifference\n";
    for (int i = 0; i < n; ++i)
      out << points[where[i]].strike << "," << market[i] << "," << fitted[i] << "," << div[i] << "\n";
  } else {
    out << "# Calibration with " << num_maturities << " maturities, " << num_strikes << " strikes per maturity, "
        << "Time=" << total_time_s << " s, "
        << "FinalError=" << result->final_error << ", "
        << "IterationCount=" << result->iterations << ", "
        << "TotalPdeSolves=" << result->pde_solves << ", "
        << "init_kappa=" << initial->kappa << ", "
        << "init_eta=" << initial->eta << ", "
        << "init_sigma=" << initial->sigma << ", "
        << "init_rho=" << initial->rho << ", "
        << "init_v0=" << initial->V0 << ", "
        << "kappa=" << result->params[0] << ", "
        << "eta=" << result->params[1] << ", "
        << "sigma=" << result->params[2] << ", "
        << "rho=" << result->params[3] << ", "
        << "v0=" << result->params[4] << "\n";
    out << "Maturity,Strike,MarketPrice,FittedPrice,MarketIV,FittedIV,IVDifference\n";
    for (int i = 0; i < n; ++i) {
      const hadi_point& pt = points[where[i]];
      out << pt.maturity << "," << pt.strike << "," << market[i] << "," << fitted[i] << "," << miv[i] << ","
          << fiv[i] << "," << div[i] << "\n";
    }
  }
  out.close();
  return out.fail() ? HADI_ERR_ARG : HADI_OK;
}


// ---- convergence-study harness ---------------------------------------------------
// ConvergenceExporter::testWithRelatedGridSizes (src/solver.cpp:61-150): one call priced on the grids m1 = 2 m2 for
// every m2 of the list, N time steps, relative error against a reference price, mean wall time of `repeats` solves
// (the reference: N = 20, theta = 0.8, 20 repeats) — here with any stepping scheme of the library and every solve on
// the GPU.  prices / rel_errors / seconds have n_sizes entries; any of them may be NULL.
int hadi_convergence_study(hadi_ctx* ctx, const hadi_model* model, double K, double T, int N, double theta, int scheme,
                           int n_sizes, const int* m2_sizes, double ref_price, int repeats, double* prices,
                           double* rel_errors, double* seconds) {
  if (!ctx || !model || !m2_sizes || n_sizes < 0 || N < 1 || !(T > 0) || repeats < 1) return HADI_ERR_ARG;
  for (int k = 0; k < n_sizes; ++k) {
    const int m2 = m2_sizes[k], m1 = 2 * m2;
    hadi_numerics num{};
    num.m1 = m1; num.m2 = m2; num.theta = theta;
    num.style = HADI_EUROPEAN; num.payoff = HADI_CALL; num.scheme = scheme;
    const hadi_point pt{K, T, N, T / N, 0};
    double price = 0.0, total = 0.0;
    for (int r = 0; r < repeats; ++r) {
      const auto t0 = std::chrono::steady_clock::now();
      const int rc = hadi_price_batch(ctx, model, &num, 1, &pt, &price, nullptr, nullptr);
      if (rc != HADI_OK) return rc;
      total += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    if (prices) prices[k] = price;
    if (rel_errors) rel_errors[k] = std::abs(price - ref_price) / ref_price;
    if (seconds) seconds[k] = total / repeats;
  }
  return HADI_OK;
}

// ConvergenceExporter::exportToCSV (src/solver.cpp:281-295): <name>_convergence.csv, header m1,m2,price,error,time,
// numbers in scientific notation with ten digits (the stream state the reference leaves set after the first price).
int hadi_write_convergence_csv(const char* path, int n_sizes, const int* m2_sizes, const double* prices,
                               const double* rel_errors, const double* seconds) {
  if (!path || n_sizes < 0 || (n_sizes > 0 && (!m2_sizes || !prices || !rel_errors || !seconds))) return HADI_ERR_ARG;
  std::ofstream file(path);
  if (!file) return HADI_ERR_ARG;
  file << "m1,m2,price,error,time\n";
  for (int i = 0; i < n_sizes; ++i) {
    file << 2 * m2_sizes[i] << "," << m2_sizes[i] << "," << std::scientific << std::setprecision(10) << prices[i] << ","
         << rel_errors[i] << "," << seconds[i] << "\n";
  }
  return file.good() ? HADI_OK : HADI_ERR_ARG;
}

}  // extern "C"
