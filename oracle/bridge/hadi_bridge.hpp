// hadi_bridge.hpp — what a maintainer of the reference would add to route its solver entry points to libhadi.so
// (INTEGRATION.md section 3).  Header-only; needs include/hadi.h, <Kokkos_Core.hpp> and the reference's
// CalibrationPoint (src/heston_calibration.cpp:2165-2171 — the struct is defined in that .cpp, so a caller outside it
// declares the same five fields; the layout is asserted against hadi_point).
#pragma once
#include <Kokkos_Core.hpp>

#include <stdexcept>
#include <string>
#include <vector>

#include "hadi.h"

struct HadiBridge {
  hadi_ctx* ctx = nullptr;
  explicit HadiBridge(int device = 0) {
    if (hadi_create(&ctx, device) != HADI_OK) throw std::runtime_error("hadi_create: no usable CUDA device");
  }
  ~HadiBridge() { hadi_destroy(ctx); }
  HadiBridge(const HadiBridge&) = delete;
  HadiBridge& operator=(const HadiBridge&) = delete;
  void check(int rc) const {
    if (rc != HADI_OK) throw std::runtime_error(std::string("libhadi: ") + hadi_last_error(ctx));
  }
};

// compute_base_prices_multi_maturity (src/heston_calibration.cpp:2339): European, no dividends
template <class PointView, class PriceView>
inline void hadi_bridge_base_prices_multi_maturity(HadiBridge& hb, double S_0, double V_0, double r_d, double r_f,
                                                   double rho, double sigma, double kappa, double eta, int m1, int m2,
                                                   double theta, const PointView& points, int n, PriceView& base_prices) {
  static_assert(sizeof(*points.data()) == sizeof(hadi_point), "CalibrationPoint and hadi_point share one layout");
  const hadi_model mdl{S_0, V_0, r_d, r_f, kappa, eta, sigma, rho};
  const hadi_numerics num{m1, m2, theta, HADI_EUROPEAN, HADI_CALL, HADI_DOUGLAS, 0, nullptr, nullptr, nullptr};
  hb.check(hadi_price_batch(hb.ctx, &mdl, &num, n, reinterpret_cast<const hadi_point*>(points.data()),
                            base_prices.data(), nullptr, nullptr));
}

// compute_jacobian_multi_maturity (src/heston_calibration.cpp:2174): J [n][5] row-major, columns (kappa, eta, sigma, rho, v0)
template <class PointView, class JView, class PriceView>
inline void hadi_bridge_jacobian_multi_maturity(HadiBridge& hb, double S_0, double V_0, double r_d, double r_f,
                                                double rho, double sigma, double kappa, double eta, int m1, int m2,
                                                double theta, const PointView& points, int n, JView& J,
                                                PriceView& base_prices, double eps) {
  const hadi_model mdl{S_0, V_0, r_d, r_f, kappa, eta, sigma, rho};
  const hadi_numerics num{m1, m2, theta, HADI_EUROPEAN, HADI_CALL, HADI_DOUGLAS, 0, nullptr, nullptr, nullptr};
  std::vector<double> Jrow((size_t)5 * n);
  hb.check(hadi_jacobian_batch(hb.ctx, &mdl, &num, n, reinterpret_cast<const hadi_point*>(points.data()), eps,
                               Jrow.data(), base_prices.data()));
  for (int i = 0; i < n; ++i)
    for (int c = 0; c < 5; ++c) J(i, c) = Jrow[(size_t)5 * i + c];
}
