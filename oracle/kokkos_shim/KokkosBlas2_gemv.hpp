// TEST INFRASTRUCTURE — stand-in for KokkosKernels' KokkosBlas::gemv, call shape of
// src/jacobian_computation.cpp:154: y = alpha * op(A) * x + beta * y, ascending-k accumulation.
#pragma once
#include <Kokkos_Core.hpp>
namespace KokkosBlas {
template <class AV, class XV, class YV>
inline void gemv(const char* trans, double alpha, const AV& A, const XV& x, double beta, const YV& y) {
  const bool tA = (trans[0] == 'T' || trans[0] == 't');
  const std::size_t M = tA ? A.extent(1) : A.extent(0);
  const std::size_t K = tA ? A.extent(0) : A.extent(1);
  for (std::size_t i = 0; i < M; ++i) {
    double acc = 0.0;
    for (std::size_t k = 0; k < K; ++k) acc += (tA ? A(k, i) : A(i, k)) * x(k);
    y(i) = (beta == 0.0) ? alpha * acc : alpha * acc + beta * y(i);
  }
}
}  // namespace KokkosBlas
