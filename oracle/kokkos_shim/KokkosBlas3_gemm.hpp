// TEST INFRASTRUCTURE — stand-in for KokkosKernels' KokkosBlas::gemm (un-vendored, version
// unpinned in the reference: CMakeLists.txt:8-13).  Only the call shape used at
// src/jacobian_computation.cpp:117 is supported: C = alpha * op(A) * op(B) + beta * C.
// Accumulation order over the inner index is ascending k ("parity unpinned" at this
// boundary: the real library's order is backend dependent; the product uses the same order).
#pragma once
#include <Kokkos_Core.hpp>
namespace KokkosBlas {
template <class AV, class BV, class CV>
inline void gemm(const char* transA, const char* transB, double alpha, const AV& A, const BV& B,
                 double beta, const CV& C) {
  const bool tA = (transA[0] == 'T' || transA[0] == 't');
  const bool tB = (transB[0] == 'T' || transB[0] == 't');
  const std::size_t M = tA ? A.extent(1) : A.extent(0);
  const std::size_t K = tA ? A.extent(0) : A.extent(1);
  const std::size_t N = tB ? B.extent(0) : B.extent(1);
  for (std::size_t i = 0; i < M; ++i)
    for (std::size_t j = 0; j < N; ++j) {
      double acc = 0.0;
      for (std::size_t k = 0; k < K; ++k) {
        const double a = tA ? A(k, i) : A(i, k);
        const double b = tB ? B(j, k) : B(k, j);
        acc += a * b;
      }
      C(i, j) = (beta == 0.0) ? alpha * acc : alpha * acc + beta * C(i, j);
    }
}
}  // namespace KokkosBlas
