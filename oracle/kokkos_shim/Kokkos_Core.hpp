// TEST INFRASTRUCTURE — not part of the product.
//
// Serial (optionally OpenMP-over-league) stand-in for the slice of the Kokkos
// API that the reference sources under /root/reference/src use.  It exists so
// that the reference's own, unmodified .cpp/.hpp files can be compiled with
// plain g++ into oracle/_ref/libhadi_ref.so and act as the parity oracle
// (SURVEY.md §8(c)).  Semantics: one team = one "thread" (team_size 1,
// team_rank 0), i.e. exactly what a Kokkos Serial build executes; league
// members (options) may be spread over OpenMP threads, each touching only its
// own instance, which leaves every result bit-identical.
//
// This header contains no arithmetic of its own besides forwarding
// Kokkos::{abs,max,min,exp,sinh,asinh} to <cmath>.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <numeric>
#include <sstream>
#include <cstddef>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#define KOKKOS_FUNCTION
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_LAMBDA [=]
#define KOKKOS_CLASS_LAMBDA [=, *this]

namespace Kokkos {

// ---------------------------------------------------------------- spaces
struct HostSpace {};
struct Serial {
  using memory_space = HostSpace;
  using execution_space = Serial;
  static const char* name() { return "SerialShim"; }
  static int impl_thread_pool_size() { return 1; }
};
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;
struct OpenMP {
  using memory_space = HostSpace;
  static const char* name() { return "OpenMPShim"; }
  static int impl_thread_pool_size() { return 1; }
};
struct Cuda {
  using memory_space = HostSpace;
  static const char* name() { return "CudaShim"; }
};
struct LayoutRight {};
struct LayoutLeft {};

inline void initialize() {}
inline void initialize(int&, char**) {}
inline void finalize() {}
inline void fence() {}
inline void fence(const std::string&) {}

struct ALL_t {};
static constexpr ALL_t ALL{};
struct AUTO_t {};
static constexpr AUTO_t AUTO{};

// ---------------------------------------------------------------- math
using std::abs;
using std::asinh;
using std::exp;
using std::log;
using std::pow;
using std::sinh;
using std::sqrt;
template <class T>
inline T max(const T& a, const T& b) { return std::max(a, b); }
template <class T>
inline T min(const T& a, const T& b) { return std::min(a, b); }
inline double max(double a, double b) { return std::max(a, b); }
inline double min(double a, double b) { return std::min(a, b); }

// ---------------------------------------------------------------- View
namespace shim {
template <class DT>
struct view_traits;
template <class T>
struct view_traits<T*> {
  using value_type = T;
  static constexpr int rank = 1;
};
template <class T>
struct view_traits<T**> {
  using value_type = T;
  static constexpr int rank = 2;
};
}  // namespace shim

template <class DataType, class... Props>
class View {
 public:
  using traits = shim::view_traits<DataType>;
  using value_type = typename traits::value_type;
  using HostMirror = View<DataType>;
  using host_mirror_type = View<DataType>;
  static constexpr int Rank = traits::rank;

  View() = default;

  // managed, zero/default-initialised like Kokkos::View
  explicit View(const std::string&, std::size_t n0, std::size_t n1 = 1) { allocate(n0, n1); }
  explicit View(const char*, std::size_t n0, std::size_t n1 = 1) { allocate(n0, n1); }

  // unmanaged
  View(value_type* p, std::size_t n0, std::size_t n1 = 1) : ptr_(p), n0_(n0), n1_(n1) {}

  // converting copy (different memory-/execution-space decorations alias the same data)
  template <class... P2>
  View(const View<DataType, P2...>& o) : own_(o.own_), ptr_(o.ptr_), n0_(o.n0_), n1_(o.n1_) {}
  template <class... P2>
  View& operator=(const View<DataType, P2...>& o) {
    own_ = o.own_;
    ptr_ = o.ptr_;
    n0_ = o.n0_;
    n1_ = o.n1_;
    return *this;
  }

  // aliasing ctor used by subview()
  View(std::shared_ptr<value_type[]> own, value_type* p, std::size_t n0, std::size_t n1)
      : own_(std::move(own)), ptr_(p), n0_(n0), n1_(n1) {}

  template <class I>
  value_type& operator()(I i) const { return ptr_[static_cast<std::size_t>(i)]; }
  template <class I, class J>
  value_type& operator()(I i, J j) const {
    return ptr_[static_cast<std::size_t>(i) * n1_ + static_cast<std::size_t>(j)];
  }
  template <class I>
  value_type& operator[](I i) const { return ptr_[static_cast<std::size_t>(i)]; }

  std::size_t extent(int d) const { return d == 0 ? n0_ : (d == 1 ? n1_ : 1); }
  int extent_int(int d) const { return static_cast<int>(extent(d)); }
  std::size_t size() const { return n0_ * n1_; }
  std::size_t span() const { return n0_ * n1_; }
  value_type* data() const { return ptr_; }
  bool is_allocated() const { return ptr_ != nullptr; }
  std::string label() const { return ""; }

  std::shared_ptr<value_type[]> own_;
  value_type* ptr_ = nullptr;
  std::size_t n0_ = 0, n1_ = 1;

 private:
  void allocate(std::size_t n0, std::size_t n1) {
    n0_ = n0;
    n1_ = (Rank == 2) ? n1 : 1;
    const std::size_t n = n0_ * n1_;
    own_ = std::shared_ptr<value_type[]>(new value_type[n > 0 ? n : 1]());
    ptr_ = own_.get();
  }
};

template <class T, class... P>
inline View<T*> subview(const View<T**, P...>& v, std::size_t i, ALL_t) {
  return View<T*>(v.own_, v.ptr_ + i * v.n1_, v.n1_, 1);
}
template <class T, class... P>
inline View<T*> subview(const View<T**, P...>& v, int i, ALL_t) {
  return View<T*>(v.own_, v.ptr_ + static_cast<std::size_t>(i) * v.n1_, v.n1_, 1);
}

template <class DT, class... P>
inline View<DT> create_mirror_view(const View<DT, P...>& v) { return View<DT>(v); }
template <class Space, class DT, class... P>
inline View<DT> create_mirror_view(const Space&, const View<DT, P...>& v) { return View<DT>(v); }
template <class DT, class... P>
inline View<DT> create_mirror(const View<DT, P...>& v) {
  View<DT> r("mirror", v.extent(0), v.extent(1));
  return r;
}
template <class Space, class DT, class... P>
inline View<DT> create_mirror_view_and_copy(const Space&, const View<DT, P...>& v) { return View<DT>(v); }

template <class DT, class... P1, class... P2>
inline void deep_copy(const View<DT, P1...>& dst, const View<DT, P2...>& src) {
  if (dst.data() == src.data()) return;
  const std::size_t n = std::min(dst.size(), src.size());
  for (std::size_t k = 0; k < n; ++k) dst.data()[k] = src.data()[k];
}
template <class DT, class... P1>
inline void deep_copy(const View<DT, P1...>& dst, const typename View<DT, P1...>::value_type& val) {
  for (std::size_t k = 0; k < dst.size(); ++k) dst.data()[k] = val;
}
template <class Exec, class DT, class... P1, class... P2>
inline void deep_copy(const Exec&, const View<DT, P1...>& dst, const View<DT, P2...>& src) {
  deep_copy(dst, src);
}

// ---------------------------------------------------------------- policies
struct TeamMember {
  int league_rank_ = 0;
  int league_size_ = 1;
  int league_rank() const { return league_rank_; }
  int league_size() const { return league_size_; }
  int team_rank() const { return 0; }
  int team_size() const { return 1; }
  void team_barrier() const {}
  template <class T>
  void team_broadcast(T&, int) const {}
};

template <class... Props>
struct TeamPolicy {
  using member_type = TeamMember;
  int league_size_ = 0;
  TeamPolicy() = default;
  TeamPolicy(int league, AUTO_t) : league_size_(league) {}
  TeamPolicy(int league, int) : league_size_(league) {}
  TeamPolicy(int league, AUTO_t, int) : league_size_(league) {}
  TeamPolicy(int league, int, int) : league_size_(league) {}
  template <class... P2>
  TeamPolicy(const TeamPolicy<P2...>& o) : league_size_(o.league_size_) {}
  int league_size() const { return league_size_; }
  int team_size() const { return 1; }
};

template <class... Props>
struct RangePolicy {
  long begin_ = 0, end_ = 0;
  RangePolicy(long b, long e) : begin_(b), end_(e) {}
};

template <int N>
struct Rank {};
template <class... Props>
struct MDRangePolicy {
  long lo_[2], hi_[2];
  MDRangePolicy(std::initializer_list<long> lo, std::initializer_list<long> hi) {
    int k = 0;
    for (long v : lo) lo_[k++] = v;
    k = 0;
    for (long v : hi) hi_[k++] = v;
  }
};

struct TeamThreadRangeT {
  long begin_, end_;
};
inline TeamThreadRangeT TeamThreadRange(const TeamMember&, long n) { return {0, n}; }
inline TeamThreadRangeT TeamThreadRange(const TeamMember&, long b, long e) { return {b, e}; }
inline TeamThreadRangeT ThreadVectorRange(const TeamMember&, long n) { return {0, n}; }
inline TeamThreadRangeT TeamVectorRange(const TeamMember&, long n) { return {0, n}; }

// ---------------------------------------------------------------- parallel_for
template <class F>
inline void parallel_for(const TeamThreadRangeT& r, const F& f) {
  for (long i = r.begin_; i < r.end_; ++i) f(static_cast<int>(i));
}

template <class... P, class F>
inline void parallel_for(const std::string&, const TeamPolicy<P...>& p, const F& f) {
  const int n = p.league_size();
#if defined(_OPENMP) && defined(HADI_SHIM_OMP_LEAGUE)
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int l = 0; l < n; ++l) {
    TeamMember m;
    m.league_rank_ = l;
    m.league_size_ = n;
    f(m);
  }
}
template <class... P, class F>
inline void parallel_for(const TeamPolicy<P...>& p, const F& f) {
  parallel_for(std::string(), p, f);
}

template <class... P, class F>
inline void parallel_for(const std::string&, const RangePolicy<P...>& p, const F& f) {
  for (long i = p.begin_; i < p.end_; ++i) f(static_cast<int>(i));
}
template <class... P, class F>
inline void parallel_for(const RangePolicy<P...>& p, const F& f) {
  for (long i = p.begin_; i < p.end_; ++i) f(static_cast<int>(i));
}

template <class... P, class F>
inline void parallel_for(const std::string&, const MDRangePolicy<P...>& p, const F& f) {
  for (long i = p.lo_[0]; i < p.hi_[0]; ++i)
    for (long j = p.lo_[1]; j < p.hi_[1]; ++j) f(static_cast<int>(i), static_cast<int>(j));
}
template <class... P, class F>
inline void parallel_for(const MDRangePolicy<P...>& p, const F& f) {
  parallel_for(std::string(), p, f);
}

template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
inline void parallel_for(const std::string&, I n, const F& f) {
  for (long i = 0; i < static_cast<long>(n); ++i) f(static_cast<int>(i));
}
template <class I, class F, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
inline void parallel_for(I n, const F& f) {
  for (long i = 0; i < static_cast<long>(n); ++i) f(static_cast<int>(i));
}

template <class... P, class F, class R>
inline void parallel_reduce(const std::string&, const RangePolicy<P...>& p, const F& f, R& result) {
  R acc = R();
  for (long i = p.begin_; i < p.end_; ++i) f(static_cast<int>(i), acc);
  result = acc;
}
template <class I, class F, class R, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
inline void parallel_reduce(const std::string&, I n, const F& f, R& result) {
  R acc = R();
  for (long i = 0; i < static_cast<long>(n); ++i) f(static_cast<int>(i), acc);
  result = acc;
}

struct Timer {
  double seconds() const { return 0.0; }
  void reset() {}
};

}  // namespace Kokkos
