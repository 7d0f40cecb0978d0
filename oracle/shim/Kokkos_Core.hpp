// TEST INFRASTRUCTURE ONLY (oracle/). Not part of the product path.
//
// Host-only stand-in for the subset of Kokkos that the reference's cedr/*.cpp
// touch, so that the UNMODIFIED reference sources under /root/reference/cedr
// compile with plain g++ in a container that has no Kokkos. It is deliberately
// not "a Kokkos": there is one memory space (host), Views are flat 1-D arrays,
// team-level loops run sequentially (which is what real Kokkos does on a host
// backend with the reference's default TeamPolicy(outer, 1, 1),
// cedr_kokkos.hpp:118), and league-level / range loops are OpenMP-parallel
// when compiled with -fopenmp. Results produced through this shim are labelled
// "reference sources + stand-in runtime" wherever they are reported.
//
// Surface covered (enumerated by grepping Kokkos:: in /root/reference/cedr):
//   View<T*,...> (label ctor, (ptr,n) ctor, converting ctor, (), [], data,
//   size, extent, extent_int, HostMirror, const_type, traits::*),
//   Device, Serial/OpenMP(+concurrency), Default(Host)ExecutionSpace,
//   LayoutRight, MemoryTraits + Unmanaged/RandomAccess/Atomic/Restrict/Aligned,
//   create_mirror_view, deep_copy, fence, initialize, finalize, abort,
//   RangePolicy, TeamPolicy(+member_type::league_rank), TeamThreadRange,
//   parallel_for x3, parallel_reduce(TeamThreadRange, f, Sum<T>),
//   reduction_identity, KOKKOS_* macros.
#ifndef CEDR_B200_ORACLE_KOKKOS_SHIM_HPP
#define CEDR_B200_ORACLE_KOKKOS_SHIM_HPP

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <type_traits>

#ifdef _OPENMP
# include <omp.h>
# define KOKKOS_ENABLE_OPENMP
#else
# define KOKKOS_ENABLE_SERIAL
#endif

// Selects the ">= v3" trait names in cedr_kokkos.hpp:31-45.
#define KOKKOS_VERSION 30100
// Selects ConstExceptGnu = <empty> in cedr_kokkos.hpp:16-21.
#define KOKKOS_COMPILER_GNU 1
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]

namespace Kokkos {

struct HostSpace { typedef HostSpace memory_space; };

struct Serial {
  typedef Serial execution_space;
  typedef HostSpace memory_space;
  static int concurrency () { return 1; }
};

struct OpenMP {
  typedef OpenMP execution_space;
  typedef HostSpace memory_space;
  static int concurrency () {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
  }
};

#ifdef _OPENMP
typedef OpenMP DefaultExecutionSpace;
typedef OpenMP DefaultHostExecutionSpace;
#else
typedef Serial DefaultExecutionSpace;
typedef Serial DefaultHostExecutionSpace;
#endif

template <typename ES, typename MS> struct Device {
  typedef ES execution_space;
  typedef MS memory_space;
};

struct LayoutRight {};

enum MemoryTraitsFlags : unsigned {
  Unmanaged = 1, RandomAccess = 2, Atomic = 4, Restrict = 8, Aligned = 16
};

template <unsigned M> struct MemoryTraits {
  enum : bool {
    is_unmanaged = (M & Unmanaged) != 0,
    is_random_access = (M & RandomAccess) != 0,
    is_atomic = (M & Atomic) != 0,
    is_restrict = (M & Restrict) != 0,
    is_aligned = (M & Aligned) != 0
  };
};

namespace shim_detail {
// Pick the MemoryTraits<> out of a View's property pack, default 0.
template <typename... P> struct PickTraits { typedef MemoryTraits<0> type; };
template <unsigned M, typename... P>
struct PickTraits<MemoryTraits<M>, P...> { typedef MemoryTraits<M> type; };
template <typename T, typename... P>
struct PickTraits<T, P...> { typedef typename PickTraits<P...>::type type; };
} // namespace shim_detail

// Flat host array with shared ownership (label ctor) or no ownership (ptr ctor).
template <typename DataType, typename... Props>
class View {
public:
  typedef typename std::remove_pointer<DataType>::type value_type;
  typedef typename std::remove_const<value_type>::type non_const_value_type;

  struct traits {
    typedef DataType scalar_array_type;
    typedef LayoutRight array_layout;
    typedef Device<DefaultExecutionSpace, HostSpace> device_type;
    typedef typename shim_detail::PickTraits<Props...>::type memory_traits;
  };

  // Must be a view of the NON-const value type even for const views
  // (cedr_qlt.cpp:143 assigns a_h = a_h_).
  typedef View<non_const_value_type*> HostMirror;
  typedef View<const non_const_value_type*, Props...> const_type;

  View () : ptr_(nullptr), n_(0) {}
  View (const std::string& /*label*/, size_t n)
    : owner_(new non_const_value_type[n](),
             std::default_delete<non_const_value_type[]>()),
      ptr_(owner_.get()), n_(n) {}
  View (value_type* p, size_t n) : ptr_(p), n_(n) {}
  // Conversions across const / memory-trait variants (cedr_qlt.cpp:147-151).
  template <typename DT2, typename... P2>
  View (const View<DT2, P2...>& v) : owner_(v.owner_), ptr_(v.ptr_), n_(v.n_) {}

  value_type& operator() (size_t i) const { return ptr_[i]; }
  value_type& operator[] (size_t i) const { return ptr_[i]; }
  value_type* data () const { return ptr_; }
  size_t size () const { return n_; }
  size_t extent (int) const { return n_; }
  int extent_int (int) const { return static_cast<int>(n_); }

  // Public so that the converting ctor and create_mirror_view can reach them.
  std::shared_ptr<non_const_value_type> owner_;
  value_type* ptr_;
  size_t n_;
};

// One memory space => the mirror aliases the same memory; deep_copy of a view
// onto its own mirror is then a no-op.
template <typename V>
typename V::HostMirror create_mirror_view (const V& v) {
  typename V::HostMirror m;
  m.owner_ = v.owner_;
  m.ptr_ = const_cast<typename V::non_const_value_type*>(v.ptr_);
  m.n_ = v.n_;
  return m;
}

template <typename Dst, typename Src>
void deep_copy (const Dst& dst, const Src& src) {
  const void* s = static_cast<const void*>(src.data());
  void* d = const_cast<void*>(static_cast<const void*>(dst.data()));
  if (s == d) return;
  std::memcpy(d, s, sizeof(typename Dst::value_type)*dst.size());
}

inline void fence () {}
inline void initialize (int&, char**) {}
inline void finalize () {}
inline void abort (const char* msg) {
  std::fprintf(stderr, "Kokkos::abort (shim): %s\n", msg);
  std::abort();
}

template <typename ES = DefaultExecutionSpace>
struct RangePolicy {
  int begin, end;
  RangePolicy (int b, int e) : begin(b), end(e) {}
};

template <typename ES = DefaultExecutionSpace>
struct TeamPolicy {
  struct member_type {
    int league_rank_;
    int league_rank () const { return league_rank_; }
  };
  int league_size, team_size;
  TeamPolicy (int league, int team, int /*vector*/ = 1)
    : league_size(league), team_size(team) {}
};

struct TeamThreadRangeShim { int n; };
template <typename Member>
TeamThreadRangeShim TeamThreadRange (const Member&, int n) {
  return TeamThreadRangeShim{n};
}

// cedr_caas.cpp:8-32 specialises this for its own ComposeReal2.
template <typename T> struct reduction_identity;
template <> struct reduction_identity<double> {
  static double sum () { return 0; }
};

template <typename T> struct Sum {
  T& result;
  Sum (T& r) : result(r) {}
};

template <typename ES, typename F>
void parallel_for (const RangePolicy<ES>& p, const F& f) {
#ifdef _OPENMP
# pragma omp parallel for
#endif
  for (int i = p.begin; i < p.end; ++i) f(i);
}

template <typename ES, typename F>
void parallel_for (const TeamPolicy<ES>& p, const F& f) {
#ifdef _OPENMP
# pragma omp parallel for
#endif
  for (int i = 0; i < p.league_size; ++i) {
    typename TeamPolicy<ES>::member_type m{i};
    f(m);
  }
}

// Team-level loops: sequential, in index order (== real Kokkos host backend
// with team size 1). This fixes the summation order of the reference CAAS.
template <typename F>
void parallel_for (const TeamThreadRangeShim& r, const F& f) {
  for (int i = 0; i < r.n; ++i) f(i);
}

template <typename F, typename T>
void parallel_reduce (const TeamThreadRangeShim& r, const F& f, Sum<T> s) {
  T acc = reduction_identity<T>::sum();
  for (int i = 0; i < r.n; ++i) f(i, acc);
  s.result = acc;
}

} // namespace Kokkos

#endif
