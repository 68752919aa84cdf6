// Kokkos_Core.hpp — a SERIAL-ONLY stand-in for the subset of Kokkos that the reference's preqx C++ path uses
// (src/share/cxx of /root/reference). TEST INFRASTRUCTURE: it exists so that the reference's own functor sources —
// CaarFunctorImpl.hpp, EulerStepFunctorImpl.hpp, HyperviscosityFunctorImpl.{hpp,cpp}, RemapFunctor.hpp, PpmRemap.hpp,
// SphereOperators.hpp, BoundaryExchange.cpp, prim_driver.cpp ... — can be compiled WHERE THEY LIE into
// oracle/_ref/libref_hommexx_*.so (oracle/Makefile, target ref_full) and run on one host thread as the pin of
// oracle/oracle.c. Kokkos itself is not in this image. Nothing here is a port of Kokkos: every parallel pattern is the
// plain sequential loop its Serial back end is defined to be equivalent to, a team has one thread and one vector
// lane, and a View is a reference-counted LayoutRight array. Only what the reference touches is provided.
#ifndef REF_SHIM_KOKKOS_CORE_HPP
#define REF_SHIM_KOKKOS_CORE_HPP

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <iostream>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <initializer_list>
#include <memory>
#include <string>
#include <type_traits>
#include <utility>

#define KOKKOS_ENABLE_SERIAL
// REF_SHIM_OPENMP (the timing build only, bench.py's CPU arm): league iterations and flat ranges are spread over the
// host threads with OpenMP, as Kokkos' own OpenMP back end does with one thread per team; everything inside a team
// stays sequential. The parity builds never define it.
#ifdef REF_SHIM_OPENMP
#include <omp.h>
#define KOKKOS_ENABLE_OPENMP
#define REF_SHIM_PRAGMA(x) _Pragma(#x)
#define REF_SHIM_PARALLEL_FOR REF_SHIM_PRAGMA(omp parallel for schedule(static))
#else
#define REF_SHIM_PARALLEL_FOR
#endif
#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline __attribute__((always_inline))
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define KOKKOS_RESTRICT __restrict__

namespace Kokkos {

struct LayoutRight {};
struct LayoutLeft {};
struct Serial;

struct HostSpace {
  using memory_space = HostSpace;
  using execution_space = Serial;
  using device_type = HostSpace;
};
struct ScratchMemorySpace {
  using memory_space = ScratchMemorySpace;
  using execution_space = Serial;
  using device_type = ScratchMemorySpace;
  void* get_shmem(size_t) const { return nullptr; }
};
struct Serial {
  using execution_space = Serial;
  using memory_space = HostSpace;
  using scratch_memory_space = ScratchMemorySpace;
  using array_layout = LayoutRight;
  using device_type = Serial;
  static int concurrency() { return 1; }
  static int impl_thread_pool_size() { return 1; }
  static int thread_pool_size() { return 1; }
  static void fence() {}
  static const char* name() { return "Serial (reference shim)"; }
  static void print_configuration(std::ostream&, bool = false) {}
};
#ifdef REF_SHIM_OPENMP
struct OpenMP {
  using execution_space = OpenMP;
  using memory_space = HostSpace;
  using scratch_memory_space = ScratchMemorySpace;
  using array_layout = LayoutRight;
  using device_type = OpenMP;
  static int concurrency() { return omp_get_max_threads(); }
  static int impl_thread_pool_size() { return omp_get_max_threads(); }
  static int thread_pool_size() { return omp_get_max_threads(); }
  static void fence() {}
  static const char* name() { return "OpenMP (reference shim)"; }
  static void print_configuration(std::ostream&, bool = false) {}
};
using DefaultExecutionSpace = OpenMP;
using DefaultHostExecutionSpace = OpenMP;
#else
using DefaultExecutionSpace = Serial;
using DefaultHostExecutionSpace = Serial;
#endif

enum MemoryTraitsFlags { Unmanaged = 0x01, RandomAccess = 0x02, Atomic = 0x04, Restrict = 0x08, Aligned = 0x10 };
template <unsigned T>
struct MemoryTraits {
  enum : unsigned { value = T };
  enum : bool {
    Unmanaged = (T & Kokkos::Unmanaged) != 0,
    RandomAccess = (T & Kokkos::RandomAccess) != 0,
    Atomic = (T & Kokkos::Atomic) != 0,
    Restrict = (T & Kokkos::Restrict) != 0
  };
};

template <class T>
struct IndexType {};

inline void initialize() {}
inline void initialize(int&, char**) {}
inline void finalize() {}
inline bool is_initialized() { return true; }
inline void fence() {}
[[noreturn]] inline void abort(const char* msg) {
  std::fprintf(stderr, "Kokkos::abort: %s\n", msg);
  std::abort();
}

template <class T, size_t N>
struct Array {
  T m_data[N > 0 ? N : 1];
  using value_type = T;
  KOKKOS_INLINE_FUNCTION T& operator[](size_t i) { return m_data[i]; }
  KOKKOS_INLINE_FUNCTION const T& operator[](size_t i) const { return m_data[i]; }
  KOKKOS_INLINE_FUNCTION T* data() { return m_data; }
  KOKKOS_INLINE_FUNCTION const T* data() const { return m_data; }
  static constexpr size_t size() { return N; }
};

template <class T>
T atomic_fetch_add(T* p, T v) {
  const T old = *p;
  *p += v;
  return old;
}

template <class T>
struct reduction_identity;
template <class T>
struct Sum {};

// ---- View -------------------------------------------------------------------------------------------------
namespace Impl {
enum : int { MEMORY_ALIGNMENT = 64 };

template <class T>
struct PointerDepth { enum : int { value = 0 }; using type = T; };
template <class T>
struct PointerDepth<T*> { enum : int { value = 1 + PointerDepth<T>::value }; using type = typename PointerDepth<T>::type; };

// DataType = value ** ... [S0][S1]... : the run-time extents come first, then the compile-time ones
template <class DataType>
struct Analyze {
  using no_extents = typename std::remove_all_extents<DataType>::type;
  using value_type = typename PointerDepth<no_extents>::type;  // possibly const
  enum : int { rank_dynamic = PointerDepth<no_extents>::value, rank_static = std::rank<DataType>::value,
               rank = rank_dynamic + rank_static };
  template <int I>
  static constexpr size_t static_extent() { return std::extent<DataType, I>::value; }
  static void fill_static(size_t* ext) { fill<0>(ext, std::integral_constant<bool, (rank_static > 0)>()); }

 private:
  template <int I>
  static void fill(size_t* ext, std::true_type) {
    ext[rank_dynamic + I] = std::extent<DataType, I>::value;
    fill<I + 1>(ext, std::integral_constant<bool, (I + 1 < rank_static)>());
  }
  template <int I>
  static void fill(size_t*, std::false_type) {}
};

template <class DataType, class NewValue>
struct ReplaceValue;  // rebuild DataType with another value type (const <-> non-const)
template <class V, class NewValue>
struct ReplaceValueNoExt { using type = NewValue; };
template <class V, class NewValue>
struct ReplaceValueNoExt<V*, NewValue> { using type = typename ReplaceValueNoExt<V, NewValue>::type*; };
template <class DataType, class NewValue>
struct ReplaceValue { using type = typename ReplaceValueNoExt<DataType, NewValue>::type; };
template <class DataType, class NewValue, size_t N>
struct ReplaceValue<DataType[N], NewValue> { using type = typename ReplaceValue<DataType, NewValue>::type[N]; };

template <class... P>
struct HasUnmanaged : std::false_type {};
template <unsigned T, class... P>
struct HasUnmanaged<MemoryTraits<T>, P...> : std::integral_constant<bool, (T & Unmanaged) != 0 || HasUnmanaged<P...>::value> {};
template <class A, class... P>
struct HasUnmanaged<A, P...> : HasUnmanaged<P...> {};

template <class... P>
struct FindTraits { using type = MemoryTraits<0>; };
template <unsigned T, class... P>
struct FindTraits<MemoryTraits<T>, P...> { using type = MemoryTraits<T>; };
template <class A, class... P>
struct FindTraits<A, P...> : FindTraits<P...> {};

template <class... P>
struct FindSpace { using type = HostSpace; };
template <class... P>
struct FindSpace<ScratchMemorySpace, P...> { using type = ScratchMemorySpace; };
template <class A, class... P>
struct FindSpace<A, P...> : FindSpace<P...> {};
}  // namespace Impl

struct ALL_t {};
constexpr ALL_t ALL{};

template <class DataType, class... Props>
class View {
  using A = Impl::Analyze<DataType>;

 public:
  using data_type = DataType;
  using value_type = typename A::value_type;
  using non_const_value_type = typename std::remove_const<value_type>::type;
  using const_value_type = typename std::add_const<value_type>::type;
  using const_data_type = typename Impl::ReplaceValue<DataType, const_value_type>::type;
  using non_const_data_type = typename Impl::ReplaceValue<DataType, non_const_value_type>::type;
  using array_layout = LayoutRight;
  using memory_space = typename Impl::FindSpace<Props...>::type;
  using execution_space = Serial;
  using device_type = memory_space;
  using memory_traits = typename Impl::FindTraits<Props...>::type;
  using size_type = size_t;
  using pointer_type = value_type*;
  using reference_type = value_type&;
  using HostMirror = View<non_const_data_type, LayoutRight, HostSpace>;
  using traits = View;
  enum : int { Rank = A::rank, rank_dynamic = A::rank_dynamic };
  enum : bool { is_managed = !Impl::HasUnmanaged<Props...>::value };
  static constexpr int rank() { return Rank; }

  View() { init_extents(); }

  // allocating: label + the run-time extents
  template <class... Ints>
  explicit View(const std::string& label, Ints... n) : m_label(label) {
    set_extents(n...);
    // 64-byte aligned and value-initialised, as Kokkos allocations are (the AVX vector packs need the alignment)
    const size_t cnt = span();
    non_const_value_type* p = nullptr;
    if (cnt) {
      void* raw = nullptr;
      if (posix_memalign(&raw, Impl::MEMORY_ALIGNMENT, cnt * sizeof(non_const_value_type)) != 0) std::abort();
      p = static_cast<non_const_value_type*>(raw);
      for (size_t i = 0; i < cnt; ++i) new (p + i) non_const_value_type();
    }
    m_owner = std::shared_ptr<void>(p, [cnt](void* q) {
      non_const_value_type* t = static_cast<non_const_value_type*>(q);
      for (size_t i = 0; i < cnt; ++i) t[i].~non_const_value_type();
      std::free(q);
    });
    m_ptr = p;
  }
  explicit View(const char* label) : View(std::string(label)) {}
  // wrapping user memory
  template <class... Ints>
  explicit View(value_type* ptr, Ints... n) : m_ptr(ptr) { set_extents(n...); }
  // scratch views are never used on the host path (Memory<ExeSpace>::get_shmem returns null there)
  explicit View(const ScratchMemorySpace&) { init_extents(); }

  // same shape, compatible value type (adds const, changes traits / managed-ness)
  template <class ODT, class... OP,
            class = typename std::enable_if<
                std::is_convertible<typename View<ODT, OP...>::value_type*, value_type*>::value &&
                int(View<ODT, OP...>::Rank) == int(Rank)>::type>
  View(const View<ODT, OP...>& o) : m_ptr(o.data()), m_owner(o.owner()), m_label(o.label()) {
    for (int i = 0; i < 8; ++i) m_ext[i] = o.extent(i);
    check_static();
  }

  template <class... Ints>
  KOKKOS_FORCEINLINE_FUNCTION reference_type operator()(Ints... idx) const {
    static_assert(sizeof...(Ints) == size_t(Rank), "View: wrong number of indices");
    return m_ptr[offset(0, size_t(0), idx...)];
  }
  KOKKOS_FORCEINLINE_FUNCTION reference_type operator()() const { return m_ptr[0]; }
  template <class I>
  KOKKOS_FORCEINLINE_FUNCTION reference_type operator[](I i) const {
    static_assert(Rank == 1, "View::operator[] needs rank 1");
    return m_ptr[i];
  }

  KOKKOS_INLINE_FUNCTION pointer_type data() const { return m_ptr; }
  KOKKOS_INLINE_FUNCTION pointer_type ptr_on_device() const { return m_ptr; }
  KOKKOS_INLINE_FUNCTION size_t extent(int i) const { return i < 8 ? m_ext[i] : 1; }
  KOKKOS_INLINE_FUNCTION int extent_int(int i) const { return int(extent(i)); }
  KOKKOS_INLINE_FUNCTION size_t size() const { return span(); }
  KOKKOS_INLINE_FUNCTION size_t span() const {
    size_t s = 1;
    for (int i = 0; i < Rank; ++i) s *= m_ext[i];
    return s;
  }
  KOKKOS_INLINE_FUNCTION constexpr bool span_is_contiguous() const { return true; }
  const std::string& label() const { return m_label; }
  const std::shared_ptr<void>& owner() const { return m_owner; }
  int use_count() const { return int(m_owner.use_count()); }

  struct Map {
    const View* v;
    template <class... Ints>
    KOKKOS_FORCEINLINE_FUNCTION reference_type reference(Ints... idx) const { return (*v)(idx...); }
  };
  KOKKOS_INLINE_FUNCTION Map implementation_map() const { return Map{this}; }

 private:
  void init_extents() {
    for (int i = 0; i < 8; ++i) m_ext[i] = (i < Rank) ? 0 : 1;
    A::fill_static(m_ext);
  }
  void check_static() const {
#ifndef NDEBUG
    size_t st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    A::fill_static(st);
    for (int i = A::rank_dynamic; i < Rank; ++i) assert(st[i] == m_ext[i]);
#endif
  }
  template <class... Ints>
  void set_extents(Ints... n) {
    init_extents();
    const size_t dyn[] = {size_t(n)..., 0};
    // Kokkos accepts the static extents repeated after the dynamic ones; only the dynamic ones are taken
    for (int i = 0; i < A::rank_dynamic && i < int(sizeof...(Ints)); ++i) m_ext[i] = dyn[i];
  }
  KOKKOS_FORCEINLINE_FUNCTION size_t offset(int, size_t acc) const { return acc; }
  template <class I0, class... Ints>
  KOKKOS_FORCEINLINE_FUNCTION size_t offset(int d, size_t acc, I0 i0, Ints... rest) const {
    assert(size_t(i0) < m_ext[d]);
    return offset(d + 1, acc * m_ext[d] + size_t(i0), rest...);
  }

  pointer_type m_ptr = nullptr;
  size_t m_ext[8];
  std::shared_ptr<void> m_owner;
  std::string m_label;
};

namespace Impl {
// ViewMapping<DstTraits, SrcTraits, void>::is_assignable — the reference overloads its host<->device sync helpers on it
// (utilities/SyncUtils.hpp): same rank, compatible value type, no run-time extent where the source has a compile-time
// one missing, and equal compile-time extents where the destination has them.
template <class Dst, class Src, int I, bool More>
struct StaticExtentsMatch {
  enum : bool {
    here = (I < int(Dst::rank_dynamic)) ||
           (std::extent<typename Dst::data_type, (I >= int(Dst::rank_dynamic) ? I - int(Dst::rank_dynamic) : 0)>::value ==
            std::extent<typename Src::data_type, (I >= int(Src::rank_dynamic) ? I - int(Src::rank_dynamic) : 0)>::value),
    value = here && StaticExtentsMatch<Dst, Src, I + 1, (I + 1 < int(Dst::Rank))>::value
  };
};
template <class Dst, class Src, int I>
struct StaticExtentsMatch<Dst, Src, I, false> { enum : bool { value = true }; };
template <class Dst, class Src, bool SameRank>
struct DimsAssignable { enum : bool { value = false }; };
template <class Dst, class Src>
struct DimsAssignable<Dst, Src, true> {
  enum : bool { value = int(Dst::rank_dynamic) >= int(Src::rank_dynamic) &&
                        StaticExtentsMatch<Dst, Src, 0, (int(Dst::Rank) > 0)>::value };
};
template <class DstTraits, class SrcTraits, class Enable = void>
struct ViewMapping {
  enum : bool {
    is_assignable = std::is_convertible<typename SrcTraits::value_type*, typename DstTraits::value_type*>::value &&
                    DimsAssignable<DstTraits, SrcTraits, int(DstTraits::Rank) == int(SrcTraits::Rank)>::value
  };
};
}  // namespace Impl

template <class T>
struct is_view : std::false_type {};
template <class D, class... P>
struct is_view<View<D, P...>> : std::true_type {};

// subview(v, i, j, ..., ALL, ALL): leading indices, trailing ALLs — the only form the reference uses
// (mpi/BoundaryExchange.hpp:288-401); the result is the contiguous LayoutRight block behind those indices.
namespace Impl {
template <class T, int N>
struct AddPointers { using type = typename AddPointers<T, N - 1>::type*; };
template <class T>
struct AddPointers<T, 0> { using type = T; };
template <class... Args>
struct CountAll { enum : int { value = 0 }; };
template <class A, class... Args>
struct CountAll<A, Args...> { enum : int { value = (std::is_same<A, ALL_t>::value ? 1 : 0) + CountAll<Args...>::value }; };
inline size_t sub_index(ALL_t) { return 0; }
template <class I>
inline size_t sub_index(I i) { return size_t(i); }
template <class V, size_t... K>
V make_sub(typename V::pointer_type p, const size_t* ext, std::index_sequence<K...>) { return V(p, ext[K]...); }
}  // namespace Impl
template <class D, class... P, class... Args>
View<typename Impl::AddPointers<typename View<D, P...>::value_type, Impl::CountAll<Args...>::value>::type, LayoutRight,
     typename View<D, P...>::memory_space, MemoryTraits<Unmanaged>>
subview(const View<D, P...>& v, Args... args) {
  using Src = View<D, P...>;
  constexpr int R = Impl::CountAll<Args...>::value, K = int(sizeof...(Args)) - R;
  static_assert(int(sizeof...(Args)) == int(Src::Rank), "subview: one argument per dimension");
  using Sub = View<typename Impl::AddPointers<typename Src::value_type, R>::type, LayoutRight, typename Src::memory_space,
                   MemoryTraits<Unmanaged>>;
  const size_t idx[] = {Impl::sub_index(args)..., 0};
  size_t off = 0, ext[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  for (int d = 0; d < int(Src::Rank); ++d) {
    off = off * v.extent(d) + idx[d];
    if (d >= K) ext[d - K] = v.extent(d);
  }
  return Impl::make_sub<Sub>(v.data() + off, ext, std::make_index_sequence<size_t(R)>());
}

template <class D, class... P>
typename View<D, P...>::HostMirror create_mirror_view(const View<D, P...>& v) {
  return typename View<D, P...>::HostMirror(v);
}
template <class Space, class D, class... P>
typename View<D, P...>::HostMirror create_mirror_view(const Space&, const View<D, P...>& v) {
  return typename View<D, P...>::HostMirror(v);
}
template <class D, class... P>
typename View<D, P...>::HostMirror create_mirror(const View<D, P...>& v) {
  typename View<D, P...>::HostMirror m(v);
  return m;
}

template <class DD, class... DP, class SD, class... SP>
void deep_copy(const View<DD, DP...>& dst, const View<SD, SP...>& src) {
  assert(dst.span() == src.span());
  if ((const void*)dst.data() == (const void*)src.data()) return;
  using V = typename View<DD, DP...>::non_const_value_type;
  static_assert(std::is_same<V, typename View<SD, SP...>::non_const_value_type>::value, "deep_copy: value types differ");
  const size_t n = dst.span();
  for (size_t i = 0; i < n; ++i) dst.data()[i] = src.data()[i];
}
template <class DD, class... DP>
void deep_copy(const View<DD, DP...>& dst, const typename View<DD, DP...>::non_const_value_type& value) {
  const size_t n = dst.span();
  for (size_t i = 0; i < n; ++i) dst.data()[i] = value;
}

// ---- policies -----------------------------------------------------------------------------------------------
namespace Impl {
template <class... Args>
struct FindTag { using type = void; };
template <class A, class... Args>
struct FindTag<A, Args...> {
  using type = typename std::conditional<std::is_same<A, Serial>::value || std::is_same<A, DefaultExecutionSpace>::value ||
                                             std::is_void<A>::value,
                                         typename FindTag<Args...>::type, A>::type;
};
template <class T, class... Args>
struct FindTag<IndexType<T>, Args...> : FindTag<Args...> {};

struct TeamMember {
  int m_league_rank, m_league_size;
  KOKKOS_INLINE_FUNCTION int league_rank() const { return m_league_rank; }
  KOKKOS_INLINE_FUNCTION int league_size() const { return m_league_size; }
  KOKKOS_INLINE_FUNCTION int team_rank() const { return 0; }
  KOKKOS_INLINE_FUNCTION int team_size() const { return 1; }
  KOKKOS_INLINE_FUNCTION void team_barrier() const {}
  KOKKOS_INLINE_FUNCTION ScratchMemorySpace team_shmem() const { return ScratchMemorySpace(); }
  KOKKOS_INLINE_FUNCTION ScratchMemorySpace team_scratch(int) const { return ScratchMemorySpace(); }
  KOKKOS_INLINE_FUNCTION ScratchMemorySpace thread_scratch(int) const { return ScratchMemorySpace(); }
};
template <class ExecSpace>
struct TeamPolicyInternal { using member_type = TeamMember; };

template <class iType, class Member>
struct TeamThreadRangeBoundariesStruct {
  iType start, end;
  const Member& thread;
  TeamThreadRangeBoundariesStruct(const Member& t, iType n) : start(0), end(n), thread(t) {}
  TeamThreadRangeBoundariesStruct(const Member& t, iType a, iType b) : start(a), end(b), thread(t) {}
};
template <class iType, class Member>
struct ThreadVectorRangeBoundariesStruct {
  iType start, end;
  ThreadVectorRangeBoundariesStruct(iType n) : start(0), end(n) {}
  ThreadVectorRangeBoundariesStruct(const Member&, iType n) : start(0), end(n) {}
  ThreadVectorRangeBoundariesStruct(const Member&, iType a, iType b) : start(a), end(b) {}
};
struct ThreadSingle { const TeamMember& m; };
struct TeamSingle { const TeamMember& m; };

template <class Tag, class F, class... A>
KOKKOS_FORCEINLINE_FUNCTION typename std::enable_if<std::is_void<Tag>::value>::type call(const F& f, A&&... a) {
  f(std::forward<A>(a)...);
}
template <class Tag, class F, class... A>
KOKKOS_FORCEINLINE_FUNCTION typename std::enable_if<!std::is_void<Tag>::value>::type call(const F& f, A&&... a) {
  f(Tag(), std::forward<A>(a)...);
}

struct FunctorPatternInterface { struct SCAN {}; };
template <class Pattern, class Policy, class Functor>
struct FunctorAnalysis { using value_type = double; };
}  // namespace Impl

template <class... Args>
class TeamPolicy {
 public:
  using member_type = Impl::TeamMember;
  using work_tag = typename Impl::FindTag<Args...>::type;
  using execution_space = Serial;
  TeamPolicy() = default;
  TeamPolicy(int league, int /*team*/, int /*vector*/ = 1) : m_league(league) {}
  template <class Space>
  TeamPolicy(const Space&, int league, int /*team*/, int /*vector*/ = 1) : m_league(league) {}
  int league_size() const { return m_league; }
  int team_size() const { return 1; }
  int vector_length() const { return 1; }
  int chunk_size() const { return 1; }
  TeamPolicy& set_chunk_size(int) { return *this; }
  template <class... X>
  TeamPolicy& set_scratch_size(X...) { return *this; }
  template <class F>
  static int team_size_max(const F&) { return 1; }
  template <class F>
  static int team_size_recommended(const F&) { return 1; }

 private:
  int m_league = 0;
};

template <class... Args>
class RangePolicy {
 public:
  using work_tag = typename Impl::FindTag<Args...>::type;
  using execution_space = Serial;
  RangePolicy(long b, long e) : m_begin(b), m_end(e) {}
  long begin() const { return m_begin; }
  long end() const { return m_end; }
  RangePolicy& set_chunk_size(int) { return *this; }

 private:
  long m_begin, m_end;
};

namespace Experimental {
namespace Iterate { struct Right {}; struct Left {}; struct Default {}; }
template <int N, class Outer = Iterate::Default, class Inner = Iterate::Default>
struct Rank { enum : int { rank = N }; };
template <class... Args>
class MDRangePolicy;
template <class Space, int N, class O, class I, class... Rest>
class MDRangePolicy<Space, Rank<N, O, I>, Rest...> {
 public:
  enum : int { rank = N };
  using work_tag = void;
  using point_type = Kokkos::Array<long, N>;
  MDRangePolicy(const std::initializer_list<long>& lo, const std::initializer_list<long>& hi,
                const std::initializer_list<long>& = {}) {
    int i = 0;
    for (long v : lo) m_lo[i++] = v;
    i = 0;
    for (long v : hi) m_hi[i++] = v;
  }
  long m_lo[N], m_hi[N];
};
}  // namespace Experimental

template <class Member>
KOKKOS_INLINE_FUNCTION Impl::TeamThreadRangeBoundariesStruct<int, Member> TeamThreadRange(const Member& t, int n) {
  return Impl::TeamThreadRangeBoundariesStruct<int, Member>(t, n);
}
template <class Member>
KOKKOS_INLINE_FUNCTION Impl::TeamThreadRangeBoundariesStruct<int, Member> TeamThreadRange(const Member& t, int a, int b) {
  return Impl::TeamThreadRangeBoundariesStruct<int, Member>(t, a, b);
}
template <class Member>
KOKKOS_INLINE_FUNCTION Impl::ThreadVectorRangeBoundariesStruct<int, Member> ThreadVectorRange(const Member& t, int n) {
  return Impl::ThreadVectorRangeBoundariesStruct<int, Member>(t, n);
}
KOKKOS_INLINE_FUNCTION Impl::ThreadSingle PerThread(const Impl::TeamMember& m) { return Impl::ThreadSingle{m}; }
KOKKOS_INLINE_FUNCTION Impl::TeamSingle PerTeam(const Impl::TeamMember& m) { return Impl::TeamSingle{m}; }
template <class F>
KOKKOS_FORCEINLINE_FUNCTION void single(const Impl::ThreadSingle&, const F& f) { f(); }
template <class F>
KOKKOS_FORCEINLINE_FUNCTION void single(const Impl::TeamSingle&, const F& f) { f(); }

// ---- parallel patterns: the sequential loops the Serial back end is equivalent to ------------------------------
template <class... A, class F>
void parallel_for(const TeamPolicy<A...>& p, const F& f) {
  using Tag = typename TeamPolicy<A...>::work_tag;
  const int n = p.league_size();
  REF_SHIM_PARALLEL_FOR
  for (int l = 0; l < n; ++l) Impl::call<Tag>(f, Impl::TeamMember{l, n});
}
template <class... A, class F>
void parallel_for(const RangePolicy<A...>& p, const F& f) {
  using Tag = typename RangePolicy<A...>::work_tag;
  const long b = p.begin(), e = p.end();
  REF_SHIM_PARALLEL_FOR
  for (long i = b; i < e; ++i) Impl::call<Tag>(f, int(i));
}
template <class Space, class O, class I, class... R, class F>
void parallel_for(const Experimental::MDRangePolicy<Space, Experimental::Rank<2, O, I>, R...>& p, const F& f) {
  const long lo0 = p.m_lo[0], hi0 = p.m_hi[0];
  REF_SHIM_PARALLEL_FOR
  for (long i = lo0; i < hi0; ++i)
    for (long j = p.m_lo[1]; j < p.m_hi[1]; ++j) f(int(i), int(j));
}
template <class Space, class O, class I, class... R, class F>
void parallel_for(const Experimental::MDRangePolicy<Space, Experimental::Rank<3, O, I>, R...>& p, const F& f) {
  const long lo0 = p.m_lo[0], hi0 = p.m_hi[0];
  REF_SHIM_PARALLEL_FOR
  for (long i = lo0; i < hi0; ++i)
    for (long j = p.m_lo[1]; j < p.m_hi[1]; ++j)
      for (long k = p.m_lo[2]; k < p.m_hi[2]; ++k) f(int(i), int(j), int(k));
}
template <class Policy, class F>
void parallel_for(const std::string&, const Policy& p, const F& f) { parallel_for(p, f); }
template <class F>
void parallel_for(size_t n, const F& f) {
  for (size_t i = 0; i < n; ++i) f(int(i));
}
template <class iType, class M, class F>
KOKKOS_FORCEINLINE_FUNCTION void parallel_for(const Impl::TeamThreadRangeBoundariesStruct<iType, M>& r, const F& f) {
  for (iType i = r.start; i < r.end; ++i) f(i);
}
template <class iType, class M, class F>
KOKKOS_FORCEINLINE_FUNCTION void parallel_for(const Impl::ThreadVectorRangeBoundariesStruct<iType, M>& r, const F& f) {
  for (iType i = r.start; i < r.end; ++i) f(i);
}

template <class iType, class M, class F, class V>
KOKKOS_FORCEINLINE_FUNCTION void parallel_reduce(const Impl::TeamThreadRangeBoundariesStruct<iType, M>& r, const F& f, V& result) {
  result = V();
  for (iType i = r.start; i < r.end; ++i) f(i, result);
}
template <class iType, class M, class F, class V>
KOKKOS_FORCEINLINE_FUNCTION void parallel_reduce(const Impl::ThreadVectorRangeBoundariesStruct<iType, M>& r, const F& f, V& result) {
  result = V();
  for (iType i = r.start; i < r.end; ++i) f(i, result);
}
template <class... A, class F, class V>
void parallel_reduce(const RangePolicy<A...>& p, const F& f, V& result) {
  using Tag = typename RangePolicy<A...>::work_tag;
  result = V();
  for (long i = p.begin(); i < p.end(); ++i) Impl::call<Tag>(f, int(i), result);
}
template <class... A, class F, class V>
void parallel_reduce(const TeamPolicy<A...>& p, const F& f, V& result) {
  using Tag = typename TeamPolicy<A...>::work_tag;
  result = V();
  for (int l = 0; l < p.league_size(); ++l) Impl::call<Tag>(f, Impl::TeamMember{l, p.league_size()}, result);
}
template <class Policy, class F, class V>
void parallel_reduce(const std::string&, const Policy& p, const F& f, V& result) { parallel_reduce(p, f, result); }

// exclusive scan: lambda(i, accumulator, final)
namespace Impl {
template <class F>
struct ScanValue;  // value type = the lambda's second argument
template <class C, class R, class I, class V, class B>
struct ScanValue<R (C::*)(I, V&, B) const> { using type = V; };
template <class C, class R, class I, class V, class B>
struct ScanValue<R (C::*)(I, V&, B)> { using type = V; };
}  // namespace Impl
template <class iType, class M, class F>
KOKKOS_FORCEINLINE_FUNCTION void parallel_scan(const Impl::ThreadVectorRangeBoundariesStruct<iType, M>& r, const F& f) {
  using V = typename Impl::ScanValue<decltype(&F::operator())>::type;
  V accum = V();
  for (iType i = r.start; i < r.end; ++i) f(i, accum, true);
}
template <class iType, class M, class F>
KOKKOS_FORCEINLINE_FUNCTION void parallel_scan(const Impl::TeamThreadRangeBoundariesStruct<iType, M>& r, const F& f) {
  using V = typename Impl::ScanValue<decltype(&F::operator())>::type;
  V accum = V();
  for (iType i = r.start; i < r.end; ++i) f(i, accum, true);
}

}  // namespace Kokkos

#endif  // REF_SHIM_KOKKOS_CORE_HPP
