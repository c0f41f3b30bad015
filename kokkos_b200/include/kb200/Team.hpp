// kb200/Team.hpp -- hierarchical parallelism on the B200 execution space: TeamPolicy parallel_for /
// parallel_reduce, the team handle, nested ranges, team/vector collectives, `single`, team scratch.
//
// Same vocabulary as the reference: member type (core/src/Cuda/Kokkos_Cuda_Team.hpp:72-349), nested
// policies TeamThreadRange / TeamVectorRange / ThreadVectorRange, PerTeam / PerThread, nested
// parallel_for / parallel_reduce / parallel_scan and single (Kokkos_Cuda_Team.hpp:368-1070),
// ScratchMemorySpace with levels 0 (shared memory) and 1 (global arena) (core/src/Kokkos_ScratchSpace.hpp:37-170),
// TeamPolicy launch (core/src/Cuda/Kokkos_Cuda_Parallel_Team.hpp:431-587,589-980).
//
// Mapping: block = (vector_length, team_size); a team "thread" is a group of vector_length lanes of one
// warp; the league is walked by a persistent grid.  All collectives are register/shuffle based with one
// shared-memory hop across warps (Collectives.hpp); the level-1 scratch is a per-block slice of an
// instance-owned arena (no atomicCAS slot pool as in Cuda_Parallel_Team.hpp:392-429).
#ifndef KB200_TEAM_HPP
#define KB200_TEAM_HPP

#include "Parallel.hpp"

// Team-level functions are called from KB200_LAMBDA (= __host__ __device__) functors, so they must be
// __host__ __device__ themselves; their bodies only exist in the device pass (KOKKOS_IF_ON_DEVICE in the reference).
#ifdef __CUDA_ARCH__
#define KB200_TEAM_DEVICE_ONLY(...) __VA_ARGS__
#else
#define KB200_TEAM_DEVICE_ONLY(...)
#endif
#define KB200_TEAM_FUNCTION __host__ __device__ __forceinline__

namespace kb200 {
namespace Impl {
namespace tm {  // device builtins behind host-device wrappers (the host pass never executes them)
KB200_TEAM_FUNCTION int tx() { KB200_TEAM_DEVICE_ONLY(return (int)threadIdx.x;) return 0; }
KB200_TEAM_FUNCTION int ty() { KB200_TEAM_DEVICE_ONLY(return (int)threadIdx.y;) return 0; }
KB200_TEAM_FUNCTION int nx() { KB200_TEAM_DEVICE_ONLY(return (int)blockDim.x;) return 1; }
KB200_TEAM_FUNCTION int ny() { KB200_TEAM_DEVICE_ONLY(return (int)blockDim.y;) return 1; }
KB200_TEAM_FUNCTION void sync() { KB200_TEAM_DEVICE_ONLY(__syncthreads();) }
KB200_TEAM_FUNCTION void syncwarp();
// lane mask of this thread's vector group: vector-level shuffles must name only the group, because the other
// team threads sharing the warp may be in a different iteration (Cuda_Team.hpp does the same with blockDim.x masks)
KB200_TEAM_FUNCTION unsigned vmask() {
  KB200_TEAM_DEVICE_ONLY(
    const int vl = (int)blockDim.x;
    if (vl >= 32) return 0xffffffffu;
    const int lane = (int)(threadIdx.x + blockDim.x * threadIdx.y) & 31;
    return ((1u << vl) - 1u) << ((lane / vl) * vl);
  )
  return 0xffffffffu;
}
template <class T> KB200_TEAM_FUNCTION T up(const T& v, int d) { KB200_TEAM_DEVICE_ONLY(return ::kb200::Impl::shfl_up(v, (unsigned)d, vmask());) return v; }
template <class T> KB200_TEAM_FUNCTION T down(const T& v, int d) { KB200_TEAM_DEVICE_ONLY(return ::kb200::Impl::shfl_down(v, (unsigned)d, vmask());) return v; }
template <class T> KB200_TEAM_FUNCTION T idx(const T& v, int l) { KB200_TEAM_DEVICE_ONLY(return ::kb200::Impl::shfl_idx(v, l, vmask());) return v; }
KB200_TEAM_FUNCTION void syncwarp() { KB200_TEAM_DEVICE_ONLY(__syncwarp(vmask());) }
template <class Red> KB200_TEAM_FUNCTION void block_red(const Red& red, typename Red::value_type& v, void* smem) {
  KB200_TEAM_DEVICE_ONLY(block_reduce(red, v, smem);)
}
template <class T> KB200_TEAM_FUNCTION T fetch_add(T* p, T v) {
  KB200_TEAM_DEVICE_ONLY(
    if constexpr (sizeof(T) == 8 && std::is_integral<T>::value) return (T)atomicAdd((unsigned long long*)p, (unsigned long long)v);
    else return atomicAdd(p, v);
  )
  T o = *p; *p += v; return o;
}
}  // namespace tm
}  // namespace Impl

// ------------------------------------------------------------------------------------------ scratch space
template <class ExecSpace>
class ScratchMemorySpace {
 public:
  using is_scratch_tag = void;
  using memory_space = ScratchMemorySpace;
  using execution_space = ExecSpace;
  using device_type = Device<ExecSpace, ScratchMemorySpace>;
  using array_layout = LayoutLeft;
  using size_type = unsigned int;
  static constexpr int ALIGN = 8;  // Kokkos_ScratchSpace.hpp:44

  KB200_INLINE_FUNCTION ScratchMemorySpace() : m_iter{nullptr, nullptr}, m_end{nullptr, nullptr}, m_mult(1), m_slot(0), m_default_level(0) {}
  KB200_INLINE_FUNCTION ScratchMemorySpace(void* p0, size_t s0, void* p1, size_t s1)
      : m_iter{(char*)p0, (char*)p1}, m_end{(char*)p0 + s0, (char*)p1 + s1}, m_mult(1), m_slot(0), m_default_level(0) {}

  template <class IntType>
  KB200_INLINE_FUNCTION void* get_shmem(const IntType& size, int level = -1) const { return carve((size_t)size, 1, level); }
  template <class IntType>
  KB200_INLINE_FUNCTION void* get_shmem_aligned(const IntType& size, const ptrdiff_t alignment, int level = -1) const {
    return carve((size_t)size, (size_t)alignment, level);
  }
  // One pool per level (Kokkos_ScratchSpace.hpp:95-170): a TEAM request of `size` takes `size` bytes and every thread gets the
  // same pointer; a THREAD request takes size x team_size bytes and thread r gets the r-th slice.  The handle returned by
  // team_scratch()/thread_scratch() is this one object switched between the two modes, so successive requests advance
  // the same cursor and nullptr is returned once the pool is exhausted.
  KB200_INLINE_FUNCTION const ScratchMemorySpace& impl_set_mode(int level, int multiplier, int slot) const {
    m_default_level = level; m_mult = multiplier; m_slot = slot;
    return *this;
  }
  KB200_INLINE_FUNCTION ScratchMemorySpace set_level(int level) const { ScratchMemorySpace s(*this); s.m_default_level = level; return s; }

 private:
  KB200_INLINE_FUNCTION void* carve(size_t size, size_t alignment, int level) const {
    if (level == -1) level = m_default_level;
    char* cur = m_iter[level];
    if (alignment > 1) {
      const size_t mis = reinterpret_cast<uintptr_t>(cur) % alignment;
      if (mis) cur += alignment - mis;
    }
    const size_t need = size * (size_t)m_mult;
    if (cur > m_end[level] || need > (size_t)(m_end[level] - cur)) return nullptr;  // pool exhausted: cursor unchanged
    m_iter[level] = cur + need;
    return cur + (size_t)m_slot * size;
  }
  mutable char* m_iter[2];
  char* m_end[2];
  mutable int m_mult, m_slot;
  mutable int m_default_level;
};

// ------------------------------------------------------------------------------------------ team handle
class B200TeamMember {
 public:
  using execution_space = B200;
  using scratch_memory_space = ScratchMemorySpace<B200>;

  KB200_TEAM_FUNCTION B200TeamMember(void* collective_smem, void* l0, size_t l0_team, size_t l0_thread, void* l1, size_t l1_team,
                                       size_t l1_thread, int league_rank, int league_size)
      : m_collective(collective_smem), m_l0((char*)l0), m_l1((char*)l1), m_l0_team(l0_team), m_l0_thread(l0_thread),
        m_l1_team(l1_team), m_l1_thread(l1_thread), m_league_rank(league_rank), m_league_size(league_size),
        m_scr(l0, l0_team + l0_thread * (size_t)Impl::tm::ny(), l1, l1_team + l1_thread * (size_t)Impl::tm::ny()) {}

  KB200_TEAM_FUNCTION int league_rank() const { return m_league_rank; }
  KB200_TEAM_FUNCTION int league_size() const { return m_league_size; }
  // raw scratch pools of this team (level 0 = shared memory, 1 = global arena): base pointer and bytes, for layers that
  // hand out their own scratch-space handle type over the same memory (kokkos_b200/adapter)
  KB200_TEAM_FUNCTION void* impl_scratch_ptr(int level) const { return level == 0 ? (void*)m_l0 : (void*)m_l1; }
  KB200_TEAM_FUNCTION size_t impl_scratch_bytes(int level) const {
    return level == 0 ? m_l0_team + m_l0_thread * (size_t)Impl::tm::ny() : m_l1_team + m_l1_thread * (size_t)Impl::tm::ny();
  }
  KB200_TEAM_FUNCTION int team_rank() const { return Impl::tm::ty(); }
  KB200_TEAM_FUNCTION int team_size() const { return Impl::tm::ny(); }
  KB200_TEAM_FUNCTION int impl_vector_lane() const { return Impl::tm::tx(); }
  KB200_TEAM_FUNCTION int impl_vector_length() const { return Impl::tm::nx(); }
  KB200_TEAM_FUNCTION void team_barrier() const { Impl::tm::sync(); }

  // scratch handles: ONE pool per level whose cursor lives in the member (Cuda_Team.hpp:86-101 returns the member's scratch
  // space switched to team or thread mode); TestTeamBasic.hpp:165-180 and TestTeam.hpp:920-1036,1130-1180 pin the behaviour
  KB200_TEAM_FUNCTION const scratch_memory_space& team_shmem() const { return team_scratch(0); }
  KB200_TEAM_FUNCTION const scratch_memory_space& team_scratch(int level) const { return m_scr.impl_set_mode(level, 1, 0); }
  KB200_TEAM_FUNCTION const scratch_memory_space& thread_scratch(int level) const {
    return m_scr.impl_set_mode(level, Impl::tm::ny(), Impl::tm::ty());
  }

  // the collective area is sized per launch (Impl::team_collective_bytes): a nested collective on a larger type than it was
  // sized for must fail loudly, not overwrite level-0 scratch
  KB200_TEAM_FUNCTION void impl_need_collective(size_t bytes) const {
    KB200_TEAM_DEVICE_ONLY(
      if (bytes > (size_t)(m_l0 - (char*)m_collective)) {
        if (threadIdx.x == 0 && threadIdx.y == 0) printf("kb200: team collective on a value type too large for this launch (%llu bytes needed)\n", (unsigned long long)bytes);
        __trap();
      }
    )
  }

  // ---- team collectives (all threads of the team must call) ----
  template <class T>
  KB200_TEAM_FUNCTION void team_broadcast(T& val, int thread_id) const {
    impl_need_collective(sizeof(T));
    T* s = reinterpret_cast<T*>(m_collective);
    Impl::tm::sync();
    if (Impl::tm::ty() == thread_id && Impl::tm::tx() == 0) *s = val;
    Impl::tm::sync();
    val = *s;
    Impl::tm::sync();
  }
  template <class Closure, class T>
  KB200_TEAM_FUNCTION void team_broadcast(const Closure& f, T& val, int thread_id) const {
    f(val);
    team_broadcast(val, thread_id);
  }
  // every thread ends with the team-wide result
  template <class Red>
  KB200_TEAM_FUNCTION void team_reduce(const Red& red, typename Red::value_type& value) const {
    using V = typename Red::value_type;
    impl_need_collective(32 * sizeof(V));
    V v = value;
    if (Impl::tm::tx() != 0) red.init(v);  // a thread's value counts once, not once per vector lane
    Impl::tm::sync();
    Impl::tm::block_red(red, v, m_collective);
    V* s = reinterpret_cast<V*>(m_collective);
    Impl::tm::sync();
    if (Impl::tm::tx() == 0 && Impl::tm::ty() == 0) *s = v;
    Impl::tm::sync();
    value = *s;
    Impl::tm::sync();
  }
  template <class Red>
  KB200_TEAM_FUNCTION void team_reduce(const Red& red) const {
    typename Red::value_type v = red.reference();
    team_reduce(red, v);
    red.reference() = v;
  }
  // exclusive scan of `value` over team_rank; optional global accumulator gets the team total (Cuda_Team.hpp:249-256)
  template <class T>
  KB200_TEAM_FUNCTION T team_scan(const T& value, T* const global_accum = nullptr) const {
    impl_need_collective((size_t)(Impl::tm::ny() + 1) * sizeof(T));
    T* s = reinterpret_cast<T*>(m_collective);  // team_size + 1 entries: the area holds team_size + 2 values of 16 B
    Impl::tm::sync();
    if (Impl::tm::tx() == 0) s[Impl::tm::ty() + 1] = value;
    if (Impl::tm::tx() == 0 && Impl::tm::ty() == 0) s[0] = T();
    Impl::tm::sync();
    if (Impl::tm::tx() == 0 && Impl::tm::ty() == 0) {
      T run = T();
      for (int k = 1; k <= Impl::tm::ny(); ++k) { const T c = s[k]; s[k] = run; run += c; }
      T base = T();
      if (global_accum) base = Impl::tm::fetch_add(global_accum, run);
      s[0] = base;
    }
    Impl::tm::sync();
    const T r = s[0] + s[Impl::tm::ty() + 1];
    Impl::tm::sync();
    return r;
  }

 private:
  void* m_collective;
  char *m_l0, *m_l1;
  size_t m_l0_team, m_l0_thread, m_l1_team, m_l1_thread;
  int m_league_rank, m_league_size;
  scratch_memory_space m_scr;  // both levels; cursor state is mutable inside
};

// ------------------------------------------------------------------------------------------ nested policies
namespace Impl {
template <class I> struct TeamThreadRangeStruct { I begin, end; const B200TeamMember& member; };
template <class I> struct ThreadVectorRangeStruct { I begin, end; const B200TeamMember& member; };
template <class I> struct TeamVectorRangeStruct { I begin, end; const B200TeamMember& member; };
struct ThreadSingleStruct { const B200TeamMember& member; };
struct VectorSingleStruct { const B200TeamMember& member; };
}  // namespace Impl

template <class I>
KB200_TEAM_FUNCTION Impl::TeamThreadRangeStruct<I> TeamThreadRange(const B200TeamMember& m, I count) { return {I(0), count, m}; }
template <class I1, class I2>
KB200_TEAM_FUNCTION Impl::TeamThreadRangeStruct<std::common_type_t<I1, I2>> TeamThreadRange(const B200TeamMember& m, I1 b, I2 e) {
  using I = std::common_type_t<I1, I2>;
  return {I(b), I(e), m};
}
template <class I>
KB200_TEAM_FUNCTION Impl::ThreadVectorRangeStruct<I> ThreadVectorRange(const B200TeamMember& m, I count) { return {I(0), count, m}; }
template <class I1, class I2>
KB200_TEAM_FUNCTION Impl::ThreadVectorRangeStruct<std::common_type_t<I1, I2>> ThreadVectorRange(const B200TeamMember& m, I1 b, I2 e) {
  using I = std::common_type_t<I1, I2>;
  return {I(b), I(e), m};
}
template <class I>
KB200_TEAM_FUNCTION Impl::TeamVectorRangeStruct<I> TeamVectorRange(const B200TeamMember& m, I count) { return {I(0), count, m}; }
template <class I1, class I2>
KB200_TEAM_FUNCTION Impl::TeamVectorRangeStruct<std::common_type_t<I1, I2>> TeamVectorRange(const B200TeamMember& m, I1 b, I2 e) {
  using I = std::common_type_t<I1, I2>;
  return {I(b), I(e), m};
}
KB200_TEAM_FUNCTION Impl::ThreadSingleStruct PerTeam(const B200TeamMember& m) { return {m}; }
KB200_TEAM_FUNCTION Impl::VectorSingleStruct PerThread(const B200TeamMember& m) { return {m}; }

// ---- nested parallel_for
template <class I, class L>
KB200_TEAM_FUNCTION void parallel_for(const Impl::TeamThreadRangeStruct<I>& r, const L& f) {
  for (I i = r.begin + (I)Impl::tm::ty(); i < r.end; i += (I)Impl::tm::ny()) f(i);
}
template <class I, class L>
KB200_TEAM_FUNCTION void parallel_for(const Impl::ThreadVectorRangeStruct<I>& r, const L& f) {
  for (I i = r.begin + (I)Impl::tm::tx(); i < r.end; i += (I)Impl::tm::nx()) f(i);
  Impl::tm::syncwarp();
}
template <class I, class L>
KB200_TEAM_FUNCTION void parallel_for(const Impl::TeamVectorRangeStruct<I>& r, const L& f) {
  for (I i = r.begin + (I)(Impl::tm::ty() * Impl::tm::nx() + Impl::tm::tx()); i < r.end; i += (I)(Impl::tm::ny() * Impl::tm::nx())) f(i);
}

namespace Impl {
// reduce across the vector lanes of one thread (power-of-two groups inside a warp); every lane gets the result
template <class Red>
KB200_TEAM_FUNCTION void vector_reduce(const Red& red, typename Red::value_type& v) {
  using V = typename Red::value_type;
  // ordered fold inside the group: shfl_down tree, then broadcast from the group's first lane
  const int vl = Impl::tm::nx();
  const int lane = (Impl::tm::tx() + Impl::tm::nx() * Impl::tm::ty()) & 31;
  for (int d = 1; d < vl; d <<= 1) {
    V hi = tm::down(v, d);
    if (Impl::tm::tx() + d < vl) red.join(v, hi);
  }
  v = tm::idx(v, lane - Impl::tm::tx());
}
template <class T> struct NestedSum {
  using value_type = T;
  KB200_TEAM_FUNCTION void init(T& v) const { v = T(); }
  KB200_TEAM_FUNCTION void join(T& d, const T& s) const { d += s; }
  KB200_TEAM_FUNCTION void final(T&) const {}
};
// a nested functor that brings its own init/join acts as its reducer (TestTeam.hpp:1844-1870); final() is not applied at this
// level, exactly like the reducer-object form
template <class L, class T> struct NestedFunctorReducer {
  using value_type = T;
  const L& f;
  KB200_TEAM_FUNCTION void init(T& v) const { if constexpr (has_init<L, T>::value) f.init(v); else v = T(); }
  KB200_TEAM_FUNCTION void join(T& d, const T& s) const { if constexpr (has_join<L, T>::value) f.join(d, s); else d += s; }
  KB200_TEAM_FUNCTION void final(T&) const {}
};
template <class L, class T>
using nested_reducer_t = std::conditional_t<has_join<L, T>::value || has_init<L, T>::value, NestedFunctorReducer<L, T>, NestedSum<T>>;
template <class L, class T>
KB200_TEAM_FUNCTION nested_reducer_t<L, T> make_nested_reducer(const L& f) {
  if constexpr (has_join<L, T>::value || has_init<L, T>::value) return NestedFunctorReducer<L, T>{f};
  else return NestedSum<T>{};
}
}  // namespace Impl

// ---- nested parallel_reduce: result is a scalar (sum, or the functor's own init/join) or a reducer
template <class I, class L, class R>
KB200_TEAM_FUNCTION void parallel_reduce(const Impl::ThreadVectorRangeStruct<I>& r, const L& f, R&& result) {
  using RD = std::decay_t<R>;
  if constexpr (is_reducer_v<RD>) {
    typename RD::value_type v;
    result.init(v);
    for (I i = r.begin + (I)Impl::tm::tx(); i < r.end; i += (I)Impl::tm::nx()) f(i, v);
    Impl::vector_reduce(Impl::ReducerAdapter<RD>{result}, v);
    result.reference() = v;
  } else {
    const auto red = Impl::make_nested_reducer<L, RD>(f);
    RD v;
    red.init(v);
    for (I i = r.begin + (I)Impl::tm::tx(); i < r.end; i += (I)Impl::tm::nx()) f(i, v);
    Impl::vector_reduce(red, v);
    result = v;
  }
}
template <class I, class L, class R>
KB200_TEAM_FUNCTION void parallel_reduce(const Impl::TeamThreadRangeStruct<I>& r, const L& f, R&& result) {
  using RD = std::decay_t<R>;
  if constexpr (is_reducer_v<RD>) {
    typename RD::value_type v;
    result.init(v);
    for (I i = r.begin + (I)Impl::tm::ty(); i < r.end; i += (I)Impl::tm::ny()) f(i, v);
    r.member.team_reduce(Impl::ReducerAdapter<RD>{result}, v);
    result.reference() = v;
  } else {
    const auto red = Impl::make_nested_reducer<L, RD>(f);
    RD v;
    red.init(v);
    for (I i = r.begin + (I)Impl::tm::ty(); i < r.end; i += (I)Impl::tm::ny()) f(i, v);
    r.member.team_reduce(red, v);
    result = v;
  }
}
template <class I, class L, class R>
KB200_TEAM_FUNCTION void parallel_reduce(const Impl::TeamVectorRangeStruct<I>& r, const L& f, R&& result) {
  using RD = std::decay_t<R>;
  const I start = r.begin + (I)(Impl::tm::ty() * Impl::tm::nx() + Impl::tm::tx()), step = (I)(Impl::tm::ny() * Impl::tm::nx());
  if constexpr (is_reducer_v<RD>) {
    typename RD::value_type v;
    result.init(v);
    for (I i = start; i < r.end; i += step) f(i, v);
    Impl::vector_reduce(Impl::ReducerAdapter<RD>{result}, v);
    r.member.team_reduce(Impl::ReducerAdapter<RD>{result}, v);
    result.reference() = v;
  } else {
    const auto red = Impl::make_nested_reducer<L, RD>(f);
    RD v;
    red.init(v);
    for (I i = start; i < r.end; i += step) f(i, v);
    Impl::vector_reduce(red, v);
    r.member.team_reduce(red, v);
    result = v;
  }
}

// ---- nested parallel_reduce with several results: parallel_reduce(nested_range, f(i, v0&, v1&, ...), r0, r1, ...) where each
//      r_k is a scalar (sum) or a reducer (core/unit_test/TestTeamCombinedReducers.hpp): the values travel as one CombinedValue
//      through the single-result nested reduction above
namespace Impl {
template <class CV, class... Rs>
struct NestedCombinedReducer {
  using reducer = NestedCombinedReducer;
  using value_type = CV;
  CombinedReducers<typename combined_slot<Rs>::reducer_type...> rs;
  CV* out;
  KB200_TEAM_FUNCTION void init(CV& v) const { rs.init(v); }
  KB200_TEAM_FUNCTION void join(CV& d, const CV& s) const { rs.join(d, s); }
  KB200_TEAM_FUNCTION CV& reference() const { return *out; }
};
template <class R, class V>
KB200_TEAM_FUNCTION void nested_store(R& r, const V& v) {
  if constexpr (is_reducer_v<std::decay_t<R>>) r.reference() = v; else r = v;
}
template <class CV, class... Rs, size_t... Is>
KB200_TEAM_FUNCTION void nested_store_all(const CV& v, std::index_sequence<Is...>, Rs&... rs) {
  (nested_store(rs, combined_get<(int)Is>(v)), ...);
}
template <class T> struct is_nested_range : std::false_type {};
template <class I> struct is_nested_range<ThreadVectorRangeStruct<I>> : std::true_type {};
template <class I> struct is_nested_range<TeamThreadRangeStruct<I>> : std::true_type {};
template <class I> struct is_nested_range<TeamVectorRangeStruct<I>> : std::true_type {};
}  // namespace Impl
template <class Range, class L, class R0, class R1, class... Rs, std::enable_if_t<Impl::is_nested_range<Range>::value, int> = 0>
KB200_TEAM_FUNCTION void parallel_reduce(const Range& r, const L& f, R0&& r0, R1&& r1, Rs&&... rs) {
  using CV = Impl::CombinedValue<typename Impl::combined_slot<R0>::value_type, typename Impl::combined_slot<R1>::value_type,
                                 typename Impl::combined_slot<Rs>::value_type...>;
  constexpr int N = 2 + (int)sizeof...(Rs);
  CV result;
  Impl::NestedCombinedReducer<CV, R0, R1, Rs...> red{
      Impl::make_combined_reducers(Impl::combined_slot<R0>::reducer(r0), Impl::combined_slot<R1>::reducer(r1), Impl::combined_slot<Rs>::reducer(rs)...), &result};
  parallel_reduce(r, [&](const decltype(r.begin) i, CV& v) { Impl::CombinedFunctor<const L&, CV, N>{f}(i, v); }, red);
  Impl::nested_store_all(result, std::make_index_sequence<N>{}, r0, r1, rs...);
}

// ---- nested parallel_scan, f(i, partial, final); ThreadVectorRange and TeamThreadRange
//      forms: (range, f)  sum;  (range, f, return_value)  sum + total;  (range, f, reducer)  the reducer's init/join, total in
//      reducer.reference()   (core/src/Cuda/Kokkos_Cuda_Team.hpp:844-1070)
namespace Impl {
// scan over the vector lanes of one thread with an arbitrary (ordered) join; returns the total in every lane
template <class I, class L, class V, class Red>
KB200_TEAM_FUNCTION V vector_scan(const ThreadVectorRangeStruct<I>& r, const L& f, const Red& red) {
  const int vl = tm::nx();
  const int lane = (tm::tx() + tm::nx() * tm::ty()) & 31;
  V carry;
  red.init(carry);
  for (I base = r.begin; base < r.end; base += (I)vl) {
    const I i = base + (I)tm::tx();
    V c;
    red.init(c);
    if (i < r.end) f(i, c, false);
    V incl = c;  // inclusive scan over the group, lower lanes on the left
    for (int d = 1; d < vl; d <<= 1) {
      V lo = tm::up(incl, d);
      if (tm::tx() >= d) { red.join(lo, incl); incl = lo; }
    }
    V ex = carry;  // exclusive prefix = carry (+) inclusive prefix of the lane before me
    V prev = tm::up(incl, 1);
    if (tm::tx() > 0) red.join(ex, prev);
    if (i < r.end) f(i, ex, true);
    V last = tm::idx(incl, lane - tm::tx() + vl - 1);
    red.join(carry, last);
  }
  return carry;
}
}  // namespace Impl
template <class I, class L>
KB200_TEAM_FUNCTION void parallel_scan(const Impl::ThreadVectorRangeStruct<I>& r, const L& f) {
  using V = std::remove_reference_t<typename Impl::scan_arg_of<L, I>::type>;
  (void)Impl::vector_scan<I, L, V>(r, f, Impl::NestedSum<V>{});
}
template <class I, class L, class R>
KB200_TEAM_FUNCTION void parallel_scan(const Impl::ThreadVectorRangeStruct<I>& r, const L& f, R&& result) {
  using RD = std::decay_t<R>;
  if constexpr (is_reducer_v<RD>) {
    using V = typename RD::value_type;
    const V total = Impl::vector_scan<I, L, V>(r, f, Impl::ReducerAdapter<RD>{result});
    result.reference() = total;
  } else {
    result = Impl::vector_scan<I, L, RD>(r, f, Impl::NestedSum<RD>{});
  }
}
template <class I, class L>
KB200_TEAM_FUNCTION void parallel_scan(const Impl::TeamThreadRangeStruct<I>& r, const L& f) {
  using V = std::remove_reference_t<typename Impl::scan_arg_of<L, I>::type>;
  V carry = V();
  const I ts = (I)Impl::tm::ny();
  for (I base = r.begin; base < r.end; base += ts) {  // uniform trip count: team_scan is a team collective
    const I i = base + (I)Impl::tm::ty();
    V c = V();
    if (i < r.end) f(i, c, false);
    V ex = carry + r.member.team_scan(c);
    if (i < r.end) f(i, ex, true);
    V tot = c;
    r.member.team_reduce(Impl::NestedSum<V>{}, tot);
    carry += tot;
  }
}
template <class I, class L, class V>
KB200_TEAM_FUNCTION void parallel_scan(const Impl::TeamThreadRangeStruct<I>& r, const L& f, V& result) {
  V carry = V();
  const I ts = (I)Impl::tm::ny();
  for (I base = r.begin; base < r.end; base += ts) {
    const I i = base + (I)Impl::tm::ty();
    V c = V();
    if (i < r.end) f(i, c, false);
    V ex = carry + r.member.team_scan(c);
    if (i < r.end) f(i, ex, true);
    V tot = c;
    r.member.team_reduce(Impl::NestedSum<V>{}, tot);
    carry += tot;
  }
  result = carry;
}

// ---- single
template <class L>
KB200_TEAM_FUNCTION void single(const Impl::VectorSingleStruct&, const L& f) {
  if (Impl::tm::tx() == 0) f();
  Impl::tm::syncwarp();
}
template <class L>
KB200_TEAM_FUNCTION void single(const Impl::ThreadSingleStruct&, const L& f) {
  if (Impl::tm::tx() == 0 && Impl::tm::ty() == 0) f();
}
template <class L, class T>
KB200_TEAM_FUNCTION void single(const Impl::VectorSingleStruct&, const L& f, T& val) {
  const int lane = (Impl::tm::tx() + Impl::tm::nx() * Impl::tm::ty()) & 31;
  if (Impl::tm::tx() == 0) f(val);
  val = Impl::tm::idx(val, lane - Impl::tm::tx());
}
template <class L, class T>
KB200_TEAM_FUNCTION void single(const Impl::ThreadSingleStruct& s, const L& f, T& val) {
  if (Impl::tm::tx() == 0 && Impl::tm::ty() == 0) f(val);
  s.member.team_broadcast(val, 0);
}

// ------------------------------------------------------------------------------------------ TeamPolicy launch
namespace Impl {

// Shared memory in front of level-0 scratch for the team collectives: team_scan keeps team_size + 2 values of <= 16 bytes,
// block-wide reductions 34 values of the reduction type.  Sized per launch: a fixed worst case (16 KiB) would shrink the L1 data
// cache of every small-team kernel (13 CTAs x 16 KiB of carve-out left 43 KB of L1 for the SpMV gather: 2.8x slower than with it).
constexpr size_t kTeamCollectiveMin = 4096;  // nested team reductions of value types up to 120 bytes whatever the top-level type
KB200_FUNCTION constexpr size_t team_collective_bytes(int team_size, size_t value_bytes) {
  size_t c = kTeamCollectiveMin;
  if ((size_t)(team_size + 2) * 16 > c) c = (size_t)(team_size + 2) * 16;
  if (value_bytes * 34 > c) c = value_bytes * 34;
  return (c + 15) & ~(size_t)15;
}

struct TeamLaunchParams {
  int league_size;
  size_t coll_bytes;    // collective area in front of level-0 scratch (team_collective_bytes)
  size_t l0_team, l0_thread, l1_team, l1_thread;
  char* l1_arena;       // grid * l1_per_team bytes
  size_t l1_per_team;
};

template <class F, class Tag>
__global__ void team_for_kernel(const __grid_constant__ F f, const TeamLaunchParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  for (int lr = blockIdx.x; lr < p.league_size; lr += gridDim.x) {
    B200TeamMember m(smem, smem + p.coll_bytes, p.l0_team, p.l0_thread, p.l1_arena + (size_t)blockIdx.x * p.l1_per_team,
                     p.l1_team, p.l1_thread, lr, p.league_size);
    if constexpr (std::is_void<Tag>::value) f(m); else f(Tag{}, m);
    // user scratch is reused by the next league member (the collective area is fenced by the collectives themselves)
    if (p.l0_team + p.l0_thread + p.l1_team + p.l1_thread != 0 && lr + (int)gridDim.x < p.league_size) __syncthreads();
  }
}

template <class F, class Tag, class Red>
__global__ void team_reduce_kernel(const __grid_constant__ F f, const __grid_constant__ Red red, const TeamLaunchParams p,
                                   const ReduceScratch scratch) {
  using V = typename Red::value_type;
  extern __shared__ __align__(16) unsigned char smem[];
  if (p.league_size <= 0) return reduce_store_identity(red, scratch);
  V acc;
  red.init(acc);
  for (int lr = blockIdx.x; lr < p.league_size; lr += gridDim.x) {
    B200TeamMember m(smem, smem + p.coll_bytes, p.l0_team, p.l0_thread, p.l1_arena + (size_t)blockIdx.x * p.l1_per_team,
                     p.l1_team, p.l1_thread, lr, p.league_size);
    if constexpr (std::is_void<Tag>::value) f(m, acc); else f(Tag{}, m, acc);
    if (lr + (int)gridDim.x < p.league_size) __syncthreads();
  }
  if (threadIdx.x != 0) red.init(acc);  // one contribution per team thread (Cuda_Parallel_Team.hpp:678-800)
  __syncthreads();
  block_reduce(red, acc, smem);
  __syncthreads();
  grid_reduce_and_store(red, acc, scratch, smem);
}

template <class Policy>
struct TeamShape {
  int team, vec, threads, grid;
  size_t smem;
  TeamLaunchParams p;
  // level-0 scratch a functor asks for itself: team_shmem_size(team_size) / shmem_size(team_size)
  // (core/src/impl/Kokkos_FunctorAnalysis.hpp FunctorTeamShmemSize; used by core/unit_test/TestTeamVector.hpp:460-463)
  template <class F, class = void> struct has_team_shmem_size : std::false_type {};
  template <class F> struct has_team_shmem_size<F, std::void_t<decltype(std::declval<const F&>().team_shmem_size(0))>> : std::true_type {};
  template <class F, class = void> struct has_shmem_size : std::false_type {};
  template <class F> struct has_shmem_size<F, std::void_t<decltype(std::declval<const F&>().shmem_size(0))>> : std::true_type {};
  template <class F>
  static size_t functor_shmem(const F& f, int team_size) {
    if constexpr (has_team_shmem_size<F>::value) return (size_t)f.team_shmem_size(team_size);
    else if constexpr (has_shmem_size<F>::value) return (size_t)f.shmem_size(team_size);
    else return 0;
  }
  template <class F>
  int setup(const Policy& pol, const void* kernel, size_t value_bytes, const F& f) {
    HostRuntime rt(pol.space().impl_instance());
    vec = pol.impl_auto_vector_length() ? 1 : pol.impl_vector_length();
    team = pol.impl_auto_team_size() ? (256 / vec > 0 ? 256 / vec : 1) : pol.team_size();
    threads = team * vec;
    if (threads > 1024) throw std::runtime_error("kb200::TeamPolicy: team_size * vector_length exceeds 1024");
    p.league_size = pol.league_size();
    p.l0_team = pol.team_scratch_size(0) + functor_shmem(f, team); p.l0_thread = pol.thread_scratch_size(0);
    p.l1_team = pol.team_scratch_size(1); p.l1_thread = pol.thread_scratch_size(1);
    const size_t l0 = p.l0_team + p.l0_thread * team;
    p.coll_bytes = team_collective_bytes(team, value_bytes);
    smem = p.coll_bytes + l0 + 16;
    if (smem > (size_t)220 * 1024)
      throw std::runtime_error("kb200::TeamPolicy: requested too much level-0 scratch (shared memory) for this team size");
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int bps = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kernel, threads, smem);
    if (bps < 1) bps = 1;
    bps = pol.impl_occupancy_cap(bps);
    const long long cap = (long long)rt.sm_count() * bps;
    grid = (int)(p.league_size < cap ? (p.league_size > 0 ? p.league_size : 1) : cap);
    p.l1_per_team = (p.l1_team + p.l1_thread * team + 255) & ~(size_t)255;
    p.l1_arena = nullptr;
    if (p.l1_per_team) {
      void* arena = nullptr;
      int rc = b200_scratch_get(pol.space().impl_instance(), B200_SCRATCH_TEAM_L1, p.l1_per_team * grid, &arena, nullptr);
      if (rc) return rc;
      p.l1_arena = (char*)arena;
    }
    return 0;
  }
};
}  // namespace Impl

namespace Impl {
// value type of a TeamPolicy reduction functor without a result argument at hand: F::value_type, else the last parameter
// of its (non-template) call operator
template <class F, class = void>
struct team_reduce_value_of {
  template <class C, class R, class A0, class A1> static std::remove_reference_t<A1> pick(R (C::*)(A0, A1) const);
  template <class C, class R, class T, class A0, class A1> static std::remove_reference_t<A1> pick(R (C::*)(T, A0, A1) const);
  using type = decltype(pick(&F::operator()));
};
template <class F>
struct team_reduce_value_of<F, std::void_t<typename F::value_type>> { using type = typename F::value_type; };
template <class F, class = void> struct team_reduce_value_known : std::false_type {};
template <class F> struct team_reduce_value_known<F, std::void_t<typename team_reduce_value_of<F>::type>> : std::true_type {};
template <class F, bool = team_reduce_value_known<F>::value> struct team_reduce_scalar_value_known : std::false_type {};
template <class F> struct team_reduce_scalar_value_known<F, true> : std::integral_constant<bool, !std::is_array<typename team_reduce_value_of<F>::type>::value> {};

// team size bound from the attributes of the kernel that would be launched and the level-0 scratch request
template <class Policy, class F>
int team_size_from_attr(const Policy& pol, const F& f, const cudaFuncAttributes& attr, size_t value_bytes, bool halve) {
  const int vec = pol.impl_vector_length() > 0 ? pol.impl_vector_length() : 1;
  int max_threads = attr.maxThreadsPerBlock > 0 ? attr.maxThreadsPerBlock : 1024;
  if (halve) max_threads /= 2;
  constexpr unsigned lb = Policy::launch_bounds::maxTperB;
  if (lb > 0 && (int)lb < max_threads) max_threads = (int)lb;
  int team = max_threads / vec;
  // level-0 scratch (policy request + what the functor asks for itself) must fit next to the collective area
  while (team > 1) {
    const size_t coll = team_collective_bytes(team, value_bytes);
    const size_t l0 = pol.team_scratch_size(0) + TeamShape<Policy>::functor_shmem(f, team) + pol.thread_scratch_size(0) * (size_t)team;
    if (coll + l0 + 16 <= (size_t)220 * 1024) break;
    team /= 2;
  }
  if (team * vec >= 32) team = (team * vec / 32) * 32 / vec;  // whole warps
  return team > 0 ? team : 1;
}

// the reduction kernel is known exactly (functor wrapper + reducer type): used by the Kokkos::B200 adapter, where the reference's
// FunctorAnalysis supplies the reducer
template <class Red, class Policy, class F>
int team_size_limit_reduce(const Policy& pol, const F& f) {
  using Tag = typename Policy::work_tag;
  cudaFuncAttributes attr{};
  throw_on_error(b200_report_error((int)cudaFuncGetAttributes(&attr, team_reduce_kernel<F, Tag, Red>), "kb200::team_size_max"));
  return team_size_from_attr(pol, f, attr, sizeof(typename Red::value_type), false);
}

template <class Policy, class F, class PatternTag>
int team_size_limit(const Policy& pol, const F& f, const PatternTag&) {
  using Tag = typename Policy::work_tag;
  cudaFuncAttributes attr{};
  size_t value_bytes = 16;
  bool halve = false;
  if constexpr (std::is_same<PatternTag, ParallelReduceTag>::value) {
    if constexpr (team_reduce_scalar_value_known<F>::value) {
      using V = typename team_reduce_value_of<F>::type;
      using Red = std::conditional_t<brings_reduction_members<F, V, Tag>::value, FunctorReducer<F, V, Tag>, DefaultSumReducer<V>>;
      value_bytes = sizeof(V);
      throw_on_error(b200_report_error((int)cudaFuncGetAttributes(&attr, team_reduce_kernel<F, Tag, Red>), "kb200::team_size_max"));
    } else {  // value type only known with the result argument: bound it by the for-kernel of the same functor, halved
      attr.maxThreadsPerBlock = 1024;
      halve = true;
    }
  } else {
    throw_on_error(b200_report_error((int)cudaFuncGetAttributes(&attr, team_for_kernel<F, Tag>), "kb200::team_size_max"));
  }
  return team_size_from_attr(pol, f, attr, value_bytes, halve);
}
}  // namespace Impl

template <class... P, class F>
void parallel_for(const std::string& /*label*/, const TeamPolicy<P...>& pol, const F& f) {
  using Tag = typename TeamPolicy<P...>::work_tag;
  if (pol.league_size() <= 0) return;
  auto k = Impl::team_for_kernel<F, Tag>;
  Impl::TeamShape<TeamPolicy<P...>> sh;
  Impl::throw_on_error(sh.setup(pol, (const void*)k, 16, f));
  Impl::HostRuntime rt(pol.space().impl_instance());
  k<<<sh.grid, dim3(sh.vec, sh.team, 1), sh.smem, rt.stream()>>>(f, sh.p);
  Impl::throw_on_error(rt.check_launch("kb200::team_for_kernel"));
}
template <class... P, class F>
void parallel_for(const TeamPolicy<P...>& pol, const F& f) { parallel_for(std::string(), pol, f); }

namespace Impl {
// TeamPolicy reduction with a runtime-length array value (value_type = T[], value_count): per-thread accumulator arrays as in
// ArrayReduceKernel.hpp; one contribution per team thread (vector lanes other than 0 are reset to the identity)
template <class F, class Tag, class T, int CAP>
__global__ void array_team_reduce_kernel(const __grid_constant__ F f, const TeamLaunchParams p, const int count, T* partials, unsigned* ticket,
                                         T* result) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ T red_smem[32 * CAP];
  T acc[CAP];
  ArrayOps<F, T>::init(f, acc, count);
  for (int lr = blockIdx.x; lr < p.league_size; lr += gridDim.x) {
    B200TeamMember m(smem, smem + p.coll_bytes, p.l0_team, p.l0_thread, p.l1_arena + (size_t)blockIdx.x * p.l1_per_team,
                     p.l1_team, p.l1_thread, lr, p.league_size);
    if constexpr (std::is_void<Tag>::value) f(m, acc); else f(Tag{}, m, acc);
    if (lr + (int)gridDim.x < p.league_size) __syncthreads();
  }
  if (threadIdx.x != 0) ArrayOps<F, T>::init(f, acc, count);
  __syncthreads();
  array_block_reduce<F, T, CAP>(f, acc, count, red_smem);
  __syncthreads();
  array_grid_reduce_and_store<F, T, CAP>(f, acc, count, partials, ticket, result, red_smem);
}
template <class T, class Tag, class F, class... P, int CAP>
int array_reduce_launch(const TeamPolicy<P...>& pol, const F& f, int count, T* rh, T* rd, std::integral_constant<int, CAP>) {
  auto k = array_team_reduce_kernel<F, Tag, T, CAP>;
  TeamShape<TeamPolicy<P...>> sh;
  int rc = sh.setup(pol, (const void*)k, 16, f);
  if (rc) return rc;
  b200_instance* inst = pol.space().impl_instance();
  HostRuntime rt(inst);
  return array_reduce_run<T>(inst, count, sh.grid, rh, rd, [&](T* partials, unsigned* ticket, T* dst) {
    k<<<sh.grid, dim3(sh.vec, sh.team, 1), sh.smem, rt.stream()>>>(f, sh.p, count, partials, ticket, dst);
  });
}
}  // namespace Impl

namespace Impl {
template <class... P, class F, class Red>
void reduce_dispatch(const TeamPolicy<P...>& pol, const F& f, const Red& red, ResultTarget<typename Red::value_type> t) {
  using Tag = typename TeamPolicy<P...>::work_tag;
  using V = typename Red::value_type;
  auto k = team_reduce_kernel<F, Tag, Red>;
  TeamShape<TeamPolicy<P...>> sh;
  throw_on_error(sh.setup(pol, (const void*)k, sizeof(V), f));
  HostRuntime rt(pol.space().impl_instance());
  ReduceScratch s;
  void *slot_dev = nullptr, *slot_host = nullptr;
  throw_on_error(rt.reduce_scratch((size_t)sh.grid * sizeof(V), sizeof(V), t.host != nullptr, &s.partials, &s.ticket, &slot_dev, &slot_host));
  s.result0 = t.host ? slot_dev : (void*)t.dev;
  s.result1 = t.host ? (void*)t.dev : nullptr;
  k<<<sh.grid, dim3(sh.vec, sh.team, 1), sh.smem, rt.stream()>>>(f, red, sh.p, s);
  throw_on_error(rt.check_launch("kb200::team_reduce_kernel"));
  if (t.host) {
    throw_on_error(rt.fence("kb200::parallel_reduce(TeamPolicy): fence to hand the scalar result to the host"));
    memcpy(t.host, slot_host, sizeof(V));
  }
}
}  // namespace Impl

}  // namespace kb200
#endif
