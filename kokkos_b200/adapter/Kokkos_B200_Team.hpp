// Kokkos_B200_Team.hpp -- TeamPolicy on `Kokkos::B200` (included by Kokkos_B200_Space.hpp).
//   Impl::TeamPolicyInternal<Kokkos::B200, Props...>      model: core/src/Cuda/Kokkos_Cuda_Parallel_Team.hpp:50-390 (member set required by
//                                                         core/src/Kokkos_ExecPolicy.hpp:365-508)
//   member_type = Impl::B200AdapterTeamMember             the kb200 team handle (league/team ranks, barrier, team_reduce/scan/broadcast)
//                                                         with the REFERENCE's scratch-space type for team_scratch()/thread_scratch()
//                                                         (Cuda/Kokkos_Cuda_Team.hpp:86-101), so Views over scratch memory work unchanged
//   Kokkos::TeamThreadRange / TeamVectorRange / ThreadVectorRange / PerTeam / PerThread and the nested
//   parallel_for / parallel_reduce / parallel_scan / single overloads on that handle      (Cuda/Kokkos_Cuda_Team.hpp:368-1070)
//   Impl::ParallelFor / ParallelReduce <..., TeamPolicy<...>, Kokkos::B200>                 (Cuda/Kokkos_Cuda_Parallel_Team.hpp:431-1000)
// Kernels: kb200::Impl::team_for_kernel / team_reduce_kernel (kokkos_b200/include/kb200/Team.hpp): team -> CTA, vector lanes ->
// threadIdx.x, register-resident nested reductions, persistent grid over the league.
#ifndef KOKKOS_B200_TEAM_HPP
#define KOKKOS_B200_TEAM_HPP

namespace Kokkos {
namespace Impl {

class B200AdapterTeamMember : public kb200::B200TeamMember {
 public:
  using execution_space      = Kokkos::B200;
  using scratch_memory_space = Kokkos::ScratchMemorySpace<Kokkos::B200>;
  using team_handle          = B200AdapterTeamMember;

  KOKKOS_INLINE_FUNCTION explicit B200AdapterTeamMember(const kb200::B200TeamMember& m)
      : kb200::B200TeamMember(m),
        m_kscratch(m.impl_scratch_ptr(0), m.impl_scratch_bytes(0), m.impl_scratch_ptr(1), m.impl_scratch_bytes(1)) {}

  KOKKOS_INLINE_FUNCTION const scratch_memory_space& team_shmem() const { return m_kscratch.set_team_thread_mode(0, 1, 0); }
  KOKKOS_INLINE_FUNCTION const scratch_memory_space& team_scratch(const int level) const { return m_kscratch.set_team_thread_mode(level, 1, 0); }
  KOKKOS_INLINE_FUNCTION const scratch_memory_space& thread_scratch(const int level) const {
    return m_kscratch.set_team_thread_mode(level, team_size(), team_rank());
  }
  // reference reducers publish reference(); value-returning form of team_reduce as the reference's handle has it
  using kb200::B200TeamMember::team_reduce;

 private:
  mutable scratch_memory_space m_kscratch;
};

template <class... Properties>
class TeamPolicyInternal<Kokkos::B200, Properties...> : public PolicyTraits<Properties...> {
 public:
  using execution_policy = TeamPolicyInternal;
  using traits           = PolicyTraits<Properties...>;
  using execution_space  = Kokkos::B200;
  using member_type      = B200AdapterTeamMember;

  template <class ExecSpace, class... OtherProperties>
  friend class TeamPolicyInternal;

  template <class... OtherProperties>
  TeamPolicyInternal(const TeamPolicyInternal<Kokkos::B200, OtherProperties...>& p)
      : m_space(p.m_space), m_league_size(p.m_league_size), m_team_size(p.m_team_size), m_vector_length(p.m_vector_length),
        m_chunk_size(p.m_chunk_size), m_tune_team(p.m_tune_team), m_tune_vector(p.m_tune_vector) {
    for (int l = 0; l < 2; ++l) { m_team_scratch_size[l] = p.m_team_scratch_size[l]; m_thread_scratch_size[l] = p.m_thread_scratch_size[l]; }
  }

  TeamPolicyInternal(const execution_space& space_, int league_size_, int team_size_request, int vector_length_request = 1)
      : m_space(space_), m_league_size(league_size_), m_team_size(team_size_request),
        m_vector_length(vector_length_request > 0 ? impl_determine_vector_length(vector_length_request) : vector_length_request),
        m_chunk_size(32), m_tune_team(team_size_request <= 0), m_tune_vector(vector_length_request <= 0) {
    if (league_size_ < 0) Kokkos::abort("Kokkos::abort: Requested league size is negative");
    if (m_team_size > 0 && m_vector_length > 0 && m_team_size * m_vector_length > 1024)
      Impl::throw_runtime_exception("Kokkos::TeamPolicy<B200>: requested team_size * vector_length exceeds 1024 threads");
  }
  TeamPolicyInternal(const execution_space& space_, int league_size_, const Kokkos::AUTO_t&, int vector_length_request = 1)
      : TeamPolicyInternal(space_, league_size_, -1, vector_length_request) {}
  TeamPolicyInternal(const execution_space& space_, int league_size_, const Kokkos::AUTO_t&, const Kokkos::AUTO_t&)
      : TeamPolicyInternal(space_, league_size_, -1, -1) {}
  TeamPolicyInternal(const execution_space& space_, int league_size_, int team_size_request, const Kokkos::AUTO_t&)
      : TeamPolicyInternal(space_, league_size_, team_size_request, -1) {}
  TeamPolicyInternal(int league_size_, int team_size_request, int vector_length_request = 1)
      : TeamPolicyInternal(execution_space(), league_size_, team_size_request, vector_length_request) {}
  TeamPolicyInternal(int league_size_, const Kokkos::AUTO_t&, int vector_length_request = 1)
      : TeamPolicyInternal(execution_space(), league_size_, -1, vector_length_request) {}
  TeamPolicyInternal(int league_size_, const Kokkos::AUTO_t&, const Kokkos::AUTO_t&) : TeamPolicyInternal(execution_space(), league_size_, -1, -1) {}
  TeamPolicyInternal(int league_size_, int team_size_request, const Kokkos::AUTO_t&)
      : TeamPolicyInternal(execution_space(), league_size_, team_size_request, -1) {}

  const execution_space& space() const { return m_space; }
  inline static int vector_length_max() { return 32; }
  inline static int impl_determine_vector_length(int requested) {  // clamp to a warp, round DOWN to a power of two (Cuda_Parallel_Team.hpp:176-181)
    int v = requested > 32 ? 32 : requested, p2 = 1;
    while (p2 * 2 <= v) p2 *= 2;
    return p2;
  }
  inline static int scratch_size_max(int level) { return level == 0 ? 200 * 1024 : (1 << 30); }
  inline int impl_vector_length() const { return m_vector_length; }
  inline int team_size() const { return m_team_size; }
  inline int league_size() const { return m_league_size; }
  inline bool impl_auto_team_size() const { return m_tune_team; }
  inline bool impl_auto_vector_length() const { return m_tune_vector; }
  inline void impl_set_team_size(size_t team_size) { m_team_size = (int)team_size; }
  inline void impl_set_vector_length(size_t vector_length) { m_vector_length = (int)vector_length; }
  size_t scratch_size(int level, int team_size_ = -1) const {
    if (team_size_ < 0) team_size_ = m_team_size > 0 ? m_team_size : 1;
    return m_team_scratch_size[level] + (size_t)team_size_ * m_thread_scratch_size[level];
  }
  size_t team_scratch_size(int level) const { return m_team_scratch_size[level]; }
  size_t thread_scratch_size(int level) const { return m_thread_scratch_size[level]; }
  inline int chunk_size() const { return m_chunk_size; }
  inline TeamPolicyInternal& set_chunk_size(typename traits::index_type chunk_size_) { m_chunk_size = (int)chunk_size_; return *this; }
  inline TeamPolicyInternal& set_scratch_size(int level, const PerTeamValue& per_team) { m_team_scratch_size[level] = per_team.value; return *this; }
  inline TeamPolicyInternal& set_scratch_size(int level, const PerThreadValue& per_thread) { m_thread_scratch_size[level] = per_thread.value; return *this; }
  inline TeamPolicyInternal& set_scratch_size(int level, const PerTeamValue& per_team, const PerThreadValue& per_thread) {
    m_team_scratch_size[level] = per_team.value; m_thread_scratch_size[level] = per_thread.value; return *this;
  }

  // ---- translation to the kernel layer's policy (work tag handled by the functor wrappers below) ----
  using kb_policy = kb200::TeamPolicy<kb200::B200, kb200::LaunchBounds<traits::launch_bounds::maxTperB, traits::launch_bounds::minBperSM>>;
  kb_policy impl_to_kb() const {
    kb_policy k = m_team_size > 0 ? (m_vector_length > 0 ? kb_policy(m_space.impl_kb200(), m_league_size, m_team_size, m_vector_length)
                                                          : kb_policy(m_space.impl_kb200(), m_league_size, m_team_size, 1))
                                  : kb_policy(m_space.impl_kb200(), m_league_size, kb200::AUTO, m_vector_length > 0 ? m_vector_length : 1);
    for (int l = 0; l < 2; ++l) k.set_scratch_size(l, kb200::PerTeam(m_team_scratch_size[l]), kb200::PerThread(m_thread_scratch_size[l]));
    return k;
  }

  template <class FunctorType> int team_size_max(const FunctorType& f, const ParallelForTag&) const;
  template <class FunctorType> int team_size_max(const FunctorType& f, const ParallelReduceTag&) const;
  template <class FunctorType, class ReducerType> int team_size_max(const FunctorType& f, const ReducerType&, const ParallelReduceTag& t) const { return team_size_max(f, t); }
  template <class FunctorType> int team_size_recommended(const FunctorType& f, const ParallelForTag& t) const {
    const int mx = team_size_max(f, t), dflt = 256 / (m_vector_length > 0 ? m_vector_length : 1);
    return dflt < mx ? (dflt > 0 ? dflt : 1) : mx;
  }
  template <class FunctorType> int team_size_recommended(const FunctorType& f, const ParallelReduceTag& t) const {
    const int mx = team_size_max(f, t), dflt = 256 / (m_vector_length > 0 ? m_vector_length : 1);
    return dflt < mx ? (dflt > 0 ? dflt : 1) : mx;
  }
  template <class FunctorType, class ReducerType> int team_size_recommended(const FunctorType& f, const ReducerType&, const ParallelReduceTag& t) const { return team_size_recommended(f, t); }

 private:
  execution_space m_space;
  int m_league_size, m_team_size, m_vector_length;
  size_t m_team_scratch_size[2] = {0, 0};
  size_t m_thread_scratch_size[2] = {0, 0};
  int m_chunk_size;
  bool m_tune_team, m_tune_vector;
};

namespace B200Adapter {
// the kernels call f(kb200 member [, acc]); the user functor wants the adapter member (and possibly a work tag first)
template <class F, class Tag>
struct TeamForWrap {
  F f;
  KOKKOS_INLINE_FUNCTION void operator()(const kb200::B200TeamMember& m) const {
    const B200AdapterTeamMember am(m);
    if constexpr (std::is_void_v<Tag>) f(am); else f(Tag(), am);
  }
  size_t team_shmem_size(int team_size) const { return (size_t)Kokkos::Impl::FunctorTeamShmemSize<F>::value(f, team_size); }
};
// V: the scalar value type, or T[] for a runtime-length array reduction (the accumulator then arrives as T*)
template <class F, class Tag, class V>
struct TeamReduceWrap {
  using value_type = V;
  using T          = std::remove_extent_t<V>;
  F f;
  template <class Acc>
  KOKKOS_INLINE_FUNCTION void operator()(const kb200::B200TeamMember& m, Acc&& acc) const {
    const B200AdapterTeamMember am(m);
    if constexpr (std::is_void_v<Tag>) f(am, static_cast<Acc&&>(acc)); else f(Tag(), am, static_cast<Acc&&>(acc));
  }
  // array reductions: the functor's own init/join/final if it has them, else value-init and += (what ArrayOps would do)
  KOKKOS_INLINE_FUNCTION void init(T* a) const {
    if constexpr (kb200::Impl::has_array_init<F, T>::value) f.init(a);
    else for (int c = 0; c < (int)f.value_count; ++c) a[c] = T();
  }
  KOKKOS_INLINE_FUNCTION void join(T* d, const T* s) const {
    if constexpr (kb200::Impl::has_array_join<F, T>::value) f.join(d, s);
    else for (int c = 0; c < (int)f.value_count; ++c) d[c] += s[c];
  }
  KOKKOS_INLINE_FUNCTION void final(T* a) const {
    if constexpr (kb200::Impl::has_array_final<F, T>::value) f.final(a);
  }
  size_t team_shmem_size(int team_size) const { return (size_t)Kokkos::Impl::FunctorTeamShmemSize<F>::value(f, team_size); }
};
// stands in for a reduction functor where only launch limits are asked for (no call operator of the user's is instantiated)
template <class F>
struct TeamNullWrap {
  F f;
  KOKKOS_INLINE_FUNCTION void operator()(const kb200::B200TeamMember&) const {}
  size_t team_shmem_size(int team_size) const { return (size_t)Kokkos::Impl::FunctorTeamShmemSize<F>::value(f, team_size); }
};
}  // namespace B200Adapter

template <class... Properties>
template <class FunctorType>
int TeamPolicyInternal<Kokkos::B200, Properties...>::team_size_max(const FunctorType& f, const ParallelForTag&) const {
  using W = B200Adapter::TeamForWrap<FunctorType, typename traits::work_tag>;
  return impl_to_kb().team_size_max(W{f}, kb200::ParallelForTag());
}
template <class... Properties>
template <class FunctorType>
int TeamPolicyInternal<Kokkos::B200, Properties...>::team_size_max(const FunctorType& f, const ParallelReduceTag&) const {
  // the reference's own analysis names the value type and reducer a result-less call would use (model:
  // Cuda/Kokkos_Cuda_Parallel_Team.hpp:111-128); the bound then comes from the reduction kernel that would be launched
  using Analysis = Impl::FunctorAnalysis<Impl::FunctorPatternInterface::REDUCE, TeamPolicyInternal, FunctorType, void>;
  if constexpr (Analysis::StaticValueSize != 0) {
    using V = typename Analysis::value_type;
    using W = B200Adapter::TeamReduceWrap<FunctorType, typename traits::work_tag, V>;
    using R = B200Adapter::Red<typename Analysis::Reducer>;
    return kb200::Impl::team_size_limit_reduce<R>(impl_to_kb(), W{f});
  } else {  // runtime-length array value: bound by the for-kernel of the same functor, halved
    const int mx = impl_to_kb().team_size_max(B200Adapter::TeamNullWrap<FunctorType>{f}, kb200::ParallelForTag());
    return mx > 1 ? mx / 2 : 1;
  }
}

template <class FunctorType, class... Properties>
class ParallelFor<FunctorType, Kokkos::TeamPolicy<Properties...>, Kokkos::B200> {
 public:
  using Policy       = TeamPolicy<Properties...>;
  using functor_type = FunctorType;
  ParallelFor(const FunctorType& arg_functor, const Policy& arg_policy) : m_functor(arg_functor), m_policy(arg_policy) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const {
    B200Adapter::before_launch();
    using W = B200Adapter::TeamForWrap<FunctorType, typename Policy::work_tag>;
    kb200::parallel_for(m_policy.impl_to_kb(), W{m_functor});
  }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
};

template <class CombinedFunctorReducerType, class... Properties>
class ParallelReduce<CombinedFunctorReducerType, Kokkos::TeamPolicy<Properties...>, Kokkos::B200> {
 public:
  using Policy       = TeamPolicy<Properties...>;
  using FunctorType  = typename CombinedFunctorReducerType::functor_type;
  using ReducerType  = typename CombinedFunctorReducerType::reducer_type;
  using pointer_type = typename ReducerType::pointer_type;
  using value_type   = typename ReducerType::value_type;
  using functor_type = FunctorType;
  using reducer_type = ReducerType;

  template <class ViewType>
  ParallelReduce(const CombinedFunctorReducerType& arg_functor_reducer, const Policy& arg_policy, const ViewType& arg_result)
      : m_functor_reducer(arg_functor_reducer),
        m_policy(arg_policy),
        m_result_ptr(arg_result.data()),
        m_result_ptr_device_accessible(MemorySpaceAccess<Kokkos::CudaSpace, typename ViewType::memory_space>::accessible) {}
  Policy const& get_policy() const { return m_policy; }

  void execute() const {
    B200Adapter::before_launch();
    value_type* const host = m_result_ptr_device_accessible ? nullptr : (value_type*)m_result_ptr;
    value_type* const dev  = m_result_ptr_device_accessible ? (value_type*)m_result_ptr : nullptr;
    if constexpr (B200Adapter::is_array_reduction<ReducerType>) {
      using W = B200Adapter::TeamReduceWrap<FunctorType, typename Policy::work_tag, value_type[]>;
      B200Adapter::array_reduce<value_type, void>(m_policy.impl_to_kb(), W{m_functor_reducer.get_functor()},
                                                  (int)m_functor_reducer.get_reducer().value_count(), host, dev);
    } else {
      using R = B200Adapter::Red<ReducerType>;
      using W = B200Adapter::TeamReduceWrap<FunctorType, typename Policy::work_tag, value_type>;
      kb200::Impl::reduce_dispatch(m_policy.impl_to_kb(), W{m_functor_reducer.get_functor()}, R{m_functor_reducer.get_reducer()},
                                   kb200::Impl::ResultTarget<value_type>{host, dev});
    }
  }

 private:
  const CombinedFunctorReducerType m_functor_reducer;
  const Policy m_policy;
  const pointer_type m_result_ptr;
  const bool m_result_ptr_device_accessible;
};

// TeamThreadMDRange / ThreadVectorMDRange / TeamVectorMDRange: the reference's generic nested-loop machinery (impl/Kokkos_TeamMDPolicy.hpp)
// only needs to know on which nest levels to parallelise; as on the reference's GPU backends (Cuda/Kokkos_Cuda_MDRangePolicy.hpp:52-55)
template <typename Rank, TeamMDRangeThreadAndVector ThreadAndVector>
struct ThreadAndVectorNestLevel<Rank, Kokkos::B200, ThreadAndVector> : AcceleratorBasedNestLevel<Rank, ThreadAndVector> {};

}  // namespace Impl

// ---- nested policies and patterns on the B200 team handle: the kernel layer's own, under the names a Kokkos user writes ----
// (these and the pattern forwarders below are declared at the top of Kokkos_B200_Space.hpp, before the reference's headers)
template <class I>
KB200_TEAM_FUNCTION kb200::Impl::TeamThreadRangeStruct<I> TeamThreadRange(const Impl::B200AdapterTeamMember& m, I count) { return kb200::TeamThreadRange(m, count); }
template <class I1, class I2>
KB200_TEAM_FUNCTION kb200::Impl::TeamThreadRangeStruct<std::common_type_t<I1, I2>> TeamThreadRange(const Impl::B200AdapterTeamMember& m, I1 b, I2 e) { return kb200::TeamThreadRange(m, b, e); }
template <class I>
KB200_TEAM_FUNCTION kb200::Impl::TeamVectorRangeStruct<I> TeamVectorRange(const Impl::B200AdapterTeamMember& m, I count) { return kb200::TeamVectorRange(m, count); }
template <class I1, class I2>
KB200_TEAM_FUNCTION kb200::Impl::TeamVectorRangeStruct<std::common_type_t<I1, I2>> TeamVectorRange(const Impl::B200AdapterTeamMember& m, I1 b, I2 e) { return kb200::TeamVectorRange(m, b, e); }
template <class I>
KB200_TEAM_FUNCTION kb200::Impl::ThreadVectorRangeStruct<I> ThreadVectorRange(const Impl::B200AdapterTeamMember& m, I count) { return kb200::ThreadVectorRange(m, count); }
template <class I1, class I2>
KB200_TEAM_FUNCTION kb200::Impl::ThreadVectorRangeStruct<std::common_type_t<I1, I2>> ThreadVectorRange(const Impl::B200AdapterTeamMember& m, I1 b, I2 e) { return kb200::ThreadVectorRange(m, b, e); }
KB200_TEAM_FUNCTION kb200::Impl::ThreadSingleStruct PerTeam(const Impl::B200AdapterTeamMember& m) { return kb200::PerTeam(m); }
KB200_TEAM_FUNCTION kb200::Impl::VectorSingleStruct PerThread(const Impl::B200AdapterTeamMember& m) { return kb200::PerThread(m); }

template <class Range, class L, std::enable_if_t<kb200::Impl::is_nested_range<Range>::value, int>>
KB200_TEAM_FUNCTION void parallel_for(const Range& r, const L& f) { kb200::parallel_for(r, f); }
template <class Range, class L, class... R, std::enable_if_t<kb200::Impl::is_nested_range<Range>::value, int>>
KB200_TEAM_FUNCTION void parallel_reduce(const Range& r, const L& f, R&&... result) { kb200::parallel_reduce(r, f, static_cast<R&&>(result)...); }
template <class Range, class L, class... R, std::enable_if_t<kb200::Impl::is_nested_range<Range>::value, int>>
KB200_TEAM_FUNCTION void parallel_scan(const Range& r, const L& f, R&&... result) { kb200::parallel_scan(r, f, static_cast<R&&>(result)...); }
// Nested MDRange reductions that spread over vector lanes: the reference's generic overloads (Kokkos_ExecPolicy.hpp:1149-1214) fold the
// lanes only for the execution spaces they list by name, so the B200 handle gets its own, more specialised pair that always folds
// them (lanes first, then threads: team_reduce counts a thread's value once, from lane 0)
template <typename Rank, typename Lambda, typename ReducerValueType>
KOKKOS_INLINE_FUNCTION void parallel_reduce(ThreadVectorMDRange<Rank, Impl::B200AdapterTeamMember> const& policy, Lambda const& lambda,
                                            ReducerValueType& val) {
  static_assert(!std::is_array_v<ReducerValueType> && !std::is_pointer_v<ReducerValueType> && !Kokkos::is_reducer_v<ReducerValueType>,
                "Only scalar return types are allowed!");
  val = ReducerValueType{};
  Impl::md_parallel_impl<Rank>(policy, lambda, val);
  kb200::Impl::vector_reduce(kb200::Impl::NestedSum<ReducerValueType>{}, val);
}
template <typename Rank, typename Lambda, typename ReducerValueType>
KOKKOS_INLINE_FUNCTION void parallel_reduce(TeamVectorMDRange<Rank, Impl::B200AdapterTeamMember> const& policy, Lambda const& lambda,
                                            ReducerValueType& val) {
  static_assert(!std::is_array_v<ReducerValueType> && !std::is_pointer_v<ReducerValueType> && !Kokkos::is_reducer_v<ReducerValueType>,
                "Only scalar return types are allowed!");
  val = ReducerValueType{};
  Impl::md_parallel_impl<Rank>(policy, lambda, val);
  kb200::Impl::vector_reduce(kb200::Impl::NestedSum<ReducerValueType>{}, val);
  policy.team.team_reduce(kb200::Impl::NestedSum<ReducerValueType>{}, val);
}
template <class L>
KB200_TEAM_FUNCTION void single(const kb200::Impl::VectorSingleStruct& s, const L& f) { kb200::single(s, f); }
template <class L>
KB200_TEAM_FUNCTION void single(const kb200::Impl::ThreadSingleStruct& s, const L& f) { kb200::single(s, f); }
template <class L, class T>
KB200_TEAM_FUNCTION void single(const kb200::Impl::VectorSingleStruct& s, const L& f, T& val) { kb200::single(s, f, val); }
template <class L, class T>
KB200_TEAM_FUNCTION void single(const kb200::Impl::ThreadSingleStruct& s, const L& f, T& val) { kb200::single(s, f, val); }

}  // namespace Kokkos
#endif
