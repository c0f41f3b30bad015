// Kokkos_B200_UniqueToken.hpp -- Kokkos::Experimental::UniqueToken<Kokkos::B200, Scope> and
// Impl::ParallelFor<F, WorkGraphPolicy<...>, Kokkos::B200>: the two remaining per-backend pieces a kernel running on this space can
// ask the reference for.  Reference shapes: Cuda/Kokkos_Cuda_UniqueToken.hpp:37-146, Cuda/Kokkos_Cuda_WorkGraphPolicy.hpp:27-94.
// The token pool is the kernel layer's own (kb200/UniqueToken.hpp: one lock word per token, exchange-acquire with the successful
// lanes kept inside the loop so a warp never waits on its own members); AcquireUniqueToken / AcquireTeamUniqueToken are the
// reference's generic classes on top of it (Kokkos_UniqueToken.hpp:95-190).
#ifndef KOKKOS_B200_UNIQUE_TOKEN_HPP
#define KOKKOS_B200_UNIQUE_TOKEN_HPP

namespace Kokkos {
namespace Experimental {

template <>
class UniqueToken<Kokkos::B200, UniqueTokenScope::Global> {
 public:
  using execution_space = Kokkos::B200;
  using size_type       = int32_t;

  explicit UniqueToken(execution_space const& space = execution_space()) : m_pool(space.impl_kb200()) {}

  /// upper bound for acquired values, 0 <= value < size()
  KOKKOS_INLINE_FUNCTION size_type size() const noexcept { return (size_type)m_pool.size(); }
  KOKKOS_INLINE_FUNCTION size_type acquire() const { return (size_type)m_pool.acquire(); }
  KOKKOS_INLINE_FUNCTION void release(size_type idx) const noexcept { m_pool.release(idx); }

 private:
  kb200::Experimental::UniqueToken<kb200::B200, kb200::Experimental::UniqueTokenScope::Global> m_pool;
};

template <>
class UniqueToken<Kokkos::B200, UniqueTokenScope::Instance> {
 public:
  using execution_space = Kokkos::B200;
  using size_type       = int32_t;

  UniqueToken() : m_pool(execution_space().impl_kb200()) {}
  explicit UniqueToken(execution_space const& space) : m_pool(space.impl_kb200()) {}
  explicit UniqueToken(size_type max_size) : m_pool((int)max_size, execution_space().impl_kb200()) {}
  UniqueToken(size_type max_size, execution_space const& space) : m_pool((int)max_size, space.impl_kb200()) {}

  KOKKOS_INLINE_FUNCTION size_type size() const noexcept { return (size_type)m_pool.size(); }
  KOKKOS_INLINE_FUNCTION size_type acquire() const { return (size_type)m_pool.acquire(); }
  KOKKOS_INLINE_FUNCTION void release(size_type idx) const noexcept { m_pool.release(idx); }

 private:
  kb200::Experimental::UniqueToken<kb200::B200, kb200::Experimental::UniqueTokenScope::Instance> m_pool;
};

}  // namespace Experimental

namespace Impl {

// WorkGraphPolicy: a dependency graph drained by spinning workers (Kokkos_WorkGraphPolicy.hpp:102-160 pop_work / completed_work).
// Any single resident worker can finish the whole graph, so the launch is an ordinary range kernel over worker slots; every fourth
// slot takes part, which keeps the contention on the queue heads where the reference's measurements put it
// (Cuda/Kokkos_Cuda_WorkGraphPolicy.hpp:56-60), and the slots that find the graph completed return at once.
template <class FunctorType, class... Traits>
class ParallelFor<FunctorType, Kokkos::WorkGraphPolicy<Traits...>, Kokkos::B200> {
 public:
  using Policy = Kokkos::WorkGraphPolicy<Traits...>;

  struct Drain {
    Policy policy;
    FunctorType functor;
    KOKKOS_INLINE_FUNCTION void operator()(const int slot) const {
      if (slot & 3) return;
      for (std::int32_t w = Policy::END_TOKEN; Policy::COMPLETED_TOKEN != (w = policy.pop_work());) {
        if (Policy::END_TOKEN != w) {
          if constexpr (std::is_void_v<typename Policy::work_tag>) functor(w);
          else functor(typename Policy::work_tag{}, w);
          policy.completed_work(w);
        }
      }
    }
  };

  ParallelFor(const FunctorType& arg_functor, const Policy& arg_policy) : m_policy(arg_policy), m_functor(arg_functor) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const {
    B200Adapter::before_launch();
    const Kokkos::B200 space = m_policy.space();
    const int slots = space.concurrency() / 16;  // 128 slots, 32 of them draining, per SM
    kb200::parallel_for(kb200::RangePolicy<kb200::B200>(space.impl_kb200(), 0, slots > 4 ? slots : 4), Drain{m_policy, m_functor});
  }

 private:
  Policy m_policy;
  FunctorType m_functor;
};

}  // namespace Impl
}  // namespace Kokkos
#endif
