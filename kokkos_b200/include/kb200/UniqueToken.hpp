// kb200/UniqueToken.hpp -- Kokkos::Experimental::UniqueToken / AcquireUniqueToken / AcquireTeamUniqueToken for the B200 space
// (core/src/Kokkos_UniqueToken.hpp:30-150, Cuda/Kokkos_Cuda_UniqueToken.hpp:30-140; test: core/unit_test/TestUniqueToken.hpp).
// A token is an index in [0, size()) held by at most one thread at a time: what ScatterView-style codes use to pick a
// private scratch slot.  size() defaults to the space's concurrency (resident threads), so acquire() always terminates.
//
// The lock words are one device array.  A thread starts probing at its own linear id inside the grid: every launcher of
// this backend runs persistent grids no larger than the resident capacity, so under the default size the first probe hits a
// free word and costs one atomic exchange; oversubscribed grids (Schedule<Dynamic>) fall back to probing with the grid's
// thread count as stride, the reference's scheme.
#ifndef KB200_UNIQUETOKEN_HPP
#define KB200_UNIQUETOKEN_HPP

#include "Atomic.hpp"
#include "Team.hpp"
#include "View.hpp"

namespace kb200 {
namespace Experimental {

enum class UniqueTokenScope : int { Instance, Global };

namespace Impl2 {
inline View<unsigned*>& global_token_locks(const B200& space) {
  // intentionally never destroyed: Global tokens may be released from kernels that run until finalize()
  static View<unsigned*>* locks = new View<unsigned*>("kb200::UniqueToken::global_locks", (size_t)space.concurrency());
  return *locks;
}
}  // namespace Impl2

template <class ExecutionSpace = B200, UniqueTokenScope Scope = UniqueTokenScope::Instance>
class UniqueToken {
 public:
  using execution_space = B200;
  using size_type = int;

  explicit UniqueToken(const execution_space& space = execution_space()) {
    if constexpr (Scope == UniqueTokenScope::Global) m_locks = Impl2::global_token_locks(space);
    else m_locks = View<unsigned*>("kb200::UniqueToken::locks", (size_t)space.concurrency());
  }
  // a caller-chosen number of tokens (Instance scope only, as in the reference).  As there, the caller must not let more
  // threads hold or wait for tokens at once than there are tokens: a warp leaves acquire() together, so waiting lanes can
  // keep their warp-mates' tokens from ever being released.
  template <UniqueTokenScope S = Scope, class = std::enable_if_t<S == UniqueTokenScope::Instance>>
  UniqueToken(size_type max_size, const execution_space& = execution_space()) : m_locks("kb200::UniqueToken::locks", (size_t)max_size) {}

  KB200_INLINE_FUNCTION size_type size() const noexcept { return (size_type)m_locks.extent(0); }

  // blocks until a token is free
  KB200_INLINE_FUNCTION size_type acquire() const {
#ifdef __CUDA_ARCH__
    const unsigned n = (unsigned)m_locks.extent(0);
    const unsigned threads = blockDim.x * blockDim.y * blockDim.z;
    const unsigned long long me = (unsigned long long)blockIdx.x * threads + (threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z));
    unsigned idx = (unsigned)(me % n);
    const unsigned stride = threads % n ? threads % n : 1u;
    // keep the successful lanes out of the loop body instead of leaving it, so a warp never waits on its own members
    bool done = false;
    while (!done) {
      if (atomic_exchange(&m_locks(idx), 1u) == 0u) done = true;
      else { idx += stride; if (idx >= n) idx -= n; }
    }
    __threadfence();
    return (size_type)idx;
#else
    return 0;
#endif
  }
  KB200_INLINE_FUNCTION void release(size_type idx) const noexcept {
#ifdef __CUDA_ARCH__
    __threadfence();
    (void)atomic_exchange(&m_locks((size_t)idx), 0u);
#else
    (void)idx;
#endif
  }

 private:
  View<unsigned*> m_locks;
};

// RAII acquire / release (Kokkos_UniqueToken.hpp:95-125)
template <class ExecutionSpace = B200, UniqueTokenScope Scope = UniqueTokenScope::Instance>
class AcquireUniqueToken {
 public:
  using token_type = UniqueToken<ExecutionSpace, Scope>;
  using size_type = typename token_type::size_type;
  KB200_INLINE_FUNCTION explicit AcquireUniqueToken(const token_type& t) : m_token(t), m_value(t.acquire()) {}
  KB200_INLINE_FUNCTION ~AcquireUniqueToken() { m_token.release(m_value); }
  AcquireUniqueToken(const AcquireUniqueToken&) = delete;
  AcquireUniqueToken& operator=(const AcquireUniqueToken&) = delete;
  KB200_INLINE_FUNCTION size_type value() const { return m_value; }

 private:
  const token_type& m_token;
  size_type m_value;
};

// one token per TEAM: rank 0 acquires, the value is broadcast; released after a team barrier (Kokkos_UniqueToken.hpp:127-190)
template <class TeamPolicyT>
class AcquireTeamUniqueToken {
 public:
  using token_type = UniqueToken<B200>;
  using size_type = typename token_type::size_type;
  using team_member_type = typename TeamPolicyT::member_type;
  // the value travels through the member's collective area, no team scratch is taken
  static constexpr size_t shmem_size() { return 0; }
  KB200_INLINE_FUNCTION AcquireTeamUniqueToken(const token_type& t, const team_member_type& team) : m_token(t), m_team(team), m_value(0) {
#ifdef __CUDA_ARCH__
    if (team.team_rank() == 0 && team.impl_vector_lane() == 0) m_value = m_token.acquire();
    team.team_broadcast(m_value, 0);
#endif
  }
  KB200_INLINE_FUNCTION ~AcquireTeamUniqueToken() {
#ifdef __CUDA_ARCH__
    m_team.team_barrier();
    if (m_team.team_rank() == 0 && m_team.impl_vector_lane() == 0) m_token.release(m_value);
#endif
  }
  AcquireTeamUniqueToken(const AcquireTeamUniqueToken&) = delete;
  AcquireTeamUniqueToken& operator=(const AcquireTeamUniqueToken&) = delete;
  KB200_INLINE_FUNCTION size_type value() const { return m_value; }

 private:
  const token_type& m_token;
  const team_member_type& m_team;
  size_type m_value;
};

}  // namespace Experimental
}  // namespace kb200
#endif
