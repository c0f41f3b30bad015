// kb200/TeamMDRange.hpp -- multi-dimensional nested policies inside a team: TeamThreadMDRange, ThreadVectorMDRange,
// TeamVectorMDRange<Rank<N[, Direction]>, TeamHandle>(team, n0, ..., nN-1) with nested parallel_for / parallel_reduce
// (core/src/Kokkos_ExecPolicy.hpp:1000-1225, core/src/impl/Kokkos_TeamMDPolicy.hpp:30-290; tests: core/unit_test/TestTeamMDRange.hpp).
//
// The reference parallelises ONE chosen nest level over the team's threads (and one over the vector lanes) and runs the
// other levels as sequential loops in every thread.  Here the whole index box is flattened and the flat index is dealt out
// over the participating lanes -- threads (TeamThread), the lanes of one thread (ThreadVector) or every lane of the team
// (TeamVector) -- so short leading extents do not leave lanes idle; Direction picks which index runs fastest (Left: the
// first, the layout of device Views).  As in the reference, a TeamThreadMDRange body runs once per thread in EVERY vector
// lane of that thread, and parallel_reduce sums into a scalar that every participating lane receives.
#ifndef KB200_TEAMMDRANGE_HPP
#define KB200_TEAMMDRANGE_HPP

#include "Team.hpp"

namespace kb200 {
namespace Impl {
enum class TeamMDKind { TeamThread, ThreadVector, TeamVector };

template <class RankT, class TeamHandle, TeamMDKind KIND>
struct TeamMDRangeBase {
  static constexpr int rank = RankT::rank;
  static_assert(rank >= 2 && rank <= 8, "kb200: team MD ranges have 2..8 dimensions");
  // device Views are LayoutLeft: Default iterates with the first index fastest
  static constexpr Iterate direction = RankT::outer_direction == Iterate::Default ? Iterate::Left : RankT::outer_direction;
  using TeamHandleType = TeamHandle;
  using BoundaryType = int;

  template <class... Args>
  KB200_TEAM_FUNCTION TeamMDRangeBase(const TeamHandle& team_, Args&&... args) : team(team_), boundaries{static_cast<BoundaryType>(args)...} {
    static_assert(sizeof...(Args) == (size_t)rank, "kb200: one extent per dimension");
  }
  const TeamHandle& team;
  BoundaryType boundaries[rank];

  // this lane's first flat index and stride
  KB200_TEAM_FUNCTION long long first() const {
    if constexpr (KIND == TeamMDKind::TeamThread) return tm::ty();
    else if constexpr (KIND == TeamMDKind::ThreadVector) return tm::tx();
    else return (long long)tm::ty() * tm::nx() + tm::tx();
  }
  KB200_TEAM_FUNCTION long long step() const {
    if constexpr (KIND == TeamMDKind::TeamThread) return tm::ny();
    else if constexpr (KIND == TeamMDKind::ThreadVector) return tm::nx();
    else return (long long)tm::ny() * tm::nx();
  }
  template <class L, size_t... Is, class... Extra>
  KB200_TEAM_FUNCTION void walk(const L& f, std::index_sequence<Is...>, Extra&... extra) const {
    long long total = 1;
    for (int d = 0; d < rank; ++d) total *= boundaries[d] > 0 ? boundaries[d] : 0;
    const long long s = step();
    for (long long flat = first(); flat < total; flat += s) {
      int idx[rank];
      long long rem = flat;
      if constexpr (direction == Iterate::Left) {
        for (int d = 0; d < rank; ++d) { idx[d] = (int)(rem % boundaries[d]); rem /= boundaries[d]; }
      } else {
        for (int d = rank - 1; d >= 0; --d) { idx[d] = (int)(rem % boundaries[d]); rem /= boundaries[d]; }
      }
      f(idx[Is]..., extra...);
    }
  }
};
}  // namespace Impl

template <class RankT, class TeamHandle>
struct TeamThreadMDRange : Impl::TeamMDRangeBase<RankT, TeamHandle, Impl::TeamMDKind::TeamThread> {
  using Impl::TeamMDRangeBase<RankT, TeamHandle, Impl::TeamMDKind::TeamThread>::TeamMDRangeBase;
};
template <class RankT, class TeamHandle>
struct ThreadVectorMDRange : Impl::TeamMDRangeBase<RankT, TeamHandle, Impl::TeamMDKind::ThreadVector> {
  using Impl::TeamMDRangeBase<RankT, TeamHandle, Impl::TeamMDKind::ThreadVector>::TeamMDRangeBase;
};
template <class RankT, class TeamHandle>
struct TeamVectorMDRange : Impl::TeamMDRangeBase<RankT, TeamHandle, Impl::TeamMDKind::TeamVector> {
  using Impl::TeamMDRangeBase<RankT, TeamHandle, Impl::TeamMDKind::TeamVector>::TeamMDRangeBase;
};
template <class TeamHandle, class... Args>
TeamThreadMDRange(const TeamHandle&, Args&&...) -> TeamThreadMDRange<Rank<sizeof...(Args), Iterate::Default>, TeamHandle>;
template <class TeamHandle, class... Args>
ThreadVectorMDRange(const TeamHandle&, Args&&...) -> ThreadVectorMDRange<Rank<sizeof...(Args), Iterate::Default>, TeamHandle>;
template <class TeamHandle, class... Args>
TeamVectorMDRange(const TeamHandle&, Args&&...) -> TeamVectorMDRange<Rank<sizeof...(Args), Iterate::Default>, TeamHandle>;

// ---- parallel_for
template <class R, class TH, class L>
KB200_TEAM_FUNCTION void parallel_for(const TeamThreadMDRange<R, TH>& p, const L& f) { p.walk(f, std::make_index_sequence<(size_t)R::rank>{}); }
template <class R, class TH, class L>
KB200_TEAM_FUNCTION void parallel_for(const ThreadVectorMDRange<R, TH>& p, const L& f) {
  p.walk(f, std::make_index_sequence<(size_t)R::rank>{});
  Impl::tm::syncwarp();
}
template <class R, class TH, class L>
KB200_TEAM_FUNCTION void parallel_for(const TeamVectorMDRange<R, TH>& p, const L& f) { p.walk(f, std::make_index_sequence<(size_t)R::rank>{}); }

// ---- parallel_reduce: scalar sum; every participating lane receives the result (Kokkos_ExecPolicy.hpp:1130-1215)
template <class R, class TH, class L, class V>
KB200_TEAM_FUNCTION void parallel_reduce(const TeamThreadMDRange<R, TH>& p, const L& f, V& val) {
  static_assert(!std::is_array<V>::value && !std::is_pointer<V>::value && !is_reducer_v<V>, "Only scalar return types are allowed!");
  val = V{};
  p.walk(f, std::make_index_sequence<(size_t)R::rank>{}, val);
  p.team.team_reduce(Impl::NestedSum<V>{}, val);
}
template <class R, class TH, class L, class V>
KB200_TEAM_FUNCTION void parallel_reduce(const ThreadVectorMDRange<R, TH>& p, const L& f, V& val) {
  static_assert(!std::is_array<V>::value && !std::is_pointer<V>::value && !is_reducer_v<V>, "Only scalar return types are allowed!");
  val = V{};
  p.walk(f, std::make_index_sequence<(size_t)R::rank>{}, val);
  Impl::vector_reduce(Impl::NestedSum<V>{}, val);
}
template <class R, class TH, class L, class V>
KB200_TEAM_FUNCTION void parallel_reduce(const TeamVectorMDRange<R, TH>& p, const L& f, V& val) {
  static_assert(!std::is_array<V>::value && !std::is_pointer<V>::value && !is_reducer_v<V>, "Only scalar return types are allowed!");
  val = V{};
  p.walk(f, std::make_index_sequence<(size_t)R::rank>{}, val);
  Impl::vector_reduce(Impl::NestedSum<V>{}, val);
  p.team.team_reduce(Impl::NestedSum<V>{}, val);
}

}  // namespace kb200
#endif
