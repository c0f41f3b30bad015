// kb200/Policy.hpp -- execution policies of the hot path, same names and template-property protocol as
// core/src/Kokkos_ExecPolicy.hpp (RangePolicy :76-338, ChunkSize :45-53, TeamPolicy :614-725) and
// core/src/KokkosExp_MDRangePolicy.hpp:169-427 (MDRangePolicy, Rank, Iterate), with the policy traits of
// core/src/impl/Kokkos_AnalyzePolicy.hpp:165-190: execution space, Schedule<Static|Dynamic>, IndexType<T>,
// LaunchBounds<maxT,minB> and a work tag may be given in any order.
//
// B200 specifics:
//   * the default index type is 64-bit (the reference's Cuda default is `unsigned`, which is what made
//     ranges > 2^31 a special case: CHANGELOG.md:70, TestReduce.hpp:650-675);
//   * Schedule and ChunkSize are accepted and ignored, exactly as the reference's Cuda backend does
//     (SURVEY.md section 8a, a23);
//   * bound errors abort like the reference (Kokkos_ExecPolicy.hpp:235-251).
#ifndef KB200_POLICY_HPP
#define KB200_POLICY_HPP

#include "B200.hpp"
#include "Array.hpp"
#include <array>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <tuple>
#include <initializer_list>
#include <type_traits>

namespace kb200 {

struct Static {};
struct Dynamic {};
template <class T> struct Schedule { using type = T; using schedule_type = Schedule; };
template <class T> struct IndexType { using type = T; using index_type = IndexType; };
template <unsigned MaxT = 0, unsigned MinB = 0>
struct LaunchBounds { static constexpr unsigned maxTperB = MaxT, minBperSM = MinB; using launch_bounds = LaunchBounds; };
struct ChunkSize { int value; explicit ChunkSize(int v) : value(v) {} };
struct AUTO_t { constexpr const AUTO_t& operator()() const { return *this; } };
constexpr AUTO_t AUTO{};

enum class Iterate { Default, Left, Right };
template <unsigned N, Iterate OuterDir = Iterate::Default, Iterate InnerDir = Iterate::Default>
struct Rank { static constexpr int rank = (int)N; static constexpr Iterate outer_direction = OuterDir, inner_direction = InnerDir; };

namespace Experimental {
// Occupancy control (core/src/Kokkos_ExecPolicy.hpp:  DesiredOccupancy / MaximizeOccupancy / prefer(); tests:
// core/unit_test/TestOccupancyControlTrait.hpp, TestCommonPolicyConstructors.hpp).  On this backend every launcher sizes a
// persistent grid as SMs x resident blocks, so a desired occupancy of p % simply keeps ceil(resident * p / 100) blocks per SM
// (the reference's Cuda backend pads the launch with dynamic shared memory to the same end, Cuda_KernelLaunch.hpp:160-210).
struct DesiredOccupancy {
  int m_occ = 100;
  DesiredOccupancy() = default;
  explicit constexpr DesiredOccupancy(int occ) : m_occ(occ < 0 ? 0 : (occ > 100 ? 100 : occ)) {}
  explicit constexpr operator int() const { return m_occ; }
  constexpr int value() const { return m_occ; }
};
struct MaximizeOccupancy { explicit MaximizeOccupancy() = default; };
}  // namespace Experimental

struct ParallelForTag {};     // core/src/Kokkos_Core_fwd.hpp: pattern tags for team_size_max / team_size_recommended
struct ParallelReduceTag {};
struct ParallelScanTag {};

namespace Impl {
template <class T> struct is_schedule : std::false_type {};
template <class T> struct is_schedule<Schedule<T>> : std::true_type {};
template <class T> struct is_index_type : std::false_type {};
template <class T> struct is_index_type<IndexType<T>> : std::true_type {};
template <class T> struct is_launch_bounds : std::false_type {};
template <unsigned A, unsigned B> struct is_launch_bounds<LaunchBounds<A, B>> : std::true_type {};
template <class T> struct is_rank : std::false_type {};
template <unsigned N, Iterate A, Iterate B> struct is_rank<Rank<N, A, B>> : std::true_type {};
template <class T> struct is_exec_space : std::is_same<T, B200> {};
template <class T> struct is_occupancy : std::integral_constant<bool, std::is_same<T, Experimental::DesiredOccupancy>::value || std::is_same<T, Experimental::MaximizeOccupancy>::value> {};

template <class... P> struct policy_traits;
template <> struct policy_traits<> {
  using schedule = Schedule<Static>; using index = void; using bounds = LaunchBounds<>; using tag = void; using rank = void;
  static constexpr bool desired_occupancy = false;
};
template <class F, class... R>
struct policy_traits<F, R...> {
  using next = policy_traits<R...>;
  // a bare integral type is an index type too (RangePolicy<Space, long>: impl/Kokkos_AnalyzePolicy.hpp:165-190)
  static constexpr bool is_index = is_index_type<F>::value || std::is_integral<F>::value;
  static constexpr bool known = is_schedule<F>::value || is_index || is_launch_bounds<F>::value ||
                                is_rank<F>::value || is_exec_space<F>::value || is_occupancy<F>::value;
  static constexpr bool desired_occupancy = std::is_same<F, Experimental::DesiredOccupancy>::value || next::desired_occupancy;
  using schedule = std::conditional_t<is_schedule<F>::value, F, typename next::schedule>;
  using index = std::conditional_t<is_index, std::conditional_t<std::is_integral<F>::value, IndexType<F>, F>, typename next::index>;
  using bounds = std::conditional_t<is_launch_bounds<F>::value, F, typename next::bounds>;
  using rank = std::conditional_t<is_rank<F>::value, F, typename next::rank>;
  using tag = std::conditional_t<!known, F, typename next::tag>;  // anything unrecognised is the work tag
};
template <class IT> struct index_of { using type = typename IT::type; };
template <> struct index_of<void> { using type = long long; };

// defined in Team.hpp, after the team kernels
template <class Policy, class F, class PatternTag>
int team_size_limit(const Policy& pol, const F& f, const PatternTag&);

// true when converting `bound` to Index does not preserve its value (sign change or narrowing):
// core/src/Kokkos_ExecPolicy.hpp:258-290, core/src/KokkosExp_MDRangePolicy.hpp:60-95
template <class Index, class T>
constexpr bool index_conversion_unsafe(const T bound) {
  if constexpr (std::is_convertible<Index, T>::value) {
    bool unsafe = false;
    if constexpr (std::is_arithmetic<T>::value && std::is_signed<T>::value != std::is_signed<Index>::value) {
      if constexpr (std::is_signed<T>::value) unsafe = unsafe || bound < static_cast<T>(std::numeric_limits<Index>::min());
      if constexpr (std::is_signed<Index>::value) unsafe = unsafe || bound > static_cast<T>(std::numeric_limits<Index>::max());
    }
    return unsafe || static_cast<T>(static_cast<Index>(bound)) != bound;
  } else {
    return false;
  }
}

// defaults of the MDRange tiling on this device (the role of Impl::get_tile_size_properties, KokkosExp_MDRangePolicy.hpp:98-128
// and Cuda/Kokkos_Cuda_MDRangePolicy.hpp:25-35): 32 threads along the contiguous dimension (one warp = one 256-byte row),
// `default_tile_size` along every other dimension while the tile stays under max_total_tile_size
struct TileSizeProperties {
  int max_threads;
  int default_largest_tile_size;
  int default_tile_size;
  int max_total_tile_size;
  int max_threads_dimensions[3];
};
inline int md_default_tile_size() {
  static const int v = [] {
    const char* e = std::getenv("KB200_MD_DEFAULT_TILE");
    const int t = e ? std::atoi(e) : 0;
    return t > 0 && t <= 32 ? t : 4;
  }();
  return v;
}
template <class Space>
TileSizeProperties get_tile_size_properties(const Space&) {
  return TileSizeProperties{1024, 32, md_default_tile_size(), 1024, {1024, 1024, 64}};
}

[[noreturn]] inline void policy_abort(const char* msg) {
  std::fprintf(stderr, "%s", msg);
  if (msg[0] && msg[std::strlen(msg) - 1] != '\n') std::fprintf(stderr, "\n");
  std::abort();
}
// Common base of the execution policies (the role of Impl::PolicyTraits, core/src/impl/Kokkos_AnalyzePolicy.hpp:165-190):
// publishes the analysed traits and stores the desired occupancy only in policy types that carry the trait (empty otherwise).
template <bool Has> struct OccupancyStorage {};
template <> struct OccupancyStorage<true> { Experimental::DesiredOccupancy m_desired_occupancy; };
template <class... Props>
struct PolicyTraits : OccupancyStorage<policy_traits<Props...>::desired_occupancy> {
  using analysed = policy_traits<Props...>;
  using execution_space = B200;
  using schedule_type = typename analysed::schedule;
  using work_tag = typename analysed::tag;
  using launch_bounds = typename analysed::bounds;
  using index_type = typename index_of<typename analysed::index>::type;
  static constexpr bool experimental_contains_desired_occupancy = analysed::desired_occupancy;
  PolicyTraits() = default;
  PolicyTraits(const PolicyTraits&) = default;
  PolicyTraits& operator=(const PolicyTraits&) = default;
  // from a policy with other traits: the occupancy travels when both sides store one
  template <class... Other>
  PolicyTraits(const PolicyTraits<Other...>& o) {
    if constexpr (experimental_contains_desired_occupancy && PolicyTraits<Other...>::experimental_contains_desired_occupancy)
      this->m_desired_occupancy = o.m_desired_occupancy;
    (void)o;
  }
  template <bool B = experimental_contains_desired_occupancy, class = std::enable_if_t<B>>
  Experimental::DesiredOccupancy impl_get_desired_occupancy() const { return this->m_desired_occupancy; }
  template <bool B = experimental_contains_desired_occupancy, class = std::enable_if_t<B>>
  void impl_set_desired_occupancy(Experimental::DesiredOccupancy occ) { this->m_desired_occupancy = occ; }
  // resident blocks per SM a launcher should use, given what the kernel allows
  int impl_occupancy_cap(int resident_blocks_per_sm) const {
    if constexpr (experimental_contains_desired_occupancy) {
      const int want = (resident_blocks_per_sm * this->m_desired_occupancy.value() + 99) / 100;
      return want < 1 ? 1 : (want < resident_blocks_per_sm ? want : resident_blocks_per_sm);
    } else {
      return resident_blocks_per_sm;
    }
  }
};
}  // namespace Impl

namespace Experimental {
namespace Impl2 {
template <template <class...> class Policy, class Done, class... Rest> struct without_occupancy;
template <template <class...> class Policy, class... Done>
struct without_occupancy<Policy, std::tuple<Done...>> { using type = Policy<Done...>; };
template <template <class...> class Policy, class... Done, class F, class... Rest>
struct without_occupancy<Policy, std::tuple<Done...>, F, Rest...>
    : std::conditional_t<kb200::Impl::is_occupancy<F>::value, without_occupancy<Policy, std::tuple<Done...>, Rest...>,
                         without_occupancy<Policy, std::tuple<Done..., F>, Rest...>> {};
}  // namespace Impl2
// prefer(policy, DesiredOccupancy{p}) / prefer(policy, MaximizeOccupancy{}): a policy of the matching type with the hint applied
template <template <class...> class Policy, class... Args>
auto prefer(const Policy<Args...>& p, DesiredOccupancy occ) {
  if constexpr (Policy<Args...>::experimental_contains_desired_occupancy) {
    Policy<Args...> q(p);
    q.impl_set_desired_occupancy(occ);
    return q;
  } else {
    Policy<Args..., DesiredOccupancy> q{p};
    q.impl_set_desired_occupancy(occ);
    return q;
  }
}
template <template <class...> class Policy, class... Args>
auto prefer(const Policy<Args...>& p, MaximizeOccupancy) {
  if constexpr (Policy<Args...>::experimental_contains_desired_occupancy) {
    typename Impl2::without_occupancy<Policy, std::tuple<>, Args...>::type q{p};
    return q;
  } else {
    return p;
  }
}
}  // namespace Experimental

// ------------------------------------------------------------------------------------------ RangePolicy
template <class... Props>
class RangePolicy : public Impl::PolicyTraits<Props...> {
  using traits = Impl::policy_traits<Props...>;
  template <class...> friend class RangePolicy;

 public:
  using execution_space = B200;
  using execution_policy = RangePolicy;
  using work_tag = typename traits::tag;
  using schedule_type = typename traits::schedule;
  using launch_bounds = typename traits::bounds;
  using index_type = typename Impl::index_of<typename traits::index>::type;
  using member_type = index_type;

  RangePolicy() : m_begin(0), m_end(0) {}
  // same range under other traits (what Experimental::prefer / require produce)
  template <class... Other, class = std::enable_if_t<!std::is_same<RangePolicy<Other...>, RangePolicy>::value>>
  RangePolicy(const RangePolicy<Other...>& o)
      : Impl::PolicyTraits<Props...>(static_cast<const Impl::PolicyTraits<Other...>&>(o)), m_space(o.m_space), m_begin((index_type)o.m_begin), m_end((index_type)o.m_end), m_chunk((index_type)o.m_chunk) {}
  // bounds of any integral type: each is checked for a value-preserving conversion to index_type before use
  // (core/src/Kokkos_ExecPolicy.hpp:235-290; core/unit_test/TestRangePolicyConstructors.hpp pins the diagnostics)
  template <class B, class E, class = std::enable_if_t<std::is_convertible<B, index_type>::value && std::is_convertible<E, index_type>::value &&
                                                       !std::is_same<std::decay_t<B>, B200>::value>>
  RangePolicy(const B b, const E e) : m_begin(checked(b)), m_end(checked(e)) { check(); }
  template <class B, class E, class = std::enable_if_t<std::is_convertible<B, index_type>::value && std::is_convertible<E, index_type>::value>>
  RangePolicy(const B200& s, const B b, const E e) : m_space(s), m_begin(checked(b)), m_end(checked(e)) { check(); }
  template <class B, class E, class = std::enable_if_t<std::is_convertible<B, index_type>::value && std::is_convertible<E, index_type>::value &&
                                                       !std::is_same<std::decay_t<B>, B200>::value>>
  RangePolicy(const B b, const E e, ChunkSize c) : m_begin(checked(b)), m_end(checked(e)), m_chunk(c.value) { check(); }
  template <class B, class E, class = std::enable_if_t<std::is_convertible<B, index_type>::value && std::is_convertible<E, index_type>::value>>
  RangePolicy(const B200& s, const B b, const E e, ChunkSize c) : m_space(s), m_begin(checked(b)), m_end(checked(e)), m_chunk(c.value) { check(); }

  const B200& space() const { return m_space; }
  KB200_INLINE_FUNCTION index_type begin() const { return m_begin; }
  KB200_INLINE_FUNCTION index_type end() const { return m_end; }
  index_type chunk_size() const { return m_chunk; }
  RangePolicy& set_chunk_size(int c) { m_chunk = c; return *this; }

 private:
  void check() {
    if (m_end < m_begin) {
      const std::string msg = std::string(KB200_NS_STR "::RangePolicy bounds error: The lower bound (") + std::to_string(m_begin) +
                              ") is greater than the upper bound (" + std::to_string(m_end) + ").\n";
      Impl::policy_abort(msg.c_str());
    }
  }
  template <class T>
  static index_type checked(const T bound) {
    if (Impl::index_conversion_unsafe<index_type>(bound)) {
      const std::string msg = std::string(KB200_NS_STR "::RangePolicy bound type error: an unsafe implicit conversion is performed on a bound (") +
                              std::to_string(bound) + "), which may not preserve its original value.\n";
      Impl::policy_abort(msg.c_str());
    }
    return static_cast<index_type>(bound);
  }
  B200 m_space;
  index_type m_begin, m_end;
  index_type m_chunk = 0;
};

// ------------------------------------------------------------------------------------------ MDRangePolicy
template <class... Props>
class MDRangePolicy : public Impl::PolicyTraits<Props...> {
  using traits = Impl::policy_traits<Props...>;
  template <class...> friend class MDRangePolicy;
  static_assert(!std::is_void<typename traits::rank>::value, "kb200::MDRangePolicy needs a Rank<N> property");

 public:
  static constexpr int rank = traits::rank::rank;
  static_assert(rank >= 2 && rank <= 6, "MDRangePolicy rank must be 2..6");
  using execution_space = B200;
  using execution_policy = MDRangePolicy;
  using work_tag = typename traits::tag;
  using launch_bounds = typename traits::bounds;
  using index_type = typename Impl::index_of<typename traits::index>::type;
  using point_type = Array<index_type, rank>;
  using tile_type = Array<index_type, rank>;
  // device iteration is Left/Left like the reference's Cuda backend (Cuda/Kokkos_Cuda_MDRangePolicy.hpp:25-35):
  // dimension 0 is the fastest-varying one and maps to threadIdx.x
  static constexpr Iterate outer_direction = Iterate::Left, inner_direction = Iterate::Left;

  MDRangePolicy() = default;
  template <class... Other, class = std::enable_if_t<!std::is_same<MDRangePolicy<Other...>, MDRangePolicy>::value && MDRangePolicy<Other...>::rank == rank>>
  MDRangePolicy(const MDRangePolicy<Other...>& o) : Impl::PolicyTraits<Props...>(static_cast<const Impl::PolicyTraits<Other...>&>(o)), m_space(o.m_space) {
    for (int d = 0; d < rank; ++d) {
      m_lower[d] = (index_type)o.m_lower[d]; m_upper[d] = (index_type)o.m_upper[d];
      m_tile[d] = (index_type)o.m_tile[d]; m_tile_end[d] = (index_type)o.m_tile_end[d];
    }
    m_num_tiles = (index_type)o.m_num_tiles; m_prod_tile_dims = (index_type)o.m_prod_tile_dims; m_tune_tile_size = o.m_tune_tile_size;
  }
  // braced lists of one element type: {0, 0, 0}, {n0, n1, n2} [, {t0, t1, t2}]
  template <class L, class U>
  MDRangePolicy(std::initializer_list<L> lower, std::initializer_list<U> upper) { init(lower, upper, std::initializer_list<index_type>{}); }
  template <class L, class U, class Tl>
  MDRangePolicy(std::initializer_list<L> lower, std::initializer_list<U> upper, std::initializer_list<Tl> tile) { init(lower, upper, tile); }
  template <class L, class U>
  MDRangePolicy(const B200& s, std::initializer_list<L> lower, std::initializer_list<U> upper) : m_space(s) { init(lower, upper, std::initializer_list<index_type>{}); }
  template <class L, class U, class Tl>
  MDRangePolicy(const B200& s, std::initializer_list<L> lower, std::initializer_list<U> upper, std::initializer_list<Tl> tile) : m_space(s) { init(lower, upper, tile); }
  // point_type / tile_type (also what a double-braced {{0, 1}} or a list of mixed integer types binds to)
  MDRangePolicy(const point_type& lower, const point_type& upper, const tile_type& tile = tile_type{}) { init_arrays(lower, upper, tile); }
  // (template on the space type: a braced list such as {{0, 0, 0}} must never be considered for the execution-space parameter)
  template <class S, class = std::enable_if_t<std::is_same<std::decay_t<S>, B200>::value>>
  MDRangePolicy(const S& s, const point_type& lower, const point_type& upper, const tile_type& tile = tile_type{}) : m_space(s) { init_arrays(lower, upper, tile); }
  // Array of any integer type; the tile may name fewer dimensions than the rank (KokkosExp_MDRangePolicy.hpp:262-310,
  // core/unit_test/TestMDRangePolicyConstructors.hpp:38-77)
  template <class LT, size_t LN, class UT, size_t UN, class TT = index_type, size_t TN = (size_t)rank,
            class = std::enable_if_t<!(std::is_same<LT, index_type>::value && std::is_same<UT, index_type>::value && std::is_same<TT, index_type>::value &&
                                       TN == (size_t)rank)>>
  MDRangePolicy(const Array<LT, LN>& lower, const Array<UT, UN>& upper, const Array<TT, TN>& tile = Array<TT, TN>{}) {
    init_arrays(to_point(lower, "lower"), to_point(upper, "upper"), to_tile(tile));
  }
  template <class S, class LT, size_t LN, class UT, size_t UN, class TT = index_type, size_t TN = (size_t)rank,
            class = std::enable_if_t<std::is_same<std::decay_t<S>, B200>::value &&
                                     !(std::is_same<LT, index_type>::value && std::is_same<UT, index_type>::value && std::is_same<TT, index_type>::value &&
                                       TN == (size_t)rank)>>
  MDRangePolicy(const S& s, const Array<LT, LN>& lower, const Array<UT, UN>& upper, const Array<TT, TN>& tile = Array<TT, TN>{}) : m_space(s) {
    init_arrays(to_point(lower, "lower"), to_point(upper, "upper"), to_tile(tile));
  }

  const B200& space() const { return m_space; }
  point_type m_lower{}, m_upper{};
  tile_type m_tile{}, m_tile_end{};
  index_type m_num_tiles = 0, m_prod_tile_dims = 1;
  bool m_tune_tile_size = false;
  static constexpr int max_tile_product = 1024;  // one tile = one thread block
  int max_total_tile_size() const { return Impl::get_tile_size_properties(m_space).max_total_tile_size; }
  bool impl_tune_tile_size() const { return m_tune_tile_size; }
  // The launcher calls this when a kernel cannot run with the DEFAULT tile (a register-heavy functor: 512 threads leave 128
  // registers each): halve the slowest dimension that is still wider than 1 and recount the tiles.  False when the tile was
  // chosen by the caller or nothing is left to shrink.
  bool impl_shrink_default_tile() {
    if (!m_tune_tile_size) return false;
    int d = rank - 1;
    while (d >= 0 && m_tile[d] <= 1) --d;
    if (d < 0) return false;
    m_tile[d] = (m_tile[d] + 1) / 2;
    m_num_tiles = 1; m_prod_tile_dims = 1;
    for (int r = 0; r < rank; ++r) {
      const index_type len = m_upper[r] - m_lower[r];
      m_tile_end[r] = (len + m_tile[r] - 1) / m_tile[r];
      m_num_tiles *= m_tile_end[r];
      m_prod_tile_dims *= m_tile[r];
    }
    return true;
  }
  // the tile the policy would pick on its own for these extents (KokkosExp_MDRangePolicy.hpp:335-352)
  tile_type tile_size_recommended() const {
    const Impl::TileSizeProperties pr = Impl::get_tile_size_properties(m_space);
    tile_type rec{};
    for (int d = 0; d < rank; ++d) rec[d] = d == 0 ? (index_type)pr.default_largest_tile_size : (index_type)pr.default_tile_size;
    return rec;
  }

 private:
  template <class T>
  static index_type checked(const T bound, int dim) {
    if (Impl::index_conversion_unsafe<index_type>(bound)) {
      const std::string msg = std::string(KB200_NS_STR "::MDRangePolicy bound type error: an unsafe implicit conversion is performed on a bound (") +
                              std::to_string(bound) + ") in dimension (" + std::to_string(dim) + "), which may not preserve its original value.\n";
      Impl::policy_abort(msg.c_str());
    }
    return static_cast<index_type>(bound);
  }
  template <class T, size_t N>
  static point_type to_point(const Array<T, N>& a, const char*) {
    static_assert(N == (size_t)rank, "kb200::MDRangePolicy: bound arrays must have one entry per dimension");
    point_type p{};
    for (int d = 0; d < rank; ++d) p[d] = checked(a[d], d);
    return p;
  }
  template <class T, size_t N>
  static tile_type to_tile(const Array<T, N>& a) {
    static_assert(N <= (size_t)rank, "kb200::MDRangePolicy: the tile array has more entries than the policy has dimensions");
    tile_type t{};
    if constexpr (N > 0)
      for (size_t d = 0; d < N; ++d) t[d] = checked(a[d], (int)d);
    return t;
  }
  template <class L, class U, class Tl>
  void init(std::initializer_list<L> lo, std::initializer_list<U> up, std::initializer_list<Tl> tl) {
    if ((int)lo.size() != rank || (int)up.size() != rank || ((int)tl.size() != rank && tl.size() != 0))
      Impl::policy_abort(KB200_NS_STR "::MDRangePolicy: Constructor initializer lists have wrong size");
    point_type l{}, u{};
    tile_type t{};
    int k = 0; for (auto v : lo) { l[k] = checked(v, k); ++k; }
    k = 0; for (auto v : up) { u[k] = checked(v, k); ++k; }
    k = 0; for (auto v : tl) { t[k] = checked(v, k); ++k; }
    init_arrays(l, u, t);
  }
  void init_arrays(const point_type& l, const point_type& u, const tile_type& t) {
    m_lower = l; m_upper = u; m_tile = t;
    const Impl::TileSizeProperties pr = Impl::get_tile_size_properties(m_space);
    m_num_tiles = 1; m_prod_tile_dims = 1;
    for (int d = 0; d < rank; ++d) {
      if (m_upper[d] < m_lower[d]) {
        const std::string msg = std::string(KB200_NS_STR "::MDRangePolicy bounds error: The lower bound (") + std::to_string(m_lower[d]) +
                                ") is greater than its upper bound (" + std::to_string(m_upper[d]) + ") in dimension " + std::to_string(d) + ".\n";
        Impl::policy_abort(msg.c_str());
      }
      const index_type len = m_upper[d] - m_lower[d];
      if (m_tile[d] <= 0) {  // default: see TileSizeProperties
        m_tune_tile_size = true;
        if (d == 0) m_tile[d] = (index_type)pr.default_largest_tile_size;
        else m_tile[d] = (long long)m_prod_tile_dims * pr.default_tile_size < (long long)pr.max_total_tile_size ? (index_type)pr.default_tile_size : (index_type)1;
      }
      if (m_tile[d] > len && len > 0) m_tile[d] = len;  // no idle threads along a short dimension
      if (m_tile[d] < 1) m_tile[d] = 1;
      m_tile_end[d] = (len + m_tile[d] - 1) / m_tile[d];
      m_num_tiles *= m_tile_end[d];
      m_prod_tile_dims *= m_tile[d];
    }
    if (m_prod_tile_dims > (index_type)max_tile_product) Impl::policy_abort(KB200_NS_STR "::MDRangePolicy: tile dimensions exceed the maximum of 1024 threads per tile");
  }
  B200 m_space;
};

// ------------------------------------------------------------------------------------------ TeamPolicy
struct PerTeamValue { size_t value; };
struct PerThreadValue { size_t value; };
inline PerTeamValue PerTeam(size_t v) { return PerTeamValue{v}; }
inline PerThreadValue PerThread(size_t v) { return PerThreadValue{v}; }

class B200TeamMember;

template <class... Props>
class TeamPolicy : public Impl::PolicyTraits<Props...> {
  using traits = Impl::policy_traits<Props...>;
  template <class...> friend class TeamPolicy;

 public:
  using execution_space = B200;
  using execution_policy = TeamPolicy;
  using work_tag = typename traits::tag;
  using launch_bounds = typename traits::bounds;
  using index_type = typename Impl::index_of<typename traits::index>::type;
  using member_type = B200TeamMember;

  TeamPolicy() {}
  template <class... Other, class = std::enable_if_t<!std::is_same<TeamPolicy<Other...>, TeamPolicy>::value>>
  TeamPolicy(const TeamPolicy<Other...>& o)
      : Impl::PolicyTraits<Props...>(static_cast<const Impl::PolicyTraits<Other...>&>(o)), m_space(o.m_space), m_league(o.m_league), m_team(o.m_team), m_vec(o.m_vec), m_chunk(o.m_chunk) {
    for (int l = 0; l < 2; ++l) { m_team_scratch[l] = o.m_team_scratch[l]; m_thread_scratch[l] = o.m_thread_scratch[l]; }
  }
  TeamPolicy(int league, int team, int vec = 1) : m_league(league), m_team(team), m_vec(vec) { check(); }
  TeamPolicy(int league, const AUTO_t&, int vec = 1) : m_league(league), m_team(-1), m_vec(vec) { check(); }
  TeamPolicy(int league, const AUTO_t&, const AUTO_t&) : m_league(league), m_team(-1), m_vec(-1) { check(); }
  TeamPolicy(int league, int team, const AUTO_t&) : m_league(league), m_team(team), m_vec(-1) { check(); }
  TeamPolicy(const B200& s, int league, int team, int vec = 1) : m_space(s), m_league(league), m_team(team), m_vec(vec) { check(); }
  TeamPolicy(const B200& s, int league, const AUTO_t&, int vec = 1) : m_space(s), m_league(league), m_team(-1), m_vec(vec) { check(); }

  const B200& space() const { return m_space; }
  int league_size() const { return m_league; }
  int team_size() const { return m_team; }              // -1 = AUTO (resolved at launch)
  int impl_vector_length() const { return m_vec; }      // -1 = AUTO
  bool impl_auto_team_size() const { return m_team < 0; }
  bool impl_auto_vector_length() const { return m_vec < 0; }
  static int vector_length_max() { return 32; }         // Cuda_Parallel_Team.hpp:174
  static int scratch_size_max(int level) { return level == 0 ? 200 * 1024 : (1 << 30); }
  size_t scratch_size(int level, int team_size = -1) const {
    const int ts = team_size > 0 ? team_size : (m_team > 0 ? m_team : 1);
    return m_team_scratch[level] + m_thread_scratch[level] * (size_t)ts;
  }
  int chunk_size() const { return m_chunk; }   // accepted and ignored by the launch (league members are strided over a persistent grid)
  TeamPolicy& set_chunk_size(int c) { m_chunk = c; return *this; }
  size_t team_scratch_size(int level) const { return m_team_scratch[level]; }
  size_t thread_scratch_size(int level) const { return m_thread_scratch[level]; }
  TeamPolicy& set_scratch_size(int level, PerTeamValue t) { chk_level(level); m_team_scratch[level] = t.value; return *this; }
  TeamPolicy& set_scratch_size(int level, PerThreadValue t) { chk_level(level); m_thread_scratch[level] = t.value; return *this; }
  TeamPolicy& set_scratch_size(int level, PerTeamValue a, PerThreadValue b) { chk_level(level); m_team_scratch[level] = a.value; m_thread_scratch[level] = b.value; return *this; }
  TeamPolicy& set_scratch_size(int level, PerThreadValue b, PerTeamValue a) { return set_scratch_size(level, a, b); }
  // team_size_max / team_size_recommended (Cuda_Parallel_Team.hpp:96-172,346-389): derived from the attributes of the kernel
  // that would actually be launched for this functor (registers, launch bounds) and from the level-0 scratch request
  template <class F, class PatternTag> int team_size_max(const F& f, const PatternTag& t) const { return Impl::team_size_limit(*this, f, t); }
  template <class F, class PatternTag> int team_size_recommended(const F& f, const PatternTag& t) const {
    const int mx = Impl::team_size_limit(*this, f, t), dflt = impl_default_team_size();
    return dflt < mx ? dflt : mx;
  }
  // reducer-taking forms (Kokkos_ExecPolicy.hpp:365-508)
  template <class F, class R> int team_size_max(const F& f, const R&, const ParallelReduceTag& t) const { return Impl::team_size_limit(*this, f, t); }
  template <class F, class R> int team_size_recommended(const F& f, const R&, const ParallelReduceTag& t) const { return team_size_recommended(f, t); }
  int impl_default_team_size() const { const int v = m_vec > 0 ? m_vec : 1; return 256 / v > 0 ? 256 / v : 1; }

 private:
  void chk_level(int level) const { if (level < 0 || level > 1) Impl::policy_abort("kb200::TeamPolicy: scratch level must be 0 or 1"); }
  void check() {
    if (m_league < 0) Impl::policy_abort("kb200::TeamPolicy: negative league size");
    // as the reference's Cuda backend: clamp the request to 32 and round it DOWN to a power of two
    // (Cuda/Kokkos_Cuda_Parallel_Team.hpp verify_requested_vector_length; TestTeamVector.hpp:1037 asks for 33 and 19)
    if (m_vec > 32) m_vec = 32;
    if (m_vec > 0) { int p2 = 1; while (p2 * 2 <= m_vec) p2 *= 2; m_vec = p2; }
    if (m_team > 0 && m_vec > 0 && m_team * m_vec > 1024) throw std::runtime_error("kb200::TeamPolicy: requested team_size * vector_length exceeds 1024 threads");
  }
  B200 m_space;
  int m_league = 0, m_team = -1, m_vec = 1;
  int m_chunk = 32;  // Cuda/Kokkos_Cuda_Parallel_Team.hpp: default chunk = warp size
  size_t m_team_scratch[2] = {0, 0}, m_thread_scratch[2] = {0, 0};
};

}  // namespace kb200
#endif
