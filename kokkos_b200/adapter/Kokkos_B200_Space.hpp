// Kokkos_B200_Space.hpp -- `Kokkos::B200`: the B200-native execution space attached to an UNMODIFIED Kokkos (kokkos/kokkos
// 4.6.99) the way the reference's own backends attach: include this header after <Kokkos_Core.hpp>
// (it does so itself) in a CUDA-enabled build of the reference and name the space explicitly,
//     Kokkos::parallel_reduce(Kokkos::RangePolicy<Kokkos::B200>(0, n), KOKKOS_LAMBDA(int i, double& u) { u += a(i); }, sum);
// Nothing in the reference tree is edited; `Kokkos::Cuda` and `Kokkos::OpenMP` stay usable in the same binary (comparator
// and oracle).  What is provided, with the reference shape each piece stands in for:
//   class Kokkos::B200                               execution-space concept, model core/src/Cuda/Kokkos_Cuda.hpp:95-247
//                                                    (checked by impl/Kokkos_ExecSpaceManager.hpp:96-108)
//   Impl::ParallelFor / ParallelReduce / ParallelScan / ParallelScanWithTotal <..., RangePolicy<...>, B200>
//                                                    Cuda/Kokkos_Cuda_Parallel_Range.hpp:38-116,118-388,390-701,704-1047
//   Impl::ParallelFor / ParallelReduce <..., MDRangePolicy<...>, B200>   Cuda/Kokkos_Cuda_Parallel_MDRange.hpp:52-246,248-497
//   default_{outer,inner}_direction<B200>, Impl::get_tile_size_properties<B200>   Cuda/Kokkos_Cuda_MDRangePolicy.hpp:25-48
//   Impl::ZeroMemset<B200>                           Cuda/Kokkos_Cuda_ZeroMemset.hpp:26-33
//   Tools::Experimental::DeviceTypeTraits<B200>      Cuda/Kokkos_Cuda.hpp:249-258
//   registration with the space factory              Cuda/Kokkos_Cuda_Instance.cpp:746-747
// The kernels are the hand-written sm_100a kernels of kokkos_b200/include/kb200/impl (persistent 256-bit tile skeletons,
// shuffle/redux + ordered ticket combine, single-pass look-back scan): each specialisation only translates the reference's
// policy and functor/reducer into the kernel-side "Body"/"Red" concepts and launches through libkokkos_b200.so.
// Memory: Kokkos::CudaSpace (the reference's), so every View type of a Cuda build works unchanged.
// TeamPolicy and the nested Team*MDRange policies: Kokkos_B200_Team.hpp.
#ifndef KOKKOS_B200_SPACE_HPP
#define KOKKOS_B200_SPACE_HPP

#ifdef KB200_AS_KOKKOS
#error "the adapter uses the kb200:: layer under its own name; do not define KB200_AS_KOKKOS here"
#endif
// The kernel layer first (it is self-contained), so that the forwarding overloads of the nested patterns can be DECLARED before the
// reference's headers are parsed: impl/Kokkos_TeamMDPolicy.hpp calls parallel_for / parallel_reduce unqualified from inside
// namespace Kokkos::Impl on whatever TeamThreadRange(team, n) returns, and only names already declared at that point (or found
// by argument-dependent lookup, which does not reach namespace Kokkos for kb200 types) are considered.
#include <Kokkos_B200.hpp>
namespace Kokkos {
template <class Range, class L, std::enable_if_t<kb200::Impl::is_nested_range<Range>::value, int> = 0>
KB200_TEAM_FUNCTION void parallel_for(const Range& r, const L& f);
template <class Range, class L, class... R, std::enable_if_t<kb200::Impl::is_nested_range<Range>::value, int> = 0>
KB200_TEAM_FUNCTION void parallel_reduce(const Range& r, const L& f, R&&... result);
template <class Range, class L, class... R, std::enable_if_t<kb200::Impl::is_nested_range<Range>::value, int> = 0>
KB200_TEAM_FUNCTION void parallel_scan(const Range& r, const L& f, R&&... result);
// The nested-policy factories and `single`, for the same reason: Kokkos_CopyViews.hpp (local_deep_copy) and
// Kokkos_AcquireUniqueTokenImpl.hpp call them QUALIFIED (Kokkos::TeamVectorRange(team, n), Kokkos::single(Kokkos::PerTeam(team), f)),
// which binds to the overloads declared at that point of the reference's headers.  Defined in Kokkos_B200_Team.hpp.
namespace Impl {
class B200AdapterTeamMember;
}
template <class I>
KB200_TEAM_FUNCTION kb200::Impl::TeamThreadRangeStruct<I> TeamThreadRange(const Impl::B200AdapterTeamMember& m, I count);
template <class I1, class I2>
KB200_TEAM_FUNCTION kb200::Impl::TeamThreadRangeStruct<std::common_type_t<I1, I2>> TeamThreadRange(const Impl::B200AdapterTeamMember& m, I1 b, I2 e);
template <class I>
KB200_TEAM_FUNCTION kb200::Impl::TeamVectorRangeStruct<I> TeamVectorRange(const Impl::B200AdapterTeamMember& m, I count);
template <class I1, class I2>
KB200_TEAM_FUNCTION kb200::Impl::TeamVectorRangeStruct<std::common_type_t<I1, I2>> TeamVectorRange(const Impl::B200AdapterTeamMember& m, I1 b, I2 e);
template <class I>
KB200_TEAM_FUNCTION kb200::Impl::ThreadVectorRangeStruct<I> ThreadVectorRange(const Impl::B200AdapterTeamMember& m, I count);
template <class I1, class I2>
KB200_TEAM_FUNCTION kb200::Impl::ThreadVectorRangeStruct<std::common_type_t<I1, I2>> ThreadVectorRange(const Impl::B200AdapterTeamMember& m, I1 b, I2 e);
KB200_TEAM_FUNCTION kb200::Impl::ThreadSingleStruct PerTeam(const Impl::B200AdapterTeamMember& m);
KB200_TEAM_FUNCTION kb200::Impl::VectorSingleStruct PerThread(const Impl::B200AdapterTeamMember& m);
template <class L>
KB200_TEAM_FUNCTION void single(const kb200::Impl::VectorSingleStruct& s, const L& f);
template <class L>
KB200_TEAM_FUNCTION void single(const kb200::Impl::ThreadSingleStruct& s, const L& f);
template <class L, class T>
KB200_TEAM_FUNCTION void single(const kb200::Impl::VectorSingleStruct& s, const L& f, T& val);
template <class L, class T>
KB200_TEAM_FUNCTION void single(const kb200::Impl::ThreadSingleStruct& s, const L& f, T& val);
}  // namespace Kokkos

#include <Kokkos_Core.hpp>
// a backend header: allowed to see the reference's implementation headers (as its own backends do)
#ifndef KOKKOS_IMPL_PUBLIC_INCLUDE
#define KOKKOS_IMPL_PUBLIC_INCLUDE
#define KOKKOS_B200_UNDEF_PUBLIC_INCLUDE
#endif
#include <impl/Kokkos_ExecSpaceManager.hpp>
#ifdef KOKKOS_B200_UNDEF_PUBLIC_INCLUDE
#undef KOKKOS_IMPL_PUBLIC_INCLUDE
#undef KOKKOS_B200_UNDEF_PUBLIC_INCLUDE
#endif
#if !defined(KOKKOS_ENABLE_CUDA)
#error "Kokkos::B200 needs a CUDA-enabled build of Kokkos (it reuses Kokkos::CudaSpace and the CUDA function annotations)"
#endif

#include <iosfwd>
#include <map>
#include <memory>
#include <mutex>
#include <string>

namespace Kokkos {

namespace Impl {
namespace B200Adapter {
inline std::map<const void*, std::shared_ptr<Kokkos::Cuda>>& cuda_companions() {
  static std::map<const void*, std::shared_ptr<Kokkos::Cuda>> companions;
  return companions;
}
inline std::mutex& cuda_companions_mutex() {
  static std::mutex m;
  return m;
}
}  // namespace B200Adapter
}  // namespace Impl

class B200 {
 public:
  using execution_space      = B200;
  using memory_space         = CudaSpace;
  using device_type          = Kokkos::Device<execution_space, memory_space>;
  using size_type            = memory_space::size_type;
  using array_layout         = LayoutLeft;
  using scratch_memory_space = ScratchMemorySpace<B200>;

  B200() : m_space() {}  // the default instance (created by Kokkos::initialize through the space factory)
  explicit B200(cudaStream_t stream) : m_space(stream) {}
  // deduced parameter: a braced list ({{0, 0}} in MDRangePolicy calls) must never be tried as an execution space
  template <class S, std::enable_if_t<std::is_same_v<S, kb200::B200>, int> = 0>
  explicit B200(const S& s) : m_space(s) {}

  static void impl_initialize(InitializationSettings const& settings) {
    int device = 0;
    if (Cuda::impl_is_initialized()) device = Cuda().cuda_device();  // same device as the reference's Cuda space
    else if (settings.has_device_id()) device = settings.get_device_id();
    kb200::initialize(kb200::InitializationSettings().set_device_id(device));
  }
  static void impl_finalize() {
    {
      std::lock_guard<std::mutex> lock(Impl::B200Adapter::cuda_companions_mutex());
      Impl::B200Adapter::cuda_companions().clear();
    }
    kb200::finalize();
  }
  static int impl_is_initialized() { return kb200::is_initialized() ? 1 : 0; }
  static void impl_static_fence(const std::string& name) {
    Kokkos::Tools::Experimental::Impl::profile_fence_event<B200>(
        name, Kokkos::Tools::Experimental::SpecialSynchronizationCases::GlobalDeviceSynchronization, [&]() { kb200::fence(name); });
  }
  void fence(const std::string& name = "Kokkos::B200::fence(): Unnamed Instance Fence") const {
    Kokkos::Tools::Experimental::Impl::profile_fence_event<B200>(
        name, Kokkos::Tools::Experimental::Impl::DirectFenceIDHandle{impl_instance_id()}, [&]() { m_space.fence(name); });
  }
  int concurrency() const { return m_space.concurrency(); }
  void print_configuration(std::ostream& os, bool verbose = false) const { m_space.print_configuration(os, verbose); }
  static const char* name() { return "B200"; }
  uint32_t impl_instance_id() const noexcept { return m_space.impl_instance_id(); }
  cudaStream_t cuda_stream() const { return m_space.cuda_stream(); }
  int cuda_device() const { return m_space.cuda_device(); }
  const kb200::B200& impl_kb200() const { return m_space; }
  // Views in CudaSpace that are allocated, zero-filled or copied "on" this instance (view_alloc(label, exec), the execution-space
  // overloads of deep_copy) go through the reference's CudaSpace / DeepCopy, whose overloads take a Kokkos::Cuda
  // (Cuda/Kokkos_CudaSpace.hpp:83-86,476-478): hand them one that wraps the SAME stream, created on first use
  // (an execution space object may not be larger than two pointers, impl/Kokkos_ExecSpaceManager.hpp:105, so the companion lives in a
  // registry keyed by the instance; impl_finalize() drops it before the reference finalizes its Cuda space)
  operator const Kokkos::Cuda&() const {
    std::lock_guard<std::mutex> lock(Impl::B200Adapter::cuda_companions_mutex());
    std::shared_ptr<Kokkos::Cuda>& slot = Impl::B200Adapter::cuda_companions()[(const void*)m_space.impl_instance()];
    if (!slot || slot->cuda_stream() != cuda_stream()) slot = std::make_shared<Kokkos::Cuda>(cuda_stream());
    return *slot;
  }

 private:
  friend bool operator==(B200 const& a, B200 const& b) { return a.m_space == b.m_space; }
  friend bool operator!=(B200 const& a, B200 const& b) { return !(a == b); }
  kb200::B200 m_space;
};

namespace Tools {
namespace Experimental {
template <>
struct DeviceTypeTraits<B200> {
  static constexpr DeviceType id = DeviceType::Unknown;  // the enum has no free slot (impl/Kokkos_Profiling_Interface.hpp:40-51)
  static int device_id(const B200& exec) { return exec.cuda_device(); }
};
}  // namespace Experimental
}  // namespace Tools

template <>
struct default_outer_direction<Kokkos::B200> {
  using type                     = Iterate;
  static constexpr Iterate value = Iterate::Left;
};
template <>
struct default_inner_direction<Kokkos::B200> {
  using type                     = Iterate;
  static constexpr Iterate value = Iterate::Left;
};

namespace Impl {

// device code (active memory space CudaSpace) may touch B200 team scratch, as Cuda/Kokkos_Cuda.hpp:267-273 says for Cuda's
template <>
struct MemorySpaceAccess<Kokkos::CudaSpace, Kokkos::B200::scratch_memory_space> {
  enum : bool { assignable = false };
  enum : bool { accessible = true };
  enum : bool { deepcopy = false };
};

// one definition per program (C++17 inline variable): registers the space with Kokkos::initialize / finalize / fence
inline int g_b200_space_factory_initialized = initialize_space_factory<Kokkos::B200>("151_B200");

template <>
inline TileSizeProperties get_tile_size_properties<Kokkos::B200>(const Kokkos::B200& space) {
  const kb200::Impl::TileSizeProperties k = kb200::Impl::get_tile_size_properties(space.impl_kb200());
  TileSizeProperties properties;
  properties.max_threads               = k.max_threads;
  properties.default_largest_tile_size = k.default_largest_tile_size;
  properties.default_tile_size         = k.default_tile_size;
  properties.max_total_tile_size       = k.max_total_tile_size;
  return properties;
}

template <>
struct ZeroMemset<Kokkos::B200> {
  ZeroMemset(const Kokkos::B200& exec, void* dst, size_t cnt) {
    kb200::Impl::throw_on_error(b200_memset_async(exec.impl_kb200().impl_instance(), dst, 0, cnt));
  }
};

namespace B200Adapter {
// What the reference's Cuda launch path does before every kernel (Cuda/Kokkos_Cuda_KernelLaunch.hpp:693): without relocatable
// device code each translation unit has its own copy of desul's lock-array pointers (the fallback for atomics on types wider
// than 8 bytes, e.g. Kokkos::complex<double>); user functors run on this space may use them too.
inline void before_launch() { desul::ensure_cuda_lock_arrays_on_device(); }

// ---- policy translation -------------------------------------------------------------------------------------------
template <class S>
struct schedule_of { using type = kb200::Schedule<kb200::Static>; };
template <>
struct schedule_of<Kokkos::Schedule<Kokkos::Dynamic>> { using type = kb200::Schedule<kb200::Dynamic>; };

template <class Policy>
using kb_range_policy =
    kb200::RangePolicy<kb200::B200, kb200::IndexType<typename Policy::index_type>, typename schedule_of<typename Policy::schedule_type>::type,
                       kb200::LaunchBounds<Policy::launch_bounds::maxTperB, Policy::launch_bounds::minBperSM>, typename Policy::work_tag>;
template <class Policy>
kb_range_policy<Policy> to_kb(const Policy& p) {
  return kb_range_policy<Policy>(p.space().impl_kb200(), p.begin(), p.end());
}

constexpr kb200::Iterate to_kb(Kokkos::Iterate d) {
  return d == Kokkos::Iterate::Left ? kb200::Iterate::Left : (d == Kokkos::Iterate::Right ? kb200::Iterate::Right : kb200::Iterate::Default);
}
template <class Policy>
using kb_md_policy = kb200::MDRangePolicy<kb200::B200, kb200::Rank<(unsigned)Policy::rank, to_kb(Policy::outer_direction), to_kb(Policy::inner_direction)>,
                                          kb200::IndexType<typename Policy::index_type>,
                                          kb200::LaunchBounds<Policy::launch_bounds::maxTperB, Policy::launch_bounds::minBperSM>, typename Policy::work_tag>;
template <class Policy>
kb_md_policy<Policy> to_kb_md(const Policy& p) {
  using I = typename Policy::index_type;
  kb200::Array<I, (size_t)Policy::rank> lo, up, tile;
  for (int r = 0; r < Policy::rank; ++r) {
    lo[r]   = (I)p.m_lower[r];
    up[r]   = (I)p.m_upper[r];
    tile[r] = (I)p.m_tile[r];
  }
  return kb_md_policy<Policy>(p.space().impl_kb200(), lo, up, tile);
}

// ---- the reference's pointer-style reducer as the kernels' value-style "Red" ------------------------------------------
template <class ReducerType>
struct Red {
  using value_type = typename ReducerType::value_type;
  ReducerType r;
  KOKKOS_FORCEINLINE_FUNCTION void init(value_type& v) const { r.init(&v); }
  KOKKOS_FORCEINLINE_FUNCTION void join(value_type& d, const value_type& s) const { r.join(&d, &s); }
  KOKKOS_FORCEINLINE_FUNCTION void final(value_type& v) const { r.final(&v); }
};

// runtime-length array reductions (value_type[] + value_count, FunctorAnalysis.hpp:865-958): the functor's own array
// init/join/final are used by the kernel layer's array-reduce kernels (per-thread accumulator arrays, kb200/impl/ArrayReduceKernel.hpp)
template <class ReducerType>
inline constexpr bool is_array_reduction = (ReducerType::static_value_size() == 0);

template <class T, class Tag, class KbPolicy, class Functor>
void array_reduce(const KbPolicy& policy, const Functor& f, int count, T* host, T* dev) {
  int rc;
  if (count <= 8) rc = kb200::Impl::array_reduce_launch<T, Tag>(policy, f, count, host, dev, std::integral_constant<int, 8>{});
  else if (count <= 32) rc = kb200::Impl::array_reduce_launch<T, Tag>(policy, f, count, host, dev, std::integral_constant<int, 32>{});
  else if (count <= 64) rc = kb200::Impl::array_reduce_launch<T, Tag>(policy, f, count, host, dev, std::integral_constant<int, 64>{});
  else rc = kb200::Impl::array_reduce_launch_big<T, Tag>(policy, f, count, host, dev);  // any length: accumulators in global memory
  kb200::Impl::throw_on_error(rc);
}
}  // namespace B200Adapter

// =========================================================================================== RangePolicy
template <class FunctorType, class... Traits>
class ParallelFor<FunctorType, Kokkos::RangePolicy<Traits...>, Kokkos::B200> {
 public:
  using Policy       = Kokkos::RangePolicy<Traits...>;
  using functor_type = FunctorType;

  ParallelFor(const FunctorType& arg_functor, const Policy& arg_policy) : m_functor(arg_functor), m_policy(arg_policy) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const {
    B200Adapter::before_launch(); kb200::parallel_for(B200Adapter::to_kb(m_policy), m_functor); }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
};

template <class CombinedFunctorReducerType, class... Traits>
class ParallelReduce<CombinedFunctorReducerType, Kokkos::RangePolicy<Traits...>, Kokkos::B200> {
 public:
  using Policy       = Kokkos::RangePolicy<Traits...>;
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
      B200Adapter::array_reduce<value_type, typename Policy::work_tag>(B200Adapter::to_kb(m_policy), m_functor_reducer.get_functor(),
                                                                       (int)m_functor_reducer.get_reducer().value_count(), host, dev);
    } else {
      using R = B200Adapter::Red<ReducerType>;
      kb200::Impl::reduce_dispatch(B200Adapter::to_kb(m_policy), m_functor_reducer.get_functor(), R{m_functor_reducer.get_reducer()},
                                   kb200::Impl::ResultTarget<value_type>{host, dev});
    }
  }

 private:
  const CombinedFunctorReducerType m_functor_reducer;
  const Policy m_policy;
  const pointer_type m_result_ptr;
  const bool m_result_ptr_device_accessible;
};

template <class FunctorType, class... Traits>
class ParallelScan<FunctorType, Kokkos::RangePolicy<Traits...>, Kokkos::B200> {
 public:
  using Policy = Kokkos::RangePolicy<Traits...>;
  using Analysis = Kokkos::Impl::FunctorAnalysis<FunctorPatternInterface::SCAN, Policy, FunctorType, void>;
  using value_type = typename Analysis::value_type;
  using functor_type = FunctorType;

  ParallelScan(const FunctorType& arg_functor, const Policy& arg_policy) : m_functor(arg_functor), m_policy(arg_policy) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const {
    B200Adapter::before_launch();
    using KP  = B200Adapter::kb_range_policy<Policy>;
    using Red = kb200::Impl::FunctorReducer<FunctorType, value_type, typename Policy::work_tag>;
    kb200::Impl::throw_on_error(
        kb200::Impl::GenericScan<KP, FunctorType, Red>::run(B200Adapter::to_kb(m_policy), m_functor, Red{m_functor}, (value_type*)nullptr, (value_type*)nullptr));
  }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
};

template <class FunctorType, class ReturnType, class... Traits>
class ParallelScanWithTotal<FunctorType, Kokkos::RangePolicy<Traits...>, ReturnType, Kokkos::B200> {
 public:
  using Policy = Kokkos::RangePolicy<Traits...>;
  using Analysis = Kokkos::Impl::FunctorAnalysis<FunctorPatternInterface::SCAN, Policy, FunctorType, ReturnType>;
  using value_type = typename Analysis::value_type;
  using functor_type = FunctorType;

  template <class ViewType>
  ParallelScanWithTotal(const FunctorType& arg_functor, const Policy& arg_policy, const ViewType& arg_result_view)
      : m_functor(arg_functor),
        m_policy(arg_policy),
        m_result_ptr(arg_result_view.data()),
        m_result_ptr_device_accessible(MemorySpaceAccess<Kokkos::CudaSpace, typename ViewType::memory_space>::accessible) {}
  Policy const& get_policy() const { return m_policy; }
  void execute() const {
    B200Adapter::before_launch();
    using KP  = B200Adapter::kb_range_policy<Policy>;
    using Red = kb200::Impl::FunctorReducer<FunctorType, value_type, typename Policy::work_tag>;
    value_type* const host = m_result_ptr_device_accessible ? nullptr : (value_type*)m_result_ptr;
    value_type* const dev  = m_result_ptr_device_accessible ? (value_type*)m_result_ptr : nullptr;
    kb200::Impl::throw_on_error(kb200::Impl::GenericScan<KP, FunctorType, Red>::run(B200Adapter::to_kb(m_policy), m_functor, Red{m_functor}, host, dev));
  }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
  value_type* const m_result_ptr;
  const bool m_result_ptr_device_accessible;
};

// =========================================================================================== MDRangePolicy
template <class FunctorType, class... Traits>
class ParallelFor<FunctorType, Kokkos::MDRangePolicy<Traits...>, Kokkos::B200> {
 public:
  using Policy       = Kokkos::MDRangePolicy<Traits...>;
  using functor_type = FunctorType;

  ParallelFor(const FunctorType& arg_functor, const Policy& arg_policy) : m_functor(arg_functor), m_policy(arg_policy) {}
  Policy const& get_policy() const { return m_policy; }
  template <typename P, typename F>
  static int max_tile_size_product(const P&, const F&) { return 1024; }
  void execute() const {
    B200Adapter::before_launch(); kb200::parallel_for(B200Adapter::to_kb_md(m_policy), m_functor); }

 private:
  const FunctorType m_functor;
  const Policy m_policy;
};

template <class CombinedFunctorReducerType, class... Traits>
class ParallelReduce<CombinedFunctorReducerType, Kokkos::MDRangePolicy<Traits...>, Kokkos::B200> {
 public:
  using Policy       = Kokkos::MDRangePolicy<Traits...>;
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
  template <typename P, typename F>
  static int max_tile_size_product(const P&, const F&) { return 512; }

  void execute() const {
    B200Adapter::before_launch();
    value_type* const host = m_result_ptr_device_accessible ? nullptr : (value_type*)m_result_ptr;
    value_type* const dev  = m_result_ptr_device_accessible ? (value_type*)m_result_ptr : nullptr;
    if constexpr (B200Adapter::is_array_reduction<ReducerType>) {
      B200Adapter::array_reduce<value_type, typename Policy::work_tag>(B200Adapter::to_kb_md(m_policy), m_functor_reducer.get_functor(),
                                                                       (int)m_functor_reducer.get_reducer().value_count(), host, dev);
    } else {
      using R = B200Adapter::Red<ReducerType>;
      kb200::Impl::reduce_dispatch(B200Adapter::to_kb_md(m_policy), m_functor_reducer.get_functor(), R{m_functor_reducer.get_reducer()},
                                   kb200::Impl::ResultTarget<value_type>{host, dev});
    }
  }

 private:
  const CombinedFunctorReducerType m_functor_reducer;
  const Policy m_policy;
  const pointer_type m_result_ptr;
  const bool m_result_ptr_device_accessible;
};

}  // namespace Impl
}  // namespace Kokkos

#include "Kokkos_B200_Team.hpp"
#include "Kokkos_B200_UniqueToken.hpp"

#endif
