// kb200/View.hpp -- View<T*>-style device allocations for the B200 execution space.
//
// Covers what the hot path needs of core/src/Kokkos_View.hpp + core/src/Cuda/Kokkos_CudaSpace.{hpp,cpp}:
//   * View<T>, View<T*> ... View<T********>  (rank 0..8, run-time extents), LayoutLeft / LayoutRight,
//     memory spaces B200Space (device, CudaSpace::allocate -> b200_malloc) and HostSpace,
//     MemoryTraits<Unmanaged> / pointer-wrapping constructors, labels, ref-counted ownership
//     (impl/Kokkos_SharedAlloc.*), zero-initialisation at allocation (View/Kokkos_ViewAlloc.hpp:100-166 ->
//     ZeroMemset<B200> = b200_memset_async + fence), WithoutInitializing;
//   * trailing static extents (View<T*[3]>), LayoutStride Views carrying explicit strides;
//   * deep_copy between spaces / from a scalar (element-wise through strides where needed), create_mirror_view[_and_copy],
//     subview: rank-1 ranges keep the parent's type, the general form (index / range / ALL per dimension) yields LayoutStride.
// Device Views default to LayoutLeft as in the reference's Cuda backend (first index fastest).
// DualView/DynRankView, mdspan interop, resize/realloc and layout-transposing host<->device copies are out of scope
// (SURVEY.md section 2 rows 11, 24).
#ifndef KB200_VIEW_HPP
#define KB200_VIEW_HPP

#include "B200.hpp"
#include <cstring>
#include <string>
#include <type_traits>
#include <utility>

namespace kb200 {

struct LayoutLeft {};
struct LayoutRight {};
// LayoutStride (core/src/Kokkos_Layout.hpp:148-230): explicit per-dimension strides.  Views of this layout come out of
// multi-dimensional subview() and carry their strides; they are also what scratch sizing (shmem_size(order_dimensions(...))) takes.
struct ALL_t {};
constexpr ALL_t ALL{};
struct LayoutStride {
  size_t dimension[8] = {1, 1, 1, 1, 1, 1, 1, 1};
  size_t stride[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  template <class IOrder, class IDim>
  static LayoutStride order_dimensions(int rank, const IOrder* order, const IDim* dims) {
    LayoutStride l;
    size_t n = 1;
    for (int r = 0; r < rank; ++r) {
      l.stride[order[r]] = n;
      l.dimension[order[r]] = (size_t)dims[order[r]];
      n *= (size_t)dims[order[r]];
    }
    return l;
  }
};
struct HostSpace {
  using memory_space = HostSpace;
  static constexpr const char* name() { return "Host"; }
};
struct B200Space {
  using memory_space = B200Space;
  using execution_space = B200;
  static constexpr const char* name() { return "B200"; }
};
struct B200HostPinnedSpace {  // CudaHostPinnedSpace: host memory the device can read/write
  using memory_space = B200HostPinnedSpace;
  static constexpr const char* name() { return "B200HostPinned"; }
};
enum MemoryTraitsFlags { Unmanaged = 1, RandomAccess = 2, Atomic = 4, Restrict = 8 };
template <unsigned F>
struct MemoryTraits { static constexpr unsigned flags = F; };
using MemoryUnmanaged = MemoryTraits<Unmanaged>;

struct WithoutInitializing_t {};
constexpr WithoutInitializing_t WithoutInitializing{};
struct ViewAllocProp {  // view_alloc(WithoutInitializing, "label")
  std::string label;
  bool init = true;
  const B200* space = nullptr;
};
inline ViewAllocProp view_alloc(WithoutInitializing_t, const std::string& l) { return ViewAllocProp{l, false, nullptr}; }
inline ViewAllocProp view_alloc(const std::string& l, WithoutInitializing_t) { return ViewAllocProp{l, false, nullptr}; }
inline ViewAllocProp view_alloc(const std::string& l) { return ViewAllocProp{l, true, nullptr}; }
inline ViewAllocProp view_alloc(const B200& s, const std::string& l) { return ViewAllocProp{l, true, &s}; }
inline ViewAllocProp view_alloc(const B200& s, WithoutInitializing_t, const std::string& l) { return ViewAllocProp{l, false, &s}; }

// element proxy of View<..., MemoryTraits<Atomic>> (defined in Atomic.hpp): every read-modify-write through it is atomic
template <class T>
struct AtomicDataElement;

namespace Impl {
template <class D> struct data_type_rank { static constexpr int value = 0; using type = D; };
template <class D> struct data_type_rank<D*> { static constexpr int value = 1 + data_type_rank<D>::value; using type = typename data_type_rank<D>::type; };
// static extents: double[100], double*[3] ... (dynamic dimensions first, as in Kokkos data types)
template <class D, size_t N> struct data_type_rank<D[N]> { static constexpr int value = 1 + data_type_rank<D>::value; using type = typename data_type_rank<D>::type; };
template <class D> struct data_type_dynamic_rank { static constexpr int value = 0; };
template <class D> struct data_type_dynamic_rank<D*> { static constexpr int value = 1 + data_type_dynamic_rank<D>::value; };
template <class D, size_t N> struct data_type_dynamic_rank<D[N]> { static constexpr int value = data_type_dynamic_rank<D>::value; };
// static extent of dimension r (0 = dynamic); for D = T*..*[N1][N2] the static dimensions are the trailing ones, outermost first
template <class D> struct data_type_static { KB200_FUNCTION static constexpr size_t get(int) { return 0; } };
template <class D> struct data_type_static<D*> { KB200_FUNCTION static constexpr size_t get(int r) { return data_type_static<D>::get(r); } };
template <class D, size_t N> struct data_type_static<D[N]> {
  KB200_FUNCTION static constexpr size_t get(int r) {  // r counts static dimensions from the outermost
    return r == 0 ? N : data_type_static<D>::get(r - 1);
  }
};

template <class... P> struct view_props;
template <> struct view_props<> {
  using layout = void; using space = void; static constexpr unsigned traits = 0;
};
template <class First, class... Rest>
struct view_props<First, Rest...> {
  using next = view_props<Rest...>;
  static constexpr bool is_layout = std::is_same<First, LayoutLeft>::value || std::is_same<First, LayoutRight>::value || std::is_same<First, LayoutStride>::value;
  static constexpr bool is_space = std::is_same<First, HostSpace>::value || std::is_same<First, B200Space>::value ||
                                   std::is_same<First, B200HostPinnedSpace>::value || std::is_same<First, B200>::value ||
                                   std::is_same<First, ScratchMemorySpace<B200>>::value;
  using layout = std::conditional_t<is_layout, First, typename next::layout>;
  using space_raw = std::conditional_t<is_space, First, typename next::space>;
  using space = std::conditional_t<std::is_same<space_raw, B200>::value, B200Space, space_raw>;
  template <class T, class = void> struct flags_of { static constexpr unsigned value = 0; };
  template <class T> struct flags_of<T, std::void_t<decltype(T::flags)>> { static constexpr unsigned value = T::flags; };
  static constexpr unsigned traits = flags_of<First>::value | next::traits;
};

// host-side allocation record (SharedAllocationRecord's role); never dereferenced on the device
struct AllocRecord {
  int refcount = 1;
  void* ptr = nullptr;
  int kind = 0;  // 0 device, 1 host malloc, 2 host pinned
  std::shared_ptr<b200_instance> inst;
  std::string label;
};
inline void release(AllocRecord* r) {
  if (!r) return;
  if (__atomic_sub_fetch(&r->refcount, 1, __ATOMIC_ACQ_REL) == 0) {
    if (r->ptr) {
      if (r->kind == 0) b200_free(r->inst.get(), r->ptr);
      else if (r->kind == 1) std::free(r->ptr);
      else b200_free_host_pinned(r->ptr);
    }
    delete r;
  }
}
}  // namespace Impl

namespace Impl {
// only LayoutStride Views pay for stride storage (empty base otherwise)
template <bool Strided> struct ViewStrides {};
template <> struct ViewStrides<true> { size_t m_stride[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };
template <class... Props> struct view_is_strided : std::is_same<typename view_props<Props...>::layout, LayoutStride> {};
}  // namespace Impl

template <class DataType, class... Props>
class View : public Impl::ViewStrides<Impl::view_is_strided<Props...>::value> {
  using props = Impl::view_props<Props...>;

 public:
  static constexpr int rank = Impl::data_type_rank<DataType>::value;
  static_assert(rank <= 8, "kb200::View supports rank 0..8");
  using value_type = typename Impl::data_type_rank<DataType>::type;
  using non_const_value_type = std::remove_const_t<value_type>;
  using memory_space = std::conditional_t<std::is_void<typename props::space>::value, B200Space, typename props::space>;
  using array_layout = std::conditional_t<std::is_void<typename props::layout>::value,
                                          std::conditional_t<std::is_same<memory_space, HostSpace>::value, LayoutRight, LayoutLeft>,
                                          typename props::layout>;
  using execution_space = B200;
  using size_type = size_t;
  using pointer_type = value_type*;
  static constexpr bool is_atomic = (props::traits & Atomic) != 0;  // impl/Kokkos_Atomic_View.hpp: operator() yields an atomic proxy
  using reference_type = std::conditional_t<is_atomic, AtomicDataElement<value_type>, value_type&>;
  static constexpr bool is_managed = !(props::traits & Unmanaged);
  static constexpr bool is_device = !std::is_same<memory_space, HostSpace>::value;
  static constexpr bool is_strided = std::is_same<array_layout, LayoutStride>::value;
  using HostMirror = View<std::remove_const_t<DataType>, array_layout, HostSpace>;
  using host_mirror_space = HostSpace;
  // bytes an allocation of these extents takes (core/src/Kokkos_View.hpp required_allocation_size)
  static constexpr size_t required_allocation_size(size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0,
                                                   size_t n6 = 0, size_t n7 = 0) {
    return (rank > 0 ? n0 : 1) * (rank > 1 ? n1 : 1) * (rank > 2 ? n2 : 1) * (rank > 3 ? n3 : 1) * (rank > 4 ? n4 : 1) * (rank > 5 ? n5 : 1) *
           (rank > 6 ? n6 : 1) * (rank > 7 ? n7 : 1) * sizeof(value_type);
  }
  using traits = View;  // View::traits::memory_space, ::array_layout, ::value_type ... (the reference's ViewTraits members)
  using non_const_type = View<DataType, Props...>;

  KB200_INLINE_FUNCTION View() : m_data(nullptr), m_rec(nullptr) { for (int r = 0; r < 8; ++r) m_ext[r] = 0; }

  // allocating constructors
  explicit View(const std::string& label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) { allocate(ViewAllocProp{label, true, nullptr}, n0, n1, n2, n3, n4, n5, n6, n7); }
  explicit View(const char* label, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) { allocate(ViewAllocProp{label, true, nullptr}, n0, n1, n2, n3, n4, n5, n6, n7); }
  explicit View(const ViewAllocProp& p, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) { allocate(p, n0, n1, n2, n3, n4, n5, n6, n7); }
  // wrapping (unmanaged) constructor
  KB200_INLINE_FUNCTION View(pointer_type ptr, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) : m_data(ptr), m_rec(nullptr) {
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
  }

  // View over team/thread scratch memory: View<T*, ScratchSpace, Unmanaged>(team.team_scratch(level), n)
  template <class S, class = typename S::is_scratch_tag>
  KB200_INLINE_FUNCTION View(const S& scratch, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) : m_rec(nullptr) {
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
    m_data = static_cast<pointer_type>(scratch.get_shmem_aligned(size() * sizeof(value_type), (ptrdiff_t)scratch_value_alignment));
  }
  // scratch Views are aligned to max(sizeof(T), alignof(T), 8) and shmem_size() reserves that much slack
  // (core/src/View/Kokkos_ViewMapping / Kokkos_View.hpp scratch_value_alignment; TestTeam.hpp:1601-1640 checks it)
  static constexpr size_t scratch_value_alignment =
      sizeof(value_type) > (alignof(value_type) > 8 ? alignof(value_type) : 8) ? sizeof(value_type) : (alignof(value_type) > 8 ? alignof(value_type) : 8);
  static size_t shmem_size(const LayoutStride& l) {
    size_t n = 1;
    for (int r = 0; r < rank; ++r) n *= l.dimension[r];
    return n * sizeof(value_type) + scratch_value_alignment;
  }
  static constexpr size_t shmem_size(size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
    return (rank > 0 ? n0 : 1) * (rank > 1 ? n1 : 1) * (rank > 2 ? n2 : 1) * (rank > 3 ? n3 : 1) * (rank > 4 ? n4 : 1) * (rank > 5 ? n5 : 1) * (rank > 6 ? n6 : 1) * (rank > 7 ? n7 : 1) * sizeof(value_type) + scratch_value_alignment;
  }

  KB200_INLINE_FUNCTION View(const View& o) : m_data(o.m_data), m_rec(o.m_rec) { copy_ext(o); retain(); }
  KB200_INLINE_FUNCTION View(View&& o) noexcept : m_data(o.m_data), m_rec(o.m_rec) { copy_ext(o); o.m_rec = nullptr; o.m_data = nullptr; }
  // const-adding / trait-changing conversions between compatible Views
  template <class D2, class... P2, class = std::enable_if_t<std::is_convertible<typename View<D2, P2...>::pointer_type, pointer_type>::value &&
                                                            View<D2, P2...>::rank == rank>>
  KB200_INLINE_FUNCTION View(const View<D2, P2...>& o) : m_data(o.data()), m_rec(is_managed ? o.impl_record() : nullptr) {
    using Src = View<D2, P2...>;
    // layout compatibility, as the reference's ViewMapping assignability rules (core/src/View/Kokkos_ViewMapping.hpp): same
    // layout, anything into LayoutStride, and Left <-> Right only where both describe the same memory (rank <= 1)
    static_assert(is_strided || std::is_same<array_layout, typename Src::array_layout>::value || (rank <= 1 && !Src::is_strided),
                  "View assignment must have compatible layouts (LayoutLeft <-> LayoutRight differ for rank > 1; a strided View "
                  "cannot be assigned to a contiguous layout)");
    for (int r = 0; r < 8; ++r) m_ext[r] = o.extent(r);
    if constexpr (is_strided)
      for (int r = 0; r < 8; ++r) this->m_stride[r] = r < rank ? o.stride(r) : 0;
    retain();
  }
  KB200_INLINE_FUNCTION View& operator=(const View& o) {
    if (this != &o) { drop(); m_data = o.m_data; m_rec = o.m_rec; copy_ext(o); retain(); }
    return *this;
  }
  KB200_INLINE_FUNCTION View& operator=(View&& o) noexcept {
    if (this != &o) { drop(); m_data = o.m_data; m_rec = o.m_rec; copy_ext(o); o.m_rec = nullptr; o.m_data = nullptr; }
    return *this;
  }
  KB200_INLINE_FUNCTION ~View() { drop(); }

  // element access
  template <int R = rank, std::enable_if_t<R == 0, int> = 0>
  KB200_FORCEINLINE_FUNCTION reference_type operator()() const { return ref(0); }
  template <class I0, int R = rank, std::enable_if_t<R == 1, int> = 0>
  KB200_FORCEINLINE_FUNCTION reference_type operator()(const I0 i0) const {
    if constexpr (is_strided) return ref((size_t)i0 * this->m_stride[0]);
    else return ref((size_t)i0);
  }
  template <class I0, int R = rank, std::enable_if_t<R == 1, int> = 0>
  KB200_FORCEINLINE_FUNCTION reference_type operator[](const I0 i0) const { return (*this)(i0); }
  template <class I0, class I1, int R = rank, std::enable_if_t<R == 2, int> = 0>
  KB200_FORCEINLINE_FUNCTION reference_type operator()(const I0 i0, const I1 i1) const {
    if constexpr (is_strided) return ref((size_t)i0 * this->m_stride[0] + (size_t)i1 * this->m_stride[1]);
    else if constexpr (std::is_same<array_layout, LayoutLeft>::value) return ref((size_t)i0 + m_ext[0] * (size_t)i1);
    else return ref((size_t)i1 + m_ext[1] * (size_t)i0);
  }
  template <class I0, class I1, class I2, int R = rank, std::enable_if_t<R == 3, int> = 0>
  KB200_FORCEINLINE_FUNCTION reference_type operator()(const I0 i0, const I1 i1, const I2 i2) const {
    if constexpr (is_strided) return ref((size_t)i0 * this->m_stride[0] + (size_t)i1 * this->m_stride[1] + (size_t)i2 * this->m_stride[2]);
    else if constexpr (std::is_same<array_layout, LayoutLeft>::value) return ref((size_t)i0 + m_ext[0] * ((size_t)i1 + m_ext[1] * (size_t)i2));
    else return ref((size_t)i2 + m_ext[2] * ((size_t)i1 + m_ext[1] * (size_t)i0));
  }

  // rank 4..8: generic mixed-radix offset (LayoutLeft: first index fastest; LayoutRight: last index fastest)
  template <class... Is, int R = rank, std::enable_if_t<(R >= 4) && sizeof...(Is) == (size_t)R, int> = 0>
  KB200_FORCEINLINE_FUNCTION reference_type operator()(const Is... is) const {
    const size_t ix[sizeof...(Is)] = {(size_t)is...};
    size_t off = 0;
    if constexpr (is_strided) {
      for (int r = 0; r < rank; ++r) off += ix[r] * this->m_stride[r];
    } else if constexpr (std::is_same<array_layout, LayoutLeft>::value) {
      for (int r = rank - 1; r >= 0; --r) off = off * m_ext[r] + ix[r];  // (constant trip count: unrolled)
    } else {
      for (int r = 0; r < rank; ++r) off = off * m_ext[r] + ix[r];
    }
    return ref(off);
  }

  // access(i0, ..., i7): element access that ignores the indices beyond the rank (core/src/Kokkos_View.hpp View::access)
  template <class... Is>
  KB200_FORCEINLINE_FUNCTION reference_type access(const Is... is) const {
    static_assert(sizeof...(Is) >= (size_t)rank && sizeof...(Is) <= 8, "kb200::View::access: between rank and 8 indices");
    const size_t ix[sizeof...(Is) > 0 ? sizeof...(Is) : 1] = {(size_t)is...};
    size_t off = 0;
    if constexpr (is_strided) {
      for (int r = 0; r < rank; ++r) off += ix[r] * this->m_stride[r];
    } else if constexpr (std::is_same<array_layout, LayoutLeft>::value) {
      for (int r = rank - 1; r >= 0; --r) off = off * m_ext[r] + ix[r];
    } else {
      for (int r = 0; r < rank; ++r) off = off * m_ext[r] + ix[r];
    }
    return ref(off);
  }

  KB200_FORCEINLINE_FUNCTION pointer_type data() const { return m_data; }
  KB200_FORCEINLINE_FUNCTION size_t extent(int r) const { return r < rank ? m_ext[r] : 1; }
  KB200_FORCEINLINE_FUNCTION int extent_int(int r) const { return (int)extent(r); }
  KB200_FORCEINLINE_FUNCTION size_t size() const {
    size_t s = 1;
    for (int r = 0; r < rank; ++r) s *= m_ext[r];
    return s;
  }
  KB200_FORCEINLINE_FUNCTION size_t span() const {
    if constexpr (is_strided) {
      size_t last = 0;
      for (int r = 0; r < rank; ++r) { if (m_ext[r] == 0) return 0; last += (m_ext[r] - 1) * this->m_stride[r]; }
      return last + 1;
    } else {
      return size();
    }
  }
  KB200_FORCEINLINE_FUNCTION bool span_is_contiguous() const { return span() == size(); }
  KB200_FORCEINLINE_FUNCTION bool is_allocated() const { return m_data != nullptr; }
  KB200_FORCEINLINE_FUNCTION size_t stride(int r) const {
    if constexpr (is_strided) return this->m_stride[r];
    else if (std::is_same<array_layout, LayoutLeft>::value) { size_t s = 1; for (int k = 0; k < r; ++k) s *= m_ext[k]; return s; }
    size_t s = 1; for (int k = rank - 1; k > r; --k) s *= m_ext[k]; return s;
  }
  std::string label() const { return m_rec ? m_rec->label : std::string(); }
  int use_count() const { return m_rec ? m_rec->refcount : 0; }
  KB200_INLINE_FUNCTION Impl::AllocRecord* impl_record() const { return m_rec; }
  void impl_window(size_t offset, size_t count) {
    if constexpr (is_strided) m_data += offset * this->m_stride[0];
    else m_data += offset;
    m_ext[0] = count;
  }
  // subview construction: share `parent`'s allocation record, point at `ptr`, take extents (and strides) as given
  template <class Parent>
  void impl_assign_strided(const Parent& parent, pointer_type ptr, const size_t* ext, const size_t* strides) {
    drop();
    m_data = ptr;
    m_rec = is_managed ? parent.impl_record() : nullptr;
    for (int r = 0; r < 8; ++r) m_ext[r] = r < rank ? ext[r] : 1;
    if constexpr (is_strided)
      for (int r = 0; r < 8; ++r) this->m_stride[r] = r < rank ? strides[r] : 0;
    retain();
  }

 private:
  KB200_FORCEINLINE_FUNCTION reference_type ref(size_t off) const {
    if constexpr (is_atomic) return reference_type(m_data + off);
    else return m_data[off];
  }
  KB200_INLINE_FUNCTION void set_extents(size_t n0, size_t n1, size_t n2, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
    m_ext[0] = rank > 0 ? n0 : 1; m_ext[1] = rank > 1 ? n1 : 1; m_ext[2] = rank > 2 ? n2 : 1;
    m_ext[3] = rank > 3 ? n3 : 1; m_ext[4] = rank > 4 ? n4 : 1; m_ext[5] = rank > 5 ? n5 : 1;
    m_ext[6] = rank > 6 ? n6 : 1; m_ext[7] = rank > 7 ? n7 : 1;
    // static extents occupy the dimensions after the dynamic ones
    constexpr int dyn = Impl::data_type_dynamic_rank<DataType>::value;
    for (int r = dyn; r < rank; ++r) m_ext[r] = Impl::data_type_static<DataType>::get(r - dyn);
    if constexpr (is_strided) {  // built from extents alone: contiguous, first index fastest
      size_t st = 1;
      for (int r = 0; r < rank; ++r) { this->m_stride[r] = st; st *= m_ext[r]; }
    }
  }
  KB200_INLINE_FUNCTION void copy_ext(const View& o) {
    for (int r = 0; r < 8; ++r) m_ext[r] = o.m_ext[r];
    if constexpr (is_strided)
      for (int r = 0; r < 8; ++r) this->m_stride[r] = o.m_stride[r];
  }
  KB200_INLINE_FUNCTION void retain() {
#ifndef __CUDA_ARCH__
    if (m_rec) __atomic_add_fetch(&m_rec->refcount, 1, __ATOMIC_RELAXED);
#endif
  }
  KB200_INLINE_FUNCTION void drop() {
#ifndef __CUDA_ARCH__
    if (m_rec) { Impl::release(m_rec); m_rec = nullptr; }
#endif
  }
  void allocate(const ViewAllocProp& p, size_t n0, size_t n1, size_t n2, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
    static_assert(!std::is_const<value_type>::value, "cannot allocate a View of const");
    set_extents(n0, n1, n2, n3, n4, n5, n6, n7);
    const size_t bytes = size() * sizeof(value_type);
    m_rec = new Impl::AllocRecord();
    m_rec->label = p.label;
    void* ptr = nullptr;
    if (std::is_same<memory_space, B200Space>::value) {
      B200 space = p.space ? *p.space : B200();
      m_rec->inst = Impl::default_instance();
      m_rec->kind = 0;
      if (p.space) m_rec->inst = std::shared_ptr<b200_instance>(std::shared_ptr<b200_instance>(), space.impl_instance());
      int rc = b200_malloc(space.impl_instance(), bytes, &ptr);
      if (rc) { delete m_rec; m_rec = nullptr; Impl::throw_on_error(rc); }
      if (p.init && bytes) {  // value-initialisation of trivially constructible types = zero memset + fence
        Impl::throw_on_error(b200_memset_async(space.impl_instance(), ptr, 0, bytes));
        space.fence("kb200::View: fence after zero-initialisation");
      }
    } else if (std::is_same<memory_space, B200HostPinnedSpace>::value) {
      m_rec->kind = 2;
      int rc = b200_malloc_host_pinned(bytes, &ptr);
      if (rc) { delete m_rec; m_rec = nullptr; Impl::throw_on_error(rc); }
      if (p.init && bytes) std::memset(ptr, 0, bytes);
    } else {
      m_rec->kind = 1;
      if (bytes) {
        ptr = p.init ? std::calloc(bytes, 1) : std::malloc(bytes);
        if (!ptr) { delete m_rec; m_rec = nullptr; throw RawMemoryAllocationFailure("kb200::View: host allocation failed"); }
      }
    }
    m_rec->ptr = ptr;
    m_data = static_cast<pointer_type>(ptr);
  }

  pointer_type m_data;
  size_t m_ext[8];
  Impl::AllocRecord* m_rec;
};

// two Views are equal when they are the same window on the same memory (core/src/Kokkos_View.hpp operator==)
template <class D1, class... P1, class D2, class... P2>
KB200_INLINE_FUNCTION bool operator==(const View<D1, P1...>& a, const View<D2, P2...>& b) {
  using A = View<D1, P1...>; using B = View<D2, P2...>;
  if (!std::is_same<typename A::value_type, typename B::value_type>::value || !std::is_same<typename A::array_layout, typename B::array_layout>::value ||
      !std::is_same<typename A::memory_space, typename B::memory_space>::value || (int)A::rank != (int)B::rank)
    return false;
  if ((const void*)a.data() != (const void*)b.data()) return false;
  for (int r = 0; r < A::rank; ++r)
    if (a.extent(r) != b.extent(r) || a.stride(r) != b.stride(r)) return false;
  return true;
}
template <class D1, class... P1, class D2, class... P2>
KB200_INLINE_FUNCTION bool operator!=(const View<D1, P1...>& a, const View<D2, P2...>& b) { return !(a == b); }

// Kokkos::ViewTraits<DataType, Props...>: the analysed properties; here the View publishes them itself
template <class DataType, class... Props>
using ViewTraits = View<DataType, Props...>;

template <class T> struct is_view : std::false_type {};
template <class D, class... P> struct is_view<View<D, P...>> : std::true_type {};
template <class T> constexpr bool is_view_v = is_view<std::decay_t<T>>::value;

// ---------------------------------------------------------------- subview (rank-1 ranges)
// shares ownership with the parent (the reference's subviews share the allocation record)
template <class D, class... P, class I0, class I1>
View<D, P...> subview(const View<D, P...>& v, const std::pair<I0, I1>& r) {
  static_assert(View<D, P...>::rank == 1, "kb200::subview: rank-1 Views only");
  if ((size_t)r.second > v.extent(0) || (size_t)r.first > (size_t)r.second) throw std::runtime_error("kb200::subview: range out of bounds");
  View<D, P...> s(v);
  s.impl_window((size_t)r.first, (size_t)(r.second - r.first));
  return s;
}

// ---------------------------------------------------------------- subview, general form (core/src/Kokkos_View.hpp subview / ViewMapping)
// One argument per dimension: an index (the dimension is dropped), a pair [begin, end) or ALL.  The result is a LayoutStride
// View of the kept dimensions sharing the parent's allocation.
namespace Impl {
template <class T, int K> struct add_pointers { using type = typename add_pointers<T, K - 1>::type*; };
template <class T> struct add_pointers<T, 0> { using type = T; };
template <class A> struct subview_arg_keeps : std::integral_constant<int, std::is_integral<std::decay_t<A>>::value ? 0 : 1> {};
template <class A>
inline void subview_arg(const A& a, size_t extent, size_t& begin, size_t& count, bool& keep) {
  if constexpr (std::is_integral<A>::value) { begin = (size_t)a; count = 1; keep = false; if (begin >= extent) throw std::runtime_error("kb200::subview: index out of bounds"); }
  else if constexpr (std::is_same<A, ALL_t>::value) { begin = 0; count = extent; keep = true; }
  else {
    begin = (size_t)a.first; count = (size_t)a.second >= begin ? (size_t)a.second - begin : 0; keep = true;
    if ((size_t)a.second > extent || (size_t)a.first > (size_t)a.second) throw std::runtime_error("kb200::subview: range out of bounds");
  }
}
template <class V> struct view_space_of { using type = typename V::memory_space; };
}  // namespace Impl

template <class D, class... P, class A0, class A1, class... As>
auto subview(const View<D, P...>& v, const A0& a0, const A1& a1, const As&... as) {
  using V = View<D, P...>;
  static_assert(V::rank == 2 + (int)sizeof...(As), "kb200::subview: one argument per dimension");
  constexpr int kept = Impl::subview_arg_keeps<A0>::value + Impl::subview_arg_keeps<A1>::value + (0 + ... + Impl::subview_arg_keeps<As>::value);
  using SubData = typename Impl::add_pointers<typename V::value_type, kept>::type;
  using Sub = std::conditional_t<V::is_managed, View<SubData, LayoutStride, typename V::memory_space>,
                                 View<SubData, LayoutStride, typename V::memory_space, MemoryTraits<Unmanaged>>>;
  size_t begin[8] = {0}, count[8] = {0};
  bool keep[8] = {false};
  int r = 0;
  Impl::subview_arg(a0, v.extent(0), begin[0], count[0], keep[0]);
  Impl::subview_arg(a1, v.extent(1), begin[1], count[1], keep[1]);
  r = 2;
  ((Impl::subview_arg(as, v.extent(r), begin[r], count[r], keep[r]), ++r), ...);
  size_t off = 0, ext[8] = {1, 1, 1, 1, 1, 1, 1, 1}, str[8] = {0};
  int k = 0;
  for (int d = 0; d < V::rank; ++d) {
    off += begin[d] * v.stride(d);
    if (keep[d]) { ext[k] = count[d]; str[k] = v.stride(d); ++k; }
  }
  Sub sub;
  sub.impl_assign_strided(v, v.data() + off, ext, str);
  return sub;
}

// ---------------------------------------------------------------- deep_copy
namespace Impl {
// element-wise copy for strided or layout-changing pairs (defined in StdAlgorithms.hpp, after parallel_for)
template <class V1, class V2>
void strided_copy(const B200& space, const V1& dst, const V2& src);

template <class DstSpace, class SrcSpace>
inline void copy_bytes(const B200& space, void* dst, const void* src, size_t bytes) {
  constexpr bool dd = !std::is_same<DstSpace, HostSpace>::value && !std::is_same<DstSpace, B200HostPinnedSpace>::value;
  constexpr bool sd = !std::is_same<SrcSpace, HostSpace>::value && !std::is_same<SrcSpace, B200HostPinnedSpace>::value;
  if (bytes == 0) return;
  if (dd && sd) throw_on_error(b200_memcpy_d2d_async(space.impl_instance(), dst, src, bytes));
  else if (dd) throw_on_error(b200_memcpy_h2d_async(space.impl_instance(), dst, src, bytes));
  else if (sd) throw_on_error(b200_memcpy_d2h_async(space.impl_instance(), dst, src, bytes));
  else std::memcpy(dst, src, bytes);
}
}  // namespace Impl

// deep_copy(dst, src): same extents, same layout, contiguous (core/src/Kokkos_CopyViews.hpp:897-1100 fast path);
// blocking, like the reference's two-argument form
template <class D1, class... P1, class D2, class... P2>
void deep_copy(const B200& space, const View<D1, P1...>& dst, const View<D2, P2...>& src) {
  using V1 = View<D1, P1...>; using V2 = View<D2, P2...>;
  static_assert(std::is_same<typename V1::non_const_value_type, typename V2::non_const_value_type>::value, "deep_copy: value types differ");
  static_assert(V1::rank == V2::rank, "deep_copy: ranks differ");
  for (int r = 0; r < V1::rank; ++r)
    if (dst.extent(r) != src.extent(r)) throw std::runtime_error("kb200::deep_copy: extents differ");
  // strided operands or different layouts: element-wise ViewCopy kernel (Kokkos_CopyViews.hpp:300-560); otherwise one memcpy
  if constexpr (V1::is_strided || V2::is_strided || (V1::rank > 1 && !std::is_same<typename V1::array_layout, typename V2::array_layout>::value)) {
    if (!(V1::is_strided == V2::is_strided && std::is_same<typename V1::array_layout, typename V2::array_layout>::value && dst.span_is_contiguous() &&
          src.span_is_contiguous() && !V1::is_strided)) {
      Impl::strided_copy(space, dst, src);
      return;
    }
  }
  Impl::copy_bytes<typename V1::memory_space, typename V2::memory_space>(space, (void*)dst.data(), (const void*)src.data(),
                                                                         dst.size() * sizeof(typename V1::value_type));
}
template <class D1, class... P1, class D2, class... P2>
void deep_copy(const View<D1, P1...>& dst, const View<D2, P2...>& src) {
  B200 space;
  deep_copy(space, dst, src);
  space.fence("kb200::deep_copy: fence after copy");
}

// deep_copy(scalar, rank-0 View) / deep_copy(rank-0 View, scalar)  (core/src/Kokkos_CopyViews.hpp:1116-1160)
template <class T, class D, class... P, class = std::enable_if_t<View<D, P...>::rank == 0 && std::is_same<T, typename View<D, P...>::non_const_value_type>::value>>
void deep_copy(T& dst, const View<D, P...>& src) {
  B200 space;
  Impl::copy_bytes<HostSpace, typename View<D, P...>::memory_space>(space, &dst, src.data(), sizeof(T));
  space.fence("kb200::deep_copy(scalar, View)");
}
namespace Impl {
// Kokkos::Impl::DeepCopy<DstSpace, SrcSpace>(dst, src, bytes): raw byte copy between memory spaces, used as a constructor call
// (core/src/Cuda/Kokkos_CudaSpace.hpp:473-590); blocking like the reference's no-exec-space form
template <class DstSpace, class SrcSpace, class Exec = B200>
struct DeepCopy {
  DeepCopy(void* dst, const void* src, size_t bytes) {
    B200 space;
    copy_bytes<typename DstSpace::memory_space, typename SrcSpace::memory_space>(space, dst, src, bytes);
    space.fence("kb200::Impl::DeepCopy");
  }
  DeepCopy(const B200& space, void* dst, const void* src, size_t bytes) {
    copy_bytes<typename DstSpace::memory_space, typename SrcSpace::memory_space>(space, dst, src, bytes);
  }
};
}  // namespace Impl

// resize(view, n0...) keeps the common index box, realloc(view, n0...) does not (core/src/Kokkos_CopyViews.hpp:1580-1790)
// (pairs are spelled std::pair<size_t, size_t>, not make_pair: nvcc prints the first spelling it saw of a type into the host
// stubs of extended lambdas, and make_pair's spelling contains a GCC built-in trait that GCC then rejects in a signature)
template <class D, class... P>
void realloc(View<D, P...>& v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
  using V = View<D, P...>;
  static_assert(V::is_managed, "kb200::realloc: the View must own its allocation");
  const size_t n[8] = {n0, n1, n2, n3, n4, n5, n6, n7};
  bool same = v.is_allocated();
  for (int r = 0; r < Impl::data_type_dynamic_rank<D>::value; ++r) same = same && v.extent(r) == n[r];
  if (same) { deep_copy(v, typename V::non_const_value_type()); return; }
  v = V(v.label(), n0, n1, n2, n3, n4, n5, n6, n7);
}
template <class D, class... P>
void resize(View<D, P...>& v, size_t n0 = 0, size_t n1 = 0, size_t n2 = 0, size_t n3 = 0, size_t n4 = 0, size_t n5 = 0, size_t n6 = 0, size_t n7 = 0) {
  using V = View<D, P...>;
  static_assert(V::is_managed, "kb200::resize: the View must own its allocation");
  static_assert(V::rank >= 1 && V::rank <= 3, "kb200::resize: rank 1..3");
  const size_t n[8] = {n0, n1, n2, n3, n4, n5, n6, n7};
  bool same = true;
  for (int r = 0; r < Impl::data_type_dynamic_rank<D>::value; ++r) same = same && v.extent(r) == n[r];
  if (same) return;
  V fresh(v.label(), n0, n1, n2, n3, n4, n5, n6, n7);
  size_t c[3];
  for (int r = 0; r < 3; ++r) c[r] = r < V::rank ? (v.extent(r) < fresh.extent(r) ? v.extent(r) : fresh.extent(r)) : 1;
  if (c[0] * c[1] * c[2] > 0) {
    if constexpr (V::rank == 1) deep_copy(subview(fresh, std::pair<size_t, size_t>(0, c[0])), subview(v, std::pair<size_t, size_t>(0, c[0])));
    else if constexpr (V::rank == 2)
      deep_copy(subview(fresh, std::pair<size_t, size_t>(0, c[0]), std::pair<size_t, size_t>(0, c[1])), subview(v, std::pair<size_t, size_t>(0, c[0]), std::pair<size_t, size_t>(0, c[1])));
    else
      deep_copy(subview(fresh, std::pair<size_t, size_t>(0, c[0]), std::pair<size_t, size_t>(0, c[1]), std::pair<size_t, size_t>(0, c[2])),
                subview(v, std::pair<size_t, size_t>(0, c[0]), std::pair<size_t, size_t>(0, c[1]), std::pair<size_t, size_t>(0, c[2])));
  }
  v = fresh;
}

template <class D, class... P>
typename View<D, P...>::HostMirror create_mirror_view(const View<D, P...>& v) {
  return typename View<D, P...>::HostMirror(view_alloc(WithoutInitializing, v.label() + "_mirror"), v.extent(0), v.extent(1), v.extent(2), v.extent(3), v.extent(4), v.extent(5), v.extent(6), v.extent(7));
}
template <class D, class... P>
typename View<D, P...>::HostMirror create_mirror(const View<D, P...>& v) { return create_mirror_view(v); }
template <class Space, class D, class... P>
auto create_mirror_view_and_copy(const Space&, const View<D, P...>& v) {
  using Dst = View<std::remove_const_t<D>, typename View<D, P...>::array_layout, typename Space::memory_space>;
  Dst d(view_alloc(WithoutInitializing, v.label() + "_copy"), v.extent(0), v.extent(1), v.extent(2), v.extent(3), v.extent(4), v.extent(5), v.extent(6), v.extent(7));
  deep_copy(d, v);
  return d;
}

}  // namespace kb200
#endif
