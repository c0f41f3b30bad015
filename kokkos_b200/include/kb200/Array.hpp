// kb200/Array.hpp -- fixed-size aggregate usable in device code, the role of core/src/Kokkos_Array.hpp:83-215
// (Kokkos::Array<T, N>: aggregate-initialisable, constexpr element access, zero-length specialisation, to_array,
// structured bindings).  Used for MDRangePolicy bounds/tiles and as a reduction value type in functors.
#ifndef KB200_ARRAY_HPP
#define KB200_ARRAY_HPP

#include "Macros.hpp"
#include <cstddef>
#include <tuple>
#include <type_traits>
#include <utility>

namespace kb200 {

template <class T, size_t N>
struct Array {
  T m_elems[N];  // public: this is an aggregate (Array<int, 2>{{0, 1}} and Array<int, 2>{0, 1} both work)

  using value_type = T;
  using size_type = size_t;
  using difference_type = std::ptrdiff_t;
  using reference = T&;
  using const_reference = const T&;
  using pointer = T*;
  using const_pointer = const T*;

  KB200_FORCEINLINE_FUNCTION static constexpr size_type size() { return N; }
  KB200_FORCEINLINE_FUNCTION static constexpr bool empty() { return false; }
  KB200_FORCEINLINE_FUNCTION constexpr size_type max_size() const { return N; }
  template <class I>
  KB200_FORCEINLINE_FUNCTION constexpr reference operator[](const I& i) {
    static_assert(std::is_integral<I>::value || std::is_enum<I>::value, "kb200::Array: index must be integral");
    return m_elems[i];
  }
  template <class I>
  KB200_FORCEINLINE_FUNCTION constexpr const_reference operator[](const I& i) const {
    static_assert(std::is_integral<I>::value || std::is_enum<I>::value, "kb200::Array: index must be integral");
    return m_elems[i];
  }
  KB200_FORCEINLINE_FUNCTION constexpr pointer data() { return m_elems; }
  KB200_FORCEINLINE_FUNCTION constexpr const_pointer data() const { return m_elems; }
  KB200_FORCEINLINE_FUNCTION constexpr pointer begin() { return m_elems; }
  KB200_FORCEINLINE_FUNCTION constexpr const_pointer begin() const { return m_elems; }
  KB200_FORCEINLINE_FUNCTION constexpr pointer end() { return m_elems + N; }
  KB200_FORCEINLINE_FUNCTION constexpr const_pointer end() const { return m_elems + N; }
  KB200_FORCEINLINE_FUNCTION constexpr reference front() { return m_elems[0]; }
  KB200_FORCEINLINE_FUNCTION constexpr const_reference front() const { return m_elems[0]; }
  KB200_FORCEINLINE_FUNCTION constexpr reference back() { return m_elems[N - 1]; }
  KB200_FORCEINLINE_FUNCTION constexpr const_reference back() const { return m_elems[N - 1]; }
  KB200_FORCEINLINE_FUNCTION constexpr void fill(const T& v) {
    for (size_t i = 0; i < N; ++i) m_elems[i] = v;
  }

  friend KB200_FORCEINLINE_FUNCTION constexpr bool operator==(const Array& a, const Array& b) {
    for (size_t i = 0; i < N; ++i)
      if (!(a.m_elems[i] == b.m_elems[i])) return false;
    return true;
  }
  friend KB200_FORCEINLINE_FUNCTION constexpr bool operator!=(const Array& a, const Array& b) { return !(a == b); }
};

template <class T>
struct Array<T, 0> {
  using value_type = T;
  using size_type = size_t;
  using difference_type = std::ptrdiff_t;
  using reference = T&;
  using const_reference = const T&;
  using pointer = T*;
  using const_pointer = const T*;

  KB200_FORCEINLINE_FUNCTION static constexpr size_type size() { return 0; }
  KB200_FORCEINLINE_FUNCTION static constexpr bool empty() { return true; }
  KB200_FORCEINLINE_FUNCTION constexpr size_type max_size() const { return 0; }
  KB200_FORCEINLINE_FUNCTION constexpr pointer data() { return nullptr; }
  KB200_FORCEINLINE_FUNCTION constexpr const_pointer data() const { return nullptr; }
  KB200_FORCEINLINE_FUNCTION constexpr pointer begin() { return nullptr; }
  KB200_FORCEINLINE_FUNCTION constexpr const_pointer begin() const { return nullptr; }
  KB200_FORCEINLINE_FUNCTION constexpr pointer end() { return nullptr; }
  KB200_FORCEINLINE_FUNCTION constexpr const_pointer end() const { return nullptr; }
  friend KB200_FORCEINLINE_FUNCTION constexpr bool operator==(const Array&, const Array&) { return true; }
  friend KB200_FORCEINLINE_FUNCTION constexpr bool operator!=(const Array&, const Array&) { return false; }
};

template <class T, class... Us>
Array(T, Us...) -> Array<T, 1 + sizeof...(Us)>;

// swap usable in device code (core/src/Kokkos_Swap.hpp:28-60): values, C arrays (element-wise) and Array
template <class T>
KB200_FORCEINLINE_FUNCTION constexpr std::enable_if_t<std::is_move_constructible<T>::value && std::is_move_assignable<T>::value> kokkos_swap(T& a, T& b) noexcept(
    std::is_nothrow_move_constructible<T>::value && std::is_nothrow_move_assignable<T>::value) {
  T t(std::move(a));
  a = std::move(b);
  b = std::move(t);
}
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr void kokkos_swap(T (&a)[N], T (&b)[N]) {
  for (size_t i = 0; i < N; ++i) kokkos_swap(a[i], b[i]);
}
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr void kokkos_swap(Array<T, N>& a, Array<T, N>& b) {
  if constexpr (N > 0)
    for (size_t i = 0; i < N; ++i) kokkos_swap(a.m_elems[i], b.m_elems[i]);
}

namespace Impl {
template <class T, size_t N, size_t... I>
KB200_FORCEINLINE_FUNCTION constexpr Array<std::remove_cv_t<T>, N> to_array_copy(T (&a)[N], std::index_sequence<I...>) { return {{a[I]...}}; }
template <class T, size_t N, size_t... I>
KB200_FORCEINLINE_FUNCTION constexpr Array<std::remove_cv_t<T>, N> to_array_move(T (&&a)[N], std::index_sequence<I...>) { return {{std::move(a[I])...}}; }
}  // namespace Impl
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr auto to_array(T (&a)[N]) { return Impl::to_array_copy(a, std::make_index_sequence<N>{}); }
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr auto to_array(T (&&a)[N]) { return Impl::to_array_move(std::move(a), std::make_index_sequence<N>{}); }

template <size_t I, class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr T& get(Array<T, N>& a) noexcept { static_assert(I < N, "kb200::get<I>(Array): index out of range"); return a.m_elems[I]; }
template <size_t I, class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr const T& get(const Array<T, N>& a) noexcept { static_assert(I < N, "kb200::get<I>(Array): index out of range"); return a.m_elems[I]; }
template <size_t I, class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr T&& get(Array<T, N>&& a) noexcept { return std::move(get<I>(a)); }

template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr T* begin(Array<T, N>& a) noexcept { return a.data(); }
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr const T* begin(const Array<T, N>& a) noexcept { return a.data(); }
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr T* end(Array<T, N>& a) noexcept { return a.data() + N; }
template <class T, size_t N>
KB200_FORCEINLINE_FUNCTION constexpr const T* end(const Array<T, N>& a) noexcept { return a.data() + N; }

}  // namespace kb200

// (in Kokkos-namespace mode `kb200` is a macro for `Kokkos`, so these specialise std::tuple_size<Kokkos::Array<...>>)
template <class T, std::size_t N>
struct std::tuple_size<kb200::Array<T, N>> : std::integral_constant<std::size_t, N> {};
template <std::size_t I, class T, std::size_t N>
struct std::tuple_element<I, kb200::Array<T, N>> { using type = T; };

#endif
