// kb200/Complex.hpp -- complex numbers usable in device code: the role of core/src/Kokkos_Complex.hpp:40-930
// (Kokkos::complex<T>: std::complex-compatible layout aligned to 2*sizeof(T) so that a complex<double> is one 16-byte
// word for vector loads and for the 128-bit compare-and-swap of Atomic.hpp; arithmetic with complex and real operands;
// abs/conj/exp/sqrt/pow/polar; stream I/O; reduction identities).
#ifndef KB200_COMPLEX_HPP
#define KB200_COMPLEX_HPP

#include "Macros.hpp"
#include <cmath>
#include <complex>
#include <iosfwd>
#include <istream>
#include <ostream>
#include <type_traits>
#include <utility>

namespace kb200 {

template <class T>
class alignas(2 * sizeof(T)) complex {
  static_assert(std::is_floating_point<T>::value && std::is_same<T, std::remove_cv_t<T>>::value, "kb200::complex needs a cv-unqualified floating-point type");
  T re_{};
  T im_{};

 public:
  using value_type = T;

  constexpr complex() = default;
  constexpr complex(const complex&) noexcept = default;
  constexpr complex& operator=(const complex&) noexcept = default;
  KB200_FORCEINLINE_FUNCTION constexpr complex(const T& re) noexcept : re_(re), im_(T()) {}
  KB200_FORCEINLINE_FUNCTION constexpr complex(const T& re, const T& im) noexcept : re_(re), im_(im) {}
  template <class U, class = std::enable_if_t<std::is_convertible<U, T>::value>>
  KB200_FORCEINLINE_FUNCTION constexpr complex(const complex<U>& o) noexcept : re_((T)o.real()), im_((T)o.imag()) {}
  complex(const std::complex<T>& s) noexcept : re_(s.real()), im_(s.imag()) {}
  operator std::complex<T>() const noexcept { return std::complex<T>(re_, im_); }
  complex& operator=(const std::complex<T>& s) noexcept { re_ = s.real(); im_ = s.imag(); return *this; }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator=(const T& re) noexcept { re_ = re; im_ = T(); return *this; }

  KB200_FORCEINLINE_FUNCTION constexpr T real() const noexcept { return re_; }
  KB200_FORCEINLINE_FUNCTION constexpr T imag() const noexcept { return im_; }
  KB200_FORCEINLINE_FUNCTION constexpr T& real() noexcept { return re_; }
  KB200_FORCEINLINE_FUNCTION constexpr T& imag() noexcept { return im_; }
  KB200_FORCEINLINE_FUNCTION constexpr void real(T v) noexcept { re_ = v; }
  KB200_FORCEINLINE_FUNCTION constexpr void imag(T v) noexcept { im_ = v; }

  KB200_FORCEINLINE_FUNCTION constexpr complex& operator+=(const complex& o) noexcept { re_ += o.re_; im_ += o.im_; return *this; }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator+=(const T& o) noexcept { re_ += o; return *this; }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator-=(const complex& o) noexcept { re_ -= o.re_; im_ -= o.im_; return *this; }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator-=(const T& o) noexcept { re_ -= o; return *this; }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator*=(const complex& o) noexcept {
    const T r = re_ * o.re_ - im_ * o.im_, i = re_ * o.im_ + im_ * o.re_;
    re_ = r; im_ = i;
    return *this;
  }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator*=(const T& o) noexcept { re_ *= o; im_ *= o; return *this; }
  // division scaled by the 1-norm of the divisor so that intermediate squares neither overflow nor underflow needlessly
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator/=(const complex& y) noexcept {
    const T s = (y.re_ < T(0) ? -y.re_ : y.re_) + (y.im_ < T(0) ? -y.im_ : y.im_);
    if (s == T(0)) {  // x / 0: IEEE semantics component-wise
      re_ /= s; im_ /= s;
    } else {
      const T yr = y.re_ / s, yi = y.im_ / s, d = yr * yr + yi * yi;
      const T xr = re_ / s, xi = im_ / s;
      re_ = (xr * yr + xi * yi) / d;
      im_ = (xi * yr - xr * yi) / d;
    }
    return *this;
  }
  KB200_FORCEINLINE_FUNCTION constexpr complex& operator/=(const T& o) noexcept { re_ /= o; im_ /= o; return *this; }

  template <size_t I>
  friend KB200_FORCEINLINE_FUNCTION constexpr const T& get(const complex& z) noexcept { static_assert(I < 2, ""); return I == 0 ? z.re_ : z.im_; }
  template <size_t I>
  friend KB200_FORCEINLINE_FUNCTION constexpr T& get(complex& z) noexcept { static_assert(I < 2, ""); return I == 0 ? z.re_ : z.im_; }
};

// ---- comparisons (mixed element types compare by value, as the reference's do)
template <class A, class B> KB200_FORCEINLINE_FUNCTION constexpr bool operator==(const complex<A>& x, const complex<B>& y) noexcept { return x.real() == y.real() && x.imag() == y.imag(); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<B>::value>>
KB200_FORCEINLINE_FUNCTION constexpr bool operator==(const complex<A>& x, const B& y) noexcept { return x.real() == y && x.imag() == A(0); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value>>
KB200_FORCEINLINE_FUNCTION constexpr bool operator==(const A& x, const complex<B>& y) noexcept { return y == x; }
template <class A, class B> bool operator==(const std::complex<A>& x, const complex<B>& y) noexcept { return x.real() == y.real() && x.imag() == y.imag(); }
template <class A, class B> bool operator==(const complex<A>& x, const std::complex<B>& y) noexcept { return y == x; }
template <class A, class B> KB200_FORCEINLINE_FUNCTION constexpr bool operator!=(const complex<A>& x, const complex<B>& y) noexcept { return !(x == y); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<B>::value>>
KB200_FORCEINLINE_FUNCTION constexpr bool operator!=(const complex<A>& x, const B& y) noexcept { return !(x == y); }
template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value>>
KB200_FORCEINLINE_FUNCTION constexpr bool operator!=(const A& x, const complex<B>& y) noexcept { return !(y == x); }
template <class A, class B> bool operator!=(const std::complex<A>& x, const complex<B>& y) noexcept { return !(x == y); }
template <class A, class B> bool operator!=(const complex<A>& x, const std::complex<B>& y) noexcept { return !(y == x); }

// ---- arithmetic; the result element type is the common type of the operands
#define KB200_COMPLEX_BINOP(OP)                                                                                                         \
  template <class A, class B>                                                                                                           \
  KB200_FORCEINLINE_FUNCTION constexpr complex<std::common_type_t<A, B>> operator OP(const complex<A>& x, const complex<B>& y) noexcept { \
    complex<std::common_type_t<A, B>> r(x);                                                                                             \
    r OP## = complex<std::common_type_t<A, B>>(y);                                                                                      \
    return r;                                                                                                                           \
  }                                                                                                                                     \
  template <class A, class B, class = std::enable_if_t<std::is_arithmetic<B>::value>>                                                   \
  KB200_FORCEINLINE_FUNCTION constexpr complex<std::common_type_t<A, B>> operator OP(const complex<A>& x, const B& y) noexcept {          \
    complex<std::common_type_t<A, B>> r(x);                                                                                             \
    r OP## = (std::common_type_t<A, B>)y;                                                                                               \
    return r;                                                                                                                           \
  }                                                                                                                                     \
  template <class A, class B, class = std::enable_if_t<std::is_arithmetic<A>::value>>                                                   \
  KB200_FORCEINLINE_FUNCTION constexpr complex<std::common_type_t<A, B>> operator OP(const A& x, const complex<B>& y) noexcept {          \
    complex<std::common_type_t<A, B>> r((std::common_type_t<A, B>)x);                                                                   \
    r OP## = complex<std::common_type_t<A, B>>(y);                                                                                      \
    return r;                                                                                                                           \
  }
KB200_COMPLEX_BINOP(+)
KB200_COMPLEX_BINOP(-)
KB200_COMPLEX_BINOP(*)
KB200_COMPLEX_BINOP(/)
#undef KB200_COMPLEX_BINOP
template <class T> KB200_FORCEINLINE_FUNCTION constexpr complex<T> operator+(const complex<T>& x) noexcept { return x; }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr complex<T> operator-(const complex<T>& x) noexcept { return complex<T>(-x.real(), -x.imag()); }

// ---- value functions
template <class T> KB200_FORCEINLINE_FUNCTION constexpr T real(const complex<T>& x) noexcept { return x.real(); }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr T imag(const complex<T>& x) noexcept { return x.imag(); }
template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>> KB200_FORCEINLINE_FUNCTION constexpr T real(const T& x) noexcept { return x; }
template <class T, class = std::enable_if_t<std::is_arithmetic<T>::value>> KB200_FORCEINLINE_FUNCTION constexpr T imag(const T&) noexcept { return T(0); }
template <class T> KB200_FORCEINLINE_FUNCTION constexpr complex<T> conj(const complex<T>& x) noexcept { return complex<T>(x.real(), -x.imag()); }
template <class T> KB200_FORCEINLINE_FUNCTION T abs(const complex<T>& x) { return (T)::hypot(x.real(), x.imag()); }
template <class T> KB200_FORCEINLINE_FUNCTION T arg(const complex<T>& x) { return (T)::atan2(x.imag(), x.real()); }
template <class T> KB200_FORCEINLINE_FUNCTION T norm(const complex<T>& x) { return x.real() * x.real() + x.imag() * x.imag(); }
template <class T> KB200_FORCEINLINE_FUNCTION complex<T> polar(const T& r, const T& theta = T()) { return complex<T>(r * (T)::cos(theta), r * (T)::sin(theta)); }
template <class T> KB200_FORCEINLINE_FUNCTION complex<T> exp(const complex<T>& x) { return polar((T)::exp(x.real()), x.imag()); }
template <class T> KB200_FORCEINLINE_FUNCTION complex<T> log(const complex<T>& x) { return complex<T>((T)::log(abs(x)), arg(x)); }
template <class T>
KB200_FORCEINLINE_FUNCTION complex<T> sqrt(const complex<T>& z) {
  const T x = z.real(), y = z.imag();
  if (x == T(0)) {
    const T t = (T)::sqrt((y < T(0) ? -y : y) / 2);
    return complex<T>(t, y < T(0) ? -t : t);
  }
  const T t = (T)::sqrt(2 * (abs(z) + (x < T(0) ? -x : x))), u = t / 2;
  return x > T(0) ? complex<T>(u, y / t) : complex<T>((y < T(0) ? -y : y) / t, y < T(0) ? -u : u);
}
template <class T> KB200_FORCEINLINE_FUNCTION complex<T> pow(const complex<T>& x, const T& y) { return x == T(0) ? (y == T(0) ? complex<T>(1) : complex<T>()) : polar((T)::pow(abs(x), y), y * arg(x)); }
template <class T> KB200_FORCEINLINE_FUNCTION complex<T> pow(const complex<T>& x, const complex<T>& y) { return x == T(0) ? (y == T(0) ? complex<T>(1) : complex<T>()) : exp(y * log(x)); }
template <class T> KB200_FORCEINLINE_FUNCTION complex<T> pow(const T& x, const complex<T>& y) { return pow(complex<T>(x), y); }

template <class T>
std::ostream& operator<<(std::ostream& os, const complex<T>& x) { return os << std::complex<T>(x); }
template <class T>
std::istream& operator>>(std::istream& is, complex<T>& x) {
  std::complex<T> s;
  is >> s;
  x = s;
  return is;
}

}  // namespace kb200

// structured bindings: auto [re, im] = z;
template <class T>
struct std::tuple_size<kb200::complex<T>> : std::integral_constant<std::size_t, 2> {};
template <std::size_t I, class T>
struct std::tuple_element<I, kb200::complex<T>> { using type = T; };

#endif
