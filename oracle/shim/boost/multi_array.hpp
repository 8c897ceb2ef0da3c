// Minimal functional stand-in for boost::multi_array<T,2>, just large enough for the
// reference's src/needle.h and src/align.h to compile unmodified (oracle/_ref only).
// Supports: default ctor, resize(boost::extents[a][b]), shape(), operator[][] and the
// nested `index` typedef.  Test infrastructure -- never linked into the product.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
namespace detail_shim {
struct extent2 { std::size_t a, b; };
struct extent1 { std::size_t a; extent2 operator[](std::size_t b) const { return extent2{a, b}; } };
struct extent_gen { extent1 operator[](std::size_t a) const { return extent1{a}; } };
}  // namespace detail_shim
static const detail_shim::extent_gen extents = detail_shim::extent_gen();

template <typename T, std::size_t NDims>
class multi_array;

template <typename T>
class multi_array<T, 2> {
 public:
  typedef std::ptrdiff_t index;
  typedef std::size_t size_type;
  multi_array() { shape_[0] = shape_[1] = 0; }
  void resize(detail_shim::extent2 const& e) {
    shape_[0] = e.a;
    shape_[1] = e.b;
    data_.assign(e.a * e.b, T());
  }
  const size_type* shape() const { return shape_; }
  T* operator[](index i) { return data_.data() + static_cast<std::size_t>(i) * shape_[1]; }
  const T* operator[](index i) const { return data_.data() + static_cast<std::size_t>(i) * shape_[1]; }

 private:
  size_type shape_[2];
  std::vector<T> data_;
};
}  // namespace boost
