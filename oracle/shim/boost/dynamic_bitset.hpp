// Minimal functional stand-in for boost::dynamic_bitset<>, just large enough for the
// reference's src/needle.h (ctor(n, value) and operator[]) -- oracle/_ref only.
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <typename Block = unsigned long>
class dynamic_bitset {
 public:
  dynamic_bitset(std::size_t n, bool v) : bits_(n, v) {}
  std::vector<bool>::reference operator[](std::size_t i) { return bits_[i]; }
  bool operator[](std::size_t i) const { return bits_[i]; }
  std::size_t size() const { return bits_.size(); }

 private:
  std::vector<bool> bits_;
};
}  // namespace boost
