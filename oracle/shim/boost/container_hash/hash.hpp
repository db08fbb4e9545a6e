// Build shim for oracle/_ref only (test infrastructure, not shipped).
// The reference uses exactly one Boost symbol, boost::hash_combine
// (include/graphite/block.hpp:23-24, include/graphite/schur.hpp:73-75).
// Boost is absent from this image, so this provides that one function.
#pragma once
#include <cstddef>
#include <functional>
namespace boost {
template <class T> inline void hash_combine(std::size_t &seed, const T &v) {
  std::hash<T> h;
  seed ^= h(v) + 0x9e3779b97f4a7c15ULL + (seed << 6) + (seed >> 2);
}
} // namespace boost
