// Stand-in for the reference's src/utility/span.h (tcb::span polyfill) when the mirror classes are built OUTSIDE the reference
// tree: with C++20 tcb::span is simply std::span.  Inside the reference tree its own utility/span.h is found first instead.
#pragma once
#include <cstddef>
#include <span>
namespace tcb {
inline constexpr std::size_t dynamic_extent = std::dynamic_extent;
template <typename T, std::size_t Extent = std::dynamic_extent>
using span = std::span<T, Extent>;
}  // namespace tcb
