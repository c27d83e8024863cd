/* RAFT stand-in (see raft/core/error.hpp in this shim): only the two rounding helpers the reference uses. */
#pragma once
namespace raft {
template <typename S, typename T>
constexpr inline S div_rounding_up_unsafe(const S& dividend, const T& divisor) noexcept
{
  return (dividend + divisor - 1) / divisor;
}
template <typename I>
constexpr inline I div_rounding_up_safe(I dividend, I divisor) noexcept
{
  return dividend == 0 ? 0 : 1 + (dividend - 1) / divisor;
}
}  // namespace raft
