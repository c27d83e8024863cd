/* Minimal stand-in for RAFT's error header so the reference's gather/scatter/memory/comm sources build
 * without the (un-vendored, network-fetched) RAFT dependency.  TEST INFRASTRUCTURE (oracle/_ref build). */
#pragma once
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <functional>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

namespace raft {
class exception : public std::exception {
 public:
  explicit exception() noexcept : msg_() {}
  explicit exception(char const* const message) noexcept : msg_(message) {}
  explicit exception(std::string const& message) noexcept : msg_(message) {}
  char const* what() const noexcept override { return msg_.c_str(); }

 private:
  std::string msg_;
};
struct logic_error : public exception {
  explicit logic_error(char const* const message) : exception(message) {}
  explicit logic_error(std::string const& message) : exception(message) {}
};
}  // namespace raft
