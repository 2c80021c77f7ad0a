// stand-in for <dune/common/exceptions.hh>
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace Dune {
struct Exception : std::runtime_error
{
  Exception() : std::runtime_error("") {}
  void message(const std::string& m) { msg_ = m; }
  const char* what() const noexcept override { return msg_.c_str(); }
private:
  std::string msg_;
};
struct InvalidStateException : Exception {};
struct NotImplemented : Exception {};
template <typename T>
const T& resolveRef(const T& t) { return t; }
}  // namespace Dune
#define DUNE_THROW(E, m)        \
  do {                          \
    E th__ex;                   \
    std::ostringstream th__out; \
    th__out << m;               \
    th__ex.message(th__out.str()); \
    throw th__ex;               \
  } while (0)
