// oracle/shims: STDLOG(level) << ... swallowed.
#pragma once
#include <iostream>
namespace orc_shim {
struct NullStream {
  template <typename T> NullStream& operator<<(const T&) { return *this; }
  NullStream& operator<<(std::ostream& (*)(std::ostream&)) { return *this; }
};
}  // namespace orc_shim
#define STDLOG(level) orc_shim::NullStream()
