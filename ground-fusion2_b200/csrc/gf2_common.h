// gf2_common.h — error reporting shared by the translation units of libgf2_b200.so.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include "../../include/gf2_abi.h"

namespace gf2 {
inline char* err_buf() { static thread_local char buf[512] = {0}; return buf; }
inline const char* last_error() { return err_buf(); }
inline int fail(int code, const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_buf(), 512, fmt, ap); va_end(ap);
  return code;
}
}  // namespace gf2
