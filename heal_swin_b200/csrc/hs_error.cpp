#include "hs_common.h"

namespace hs {

char* error_buffer() {
  static thread_local char buf[512] = "";
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace hs

extern "C" const char* hs_last_error(void) { return hs::error_buffer(); }
extern "C" int hs_version(void) { return 111; }
