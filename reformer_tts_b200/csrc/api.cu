// Library-level entry points: ABI version and the thread-local error string.
#include "host_util.h"
#include "rtts_b200.h"

namespace rtts {
char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace rtts

extern "C" const char* rtts_last_error(void) { return rtts::error_buffer(); }
extern "C" int rtts_abi_version(void) { return 1; }
