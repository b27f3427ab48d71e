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
// sha256 (first 16 hex digits) over the sources the library was compiled from: csrc/*.cu, *.cuh, *.h and include/rtts_b200.h,
// computed by csrc/build.py (source_hash) and passed as -DRTTS_BUILD_ID.  tests/test_host.py compares it with a hash of the
// checked-out sources, so a stale prebuilt .so cannot pass for the current code.
#ifndef RTTS_BUILD_ID
#define RTTS_BUILD_ID "unknown"
#endif
extern "C" const char* rtts_build_id(void) { return RTTS_BUILD_ID; }
