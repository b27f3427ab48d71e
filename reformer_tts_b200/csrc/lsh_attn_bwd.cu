// placeholder: replaced by the real backward kernels
#include "host_util.h"
#include "rtts_b200.h"
using namespace rtts;
extern "C" int rtts_lsh_attn_bwd(const void*, const void*, int64_t, const int32_t*, const uint8_t*, const rtts_lsh_spec*,
                                 const void*, const float*, const float*, float*, float*, float*, int, int, int, int, int, int, void*) {
  return fail(kErrUnsupported, "rtts_lsh_attn_bwd: not built yet");
}
extern "C" int rtts_lsh_grad_reduce(const void*, int64_t, const float*, const float*, const float*, const rtts_lsh_spec*, void*,
                                    void*, int, int, int, int, int, void*) {
  return fail(kErrUnsupported, "rtts_lsh_grad_reduce: not built yet");
}
