// Parameters shared by the forward attention kernels (lsh_attn_fwd.cu: bucket 128 and the generic bucket-64 kernel;
// lsh_attn_fwd64.cu: the block-streaming bucket-64 kernel).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace rtts {

constexpr int kPadFlag = 0x40000000;      // position entry of a padded token: pos | kPadFlag

struct AttnFwdParams {
  const __nv_bfloat16* qk;
  const __nv_bfloat16* v;
  int64_t ld;
  const int32_t* sticker;
  const float* sumsq;   // [B,H,T] |qk row|^2
  const uint8_t* mask;
  __nv_bfloat16* o_rounds;
  float* lse_rounds;
  long long* trace;     // debug: per-role clock64 stamps of CTA 0 (nullable)
  int T, H, R;
  int tiles_per_row;  // R*T / 128
  float score_scale_log2;  // score_scale * log2(e)
  float mask_value_log2, self_value_log2;
  int key_norm, mask_mode, causal;
  int pos16;          // T <= 2048: positions are exact in fp16, the position mask is evaluated two keys per instruction
};

// lsh_attn_fwd64.cu (block-streaming kernel) / lsh_attn_fwd64p.cu (paired-chunk kernel: the default for bucket 64, T <= 2048)
int launch_attn_fwd64(const AttnFwdParams& p, int B, cudaStream_t stream);
int launch_attn_fwd64p(const AttnFwdParams& p, int B, cudaStream_t stream);

}  // namespace rtts
