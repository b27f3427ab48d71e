// Row-wise HBM-bound kernels around the GEMMs: LayerNorm forward / backward
// (ref:reformer_tts/model/reformer.py:25-33: nn.LayerNorm(dim), eps 1e-5, affine), the fp32->bf16 cast with
// bias-gradient column sums, and the attention-backward delta.  One warp per row, 128-bit accesses.
#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"

namespace rtts {

constexpr int kRowThreads = 256;   // 8 rows per CTA
constexpr int kMaxDim = 1024;      // dim / 128 float4 per lane <= 8

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x fp32 [rows, dim] -> y bf16.  Two-pass variance on the register-resident row (matches torch's
// mean / biased variance to fp32 rounding).
template <int VEC>  // float4 per lane = dim / 128
__global__ void __launch_bounds__(kRowThreads) layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                                                    float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                                    int rows, float eps) {
  constexpr int dim = VEC * 128;
  const int row = blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * dim);
  float4 v[VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    v[i] = __ldg(xr + i * 32 + lane);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.f / dim);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / dim) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<int64_t>(row) * dim);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + i * 32 + lane);
    uint2 o;
    o.x = pack_bf16((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
    o.y = pack_bf16((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
    yr[i * 32 + lane] = o;
  }
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)); dgamma += sum dy*xhat; dbeta += sum dy.
// Parameter gradients: per-CTA partial sums in shared memory, then one atomicAdd per column per CTA.
template <int VEC>
__global__ void __launch_bounds__(kRowThreads) layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                    const float* __restrict__ gamma, const float* __restrict__ mean,
                                                                    const float* __restrict__ rstd, float* dx,
                                                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int rows,
                                                                    int rows_per_cta, const float* dx_add) {
  constexpr int dim = VEC * 128;
  __shared__ float sg[dim], sb[dim];
  for (int i = threadIdx.x; i < dim; i += kRowThreads) { sg[i] = 0.f; sb[i] = 0.f; }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 ag[VEC], ab[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) { ag[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); }
  const int row_begin = blockIdx.x * rows_per_cta;
  const int row_end = min(rows, row_begin + rows_per_cta);
  for (int row = row_begin + warp; row < row_end; row += kRowThreads / 32) {
    const float4* xr = reinterpret_cast<const float4*>(x + static_cast<int64_t>(row) * dim);
    const float4* dr = reinterpret_cast<const float4*>(dy + static_cast<int64_t>(row) * dim);
    const float mu = __ldg(mean + row), rs = __ldg(rstd + row);
    float4 xh[VEC], gd[VEC], av[VEC];
    // dx_add (nullable, may alias dx): the gradient this one is accumulated onto (dx2 += df of the reversible blocks) - its row is
    // requested with the other loads and added in the store, instead of a separate pass over both tensors
    if (dx_add != nullptr) {
#pragma unroll
      for (int i = 0; i < VEC; ++i) av[i] = *(reinterpret_cast<const float4*>(dx_add + static_cast<int64_t>(row) * dim) + i * 32 + lane);
    } else {
#pragma unroll
      for (int i = 0; i < VEC; ++i) av[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      const float4 xv = __ldg(xr + i * 32 + lane), dv = __ldg(dr + i * 32 + lane);
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i * 32 + lane);
      xh[i] = make_float4((xv.x - mu) * rs, (xv.y - mu) * rs, (xv.z - mu) * rs, (xv.w - mu) * rs);
      gd[i] = make_float4(dv.x * g.x, dv.y * g.y, dv.z * g.z, dv.w * g.w);
      s1 += gd[i].x + gd[i].y + gd[i].z + gd[i].w;
      s2 += gd[i].x * xh[i].x + gd[i].y * xh[i].y + gd[i].z * xh[i].z + gd[i].w * xh[i].w;
      ag[i].x += dv.x * xh[i].x; ag[i].y += dv.y * xh[i].y; ag[i].z += dv.z * xh[i].z; ag[i].w += dv.w * xh[i].w;
      ab[i].x += dv.x; ab[i].y += dv.y; ab[i].z += dv.z; ab[i].w += dv.w;
    }
    s1 = warp_sum(s1) * (1.f / dim);
    s2 = warp_sum(s2) * (1.f / dim);
    float4* out = reinterpret_cast<float4*>(dx + static_cast<int64_t>(row) * dim);
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
      out[i * 32 + lane] = make_float4(av[i].x + rs * (gd[i].x - s1 - xh[i].x * s2), av[i].y + rs * (gd[i].y - s1 - xh[i].y * s2),
                                       av[i].z + rs * (gd[i].z - s1 - xh[i].z * s2), av[i].w + rs * (gd[i].w - s1 - xh[i].w * s2));
    }
  }
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int c = (i * 32 + lane) * 4;
    atomicAdd(&sg[c], ag[i].x); atomicAdd(&sg[c + 1], ag[i].y); atomicAdd(&sg[c + 2], ag[i].z); atomicAdd(&sg[c + 3], ag[i].w);
    atomicAdd(&sb[c], ab[i].x); atomicAdd(&sb[c + 1], ab[i].y); atomicAdd(&sb[c + 2], ab[i].z); atomicAdd(&sb[c + 3], ab[i].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dim; i += kRowThreads) {
    atomicAdd(dgamma + i, sg[i]);
    atomicAdd(dbeta + i, sb[i]);
  }
}

// y = bf16(x); colsum[c] += sum_rows x[r,c]  (colsum may be NULL).  A CTA (256 threads) owns a strip of up to 1024 columns:
// `tpr` threads (float4 each) cover one row of the strip and 256 / tpr rows are in flight per pass, eight passes of loads are
// issued before any is consumed.  Column sums: registers -> shared-memory reduction over the row lanes -> ONE atomic per column
// per CTA, and the grid is only ~4 CTAs per SM, so an address sees a few hundred atomics instead of thousands.
// KEEP: inverted dropout first (x * keep * scale; keep = uint8 mask of x's shape), the backward of a fused dropout epilogue.
constexpr int kCastThreads = 512;
template <bool KEEP>
__global__ void __launch_bounds__(kCastThreads) cast_bf16_colsum_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                               float* __restrict__ colsum, int rows, int cols, int rows_per_cta, int tpr,
                                                               const uint8_t* __restrict__ keep, float keep_scale) {
  auto masked = [&](float4 v, int64_t off) {
    if (KEEP) {
      const uint32_t kw = __ldg(reinterpret_cast<const uint32_t*>(keep + off));
      v.x = (kw & 0x000000ffu) ? v.x * keep_scale : 0.f; v.y = (kw & 0x0000ff00u) ? v.y * keep_scale : 0.f;
      v.z = (kw & 0x00ff0000u) ? v.z * keep_scale : 0.f; v.w = (kw & 0xff000000u) ? v.w * keep_scale : 0.f;
    }
    return v;
  };
  __shared__ float4 red[kCastThreads];
  const int lane_row = threadIdx.x / tpr, lanes = kCastThreads / tpr;
  const int c = (blockIdx.x * tpr + threadIdx.x % tpr) * 4;
  const bool live = c < cols;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float4 acc = make_float4(0, 0, 0, 0);
  if (live) {
    int r = r0 + lane_row;
    for (; r + 7 * lanes < r1; r += 8 * lanes) {       // eight independent 16-byte loads in flight per thread
      float4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(r + i * lanes) * cols + c));
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = masked(v[i], static_cast<int64_t>(r + i * lanes) * cols + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint2 o;
        o.x = pack_bf16(v[i].x, v[i].y);
        o.y = pack_bf16(v[i].z, v[i].w);
        *reinterpret_cast<uint2*>(y + static_cast<int64_t>(r + i * lanes) * cols + c) = o;
        acc.x += v[i].x; acc.y += v[i].y; acc.z += v[i].z; acc.w += v[i].w;
      }
    }
    for (; r < r1; r += lanes) {
      const float4 v = masked(__ldg(reinterpret_cast<const float4*>(x + static_cast<int64_t>(r) * cols + c)), static_cast<int64_t>(r) * cols + c);
      uint2 o;
      o.x = pack_bf16(v.x, v.y);
      o.y = pack_bf16(v.z, v.w);
      *reinterpret_cast<uint2*>(y + static_cast<int64_t>(r) * cols + c) = o;
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (colsum == nullptr) return;
  red[threadIdx.x] = acc;
  __syncthreads();
  if (lane_row == 0 && live) {
    for (int l = 1; l < lanes; ++l) {
      const float4 o = red[l * tpr + threadIdx.x];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    // One 16-byte vector reduction per thread where the target allows it.  ncu (profiles/r2_hbm_kernels_ncu_full.txt) had this kernel
    // at 15 % of DRAM peak with 7 % of the issue slots busy: 592 CTAs x 4 scalar atomics on each of the same 512 addresses
    // serialise in L2, and the CTAs sit at the barrier above waiting for them.
    if ((reinterpret_cast<uintptr_t>(colsum + c) & 15) == 0)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(acc.x), "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
    else {
      atomicAdd(colsum + c, acc.x); atomicAdd(colsum + c + 1, acc.y);
      atomicAdd(colsum + c + 2, acc.z); atomicAdd(colsum + c + 3, acc.w);
    }
  }
}

// colsum[c] += sum_rows x[r,c] for a bf16 matrix (bias gradients of gradient matrices that are already bf16): the same strip
// scheme as above - `tpr` threads (8 bf16 each) cover one row of the strip, eight 16-byte loads in flight per thread, shared-memory
// reduction over the row lanes, two 16-byte vector reductions per thread and CTA.
__global__ void __launch_bounds__(kCastThreads) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, float* __restrict__ colsum, int rows,
                                                                    int cols, int rows_per_cta, int tpr) {
  __shared__ float red[kCastThreads][9];
  const int lane_row = threadIdx.x / tpr, lanes = kCastThreads / tpr;
  const int c = (blockIdx.x * tpr + threadIdx.x % tpr) * 8;
  const bool live = c < cols;
  const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add = [&](const uint4& u) {
    acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
    acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
  };
  if (live) {
    int r = r0 + lane_row;
    for (; r + 7 * lanes < r1; r += 8 * lanes) {
      uint4 v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r + i * lanes) * ld + c));
#pragma unroll
      for (int i = 0; i < 8; ++i) add(v[i]);
    }
    for (; r < r1; r += lanes) add(__ldg(reinterpret_cast<const uint4*>(x + static_cast<int64_t>(r) * ld + c)));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = acc[i];
  __syncthreads();
  if (lane_row == 0 && live) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += red[l * tpr + threadIdx.x][i];
    if ((reinterpret_cast<uintptr_t>(colsum + c) & 15) == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c), "f"(acc[0]), "f"(acc[1]), "f"(acc[2]), "f"(acc[3]) : "memory");
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum + c + 4), "f"(acc[4]), "f"(acc[5]), "f"(acc[6]), "f"(acc[7]) : "memory");
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(colsum + c + i, acc[i]);
    }
  }
}

// delta[(b*H+h)*T + t] = <dout[b,t,h,:], out[b,t,h,:]>, bf16 inputs, 8 lanes per 64-wide head slice.
__global__ void __launch_bounds__(256) lsh_delta_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ out,
                                                        int64_t ld, float* __restrict__ delta, int T, int H, int64_t rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 3;   // (b*T + t)*H + h : coalesced reads
  const int c = threadIdx.x & 7;
  float s = 0.f;
  int64_t b = 0; int t = 0, h = 0;
  if (row < rows) {
    const int64_t bt = row / H;
    h = static_cast<int>(row - bt * H);
    b = bt / T;
    t = static_cast<int>(bt - b * T);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(dout + bt * ld + h * 64) + c);
    const uint4 o = __ldg(reinterpret_cast<const uint4*>(out + bt * ld + h * 64) + c);
    s = bf16_lo(a.x) * bf16_lo(o.x) + bf16_hi(a.x) * bf16_hi(o.x) + bf16_lo(a.y) * bf16_lo(o.y) + bf16_hi(a.y) * bf16_hi(o.y) +
        bf16_lo(a.z) * bf16_lo(o.z) + bf16_hi(a.z) * bf16_hi(o.z) + bf16_lo(a.w) * bf16_lo(o.w) + bf16_hi(a.w) * bf16_hi(o.w);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (row < rows && c == 0) delta[(b * H + h) * T + t] = s;
}

// sumsq[(b*H+h)*T + t] = |qk[b,t,h,:]|^2 (fp32), 8 lanes per 64-wide head slice.  Same quantity rtts_lsh_hash can emit.
__global__ void __launch_bounds__(256) lsh_sumsq_kernel(const __nv_bfloat16* __restrict__ qk, int64_t ld, float* __restrict__ sumsq, int T, int H,
                                                        int64_t rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 3;   // (b*T + t)*H + h
  const int c = threadIdx.x & 7;
  float s = 0.f;
  int64_t b = 0; int t = 0, h = 0;
  if (row < rows) {
    const int64_t bt = row / H;
    h = static_cast<int>(row - bt * H);
    b = bt / T;
    t = static_cast<int>(bt - b * T);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(qk + bt * ld + h * 64) + c);
    const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) s += bf16_lo(w[e]) * bf16_lo(w[e]) + bf16_hi(w[e]) * bf16_hi(w[e]);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (row < rows && c == 0) sumsq[(b * H + h) * T + t] = s;
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_layernorm_fwd(const float* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                                  int rows, int dim, float eps, void* stream) {
  RTTS_REQUIRE(x && gamma && beta && y, "rtts_layernorm_fwd: null pointer");
  RTTS_REQUIRE(rows > 0 && dim % 128 == 0 && dim <= kMaxDim, "rtts_layernorm_fwd: dim=%d must be a multiple of 128 and <= %d", dim, kMaxDim);
  const int blocks = (rows + kRowThreads / 32 - 1) / (kRowThreads / 32);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* yb = static_cast<__nv_bfloat16*>(y);
  switch (dim / 128) {
    case 1: layernorm_fwd_kernel<1><<<blocks, kRowThreads, 0, s>>>(x, gamma, beta, yb, mean, rstd, rows, eps); break;
    case 2: layernorm_fwd_kernel<2><<<blocks, kRowThreads, 0, s>>>(x, gamma, beta, yb, mean, rstd, rows, eps); break;
    case 4: layernorm_fwd_kernel<4><<<blocks, kRowThreads, 0, s>>>(x, gamma, beta, yb, mean, rstd, rows, eps); break;
    case 8: layernorm_fwd_kernel<8><<<blocks, kRowThreads, 0, s>>>(x, gamma, beta, yb, mean, rstd, rows, eps); break;
    default: return fail(kErrUnsupported, "rtts_layernorm_fwd: dim=%d unsupported (128, 256, 512, 1024)", dim);
  }
  return check_launch("rtts_layernorm_fwd");
}

extern "C" int rtts_layernorm_bwd(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                  float* dx, float* dgamma, float* dbeta, int rows, int dim, void* stream) {
  return rtts_layernorm_bwd_acc(dy, x, gamma, mean, rstd, nullptr, dx, dgamma, dbeta, rows, dim, stream);
}

extern "C" int rtts_layernorm_bwd_acc(const float* dy, const float* x, const float* gamma, const float* mean, const float* rstd,
                                      const float* dx_add, float* dx, float* dgamma, float* dbeta, int rows, int dim, void* stream) {
  RTTS_REQUIRE(dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "rtts_layernorm_bwd: null pointer");
  RTTS_REQUIRE(rows > 0 && dim % 128 == 0 && dim <= kMaxDim, "rtts_layernorm_bwd: dim=%d must be a multiple of 128 and <= %d", dim, kMaxDim);
  // ~2 CTAs per SM; each walks a contiguous slab of rows so parameter-gradient atomics stay at 2*dim per CTA
  const int ctas = min(rows / 8 > 0 ? rows / 8 : 1, 2 * kNumSMs);
  const int rows_per_cta = (rows + ctas - 1) / ctas;
  const int blocks = (rows + rows_per_cta - 1) / rows_per_cta;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  switch (dim / 128) {
    case 1: layernorm_bwd_kernel<1><<<blocks, kRowThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, rows_per_cta, dx_add); break;
    case 2: layernorm_bwd_kernel<2><<<blocks, kRowThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, rows_per_cta, dx_add); break;
    case 4: layernorm_bwd_kernel<4><<<blocks, kRowThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, rows_per_cta, dx_add); break;
    case 8: layernorm_bwd_kernel<8><<<blocks, kRowThreads, 0, s>>>(dy, x, gamma, mean, rstd, dx, dgamma, dbeta, rows, rows_per_cta, dx_add); break;
    default: return fail(kErrUnsupported, "rtts_layernorm_bwd: dim=%d unsupported (128, 256, 512, 1024)", dim);
  }
  return check_launch("rtts_layernorm_bwd");
}

extern "C" int rtts_colsum_bf16(const void* x, int64_t ld, float* colsum, int rows, int cols, void* stream) {
  RTTS_REQUIRE(x && colsum, "rtts_colsum_bf16: null pointer");
  RTTS_REQUIRE(rows > 0 && cols > 0 && cols % 8 == 0 && ld % 8 == 0 && ld >= cols, "rtts_colsum_bf16: cols=%d and ld must be multiples of 8", cols);
  RTTS_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "rtts_colsum_bf16: x must be 16-byte aligned");
  const int vec = cols / 8;                                   // 16-byte pieces per row
  int tpr = 1;
  while (tpr < vec && tpr < 128) tpr *= 2;                    // threads per row of a strip (power of two <= 128: at most 1024 columns per CTA)
  const int strips = (vec + tpr - 1) / tpr;
  int ctas_y = (4 * kNumSMs + strips - 1) / strips;           // ~4 CTAs per SM
  const int lanes = kCastThreads / tpr;
  const int max_y = (rows + 8 * lanes - 1) / (8 * lanes);
  if (ctas_y > max_y) ctas_y = max_y;
  if (ctas_y < 1) ctas_y = 1;
  const int rows_per_cta = (rows + ctas_y - 1) / ctas_y;
  colsum_bf16_kernel<<<dim3(strips, ctas_y), kCastThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(x), ld, colsum, rows, cols,
                                                                                                   rows_per_cta, tpr);
  return check_launch("rtts_colsum_bf16");
}

extern "C" int rtts_cast_bf16_colsum(const float* x, void* y, float* colsum, int rows, int cols, void* stream) {
  return rtts_cast_bf16_colsum_dropout(x, nullptr, 1.f, y, colsum, rows, cols, stream);
}

extern "C" int rtts_cast_bf16_colsum_dropout(const float* x, const uint8_t* keep_mask, float keep_scale, void* y, float* colsum, int rows,
                                             int cols, void* stream) {
  RTTS_REQUIRE(x && y, "rtts_cast_bf16_colsum: null pointer");
  RTTS_REQUIRE(!keep_mask || (reinterpret_cast<uintptr_t>(keep_mask) & 3) == 0, "rtts_cast_bf16_colsum_dropout: keep mask must be 4-byte aligned");
  RTTS_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "rtts_cast_bf16_colsum: cols=%d must be a multiple of 4", cols);
  int tpr = 256;                                  // threads per row of a strip: a power of two covering min(cols, 1024) columns
  while (tpr > 1 && (tpr / 2) * 4 >= cols) tpr /= 2;
  const int strips = (cols / 4 + tpr - 1) / tpr;
  int row_ctas = (2 * kNumSMs + strips - 1) / strips;       // two fat CTAs per SM: half the atomics per address of four thin ones
  if (row_ctas > rows) row_ctas = rows;
  const int rows_per_cta = (rows + row_ctas - 1) / row_ctas;
  dim3 grid(strips, (rows + rows_per_cta - 1) / rows_per_cta);
  if (keep_mask != nullptr)
    cast_bf16_colsum_kernel<true><<<grid, kCastThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y), colsum, rows, cols,
                                                                                      rows_per_cta, tpr, keep_mask, keep_scale);
  else
    cast_bf16_colsum_kernel<false><<<grid, kCastThreads, 0, static_cast<cudaStream_t>(stream)>>>(x, static_cast<__nv_bfloat16*>(y), colsum, rows, cols,
                                                                                       rows_per_cta, tpr, nullptr, 1.f);
  return check_launch("rtts_cast_bf16_colsum");
}

extern "C" int rtts_lsh_delta(const void* dout, const void* out, int64_t ld, float* delta, int B, int T, int H, int dh,
                              void* stream) {
  RTTS_REQUIRE(dout && out && delta, "rtts_lsh_delta: null pointer");
  RTTS_REQUIRE(dh == 64 && ld % 8 == 0, "rtts_lsh_delta: head size 64 and 16-byte rows required");
  const int64_t rows = static_cast<int64_t>(B) * T * H;
  const int64_t blocks = (rows * 8 + 255) / 256;
  lsh_delta_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dout), static_cast<const __nv_bfloat16*>(out), ld, delta, T, H, rows);
  return check_launch("rtts_lsh_delta");
}

extern "C" int rtts_lsh_sumsq(const void* qk, int64_t ld, float* sumsq, int B, int T, int H, int dh, void* stream) {
  RTTS_REQUIRE(qk && sumsq, "rtts_lsh_sumsq: null pointer");
  RTTS_REQUIRE(dh == 64 && ld % 8 == 0, "rtts_lsh_sumsq: head size 64 and 16-byte rows required");
  const int64_t rows = static_cast<int64_t>(B) * T * H;
  const int64_t blocks = (rows * 8 + 255) / 256;
  lsh_sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const __nv_bfloat16*>(qk), ld, sumsq, T, H,
                                                                                                  rows);
  return check_launch("rtts_lsh_sumsq");
}
