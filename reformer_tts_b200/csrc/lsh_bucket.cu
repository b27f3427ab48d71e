// LSH bucketing: random-rotation hash + argmax (rtts_lsh_hash) and the per-round stable counting
// sort with its inverse (rtts_lsh_sort).  Integer outputs are bit-exact with the reference's
// torch.argmax / torch.sort pipeline (rp R2-R3; hf:717-779); see include/rtts_b200.h.
#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"

namespace rtts {

// ------------------------------------------------------------------------------------------------
// Hash.  One thread per (batch, head, token): the 64-wide bf16 row sits in registers as fp32, the
// rotation matrix streams through shared memory in blocks of kProjBlock projections, and the
// argmax over cat([r, -r]) is tracked as (max r, first index) and (min r, first index): torch.argmax
// returns the FIRST maximum, and every +r precedes every -r, so a tie between the halves goes to +r.
// HBM-bound by design: reads 128 B per row once, writes R int32 per row.
// ------------------------------------------------------------------------------------------------
constexpr int kHashThreads = 128;
constexpr int kProjBlock = 64;
constexpr int kRotLd = 68;  // floats per projection row in smem (16 B aligned, skews banks)

template <int DH>
__global__ void __launch_bounds__(kHashThreads) lsh_hash_kernel(const __nv_bfloat16* __restrict__ qk, int64_t ld,
                                                                const float* __restrict__ rot, int rot_heads,
                                                                const uint8_t* __restrict__ pad_mask, int use_pad_bucket,
                                                                int32_t* __restrict__ buckets, float* __restrict__ sumsq, int T, int H,
                                                                int R, int n_buckets) {
  __shared__ __align__(16) float rot_s[kProjBlock * kRotLd];
  const int half = n_buckets >> 1;
  const int P = R * half;  // projections per vector
  const int t = blockIdx.x * kHashThreads + threadIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const bool live = t < T;

  float x[DH];
  if (live) {
    const uint4* src = reinterpret_cast<const uint4*>(qk + (static_cast<int64_t>(b) * T + t) * ld + h * DH);
#pragma unroll
    for (int c = 0; c < DH / 8; ++c) {
      uint4 u = __ldg(src + c);
      x[c * 8 + 0] = bf16_lo(u.x); x[c * 8 + 1] = bf16_hi(u.x);
      x[c * 8 + 2] = bf16_lo(u.y); x[c * 8 + 3] = bf16_hi(u.y);
      x[c * 8 + 4] = bf16_lo(u.z); x[c * 8 + 5] = bf16_hi(u.z);
      x[c * 8 + 6] = bf16_lo(u.w); x[c * 8 + 7] = bf16_hi(u.w);
    }
  } else {
#pragma unroll
    for (int k = 0; k < DH; ++k) x[k] = 0.f;
  }
  if (sumsq != nullptr && live) {     // |x|^2 of the row, reused by the attention kernels for the key normalisation
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < DH; ++k) ss = fmaf(x[k], x[k], ss);
    sumsq[(static_cast<int64_t>(b) * H + h) * T + t] = ss;
  }
  const float* rot_h = rot + static_cast<int64_t>(rot_heads == 1 ? 0 : h) * DH * P;
  const bool padded = use_pad_bucket && pad_mask != nullptr && live && pad_mask[static_cast<int64_t>(b) * T + t] == 0;
  const int stride = use_pad_bucket ? n_buckets + 1 : n_buckets;
  int32_t* out = buckets + (static_cast<int64_t>(b) * H + h) * R * T + t;

  float vmax = 0.f, vmin = 0.f;
  int imax = 0, imin = 0;
  for (int p0 = 0; p0 < P; p0 += kProjBlock) {
    const int pn = min(kProjBlock, P - p0);
    __syncthreads();
    // rot is [DH][P] per head: consecutive threads read consecutive projections of one k.
    for (int idx = threadIdx.x; idx < DH * pn; idx += kHashThreads) {
      const int k = idx / pn, pl = idx - k * pn;
      rot_s[pl * kRotLd + k] = __ldg(rot_h + static_cast<int64_t>(k) * P + p0 + pl);
    }
    __syncthreads();
    for (int pl = 0; pl < pn; ++pl) {
      const float4* rr = reinterpret_cast<const float4*>(rot_s + pl * kRotLd);
      float acc = 0.f;
#pragma unroll
      for (int k4 = 0; k4 < DH / 4; ++k4) {
        const float4 w = rr[k4];
        acc = fmaf(x[4 * k4 + 0], w.x, acc);
        acc = fmaf(x[4 * k4 + 1], w.y, acc);
        acc = fmaf(x[4 * k4 + 2], w.z, acc);
        acc = fmaf(x[4 * k4 + 3], w.w, acc);
      }
      const int p = p0 + pl;
      const int r = p / half, i = p - r * half;
      if (i == 0) { vmax = acc; vmin = acc; imax = 0; imin = 0; }
      else {
        if (acc > vmax) { vmax = acc; imax = i; }
        if (acc < vmin) { vmin = acc; imin = i; }
      }
      if (i == half - 1 && live) {
        int id = (vmax >= -vmin) ? imax : half + imin;
        if (padded) id = n_buckets;
        out[static_cast<int64_t>(r) * T] = r * stride + id;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sort.  One CTA per (row, round) segment of T keys.  Bucket ids within a round are < ids_per_round
// (<= kMaxIds), so a counting sort is exact: histogram -> exclusive scan -> stable scatter.  The
// scatter walks the segment in position order, 256 keys at a time; inside a tile warps take turns
// (8 barriers) and inside a warp __match_any_sync ranks equal keys by lane, which keeps equal
// buckets in ascending position = the order torch.sort gives for the unique key T*bucket + pos.
// ------------------------------------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kMaxIds = 1032;

__global__ void __launch_bounds__(kSortThreads) lsh_sort_kernel(const int32_t* __restrict__ buckets,
                                                                int32_t* __restrict__ sticker, int32_t* __restrict__ undo,
                                                                int T, int R, int ids_per_round) {
  __shared__ int32_t cursor[kMaxIds];
  __shared__ int32_t warp_tot[kSortThreads / 32];
  const int r = blockIdx.x, row = blockIdx.y;
  const int64_t seg = (static_cast<int64_t>(row) * R + r) * T;
  const int32_t* key = buckets + seg;
  const int base_id = r * ids_per_round;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  for (int i = tid; i < ids_per_round; i += kSortThreads) cursor[i] = 0;
  __syncthreads();
  for (int i = tid; i < T; i += kSortThreads) atomicAdd(&cursor[key[i] - base_id], 1);
  __syncthreads();
  // exclusive scan of cursor[0..ids_per_round) by one warp-synchronous block pass
  {
    int carry = 0;
    for (int i0 = 0; i0 < ids_per_round; i0 += kSortThreads) {
      const int i = i0 + tid;
      const int v = i < ids_per_round ? cursor[i] : 0;
      int s = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += n;
      }
      if (lane == 31) warp_tot[warp] = s;
      __syncthreads();
      int woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < kSortThreads / 32; ++w) {
        if (w < warp) woff += warp_tot[w];
        tot += warp_tot[w];
      }
      if (i < ids_per_round) cursor[i] = carry + woff + s - v;
      carry += tot;
      __syncthreads();
    }
  }
  for (int i0 = 0; i0 < T; i0 += kSortThreads) {
    const int i = i0 + tid;
    const bool live = i < T;
    const int id = live ? key[i] - base_id : -1 - lane;  // dead lanes never match a live key
    const unsigned peers = __match_any_sync(0xffffffffu, id);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const bool leader = rank == 0;
    for (int w = 0; w < kSortThreads / 32; ++w) {
      if (w == warp && live) {
        const int slot = cursor[id] + rank;
        sticker[seg + slot] = r * T + i;
        undo[seg + i] = r * T + slot;
      }
      __syncwarp();
      if (w == warp && live && leader) cursor[id] += __popc(peers);
      __syncthreads();
    }
  }
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_lsh_hash(const void* qk, int64_t ld, const float* rot, int rot_heads, const uint8_t* pad_mask,
                             int use_pad_bucket, int32_t* buckets, float* sumsq, int B, int T, int H, int dh, int R,
                             int n_buckets, void* stream) {
  RTTS_REQUIRE(qk && rot && buckets, "rtts_lsh_hash: null pointer");
  RTTS_REQUIRE(dh == 64, "rtts_lsh_hash: head size %d unsupported (64 only)", dh);
  RTTS_REQUIRE(n_buckets >= 2 && n_buckets % 2 == 0, "rtts_lsh_hash: n_buckets must be even, got %d", n_buckets);
  RTTS_REQUIRE(rot_heads == 1 || rot_heads == H, "rtts_lsh_hash: rot_heads must be 1 or H");
  RTTS_REQUIRE(ld % 8 == 0 && (reinterpret_cast<uintptr_t>(qk) & 15) == 0, "rtts_lsh_hash: qk must be 16-byte aligned");
  RTTS_REQUIRE(B > 0 && T > 0 && H > 0 && R > 0 && B < 65536 && H < 65536, "rtts_lsh_hash: bad sizes");
  dim3 grid((T + kHashThreads - 1) / kHashThreads, H, B);
  lsh_hash_kernel<64><<<grid, kHashThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qk), ld, rot, rot_heads, pad_mask, use_pad_bucket, buckets, sumsq, T, H, R, n_buckets);
  return check_launch("rtts_lsh_hash");
}

extern "C" int rtts_lsh_sort(const int32_t* buckets, int32_t* sticker, int32_t* undo, int rows, int T, int R,
                             int ids_per_round, void* stream) {
  RTTS_REQUIRE(buckets && sticker && undo, "rtts_lsh_sort: null pointer");
  RTTS_REQUIRE(ids_per_round > 0 && ids_per_round <= kMaxIds, "rtts_lsh_sort: ids_per_round %d > %d", ids_per_round,
               kMaxIds);
  RTTS_REQUIRE(rows > 0 && rows < 65536 && T > 0 && R > 0, "rtts_lsh_sort: bad sizes");
  dim3 grid(R, rows);
  lsh_sort_kernel<<<grid, kSortThreads, 0, static_cast<cudaStream_t>(stream)>>>(buckets, sticker, undo, T, R,
                                                                               ids_per_round);
  return check_launch("rtts_lsh_sort");
}
