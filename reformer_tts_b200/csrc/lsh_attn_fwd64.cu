// Fused chunked shared-QK attention, forward, bucket size 64: the block-streaming kernel behind rtts_lsh_attn_fwd.
//
// The sorted slots of one (batch, head) row form a cyclic sequence of 64-slot chunks (rp R6: every chunk attends to itself
// and to the chunk before it).  The unit of work here is a BLOCK = the 64 keys of one chunk c against the 128 queries that
// can see them, the chunk itself (lanes 0-63 of the MMA) and its successor c+1 (lanes 64-127):
//     S_c  = [X_c ; X_c+1] X_c^T      tcgen05.mma M=128, N=64, K=64: every one of the 128x64 scores is used
//     P_c  = exp2(S_c * key_scale - bound[query])   (masks applied on the packed bf16 pairs), written back over S_c in TMEM
//     D_c  = P_c [V_c | 1]            tcgen05.mma, A = P from TMEM, N = 64 (+16 columns of row sums)
// so lanes 0-63 of D_c hold the "own chunk" part of chunk c's output and lanes 64-127 the look-back part of chunk c+1's.
// (The tile form - 128 queries against 192 keys - spends a third of the tensor and TMEM work on key blocks a query cannot
// see; this form has none.)  The two parts of a chunk's output come out of consecutive blocks on different TMEM lanes: the
// warps that own lanes 64-127 pass their part through shared memory (fp32) to the warps that own lanes 0-63, which add the
// two, normalise, and store.  The A operand of S_c is 128 CONSECUTIVE rows of the shared-memory ring of gathered chunks
// (chunk c then chunk c+1), which is what fixes "successor on the upper lanes".
//
// Softmax is single-pass as in lsh_attn_fwd.cu: keys are unit vectors after normalisation, so |q| * scale bounds every score
// and is known before any score is read; both parts of a row use the same bound, so they add without rescaling.  A chunk with
// a bound >= 60 (a visible key could underflow against it) is flagged by the loader and runs the reference's arithmetic per
// block (row maximum of the block first, fill values for masked / self entries); the combining warps then rescale the two
// parts by their maxima.  Thread = one query row of a block = all 64 keys, so neither mode needs any cross-thread exchange.
//
// Pipeline per CTA (persistent, one CTA per SM, a contiguous run of chunks; entry e of the run lives in ring slot e % 8):
//   4 loader warps   warp w gathers entries e = w (mod 4): sticker -> position -> 16-byte cp.async of the qk and v rows into
//                    SWIZZLE_128B chunks, per-row metadata (key scale, -bound, positions as int / fp16, query limit)
//   S issuer         one elected thread: S_e into TMEM buffer e % 4 as soon as entries e, e+1 have landed
//   G softmax groups 4 warps each (thread = TMEM lane = query row), group g takes blocks e = g (mod G)
//   PV issuer        D_e = P_e [V_e | 1] into TMEM buffer e % 3, one commit that frees S buffer / publishes D
//   2 x 4 epilogue warps (set s takes entries e = s mod 2)  lanes 64-127: look-back part -> shared memory;  lanes 0-63: add, normalise, bf16, whole-row scatter stores
// A run starts with the predecessor of its first chunk (and of every chunk that opens a (batch, head) row: the row's LAST
// chunk, rp's roll) as an entry that only serves as keys.
#include <cfloat>
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_util.h"
#include "lsh_attn_params.h"
#include "rtts_b200.h"

namespace rtts {
namespace f64 {

constexpr int kDh = 64;
constexpr int kC = 64;                  // chunk = bucket size
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kExactBound = 60.f;

#ifndef RTTS_F64_GROUPS
#define RTTS_F64_GROUPS 2
#endif
constexpr int kGroups = RTTS_F64_GROUPS;          // softmax groups of 4 warps
constexpr int kSoftWarps = 4 * kGroups;
constexpr int kFirstEpiWarp = kSoftWarps;         // sets of 4 warps: warp & 3 = TMEM lane quarter; quarters 0,1 combine + store, 2,3 hand over
#ifndef RTTS_F64_EPISETS
#define RTTS_F64_EPISETS 2
#endif
constexpr int kEpiSets = RTTS_F64_EPISETS;        // epilogue set s takes entries e = s (mod kEpiSets)
static_assert(kEpiSets == 1 || kEpiSets == 2, "the hand-over buffers are indexed by entry parity");
// The last 8 warps: sub-partitions 0 and 1 also host the combining epilogue warps (TMEM lanes 0-63), so they get the two MMA
// issuers and two idle warps, and sub-partitions 2 and 3 the four loader warps (measured: 162 -> 150 us):
//   +0 S issuer, +1 PV issuer, +2 +3 loaders 0 1, +4 +5 idle, +6 +7 loaders 2 3
constexpr int kFirstTailWarp = kFirstEpiWarp + 4 * kEpiSets;
constexpr int kLoaderWarps = 4;
constexpr int kSWarp = kFirstTailWarp;
constexpr int kPVWarp = kFirstTailWarp + 1;
constexpr int kThreads = (kFirstTailWarp + 8) * 32;
__device__ __forceinline__ int loader_index(int warp) {      // 0..3 for the loader warps, -1 otherwise
  const int t = warp - kFirstTailWarp;
  return (t >= 0 && (t & 2)) ? ((t >> 2) * 2 + (t & 1)) : -1;
}

constexpr int kSlots = 8;               // ring of gathered chunks (K and V rows), released when PV of the entry has completed
constexpr int kMetaSlots = 16;          // ring of per-row metadata, released when the epilogue of the entry is done
constexpr int kNS = 3;                  // S / P buffers in TMEM (64 columns each)
constexpr int kND = 4;                  // D buffers in TMEM (64 + 16 columns each)
constexpr uint32_t kColD0 = kNS * 64, kDCols = 80;

struct Smem {
  static constexpr int kChunkBytes = kC * 128;                           // 8 KB: 64 rows of one head
  static constexpr int kOffK = 0;                                        // kSlots + 1 chunks: slot kSlots mirrors slot 0, so that the
  static constexpr int kOffV = (kSlots + 1) * kChunkBytes;               //   128 rows of (entry e, entry e+1) are always contiguous
  static constexpr int kOffOnes = kOffV + kSlots * kChunkBytes;          // 1 KB of bf16 1.0: B operand of the row-sum MMA
  static constexpr int kOffMeta = kOffOnes + 1024;
  // per slot
  static constexpr int kMScale = 0;                                      // float[64]  key role: score_scale*log2e / |k|
  static constexpr int kMQ = 256;                                        // uint2[64]  query role: {-bound (float), limit as fp16 pair}
  static constexpr int kMPos = 768;                                      // int[64]    position | kPadFlag
  static constexpr int kMPos16 = 1024;                                   // half[64]   position (NaN: padded)
  static constexpr int kMMaxMain = 1152;                                 // float[64]  exact mode: block maximum, own-chunk part
  static constexpr int kMMaxLb = 1408;                                   // float[64]  exact mode: block maximum, look-back part
  static constexpr int kMInfo = 1920;                                    // int4 {row_bh, round * T, flags, exact flag of rows 32-63}
  static constexpr int kMetaBytes = 1936;
  static constexpr int kOffX = kOffMeta + kMetaSlots * kMetaBytes;           // hand-over: 2 buffers x 64 rows x (64 O + sum) fp32, rows padded
  static constexpr int kXRow = 272, kXBytes = kC * kXRow;
  static constexpr int kOffStage = kOffX + 2 * kXBytes;                  // output staging: one tile of 32 rows x 128 B (swizzled) per combining warp
  static constexpr int kStageRow = 128, kStageBytes = 32 * kStageRow;
  static constexpr int kOffBar = kOffStage + 2 * kEpiSets * kStageBytes;
  static constexpr int kNumBars = 2 * kSlots + kMetaSlots + 2 * kNS + 8;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmem + 16;
  static_assert(kOffOnes % 1024 == 0 && kOffX % 16 == 0 && kOffStage % 16 == 0 && kOffBar % 8 == 0, "alignment");
  static_assert(kTotal <= 232448, "shared memory budget of one CTA (227 KB)");
};
constexpr int kFlagEmit = 1, kFlagRoundStart = 2, kFlagExact = 4;

#ifdef RTTS_TRACE      // timeline build (tools/trace_fwd64.py): RTTS_DEFS=-DRTTS_TRACE python reformer_tts_b200/csrc/build.py -f
#define F64_STAMP(role, n, k) do { if (p.trace != nullptr && blockIdx.x == 0 && (n) < 64) p.trace[((role) * 64 + (n)) * 8 + (k)] = clock64(); } while (0)
#else
#define F64_STAMP(role, n, k) do { } while (0)
#endif

__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// Bounded wait (a protocol bug traps instead of hanging the GPU).
#ifndef RTTS_F64_WAIT
#define RTTS_F64_WAIT 1
#endif
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
#if RTTS_F64_WAIT == 0          // plain polling: ~10 waiting warps take half of the SM's issue slots (ncu: 55 % issue-active, mostly polls)
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    if (mbar_try_wait_a(bar_addr, parity)) return;
  }
#elif RTTS_F64_WAIT == 1        // try_wait with a suspend-time hint: the hardware parks the thread until the phase completes
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 20); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, P1;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity), "r"(20000u)
        : "memory");
    if (ok) return;
  }
#else                           // poll, sleeping between polls
  if (mbar_try_wait_a(bar_addr, parity)) return;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    __nanosleep(RTTS_F64_WAIT);
    if (mbar_try_wait_a(bar_addr, parity)) return;
  }
#endif
  __trap();
}
// one lane waits (parked by the hardware until the phase completes), the warp follows
__device__ __forceinline__ void warp_wait(uint32_t bar_addr, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait_a(bar_addr, parity);
  __syncwarp();
}
__device__ __forceinline__ void fadd2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "add.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  asm("{\n"
      ".reg .b64 ra, rb, rd;\n"
      "mov.b64 ra, {%2, %3};\n"
      "mov.b64 rb, {%4, %5};\n"
      "mul.rn.f32x2 rd, ra, rb;\n"
      "mov.b64 {%0, %1}, rd;\n"
      "}\n"
      : "=f"(d0), "=f"(d1)
      : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// ---------------------------------------------------------------------------------------------------------------------------
// 16 key columns of one query row.  r: scores (fp32 bits) -> pk: 8 registers of bf16 pairs.
// Packed form: e = exp2(s * key_scale - bound), then ONE compare per pair of keys on the fp16 positions (integers up to 2048
// are exact; a padded key holds NaN and the compares are the unordered ones, so it is always cleared):
//   causal      key position >= query position   - the future, the query itself, and a second copy of the query's own token in a
//               look-back chunk of the previous hash round
//   otherwise   key position == query position   - the query itself (and that second copy)
// exp2(mask_value - m) and exp2(self_value - m) are exact zeros, so clearing the bf16 pair is the reference's arithmetic.
template <bool CAUSAL>
__device__ __forceinline__ void soft16_packed(uint32_t* r, uint32_t a_scale, uint32_t a_p16, float neg_m, uint32_t q_pos2, uint32_t* pk) {
  float* x = reinterpret_cast<float*>(r);
  {
    uint4 s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = lds128(a_scale + q * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      ffma2(x[q * 4 + 0], x[q * 4 + 1], x[q * 4 + 0], x[q * 4 + 1], __uint_as_float(s[q].x), __uint_as_float(s[q].y), neg_m, neg_m);
      ffma2(x[q * 4 + 2], x[q * 4 + 3], x[q * 4 + 2], x[q * 4 + 3], __uint_as_float(s[q].z), __uint_as_float(s[q].w), neg_m, neg_m);
    }
  }
#ifndef RTTS_X_NOEXP
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = exp2f(x[i]);
#endif
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
  const uint4 k0 = lds128(a_p16), k1 = lds128(a_p16 + 16);
  const uint32_t kp[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  const __half2 qp = *reinterpret_cast<const __half2*>(&q_pos2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __half2 k2 = *reinterpret_cast<const __half2*>(&kp[i]);
    pk[i] &= ~(CAUSAL ? __hgeu2_mask(k2, qp) : __hequ2_mask(k2, qp));
  }
}

// Integer form of the same (positions beyond fp16's exact range: T > 2048).
__device__ __forceinline__ void soft16_int(uint32_t* r, uint32_t a_scale, uint32_t a_pos, float neg_m, int q_limit, int q_enc, uint32_t* pk) {
  float* x = reinterpret_cast<float*>(r);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 s = lds128(a_scale + q * 16);
    x[q * 4 + 0] = fmaf(x[q * 4 + 0], __uint_as_float(s.x), neg_m); x[q * 4 + 1] = fmaf(x[q * 4 + 1], __uint_as_float(s.y), neg_m);
    x[q * 4 + 2] = fmaf(x[q * 4 + 2], __uint_as_float(s.z), neg_m); x[q * 4 + 3] = fmaf(x[q * 4 + 3], __uint_as_float(s.w), neg_m);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = exp2f(x[i]);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 kq = lds128(a_pos + q * 16);
    const int kp[4] = {static_cast<int>(kq.x), static_cast<int>(kq.y), static_cast<int>(kq.z), static_cast<int>(kq.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      x[q * 4 + i] = kp[i] > q_limit ? 0.f : x[q * 4 + i];
      x[q * 4 + i] = kp[i] == q_enc ? 0.f : x[q * 4 + i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
}

// Exact mode: the reference's arithmetic (fill values for masked / self entries, true maximum of the row inside this block).
__device__ __forceinline__ float exact16_max(const uint32_t* r, uint32_t a_scale, uint32_t a_pos, int q_limit, int q_enc, float mv, float sv, float mx) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 s = lds128(a_scale + q * 16), kq = lds128(a_pos + q * 16);
    const float ks[4] = {__uint_as_float(s.x), __uint_as_float(s.y), __uint_as_float(s.z), __uint_as_float(s.w)};
    const int kp[4] = {static_cast<int>(kq.x), static_cast<int>(kq.y), static_cast<int>(kq.z), static_cast<int>(kq.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float sc = __uint_as_float(r[q * 4 + i]) * ks[i];
      sc = kp[i] > q_limit ? mv : sc;
      sc = kp[i] == q_enc ? sv : sc;
      mx = fmaxf(mx, sc);
    }
  }
  return mx;
}
__device__ __forceinline__ void exact16(uint32_t* r, uint32_t a_scale, uint32_t a_pos, float neg_m, int q_limit, int q_enc, float mv, float sv, uint32_t* pk) {
  float* x = reinterpret_cast<float*>(r);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const uint4 s = lds128(a_scale + q * 16), kq = lds128(a_pos + q * 16);
    const float ks[4] = {__uint_as_float(s.x), __uint_as_float(s.y), __uint_as_float(s.z), __uint_as_float(s.w)};
    const int kp[4] = {static_cast<int>(kq.x), static_cast<int>(kq.y), static_cast<int>(kq.z), static_cast<int>(kq.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float sc = x[q * 4 + i] * ks[i];
      sc = kp[i] > q_limit ? mv : sc;
      sc = kp[i] == q_enc ? sv : sc;
      x[q * 4 + i] = exp2f(sc + neg_m);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
}

// The run of a CTA as a sequence of ring entries: every owned chunk, preceded by its cyclic predecessor (keys only) where the
// previous entry is not that predecessor (start of the run, start of a (batch, head) row).
struct Entry {
  int row, j, round, jr, b, h;      // (batch, head) row; chunk in row; hash round; chunk in round; batch; head
  int emit, ok;
};
struct EntryIter {
  int gc, g1, cpr, cpround, R, H;
  Entry nx;                          // the next owned chunk
  bool need_pred;
  __device__ EntryIter(int g0, int g1_, int cpr_, int cpround_, int R_, int H_) : gc(g0), g1(g1_), cpr(cpr_), cpround(cpround_), R(R_), H(H_), need_pred(true) {
    nx.row = g0 / cpr_;
    nx.j = g0 - nx.row * cpr_;
    nx.round = nx.j / cpround_;
    nx.jr = nx.j - nx.round * cpround_;
    nx.b = nx.row / H_;
    nx.h = nx.row - nx.b * H_;
    nx.emit = 1;
    nx.ok = 1;
  }
  // (no integer division per entry: the loaders' serial path per entry is what bounds the kernel once everything else overlaps)
  __device__ Entry next() {
    Entry x = nx;
    if (gc >= g1) {
      x.ok = 0;
      return x;
    }
    if (need_pred) {
      x.emit = 0;
      if (nx.j == 0) { x.j = cpr - 1; x.round = R - 1; x.jr = cpround - 1; }
      else if (nx.jr == 0) { x.j = nx.j - 1; x.round = nx.round - 1; x.jr = cpround - 1; }
      else { x.j = nx.j - 1; x.jr = nx.jr - 1; }
      need_pred = false;
      return x;
    }
    ++gc;
    ++nx.j;
    if (++nx.jr == cpround) { nx.jr = 0; ++nx.round; }
    if (nx.j == cpr) {
      nx.j = 0; nx.round = 0; nx.jr = 0; ++nx.row;
      if (++nx.h == H) { nx.h = 0; ++nx.b; }
      need_pred = true;
    }
    return x;
  }
};

// (22 warps = 6 on two of the four sub-partitions, whose register files hold 16384: 80 registers per thread)
__global__ void __launch_bounds__(kThreads, 1) lsh_attn_fwd64_kernel(const AttnFwdParams p, const int num_chunks) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();     // SWIZZLE_128B chunks need 1024-byte alignment
  const uint32_t sbase = smem_u32(smem);
  // barrier rings (shared-window addresses); entry e uses index e % ring size, phase parity (e / ring size) & 1
  const uint32_t a_kfull = sbase + L::kOffBar;                 // [kSlots]     entry landed (4 loader warps)
  const uint32_t a_pv = a_kfull + kSlots * 8;                  // [kSlots]     D_e complete (tcgen05.commit): data slot / S buffer free, D published
  const uint32_t a_epi = a_pv + kSlots * 8;                    // [kMetaSlots] the 4 epilogue warps are done with entry e (D buffer, metadata)
  const uint32_t a_sfull = a_epi + kMetaSlots * 8;             // [kNS]        tcgen05.commit
  const uint32_t a_pfull = a_sfull + kNS * 8;                  // [kNS]        all softmax warps
  const uint32_t a_xfull = a_pfull + kNS * 8;                  // [2 buffers][2 warp pairs]
  const uint32_t a_xfree = a_xfull + 4 * 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int RT = p.R * p.T;
  const int cpr = RT / kC;              // chunks per (batch, head) row
  const int cpround = p.T / kC;         // chunks per hash round
  const int g0 = static_cast<int>(static_cast<int64_t>(blockIdx.x) * num_chunks / gridDim.x);
  const int g1 = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * num_chunks / gridDim.x);
  // entries of this run: owned chunks + one predecessor at the start + one per (batch, head) row opened inside the run
  const int n_entries = g1 > g0 ? (g1 - g0) + 1 + ((g1 - 1) / cpr - g0 / cpr) : 0;

  if (tid == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    int b = 0;
    for (int s = 0; s < kSlots; ++s) mbar_init(bars + b++, 1);          // the loader warp of the entry
    for (int s = 0; s < kSlots; ++s) mbar_init(bars + b++, 1);
    for (int s = 0; s < kMetaSlots; ++s) mbar_init(bars + b++, 4);
    for (int s = 0; s < kNS; ++s) mbar_init(bars + b++, 1);
    for (int s = 0; s < kNS; ++s) mbar_init(bars + b++, 4);              // the 4 warps of a softmax group
    for (int s = 0; s < 8; ++s) mbar_init(bars + b++, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 256; i += kThreads) reinterpret_cast<uint32_t*>(smem + L::kOffOnes)[i] = 0x3F803F80u;      // bf16 1.0 pairs
  fence_proxy_async_smem();
  if (warp == kSWarp) tmem_alloc(tmem_slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) __trap();          // all 512 columns are ours: addresses below are compile-time constants

  if (warp == kSWarp) {
    // ================================================= S_e = [X_e ; X_e+1] X_e^T ===================================
    if (elect_one() && n_entries > 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, kC, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const uint32_t k_lo0 = umma_desc_lo(sbase + L::kOffK, 16);
      mbar_wait_a(a_kfull, 0);
      int sb = 0;
      for (int e = 0; e < n_entries; ++e) {
        const int slot = e & (kSlots - 1);
        if (e + 1 < n_entries) mbar_wait_a(a_kfull + ((e + 1) & (kSlots - 1)) * 8, ((e + 1) / kSlots) & 1);      // the upper 64 rows of A
        if (e >= kNS) mbar_wait_a(a_pv + ((e - kNS) & (kSlots - 1)) * 8, ((e - kNS) / kSlots) & 1);               // PV(e - kNS) has consumed the buffer
        tc_fence_after_sync();
        F64_STAMP(0, e, 0);
        const uint32_t x_lo = k_lo0 + slot * (L::kChunkBytes >> 4);
        const uint32_t t_s = sb * 64;
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) umma_ss_lo(t_s, x_lo + kk * 2, x_lo + kk * 2, hi, idesc, kk > 0);
        umma_commit(reinterpret_cast<uint64_t*>(smem + L::kOffBar) + 2 * kSlots + kMetaSlots + sb);
        F64_STAMP(0, e, 1);
        if (++sb == kNS) sb = 0;
      }
    }
  } else if (warp == kPVWarp) {
    // ================================================= D_e = P_e [V_e | 1] =========================================
    if (elect_one()) {
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kDh, false, true);
      constexpr uint32_t idesc_sum = umma_idesc_bf16(128, 16, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      constexpr uint32_t hi_ones = umma_desc_hi_sw128(0);         // the 16 rows of B alias one 1 KB atom of ones
      const uint32_t v_lo0 = umma_desc_lo(sbase + L::kOffV, 0);   // MN-major operand (V rows)
      const uint32_t ones_lo = umma_desc_lo(sbase + L::kOffOnes, 16);
      int sb = 0;
      uint32_t sph = 0;
      for (int e = 0; e < n_entries; ++e) {
        const int slot = e & (kSlots - 1);
        mbar_wait_a(a_pfull + sb * 8, sph);
        F64_STAMP(0, e, 2);
        if (e >= kND) mbar_wait_a(a_epi + ((e - kND) & (kMetaSlots - 1)) * 8, ((e - kND) / kMetaSlots) & 1);      // D buffer drained
        tc_fence_after_sync();
        F64_STAMP(0, e, 3);
        const uint32_t t_p = sb * 64, t_d = kColD0 + (e & (kND - 1)) * kDCols;
        const uint32_t v_lo = v_lo0 + slot * (L::kChunkBytes >> 4);
#pragma unroll
        for (int s = 0; s < kC / 16; ++s) umma_ts_lo(t_d, t_p + 16 * s, v_lo + s * (2048 >> 4), hi, idesc_o, s > 0);
#pragma unroll
        for (int s = 0; s < kC / 16; ++s) umma_ts_lo(t_d + 64, t_p + 16 * s, ones_lo, hi_ones, idesc_sum, s > 0);
        umma_commit(reinterpret_cast<uint64_t*>(smem + L::kOffBar) + kSlots + slot);
        F64_STAMP(0, e, 4);
        if (++sb == kNS) { sb = 0; sph ^= 1u; }
      }
    }
  } else if (loader_index(warp) >= 0) {
    // ================================================= loaders =====================================================
    // Warp w gathers the entries e = w (mod 4) on its own (no cross-warp step on the per-entry path): lane = (row within a group of
    // 4, 16-byte piece), 16 passes of K and V rows; lane l also owns the metadata of rows l and l + 32.  Software-pipelined: the
    // stickers and |x|^2 / mask values of the warp's NEXT entry are requested before the current one is copied.
    const int lw = loader_index(warp);
    const int grp = lane >> 3, c = lane & 7;
    const uint32_t ld32 = static_cast<uint32_t>(p.ld);
    const float ssl2 = p.score_scale_log2;
    EntryIter it(g0, g1, cpr, cpround, p.R, p.H);
    auto fetch = [&]() {                              // this warp's next entry
      Entry x = it.next();
      it.next(); it.next(); it.next();
      return x;
    };
    // raw stickers (round * T + position) of rows lane, lane + 32 and their |x|^2 / mask values
    auto prefetch = [&](const Entry& x, int& s0, int& s1, float& q0, float& q1, uint32_t& v0, uint32_t& v1) {
      q0 = q1 = 1.f;
      v0 = v1 = 1u;
      if (!x.ok) return;
      const int32_t* stk = p.sticker + static_cast<int64_t>(x.row) * RT + x.j * kC;
      s0 = __ldg(stk + lane);
      s1 = __ldg(stk + 32 + lane);
      const int base = x.round * p.T;
      const float* sq = p.sumsq + static_cast<int64_t>(x.row) * p.T - base;
      q0 = __ldg(sq + s0);
      q1 = __ldg(sq + s1);
      if (p.mask != nullptr) {
        const uint8_t* mk = p.mask + static_cast<int64_t>(x.b) * p.T - base;
        v0 = __ldg(mk + s0);
        v1 = __ldg(mk + s1);
      }
    };
    for (int i = 0; i < lw; ++i) it.next();
    Entry e0 = fetch(), e1;
    int s0 = 0, s1 = 0, n0 = 0, n1 = 0;
    float q0, q1, nq0, nq1;
    uint32_t v0, v1, nv0, nv1;
    prefetch(e0, s0, s1, q0, q1, v0, v1);
    int pending = -1;
    auto announce = [&](int slot) {
      fence_proxy_async_smem();            // cp.async / st.shared data -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_kfull + slot * 8);
    };
    for (int e = lw; e0.ok; e += kLoaderWarps) {
      const int slot = e & (kSlots - 1), ms = e & (kMetaSlots - 1);
      if (lane == 0) F64_STAMP(1, e, 0);
      e1 = fetch();
      prefetch(e1, n0, n1, nq0, nq1, nv0, nv1);
      const int base_round = e0.round * p.T;
      // data slot: PV of the entry that lived there has completed (which implies its S); metadata slot: its epilogue is done
      {
        int ready = 1;
        if (lane == 0) {
          if (e >= kSlots) ready = mbar_try_wait_a(a_pv + slot * 8, ((e / kSlots) - 1) & 1);
          if (ready && e >= kMetaSlots) ready = mbar_try_wait_a(a_epi + ms * 8, ((e / kMetaSlots) - 1) & 1);
        }
        ready = __shfl_sync(0xffffffffu, ready, 0);
        if (!ready) {
          if (pending >= 0) {              // would block: do not sit on an entry that has already landed
            cp_async_wait<0>();
            announce(pending);
            pending = -1;
          }
          if (e >= kSlots) warp_wait(a_pv + slot * 8, ((e / kSlots) - 1) & 1);
          if (e >= kMetaSlots) warp_wait(a_epi + ms * 8, ((e / kMetaSlots) - 1) & 1);
        }
      }
      if (lane == 0) F64_STAMP(1, e, 1);
      const int p0 = s0 - base_round, p1 = s1 - base_round;
      // rows 4i + grp: the swizzle term (row & 7) alternates between grp and grp + 4 with the parity of i
      const uint32_t sKe = sbase + L::kOffK + slot * L::kChunkBytes + sw128_offset(grp, c), sKo = sbase + L::kOffK + slot * L::kChunkBytes + sw128_offset(grp + 4, c);
      constexpr uint32_t kVoff = L::kOffV - L::kOffK - L::kChunkBytes;      // slot s of the V ring sits kVoff + 8 KB behind slot s of the K ring
      const __nv_bfloat16* qk_b = p.qk + static_cast<int64_t>(e0.b) * p.T * p.ld + e0.h * kDh + c * 8;
      const __nv_bfloat16* v_b = p.v + static_cast<int64_t>(e0.b) * p.T * p.ld + e0.h * kDh + c * 8;
      const bool mirror = slot == 0 && e > 0;      // slot kSlots mirrors slot 0 (K rows): A of S(e-1) = slots kSlots-1, kSlots
      // (a rolled loop: the kernel lives or dies by its instruction-cache footprint - fully unrolled, these 16 passes alone were 30 KB
      // of code and `no_inst` was again the first stall reason of every role)
#pragma unroll 1
      for (int i2 = 0; i2 < kC / 8; ++i2) {
        const int psel = i2 < 4 ? p0 : p1;
        const int pre = __shfl_sync(0xffffffffu, psel, (8 * i2 + grp) & 31), pro = __shfl_sync(0xffffffffu, psel, (8 * i2 + 4 + grp) & 31);
        const uint32_t offe = static_cast<uint32_t>(pre) * ld32, offo = static_cast<uint32_t>(pro) * ld32;       // element offsets inside the batch entry: T * ld < 2^31 (checked by the host)
        const uint32_t de = sKe + i2 * 1024, dd = sKo + i2 * 1024;
#ifdef RTTS_X_NOCOPY
        if (offe == 0xffffffffu)
#endif
        {
        cp_async16(de, qk_b + offe);
        cp_async16(de + kVoff + L::kChunkBytes, v_b + offe);
        cp_async16(dd, qk_b + offo);
        cp_async16(dd + kVoff + L::kChunkBytes, v_b + offo);
        }
        if (mirror) {
          cp_async16(de + kSlots * L::kChunkBytes, qk_b + offe);
          cp_async16(dd + kSlots * L::kChunkBytes, qk_b + offo);
        }
      }
      cp_async_commit();
      if (lane == 0) F64_STAMP(1, e, 2);
      {
        const uint32_t a_meta = sbase + L::kOffMeta + ms * L::kMetaBytes;
        bool big = false;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int r = lane + 32 * hh, mpos = hh == 0 ? p0 : p1;
          const float ssq = hh == 0 ? q0 : q1;
          const bool valid = (hh == 0 ? v0 : v1) != 0;
          // key scale = score_scale * log2(e) / |k|  (rp R5: x / max(|x|, 1e-12); hf:1042-1056: x * rsqrt(mean(x^2) + 1e-6) / sqrt(dh)), and the
          // stabiliser |q| * score_scale * log2(e) = score_scale_log2^2 / key_scale, times (1 + 2^-10) so rounding cannot push a score above it
          float ks, bound;
          if (p.key_norm == RTTS_KEYNORM_L2) {
            const float s2 = fmaxf(ssq, 1e-24f), rs = rsqrtf(s2);
            ks = rs * ssl2;
            bound = s2 * rs * (ssl2 * 1.001f);
          } else {
            const float s2 = ssq * (1.f / kDh) + 1e-6f, rs = rsqrtf(s2);
            ks = rs * (0.125f * ssl2);
            bound = s2 * rs * (8.f * 1.001f * ssl2);
          }
          const bool bigr = bound >= kExactBound;
          big |= bigr;
          // a padded query under the query-and-key mask sees nothing: -inf clears its whole row, the epilogue then treats it like
          // every row that sees only itself (exact mode works on the integer positions and keeps the finite bound)
          if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && !valid && !bigr) bound = __int_as_float(0x7f800000);
          const __half p16 = valid ? __int2half_rn(mpos) : __ushort_as_half(0x7fff);      // padded key: NaN, cleared by the unordered compares
          const __half2 q2 = __half2half2(__int2half_rn(mpos));
          sts32(a_meta + L::kMScale + r * 4, __float_as_uint(ks));
          asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a_meta + L::kMQ + r * 8), "r"(__float_as_uint(-bound)), "r"(*reinterpret_cast<const uint32_t*>(&q2)) : "memory");
          sts32(a_meta + L::kMPos + r * 4, static_cast<uint32_t>(valid ? mpos : (mpos | kPadFlag)));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(a_meta + L::kMPos16 + r * 2), "h"(*reinterpret_cast<const unsigned short*>(&p16)) : "memory");
        }
        const bool any_big = __any_sync(0xffffffffu, big);
        if (lane == 0) {
          const int flags = (e0.emit ? kFlagEmit : 0) | (e0.jr == 0 ? kFlagRoundStart : 0) | (any_big ? kFlagExact : 0);
          sts128(a_meta + L::kMInfo, make_uint4(static_cast<uint32_t>(e0.row), static_cast<uint32_t>(base_round), static_cast<uint32_t>(flags), 0u));
        }
      }
      if (pending >= 0) {
        cp_async_wait<1>();         // everything but the group just committed has landed
        announce(pending);
      }
      pending = slot;
      if (lane == 0) F64_STAMP(1, e, 3);
      e0 = e1;
      s0 = n0; s1 = n1; q0 = nq0; q1 = nq1; v0 = nv0; v1 = nv1;
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      announce(pending);
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstTailWarp) {
    // ================================================= epilogue ====================================================
    const int q = warp & 3, eset = (warp - kFirstEpiWarp) >> 2;
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    if (q >= 2) {
      // ---- lanes 64-127: the look-back part of entry e+1's output, handed over through shared memory
      const int pair = q - 2;
      uint32_t uses[2] = {0u, 0u};
      for (int e = eset; e < n_entries; e += kEpiSets) {
        warp_wait(a_pv + (e & (kSlots - 1)) * 8, (e / kSlots) & 1);
        tc_fence_after_sync();
        if (pair == 0 && lane == 0) F64_STAMP(3, e, 0);
        bool valid = false;
        if (e + 1 < n_entries) valid = (lds32(sbase + L::kOffMeta + ((e + 1) & (kMetaSlots - 1)) * L::kMetaBytes + L::kMInfo + 8) & kFlagEmit) != 0;
        if (valid) {
          const int xb = (e + 1) & 1;
          const uint32_t t_d = t_lane + kColD0 + (e & (kND - 1)) * kDCols;
          uint32_t r0[32], r1[32], rs;
          tmem_ld32(t_d, r0);
          tmem_ld32(t_d + 32, r1);
          tmem_ld1(t_d + 64, &rs);
          warp_wait(a_xfree + (xb * 2 + pair) * 8, (uses[xb] & 1) ^ 1);      // (first use: passes on the fresh barrier)
          ++uses[xb];
          tmem_ld_wait();
          tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive_a(a_epi + (e & (kMetaSlots - 1)) * 8);
          const uint32_t a_x = sbase + L::kOffX + xb * L::kXBytes + (pair * 32 + lane) * L::kXRow;
#pragma unroll
          for (int i = 0; i < 8; ++i) sts128(a_x + i * 16, make_uint4(r0[4 * i], r0[4 * i + 1], r0[4 * i + 2], r0[4 * i + 3]));
#pragma unroll
          for (int i = 0; i < 8; ++i) sts128(a_x + 128 + i * 16, make_uint4(r1[4 * i], r1[4 * i + 1], r1[4 * i + 2], r1[4 * i + 3]));
          sts32(a_x + 256, rs);
          __syncwarp();
          if (lane == 0) mbar_arrive_a(a_xfull + (xb * 2 + pair) * 8);
          if (pair == 0 && lane == 0) F64_STAMP(3, e, 1);
        } else {
          if (lane == 0) mbar_arrive_a(a_epi + (e & (kMetaSlots - 1)) * 8);
        }
      }
    } else {
      // ---- lanes 0-63: own-chunk part from TMEM + look-back part from shared memory -> normalise -> bf16 -> store
      const int pair = q;
      uint32_t uses[2] = {0u, 0u};
      const uint32_t a_stage = sbase + L::kOffStage + (eset * 2 + pair) * L::kStageBytes + lane * L::kStageRow;
      const uint32_t l7 = lane & 7;
      for (int e = eset; e < n_entries; e += kEpiSets) {
        warp_wait(a_pv + (e & (kSlots - 1)) * 8, (e / kSlots) & 1);
        tc_fence_after_sync();
        const uint32_t a_meta = sbase + L::kOffMeta + (e & (kMetaSlots - 1)) * L::kMetaBytes;
        const uint4 info = lds128(a_meta + L::kMInfo);
        if (pair == 0 && lane == 0) F64_STAMP(3, e, 2);
        if ((info.z & kFlagEmit) == 0) {
          if (lane == 0) mbar_arrive_a(a_epi + (e & (kMetaSlots - 1)) * 8);
          continue;
        }
        const int rr = pair * 32 + lane;          // row inside the chunk
        const int xb = e & 1;
        const uint32_t t_d = t_lane + kColD0 + (e & (kND - 1)) * kDCols;
        const uint32_t a_x = sbase + L::kOffX + xb * L::kXBytes + rr * L::kXRow;
        uint32_t ra[16], rs;
        tmem_ld16(t_d, ra);
        tmem_ld1(t_d + 64, &rs);
        const int pos_enc = static_cast<int>(lds32(a_meta + L::kMPos + rr * 4));
        float row_max = -__uint_as_float(lds32(a_meta + L::kMQ + rr * 8));
        const bool exact = ((info.z & kFlagExact) | info.w) != 0;
        warp_wait(a_xfull + (xb * 2 + pair) * 8, uses[xb] & 1);
        ++uses[xb];
        if (pair == 0 && lane == 0) F64_STAMP(3, e, 3);
        tmem_ld_wait();
        float sum = __uint_as_float(rs) + __uint_as_float(lds32(a_x + 256));
        float w_main = 1.f, w_lb = 1.f;
        if (exact) {
          // the two parts were formed against their own block maxima: bring them to the common one
          const float m_main = __uint_as_float(lds32(a_meta + L::kMMaxMain + rr * 4)), m_lb = __uint_as_float(lds32(a_meta + L::kMMaxLb + rr * 4));
          row_max = fmaxf(m_main, m_lb);
          w_main = exp2f(m_main - row_max);
          w_lb = exp2f(m_lb - row_max);
          sum = __uint_as_float(rs) * w_main + __uint_as_float(lds32(a_x + 256)) * w_lb;
        }
        // all terms exactly zero: the row sees only itself (rp R8): softmax uniform over the self columns, which all hold the
        // query's own token, so out = v[own position], lse = self_value + log(#self columns)
        const bool lonely = !(sum > 0.f);
        if (lonely) {
          // a second self column exists if the own token also sits in the look-back chunk (first chunk of a hash round only);
          // the previous entry's metadata is still there: a slot is rewritten 16 entries later, which needs PV(e+7), which needs
          // the epilogue warps of this set to have arrived for entry e+2, i.e. to be past this point
          int dup = 0;
          if (info.z & kFlagRoundStart) {
            const uint32_t a_prev = sbase + L::kOffMeta + ((e - 1) & (kMetaSlots - 1)) * L::kMetaBytes + L::kMPos;
            for (int j = 0; j < kC; ++j) dup |= static_cast<int>(lds32(a_prev + j * 4)) == pos_enc;
          }
          sum = dup ? 2.f : 1.f;
          row_max = p.self_value_log2;
        }
        const float inv = 1.f / sum;
        w_main *= inv;
        w_lb *= inv;
        // a quarter (16 columns) of the row at a time: own-chunk part (registers, from TMEM) + look-back part (shared memory) -> bf16 ->
        // staging tile of this warp (row = lane, 16-byte pieces swizzled by the row so that both the row-wise writes here and the
        // piece-wise reads of the store phase are conflict-free)
        auto quarter = [&](const uint32_t* a, int qq) {
          uint4 xv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) xv[i] = lds128(a_x + qq * 64 + i * 16);
          const uint32_t* xw = reinterpret_cast<const uint32_t*>(xv);
          float o[16];
          if (!exact) {
#pragma unroll
            for (int k = 0; k < 16; k += 2) {
              fadd2(o[k], o[k + 1], __uint_as_float(a[k]), __uint_as_float(a[k + 1]), __uint_as_float(xw[k]), __uint_as_float(xw[k + 1]));
              fmul2(o[k], o[k + 1], o[k], o[k + 1], inv, inv);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) o[k] = __uint_as_float(a[k]) * w_main + __uint_as_float(xw[k]) * w_lb;
          }
#pragma unroll
          for (int i = 0; i < 2; ++i)
            sts128(a_stage + (((qq * 2 + i) ^ l7) << 4),
                   make_uint4(pack_bf16(o[8 * i], o[8 * i + 1]), pack_bf16(o[8 * i + 2], o[8 * i + 3]), pack_bf16(o[8 * i + 4], o[8 * i + 5]), pack_bf16(o[8 * i + 6], o[8 * i + 7])));
        };
        uint32_t rb[16];
#pragma unroll 1
        for (int hq = 0; hq < 2; ++hq) {      // the next quarter's TMEM load is in flight while this one is combined
          tmem_ld16(t_d + 32 * hq + 16, rb);
          quarter(ra, 2 * hq);
          tmem_ld_wait();
          if (hq == 0) tmem_ld16(t_d + 32, ra);
          quarter(rb, 2 * hq + 1);
          if (hq == 0) tmem_ld_wait();
        }
        tc_fence_before_sync();
        if (pair == 0 && lane == 0) F64_STAMP(3, e, 7);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_a(a_xfree + (xb * 2 + pair) * 8);
          mbar_arrive_a(a_epi + (e & (kMetaSlots - 1)) * 8);
        }
        const int pos = pos_enc & ~kPadFlag;
        const int row_bh = static_cast<int>(info.x);
        if (__any_sync(0xffffffffu, lonely)) {
          if (lonely) {
            const int b = row_bh / p.H, h = row_bh - b * p.H;
            const uint4* vrow = reinterpret_cast<const uint4*>(p.v + (static_cast<int64_t>(b) * p.T + pos) * p.ld + h * kDh);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) sts128(a_stage + ((ch ^ l7) << 4), __ldg(vrow + ch));
          }
          __syncwarp();
        }
        const uint32_t out_slot = static_cast<uint32_t>(static_cast<int>(info.y) + pos);      // unsorted slot inside the (batch, head) row = round * T + position
        const int64_t row_base = static_cast<int64_t>(row_bh) * RT;
        p.lse_rounds[row_base + out_slot] = (row_max + log2f(sum)) * kLn2;
        // scatter-store, one full 128-byte row per 8 lanes (four rows per instruction)
        {
          const uint32_t a_tile = a_stage - lane * 128;
          const char* obase = reinterpret_cast<const char*>(p.o_rounds) + row_base * (kDh * 2) + l7 * 16;
#pragma unroll 2
          for (int itr = 0; itr < 8; ++itr) {
            const int row = itr * 4 + (lane >> 3);
            const uint4 u = lds128(a_tile + row * 128 + ((l7 ^ (row & 7)) << 4));
            const uint32_t os = __shfl_sync(0xffffffffu, out_slot, row);
            *reinterpret_cast<uint4*>(const_cast<char*>(obase) + static_cast<uint64_t>(os) * (kDh * 2)) = u;
          }
        }
        __syncwarp();                 // the staging tile is free again
        if (pair == 0 && lane == 0) F64_STAMP(3, e, 5);
      }
    }
  } else if (warp < kSoftWarps) {
    // ================================================= softmax =====================================================
    // kGroups groups of 4 warps; group g takes the blocks e = g (mod kGroups), so that the hand-overs of one block (S ready ->
    // P ready -> PV) hide behind the arithmetic of the other.  Thread = (query row = TMEM lane, all 64 key columns of the block).
    const int q = warp & 3, g = warp >> 2;
    const bool is_main = q < 2;
    const int rr = (q & 1) * 32 + lane;           // row inside the query's chunk
    const uint32_t t_lane = static_cast<uint32_t>(q * 32) << 16;
    const float mv = p.mask_value_log2, sv = p.self_value_log2;
    int sb = g % kNS;
    uint32_t sph = 0;
    for (int e = g; e < n_entries; e += kGroups) {
      if (q == 0 && lane == 0) F64_STAMP(2, e, 0);
      warp_wait(a_sfull + sb * 8, sph);
      tc_fence_after_sync();
      if (q == 0 && lane == 0) F64_STAMP(2, e, 1);
      const uint32_t a_meta_k = sbase + L::kOffMeta + (e & (kMetaSlots - 1)) * L::kMetaBytes;                       // keys: entry e
      const uint32_t a_meta_q = is_main ? a_meta_k : sbase + L::kOffMeta + ((e + 1) & (kMetaSlots - 1)) * L::kMetaBytes;      // queries: entry e / e+1
      uint2 fl = make_uint2(0u, 0u);
      if (is_main || e + 1 < n_entries) fl = lds64(a_meta_q + L::kMInfo + 8);
      if (fl.x & kFlagEmit) {
        const uint32_t t_s = t_lane + sb * 64;
        const uint32_t a_scale = a_meta_k + L::kMScale, a_pos = a_meta_k + L::kMPos, a_p16 = a_meta_k + L::kMPos16;
        const uint2 qm = lds64(a_meta_q + L::kMQ + rr * 8);
        const float neg_m = __uint_as_float(qm.x);
        if ((fl.x & kFlagExact) | fl.y) {
          const int q_enc = static_cast<int>(lds32(a_meta_q + L::kMPos + rr * 4));
          int q_limit = p.causal ? (q_enc & ~kPadFlag) : (kPadFlag - 1);
          if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && (q_enc & kPadFlag)) q_limit = -1;
          float mx = -FLT_MAX;
#pragma unroll 1
          for (int s = 0; s < 4; ++s) {
            uint32_t r[16];
            tmem_ld16(t_s + 16 * s, r);
            tmem_ld_wait();
            mx = exact16_max(r, a_scale + s * 64, a_pos + s * 64, q_limit, q_enc, mv, sv, mx);
          }
          sts32(a_meta_q + (is_main ? L::kMMaxMain : L::kMMaxLb) + rr * 4, __float_as_uint(mx));
#pragma unroll 1
          for (int s = 0; s < 4; ++s) {
            uint32_t r[16], pk[8];
            tmem_ld16(t_s + 16 * s, r);
            tmem_ld_wait();
            exact16(r, a_scale + s * 64, a_pos + s * 64, -mx, q_limit, q_enc, mv, sv, pk);
            tmem_st8(t_s + 16 * s, pk);
          }
        } else if (!p.pos16) {
          const int q_enc = static_cast<int>(lds32(a_meta_q + L::kMPos + rr * 4));
          int q_limit = p.causal ? (q_enc & ~kPadFlag) : (kPadFlag - 1);
          if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && (q_enc & kPadFlag)) q_limit = -1;
#pragma unroll 1
          for (int s = 0; s < 4; ++s) {
            uint32_t r[16], pk[8];
            tmem_ld16(t_s + 16 * s, r);
            tmem_ld_wait();
            soft16_int(r, a_scale + s * 64, a_pos + s * 64, neg_m, q_limit, q_enc, pk);
            tmem_st8(t_s + 16 * s, pk);
          }
        } else {
          // packed form: one short rolled loop for every warp (the kernel's first limit was instruction fetch: with an unrolled
          // variant per warp role, 72 KB of code against a 6 KB L0 / 32 KB L1.5 instruction cache, half of the stall samples of
          // the working warps were `no_inst`)
#pragma unroll 1
          for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[32], pk[8];
            tmem_ld32(t_s + 32 * hh, r);
            tmem_ld_wait();
            if (p.causal) {
              soft16_packed<true>(r, a_scale + hh * 128, a_p16 + hh * 64, neg_m, qm.y, pk);
              tmem_st8(t_s + 32 * hh, pk);
              soft16_packed<true>(r + 16, a_scale + hh * 128 + 64, a_p16 + hh * 64 + 32, neg_m, qm.y, pk);
            } else {
              soft16_packed<false>(r, a_scale + hh * 128, a_p16 + hh * 64, neg_m, qm.y, pk);
              tmem_st8(t_s + 32 * hh, pk);
              soft16_packed<false>(r + 16, a_scale + hh * 128 + 64, a_p16 + hh * 64 + 32, neg_m, qm.y, pk);
            }
            tmem_st8(t_s + 32 * hh + 16, pk);
          }
        }
        if (q == 0 && lane == 0) F64_STAMP(2, e, 7);
        tmem_st_wait();
      }
      tc_fence_before_sync();       // this thread's TMEM reads of S / writes of P precede the MMAs that consume / overwrite the buffer
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_pfull + sb * 8);
      if (q == 0 && lane == 0) F64_STAMP(2, e, 2);
      sb += kGroups;
      if (sb >= kNS) { sb -= kNS; sph ^= 1u; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kSWarp) tmem_dealloc(0, 512);
}

}  // namespace f64

int launch_attn_fwd64(const AttnFwdParams& p, int B, cudaStream_t stream) {
  using L = f64::Smem;
  static bool configured = false;   // idempotent attribute set; benign if raced
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(f64::lsh_attn_fwd64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int64_t chunks = static_cast<int64_t>(B) * p.H * (static_cast<int64_t>(p.R) * p.T / f64::kC);
  if (chunks <= 0 || chunks >= (1ll << 31)) return fail(kErrBadArg, "rtts_lsh_attn_fwd: bad grid");
  const int grid = chunks < kNumSMs ? static_cast<int>(chunks) : kNumSMs;     // persistent: one CTA per SM
  f64::lsh_attn_fwd64_kernel<<<grid, f64::kThreads, L::kTotal, stream>>>(p, static_cast<int>(chunks));
  return check_launch("rtts_lsh_attn_fwd");
}

}  // namespace rtts
