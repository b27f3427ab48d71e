// Fused chunked shared-QK attention, forward (rtts_lsh_attn_fwd) and the round merge
// (rtts_lsh_merge_fwd).  See include/rtts_b200.h for the contract and DESIGN.md "K9" for the tiling.
//
// A tile is 128 consecutive SORTED query slots of one (batch, head) row: two chunks of bucket 64 or one chunk of
// bucket 128.  Its keys are the same 128 slots plus the BUCKET slots before them (look-one-back, rp R6), and the
// queries are simply the main 128 rows (shared-QK: one gather serves both operands).  Rows are gathered straight from
// the UNSORTED token-major qk / v arrays through `sticker` with 16-byte cp.async into SWIZZLE_128B shared-memory tiles
// (a row is 64 bf16 = one 128-byte swizzle row), so the R sorted copies the reference materialises never exist.
//
// Each CTA walks a CONTIGUOUS run of tiles and keeps the gathered 128-row blocks in a ring, so the look-back rows of
// tile k are the tail of the block gathered for tile k-1: every row is fetched once per hash round (the gather is what
// loads the SM's load/store pipe most).
//   S = Q K^T          tcgen05.mma  M=128, N=BUCKET (look-back block) + N=128 (main block), K=64  -> TMEM fp32
//   softmax            thread = (query row, quarter of its 2*bucket window): reads S from TMEM, applies the per-key 1/|k|
//                      scale (keys are normalised AFTER the fp32-accumulated dot) and the padding / causal / self masks
//                      from the position ids, single pass (see below), and writes P (bf16 pairs) back into TMEM over
//                      the S columns it has just consumed
//   O = P V            tcgen05.mma with A = P from TMEM, B = the gathered V block used MN-major -> TMEM
//   epilogue           O / rowsum -> bf16, scatter-stored at the UNSORTED slot (r*T + pos); lse too.
#include <cfloat>
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_util.h"
#include "lsh_attn_params.h"
#include "rtts_b200.h"

namespace rtts {

constexpr int kDh = 64;
constexpr int kQRows = 128;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// Persistent, warp-specialised pipeline; CTA c owns tiles [c*N/grid, (c+1)*N/grid) of the (row, tile-in-row) order.
//   warp 24     : TMEM allocation; one elected lane issues every tcgen05.mma  (S(k) as soon as block k lands, then PV(k-1) when P(k-1) is ready)
//   warps 16-19 : epilogue - thread = query row: O(k) from TMEM, divide by the row sum the softmax pair left in shared memory, bf16,
//                 scatter-store at the unsorted slot, lse.  Keeps the softmax pairs off the tensor pipe's latency.
//   warps 20-23 : loaders - sticker -> position -> qk and V rows by 16-byte cp.async into the ring slot of the tile, key scale
//                 from sumsq.  Software-pipelined: sticker loads run two tiles ahead, sumsq / mask loads one tile ahead, and a
//                 tile is announced (cp.async.wait_group 1) after the next one's copies are queued, so no load latency is exposed.
//   warps 0-7   : softmax pair 0 (tiles k even);  warps 8-15: softmax pair 1 (tiles k odd).  A pair is two warpgroups sharing a
//                 tile: thread (part, row) owns a quarter of the row's key window
// Shared memory: kSlots ring slots {K block, V block} of 128 rows + per-row key scale / position.  Tile k lives in slot k % kSlots
// and reads its look-back rows from the tail of slot (k-1) % kSlots.  The first tile of a CTA and the first tile of a (batch, head)
// row have no predecessor in the ring ("fresh"): their look-back rows are gathered into the tail of slot (k-1) % kSlots once the
// tile that lived there is done (a pipeline bubble once per row).
// TMEM: one 256-column region per softmax pair.  bucket 64: S in [0,192), O in [192,256); P blocks in place at [32q, 32q+16).
//       bucket 128: S in [0,256); P blocks compacted to [16q,16q+16) (first half) / [128+16(q-4), ...) (second half), O in [64,128).
// mbarriers: full[slot] (128 loader arrivals) -> s_full[pair] (tcgen05.commit) -> p_full[pair] (256 softmax arrivals)
//            -> o_full[pair] + free[slot(s)] (tcgen05.commit after PV) -> o_free[pair] (128 epilogue arrivals).
//
// Softmax is single-pass: keys are unit vectors after normalisation, so |s_ij| <= |q_i| * score_scale (Cauchy-Schwarz) and
// m_i = that bound is a valid stabiliser known before any score is read; masked / self entries give exp2() == 0 exactly.  A row
// that sums to zero can only see itself (masked to self_value): its softmax is uniform over the self columns, set analytically.
// If any query of a tile has a bound >= 60 (norms so large that a visible key could underflow against the bound) the loader
// flags the tile and the whole pair runs the exact two-pass arithmetic of the reference (row max first) instead.
// 16 softmax warps.  bucket 64: ONE group, thread = (query row, quarter of its window), works on every tile and alternates between
// the two TMEM regions.  bucket 128: two groups (thread = (row, half of the window)), one per region / tile parity - its compacted P
// layout (which makes room for O inside the S columns) lets a thread write only behind the S columns of its own half.
#ifndef RTTS_SOFTMAX_WARPS
#define RTTS_SOFTMAX_WARPS 16
#endif
constexpr int kSoftmaxWarps = RTTS_SOFTMAX_WARPS;
constexpr int kMaxParts = 4;
// Warp roles by index.  The SM's issue arbiter favours higher warp ids, so the producers that everything else waits on get
// the top ids: warps 0-7 softmax groups, warps 8-11 epilogue, warps 12-19 loaders, warp 20 MMA issuer.
constexpr int kFirstEpiWarp = kSoftmaxWarps;
constexpr int kEpiThreads = 128;       // thread = query row = TMEM lane
constexpr int kFirstLoaderWarp = kFirstEpiWarp + 4;
#ifndef RTTS_LOADER_WARPS
#define RTTS_LOADER_WARPS 4
#endif
constexpr int kLoaderWarps = RTTS_LOADER_WARPS;
#ifndef RTTS_LOADER_GROUPS
#define RTTS_LOADER_GROUPS 1
#endif
constexpr int kLoaderGroups = RTTS_LOADER_GROUPS;      // independent loader groups, group g gathers tiles k = g (mod kLoaderGroups)
constexpr int kLoaderThreads = kLoaderWarps * 32 / kLoaderGroups;      // threads of one group
constexpr int kMmaWarp = kFirstLoaderWarp + kLoaderWarps;
constexpr int kFwdThreads = (kMmaWarp + 2) * 32;      // 672 threads -> 80 registers each (setmaxnreg rebalancing was tried: the epilogue and
                                                      // loader warpgroups spill below 64 registers and the spills cost more than the softmax gains)
constexpr float kExactBound = 60.f;

template <int BUCKET>
struct AttnFwdSmem {
  static constexpr int kKeyRows = kQRows + BUCKET;
  static constexpr int kSlots = 6;
  static constexpr int kTileBytes = kQRows * 128;                    // 16 KB: 128 rows of one head
  static constexpr int kOffK = 0;                                    // per slot (1024-B aligned): [ K block | V block ]
  static constexpr int kOffV = kTileBytes;
  static constexpr int kSlotBytes = 2 * kTileBytes;
  static constexpr int kOffOnes = kSlots * kSlotBytes;               // 8 rows x 128 B of bf16 1.0 (1024-B aligned): B operand of the row-sum MMA (P . 1)
  static constexpr int kOffMeta = kOffOnes + 1024;                   // per slot: float scale[128], int pos[128], int exact_tag (+pad), half pos16[128]
  static constexpr int kMetaScale = 0, kMetaPos = kQRows * 4;
  static constexpr int kMetaGeo = 2 * kQRows * 4;                    // one 16-byte record {row_bh, base_main (round * T), round_start, exact tag}
  static constexpr int kMetaTag = kMetaGeo + 12;                     // (the first three words by loader thread 0, the tag by whichever thread sees a large bound)
  static constexpr int kMetaPos16 = 2 * kQRows * 4 + 32;             // fp16 position per row (+inf for a padded token): packed position mask
  static constexpr int kMetaBytes = 2 * kQRows * 4 + 32 + kQRows * 2;
  static constexpr int kOffPart = kOffMeta + kSlots * kMetaBytes;    // float[2 softmax groups][kParts][128 rows]: row maxima of the exact two-pass mode (exchange inside one tile)
  static constexpr int kOffDup = kOffPart + 2 * kMaxParts * kQRows * 4;   // uint[2 tile parities][2 phases][128 rows]: bit 31 = the query's own token also sits in its look-back chunk
  // per (tile parity, phase): what the epilogue needs: float max[128], int slot[128], int4 {row_bh, round_start, -, -}
  static constexpr int kOffFin = kOffDup + 4 * kQRows * 4;
  static constexpr int kFinMax = 0, kFinSlot = kQRows * 4, kFinRow = 2 * kQRows * 4;
  static constexpr int kFinBytes = 2 * kQRows * 4 + 16;
  static constexpr int kOffStage = kOffFin + 4 * kFinBytes;          // epilogue staging: 4 warps x 32 rows x 128 B (swizzled) for coalesced stores
  static constexpr int kOffBar = kOffStage + 4 * 4096;               // 2*kSlots + 8 mbarriers
  static constexpr int kOffTmem = kOffBar + (2 * kSlots + 8) * 8;
  static constexpr int kTotal = kOffTmem + 8;
  static constexpr int kDynamic = kTotal;                            // no static shared memory in the kernel: the dynamic segment starts 1024-B aligned (checked)
  static_assert(kTotal <= 232448, "shared memory budget of one CTA (227 KB)");
};

#ifdef RTTS_TRACE      // timeline build (tools/trace_fwd.py): RTTS_DEFS=-DRTTS_TRACE python reformer_tts_b200/csrc/build.py -f
#define RTTS_STAMP(role, n, k) do { if (p.trace != nullptr && blockIdx.x == 0 && (n) < 32) p.trace[((role) * 32 + (n)) * 8 + (k)] = clock64(); } while (0)
#else
#define RTTS_STAMP(role, n, k) do { } while (0)
#endif

// TMEM column (relative to the pair's region) of the P block of key chunk q (32 keys -> 16 columns of bf16 pairs).
template <int BUCKET>
__device__ __forceinline__ constexpr uint32_t p_col(int q) {
  return BUCKET == 64 ? 32u * q : (q < 4 ? 16u * q : 128u + 16u * (q - 4));
}

// One 16-column chunk of the softmax for one query row: scores r -> e = exp2(s * key_scale + neg_m), zero where the key is masked
// (MASK) or is the query itself (SELF); accumulates the row sum and packs the chunk into 8 registers of bf16 pairs.
// EXACT: neg_m is minus the true row maximum and masked / self entries take the reference's fill values instead of zero.
// a_pos / a_scale: shared addresses of key_pos / key_scale at this chunk's first column.
// DUP: also record whether a key of this chunk is the query's own token (look-back chunk of the first tile of a hash round).
template <bool MASK, bool SELF, bool EXACT, bool DUP = false>
__device__ __forceinline__ void soft_chunk(uint32_t* r, uint32_t a_pos, uint32_t a_scale, float neg_m, int q_limit, int q_enc, float mv,
                                           float sv, float* sum4, uint32_t* pk, uint32_t* dup = nullptr) {
  // Written as whole-chunk stages over r[] in place (16 independent elements per stage), so that every stage has 16-way
  // instruction-level parallelism and no stage waits on the previous element's latency.
  float* x = reinterpret_cast<float*>(r);
  {
    uint4 s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = lds128(a_scale + q * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (EXACT) {
        x[q * 4 + 0] *= __uint_as_float(s[q].x); x[q * 4 + 1] *= __uint_as_float(s[q].y);
        x[q * 4 + 2] *= __uint_as_float(s[q].z); x[q * 4 + 3] *= __uint_as_float(s[q].w);
      } else {
        x[q * 4 + 0] = fmaf(x[q * 4 + 0], __uint_as_float(s[q].x), neg_m); x[q * 4 + 1] = fmaf(x[q * 4 + 1], __uint_as_float(s[q].y), neg_m);
        x[q * 4 + 2] = fmaf(x[q * 4 + 2], __uint_as_float(s[q].z), neg_m); x[q * 4 + 3] = fmaf(x[q * 4 + 3], __uint_as_float(s[q].w), neg_m);
      }
    }
  }
  if (EXACT) {
    uint4 kq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) kq[q] = lds128(a_pos + q * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kp[4] = {static_cast<int>(kq[q].x), static_cast<int>(kq[q].y), static_cast<int>(kq[q].z), static_cast<int>(kq[q].w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float sc = x[q * 4 + i];
        sc = kp[i] > q_limit ? mv : sc;
        sc = kp[i] == q_enc ? sv : sc;
        x[q * 4 + i] = sc + neg_m;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = exp2f(x[i]);
  if (!EXACT && (MASK || SELF)) {
    uint4 kq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) kq[q] = lds128(a_pos + q * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int kp[4] = {static_cast<int>(kq[q].x), static_cast<int>(kq[q].y), static_cast<int>(kq[q].z), static_cast<int>(kq[q].w)};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (MASK) x[q * 4 + i] = kp[i] > q_limit ? 0.f : x[q * 4 + i];       // exp2(mask_value - m) == 0
        if (SELF) x[q * 4 + i] = kp[i] == q_enc ? 0.f : x[q * 4 + i];        // exp2(self_value - m) == 0
        if (DUP) *dup = kp[i] == q_enc ? 0x80000000u : *dup;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) sum4[i & 3] += x[i];
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
}

// The same chunk with the masks applied AFTER the exponentials, on the packed bf16 pairs, two keys per instruction: positions
// below 2048 are exact in fp16 (a padded key holds +inf), so `key position > query limit` is one HSET2 producing a 0xffff / 0
// mask per key and clearing P is one LOP3 - a quarter of the ISETP + FSEL pairs of the fp32 form.  The query's own column is a
// fixed column of the tile (key column BUCKET + row), found by comparing the constant column indices with `self_col` (the own
// column relative to this chunk; anything outside 0..15 never matches).  No row sum here: it comes out of the tensor pipe (P . 1).
template <bool MASK, bool SELF>
__device__ __forceinline__ void soft_chunk_packed(uint32_t* r, uint32_t a_pos16, uint32_t a_scale, float neg_m, uint32_t q_limit2, int self_col, uint32_t* pk) {
  float* x = reinterpret_cast<float*>(r);
  {
    uint4 s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = lds128(a_scale + q * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {      // packed fp32 FMA: two scores per instruction
      ffma2(x[q * 4 + 0], x[q * 4 + 1], x[q * 4 + 0], x[q * 4 + 1], __uint_as_float(s[q].x), __uint_as_float(s[q].y), neg_m, neg_m);
      ffma2(x[q * 4 + 2], x[q * 4 + 3], x[q * 4 + 2], x[q * 4 + 3], __uint_as_float(s[q].z), __uint_as_float(s[q].w), neg_m, neg_m);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = exp2f(x[i]);
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
  if (MASK) {
    const uint4 k0 = lds128(a_pos16), k1 = lds128(a_pos16 + 16);
    const uint32_t kp[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const __half2 ql = *reinterpret_cast<const __half2*>(&q_limit2);
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] &= ~__hgt2_mask(*reinterpret_cast<const __half2*>(&kp[i]), ql);      // exp2(mask_value - m) == 0
  }
  if (SELF) {
    const __half2 sc = __half2half2(__int2half_rn(self_col));
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] &= ~__heq2_mask(__floats2half2_rn(2.f * i, 2.f * i + 1.f), sc);       // exp2(self_value - m) == 0
  }
}

// Row maximum of one 16-column chunk with the reference's fill values (exact mode, first pass).
__device__ __forceinline__ float chunk_max(const uint32_t* r, uint32_t a_pos, uint32_t a_scale, int q_limit, int q_enc, float mv, float sv, float mx) {
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const uint4 s0 = lds128(a_scale + q4 * 16), p0 = lds128(a_pos + q4 * 16);
    const float ks[4] = {__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s0.w)};
    const int kp[4] = {static_cast<int>(p0.x), static_cast<int>(p0.y), static_cast<int>(p0.z), static_cast<int>(p0.w)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float sc = __uint_as_float(r[q4 * 4 + i]) * ks[i];
      sc = kp[i] > q_limit ? mv : sc;
      sc = kp[i] == q_enc ? sv : sc;
      mx = fmaxf(mx, sc);
    }
  }
  return mx;
}

template <int BUCKET>
__global__ void __launch_bounds__(kFwdThreads, 1) lsh_attn_fwd_kernel(const AttnFwdParams p, const int num_tiles) {
  using L = AttnFwdSmem<BUCKET>;
  constexpr int kKeyRows = L::kKeyRows, kSlots = L::kSlots;
  constexpr int kWin = 2 * BUCKET;        // attention window per query
  constexpr int kTail = kQRows - BUCKET;  // first look-back row inside the previous block
  constexpr uint32_t kColO = BUCKET == 64 ? 192 : 64;
  constexpr bool kAliasO = BUCKET == 128; // bucket 128: S fills all 256 columns of the region, O reuses columns S no longer needs
  // Row sums of P come out of the tensor pipe: one more accumulator, D2 = P . 1 (16 identical columns), in S columns that are
  // dead once the softmax has consumed them - the second half of P block 0 (bucket 64) / the tail of the region (bucket 128).
  // It is the sum of the bf16-ROUNDED P, i.e. exactly the normaliser of the O = P V the same instruction stream accumulates.
  constexpr uint32_t kColSum = BUCKET == 64 ? 16 : 192;
  constexpr uint32_t kTmemCols = 512;
#ifndef RTTS_PARTS64
#define RTTS_PARTS64 4
#endif
  constexpr int kParts = BUCKET == 64 ? RTTS_PARTS64 : 2;       // threads per query row in a softmax group
  constexpr int kGroups = kSoftmaxWarps / (4 * kParts);         // softmax groups: group w takes tiles k = w (mod kGroups)
  constexpr int kSoftmaxThreads = 128 * kParts;                 // threads working on one tile

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();     // SWIZZLE_128B tiles need 1024-byte alignment
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
  uint64_t* full = bars;                       // [kSlots]
  uint64_t* slot_free = bars + kSlots;         // [kSlots]
  uint64_t* s_full = bars + 2 * kSlots;        // [2]
  uint64_t* p_full = s_full + 2;               // [2]
  uint64_t* o_full = s_full + 4;               // [2]
  uint64_t* o_free = s_full + 6;               // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int RT = p.R * p.T;
  const uint32_t sbase = smem_u32(smem);       // shared-window address of the dynamic segment, computed once (per-tile code adds offsets)
#ifdef RTTS_TRACE
  const long long t_cta0 = clock64();
#endif
  // this CTA's run of tiles
  const int g0 = static_cast<int>(static_cast<int64_t>(blockIdx.x) * num_tiles / gridDim.x);
  const int g1 = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * num_tiles / gridDim.x);
  const int my_tiles = g1 - g0;
  // position of a tile inside the problem, advanced incrementally (integer divisions only once per CTA)
  struct Geo {
    int row_bh, t_in, round, t_round;     // (batch*H + head), tile in row, hash round, tile in round
  };
  const int tiles_per_round = p.T / kQRows;
  auto geo_at = [&](int g) {
    Geo x;
    x.row_bh = g / p.tiles_per_row;
    x.t_in = g - x.row_bh * p.tiles_per_row;
    x.round = x.t_in / tiles_per_round;
    x.t_round = x.t_in - x.round * tiles_per_round;
    return x;
  };
  auto geo_next = [&](Geo& x) {
    ++x.t_in;
    if (++x.t_round == tiles_per_round) { x.t_round = 0; ++x.round; }
    if (x.t_in == p.tiles_per_row) { x.t_in = 0; x.round = 0; x.t_round = 0; ++x.row_bh; }
  };

  if (tid == 0) {
    for (int s = 0; s < kSlots; ++s) {
      mbar_init(full + s, kLoaderThreads / 32);  // one arrival per warp (after __syncwarp): 256 per-thread arrivals on one word serialise
      mbar_init(slot_free + s, 1);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full + g, 1);
      mbar_init(p_full + g, kSoftmaxThreads / 32);
      mbar_init(o_full + g, 1);
      mbar_init(o_free + g, kEpiThreads / 32);
    }
    fence_mbar_init();
  }
  if (tid < kSlots) *reinterpret_cast<int*>(smem + L::kOffMeta + tid * L::kMetaBytes + L::kMetaTag) = 0;
  for (int i = tid; i < 256; i += kFwdThreads) reinterpret_cast<uint32_t*>(smem + L::kOffOnes)[i] = 0x3F803F80u;      // bf16 1.0 pairs
  for (int i = tid; i < 4 * kQRows; i += kFwdThreads) reinterpret_cast<uint32_t*>(smem + L::kOffDup)[i] = 0u;
  fence_proxy_async_smem();          // the ones tile is read by the tensor-core (async) proxy
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, kTmemCols);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // all 512 columns are ours, so the allocation starts at lane 0, column 0: addresses below are compile-time constants
  if (*tmem_slot != 0) __trap();
  constexpr uint32_t tmem = 0;

  if (warp == kMmaWarp) {
    // ================================================= MMA issuer: S = Q K^T =======================================
    // Two issuing threads (this warp: S, the next warp: PV) because one thread's chain of barrier waits, descriptor updates and
    // commits for 20 MMAs per tile (~1400 cycles) is longer than the tensor pipe needs for them (768).
    if (elect_one()) {
      constexpr uint32_t idesc_lb = umma_idesc_bf16(128, BUCKET, false, false);
      constexpr uint32_t idesc_main = umma_idesc_bf16(128, kQRows, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const uint32_t k_lo0 = umma_desc_lo(smem_u32(smem + L::kOffK), 16);     // K-major operand (Q / K rows)
      constexpr uint32_t kSlotLo = L::kSlotBytes >> 4, kTailLo = (kTail * 128) >> 4;
      for (int k = 0; k < my_tiles; ++k) {
        const int g = k & 1, st = k % kSlots, sp = st == 0 ? kSlots - 1 : st - 1;
        const uint32_t t_reg = tmem + g * 256;
        RTTS_STAMP(0, k, 4);
        mbar_wait(full + st, (k / kSlots) & 1);       // a fresh tile's look-back rows are part of this (the loader waits for PV(k-1))
        RTTS_STAMP(0, k, 0);
        if (k >= 2) {
          // S(k) overwrites the S/P columns PV(k-2) reads and the row-sum columns (bucket 128: also the O columns) its epilogue reads
          mbar_wait(o_free + g, ((k >> 1) & 1) ^ 1);
        }
        tc_fence_after_sync();
        const uint32_t q_lo = k_lo0 + st * kSlotLo, lb_lo = k_lo0 + sp * kSlotLo + kTailLo;
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) {
          umma_ss_lo(t_reg, q_lo + kk * 2, lb_lo + kk * 2, hi, idesc_lb, kk > 0);
          umma_ss_lo(t_reg + BUCKET, q_lo + kk * 2, q_lo + kk * 2, hi, idesc_main, kk > 0);
        }
        umma_commit(s_full + g);
        RTTS_STAMP(0, k, 1);
      }
    }
  } else if (warp == kMmaWarp + 1) {
    // ================================================= MMA issuer: O = P V =========================================
    if (elect_one()) {
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kDh, false, true);
      constexpr uint32_t idesc_sum = umma_idesc_bf16(128, 16, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      constexpr uint32_t hi_ones = umma_desc_hi_sw128(0);                     // 8-row group stride 0: the 16 rows of B alias one 1 KB atom of ones
      const uint32_t v_lo0 = umma_desc_lo(smem_u32(smem + L::kOffV), 0);      // MN-major operand (V rows)
      const uint32_t ones_lo = umma_desc_lo(smem_u32(smem + L::kOffOnes), 16);
      constexpr uint32_t kSlotLo = L::kSlotBytes >> 4, kTailLo = (kTail * 128) >> 4;
      int t_in_next = (g0 + 1) % p.tiles_per_row;     // tile-in-row of tile m + 1 (fresh when 0)
      for (int m = 0; m < my_tiles; ++m) {
        const int g = m & 1, st = m % kSlots, sp = st == 0 ? kSlots - 1 : st - 1;
        const uint32_t t_reg = tmem + g * 256;
        mbar_wait(p_full + g, (m >> 1) & 1);
        RTTS_STAMP(0, m, 2);
        if (!kAliasO) mbar_wait(o_free + g, ((m >> 1) & 1) ^ 1);      // epilogue of tile m-2 has drained O of this group
        RTTS_STAMP(0, m, 5);
        tc_fence_after_sync();
        const uint32_t v_lb = v_lo0 + sp * kSlotLo + kTailLo, v_main = v_lo0 + st * kSlotLo;
#pragma unroll
        for (int j = 0; j < kKeyRows / 16; ++j) {
          const uint32_t b_lo = (j * 16 < BUCKET) ? v_lb + j * (2048 >> 4) : v_main + (j * 16 - BUCKET) * (128 >> 4);
          umma_ts_lo(t_reg + kColO, t_reg + p_col<BUCKET>(j >> 1) + (j & 1) * 8, b_lo, hi, idesc_o, j > 0);
        }
#pragma unroll
        for (int j = 0; j < kKeyRows / 16; ++j)      // row sums: D2 = P . 1 (N = 16, the narrowest shape of an M = 128 MMA)
          umma_ts_lo(t_reg + kColSum, t_reg + p_col<BUCKET>(j >> 1) + (j & 1) * 8, ones_lo, hi_ones, idesc_sum, j > 0);
        RTTS_STAMP(0, m, 6);
        umma_commit(o_full + g);
        // S(m) (issued by the other warp) completed before the softmax of tile m started, so everything that reads the previous
        // block is covered by this thread's PV(m): release it.  This tile's block is released here only if no tile looks back at it.
        umma_commit(slot_free + sp);
        if (m + 1 >= my_tiles || t_in_next == 0) umma_commit(slot_free + st);
        RTTS_STAMP(0, m, 3);
        if (++t_in_next == p.tiles_per_row) t_in_next = 0;
      }
    }
  } else if (warp >= kFirstEpiWarp && warp < kFirstLoaderWarp) {
    // ================================================= epilogue ===================================================
    const int m = tid - kFirstEpiWarp * 32;         // query row = TMEM lane (warp % 4 selects the lane quarter)
    const int lane = tid & 31;
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const uint32_t a_stage = sbase + L::kOffStage + (warp & 3) * 4096;      // this warp's 32 rows x 128 B
    for (int k = 0; k < my_tiles; ++k) {
      const int g = k & 1;
      const uint32_t ph = (k >> 1) & 1;
      const uint32_t a_fin = sbase + L::kOffFin + (g * 2 + ph) * L::kFinBytes;
      // Only o_full is waited for.  PV(k) is issued after p_full(k), so o_full(k) implies that the softmax group's partial sums and
      // slots are in shared memory; and p_full must NOT be waited on here: p_full(k+2) needs only PV(k) (not this role), so with a
      // slow epilogue the barrier can run two phases ahead of a waiter, whose parity test would then never pass.  o_full cannot:
      // PV(k+2) waits for this role's o_free(k).
      if (lane == 0) mbar_wait_relaxed(o_full + g, ph, 40);      // one polling lane per warp
      __syncwarp();
      tc_fence_after_sync();
      uint32_t rs;
      tmem_ld1(t_lane + g * 256 + kColSum, &rs);                   // row sum of the (bf16) P row, from the tensor pipe
      const uint4 fin = lds128(a_fin + L::kFinRow);                // {row_bh, first tile of a hash round, -, -}
      const int row_bh = static_cast<int>(fin.x);
      uint32_t dup = 0;
      if (fin.y != 0) {                                           // only there can the own token sit in the look-back chunk a second time
        const uint32_t a_dup = sbase + L::kOffDup + ((g * 2 + ph) * kQRows + m) * 4;
        dup = lds32(a_dup);
        if (dup != 0) sts32(a_dup, 0u);                          // (this role owns the row's flag once o_full has fired)
      }
      float row_max = __uint_as_float(lds32(a_fin + L::kFinMax + m * 4));
      const int64_t row_base = static_cast<int64_t>(row_bh) * RT;
      const int own_slot = static_cast<int>(lds32(a_fin + L::kFinSlot + m * 4));
      tmem_ld_wait();
      float row_sum = __uint_as_float(rs);
      // all terms exactly zero: the row sees only itself (see the softmax role); exact-mode rows always have a positive sum
      const bool lonely = !(row_sum > 0.f);
      if (lonely) {
        row_sum = (dup >> 31) ? 2.f : 1.f;      // number of self columns
        row_max = p.self_value_log2;
      }
      const float inv_sum = 1.f / row_sum;
      if (m == 0) RTTS_STAMP(3, k, 0);
      // O row / row sum -> bf16 -> this warp's staging tile (row = lane, 16-byte chunks swizzled by the row)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t r[32];
        tmem_ld32(t_lane + g * 256 + kColO + hh * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(r[q4 * 8 + 0]) * inv_sum, __uint_as_float(r[q4 * 8 + 1]) * inv_sum);
          u.y = pack_bf16(__uint_as_float(r[q4 * 8 + 2]) * inv_sum, __uint_as_float(r[q4 * 8 + 3]) * inv_sum);
          u.z = pack_bf16(__uint_as_float(r[q4 * 8 + 4]) * inv_sum, __uint_as_float(r[q4 * 8 + 5]) * inv_sum);
          u.w = pack_bf16(__uint_as_float(r[q4 * 8 + 6]) * inv_sum, __uint_as_float(r[q4 * 8 + 7]) * inv_sum);
          sts128(a_stage + lane * 128 + (((hh * 4 + q4) ^ (lane & 7)) << 4), u);
        }
      }
      tc_fence_before_sync();
      __syncwarp();                 // staging tile complete; all TMEM reads of this warp done
      if (lane == 0) mbar_arrive(o_free + g);      // O columns of this group may be overwritten
      if (__any_sync(0xffffffffu, lonely)) {
        if (lonely) {
          // out = v[own position] (every self column holds the query's own token); P of this row is all zero, so O was zero
          const int pos = own_slot % p.T, b = row_bh / p.H, h = row_bh - b * p.H;
          const uint4* vrow = reinterpret_cast<const uint4*>(p.v + (static_cast<int64_t>(b) * p.T + pos) * p.ld + h * kDh);
          uint4 vv[8];
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) vv[ch] = __ldg(vrow + ch);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) sts128(a_stage + lane * 128 + ((ch ^ (lane & 7)) << 4), vv[ch]);
        }
        __syncwarp();
      }
      // scatter-store at the UNSORTED slot, one full 128-byte row per 8 lanes (four rows per instruction)
      const int64_t my_slot = row_base + own_slot;
      p.lse_rounds[my_slot] = (row_max + log2f(row_sum)) * kLn2;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = it * 4 + (lane >> 3), ch = lane & 7;
        const uint4 u = lds128(a_stage + row * 128 + ((ch ^ (row & 7)) << 4));
        const int64_t slot = row_base + __shfl_sync(0xffffffffu, own_slot, row);      // (registers: the shared copy may be rewritten by tile k+4 once o_free is signalled)
        reinterpret_cast<uint4*>(p.o_rounds + slot * kDh)[ch] = u;
      }
      __syncwarp();                 // the staging tile is free again
      if (m == 0) RTTS_STAMP(3, k, 1);
    }
  } else if (warp >= kFirstLoaderWarp && warp < kMmaWarp) {
    // ================================================= loaders ====================================================
    // sticker[slot] = round*T + pos and sorted slots keep the rounds contiguous, so pos = sticker - (slot / T) * T with the
    // round taken from the slot index (no integer division on the load's critical path).
    const int lg = (tid - kFirstLoaderWarp * 32) / kLoaderThreads;      // loader group: tiles lg, lg + kLoaderGroups, ...
    const int lt = tid - kFirstLoaderWarp * 32 - lg * kLoaderThreads;
    const int grp = lt >> 3, c = lt & 7;         // row groups of 8 lanes; lane c owns 16-byte chunk c of a row
    constexpr int kGroups = kLoaderThreads / 8;  // rows per pass
    constexpr int kPasses = kQRows / kGroups;    // lane c also owns the metadata of passes c (and c + 8 when there are 16 passes)
    constexpr int kLbPasses = BUCKET / kGroups;
    constexpr int kMetaPerLane = (kPasses + 7) / 8, kLbMetaPerLane = (kLbPasses + 7) / 8;
    auto load_stickers = [&](int k, const Geo& x, int* st) {   // main rows of local tile k
      if (k >= my_tiles) return;
      const int32_t* stk = p.sticker + static_cast<int64_t>(x.row_bh) * RT + x.t_in * kQRows + grp;
#pragma unroll
      for (int i = 0; i < kPasses; ++i) st[i] = __ldg(stk + i * kGroups);
    };
    auto load_meta = [&](int k, const Geo& x, int b, int pos, float& ssq, uint32_t& valid) {
      valid = 1u;
      ssq = 1.f;
      if (k >= my_tiles) return;
      ssq = __ldg(p.sumsq + static_cast<int64_t>(x.row_bh) * p.T + pos);
      if (p.mask != nullptr) valid = __ldg(p.mask + static_cast<int64_t>(b) * p.T + pos);
    };
    auto key_scale_of = [&](float ssq) {
      float inv;
      if (p.key_norm == RTTS_KEYNORM_L2) inv = 1.f / fmaxf(sqrtf(ssq), 1e-12f);
      else inv = rsqrtf(ssq * (1.f / kDh) + 1e-6f) * 0.125f;   // 1/sqrt(64)
      return inv * p.score_scale_log2;
    };
    auto select_pass = [&](const int* a, int which) {       // a[which] with a static register index (which = c or c + 8)
      int x = 0;
#pragma unroll
      for (int i = 0; i < kPasses; ++i) x = (i == which) ? a[i] : x;
      return x;
    };
    const uint32_t ld32 = static_cast<uint32_t>(p.ld);      // 32-bit row offsets: one IMAD + one IMAD.WIDE per copy instead of 64-bit multiplies
    int pos_cur[kPasses], st_nxt[kPasses], st_nn[kPasses];
    float ssq_cur[2] = {1.f, 1.f}, ssq_nxt[2] = {1.f, 1.f};
    uint32_t valid_cur[2] = {1u, 1u}, valid_nxt[2] = {1u, 1u};
    constexpr int G = kLoaderGroups;
    auto geo_skip = [&](Geo& x) {
#pragma unroll
      for (int i = 0; i < G; ++i) geo_next(x);
    };
    Geo x_cur = geo_at(g0 + lg), x_nxt = x_cur, x_nn;      // tiles k, k + G, k + 2G of this group
    geo_skip(x_nxt);
    x_nn = x_nxt;
    geo_skip(x_nn);
    int b_cur = x_cur.row_bh / p.H, b_nxt = x_nxt.row_bh / p.H;      // batch index (mask row) of tiles k and k + G
    if (lg < my_tiles) {
      load_stickers(lg, x_cur, pos_cur);
      load_stickers(lg + G, x_nxt, st_nxt);
      const int base0 = x_cur.round * p.T;
#pragma unroll
      for (int i = 0; i < kPasses; ++i) pos_cur[i] -= base0;
#pragma unroll
      for (int q = 0; q < kMetaPerLane; ++q)
        if (c + 8 * q < kPasses) load_meta(lg, x_cur, b_cur, select_pass(pos_cur, c + 8 * q), ssq_cur[q], valid_cur[q]);
    }
    // bit s: parity the next wait on slot_free[s] uses (a first wait passes on a fresh barrier).  A parity wait is only meaningful
    // for a waiter that is at most one phase behind, so every group performs the waits of ALL tiles in tile order - also of the tiles
    // another group gathers (those have long passed or pass together with its own next wait).
    uint32_t free_parity = 0xffffffffu;
    int pending = -1;                      // slot whose copies are in flight and not yet announced
    auto announce = [&](int st_i) {
      fence_proxy_async_smem();            // cp.async / st.shared data -> visible to the tensor-core (async) proxy
      __syncwarp();
      if ((tid & 31) == 0) mbar_arrive(full + st_i);
    };
    auto wait_free = [&](int s) {
      const uint32_t par = (free_parity >> s) & 1u;
      if (!mbar_try_wait(slot_free + s, par)) {      // would block:
        if (pending >= 0) {                          // do not sit on a tile that has already landed
          cp_async_wait<0>();
          announce(pending);
          pending = -1;
        }
        mbar_wait_relaxed(slot_free + s, par, 100);
      }
      free_parity ^= 1u << s;
    };
    auto wait_tile = [&](int k, bool fresh) {      // the waits tile k's loader performs
      const int s = k % kSlots;
      if (fresh) wait_free(s == 0 ? kSlots - 1 : s - 1);
      wait_free(s);
    };
    {
      Geo w = geo_at(g0);
      for (int k = 0; k < lg && k < my_tiles; ++k) { wait_tile(k, k == 0 || w.t_in == 0); geo_next(w); }
    }
    for (int k = lg; k < my_tiles; k += G) {
      const int st_i = k % kSlots, sp_i = (st_i + kSlots - 1) % kSlots;
      uint8_t* slot = smem + st_i * L::kSlotBytes;
      const int row_bh = x_cur.row_bh, t_in = x_cur.t_in;
      const int b = b_cur, h = row_bh - b * p.H;
      const __nv_bfloat16* qk_b = p.qk + static_cast<int64_t>(b) * p.T * p.ld + h * kDh + c * 8;
      const __nv_bfloat16* v_b = p.v + static_cast<int64_t>(b) * p.T * p.ld + h * kDh + c * 8;
      if (lt == 0) RTTS_STAMP(1, k, 0);
      load_stickers(k + 2 * G, x_nn, st_nn);
      if (k + G < my_tiles) {
        const int base1 = x_nxt.round * p.T;
#pragma unroll
        for (int i = 0; i < kPasses; ++i) st_nxt[i] -= base1;
      }
#pragma unroll
      for (int q = 0; q < kMetaPerLane; ++q)
        if (c + 8 * q < kPasses) load_meta(k + G, x_nxt, b_nxt, select_pass(st_nxt, c + 8 * q), ssq_nxt[q], valid_nxt[q]);
      if (lt == 0) RTTS_STAMP(1, k, 1);
      if (k == 0 || t_in == 0) {
        // no predecessor in the ring: gather the look-back rows (the BUCKET sorted slots before this tile, wrapping to the end
        // of the row) into the tail of the previous slot once the tile that lived there has been consumed
        wait_free(sp_i);
        uint8_t* prev = smem + sp_i * L::kSlotBytes;
        const uint32_t sKp = smem_u32(prev + L::kOffK), sVp = smem_u32(prev + L::kOffV);
        uint8_t* meta_p = smem + L::kOffMeta + sp_i * L::kMetaBytes;
        int first = t_in * kQRows - BUCKET;
        if (first < 0) first += RT;
        const int base_prev = (first / p.T) * p.T;
        const int32_t* stk = p.sticker + static_cast<int64_t>(row_bh) * RT + first + grp;
        int lb[kLbPasses];
#pragma unroll
        for (int i = 0; i < kLbPasses; ++i) lb[i] = __ldg(stk + i * kGroups) - base_prev;
#pragma unroll
        for (int i = 0; i < kLbPasses; ++i) {
          const int j = kTail + i * kGroups + grp;
          const uint32_t off = static_cast<uint32_t>(lb[i]) * ld32;      // element offset inside the batch entry: T * ld < 2^31 (checked by the host)
          const uint32_t so = sw128_offset(j, c);
          cp_async16(sKp + so, qk_b + off);
          cp_async16(sVp + so, v_b + off);
        }
#pragma unroll
        for (int q = 0; q < kLbMetaPerLane; ++q) {
          if (c + 8 * q >= kLbPasses) break;
          int my_pos = 0;
#pragma unroll
          for (int i = 0; i < kLbPasses; ++i) my_pos = (i == c + 8 * q) ? lb[i] : my_pos;
          float ssq;
          uint32_t valid;
          load_meta(k, x_cur, b, my_pos, ssq, valid);
          const int j = kTail + (c + 8 * q) * kGroups + grp;
          reinterpret_cast<float*>(meta_p + L::kMetaScale)[j] = key_scale_of(ssq);
          reinterpret_cast<int*>(meta_p + L::kMetaPos)[j] = valid ? my_pos : (my_pos | kPadFlag);
          reinterpret_cast<__half*>(meta_p + L::kMetaPos16)[j] = valid ? __int2half_rn(my_pos) : __ushort_as_half(0x7c00);      // padded: +inf
        }
      }
      wait_free(st_i);
      if (lt == 0) RTTS_STAMP(1, k, 2);
      const uint32_t sK = smem_u32(slot + L::kOffK), sV = smem_u32(slot + L::kOffV);
#pragma unroll
      for (int i = 0; i < kPasses; ++i) {
        const int j = i * kGroups + grp;
        const uint32_t off = static_cast<uint32_t>(pos_cur[i]) * ld32;
        const uint32_t so = sw128_offset(j, c);
        cp_async16(sK + so, qk_b + off);
        cp_async16(sV + so, v_b + off);
      }
      cp_async_commit();
      {
        uint8_t* meta = smem + L::kOffMeta + st_i * L::kMetaBytes;
#pragma unroll
        for (int q = 0; q < kMetaPerLane; ++q) {
          if (c + 8 * q >= kPasses) break;
          const int my_pos = select_pass(pos_cur, c + 8 * q), my_row = (c + 8 * q) * kGroups + grp;
          const float ks = key_scale_of(ssq_cur[q]);
          reinterpret_cast<float*>(meta + L::kMetaScale)[my_row] = ks;
          reinterpret_cast<int*>(meta + L::kMetaPos)[my_row] = valid_cur[q] ? my_pos : (my_pos | kPadFlag);
          reinterpret_cast<__half*>(meta + L::kMetaPos16)[my_row] = valid_cur[q] ? __int2half_rn(my_pos) : __ushort_as_half(0x7c00);
          // a query whose score bound could push a visible key below the exp2 underflow: the pair runs this tile in exact mode
          if (p.score_scale_log2 * p.score_scale_log2 / ks * 1.001f >= kExactBound) *reinterpret_cast<volatile int*>(meta + L::kMetaTag) = k + 1;
        }
        if (lt == 0) {
          *reinterpret_cast<int2*>(meta + L::kMetaGeo) = make_int2(row_bh, x_cur.round * p.T);
          *reinterpret_cast<int*>(meta + L::kMetaGeo + 8) = x_cur.t_round == 0;
        }
      }
      if (pending >= 0) {
        cp_async_wait<1>();         // everything but the group just committed has landed
        announce(pending);
      }
      pending = st_i;
      if (lt == 0) RTTS_STAMP(1, k, 3);
#pragma unroll
      for (int i = 0; i < kPasses; ++i) { pos_cur[i] = st_nxt[i]; st_nxt[i] = st_nn[i]; }
      ssq_cur[0] = ssq_nxt[0]; ssq_cur[1] = ssq_nxt[1];
      valid_cur[0] = valid_nxt[0]; valid_cur[1] = valid_nxt[1];
      {
        // the tiles the other groups gather between this one and this group's next
        Geo w = x_cur;
        for (int kk = k + 1; kk < k + G && kk < my_tiles; ++kk) { geo_next(w); wait_tile(kk, w.t_in == 0); }
      }
      x_cur = x_nxt;
      x_nxt = x_nn;
      geo_skip(x_nn);
      b_cur = b_nxt;
      if (x_nxt.t_in < G) b_nxt = x_nxt.row_bh / p.H;       // a new (batch, head) row started within the last G tiles
    }
    if (pending >= 0) {
      cp_async_wait<0>();
      announce(pending);
    }
  } else {
    // ================================================= softmax pairs ==============================================
    // All 16 warps work on the SAME tile (thread = (query row, quarter of its window)) and alternate between the two TMEM
    // regions: while they are in tile k (region k & 1), the tensor pipe runs PV(k-1) and S(k+1) on the other region, so that
    // chain hides under the softmax.  (Two groups owning one region each run their softmax phases in lockstep - they share the
    // issue slots equally and finish together - and then both idle through their PV -> S chains: 2100 of 6500 cycles per tile pair.)
    const int grp_id = warp / (4 * kParts);         // softmax group (bucket 128 only: 0 | 1)
    const int part = (warp >> 2) % kParts;          // which part of the row's window
    const int m = tid & 127;                        // query row = TMEM lane
    const uint32_t t_lane = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    constexpr int kPartCols = kWin / kParts;        // 32 (bucket 64) or 64 (bucket 128) key columns per thread
    const int win0 = (m / BUCKET) * BUCKET;          // first key column of this query's window
    const float mv = p.mask_value_log2, sv = p.self_value_log2;
    const bool need_mask = p.causal || p.mask != nullptr;
    auto pair_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp_id), "n"(kSoftmaxThreads) : "memory"); };
    // the block of 32 key columns holding this warp's own columns (rows 32*(m/32) .. +31) and the quarter that owns it (warp-uniform)
    const int diag_col = BUCKET + (m & ~31);
    // Every instruction of this per-tile prologue is executed by 16 warps (the kernel is bound by instruction issue), so the ring
    // position is carried incrementally instead of being re-derived from k (k % 6, k / 6 are multiply-high sequences).
    const uint32_t a_meta0 = sbase + L::kOffMeta, a_sfull = sbase + L::kOffBar + 2 * kSlots * 8;
    const float bound_c = p.score_scale_log2 * p.score_scale_log2 * 1.001f;
    int st_i = grp_id % kSlots;                     // ring slot of tile k
    for (int k = grp_id; k < my_tiles; k += kGroups) {
      const int wg = k & 1;                         // TMEM region / barrier set of this tile
      const uint32_t ph = (k >> 1) & 1;
      const uint32_t t_row = t_lane + wg * 256;
      const uint32_t a_part = sbase + L::kOffPart + grp_id * kMaxParts * kQRows * 4;      // float [kParts][128 rows]: exact-mode row maxima
      const int sp_i = st_i == 0 ? kSlots - 1 : st_i - 1;
      const uint32_t a_meta = a_meta0 + st_i * L::kMetaBytes, a_meta_p = a_meta0 + sp_i * L::kMetaBytes;
      // key column j of the tile: j < BUCKET -> look-back row kTail + j of the previous slot, else main row j - BUCKET
      const uint32_t a_scale_lb = a_meta_p + L::kMetaScale + kTail * 4, a_pos_lb = a_meta_p + L::kMetaPos + kTail * 4;
      const uint32_t a_scale_mn = a_meta + L::kMetaScale - BUCKET * 4, a_pos_mn = a_meta + L::kMetaPos - BUCKET * 4;
      const uint32_t a_p16_lb = a_meta_p + L::kMetaPos16 + kTail * 2, a_p16_mn = a_meta + L::kMetaPos16 - BUCKET * 2;
      if (m == 0 && part == 0) RTTS_STAMP(2, k, 0);
      st_i += kGroups;                              // (for the next tile)
      if (st_i >= kSlots) st_i -= kSlots;
      // ONE wait per tile, by one lane: S(k) was issued after the MMA thread had seen full[slot], so its completion also certifies
      // the loader's metadata (written long before); 16 warps x 32 lanes polling one barrier word serialise in the barrier unit.
      // (sleeping wait: polling warps would take issue slots from the loader warps)
      if ((tid & 31) == 0) mbar_wait_relaxed_a(a_sfull + wg * 8, ph, 40);
      __syncwarp();
      tc_fence_after_sync();
      if (m == 0 && part == 0) RTTS_STAMP(2, k, 3);
      const uint4 geo = lds128(a_meta + L::kMetaGeo);
      const int row_bh = static_cast<int>(geo.x), base_main = static_cast<int>(geo.y);
      const bool round_start = geo.z != 0;
      const bool exact = static_cast<int>(geo.w) == k + 1;
      const int q_enc = static_cast<int>(lds32(a_meta + L::kMetaPos + m * 4));
      int q_limit = p.causal ? (q_enc & ~kPadFlag) : (kPadFlag - 1);
      if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && (q_enc & kPadFlag)) q_limit = -1;   // padded query: all masked
      // packed position mask: the query's limit as an fp16 pair (causal: its own position; otherwise the largest finite value, which
      // only a padded key's +inf exceeds; a fully masked query: -1)
      uint32_t q_limit2;
      {
        const __half qh = q_limit < 0 ? __float2half_rn(-1.f) : (p.causal ? __int2half_rn(q_limit) : __ushort_as_half(0x7bff));
        const __half2 q2 = __half2half2(qh);
        q_limit2 = *reinterpret_cast<const uint32_t*>(&q2);
      }
      const bool packed = p.pos16 != 0 && !exact;
      // stabiliser: |q_i| * score_scale * log2(e) = score_scale_log2^2 / key_scale[own row] for both key-norm variants ...
      // (key_scale = score_scale_log2 / |x| up to the norm's epsilon), times (1 + 2^-10) so rounding cannot push a score above it
      const float row_bound = __fdividef(bound_c, __uint_as_float(lds32(a_meta + L::kMetaScale + m * 4)));
      if (m == 0 && part == 0) RTTS_STAMP(2, k, 1);

      float row_max = row_bound;
      if (exact) {
        float mx = -FLT_MAX;
#pragma unroll 1
        for (int c0 = part * kPartCols; c0 < (part + 1) * kPartCols; c0 += 16) {
          const int col = win0 + c0;
          uint32_t r[16];
          tmem_ld16(t_row + col, r);
          tmem_ld_wait();
          mx = chunk_max(r, (col < BUCKET ? a_pos_lb : a_pos_mn) + col * 4, (col < BUCKET ? a_scale_lb : a_scale_mn) + col * 4, q_limit, q_enc, mv, sv, mx);
        }
        row_max = mx;
        sts32(a_part + (part * kQRows + m) * 4, __float_as_uint(mx));
        pair_sync();
#pragma unroll
        for (int q = 0; q < kParts; ++q) row_max = q == 0 ? __uint_as_float(lds32(a_part + m * 4)) : fmaxf(row_max, __uint_as_float(lds32(a_part + (q * kQRows + m) * 4)));
        pair_sync();               // every part has read the maxima before the slots are reused for the sums
      }
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
      uint32_t dup = 0;           // bit 31: the query's own token also sits in the look-back chunk (rows with win0 == 0 only count it)
      {
        // The query's own column sits in exactly one 32-column chunk per warp (warp-uniform); the same token can appear a second
        // time only in the look-back chunk of the first tile of a hash round.  Only those chunks pay for the self comparison, and
        // the position mask is skipped altogether when nothing can be masked (non-causal, no padding mask).
        const float neg_m = -row_max;
        if (kParts == 4) {
          // 32 columns per thread = one P block: both TMEM loads are issued before anything is computed, and both chunks share
          // one variant (a 32-column block lies entirely in the look-back or the main part and holds the own column or not)
          const int col = win0 + part * kPartCols;
          uint32_t r0[16], r1[16], pk[8];
          tmem_ld16(t_row + col, r0);
          tmem_ld16(t_row + col + 16, r1);
          const uint32_t a_pos = (col < BUCKET ? a_pos_lb : a_pos_mn) + col * 4, a_scale = (col < BUCKET ? a_scale_lb : a_scale_mn) + col * 4;
          const bool self_chunk = col == diag_col || (round_start && col < BUCKET);
          const uint32_t t_p = t_row + p_col<BUCKET>(col >> 5);
          tmem_ld_wait();
          if (m == 0 && part == 0) RTTS_STAMP(2, k, 4);
          if (exact) {
            soft_chunk<true, true, true>(r0, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            tmem_st8(t_p, pk);
            soft_chunk<true, true, true>(r1, a_pos + 64, a_scale + 64, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
          } else if (self_chunk && col < BUCKET) {       // look-back block of the first tile of a hash round: the own token may be there
            soft_chunk<true, true, false, true>(r0, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk, &dup);
            tmem_st8(t_p, pk);
            soft_chunk<true, true, false, true>(r1, a_pos + 64, a_scale + 64, neg_m, q_limit, q_enc, mv, sv, sum4, pk, &dup);
          } else if (packed) {
            // (everything but the look-back block of a round's first tile, where the own token may sit a second time at an unknown column)
            const uint32_t a_p16 = (col < BUCKET ? a_p16_lb : a_p16_mn) + col * 2;
            const int own = col == diag_col ? (m & 31) : -1;        // own column relative to this 32-column block
            if (own >= 0) {
              if (need_mask) {
                soft_chunk_packed<true, true>(r0, a_p16, a_scale, neg_m, q_limit2, own, pk);
                tmem_st8(t_p, pk);
                soft_chunk_packed<true, true>(r1, a_p16 + 32, a_scale + 64, neg_m, q_limit2, own - 16, pk);
              } else {
                soft_chunk_packed<false, true>(r0, a_p16, a_scale, neg_m, q_limit2, own, pk);
                tmem_st8(t_p, pk);
                soft_chunk_packed<false, true>(r1, a_p16 + 32, a_scale + 64, neg_m, q_limit2, own - 16, pk);
              }
            } else if (need_mask) {
              soft_chunk_packed<true, false>(r0, a_p16, a_scale, neg_m, q_limit2, 0, pk);
              tmem_st8(t_p, pk);
              soft_chunk_packed<true, false>(r1, a_p16 + 32, a_scale + 64, neg_m, q_limit2, 0, pk);
            } else {
              soft_chunk_packed<false, false>(r0, a_p16, a_scale, neg_m, q_limit2, 0, pk);
              tmem_st8(t_p, pk);
              soft_chunk_packed<false, false>(r1, a_p16 + 32, a_scale + 64, neg_m, q_limit2, 0, pk);
            }
          } else if (self_chunk) {
            soft_chunk<true, true, false>(r0, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            tmem_st8(t_p, pk);
            soft_chunk<true, true, false>(r1, a_pos + 64, a_scale + 64, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
          } else if (need_mask) {
            soft_chunk<true, false, false>(r0, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            tmem_st8(t_p, pk);
            soft_chunk<true, false, false>(r1, a_pos + 64, a_scale + 64, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
          } else {
            soft_chunk<false, false, false>(r0, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            tmem_st8(t_p, pk);
            soft_chunk<false, false, false>(r1, a_pos + 64, a_scale + 64, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
          }
          if (m == 0 && part == 0) RTTS_STAMP(2, k, 5);
          tmem_st8(t_p + 8, pk);        // P over S columns this thread has already consumed
        } else {
#pragma unroll 1
          for (int c0 = part * kPartCols; c0 < (part + 1) * kPartCols; c0 += 16) {
            const int col = win0 + c0;
            uint32_t r[16], pk[8];
#ifdef RTTS_TRACE
            const long long t_a = clock64();
#endif
            tmem_ld16(t_row + col, r);
            const uint32_t a_pos = (col < BUCKET ? a_pos_lb : a_pos_mn) + col * 4, a_scale = (col < BUCKET ? a_scale_lb : a_scale_mn) + col * 4;
            const bool self_chunk = (col & ~31) == diag_col || (round_start && col < BUCKET);
            tmem_ld_wait();
#ifdef RTTS_TRACE
            const long long t_b = clock64();
#endif
            if (exact) soft_chunk<true, true, true>(r, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            else if (self_chunk && col < BUCKET) soft_chunk<true, true, false, true>(r, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk, &dup);
            else if (packed) {
              const uint32_t a_p16 = (col < BUCKET ? a_p16_lb : a_p16_mn) + col * 2;
              const int own = BUCKET + m - col;      // own column relative to this chunk (matches only inside 0..15)
              if (self_chunk && need_mask) soft_chunk_packed<true, true>(r, a_p16, a_scale, neg_m, q_limit2, own, pk);
              else if (self_chunk) soft_chunk_packed<false, true>(r, a_p16, a_scale, neg_m, q_limit2, own, pk);
              else if (need_mask) soft_chunk_packed<true, false>(r, a_p16, a_scale, neg_m, q_limit2, 0, pk);
              else soft_chunk_packed<false, false>(r, a_p16, a_scale, neg_m, q_limit2, 0, pk);
            }
            else if (self_chunk) soft_chunk<true, true, false>(r, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            else if (need_mask) soft_chunk<true, false, false>(r, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
            else soft_chunk<false, false, false>(r, a_pos, a_scale, neg_m, q_limit, q_enc, mv, sv, sum4, pk);
#ifdef RTTS_TRACE
            const long long t_c = clock64();
#endif
            tmem_st8(t_row + p_col<BUCKET>(col >> 5) + ((col >> 4) & 1) * 8, pk);      // P over S columns this thread has already consumed
#ifdef RTTS_TRACE
            if (p.trace != nullptr && blockIdx.x == 0 && k < 32 && m == 0 && part == 0) {
              p.trace[(2 * 32 + k) * 8 + 6] += t_b - t_a;               // TMEM load + wait
              p.trace[(2 * 32 + k) * 8 + 4] += t_c - t_b;               // chunk arithmetic
              p.trace[(2 * 32 + k) * 8 + 7] += clock64() - t_c;         // TMEM store issue
            }
#endif
          }
        }
      }
      if (BUCKET == 64) {
        // the 64 keys outside this query's window contribute nothing: zero their two P blocks
        const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const int dead_q = m < 64 ? 4 : 0;
        if (part < 2) tmem_st16(t_row + p_col<BUCKET>(dead_q + part), z);      // parts 0 / 1: the same blocks they own in the duplicate scan below
      }
      // No exchange between the parts of a row: each leaves its partial row sum (>= 0; bit 31 = duplicate of the own token seen in
      // the look-back chunk) for the epilogue thread of the row, which adds them up.  A row whose terms are all exactly zero sees
      // only itself (with bound < 60 a visible key cannot underflow: s - bound >= -2*bound > -126): its P row stays zero and the
      // epilogue substitutes the exact result (rp R8 "except when no other targets are available": softmax uniform over the self
      // columns, all of which hold the query's own token, so out = v[own position], lse = self_value + log(#self columns)).
      // No exchange between the parts of a row and no row sum here: the sum of the bf16 P row comes out of the tensor pipe (P . 1,
      // issued with PV) and is read by the epilogue.  A row whose terms are all exactly zero sees only itself (with bound < 60 a
      // visible key cannot underflow: s - bound >= -2*bound > -126): its P row stays zero and the epilogue substitutes the exact
      // result (rp R8 "except when no other targets are available": softmax uniform over the self columns, all of which hold the
      // query's own token, so out = v[own position], lse = self_value + log(#self columns)); `dup` tells it about a second self column.
      if (round_start && win0 == 0 && dup != 0) sts32(sbase + L::kOffDup + ((wg * 2 + ph) * kQRows + m) * 4, dup);
      if (part == 0) {
        // (buffers are per (tile parity, phase): tile k+4 writes the same ones, and its S is issued only after epilogue(k+2) - hence
        // epilogue(k) - has signalled o_free)
        const uint32_t a_fin = sbase + L::kOffFin + (wg * 2 + ph) * L::kFinBytes;
        sts32(a_fin + L::kFinMax + m * 4, __float_as_uint(row_max));
        sts32(a_fin + L::kFinSlot + m * 4, static_cast<uint32_t>(base_main + (q_enc & ~kPadFlag)));      // unsorted slot = round * T + position
        if (m == 0) sts128(a_fin + L::kFinRow, make_uint4(static_cast<uint32_t>(row_bh), round_start ? 1u : 0u, 0u, 0u));
      }
      if (m == 0 && part == 0) RTTS_STAMP(2, k, 6);
      tmem_st_wait();
      if (m == 0 && part == 0) RTTS_STAMP(2, k, 7);
      tc_fence_before_sync();       // this thread's TMEM reads of S / writes of P precede the MMAs that consume / overwrite the region
      __syncwarp();
      if ((tid & 31) == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_sfull + (2 + wg) * 8) : "memory");      // p_full[wg]
      if (m == 0 && part == 0) RTTS_STAMP(2, k, 2);
    }
  }
  // debug: per-CTA time until each role is done (trace buffer rows after the 4*32*8 stamps)
#ifdef RTTS_TRACE
  if (p.trace != nullptr && (tid & 31) == 0) p.trace[4 * 32 * 8 + blockIdx.x * 32 + warp] = clock64() - t_cta0;
#endif
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------
// Round merge (rp R11): out[t] = sum_r o_r[t] * exp(lse_r[t] - logsumexp_r lse_r[t]).  8 lanes per (b,h,t) row,
// each lane owns 16 bytes of the 128-byte head slice.  HBM-bound: reads R*(128+4) B, writes 128+4 B per row.
// ------------------------------------------------------------------------------------------------------------
// RC > 0: the number of rounds is a compile-time constant, so all R lse values and all R 16-byte slices of a row are requested
// before anything is computed (R loads in flight per lane instead of one at a time).  RC == 0: any R, rolled loops.
template <int RC>
__global__ void __launch_bounds__(256) lsh_merge_fwd_kernel(const __nv_bfloat16* __restrict__ o_rounds,
                                                            const float* __restrict__ lse_rounds,
                                                            __nv_bfloat16* __restrict__ out, int64_t ld_out,
                                                            float* __restrict__ lse, int T, int H, int R, int64_t rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 3;   // (b*H + h)*T + t
  const int c = threadIdx.x & 7;
  if (row >= rows) return;
  const int64_t bh = row / T;
  const int t = static_cast<int>(row - bh * T);
  const float* l = lse_rounds + bh * R * T + t;
  const uint4* o = reinterpret_cast<const uint4*>(o_rounds + (bh * R * T + t) * kDh) + c;
  const int64_t o_step = static_cast<int64_t>(T) * (kDh * 2 / 16);      // uint4 elements between rounds
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float mx = -FLT_MAX, den = 0.f;
  auto add_row = [&](float w, const uint4& u) {
    acc[0] = fmaf(w, bf16_lo(u.x), acc[0]); acc[1] = fmaf(w, bf16_hi(u.x), acc[1]);
    acc[2] = fmaf(w, bf16_lo(u.y), acc[2]); acc[3] = fmaf(w, bf16_hi(u.y), acc[3]);
    acc[4] = fmaf(w, bf16_lo(u.z), acc[4]); acc[5] = fmaf(w, bf16_hi(u.z), acc[5]);
    acc[6] = fmaf(w, bf16_lo(u.w), acc[6]); acc[7] = fmaf(w, bf16_hi(u.w), acc[7]);
  };
  // Weights as the reference forms them (rp R11 / hf:626-645): probs = exp(lse_r - logsumexp_r lse_r) with the total rounded to
  // fp32 FIRST.  For a row that sees only itself in every round lse_r = self_value (-5e4 / -1e5, fp32 spacing 4e-3 / 8e-3), the
  // rounded total is off by up to half a spacing and the reference's weights then sum to 1 +- 4e-3: reproduced, not "fixed".
  float tot;
  if (RC > 0) {
    float lv[RC > 0 ? RC : 1];
    uint4 ov[RC > 0 ? RC : 1];
#pragma unroll
    for (int r = 0; r < RC; ++r) lv[r] = __ldg(l + static_cast<int64_t>(r) * T);
#pragma unroll
    for (int r = 0; r < RC; ++r) ov[r] = __ldg(o + r * o_step);
#pragma unroll
    for (int r = 0; r < RC; ++r) mx = fmaxf(mx, lv[r]);
#pragma unroll
    for (int r = 0; r < RC; ++r) den += __expf(lv[r] - mx);
    tot = mx + __logf(den);
#pragma unroll
    for (int r = 0; r < RC; ++r) add_row(__expf(lv[r] - tot), ov[r]);
  } else {
    for (int r = 0; r < R; ++r) mx = fmaxf(mx, l[static_cast<int64_t>(r) * T]);
    for (int r = 0; r < R; ++r) den += __expf(l[static_cast<int64_t>(r) * T] - mx);
    tot = mx + __logf(den);
    for (int r = 0; r < R; ++r) add_row(__expf(l[static_cast<int64_t>(r) * T] - tot), __ldg(o + r * o_step));
  }
  const int64_t b = bh / H;
  const int h = static_cast<int>(bh - b * H);
  uint4 u;
  u.x = pack_bf16(acc[0], acc[1]); u.y = pack_bf16(acc[2], acc[3]);
  u.z = pack_bf16(acc[4], acc[5]); u.w = pack_bf16(acc[6], acc[7]);
  reinterpret_cast<uint4*>(out + (b * T + t) * ld_out + h * kDh)[c] = u;
  if (c == 0) lse[row] = tot;
}

static long long* g_fwd_trace = nullptr;   // debug only (rtts_debug_set_fwd_trace)
static int g_fwd_tile_kernel = 0;          // debug only (rtts_debug_set_fwd_kernel): 1 = bucket 64 on the tile kernel of this file

template <int BUCKET>
int launch_attn_fwd(const AttnFwdParams& p, int ctas, cudaStream_t stream) {
  using L = AttnFwdSmem<BUCKET>;
  static bool configured = false;   // idempotent attribute set; benign if raced
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lsh_attn_fwd_kernel<BUCKET>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = ctas < kNumSMs ? ctas : kNumSMs;     // persistent: one CTA per SM
  lsh_attn_fwd_kernel<BUCKET><<<grid, kFwdThreads, L::kDynamic, stream>>>(p, ctas);
  return check_launch("rtts_lsh_attn_fwd");
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_lsh_attn_fwd(const void* qk, const void* v, int64_t ld, const int32_t* sticker, const float* sumsq, const uint8_t* mask,
                                 const rtts_lsh_spec* spec, void* o_rounds, float* lse_rounds, int B, int T, int H, int dh,
                                 int R, int bucket, void* stream) {
  RTTS_REQUIRE(qk && v && sticker && sumsq && spec && o_rounds && lse_rounds, "rtts_lsh_attn_fwd: null pointer");
  RTTS_REQUIRE(dh == kDh, "rtts_lsh_attn_fwd: head size %d unsupported (64 only)", dh);
  RTTS_REQUIRE(bucket == 64 || bucket == 128, "rtts_lsh_attn_fwd: bucket size %d unsupported (64 or 128)", bucket);
  RTTS_REQUIRE(T % (2 * bucket) == 0, "rtts_lsh_attn_fwd: T=%d must be a multiple of 2*bucket", T);
  RTTS_REQUIRE(ld % 8 == 0 && ((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) |
                                reinterpret_cast<uintptr_t>(o_rounds)) & 15) == 0,
               "rtts_lsh_attn_fwd: tensors must be 16-byte aligned");
  RTTS_REQUIRE(static_cast<int64_t>(T) < kPadFlag && static_cast<int64_t>(T) * ld < (1ll << 31), "rtts_lsh_attn_fwd: T * ld must be below 2^31");
  AttnFwdParams p;
  p.qk = static_cast<const __nv_bfloat16*>(qk);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.ld = ld;
  p.sticker = sticker;
  p.sumsq = sumsq;
  p.mask = mask;
  p.o_rounds = static_cast<__nv_bfloat16*>(o_rounds);
  p.lse_rounds = lse_rounds;
  p.trace = g_fwd_trace;
  p.T = T; p.H = H; p.R = R;
  p.tiles_per_row = R * T / kQRows;
  p.score_scale_log2 = spec->score_scale * kLog2e;
  p.mask_value_log2 = fmaxf(spec->mask_value * kLog2e, -3.0e38f);
  p.self_value_log2 = spec->self_value * kLog2e;
  p.key_norm = spec->key_norm; p.mask_mode = spec->mask_mode; p.causal = spec->causal;
  p.pos16 = T <= 2048;       // integers up to 2048 are exact in fp16
  const int64_t ctas = static_cast<int64_t>(B) * H * p.tiles_per_row;
  RTTS_REQUIRE(ctas > 0 && ctas < (1ll << 31), "rtts_lsh_attn_fwd: bad grid");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // bucket 64: the block-streaming kernel (lsh_attn_fwd64.cu) wherever its packed fp16 position compare applies (T <= 2048: equal
  // to the tile kernel of this file within 1 % at the training shapes, with 2/3 of its tensor work); beyond, its integer-compare
  // path is 5-8 % behind the tile kernel (tools/ab_fwd.py, profiles/README.md), which keeps those lengths
  if (bucket == 64 && p.pos16 && g_fwd_tile_kernel == 0) return launch_attn_fwd64p(p, B, s);
  if (bucket == 64 && p.pos16 && g_fwd_tile_kernel == 2) return launch_attn_fwd64(p, B, s);
  return bucket == 64 ? launch_attn_fwd<64>(p, static_cast<int>(ctas), s) : launch_attn_fwd<128>(p, static_cast<int>(ctas), s);
}

extern "C" int rtts_lsh_merge_fwd(const void* o_rounds, const float* lse_rounds, void* out, int64_t ld_out, float* lse, int B,
                                  int T, int H, int dh, int R, void* stream) {
  RTTS_REQUIRE(o_rounds && lse_rounds && out && lse, "rtts_lsh_merge_fwd: null pointer");
  RTTS_REQUIRE(dh == kDh, "rtts_lsh_merge_fwd: head size %d unsupported (64 only)", dh);
  RTTS_REQUIRE(ld_out % 8 == 0, "rtts_lsh_merge_fwd: ld_out must be a multiple of 8");
  const int64_t rows = static_cast<int64_t>(B) * H * T;
  const int64_t blocks = (rows * 8 + 255) / 256;
  const auto launch = [&](auto kernel) {
    kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(o_rounds), lse_rounds, static_cast<__nv_bfloat16*>(out), ld_out, lse, T, H, R, rows);
  };
  switch (R) {      // the reference configs use 8 (RP / HF default) and 4 (long-sequence sweep) hash rounds
    case 8: launch(lsh_merge_fwd_kernel<8>); break;
    case 4: launch(lsh_merge_fwd_kernel<4>); break;
    case 2: launch(lsh_merge_fwd_kernel<2>); break;
    default: launch(lsh_merge_fwd_kernel<0>); break;
  }
  return check_launch("rtts_lsh_merge_fwd");
}

// Debug hook (not part of the product ABI): device buffer of 4*32*8 int64 receiving clock64 stamps of CTA 0.
extern "C" void rtts_debug_set_fwd_trace(void* device_buffer) { g_fwd_trace = static_cast<long long*>(device_buffer); }
// Debug hook (not part of the product ABI): 1 = run bucket 64 on the 128-query tile kernel (A/B measurements), 0 = default.
extern "C" void rtts_debug_set_fwd_kernel(int tile_kernel) { g_fwd_tile_kernel = tile_kernel; }
