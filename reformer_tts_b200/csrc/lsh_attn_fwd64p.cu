// Fused chunked shared-QK attention, forward, bucket size 64: the paired-chunk kernel behind rtts_lsh_attn_fwd.
//
// The sorted slots of one (batch, head) row form a cyclic sequence of 64-slot chunks (rp R6: every chunk attends to itself
// and to the chunk before it).  The unit of work is a TILE = two consecutive chunks (e, e+1) as queries = the 128 lanes of one
// MMA, against the three chunks (e-1, e, e+1) that hold their keys:
//     S = [X_e ; X_e+1] [X_e-1 ; X_e ; X_e+1]^T      tcgen05.mma M=128, N=64 (look-back block) + N=128 (main block), K=64
//     P = exp2(S * key_scale - bound[query])         thread = (query row, half of its 128-key window: columns 0-127 for the rows
//                                                    of chunk e, 64-191 for chunk e+1), masks on the packed bf16 pairs, written
//                                                    back over S in TMEM; the 64 columns outside the window are zeroed
//     O = P V,  rowsum = P 1                         tcgen05.mma, A = P from TMEM
//     out = O / rowsum -> bf16                       the SAME threads (FA-4 layout: no hand-over between roles)
// A third of the score MMA is spent on (query, key) blocks outside the windows; the tensor pipe has that room (it is ~30 %
// busy), and in exchange a query row lives on ONE TMEM lane for both of its key chunks: no exchange between lanes, no
// separate epilogue role, no row sum or maximum passed through shared memory.
//
// Roles (24 warps): warps 0-15 = two groups of eight softmax + epilogue warps (group g owns TMEM region g and takes the tiles
// t = g mod 2: while one group waits for its PV the other computes; two warps - window halves - per lane quarter and group, so
// every SM sub-partition runs four of them over the same loop), warps 16-19 K loaders (warp w gathers the K rows + per-row
// metadata of the ring entries e = w mod 4: sticker -> position -> 16-byte cp.async into SWIZZLE_128B chunks), warps 20-21 V
// loaders (entries e = w mod 2), warp 22: one thread issuing the look-back block of S, warp 23: one thread issuing PV + row
// sums of tile t and, chained behind them, the main block of S of tile t+2.
// Ring entries: every owned chunk, preceded by its cyclic predecessor (keys only) at the start of the run and of every
// (batch, head) row (the row's LAST chunk: rp's roll).  Entry e lives in data slot e % 8 and metadata slot e % 32.
// TMEM: region g = columns [256 g, 256 g + 256): S in [0,192), P block of keys 32q..32q+31 in place at [32q, 32q+16),
// row sums at [16,32) (dead once the softmax has consumed them), O at [192,256).
//
// Softmax is single-pass as in lsh_attn_fwd.cu (keys are unit vectors, |q| * scale bounds every score); a chunk with a bound
// >= 60 is flagged by the loader and its rows run the reference's two-pass arithmetic over their whole window.
#include <cfloat>
#include <cuda_fp16.h>

#include "common.cuh"
#include "host_util.h"
#include "lsh_attn_params.h"
#include "rtts_b200.h"

namespace rtts {
namespace f64p {

constexpr int kDh = 64;
constexpr int kC = 64;                  // chunk = bucket size
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kExactBound = 60.f;

constexpr int kSoftWarps = 16;          // two groups of eight: two warps (window halves) per TMEM lane quarter (= warp & 3 = SM sub-partition)
constexpr int kKLoaders = 4, kVLoaders = 2;      // K loader i takes the entries e = i mod 4, V loader i the entries e = i mod 2
constexpr int kLoaderWarps = kKLoaders + kVLoaders;
constexpr int kFirstLoaderWarp = kSoftWarps;
constexpr int kSWarp = kFirstLoaderWarp + kLoaderWarps;
constexpr int kPVWarp = kSWarp + 1;
constexpr int kThreads = (kPVWarp + 1) * 32;      // 768 threads: up to 80 registers each

constexpr int kSlots = 8;               // ring of gathered chunks (K and V rows), an entry is released by the PV that last reads it
constexpr int kMetaSlots = 32;          // ring of per-row metadata, no barrier of its own: entry x + 32 is written by a K loader after the K slot
                                        // of entry x + 24 was released, i.e. after both blocks of S(T') of a tile T' >= tile(x) + 8 were issued.
                                        // Its look-back block waited for o_free(T' - 2) (that group has finished every tile <= T' - 4 in
                                        // program order), its main block was chained behind PV(T' - 2), issued after PV(T' - 3), which
                                        // needed p_full(T' - 3) (the other group has finished every tile <= T' - 5): the readers of entry
                                        // x's metadata - tiles tile(x) and tile(x) + 1, dup scan of the epilogue included - are done
constexpr uint32_t kColO = 192, kColSum = 16;

struct Smem {
  static constexpr int kChunkBytes = kC * 128;                           // 8 KB: 64 rows of one head
  static constexpr int kOffK = 0;                                        // kSlots + 1 chunks: slot kSlots mirrors slot 0, so that the
  static constexpr int kOffV = (kSlots + 1) * kChunkBytes;               //   128 rows of (entry e, entry e+1) are always contiguous
  static constexpr int kOffOnes = kOffV + kSlots * kChunkBytes;          // 1 KB of bf16 1.0: B operand of the row-sum MMA
  static constexpr int kOffMeta = kOffOnes + 1024;
  // per metadata slot
  static constexpr int kMScale = 0;                                      // float[64]  key role: score_scale*log2e / |k|
  static constexpr int kMQ = 256;                                        // uint2[64]  query role: {-bound (float), position as fp16 pair}
  static constexpr int kMPos = 768;                                      // int[64]    position | kPadFlag
  static constexpr int kMPos16 = 1024;                                   // half[64]   position (NaN: padded)
  static constexpr int kMInfo = 1152;                                    // int4 {row_bh, round * T, flags, -}
  static constexpr int kMetaBytes = 1168;
  static constexpr int kOffStage = kOffMeta + kMetaSlots * kMetaBytes;   // output staging: 32 rows x 128 B (swizzled) per softmax warp
  static constexpr int kStageBytes = 32 * 64;
  static constexpr int kOffXch = kOffStage + kSoftWarps * kStageBytes;   // float[2 groups][2 halves][128 rows]: exact mode, row maxima of the halves
  static constexpr int kOffBar = kOffXch + 2 * 2 * 128 * 4;
  static constexpr int kNumBars = 4 * kSlots + 8;
  static constexpr int kOffTmem = kOffBar + kNumBars * 8;
  static constexpr int kTotal = kOffTmem + 16;
  static_assert(kOffOnes % 1024 == 0 && kOffMeta % 16 == 0 && kOffStage % 16 == 0 && kOffBar % 8 == 0, "alignment");
  static_assert(kTotal <= 232448, "shared memory budget of one CTA (227 KB)");
};
constexpr int kFlagRoundStart = 2, kFlagExact = 4;

#ifdef RTTS_TRACE      // timeline build (tools/trace_fwd64p.py): RTTS_DEFS=-DRTTS_TRACE python reformer_tts_b200/csrc/build.py -f
#define P_STAMP(role, n, k) do { if (p.trace != nullptr && blockIdx.x == 0 && (n) < 64) p.trace[((role) * 64 + (n)) * 8 + (k)] = clock64(); } while (0)
#else
#define P_STAMP(role, n, k) do { } while (0)
#endif

__device__ __forceinline__ uint2 lds64(uint32_t saddr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
// Bounded wait (a protocol bug traps instead of hanging the GPU): try_wait with a suspend-time hint, the hardware parks the
// thread until the phase completes.
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
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
  __trap();
}
// one lane waits, the warp follows
__device__ __forceinline__ void warp_wait(uint32_t bar_addr, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait_a(bar_addr, parity);
  __syncwarp();
}
// base + a * b with a 32 x 32 -> 64-bit multiply-add (one IMAD.WIDE): row addresses of the gather and of the scatter-store
__device__ __forceinline__ uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) {
  uint64_t d;
  asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
  return d;
}
__device__ __forceinline__ void cp_async16_g(uint32_t smem_dst, uint64_t gaddr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gaddr) : "memory");
}
__device__ __forceinline__ void stg128(uint64_t gaddr, const uint4& v) {
  asm volatile("st.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(gaddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
// e = exp2(s * key_scale - bound), then ONE compare per pair of keys on the fp16 positions (integers up to 2048 are exact; a
// padded key holds NaN and the compares are the unordered ones, so it is always cleared):
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
#ifdef RTTS_X_NOSOFT      // (ablation builds: tools/ablate_fwd64p.sh - results are wrong by construction, timing only)
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = 0x3c003c00u + (r[i] & 0xff);
  return;
#endif
#ifndef RTTS_X_NOEXP
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = exp2f(x[i]);
#else
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = x[i] * x[i];
#endif
#pragma unroll
  for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(x[2 * i], x[2 * i + 1]);
#ifdef RTTS_X_NOMASK
  return;
#endif
  const uint4 k0 = lds128(a_p16), k1 = lds128(a_p16 + 16);
  const uint32_t kp[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
  const __half2 qp = *reinterpret_cast<const __half2*>(&q_pos2);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __half2 k2 = *reinterpret_cast<const __half2*>(&kp[i]);
    pk[i] &= ~(CAUSAL ? __hgeu2_mask(k2, qp) : __hequ2_mask(k2, qp));
  }
}

// Exact mode: the reference's arithmetic (fill values for masked / self entries, true maximum of the row over its window).
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
  int ok;
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
    nx.ok = 1;
  }
  __device__ Entry next() {
    Entry x = nx;
    if (gc >= g1) {
      x.ok = 0;
      return x;
    }
    if (need_pred) {
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

// Tile t of the run (two owned chunks) -> ring entry of its first chunk.  Tiles never straddle a (batch, head) row (a row has an
// even number of chunks), so a predecessor entry sits in front of the first tile of the run and of every tile that opens a row.
struct TileIter {
  int t_in, tiles_per_row, e;       // tile in row; entry of the tile's first chunk
  __device__ TileIter(int tile0, int tpr) : t_in(tile0 % tpr), tiles_per_row(tpr), e(1) {}
  __device__ bool last_in_row() const { return t_in + 1 == tiles_per_row; }
  __device__ void next() {
    e += 2;
    if (++t_in == tiles_per_row) { t_in = 0; ++e; }
  }
};

__global__ void __launch_bounds__(kThreads, 1) lsh_attn_fwd64p_kernel(const AttnFwdParams p, const int num_tiles) {
  using L = Smem;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();     // SWIZZLE_128B chunks need 1024-byte alignment
  const uint32_t sbase = smem_u32(smem);
  // barriers (shared-window addresses); ring entry e uses index e % kSlots with phase parity (e / kSlots) & 1, tile t of group
  // g = t & 1 uses the group's barrier with parity (t >> 1) & 1
  const uint32_t a_kfull = sbase + L::kOffBar;                 // [kSlots]  K rows + metadata of the entry landed (its K loader warp)
  const uint32_t a_vfull = a_kfull + kSlots * 8;               // [kSlots]  V rows of the entry landed (its V loader warp)
  const uint32_t a_kfree = a_vfull + kSlots * 8;               // [kSlots]  the last S reading the entry's K rows has completed (tcgen05.commit)
  const uint32_t a_vfree = a_kfree + kSlots * 8;               // [kSlots]  the last PV reading the entry's V rows has completed (tcgen05.commit)
  const uint32_t a_sfull = a_vfree + kSlots * 8;               // [2]       tcgen05.commit
  const uint32_t a_pfull = a_sfull + 2 * 8;                    // [2]       the 4 warps of the group
  const uint32_t a_ofull = a_pfull + 2 * 8;                    // [2]       tcgen05.commit
  const uint32_t a_ofree = a_ofull + 2 * 8;                    // [2]       the 4 warps of the group have read O and the row sums
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int RT = p.R * p.T;
  const int cpr = RT / kC;              // chunks per (batch, head) row
  const int cpround = p.T / kC;         // chunks per hash round
  const int tg0 = static_cast<int>(static_cast<int64_t>(blockIdx.x) * num_tiles / gridDim.x);
  const int tg1 = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * num_tiles / gridDim.x);
  const int my_tiles = tg1 - tg0;

  if (tid == 0) {
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    int b = 0;
    for (int s = 0; s < 2 * kSlots; ++s) mbar_init(bars + b++, 1);      // kfull, vfull
    for (int s = 0; s < kSlots; ++s) mbar_init(bars + b++, 2);          // kfree: both S threads
    for (int s = 0; s < kSlots; ++s) mbar_init(bars + b++, 1);          // vfree
    for (int s = 0; s < 2; ++s) mbar_init(bars + b++, 2);               // s_full: both S threads
    for (int s = 0; s < 2; ++s) mbar_init(bars + b++, 8);               // p_full
    for (int s = 0; s < 2; ++s) mbar_init(bars + b++, 1);               // o_full
    for (int s = 0; s < 2; ++s) mbar_init(bars + b++, 8);               // o_free
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
    // ================================================= S, look-back block: [X_e ; X_e+1] X_e-1^T ==================
    // The score MMA of a tile is issued in two parts by two threads.  The main block (columns 64..191) is chained by the PV
    // thread right behind PV(t-2) - the tensor pipe executes one thread's MMAs in issue order, so it needs no barrier and is
    // complete long before the group has finished its epilogue.  The look-back block (columns 0..63) overwrites the row-sum
    // columns of tile t-2, so it waits until the group has read them (o_free) - four MMAs on the path instead of the whole S.
    if (elect_one()) {
      constexpr uint32_t idesc_lb = umma_idesc_bf16(128, kC, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const uint32_t k_lo0 = umma_desc_lo(sbase + L::kOffK, 16);
      uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
      TileIter ti(tg0, p.tiles_per_row);
      int landed = -1;                     // entries 0..landed have been waited for
      for (int t = 0; t < my_tiles; ++t) {
        const int g = t & 1, e = ti.e;
        P_STAMP(0, t, 0);
        for (int x = landed + 1; x <= e + 1; ++x) mbar_wait_a(a_kfull + (x & (kSlots - 1)) * 8, (x / kSlots) & 1);
        landed = e + 1;
        P_STAMP(0, t, 1);
        if (t >= 2) mbar_wait_a(a_ofree + g * 8, ((t >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        P_STAMP(0, t, 2);
        const uint32_t t_reg = g * 256;
        const uint32_t q_lo = k_lo0 + (e & (kSlots - 1)) * (L::kChunkBytes >> 4);            // rows of (e, e+1): slot 7 continues into the mirror of slot 0
        const uint32_t lb_lo = k_lo0 + ((e - 1) & (kSlots - 1)) * (L::kChunkBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) {
#ifdef RTTS_X_NOS
          if (kk > 0) continue;
#endif
          umma_ss_lo(t_reg, q_lo + kk * 2, lb_lo + kk * 2, hi, idesc_lb, kk > 0);
        }
        umma_commit(bars + 4 * kSlots + g);               // s_full (second arrival: the main block is committed by the PV thread)
        // K rows: entries e-1 and e are not read again; e+1 holds the look-back keys of the next tile unless this tile closes the
        // row or the run (both S threads arrive: each commit covers its own thread's MMAs)
        umma_commit(bars + 2 * kSlots + ((e - 1) & (kSlots - 1)));
        umma_commit(bars + 2 * kSlots + (e & (kSlots - 1)));
        if (ti.last_in_row() || t + 1 == my_tiles) umma_commit(bars + 2 * kSlots + ((e + 1) & (kSlots - 1)));
        P_STAMP(0, t, 3);
        ti.next();
      }
    }
  } else if (warp == kPVWarp) {
    // ================================================= O = P V, rowsum = P 1;  S main block of tile t+2 ==============
    if (elect_one()) {
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, kDh, false, true);
      constexpr uint32_t idesc_sum = umma_idesc_bf16(128, 16, false, false);
      constexpr uint32_t idesc_main = umma_idesc_bf16(128, 2 * kC, false, false);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      constexpr uint32_t hi_ones = umma_desc_hi_sw128(0);         // the 16 rows of B alias one 1 KB atom of ones
      const uint32_t v_lo0 = umma_desc_lo(sbase + L::kOffV, 0);   // MN-major operand (V rows)
      const uint32_t k_lo0 = umma_desc_lo(sbase + L::kOffK, 16);
      const uint32_t ones_lo = umma_desc_lo(sbase + L::kOffOnes, 16);
      uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
      TileIter ti(tg0, p.tiles_per_row), ts(tg0, p.tiles_per_row);      // ts: the tile whose main S block is issued next
      int k_landed = -1, v_landed = -1;
      auto issue_s_main = [&](int t2) {      // main block of tile t2: [X_e ; X_e+1] [X_e ; X_e+1]^T -> columns 64..191 of its region
        const int e = ts.e;
        for (int x = k_landed + 1; x <= e + 1; ++x) mbar_wait_a(a_kfull + (x & (kSlots - 1)) * 8, (x / kSlots) & 1);
        k_landed = e + 1;
        tc_fence_after_sync();
        const uint32_t q_lo = k_lo0 + (e & (kSlots - 1)) * (L::kChunkBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kDh / 16; ++kk) {
#ifdef RTTS_X_NOS
          if (kk > 0) continue;
#endif
          umma_ss_lo((t2 & 1) * 256 + kC, q_lo + kk * 2, q_lo + kk * 2, hi, idesc_main, kk > 0);
        }
        umma_commit(bars + 4 * kSlots + (t2 & 1));        // s_full (first arrival)
        umma_commit(bars + 2 * kSlots + ((e - 1) & (kSlots - 1)));
        umma_commit(bars + 2 * kSlots + (e & (kSlots - 1)));
        if (ts.last_in_row() || t2 + 1 == my_tiles) umma_commit(bars + 2 * kSlots + ((e + 1) & (kSlots - 1)));
        ts.next();
      };
      issue_s_main(0);
      if (my_tiles > 1) issue_s_main(1);
      for (int t = 0; t < my_tiles; ++t) {
        const int g = t & 1, e = ti.e;
        P_STAMP(0, t, 4);
        for (int x = v_landed + 1; x <= e + 1; ++x) mbar_wait_a(a_vfull + (x & (kSlots - 1)) * 8, (x / kSlots) & 1);
        v_landed = e + 1;
        mbar_wait_a(a_pfull + g * 8, (t >> 1) & 1);      // (the group's epilogue of tile t-2 precedes this in its program order: O is drained)
        tc_fence_after_sync();
        P_STAMP(0, t, 5);
        const uint32_t t_reg = g * 256;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const uint32_t v_lo = v_lo0 + ((e - 1 + c) & (kSlots - 1)) * (L::kChunkBytes >> 4);
#pragma unroll
          for (int s = 0; s < 4; ++s) {
#ifdef RTTS_X_NOPV
            if (c + s > 0) continue;
#endif
            const int j = c * 4 + s;       // 16 keys: P columns 32 * (j / 2) + 8 * (j % 2)
            umma_ts_lo(t_reg + kColO, t_reg + 32 * (j >> 1) + 8 * (j & 1), v_lo + s * (2048 >> 4), hi, idesc_o, j > 0);
          }
        }
#ifdef RTTS_X_NOSUM
        umma_ts_lo(t_reg + kColSum, t_reg, ones_lo, hi_ones, idesc_sum, 0);
#else
#pragma unroll
        for (int j = 0; j < 12; ++j) umma_ts_lo(t_reg + kColSum, t_reg + 32 * (j >> 1) + 8 * (j & 1), ones_lo, hi_ones, idesc_sum, j > 0);
#endif
        umma_commit(bars + 4 * kSlots + 4 + g);           // o_full
        // the V rows of entries e-1 and e are not read again; e+1 is the look-back chunk of the next tile unless this tile closes
        // the row or the run
        umma_commit(bars + 3 * kSlots + ((e - 1) & (kSlots - 1)));
        umma_commit(bars + 3 * kSlots + (e & (kSlots - 1)));
        if (ti.last_in_row() || t + 1 == my_tiles) umma_commit(bars + 3 * kSlots + ((e + 1) & (kSlots - 1)));
        P_STAMP(0, t, 6);
        // the S / P columns 64..191 of the region are free once PV(t) has executed - which the pipe does before anything this thread
        // issues afterwards
        if (t + 2 < my_tiles) issue_s_main(t + 2);
        ti.next();
      }
    }
  } else if (warp >= kFirstLoaderWarp && warp < kSWarp) {
    // ================================================= loaders =====================================================
    // Two K loaders and two V loaders (loader warp 2i: K rows + metadata, 2i+1: V rows; each takes the entries e = i mod 2).  The K
    // rows of an entry are released by the S that last reads them - before the softmax of that tile has even started - and the V
    // rows by the PV, so the K side runs several tiles ahead and an S never waits for a slot that a PV still holds (with one ring
    // for both, S(t+3) could not be issued before PV(t) had completed plus a full gather latency: the S issuer sat on the loads).
    // lane = (row within a group of 4, 16-byte piece), 8 passes of two 128-byte rows; lane l also owns the metadata of rows l and
    // l + 32.  Software-pipelined: the stickers (and |x|^2 / mask values) of the warp's NEXT entry are requested before the current
    // one is copied; the entry is announced as soon as its rows have landed.
    const int lw = warp - kFirstLoaderWarp;
    const bool is_v = lw >= kKLoaders;
    const int l_first = is_v ? lw - kKLoaders : lw, l_stride = is_v ? kVLoaders : kKLoaders;
    const int grp = lane >> 3, c = lane & 7;
    const uint32_t ld_bytes = static_cast<uint32_t>(p.ld) * 2u;      // row pitch: T * ld < 2^31 elements (checked by the host)
    const int src_e = grp;                                            // lane holding the position of row 8 * i2 + grp is (8 * i2 + grp) & 31
    const float ssl2 = p.score_scale_log2;
    const uint32_t a_full = is_v ? a_vfull : a_kfull, a_free = is_v ? a_vfree : a_kfree;
    EntryIter it(2 * tg0, 2 * tg1, cpr, cpround, p.R, p.H);
    auto fetch = [&]() {                              // this warp's next entry
      Entry x = it.next();
      for (int i = 1; i < l_stride; ++i) it.next();
      return x;
    };
    // raw stickers (round * T + position) of rows lane, lane + 32: requested TWO entries ahead; (K side) their |x|^2 / mask values:
    // requested one entry ahead, when the stickers they depend on have arrived - no load is waited for on the per-entry path
    auto load_stickers = [&](const Entry& x, int& s0, int& s1) {
      if (!x.ok) return;
      const int32_t* stk = p.sticker + static_cast<int64_t>(x.row) * RT + x.j * kC;
      s0 = __ldg(stk + lane);
      s1 = __ldg(stk + 32 + lane);
    };
    auto load_meta = [&](const Entry& x, int s0, int s1, float& q0, float& q1, uint32_t& v0, uint32_t& v1) {
      q0 = q1 = 1.f;
      v0 = v1 = 1u;
      if (!x.ok || is_v) return;
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
    for (int i = 0; i < l_first; ++i) it.next();
    Entry e0 = fetch(), e1 = fetch(), e2;
    int s0 = 0, s1 = 0, n0 = 0, n1 = 0, m0 = 0, m1 = 0;
    float q0, q1, nq0, nq1;
    uint32_t v0, v1, nv0, nv1;
    load_stickers(e0, s0, s1);
    load_stickers(e1, n0, n1);
    load_meta(e0, s0, s1, q0, q1, v0, v1);
    auto announce = [&](int slot) {
      fence_proxy_async_smem();            // cp.async / st.shared data -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_full + slot * 8);
    };
    for (int e = l_first; e0.ok; e += l_stride) {
      const int slot = e & (kSlots - 1), ms = e & (kMetaSlots - 1);
      if (lane == 0) P_STAMP(1, e, is_v ? 4 : 0);
      e2 = fetch();
      load_stickers(e2, m0, m1);
      load_meta(e1, n0, n1, nq0, nq1, nv0, nv1);
      const int base_round = e0.round * p.T;
      // the last MMA reading the entry that lived in the slot has completed
      if (e >= kSlots) {
        int ready = 1;
        if (lane == 0) ready = mbar_try_wait_a(a_free + slot * 8, ((e / kSlots) - 1) & 1);
        ready = __shfl_sync(0xffffffffu, ready, 0);
        if (!ready) {
          warp_wait(a_free + slot * 8, ((e / kSlots) - 1) & 1);
        }
      }
      if (lane == 0) P_STAMP(1, e, is_v ? 5 : 1);
      const int p0 = s0 - base_round, p1 = s1 - base_round;
      // rows 4i + grp: the swizzle term (row & 7) alternates between grp and grp + 4 with the parity of i
      const uint32_t ring = sbase + (is_v ? L::kOffV : L::kOffK) + slot * L::kChunkBytes;
      const uint32_t d_e = ring + sw128_offset(grp, c), d_o = ring + sw128_offset(grp + 4, c);
      // row addresses = base of the (batch, head, 16-byte piece) + position * row pitch: one IMAD.WIDE each (the generic form - 64-bit
      // pointer arithmetic on element offsets, lane selects, a rolled loop - was 590 of the loader's 920 instructions per entry)
      const uint64_t src_b = static_cast<uint64_t>(__cvta_generic_to_global((is_v ? p.v : p.qk) + static_cast<int64_t>(e0.b) * p.T * p.ld + e0.h * kDh + c * 8));
      const bool mirror = !is_v && slot == 0 && e > 0;      // slot kSlots mirrors slot 0 (K rows): the 128 rows of (slot 7, slot 0)
#pragma unroll
      for (int i2 = 0; i2 < kC / 8; ++i2) {
        const int psel = i2 < 4 ? p0 : p1;
        const uint32_t pre = static_cast<uint32_t>(__shfl_sync(0xffffffffu, psel, src_e + 8 * (i2 & 3)));
        const uint32_t pro = static_cast<uint32_t>(__shfl_sync(0xffffffffu, psel, src_e + 8 * (i2 & 3) + 4));
        const uint64_t ge = mad_wide(pre, ld_bytes, src_b), go = mad_wide(pro, ld_bytes, src_b);
#ifdef RTTS_X_NOCOPY
        if (pre == 0xffffffffu)
#endif
        {
          cp_async16_g(d_e + i2 * 1024, ge);
          cp_async16_g(d_o + i2 * 1024, go);
        }
        if (mirror) {
          cp_async16_g(d_e + i2 * 1024 + kSlots * L::kChunkBytes, ge);
          cp_async16_g(d_o + i2 * 1024 + kSlots * L::kChunkBytes, go);
        }
      }
      cp_async_commit();
      if (lane == 0) P_STAMP(1, e, is_v ? 6 : 2);
      if (!is_v) {
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
          const int flags = (e0.jr == 0 ? kFlagRoundStart : 0) | (any_big ? kFlagExact : 0);
          sts128(a_meta + L::kMInfo, make_uint4(static_cast<uint32_t>(e0.row), static_cast<uint32_t>(base_round), static_cast<uint32_t>(flags), 0u));
        }
      }
      // announce as soon as the rows have landed: a loader handles one entry per one or two tiles, so waiting ~700 cycles here costs
      // nothing, while an announcement deferred to the warp's next entry reached the PV thread a whole tile later
      cp_async_wait<0>();
      announce(slot);
      if (lane == 0) P_STAMP(1, e, is_v ? 7 : 3);
      e0 = e1; e1 = e2;
      s0 = n0; s1 = n1; n0 = m0; n1 = m1; q0 = nq0; q1 = nq1; v0 = nv0; v1 = nv1;
    }
  } else {
    // ================================================= softmax + epilogue ==========================================
    // Thread = (query row = TMEM lane, half of its 128-key window, half of its output row), for the whole life of the row:
    // scores -> P -> O / rowsum -> store.  Two warps per lane quarter and group (four compute warps per SM sub-partition): the
    // per-warp instruction stream is a chain of dependent TMEM / MUFU / shared-memory latencies, and two warps per sub-partition
    // did not cover them (softmax phase of a tile 3000-3700 cycles against 1024 cycles of MUFU time).
    const int q = warp & 3, hf = (warp >> 2) & 1, g = warp >> 3;
    const int is_u = q >> 1;                      // rows of chunk e+1 (lanes 64-127): window = S columns 64..191
    const int rr = (q & 1) * 32 + lane;           // row inside the query's chunk
    const uint32_t t_lane = (static_cast<uint32_t>(q * 32) << 16) + g * 256;
    const uint32_t t_s = t_lane + is_u * kC + hf * kC;      // this thread's 64 score columns: keys of the look-back chunk (hf = 0) / own chunk (hf = 1)
    const float mv = p.mask_value_log2, sv = p.self_value_log2;
    const uint32_t a_stage = sbase + L::kOffStage + warp * L::kStageBytes;      // 32 rows x 64 B
    const uint32_t a_xch = sbase + L::kOffXch + ((g * 2 + hf) * 128 + q * 32 + lane) * 4;      // exact mode: row maximum of this half
    const uint32_t l3 = lane & 3;
    TileIter ti(tg0, p.tiles_per_row);
    if (g) ti.next();
    for (int t = g; t < my_tiles; t += 2) {
      const uint32_t ph = (t >> 1) & 1;
      const int eq = ti.e + is_u;                 // entry of this row's chunk; its look-back chunk is entry eq - 1
      ti.next(); ti.next();
      const uint32_t a_meta_q = sbase + L::kOffMeta + (eq & (kMetaSlots - 1)) * L::kMetaBytes;
      const uint32_t a_meta_lb = sbase + L::kOffMeta + ((eq - 1) & (kMetaSlots - 1)) * L::kMetaBytes;
      const uint32_t a_km = hf ? a_meta_q : a_meta_lb;       // metadata of this thread's 64 keys
      // S(t) was issued after its threads had seen the entries land, so its completion also certifies the loaders' metadata
      if (warp == 8 * g && lane == 0) P_STAMP(2, t, 0);
      warp_wait(a_sfull + g * 8, ph);
      tc_fence_after_sync();
      if (warp == 8 * g && lane == 0) P_STAMP(2, t, 1);
      const uint4 info = lds128(a_meta_q + L::kMInfo);       // {row_bh, round * T, flags, -}
      const uint2 qm = lds64(a_meta_q + L::kMQ + rr * 8);
      float row_max = -__uint_as_float(qm.x);
      if (warp == 8 * g && lane == 0) P_STAMP(3, t, 0);
      if (info.z & kFlagExact) {
        const int q_enc = static_cast<int>(lds32(a_meta_q + L::kMPos + rr * 4));
        int q_limit = p.causal ? (q_enc & ~kPadFlag) : (kPadFlag - 1);
        if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && (q_enc & kPadFlag)) q_limit = -1;
        float mx = -FLT_MAX;
#pragma unroll 1
        for (int s4 = 0; s4 < 4; ++s4) {
          uint32_t r[16];
          tmem_ld16(t_s + 16 * s4, r);
          tmem_ld_wait();
          mx = exact16_max(r, a_km + L::kMScale + s4 * 64, a_km + L::kMPos + s4 * 64, q_limit, q_enc, mv, sv, mx);
        }
        // the two halves of a row exchange their maxima (the flag is per chunk: both warps of the pair take this path)
        sts32(a_xch, __float_as_uint(mx));
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g * 4 + q) : "memory");
        mx = fmaxf(mx, __uint_as_float(lds32(a_xch + (hf ? -512 : 512))));
        asm volatile("bar.sync %0, 64;" ::"r"(1 + g * 4 + q) : "memory");      // (the slot is rewritten two tiles later at the earliest; cheap insurance)
#pragma unroll 1
        for (int s4 = 0; s4 < 4; ++s4) {
          uint32_t r[16], pk[8];
          tmem_ld16(t_s + 16 * s4, r);
          tmem_ld_wait();
          exact16(r, a_km + L::kMScale + s4 * 64, a_km + L::kMPos + s4 * 64, -mx, q_limit, q_enc, mv, sv, pk);
          tmem_st8(t_s + 32 * (s4 >> 1) + 8 * (s4 & 1), pk);
        }
        row_max = mx;
      } else {
        // one short loop for every warp (instruction fetch is a first-order resource for warp-specialised kernels here: a 6 KB L0
        // / 32 KB L1.5 instruction cache)
        const float neg_m = __uint_as_float(qm.x);
        const uint32_t a_scale = a_km + L::kMScale, a_p16 = a_km + L::kMPos16;
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
          uint32_t ra[32], pk[8];
          tmem_ld32(t_s + 32 * hb, ra);
          tmem_ld_wait();
          if (warp == 8 * g && lane == 0) P_STAMP(3, t, 1 + 2 * hb);
          if (p.causal) {
            soft16_packed<true>(ra, a_scale + hb * 128, a_p16 + hb * 64, neg_m, qm.y, pk);
            tmem_st8(t_s + 32 * hb, pk);
            soft16_packed<true>(ra + 16, a_scale + hb * 128 + 64, a_p16 + hb * 64 + 32, neg_m, qm.y, pk);
          } else {
            soft16_packed<false>(ra, a_scale + hb * 128, a_p16 + hb * 64, neg_m, qm.y, pk);
            tmem_st8(t_s + 32 * hb, pk);
            soft16_packed<false>(ra + 16, a_scale + hb * 128 + 64, a_p16 + hb * 64 + 32, neg_m, qm.y, pk);
          }
          tmem_st8(t_s + 32 * hb + 8, pk);
          if (warp == 8 * g && lane == 0) P_STAMP(3, t, 2 + 2 * hb);
        }
      }
      {
        // the 64 keys outside this row's window contribute nothing: zero their two P blocks (S columns 128..191 for the rows of
        // chunk e, 0..63 for chunk e+1), one per half
        const uint32_t z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        tmem_st16(t_lane + (is_u ? 0 : 2 * kC) + 32 * hf, z);
      }
      tmem_st_wait();
      if (warp == 8 * g && lane == 0) P_STAMP(3, t, 5);
      tc_fence_before_sync();       // this thread's TMEM reads of S / writes of P precede the MMAs that consume / overwrite the region
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_pfull + g * 8);
      if (lane == 0 && (warp == 8 * g || warp == 8 * g + 7)) P_STAMP(2, t, 2 + (warp != 8 * g));

      // ---------------------------------------------- epilogue of the same rows ----------------------------------
      warp_wait(a_ofull + g * 8, ph);
      tc_fence_after_sync();
      if (warp == 8 * g && lane == 0) P_STAMP(2, t, 4);
      uint32_t rs, o0[32];
      tmem_ld1(t_lane + kColSum, &rs);
      tmem_ld32(t_lane + kColO + 32 * hf, o0);
      const int pos_enc = static_cast<int>(lds32(a_meta_q + L::kMPos + rr * 4));
      tmem_ld_wait();
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(a_ofree + g * 8);      // the row sums have been read: the look-back block of S(t+2) may overwrite them
      if (warp == 8 * g && lane == 0) P_STAMP(2, t, 5);
#ifdef RTTS_X_NOEPI
      if (rs != 0x12345u) continue;
#endif
      float sum = __uint_as_float(rs);
      // all terms exactly zero: the row sees only itself (rp R8): softmax uniform over the self columns, which all hold the
      // query's own token, so out = v[own position], lse = self_value + log(#self columns)
      const bool lonely = !(sum > 0.f);
      if (lonely) {
        // a second self column exists if the own token also sits in the look-back chunk (first chunk of a hash round only)
        int dup = 0;
        if (info.z & kFlagRoundStart) {
          for (int j = 0; j < kC; ++j) dup |= static_cast<int>(lds32(a_meta_lb + L::kMPos + j * 4)) == pos_enc;
        }
        sum = dup ? 2.f : 1.f;
        row_max = sv;
      }
      const float inv = 1.f / sum;
      // half an O row / row sum -> bf16 -> this warp's staging tile (row = lane, 64 B; 16-byte pieces swizzled by the row so that both
      // the row-wise writes here and the piece-wise reads of the store phase are conflict-free)
      {
        float* o = reinterpret_cast<float*>(o0);
#pragma unroll
        for (int k = 0; k < 32; k += 2) fmul2(o[k], o[k + 1], o[k], o[k + 1], inv, inv);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sts128(a_stage + lane * 64 + ((i ^ ((lane >> 1) & 3)) << 4),
                 make_uint4(pack_bf16(o[8 * i], o[8 * i + 1]), pack_bf16(o[8 * i + 2], o[8 * i + 3]), pack_bf16(o[8 * i + 4], o[8 * i + 5]), pack_bf16(o[8 * i + 6], o[8 * i + 7])));
      }
      const int pos = pos_enc & ~kPadFlag;
      const int row_bh = static_cast<int>(info.x);
      if (__any_sync(0xffffffffu, lonely)) {
        if (lonely) {
          const int b = row_bh / p.H, h = row_bh - b * p.H;
          const uint4* vrow = reinterpret_cast<const uint4*>(p.v + (static_cast<int64_t>(b) * p.T + pos) * p.ld + h * kDh) + 4 * hf;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) sts128(a_stage + lane * 64 + ((ch ^ ((lane >> 1) & 3)) << 4), __ldg(vrow + ch));
        }
      }
      __syncwarp();
      const uint32_t out_slot = static_cast<uint32_t>(static_cast<int>(info.y) + pos);      // unsorted slot inside the (batch, head) row = round * T + position
      const int64_t row_base = static_cast<int64_t>(row_bh) * RT;
      if (hf == 0) p.lse_rounds[row_base + out_slot] = (row_max + log2f(sum)) * kLn2;
      // scatter-store, this warp's half (64 B) of a row per 4 lanes (eight rows per instruction)
      {
        const uint64_t obase = static_cast<uint64_t>(__cvta_generic_to_global(p.o_rounds)) + static_cast<uint64_t>(row_base) * (kDh * 2) + hf * 64 + l3 * 16;
        const uint32_t r8 = lane >> 2;
        const uint32_t a_r = a_stage + r8 * 64 + ((l3 ^ ((r8 >> 1) & 3)) << 4);      // rows 8 * itr + r8: the swizzle term ((row >> 1) & 3) does not depend on itr
#pragma unroll
        for (int itr = 0; itr < 4; ++itr) {
          const uint4 u = lds128(a_r + itr * 512);
          const uint32_t os = __shfl_sync(0xffffffffu, out_slot, itr * 8 + r8);
#ifndef RTTS_X_NOSTORE
          stg128(mad_wide(os, kDh * 2, obase), u);
#else
          if (os == 0xffffffffu) stg128(mad_wide(os, kDh * 2, obase), u);
#endif
        }
      }
      __syncwarp();                 // the staging tile is free again
      if (warp == 8 * g && lane == 0) P_STAMP(2, t, 6);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == kSWarp) tmem_dealloc(0, 512);
}

}  // namespace f64p

int launch_attn_fwd64p(const AttnFwdParams& p, int B, cudaStream_t stream) {
  using L = f64p::Smem;
  static bool configured = false;   // idempotent attribute set; benign if raced
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(f64p::lsh_attn_fwd64p_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_attn_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int64_t tiles = static_cast<int64_t>(B) * p.H * p.tiles_per_row;
  if (tiles <= 0 || tiles >= (1ll << 30)) return fail(kErrBadArg, "rtts_lsh_attn_fwd: bad grid");
  const int grid = tiles < kNumSMs ? static_cast<int>(tiles) : kNumSMs;     // persistent: one CTA per SM
  f64p::lsh_attn_fwd64p_kernel<<<grid, f64p::kThreads, L::kTotal, stream>>>(p, static_cast<int>(tiles));
  return check_launch("rtts_lsh_attn_fwd");
}

}  // namespace rtts
