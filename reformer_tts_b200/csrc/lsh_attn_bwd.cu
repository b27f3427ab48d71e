// Fused chunked shared-QK attention, backward (rtts_lsh_attn_bwd) and the per-token gradient reduce
// (rtts_lsh_grad_reduce).  Contract: include/rtts_b200.h; tiling: DESIGN.md "K15".
//
// Across hash rounds the LSH attention is ONE softmax over the multiset of (round, key) pairs with the
// total normaliser L = logsumexp_r lse_r, so with Pt_ij = exp(s_ij - L_i) and delta_i = <dout_i, out_i>
//     dV_j += Pt_ij dout_i,   dS_ij = Pt_ij (<dout_i, v_j> - delta_i)  (0 where a mask constant was written),
//     dQ_i += dS_ij khat_j,   dKhat_j += dS_ij q_i
// and nothing per-round from the forward has to be stored: the scores are recomputed here.
//
// One CTA (128 threads, thread = key row) owns 128 consecutive SORTED key slots and the 128+bucket query
// slots that can see them (the key slots themselves, then the next chunk, which looks one chunk back).
// Everything is formed TRANSPOSED, keys on the TMEM lane axis, because per-query quantities (L, delta,
// position) are then per-column broadcasts and per-key ones (1/|k|, position) live in registers:
//   per 64-query block qb:  St  = K  Q_qb^T      (M128 N64 K64)      dPt = V dO_qb^T   (M128 N64 K64)
//                           Pt, dSt -> bf16 shared-memory tiles (K-major in q), then
//                           dV += Pt_qb dO_qb    (M128 N64 K64)      G  += dSt_qb Q_qb (M128 N64 K64)
//   per pair of blocks:     dQ  = dS K           (A = dSt tile read MN-major, M128 N64 K128)
// G = (dKhat * 1/|k|) goes through the key-normalisation Jacobian in the epilogue: dx = G - x |k|^-2 <x, G>.
// A key's dV / dx are complete inside one CTA; a query chunk's dQ is split between the CTA that owns it as
// a key chunk (dq_a) and the previous one where it is the look-back chunk (dq_b).
#include <cfloat>
#include <type_traits>

#include "common.cuh"
#include "host_util.h"
#include "rtts_b200.h"

namespace rtts {

constexpr int kBDh = 64;
constexpr int kKeyRows = 128;
constexpr float kBLog2e = 1.4426950408889634f;
constexpr int kBPadFlag = 0x40000000;
constexpr int kBlk = 128 * 128;   // bytes of one 64-column block of a 128-row bf16 tile

struct AttnBwdParams {
  const __nv_bfloat16* qk;
  const __nv_bfloat16* v;
  const __nv_bfloat16* dout;
  int64_t ld;      // token stride of qk / v
  int64_t ld_do;   // token stride of dout
  const int32_t* sticker;
  const float* sumsq;   // [B,H,T] |qk row|^2
  const uint8_t* mask;
  const float* lse;     // [B,H,T]
  const float* delta;   // [B,H,T]
  __nv_bfloat16* dqk_main;   // per key slot: query-role gradient from this CTA's keys + key-role gradient (Jacobian applied)
  __nv_bfloat16* dq_b;       // per look-ahead slot: query-role gradient from this CTA's keys
  __nv_bfloat16* dv;
  long long* trace;   // debug: clock64 stamps of CTA 0 (nullable)
  int T, H, R, tiles_per_row;
  float score_scale;
  float mask_value_log2, self_value_log2;
  int key_norm, mask_mode, causal;
};

#define RTTS_BSTAMP(k) do { if (p.trace != nullptr && blockIdx.x == 0 && tid == 0 && n == 1) p.trace[(k)] = clock64(); } while (0)
// Persistent: one CTA per SM walks tiles blockIdx.x, += gridDim.x.
//   warps 0-15 : workers - 4 warpgroups; warpgroup g owns 16 of the 64 query columns of every block / 16 of the 64 output columns
//   warp 16    : one elected lane issues every tcgen05.mma
//   warps 17-18: loaders - gather the next tile's qk / v / dout rows and per-row statistics into the other input stage while
//                the current tile is being processed (bucket 128 has room for one stage only: no overlap, but no CTA relaunch either)
constexpr int kBwdWorkers = 512;
constexpr int kBwdMmaWarp = kBwdWorkers / 32;
constexpr int kBwdLoaderThreads = 64;
constexpr int kBwdThreads = kBwdWorkers + 32 + kBwdLoaderThreads;

template <int BUCKET>
struct AttnBwdSmem {
  static constexpr int kQRows = kKeyRows + BUCKET;      // 192 | 256
  static constexpr int kQBlocks = kQRows / 64;          // 3 | 4
  static constexpr int kStages = BUCKET == 64 ? 2 : 1;
  // per input stage (tile bases 1024-B aligned):
  static constexpr int kOffX = 0;                        // qk rows of all query slots (first 128 = keys)
  static constexpr int kOffV = kOffX + kQRows * 128;     // v rows of the key slots
  static constexpr int kOffDO = kOffV + kKeyRows * 128;  // dout rows of the query slots
  static constexpr int kOffQEnc = kOffDO + kQRows * 128;   // int[kQRows] position | pad flag (self test)
  static constexpr int kOffQLim = kOffQEnc + kQRows * 4;   // int[kQRows] largest key position this query may see (mask test)
  static constexpr int kOffQStat = kOffQLim + kQRows * 4;  // float2[kQRows] (L*log2e, delta)
  static constexpr int kOffQSlot = kOffQStat + kQRows * 8; // int[kQRows] unsorted slot
  static constexpr int kOffKInv = kOffQSlot + kQRows * 4;  // float[kKeyRows] 1/|k|
  static constexpr int kStageBytes = ((kOffKInv + kKeyRows * 4 + 1023) / 1024) * 1024;
  // shared by all tiles:
  static constexpr int kOffDS = kStages * kStageBytes;     // dSt tile: always 4 blocks of [128 x 64] bf16 (pair reads need block 3); staging in the epilogue
  static constexpr int kOffDot = kOffDS + 4 * kBlk;        // float[4][kKeyRows] partial <x, G>
  static constexpr int kOffBar = kOffDot + 4 * kKeyRows * 4;   // in_full[2], in_free[2], S/dP buffers [2], block written [2], accumulators [1]
  static constexpr int kOffTmem = kOffBar + 9 * 8;
  static constexpr int kTotal = kOffTmem + 8;
  static constexpr int kDynamic = kTotal + 1024;
};

  // Gather one tile's qk / dout rows of the query slots, v rows of the key slots and the per-row statistics into an input stage.
  // Executed by `nthr` threads (lt = 0..nthr-1): groups of 8 lanes copy a row (lane c = 16-byte chunk c); lane c also owns the
  // statistics of pass c of every 8 passes.  pos = sticker - round*T with the round taken from the slot index (see the forward kernel).
template <int BUCKET, int NTHR>
__device__ __forceinline__ void bwd_gather_tile(const AttnBwdParams& p, uint8_t* smem, uint64_t* in_full, int tile_id, int st, int lt) {
  using L = AttnBwdSmem<BUCKET>;
  constexpr int kQRows = L::kQRows;
  constexpr int G = NTHR >> 3;                       // rows per pass
  constexpr int kPasses = (kQRows + G - 1) / G;      // 24 | 32 passes of 8 rows (loader warps), 3 | 4 of 64 rows (workers)
  constexpr int kStatRounds = (kPasses + 7) / 8;     // lane c owns the statistics of passes c, c + 8, ...
  const int RT = p.R * p.T;
  const int g = lt >> 3, c = lt & 7;
  uint8_t* stage = smem + st * L::kStageBytes;
  const uint32_t sX = smem_u32(stage + L::kOffX), sV = smem_u32(stage + L::kOffV), sDO = smem_u32(stage + L::kOffDO);
  int* q_enc = reinterpret_cast<int*>(stage + L::kOffQEnc);
  int* q_lim = reinterpret_cast<int*>(stage + L::kOffQLim);
  float2* q_stat = reinterpret_cast<float2*>(stage + L::kOffQStat);
  int* q_slot = reinterpret_cast<int*>(stage + L::kOffQSlot);
  float* k_inv = reinterpret_cast<float*>(stage + L::kOffKInv);
  const int row_bh = tile_id / p.tiles_per_row, tile = tile_id - row_bh * p.tiles_per_row;
  const int b = row_bh / p.H, h = row_bh - b * p.H;
  const int32_t* stk = p.sticker + static_cast<int64_t>(row_bh) * RT;
  const int first_slot = tile * kKeyRows;      // sorted slot of row 0; rows >= RT wrap to the start
  const int ahead_first = first_slot + kKeyRows >= RT ? first_slot + kKeyRows - RT : first_slot + kKeyRows;   // look-ahead chunk
  const int base_main = (first_slot / p.T) * p.T, base_ahead = (ahead_first / p.T) * p.T;
  const int64_t head_off = static_cast<int64_t>(h) * kBDh + c * 8;
  const uint32_t ld32 = static_cast<uint32_t>(p.ld), ld_do32 = static_cast<uint32_t>(p.ld_do);
  const __nv_bfloat16* qk_b = p.qk + static_cast<int64_t>(b) * p.T * p.ld + head_off;
  const __nv_bfloat16* v_b = p.v + static_cast<int64_t>(b) * p.T * p.ld + head_off;
  const __nv_bfloat16* do_b = p.dout + static_cast<int64_t>(b) * p.T * p.ld_do + head_off;
  // all sticker loads first (one exposed latency per tile), then every row copy, then the statistics while the copies fly
  int stv[kPasses];
#pragma unroll
  for (int i = 0; i < kPasses; ++i) {
    const int r = i * G + g;
    stv[i] = r < kQRows ? __ldg(stk + (r < kKeyRows ? first_slot + r : ahead_first + (r - kKeyRows))) : 0;
  }
#pragma unroll
  for (int i = 0; i < kPasses; ++i) {
    const int r = i * G + g;
    if (r < kQRows) {
      const int pos = stv[i] - (r < kKeyRows ? base_main : base_ahead);
      const uint32_t so = sw128_offset(r, c);
      const uint32_t off = static_cast<uint32_t>(pos) * ld32, off_do = static_cast<uint32_t>(pos) * ld_do32;      // T * ld < 2^31 (checked by the host)
      cp_async16(sX + so, qk_b + off);
      cp_async16(sDO + so, do_b + off_do);
      if (r < kKeyRows) cp_async16(sV + so, v_b + off);
    }
  }
  cp_async_commit();
#pragma unroll
  for (int q = 0; q < kStatRounds; ++q) {
    int my_st = 0;
#pragma unroll
    for (int i = 0; i < kPasses; ++i) my_st = (i == q * 8 + c) ? stv[i] : my_st;
    const int r = (q * 8 + c) * G + g;
    if (q * 8 + c < kPasses && r < kQRows) {
      const int my_pos = my_st - (r < kKeyRows ? base_main : base_ahead);
      const int64_t sidx = static_cast<int64_t>(row_bh) * p.T + my_pos;
      const float lse_v = __ldg(p.lse + sidx), delta_v = __ldg(p.delta + sidx);
      const bool valid = p.mask == nullptr || __ldg(p.mask + static_cast<int64_t>(b) * p.T + my_pos) != 0;
      const int enc = valid ? my_pos : (my_pos | kBPadFlag);
      int limit = p.causal ? my_pos : (kBPadFlag - 1);
      if (p.mask_mode == RTTS_MASK_QUERY_AND_KEY && !valid) limit = -1;
      q_enc[r] = enc;
      q_lim[r] = limit;
      q_stat[r] = make_float2(lse_v * kBLog2e, delta_v);
      q_slot[r] = my_st;
      if (r < kKeyRows) {
        const float ss = __ldg(p.sumsq + sidx);
        k_inv[r] = p.key_norm == RTTS_KEYNORM_L2 ? 1.f / fmaxf(sqrtf(ss), 1e-12f) : rsqrtf(ss * (1.f / kBDh) + 1e-6f) * 0.125f;
      }
    }
  }
  cp_async_wait<0>();
  fence_proxy_async_smem();
  __syncwarp();
  if ((lt & 31) == 0) mbar_arrive(in_full + st);
}

template <int BUCKET>
__global__ void __launch_bounds__(kBwdThreads, 1) lsh_attn_bwd_kernel(const AttnBwdParams p, const int num_tiles) {
  using L = AttnBwdSmem<BUCKET>;
  constexpr int kQRows = L::kQRows, kQBlocks = L::kQBlocks, kStages = L::kStages;
  constexpr uint32_t kTmemCols = 512;
  // TMEM columns: two (St, dPt) buffers so the tensor core works on block qb+1 while block qb is consumed.  Pt (bf16 pairs) is
  // written back over the St columns its thread has consumed and read from there as the A operand of dV += Pt dO.
  constexpr uint32_t cS0 = 0, cDP0 = 64, cS1 = 128, cDP1 = 192, cDV = 256, cG = 320, cDQ0 = 384, cDQ1 = 448;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sDS = smem_u32(smem + L::kOffDS);
  float* dot_part = reinterpret_cast<float*>(smem + L::kOffDot);
  uint64_t* in_full = reinterpret_cast<uint64_t*>(smem + L::kOffBar);   // [2] loader -> everyone
  uint64_t* in_free = in_full + 2;                                      // [2] 16 worker-warp arrivals: stage may be refilled
  uint64_t* bar_s = in_full + 4;                                        // [2] tcgen05.commit: (St, dPt) of a block ready
  uint64_t* bar_blk = in_full + 6;                                      // [2] 16 worker-warp arrivals: Pt / dSt of a block written
  uint64_t* bar_acc = in_full + 8;                                      // tcgen05.commit: accumulators of the tile complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L::kOffTmem);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int RT = p.R * p.T;

  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(in_full + i, (kStages == 2 ? kBwdLoaderThreads : kBwdWorkers) / 32);
      mbar_init(in_free + i, kBwdWorkers / 32);
      mbar_init(bar_s + i, 1);
      mbar_init(bar_blk + i, kBwdWorkers / 32);        // one arrival per worker warp
    }
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  if (*tmem_slot != 0) __trap();      // all 512 columns are ours: the allocation starts at lane 0, column 0
  constexpr uint32_t tmem = 0;

  constexpr bool kPrefetch = kStages == 2;       // one stage only (bucket 128): nothing to overlap, so the workers gather themselves

  if (warp > kBwdMmaWarp) {
    // ================================================= loaders ====================================================
    if (kPrefetch) {
      const int lt = tid - (kBwdMmaWarp + 1) * 32;   // 0..63
      int n = 0;
      for (int tile_id = blockIdx.x; tile_id < num_tiles; tile_id += gridDim.x, ++n) {
        const int st = n % kStages;
        mbar_wait(in_free + st, ((n / kStages) & 1) ^ 1);      // the tile that used this stage is completely done
        bwd_gather_tile<BUCKET, kBwdLoaderThreads>(p, smem, in_full, tile_id, st, lt);
      }
    }
  } else if (warp == kBwdMmaWarp) {
    // ================================================= MMA issuer =================================================
    if (elect_one()) {      // elect.sync: the compiler then emits the tcgen05.mma sequences straight-line (no per-instruction uniformity loop)
      constexpr uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
      constexpr uint32_t idesc_kn = umma_idesc_bf16(128, 64, false, true);   // A K-major (or TMEM), B MN-major
      constexpr uint32_t idesc_nn = umma_idesc_bf16(128, 64, true, true);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const uint32_t dDS_k = umma_desc_lo(sDS, 16), dDS_n = umma_desc_lo(sDS, kBlk);
      uint32_t n_blk[2] = {0, 0};      // completed uses of bar_blk[i]
      int n = 0;
      for (int tile_id = blockIdx.x; tile_id < num_tiles; tile_id += gridDim.x, ++n) {
        const int st = n % kStages;
        const uint32_t sX = smem_u32(smem + st * L::kStageBytes + L::kOffX), sV = smem_u32(smem + st * L::kStageBytes + L::kOffV),
                       sDO = smem_u32(smem + st * L::kStageBytes + L::kOffDO);
        // descriptor bases (K-major: lbo 16; MN-major: lbo 0); a k-step / block offset is an add on the address field
        const uint32_t dX_k = umma_desc_lo(sX, 16), dV_k = umma_desc_lo(sV, 16), dDO_k = umma_desc_lo(sDO, 16);
        const uint32_t dDO_n = umma_desc_lo(sDO, 0), dX_n = umma_desc_lo(sX, 0);
        auto issue_scores = [&](int qb) {      // St / dPt of query block qb into TMEM buffer qb & 1
          const uint32_t cs_ = (qb & 1) ? cS1 : cS0, cdp_ = (qb & 1) ? cDP1 : cDP0;
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_lo(tmem + cs_, dX_k + (k * 32 >> 4), dX_k + ((qb * 8192 + k * 32) >> 4), hi, idesc, k > 0);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_ss_lo(tmem + cdp_, dV_k + (k * 32 >> 4), dDO_k + ((qb * 8192 + k * 32) >> 4), hi, idesc, k > 0);
          umma_commit(bar_s + (qb & 1));
        };
        mbar_wait(in_full + st, (n / kStages) & 1);
        tc_fence_after_sync();
        // the (St, dPt) buffers are free: every block of the previous tile was consumed before its accumulate MMAs were issued
        issue_scores(0);
        issue_scores(1);
#pragma unroll 1
        for (int qb = 0; qb < kQBlocks; ++qb) {
          // Pt / dSt of block qb written, its dPt buffer consumed.  For qb == 0 this also certifies that every worker has finished
          // the previous tile's epilogue (program order), i.e. the accumulators may be overwritten.
          mbar_wait(bar_blk + (qb & 1), n_blk[qb & 1] & 1);
          ++n_blk[qb & 1];
          tc_fence_after_sync();
          const uint32_t cs_ = (qb & 1) ? cS1 : cS0;
#pragma unroll
          for (int k = 0; k < 4; ++k)      // dV += Pt dO: A = Pt from TMEM (16 queries = 8 columns of bf16 pairs per k-step)
            umma_ts_lo(tmem + cDV, tmem + cs_ + 16 * k, dDO_n + ((qb * 8192 + k * 2048) >> 4), hi, idesc_kn, (qb | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_ss_lo(tmem + cG, dDS_k + ((qb * kBlk + k * 32) >> 4), dX_n + ((qb * 8192 + k * 2048) >> 4), hi, idesc_kn, (qb | k) != 0);
          if ((qb & 1) || qb == kQBlocks - 1) {
            // dQ for query rows [pair*128, +128): A = dSt blocks (pair*2, pair*2+1) read MN-major (M = queries)
            const int pair = qb >> 1;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_ss_lo(tmem + (pair ? cDQ1 : cDQ0), dDS_n + ((pair * 2 * kBlk + k * 2048) >> 4), dX_n + (k * 2048 >> 4), hi, idesc_nn, k > 0);
          }
          if (qb + 2 < kQBlocks) issue_scores(qb + 2);     // refill the buffer that was just consumed
        }
        umma_commit(bar_acc);
      }
    }
  } else {
    // ================================================= workers ====================================================
    const int wg = tid >> 7;            // column group 0..3
    const int j = tid & 127;            // key row = TMEM lane
    const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    const float mv = p.mask_value_log2, sv = p.self_value_log2;
    const int col0 = wg * 16;                         // this thread's 16 query columns inside a block
    const int k_chunk = j / BUCKET;                   // 0|1 for bucket 64, 0 for bucket 128
    uint32_t n_s[2] = {0, 0};            // completed uses of bar_s[i]
    int n = 0;
    for (int tile_id = blockIdx.x; tile_id < num_tiles; tile_id += gridDim.x, ++n) {
      const int st = n % kStages;
      uint8_t* stage = smem + st * L::kStageBytes;
      const uint32_t sX = smem_u32(stage + L::kOffX);
      const uint32_t a_enc = smem_u32(stage + L::kOffQEnc), a_lim = smem_u32(stage + L::kOffQLim), a_stat = smem_u32(stage + L::kOffQStat);
      const int* q_slot = reinterpret_cast<const int*>(stage + L::kOffQSlot);
      const float* k_inv = reinterpret_cast<const float*>(stage + L::kOffKInv);
      const int row_bh = tile_id / p.tiles_per_row, tile = tile_id - row_bh * p.tiles_per_row;
      // the look-ahead chunk starts a new hash round (or wraps to round 0): only then can a key's token show up a second time
      const int ahead_first = tile * kKeyRows + kKeyRows >= RT ? 0 : tile * kKeyRows + kKeyRows;
      const bool ahead_crosses = ahead_first % p.T == 0;
      RTTS_BSTAMP(0);
      if (!kPrefetch) bwd_gather_tile<BUCKET, kBwdWorkers>(p, smem, in_full, tile_id, st, tid);      // (the previous tile ended with a barrier of all workers)
      mbar_wait(in_full + st, (n / kStages) & 1);
      RTTS_BSTAMP(3);
      // this thread's key row
      const float inv = k_inv[j];
      const float cs = inv * p.score_scale * kBLog2e;   // score scale in log2 units
      const float gs = inv * p.score_scale;             // folded into dSt so one tile serves dQ and G
      const int k_enc = static_cast<int>(lds32(a_enc + j * 4));

#pragma unroll 1
      for (int qb = 0; qb < kQBlocks; ++qb) {
        mbar_wait(bar_s + (qb & 1), n_s[qb & 1] & 1);
        ++n_s[qb & 1];
        tc_fence_after_sync();
        const uint32_t cs_ = (qb & 1) ? cS1 : cS0, cdp_ = (qb & 1) ? cDP1 : cDP0;
        // which query chunk is this block, and does it see this warp's key chunk?  (warp-uniform: a dead block costs nothing)
        const int q_chunk = (qb * 64) / BUCKET;
        const bool pair_live = (q_chunk == k_chunk) || (q_chunk == k_chunk + 1);
        // Self entries (key slot == query slot: the key's own column, block j / 64) and second occurrences of its token (only in
        // look-ahead blocks that belong to another hash round) need the reference's self_value; everything else is a plain
        // masked softmax term.  The test is kept warp-uniform: the warp's 32 own columns lie in one 32-column half of a block.
        const bool self_block = (qb == (j >> 6) && (((j & 63) >> 5) == (wg >> 1))) || (qb * 64 >= kKeyRows && ahead_crosses);
        uint32_t pk[8];
        uint4 w[2];
        if (pair_live) {
          uint32_t rs[16], rp[16];
          tmem_ld16(t_row + cs_ + col0, rs);
          tmem_ld16(t_row + cdp_ + col0, rp);
          const uint32_t a_col = (qb * 64 + col0) * 4;
          float pe[16], de[16];
          auto body = [&](auto mask_tag, auto self_tag) {
            constexpr bool MASK = decltype(mask_tag)::value, SELF = decltype(self_tag)::value;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const uint4 s0 = lds128(a_stat + a_col * 2 + q4 * 32), s1 = lds128(a_stat + a_col * 2 + q4 * 32 + 16);
              const float Lv[4] = {__uint_as_float(s0.x), __uint_as_float(s0.z), __uint_as_float(s1.x), __uint_as_float(s1.z)};
              const float dv_[4] = {__uint_as_float(s0.y), __uint_as_float(s0.w), __uint_as_float(s1.y), __uint_as_float(s1.w)};
              uint4 lim = make_uint4(0, 0, 0, 0), enc = make_uint4(0, 0, 0, 0);
              if (MASK) lim = lds128(a_lim + a_col + q4 * 16);
              if (SELF) enc = lds128(a_enc + a_col + q4 * 16);
              const int lv[4] = {static_cast<int>(lim.x), static_cast<int>(lim.y), static_cast<int>(lim.z), static_cast<int>(lim.w)};
              const int ev[4] = {static_cast<int>(enc.x), static_cast<int>(enc.y), static_cast<int>(enc.z), static_cast<int>(enc.w)};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = q4 * 4 + e;
                float pr = exp2f(fmaf(__uint_as_float(rs[i]), cs, -Lv[e]));
                if (MASK) pr = k_enc > lv[e] ? 0.f : pr;               // exp2(mask_value - L) == 0
                float ds = pr * ((__uint_as_float(rp[i]) - dv_[e]) * gs);
                if (SELF) {
                  const bool self = k_enc == ev[e];
                  pr = self ? exp2f(sv - Lv[e]) : pr;                    // self_value overrides the mask; non-zero only for lonely rows
                  ds = self ? 0.f : ds;                                  // a constant was written: no gradient
                }
                pe[i] = pr;
                de[i] = ds;
              }
            }
          };
          tmem_ld_wait();
          const bool need_mask = p.causal || p.mask != nullptr;
          if (self_block) body(std::true_type{}, std::true_type{});
          else if (need_mask) body(std::true_type{}, std::false_type{});
          else body(std::false_type{}, std::false_type{});
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(pe[2 * i], pe[2 * i + 1]);
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2) {
            w[q2].x = pack_bf16(de[q2 * 8 + 0], de[q2 * 8 + 1]); w[q2].y = pack_bf16(de[q2 * 8 + 2], de[q2 * 8 + 3]);
            w[q2].z = pack_bf16(de[q2 * 8 + 4], de[q2 * 8 + 5]); w[q2].w = pack_bf16(de[q2 * 8 + 6], de[q2 * 8 + 7]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = 0u;
          w[0] = make_uint4(0, 0, 0, 0);
          w[1] = make_uint4(0, 0, 0, 0);
        }
        tmem_st8(t_row + cs_ + col0, pk);          // Pt over the St columns this thread has consumed
        sts128(sDS + qb * kBlk + sw128_offset(j, wg * 2), w[0]);
        sts128(sDS + qb * kBlk + sw128_offset(j, wg * 2 + 1), w[1]);
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(bar_blk + (qb & 1));
      }
      mbar_wait(bar_acc, n & 1);
      tc_fence_after_sync();

      RTTS_BSTAMP(20);
      // ---- epilogue: warpgroup g handles output columns [16g, 16g+16) of every accumulator row.  The bf16 results go through
      // shared-memory staging tiles (the dSt blocks, dead now) so that the scattered global stores are whole 128-byte rows.
      const int64_t out_base = static_cast<int64_t>(row_bh) * RT;
      const uint32_t stage_dv = sDS, stage_dx = sDS + kBlk, stage_dqb = sDS + 2 * kBlk;
      auto stage16 = [&](uint32_t tile, const float* o) {
#pragma unroll
        for (int q = 0; q < 2; ++q)
          sts128(tile + sw128_offset(j, wg * 2 + q), make_uint4(pack_bf16(o[q * 8], o[q * 8 + 1]), pack_bf16(o[q * 8 + 2], o[q * 8 + 3]),
                                                                pack_bf16(o[q * 8 + 4], o[q * 8 + 5]), pack_bf16(o[q * 8 + 6], o[q * 8 + 7])));
      };
      {
        uint32_t r[16];
        tmem_ld16(t_row + cDV + col0, r);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]);
        stage16(stage_dv, o);
      }
      float g[16], x[16];
      {
        uint32_t r[16];
        tmem_ld16(t_row + cG + col0, r);
        tmem_ld_wait();
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint4 u = lds128(sX + sw128_offset(j, wg * 2 + c));
          const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            x[c * 8 + 2 * e] = bf16_lo(w[e]);
            x[c * 8 + 2 * e + 1] = bf16_hi(w[e]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          g[i] = __uint_as_float(r[i]);
          dot = fmaf(x[i], g[i], dot);
        }
        dot_part[wg * kKeyRows + j] = dot;
      }
      {
        // accumulator 1 = query rows 128.. (the look-ahead chunk, BUCKET rows) -> dq_b (rows >= BUCKET are never stored)
        uint32_t r[16];
        tmem_ld16(t_row + cDQ1 + col0, r);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]);
        stage16(stage_dqb, o);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kBwdWorkers) : "memory");     // workers only: partial dots complete
      {
        // key-normalisation Jacobian dx = G - x |k|^-2 <x, G>  (dh == 64: the L2 and the RMS/sqrt(dh) norms share |k|^-2 = inv^2),
        // plus this slot's query-role gradient from the keys of this CTA (accumulator 0, row j)
        const float coef = inv * inv * (dot_part[j] + dot_part[kKeyRows + j] + dot_part[2 * kKeyRows + j] + dot_part[3 * kKeyRows + j]);
        uint32_t r[16];
        tmem_ld16(t_row + cDQ0 + col0, r);
        tmem_ld_wait();
        float o[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]) + g[i] - x[i] * coef;
        stage16(stage_dx, o);
      }
      tc_fence_before_sync();       // this thread's reads of the accumulators precede the next tile's MMAs (ordered through bar_blk)
      asm volatile("bar.sync 1, %0;" ::"n"(kBwdWorkers) : "memory");     // staging tiles complete
      {
        // warp w stores rows [8w, 8w+8) of each output: 8 lanes per 128-byte row, four rows per instruction
        const int lane = tid & 31, ch = lane & 7;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
          const int row = warp * 8 + it * 4 + (lane >> 3);
          const int64_t dst_row = (out_base + q_slot[row]) * kBDh;
          const uint32_t so = sw128_offset(row, ch);
          reinterpret_cast<uint4*>(p.dv + dst_row)[ch] = lds128(stage_dv + so);
          reinterpret_cast<uint4*>(p.dqk_main + dst_row)[ch] = lds128(stage_dx + so);
          if (row < BUCKET) reinterpret_cast<uint4*>(p.dq_b + (out_base + q_slot[kKeyRows + row]) * kBDh)[ch] = lds128(stage_dqb + so);
        }
      }
      RTTS_BSTAMP(21);
      // the input stage (x rows, slots) and the staging tiles have been read by this warp; the next tile's dSt writes wait for all
      asm volatile("bar.sync 1, %0;" ::"n"(kBwdWorkers) : "memory");
      if ((tid & 31) == 0) mbar_arrive(in_free + st);
      RTTS_BSTAMP(22);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

// ---------------------------------------------------------------------------------------------------------------------
// dqk[b,t,h,:] = sum_r dq_a + dq_b(if written) + dxk ;  dv[b,t,h,:] = sum_r dv_r.  16 lanes x float4 per (b,h,t) row.
// dq_b is written for every slot when bucket == 128, and only for slots in even chunks when bucket == 64
// (the chunk index comes from `undo`).
// ---------------------------------------------------------------------------------------------------------------------
// RC > 0: the number of rounds is a compile-time constant, so the R `undo` words and all 3 R 16-byte slices of a row are requested
// before anything is added (as in lsh_merge_fwd_kernel: with a rolled loop every round waited for its own `undo` word and then
// for the conditional dq_b load - one dependent chain of three loads per round, 3.9 TB/s).  RC == 0: any R, rolled.
template <int RC>
__global__ void __launch_bounds__(256) lsh_grad_reduce_kernel(const __nv_bfloat16* __restrict__ dqk_main, const __nv_bfloat16* __restrict__ dq_b,
                                                              const __nv_bfloat16* __restrict__ dvr, const int32_t* __restrict__ undo,
                                                              __nv_bfloat16* __restrict__ dqk, __nv_bfloat16* __restrict__ dv, int64_t ld, int T, int H, int R,
                                                              int bucket, int64_t rows) {
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x) >> 3;   // (b*H + h)*T + t
  const int c = threadIdx.x & 7;
  if (row >= rows) return;
  const int64_t bh = row / T;
  const int t = static_cast<int>(row - bh * T);
  float aq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, av[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  auto add8 = [](float* acc, const uint4 u) {
    acc[0] += bf16_lo(u.x); acc[1] += bf16_hi(u.x); acc[2] += bf16_lo(u.y); acc[3] += bf16_hi(u.y);
    acc[4] += bf16_lo(u.z); acc[5] += bf16_hi(u.z); acc[6] += bf16_lo(u.w); acc[7] += bf16_hi(u.w);
  };
  if (RC > 0) {
    constexpr int N = RC > 0 ? RC : 1;
    const int64_t idx0 = bh * RC * T + t;
    bool has_b[N];
    uint4 um[N], uv[N], ub[N];
#pragma unroll
    for (int r = 0; r < N; ++r) has_b[r] = bucket != 64 || ((__ldg(undo + idx0 + static_cast<int64_t>(r) * T) >> 6) & 1) == 0;
#pragma unroll
    for (int r = 0; r < N; ++r) um[r] = __ldg(reinterpret_cast<const uint4*>(dqk_main + (idx0 + static_cast<int64_t>(r) * T) * kBDh) + c);
#pragma unroll
    for (int r = 0; r < N; ++r) uv[r] = __ldg(reinterpret_cast<const uint4*>(dvr + (idx0 + static_cast<int64_t>(r) * T) * kBDh) + c);
#pragma unroll
    for (int r = 0; r < N; ++r)
      ub[r] = has_b[r] ? __ldg(reinterpret_cast<const uint4*>(dq_b + (idx0 + static_cast<int64_t>(r) * T) * kBDh) + c) : make_uint4(0u, 0u, 0u, 0u);
    // (same order of additions as the rolled form: main, v, b per round)
#pragma unroll
    for (int r = 0; r < N; ++r) {
      add8(aq, um[r]);
      add8(av, uv[r]);
      if (has_b[r]) add8(aq, ub[r]);
    }
  } else {
    for (int r = 0; r < R; ++r) {
      const int64_t idx = (bh * R + r) * T + t;
      add8(aq, __ldg(reinterpret_cast<const uint4*>(dqk_main + idx * kBDh) + c));
      add8(av, __ldg(reinterpret_cast<const uint4*>(dvr + idx * kBDh) + c));
      bool has_b = true;
      if (bucket == 64) has_b = ((__ldg(undo + idx) >> 6) & 1) == 0;
      if (has_b) add8(aq, __ldg(reinterpret_cast<const uint4*>(dq_b + idx * kBDh) + c));
    }
  }
  const int64_t b = bh / H;
  const int h = static_cast<int>(bh - b * H);
  const int64_t o = (b * T + t) * ld + h * kBDh + c * 8;
  *reinterpret_cast<uint4*>(dqk + o) = make_uint4(pack_bf16(aq[0], aq[1]), pack_bf16(aq[2], aq[3]), pack_bf16(aq[4], aq[5]), pack_bf16(aq[6], aq[7]));
  *reinterpret_cast<uint4*>(dv + o) = make_uint4(pack_bf16(av[0], av[1]), pack_bf16(av[2], av[3]), pack_bf16(av[4], av[5]), pack_bf16(av[6], av[7]));
}

static long long* g_bwd_trace = nullptr;   // debug only (rtts_debug_set_bwd_trace)

template <int BUCKET>
int launch_attn_bwd(const AttnBwdParams& p, int ctas, cudaStream_t stream) {
  using L = AttnBwdSmem<BUCKET>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lsh_attn_bwd_kernel<BUCKET>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
    if (e != cudaSuccess) return fail(kErrCuda, "rtts_lsh_attn_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  const int grid = ctas < kNumSMs ? ctas : kNumSMs;     // persistent: one CTA per SM
  lsh_attn_bwd_kernel<BUCKET><<<grid, kBwdThreads, L::kDynamic, stream>>>(p, ctas);
  return check_launch("rtts_lsh_attn_bwd");
}

}  // namespace rtts

using namespace rtts;

extern "C" int rtts_lsh_attn_bwd(const void* qk, const void* v, int64_t ld, const int32_t* sticker, const float* sumsq, const uint8_t* mask,
                                 const rtts_lsh_spec* spec, const void* dout, int64_t ld_dout, const float* lse, const float* delta, void* dqk_main,
                                 void* dq_b, void* dv_rounds, int B, int T, int H, int dh, int R, int bucket,
                                 void* stream) {
  RTTS_REQUIRE(qk && v && sticker && sumsq && spec && dout && lse && delta && dqk_main && dq_b && dv_rounds, "rtts_lsh_attn_bwd: null pointer");
  RTTS_REQUIRE(dh == kBDh, "rtts_lsh_attn_bwd: head size %d unsupported (64 only)", dh);
  RTTS_REQUIRE(bucket == 64 || bucket == 128, "rtts_lsh_attn_bwd: bucket size %d unsupported (64 or 128)", bucket);
  RTTS_REQUIRE(T % (2 * bucket) == 0, "rtts_lsh_attn_bwd: T=%d must be a multiple of 2*bucket", T);
  RTTS_REQUIRE(static_cast<int64_t>(T) * ld < (1ll << 31) && static_cast<int64_t>(T) * ld_dout < (1ll << 31), "rtts_lsh_attn_bwd: T * ld must be below 2^31");
  RTTS_REQUIRE(ld % 8 == 0 && ld_dout % 8 == 0 && ((reinterpret_cast<uintptr_t>(qk) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(dout)) & 15) == 0,
               "rtts_lsh_attn_bwd: tensors must be 16-byte aligned");
  AttnBwdParams p;
  p.qk = static_cast<const __nv_bfloat16*>(qk);
  p.v = static_cast<const __nv_bfloat16*>(v);
  p.dout = static_cast<const __nv_bfloat16*>(dout);
  p.ld = ld; p.ld_do = ld_dout; p.sticker = sticker; p.sumsq = sumsq; p.mask = mask; p.lse = lse; p.delta = delta;
  p.dqk_main = static_cast<__nv_bfloat16*>(dqk_main); p.dq_b = static_cast<__nv_bfloat16*>(dq_b); p.dv = static_cast<__nv_bfloat16*>(dv_rounds);
  p.trace = g_bwd_trace;
  p.T = T; p.H = H; p.R = R; p.tiles_per_row = R * T / kKeyRows;
  p.score_scale = spec->score_scale;
  p.mask_value_log2 = fmaxf(spec->mask_value * kBLog2e, -3.0e38f);
  p.self_value_log2 = spec->self_value * kBLog2e;
  p.key_norm = spec->key_norm; p.mask_mode = spec->mask_mode; p.causal = spec->causal;
  const int64_t ctas = static_cast<int64_t>(B) * H * p.tiles_per_row;
  RTTS_REQUIRE(ctas > 0 && ctas < (1ll << 31), "rtts_lsh_attn_bwd: bad grid");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return bucket == 64 ? launch_attn_bwd<64>(p, static_cast<int>(ctas), s) : launch_attn_bwd<128>(p, static_cast<int>(ctas), s);
}

extern "C" int rtts_lsh_grad_reduce(const void* dqk_main, const void* dq_b, const void* dv_rounds, const int32_t* undo, void* dqk, void* dv, int64_t ld, int B, int T, int H, int dh, int R,
                                    int bucket, void* stream) {
  RTTS_REQUIRE(dqk_main && dq_b && dv_rounds && dqk && dv, "rtts_lsh_grad_reduce: null pointer");
  RTTS_REQUIRE(dh == kBDh && ld % 8 == 0, "rtts_lsh_grad_reduce: head size 64 and 16-byte rows required");
  RTTS_REQUIRE(bucket == 128 || (bucket == 64 && undo), "rtts_lsh_grad_reduce: bucket 64 needs undo");
  const int64_t rows = static_cast<int64_t>(B) * H * T;
  const int64_t blocks = (rows * 8 + 255) / 256;
  const auto launch = [&](auto kernel) {
    kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(dqk_main), static_cast<const __nv_bfloat16*>(dq_b), static_cast<const __nv_bfloat16*>(dv_rounds), undo,
        static_cast<__nv_bfloat16*>(dqk), static_cast<__nv_bfloat16*>(dv), ld, T, H, R, bucket, rows);
  };
  switch (R) {      // the reference configs use 8 (RP / HF default) and 4 (long-sequence sweep) hash rounds
    case 8: launch(lsh_grad_reduce_kernel<8>); break;
    case 4: launch(lsh_grad_reduce_kernel<4>); break;
    case 2: launch(lsh_grad_reduce_kernel<2>); break;
    default: launch(lsh_grad_reduce_kernel<0>); break;
  }
  return check_launch("rtts_lsh_grad_reduce");
}

// Debug hook (not part of the product ABI): device buffer of 32 int64 receiving clock64 stamps of CTA 0, thread 0.
extern "C" void rtts_debug_set_bwd_trace(void* device_buffer) { g_bwd_trace = static_cast<long long*>(device_buffer); }
