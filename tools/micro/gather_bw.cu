// Microbenchmark: how fast can one CTA per SM gather 128-byte rows (token stride 2048 B) from an L2-resident
// buffer into shared memory?  Row indices are computed (no dependent index load), so this is the pure gather rate.
//   LDGSTS warps: 16-byte cp.async, 8 lanes per row, `depth` commit groups in flight
//   TMA warps   : tile::gather4 (four rows per instruction) on an mbarrier
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ int row_of(int x, int tokens) { return (int)((((unsigned)x * 2654435761u) >> 8) & (unsigned)(tokens - 1)); }

__global__ void gather_kernel(const uint4* __restrict__ src, const __grid_constant__ CUtensorMap tm, int tokens, int rows_ldgsts, int rows_tma,
                              int ldgsts_warps, int tiles, int depth, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)sm_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar + i)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int col = (blockIdx.x & 15) * 64;   // one "head" per CTA
  long long t0 = clock64();
  if (warp < ldgsts_warps) {
    const int c = tid & 7, g = tid >> 3, ngrp = ldgsts_warps * 4;
    for (int t = 0; t < tiles; ++t) {
      uint8_t* dst = sm + (t & 1) * 49152;
      for (int r = g; r < rows_ldgsts; r += ngrp) {
        const int row = row_of((blockIdx.x * tiles + t) * 512 + r, tokens);
        cp_async16(smem_u32(dst + r * 128 + ((c ^ (r & 7)) << 4)), src + (size_t)row * 128 + col / 8 + c);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (depth == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
      if (depth == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
      if (depth == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  } else if (rows_tma > 0) {
    // TMA warps: rows [rows_ldgsts, rows_ldgsts + rows_tma) of each tile, 4 per instruction, spread over the lanes
    const int tw = warp - ldgsts_warps, ntw = (blockDim.x >> 5) - ldgsts_warps;
    for (int t = 0; t < tiles; ++t) {
      uint8_t* dst = sm + (t & 1) * 49152;
      uint64_t* b = bar + (t & 1);
      if (tw == 0 && lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(rows_tma * 128) : "memory");
      __syncwarp();
      for (int q = tw * 32 + lane; q < rows_tma / 4; q += ntw * 32) {
        const int r = rows_ldgsts + q * 4;
        const int x = (blockIdx.x * tiles + t) * 512 + r;
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst + r * 128)),
            "l"(&tm), "r"(smem_u32(b)), "r"(col), "r"(row_of(x, tokens)), "r"(row_of(x + 1, tokens)), "r"(row_of(x + 2, tokens)), "r"(row_of(x + 3, tokens))
            : "memory");
      }
      if (tw == 0) {   // wait for this tile before reusing the barrier two tiles later (depth 2)
        if (t >= 1) {
          uint32_t ok = 0;
          while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(smem_u32(bar + ((t - 1) & 1))), "r"(((t - 1) >> 1) & 1) : "memory");
        }
      }
      asm volatile("bar.sync 1, %0;" ::"r"(ntw * 32) : "memory");
    }
  }
  __syncthreads();
  if (tid == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int tokens = 16 * 1024, tiles = 64, sms = 148;
  uint4* src; long long* cyc;
  cudaMalloc(&src, (size_t)tokens * 2048);
  cudaMemset(src, 1, (size_t)tokens * 2048);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  void* sym; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  CUtensorMap tm;
  const cuuint64_t gdim[2] = {1024, (cuuint64_t)tokens}; const cuuint64_t gstr[1] = {2048};
  const cuuint32_t box[2] = {64, 1}; const cuuint32_t estr[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)sym)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, src, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int smem = 2 * 49152 + 1024;
  cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Cfg { int grid, lw, tw, rows_l, rows_t, depth; };
  const Cfg cfgs[] = {{148, 0, 4, 0, 384, 2}, {148, 0, 8, 0, 384, 2}, {148, 0, 12, 0, 384, 2}, {148, 0, 16, 0, 384, 2}, {8, 0, 16, 0, 384, 2},
                      {148, 4, 4, 192, 192, 2}, {148, 4, 8, 192, 192, 2}, {148, 4, 8, 128, 256, 2}, {148, 2, 8, 128, 256, 2}};
  for (const Cfg& cf : cfgs) {
    for (int rep = 0; rep < 2; ++rep)
      gather_kernel<<<cf.grid, (cf.lw + cf.tw) * 32, smem>>>(src, tm, tokens, cf.rows_l, cf.rows_t, cf.lw, tiles, cf.depth, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long hc[148];
    cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < cf.grid; ++i) avg += hc[i]; avg /= cf.grid;
    const int rows = cf.rows_l + cf.rows_t;
    printf("grid %3d ldgsts warps %2d (%3d rows) tma warps %d (%3d rows) depth %d: %6.0f cycles per %d-row tile = %5.1f B/clk/SM  %s\n", cf.grid, cf.lw,
           cf.rows_l, cf.tw, cf.rows_t, cf.depth, avg / tiles, rows, rows * 128.0 * tiles / avg, cudaGetErrorString(e));
  }
  return 0;
}
