// TMEM read / write bandwidth per SM: W warps (warp % 4 = lane quarter) each issue `iters` tcgen05.ld (x32 / x16) or
// tcgen05.st back to back; prints bytes per clock per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I reformer_tts_b200/csrc -o tools/micro/tmem_bw tools/micro/tmem_bw.cu
#include <cstdio>
#include "common.cuh"
using namespace rtts;

template <int MODE>
__global__ void k(long long* out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t t = (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 64;
  uint32_t r[32];
  for (int i = 0; i < 32; ++i) r[i] = i;
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) { tmem_ld32(t + (i & 1) * 32, r); tmem_ld_wait(); acc += r[0] + r[31]; }
    if (MODE == 1) { tmem_ld16(t + (i & 3) * 16, r); tmem_ld_wait(); acc += r[0] + r[15]; }
    if (MODE == 2) { tmem_st16(t + (i & 3) * 16, r); tmem_st_wait(); }
    if (MODE == 3) { tmem_ld32(t, r); tmem_ld32(t + 32, r); tmem_ld_wait(); acc += r[0]; }      // two loads in flight
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[1000] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc(0, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8 * 2048);
  const int iters = 2000;
  for (int warps : {4, 8, 16}) {
    for (int mode = 0; mode < 4; ++mode) {
      if (mode == 0) k<0><<<148, warps * 32>>>(d, iters);
      if (mode == 1) k<1><<<148, warps * 32>>>(d, iters);
      if (mode == 2) k<2><<<148, warps * 32>>>(d, iters);
      if (mode == 3) k<3><<<148, warps * 32>>>(d, iters);
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes_per_iter = (mode == 0 ? 32 : mode == 3 ? 64 : 16) * 4.0 * 32 * warps;
      printf("warps %2d  %s : %8.1f cycles/iter  %7.1f B/clk/SM  (%s)\n", warps, mode == 0 ? "ld.x32      " : mode == 1 ? "ld.x16      " : mode == 2 ? "st.x16      " : "2 x ld.x32  ",
             double(h[0]) / iters, bytes_per_iter * iters / double(h[0]), cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
