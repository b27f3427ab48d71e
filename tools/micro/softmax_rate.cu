// Microbenchmark: issue / pipe cost of the forward kernel's 16-column softmax chunk (FFMA scale, EX2, position mask, self mask,
// row sum, bf16 pack) on register data, W warps per CTA (one CTA per SM), no TMEM / barriers.  Prints cycles per chunk per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 --use_fast_math -o softmax_rate softmax_rate.cu && ./softmax_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}

// MODE bits: 1 = exp2, 2 = position mask (ISETP+FSEL), 4 = self mask, 8 = pack, 16 = mask by FADD.SAT + FFMA instead of ISETP+FSEL
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(float* out, long long* cyc, int iters) {
  __shared__ __align__(16) float scale[256];
  __shared__ __align__(16) int pos[256];
  __shared__ __align__(16) float posf[256];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) { scale[i] = 1.f + 1e-3f * i; pos[i] = (i * 37) & 1023; posf[i] = float((i * 37) & 1023); }
  __syncthreads();
  float x[16];
  for (int i = 0; i < 16; ++i) x[i] = 1e-3f * (threadIdx.x + i);
  float sum4[4] = {0, 0, 0, 0};
  uint32_t acc = 0;
  const int q_limit = 512 + (threadIdx.x & 127), q_enc = threadIdx.x & 1023;
  const float q_limf = float(q_limit);
  const float neg_m = -3.f;
  const uint32_t a_scale = (uint32_t)__cvta_generic_to_shared(scale), a_pos = (uint32_t)__cvta_generic_to_shared(pos),
                 a_posf = (uint32_t)__cvta_generic_to_shared(posf);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const int c = (it & 15) * 64;
    float y[16];
    uint4 s[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) s[q] = lds128(a_scale + c + q * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      y[q * 4 + 0] = fmaf(x[q * 4 + 0], __uint_as_float(s[q].x), neg_m); y[q * 4 + 1] = fmaf(x[q * 4 + 1], __uint_as_float(s[q].y), neg_m);
      y[q * 4 + 2] = fmaf(x[q * 4 + 2], __uint_as_float(s[q].z), neg_m); y[q * 4 + 3] = fmaf(x[q * 4 + 3], __uint_as_float(s[q].w), neg_m);
    }
    if (MODE & 16) {
      uint4 kq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) kq[q] = lds128(a_posf + c + q * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float kp[4] = {__uint_as_float(kq[q].x), __uint_as_float(kq[q].y), __uint_as_float(kq[q].z), __uint_as_float(kq[q].w)};
#pragma unroll
        for (int i = 0; i < 4; ++i) y[q * 4 + i] = fmaf(__saturatef(kp[i] - q_limf), -1e30f, y[q * 4 + i]);
      }
    }
    if (MODE & 1) {
#pragma unroll
      for (int i = 0; i < 16; ++i) y[i] = exp2f(y[i]);
    }
    if (MODE & 6) {
      uint4 kq[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) kq[q] = lds128(a_pos + c + q * 16);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int kp[4] = {(int)kq[q].x, (int)kq[q].y, (int)kq[q].z, (int)kq[q].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (MODE & 2) y[q * 4 + i] = kp[i] > q_limit ? 0.f : y[q * 4 + i];
          if (MODE & 4) y[q * 4 + i] = kp[i] == q_enc ? 0.f : y[q * 4 + i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) sum4[i & 3] += y[i];
    if (MODE & 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc ^= pack_bf16(y[2 * i], y[2 * i + 1]);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] += 1e-6f * sum4[i & 3];      // dependence to the next iteration (cheap, FMA pipe)
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = sum4[0] + sum4[1] + sum4[2] + sum4[3] + __uint_as_float(acc);
}

template <int MODE>
void run(const char* name) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int warps : {4, 8, 16, 32}) {
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    k<MODE><<<148, warps * 32>>>(out, cyc, iters);
    long long h = 0;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SMSP=%d  cycles/chunk/warp=%7.1f  cycles/chunk/SMSP=%7.1f\n", name, warps / 4, double(h) / iters, double(h) / iters / (warps / 4));
  }
}

int main() {
  run<0>("ffma + sum");
  run<1>("+ exp2");
  run<1 | 8>("+ exp2 + pack");
  run<1 | 2 | 8>("+ exp2 + mask(ISETP/FSEL) + pack");
  run<1 | 2 | 4 | 8>("+ exp2 + mask + self + pack");
  run<1 | 16 | 8>("+ exp2 + mask(FADD.SAT/FFMA) + pack");
  run<2 | 8>("mask(ISETP/FSEL) + pack, no exp2");
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
