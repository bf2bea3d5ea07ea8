// MUFU.EX2 issue rate per SM as a function of resident warps (independent chains), plus FFMA+EX2+FADD+pack mix as in softmax.
#include <cstdio>
#include <cuda_fp16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(int iters, float a, long long* cyc, float* sink) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -0.001f * (threadIdx.x + i);
  float s0 = 0.f, s1 = 0.f; unsigned pkacc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float e0, e1;
      if (MODE == 0) { e0 = ex2f(v[2 * i]); e1 = ex2f(v[2 * i + 1]); }
      else { e0 = ex2f(fmaf(v[2 * i], a, -1.f)); e1 = ex2f(fmaf(v[2 * i + 1], a, -1.f)); }
      s0 += e0; s1 += e1;
      if (MODE == 1) { __half2 h = __floats2half2_rn(e0, e1); pkacc ^= *reinterpret_cast<unsigned*>(&h); }
      v[2 * i] = e0 * -0.5f; v[2 * i + 1] = e1 * -0.5f;   // next iteration depends on this one: chains of length iters, 32 wide
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s0 + s1 == 123.f) sink[0] = s0 + pkacc;
}
int main() {
  long long* cyc; float* sink; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8, 16}) {
      for (int r = 0; r < 2; ++r) { if (mode == 0) k<0><<<148, warps * 32>>>(iters, 1.01f, cyc, sink); else k<1><<<148, warps * 32>>>(iters, 1.01f, cyc, sink); }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      printf("mode %d (%s) warps=%2d: %lld cycles, %.2f ex2/clk/SM, %.1f clk per 32-wide warp-chunk\n", mode, mode ? "ffma+ex2+fadd+pack" : "ex2 only",
             warps, h[0], (double)warps * 32 * 32 * iters / h[0], (double)h[0] / iters);
    }
  return 0;
}
