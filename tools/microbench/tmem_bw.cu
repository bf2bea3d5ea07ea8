// TMEM read/write throughput per SM (tcgen05.ld / tcgen05.st 32x32b.x32), as a function of the number of warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I keep_b200/csrc -o /tmp/tmem_bw tools/microbench/tmem_bw.cu
#include "ptx.cuh"
#include <cstdio>
using namespace kb;

__global__ void __launch_bounds__(512, 1) bw_kernel(int iters, int mode, long long* cycles, uint32_t* sink) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + (uint32_t((warp & 3) * 32) << 16);
  uint32_t v[32], w[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { v[i] = threadIdx.x + i; w[i] = i; }
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) {
      tmem_ld_32x32(base + ((it * 64) & 511), v);
      tmem_ld_32x32(base + ((it * 64 + 32) & 511), w);
      tmem_ld_wait_dep(v);
      tmem_ld_wait_dep(w);
      acc += v[0] ^ v[31] ^ w[0] ^ w[31];
    }
  } else {
    for (int it = 0; it < iters; ++it) {
      const uint32_t a = base + ((it * 32) & 511);
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
          "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
          ::"r"(a), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
      if ((it & 3) == 3) tmem_st_wait();
    }
    tmem_st_wait();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 0xdeadbeef) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps : {1, 2, 4, 8, 12, 16}) {
      for (int rep = 0; rep < 2; ++rep) bw_kernel<<<148, warps * 32, 0>>>(iters, mode, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[148];
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = (double)warps * iters * (mode == 0 ? 2 : 1) * 32 * 32 * 4;
      printf("%s warps=%2d: %lld cycles, %.1f B/clk/SM  (%s)\n", mode == 0 ? "tcgen05.ld" : "tcgen05.st", warps, h[0],
             bytes / (double)h[0], cudaGetErrorString(e));
    }
  return 0;
}
