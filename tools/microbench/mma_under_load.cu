// What a single-thread role (tcgen05.mma issue, mbarrier probes) costs when its SM sub-partition is shared with
// softmax-like warps: the attention kernel's MMA thread needs ~110 clk per TS MMA and 100-400 clk per mbarrier round
// trip, against 66 clk per MMA on an idle SM (mma_rate.cu). This separates the candidates:
//   load 0: no other work          load 1: MUFU.EX2 loop (special-function ops go through the MIO queue)
//   load 2: FFMA loop (FMA pipe)   load 3: tcgen05.ld loop          load 4: softmax-like mix (ld + FFMA + MUFU + st)
// on 8 load warps (two per sub-partition) or 6 (sub-partition 3 left to the issuing warp 15).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I keep_b200/csrc -o /tmp/mma_under_load tools/microbench/mma_under_load.cu
#include "ptx.cuh"
#include <cstdio>
using namespace kb;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(512, 1) k(int load, int skip_q3, int what, int reps, long long* cycles, float* sink,
                                            volatile int* stop_flag) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar, bar2;
  __shared__ int done;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); done = 0; fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tb = slot;
  if (warp == 15) {
    if (lane == 0) {
      const uint64_t db = make_smem_desc_sw128(smem_u32(smem + 16384));
      const uint32_t idesc_ts = make_idesc(kFmtF16, 128, 64, 0, 1);
      long long t_mma = 0, t_probe = 0;
      uint32_t acc = 0;
      for (int r = 0; r < reps; ++r) {
        long long t0 = clock64();
        if (what & 1)
          for (int kk = 0; kk < 13; ++kk) umma_f16_ts(tb + 448, tb + 8 * kk, db + 128 * (kk & 7), idesc_ts, kk != 0);
        long long t1 = clock64();
        if (what & 2)
          for (int kk = 0; kk < 8; ++kk) acc += mbar_test_wait(&bar2, kk & 1) ? 1u : 0u;  // 8 non-blocking probes
        long long t2 = clock64();
        t_mma += t1 - t0;
        t_probe += t2 - t1;
      }
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      cycles[blockIdx.x * 2] = t_mma;
      cycles[blockIdx.x * 2 + 1] = t_probe + (acc == 12345u);
      *(volatile int*)&done = 1;
    }
  } else if (warp >= 4 && warp < 12 && load != 0 && !(skip_q3 && (warp & 3) == 3)) {
    const uint32_t t_row = tb + (uint32_t((warp & 3) * 32) << 16) + ((warp >= 8) ? 208u : 0u);
    float a = lane * 0.001f, b = 0.5f, c = 0.25f, d = 0.125f;
    while (*(volatile int*)&done == 0) {
      if (load == 1) {
#pragma unroll
        for (int i = 0; i < 64; ++i) { a = ex2f(a); b = ex2f(b); c = ex2f(c); d = ex2f(d); }
      } else if (load == 2) {
#pragma unroll
        for (int i = 0; i < 64; ++i) { a = fmaf(a, 1.0001f, 0.5f); b = fmaf(b, 0.9999f, a); c = fmaf(c, 1.0001f, b); d = fmaf(d, 0.9999f, c); }
      } else if (load == 3) {
        uint32_t v[32];
        tmem_ld_32x32(t_row, v);
        tmem_ld_wait_dep(v);
        a += __uint_as_float(v[lane & 31]);
      } else {
        uint32_t v[32];
        tmem_ld_32x32(t_row, v);
        tmem_ld_wait_dep(v);
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float e0 = ex2f(fmaf(__uint_as_float(v[2 * i]), 0.18f, -3.f)), e1 = ex2f(fmaf(__uint_as_float(v[2 * i + 1]), 0.18f, -3.f));
          a += e0 + e1;
          __half2 h = __floats2half2_rn(e0, e1);
          pk[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        tmem_st_32x16(t_row, pk);
      }
    }
    if (a + b + c + d == 1234.5f) sink[threadIdx.x] = a;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* cyc;
  float* sink;
  int* flag;
  cudaMalloc(&cyc, 148 * 16);
  cudaMalloc(&sink, 4096);
  cudaMalloc(&flag, 4);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 64;
  const char* names[] = {"idle", "MUFU loop", "FFMA loop", "tcgen05.ld loop", "softmax-like mix"};
  for (int skip = 0; skip < 2; ++skip)
    for (int load = 0; load < 5; ++load) {
      if (skip && load == 0) continue;
      for (int what = 1; what <= 2; ++what) {
        for (int rep = 0; rep < 2; ++rep) k<<<148, 512, smem>>>(load, skip, what, reps, cyc, sink, flag);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2];
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        if (what == 1)
          printf("%-18s load warps %d: %7.1f clk per TS MMA (13-chains)   (%s)\n", names[load], skip ? 6 : 8, (double)h[0] / (reps * 13),
                 cudaGetErrorString(e));
        else
          printf("%-18s load warps %d: %7.1f clk per mbarrier probe       (%s)\n", names[load], skip ? 6 : 8, (double)h[1] / (reps * 8),
                 cudaGetErrorString(e));
      }
    }
  return 0;
}
