// tcgen05.mma issue/execution cost per instruction for the shapes the attention kernel uses (one issuing thread per SM):
//   SS  128 x N x 16 (S = Q K^T) for several N and D column offsets, 4 dependent k-steps per "block";
//   TS  128 x 64 x 16 (O += P V), A from TMEM, chains of 7 dependent MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I keep_b200/csrc -o /tmp/mma_rate tools/microbench/mma_rate.cu
#include "ptx.cuh"
#include <cstdio>
using namespace kb;

__global__ void __launch_bounds__(128, 1) mma_kernel(int mode, int N, int dcol, int chain, int reps, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t tb = slot;
    const uint64_t da = make_smem_desc_sw128(smem_u32(smem)), db = make_smem_desc_sw128(smem_u32(smem + 16384));
    const uint32_t idesc_ss = make_idesc(kFmtF16, 128, N, 0, 0), idesc_ts = make_idesc(kFmtF16, 128, 64, 0, 1);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tb + ((r & 1) ? dcol : 0);
      if (mode == 0) {
        for (int k = 0; k < chain; ++k) umma_f16_ss(d, da + 2 * (k & 3), db + 2 * (k & 3), idesc_ss, k != 0);
      } else {
        for (int k = 0; k < chain; ++k) umma_f16_ts(tb + 448, tb + 8 * k + ((r & 1) ? dcol : 0), db + 128 * (k & 7), idesc_ts, k != 0);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    cycles[blockIdx.x * 2] = t1 - t0;
    cycles[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 16);
  const int smem = 16384 + 32768 + 1024;
  cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 256;
  struct Case { int mode, N, dcol, chain; const char* what; };
  const Case cases[] = {
      {0, 64, 256, 4, "SS N=64"},   {0, 96, 256, 4, "SS N=96"},   {0, 112, 256, 4, "SS N=112"}, {0, 112, 112, 4, "SS N=112 D@112"},
      {0, 128, 256, 4, "SS N=128"}, {0, 208, 208, 4, "SS N=208 D@208"}, {0, 224, 256, 4, "SS N=224"}, {0, 256, 256, 4, "SS N=256"},
      {1, 64, 112, 7, "TS N=64 chain 7, A@0/112"}, {1, 64, 208, 13, "TS N=64 chain 13, A@0/208"}, {1, 64, 256, 7, "TS N=64 chain 7, A@0/256"},
  };
  for (const Case& c : cases) {
    for (int rep = 0; rep < 2; ++rep) mma_kernel<<<148, 128, smem>>>(c.mode, c.N, c.dcol, c.chain, reps, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-28s: issue %.1f clk/MMA, issue+drain %.1f clk/MMA  (%s)\n", c.what, (double)h[0] / (reps * c.chain),
           (double)h[1] / (reps * c.chain), cudaGetErrorString(e));
  }
  return 0;
}
