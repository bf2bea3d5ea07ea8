// Fused fp32 tails of the two towers, one launch each (north-star "fused LayerNorm + projection + L2-normalise"):
//
//   visual_head : out = normalize(W2 . gelu_erf(W0 . LayerNorm(x) + b0) + b2)      quick_start/keep_inference.py:42-46, 54-58
//                 (x = CLS row of the last ViT block; the LayerNorm is timm's final `norm`, global_pool='token')
//   pooler      : out = normalize(tanh(Wp . x + bp))                               keep_inference.py:60-62 (BertPooler)
//
// These layers are 2.75 MFLOP per tile / 1.18 MFLOP per prompt (0.002 % of a tower), they sit AFTER the last
// LayerScale / post-LN and so pass their rounding error straight to the embedding, and at batch 1 (the reference's own
// call pattern, WSI_evaluation/utils.py:67-74) their launch count is their latency. So they run in plain fp32 on the
// CUDA cores - the reference's own arithmetic, no operand rounding at all - as ONE kernel: a CTA keeps R rows in shared
// memory through LayerNorm -> Linear -> activation -> Linear -> L2-normalise; weights are read K-major-transposed
// ([K, N] fp32, prepared at finalize) so that a warp's loads are contiguous, and each thread owns up to 4 output
// columns x R rows of accumulators.
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

constexpr int kHeadThreads = 256;
constexpr int kMaxCols = 4;  // output columns per thread: widths up to 1024

struct HeadParams {
  const float* x; long long ldx; long long n;
  int K0;                                            // input width
  const float* lnw; const float* lnb; float eps;     // LayerNorm over K0 (null = none)
  const float* wt0; const float* b0; int N0; int act0;  // [K0, N0]; act 0 none, 1 gelu(erf), 2 tanh
  const float* wt1; const float* b1; int N1;         // optional second Linear [N0, N1] (null = none)
  float* out;                                        // [n, N1 or N0], unit L2 norm
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// acc[j][r] = sum_k in[r][k] * wt[k, tid + 256 j]. The weight rows of the NEXT four k are requested before the FMAs of the
// current four run (register double buffer): with 8 warps per SM and ~300 cycles to L2 the loop is otherwise latency-bound.
template <int R>
__device__ __forceinline__ void linear_rows(const float* __restrict__ wt, int K, int N, const float* in, int in_ld,
                                            float (&acc)[kMaxCols][R]) {
#pragma unroll
  for (int j = 0; j < kMaxCols; ++j)
#pragma unroll
    for (int r = 0; r < R; ++r) acc[j][r] = 0.f;
  const int tid = threadIdx.x;
  float w[kMaxCols][4], wn[kMaxCols][4];
  auto load = [&](float (&dst)[kMaxCols][4], int k) {
#pragma unroll
    for (int j = 0; j < kMaxCols; ++j) {
      const int c = tid + kHeadThreads * j;
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[j][i] = (c < N && k + i < K) ? __ldg(wt + (long long)(k + i) * N + c) : 0.f;
    }
  };
  load(w, 0);
  for (int k = 0; k < K; k += 4) {
    load(wn, k + 4);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 xv = *reinterpret_cast<const float4*>(in + r * in_ld + k);  // same address in every thread: broadcast
#pragma unroll
      for (int j = 0; j < kMaxCols; ++j) {
        acc[j][r] = fmaf(xv.x, w[j][0], acc[j][r]);
        acc[j][r] = fmaf(xv.y, w[j][1], acc[j][r]);
        acc[j][r] = fmaf(xv.z, w[j][2], acc[j][r]);
        acc[j][r] = fmaf(xv.w, w[j][3], acc[j][r]);
      }
    }
#pragma unroll
    for (int j = 0; j < kMaxCols; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) w[j][i] = wn[j][i];
  }
}

template <int R>
__global__ void __launch_bounds__(kHeadThreads) head_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float hsm[];
  float* xs = hsm;                 // [R][K0]
  float* hs = hsm + R * p.K0;      // [R][N0]
  float* red = hs + R * p.N0;      // [8 warps][R]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long row0 = (long long)blockIdx.x * R;

  // ---- rows -> shared memory (+ LayerNorm: one warp per row, two-pass statistics as torch.nn.LayerNorm) ----
  for (int r = warp; r < R; r += kHeadThreads / 32) {
    const long long row = row0 + r;
    float* xr = xs + r * p.K0;
    if (row >= p.n) {
      for (int k = lane; k < p.K0; k += 32) xr[k] = 0.f;
      continue;
    }
    const float* src = p.x + row * p.ldx;
    float s = 0.f;
    for (int k = lane; k < p.K0; k += 32) {
      const float v = src[k];
      xr[k] = v;
      s += v;
    }
    if (p.lnw != nullptr) {
      const float mean = warp_sum(s) / (float)p.K0;
      float q = 0.f;
      for (int k = lane; k < p.K0; k += 32) {
        const float d = xr[k] - mean;
        q += d * d;
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)p.K0 + p.eps);
      for (int k = lane; k < p.K0; k += 32) xr[k] = (xr[k] - mean) * rstd * __ldg(p.lnw + k) + __ldg(p.lnb + k);
    }
  }
  __syncthreads();

  float acc[kMaxCols][R];
  linear_rows<R>(p.wt0, p.K0, p.N0, xs, p.K0, acc);
#pragma unroll
  for (int j = 0; j < kMaxCols; ++j) {
    const int c = tid + kHeadThreads * j;
    if (c < p.N0) {
      const float b = p.b0 != nullptr ? __ldg(p.b0 + c) : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float v = acc[j][r] + b;
        if (p.act0 == 1) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
        else if (p.act0 == 2) v = tanhf(v);
        acc[j][r] = v;
        hs[r * p.N0 + c] = v;
      }
    }
  }
  int Nout = p.N0;
  if (p.wt1 != nullptr) {
    __syncthreads();
    linear_rows<R>(p.wt1, p.N0, p.N1, hs, p.N0, acc);
    Nout = p.N1;
#pragma unroll
    for (int j = 0; j < kMaxCols; ++j) {
      const int c = tid + kHeadThreads * j;
      const float b = (c < Nout && p.b1 != nullptr) ? __ldg(p.b1 + c) : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[j][r] += b;
    }
  }
  // ---- L2-normalise each row: F.normalize(dim=-1) = v / max(||v||, 1e-12) ----
  float sq[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxCols; ++j)
      if (tid + kHeadThreads * j < Nout) s = fmaf(acc[j][r], acc[j][r], s);
    sq[r] = warp_sum(s);
  }
  if (lane == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) red[warp * R + r] = sq[r];
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadThreads / 32; ++w) s += red[w * R + r];  // fixed order: deterministic
    const float inv = 1.0f / fmaxf(sqrtf(s), 1e-12f);
    const long long row = row0 + r;
    if (row < p.n) {
#pragma unroll
      for (int j = 0; j < kMaxCols; ++j) {
        const int c = tid + kHeadThreads * j;
        if (c < Nout) p.out[row * Nout + c] = acc[j][r] * inv;
      }
    }
  }
}

template <int R>
int launch_rows(const HeadParams& p, cudaStream_t stream) {
  const int smem = (R * (p.K0 + p.N0) + 8 * R) * 4;
  KB_TRY_ATTR(head_kernel<R>, smem);
  const unsigned grid = (unsigned)((p.n + R - 1) / R);
  head_kernel<R><<<grid, kHeadThreads, smem, stream>>>(p);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch(const HeadParams& p, cudaStream_t stream) {
  if (p.n <= 0) return KB_OK;
  if (p.K0 % 4 != 0 || p.N0 % 4 != 0 || p.K0 > 1024 || p.N0 > kMaxCols * kHeadThreads || p.N1 > kMaxCols * kHeadThreads || p.ldx % 1 != 0)
    return set_error(KB_ERR_ARG, "head: widths %d -> %d -> %d unsupported (multiples of 4, <= 1024)", p.K0, p.N0, p.N1);
  // rows per CTA: enough CTAs to cover the SMs first, then amortise the weight reads over up to 8 rows
  const long long per_sm = (p.n + num_sms() - 1) / num_sms();
  if (per_sm >= 8) return launch_rows<8>(p, stream);
  if (per_sm >= 4) return launch_rows<4>(p, stream);
  if (per_sm >= 2) return launch_rows<2>(p, stream);
  return launch_rows<1>(p, stream);
}

}  // namespace

int launch_visual_head(const float* x, int64_t ldx, int64_t n, int D, const float* lnw, const float* lnb, float eps,
                       const float* w0t, const float* b0, int N0, const float* w1t, const float* b1, int N1, float* out,
                       cudaStream_t stream) {
  HeadParams p;
  p.x = x; p.ldx = ldx; p.n = n; p.K0 = D; p.lnw = lnw; p.lnb = lnb; p.eps = eps;
  p.wt0 = w0t; p.b0 = b0; p.N0 = N0; p.act0 = 1; p.wt1 = w1t; p.b1 = b1; p.N1 = N1; p.out = out;
  return launch(p, stream);
}

int launch_pooler(const float* x, int64_t ldx, int64_t n, int D, const float* wt, const float* b, float* out,
                  cudaStream_t stream) {
  HeadParams p;
  p.x = x; p.ldx = ldx; p.n = n; p.K0 = D; p.lnw = nullptr; p.lnb = nullptr; p.eps = 0.f;
  p.wt0 = wt; p.b0 = b; p.N0 = D; p.act0 = 2; p.wt1 = nullptr; p.b1 = nullptr; p.N1 = 0; p.out = out;
  return launch(p, stream);
}

}  // namespace kb
