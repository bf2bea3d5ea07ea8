// Tile x prompt similarity and prompt screening.
//
//   logits = normalize(feats) @ cls          (detection_utils.py:90-91, subtyping_utils.py:69-70,
//   probs  = softmax(temp * logits, groups)   segment_utils.py:46-47; temp = 10: :93 / :72 / :49)
//   score_k = mean_n(top1 - top2 - |top1 + top2 - 1|)   (WSI_evaluation/utils.py:107-117)
//
// v1: fp32 shared-memory-tiled FMA GEMM (64x64x16 tiles, 4x4 register micro-tiles) with the row
// L2-norm accumulated from the A tiles as they stream through shared memory, so feats are read once.
// The operation is HBM-bound on feats: N*D*4 bytes in, N*P*4 bytes out.
#include "common.h"
#include "ptx.cuh"


namespace kb {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

__global__ void __launch_bounds__(256)
sim_gemm_kernel(const float* __restrict__ feats, long long N, int D, const float* __restrict__ cls, int P,
                float* __restrict__ logits) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float sInv[BM];
  const int t = threadIdx.x;
  const long long row0 = (long long)blockIdx.x * BM;
  const int col0 = blockIdx.y * BN;
  const int a_r = t >> 2, a_k = (t & 3) * 4;   // A loader: row, k offset (float4 along K)
  const int b_k = t >> 4, b_p = (t & 15) * 4;  // B loader: k, 4 consecutive prompts
  const int ty = t >> 4, tx = t & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float sq = 0.f;
  const bool a_ok = (row0 + a_r) < N;
  const float* a_ptr = feats + (row0 + (a_ok ? a_r : 0)) * D + a_k;
  for (int k0 = 0; k0 < D; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_ok && (k0 + a_k) < D) av = *reinterpret_cast<const float4*>(a_ptr + k0);
    sq += (av.x * av.x + av.y * av.y) + (av.z * av.z + av.w * av.w);
    As[a_k + 0][a_r] = av.x; As[a_k + 1][a_r] = av.y; As[a_k + 2][a_r] = av.z; As[a_k + 3][a_r] = av.w;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = col0 + b_p + j, k = k0 + b_k;
      Bs[b_k][b_p + j] = (p < P && k < D) ? __ldg(cls + (long long)k * P + p) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w}, br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
  sq += __shfl_xor_sync(0xffffffffu, sq, 1);
  sq += __shfl_xor_sync(0xffffffffu, sq, 2);
  if ((t & 3) == 0) sInv[a_r] = 1.0f / fmaxf(sqrtf(sq), 1e-12f);
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long r = row0 + ty * 4 + i;
    if (r >= N) continue;
    const float inv = sInv[ty * 4 + i];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int p = col0 + tx * 4 + j;
      if (p < P) logits[r * P + p] = acc[i][j] * inv;
    }
  }
}

// one thread per (row, group)
__global__ void group_softmax_kernel(const float* __restrict__ logits, long long N, int P, int group, float temp,
                                     float* __restrict__ probs) {
  const int ngroups = P / group;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * ngroups) return;
  const long long r = i / ngroups;
  const int g = (int)(i % ngroups);
  const float* src = logits + r * P + g * group;
  float* dst = probs + r * P + g * group;
  float mx = -INFINITY;
  for (int j = 0; j < group; ++j) mx = fmaxf(mx, src[j] * temp);
  float sum = 0.f;
  for (int j = 0; j < group; ++j) sum += expf(src[j] * temp - mx);
  const float inv = 1.0f / sum;
  for (int j = 0; j < group; ++j) dst[j] = expf(src[j] * temp - mx) * inv;
}

// one thread per (row, classifier); every block of 256 rows writes its partial sums to its own row of `part`
// ([gridDim.y, K]): no atomics, so the reduction order - and with it the ranking of near-tied classifiers - is fixed
__global__ void __launch_bounds__(256)
prompt_score_kernel(const float* __restrict__ logits, long long rows, int K, int C, float* __restrict__ parts) {
  // grid: x over classifiers (blocks of 32), y over row blocks of 256; thread layout 32 (k) x 8 (rows)
  const int k = blockIdx.x * 32 + (threadIdx.x & 31);
  const int ry = threadIdx.x >> 5;
  float part = 0.f;
  if (k < K) {
    for (long long r = (long long)blockIdx.y * 256 + ry; r < rows && r < (long long)(blockIdx.y + 1) * 256; r += 8) {
      const float* src = logits + r * (long long)K * C + (long long)k * C;
      float m1 = -INFINITY, m2 = -INFINITY;
      for (int j = 0; j < C; ++j) {
        const float v = src[j];
        if (v > m1) { m2 = m1; m1 = v; } else if (v > m2) { m2 = v; }
      }
      part += (m1 - m2) - fabsf(m1 + m2 - 1.0f);
    }
  }
  __shared__ float red[8][32];
  red[ry][threadIdx.x & 31] = part;
  __syncthreads();
  if (ry == 0 && k < K) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
    parts[(long long)blockIdx.y * K + k] = s;
  }
}

}  // namespace

int launch_similarity(const float* feats, int64_t N, int D, const float* cls, int P, int group, float temp,
                      float* logits, float* probs, cudaStream_t stream, float* clsT_scratch) {
  if (N <= 0 || P <= 0) return KB_OK;
  if (D % 4 != 0) return set_error(KB_ERR_ARG, "similarity: D=%d must be a multiple of 4", D);
  if (logits == nullptr && probs == nullptr) return set_error(KB_ERR_ARG, "similarity: neither logits nor probs requested");
  if (group <= 0) group = P;
  if (P % group != 0) return set_error(KB_ERR_ARG, "similarity: P=%d is not a multiple of group=%d", P, group);
  if (clsT_scratch != nullptr && D % 32 == 0 && (reinterpret_cast<uintptr_t>(feats) & 15) == 0) {
    // tensor-core path: classifier transposed to K-major [P, D] (tiny), TF32 tcgen05 GEMM with fused epilogue
    int rc = launch_transpose_f32(cls, clsT_scratch, D, P, stream);
    if (rc) return rc;
    bool fused = false;
    rc = launch_similarity_tc(feats, N, D, clsT_scratch, P, group, temp, logits, probs, &fused, stream);
    if (rc) return rc;
    if (probs != nullptr && !fused) {
      if (logits == nullptr)
        return set_error(KB_ERR_ARG, "similarity: probabilities without logits need a group size dividing 16 (got %d)", group);
      const long long n = N * (P / group);
      group_softmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(logits, N, P, group, temp, probs);
      note_launch();
      KB_CUDA_CHECK(cudaGetLastError());
    }
    return KB_OK;
  }
  if (logits == nullptr) return set_error(KB_ERR_ARG, "similarity: the fp32 FMA path needs the logits buffer");
  dim3 grid((unsigned)((N + BM - 1) / BM), (unsigned)((P + BN - 1) / BN));
  sim_gemm_kernel<<<grid, 256, 0, stream>>>(feats, N, D, cls, P, logits);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  if (probs != nullptr) {
    const long long n = N * (P / group);
    group_softmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(logits, N, P, group, temp, probs);
    note_launch();
    KB_CUDA_CHECK(cudaGetLastError());
  }
  return KB_OK;
}

// scores[k] = (sum over the row-block partials) / N : fixed order, no atomics
__global__ void score_reduce_kernel(const float* __restrict__ part, long long nparts, int K, float inv_n, float* __restrict__ scores) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= K) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  long long i = 0;
  for (; i + 3 < nparts; i += 4) {
    a0 += part[i * K + k]; a1 += part[(i + 1) * K + k]; a2 += part[(i + 2) * K + k]; a3 += part[(i + 3) * K + k];
  }
  for (; i < nparts; ++i) a0 += part[i * K + k];
  scores[k] = ((a0 + a1) + (a2 + a3)) * inv_n;
}

size_t prompt_scores_fused_workspace_bytes(int64_t N, int64_t D, int64_t K, int64_t C) {
  if (N <= 0 || D <= 0 || K <= 0 || C <= 0) return 0;
  const size_t clsT = ((size_t)K * C * D * 4 + 1023) / 1024 * 1024;
  return clsT + (size_t)((N + 127) / 128) * 4 * K * 4;
}

int launch_prompt_scores_fused(const float* feats, int64_t N, int D, const float* cls, int K, int C, float* scores,
                               void* ws, size_t ws_bytes, cudaStream_t stream) {
  const int P = K * C;
  if (ws == nullptr || ws_bytes < prompt_scores_fused_workspace_bytes(N, D, K, C))
    return set_error(KB_ERR_WORKSPACE, "prompt_scores: workspace too small for the fused path");
  float* clsT = static_cast<float*>(ws);
  float* part = reinterpret_cast<float*>(static_cast<char*>(ws) + ((size_t)P * D * 4 + 1023) / 1024 * 1024);
  int rc = launch_transpose_f32(cls, clsT, D, P, stream);
  if (rc) return rc;
  bool fused = false;
  rc = launch_similarity_tc(feats, N, D, clsT, P, C, 1.0f, nullptr, nullptr, &fused, stream, part);
  if (rc) return rc;
  const long long nparts = ((N + 127) / 128) * 4;
  score_reduce_kernel<<<(unsigned)((K + 127) / 128), 128, 0, stream>>>(part, nparts, K, 1.0f / (float)N, scores);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_prompt_score_partials(const float* logits, int64_t rows, int K, int C, float* part, cudaStream_t stream) {
  if (rows <= 0 || K <= 0) return KB_OK;
  if (C < 2) return set_error(KB_ERR_ARG, "prompt scores: need at least 2 classes per classifier (got %d)", C);
  dim3 grid((unsigned)((K + 31) / 32), (unsigned)((rows + 255) / 256));
  prompt_score_kernel<<<grid, 256, 0, stream>>>(logits, rows, K, C, part);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_score_reduce(const float* part, int64_t nparts, int K, float scale, float* scores, cudaStream_t stream) {
  if (K <= 0) return KB_OK;
  score_reduce_kernel<<<(unsigned)((K + 127) / 128), 128, 0, stream>>>(part, nparts, K, scale, scores);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
