// Tile x prompt similarity on the tensor cores:
//     logits[n,p] = <feats[n,:] / max(||feats[n]||, 1e-12), cls[:,p]>
//     probs       = softmax(temp * logits) over each group of `group` consecutive columns
// (detection_utils.py:90-93, subtyping_utils.py:69-72, segment_utils.py:46-49, keep_inference.py:104).
//
// The operation is HBM-bound on the fp32 features (N*D*4 bytes in, N*P*4 per output), so the kernel is built
// around reading them exactly once: a persistent, warp-specialised TF32 GEMM
//   * TMA streams 128x32 fp32 feature blocks and BNx32 blocks of the (pre-transposed, K-major) classifier
//     through an mbarrier ring; tcgen05.mma.kind::tf32 (UMMA 128 x BN x 8) accumulates fp32 in TMEM, so no
//     conversion pass over the features is needed;
//   * the row norms are taken from the very same shared-memory feature blocks by four "norm" warps while the
//     tensor core consumes them (a thread owns one row; the 128-byte swizzle keeps its reads conflict-free);
//   * four epilogue warps drain the double-buffered accumulator: scale by 1/||x||, write logits, and compute the
//     grouped softmax(temp * .) in registers (group sizes dividing 16; other groupings use the standalone kernel).
// TF32 keeps 10 mantissa bits of each operand: |error| of a cosine of two 768-vectors ~3e-5, against the 1e-3
// similarity gate.
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

constexpr int SBM = 128;   // rows per tile
constexpr int SBK = 32;    // fp32 elements per 128-byte swizzle row
constexpr int kSimThreads = 384;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 -, 4-7 norm, 8-11 epilogue
constexpr int A_BYTES = SBM * 128;

struct SimParams {
  long long N;
  int D, P, BN, stages, group;
  float temp;
  float* logits;
  float* probs;
  float* score_part;  // prompt screening: [ceil(N/128)*4, P/group] partial sums of the top-2 margin term, or null
  uint32_t idesc;
};

// softmax(temp * f) over aligned groups of G consecutive entries of a 16-entry register block
template <int G>
__device__ __forceinline__ void group_softmax16(const float (&f)[16], float (&e)[16], float temp) {
#pragma unroll
  for (int b = 0; b < 16; b += G) {
    float mx = f[b] * temp;
#pragma unroll
    for (int j = 1; j < G; ++j) mx = fmaxf(mx, f[b + j] * temp);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      e[b + j] = expf(f[b + j] * temp - mx);
      sum += e[b + j];
    }
    const float inv = 1.0f / sum;
#pragma unroll
    for (int j = 0; j < G; ++j) e[b + j] *= inv;
  }
}

__global__ void __launch_bounds__(kSimThreads, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const SimParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = p.BN * 128;
  const int stage_bytes = A_BYTES + b_bytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages * A_BYTES;
  float* s_inv = reinterpret_cast<float*>(smem + p.stages * stage_bytes);  // [2][128] 1/||row||
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * SBM);
  uint64_t* full_bar = bars;                 // [stages] TMA -> MMA, norm warps
  uint64_t* empty_bar = bars + p.stages;     // [stages] MMA commit + 4 norm warps -> TMA
  uint64_t* tfull_bar = bars + 2 * p.stages; // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] epilogue -> MMA, norm warps
  uint64_t* nready_bar = tempty_bar + 2;     // [2] norm warps -> epilogue
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(nready_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (int)((p.N + SBM - 1) / SBM);
  const int n_tiles = (p.P + p.BN - 1) / p.BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.D / SBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1 + 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
      mbar_init(&nready_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 41);
          mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
          tma_load_2d(&tmap_a, &full_bar[s], smem_a + s * A_BYTES, kb * SBK, m_blk * SBM);
          tma_load_2d(&tmap_b, &full_bar[s], smem_b + s * b_bytes, kb * SBK, n_blk * p.BN);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        mbar_wait(&tempty_bar[as], ((lt >> 1) & 1) ^ 1, 42);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph, 43);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * b_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 8 tf32 elements = 32 bytes per MMA: +2 in the >>4 address field
            umma_tf32_ss(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== norm warps: sum of squares of each feature row, from the staged blocks ==========
    const int r = (warp - 4) * 32 + lane;  // row of the tile owned by this thread
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int as = lt & 1;
      float sq = 0.f;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[s], ph, 44);
        const uint32_t row_addr = smem_u32(smem_a + s * A_BYTES) + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(row_addr + ((j ^ (r & 7)) << 4)));
          sq += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (++s == p.stages) { s = 0; ph ^= 1; }
      }
      mbar_wait(&tempty_bar[as], ((lt >> 1) & 1) ^ 1, 45);  // the epilogue two tiles back has read s_inv[as]
      s_inv[as * SBM + r] = 1.0f / fmaxf(sqrtf(sq), 1e-12f);
      __syncwarp();
      if (lane == 0) mbar_arrive(&nready_bar[as]);
    }
  } else if (warp >= 8) {
    // ===================== epilogue: scale, logits, grouped softmax =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const bool fused_softmax = p.probs != nullptr && (16 % p.group) == 0;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(&nready_bar[as], aph, 46);
      mbar_wait(&tfull_bar[as], aph, 47);
      tc_fence_after();
      const float inv = s_inv[as * SBM + r];
      const long long row = (long long)m_blk * SBM + r;
      const uint32_t t_row = tmem_base + as * 256 + (uint32_t(q * 32) << 16);
      for (int c0 = 0; c0 < p.BN; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(t_row + c0, v);
        tmem_ld_wait();
        const int col = n_blk * p.BN + c0;
        if (p.score_part != nullptr) {
          // prompt screening (WSI_evaluation/utils.py:107-117, rank_cls_score): per classifier (group of C columns) the
          // term top1 - top2 - |top1 + top2 - 1| of every tile, summed over the 32 rows of this warp; the [N, K*C] logits
          // never leave the SM. Whole warp participates (rows beyond N contribute 0); groups never straddle a block.
          if (col >= p.P) continue;  // warp-uniform
          const int G = p.group, ng = 16 / G;
          const bool row_ok = row < p.N;
          float term[8];
#pragma unroll
          for (int gi = 0; gi < 8; ++gi) {
            float t1 = -INFINITY, t2 = -INFINITY;
            if (gi < ng) {
              for (int i = 0; i < G; ++i) {
                const float x = __uint_as_float(v[gi * G + i]) * inv;
                t2 = fmaxf(t2, fminf(t1, x));
                t1 = fmaxf(t1, x);
              }
            }
            term[gi] = (gi < ng && row_ok && col + gi * G < p.P) ? (t1 - t2) - fabsf(t1 + t2 - 1.0f) : 0.f;
          }
          // sum over the 32 rows of the warp: exchange-and-add butterfly over lane bits 0-2 (7 shuffles leave lane l with
          // the 8-lane partial of group l & 7), then two plain steps over bits 3-4: 9 shuffles instead of 8 x 5
          {
            const bool b1 = (lane & 1) != 0, b2 = (lane & 2) != 0, b4 = (lane & 4) != 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {  // bit 2 of the group index: lanes with bit 2 set keep groups 4..7
              const float keep = b4 ? term[k + 4] : term[k], send = b4 ? term[k] : term[k + 4];
              term[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // bit 1
              const float keep = b2 ? term[k + 2] : term[k], send = b2 ? term[k] : term[k + 2];
              term[k] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            }
            const float keep = b1 ? term[1] : term[0], send = b1 ? term[0] : term[1];
            float t = keep + __shfl_xor_sync(0xffffffffu, send, 1);
            t += __shfl_xor_sync(0xffffffffu, t, 8);
            t += __shfl_xor_sync(0xffffffffu, t, 16);
            const int gi = lane & 7;  // = (bit2, bit1, bit0) of the group this lane ended up with
            if (lane < 8 && gi < ng && col + gi * G < p.P)
              p.score_part[((long long)m_blk * 4 + q) * (p.P / G) + (col / G + gi)] = t;
          }
          continue;
        }
        if (row >= p.N || col >= p.P) continue;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
        const int valid = min(16, p.P - col);
        if (p.logits != nullptr) {
          float* dst = p.logits + row * p.P + col;
          if (valid == 16 && (p.P & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(f[i], f[i + 1], f[i + 2], f[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < valid) dst[i] = f[i];
          }
        }
        if (fused_softmax) {
          // groups never straddle a 16-column block (group | 16; BN and col are multiples of 16)
          float e[16];
          switch (p.group) {
            case 1: group_softmax16<1>(f, e, p.temp); break;
            case 2: group_softmax16<2>(f, e, p.temp); break;
            case 4: group_softmax16<4>(f, e, p.temp); break;
            case 8: group_softmax16<8>(f, e, p.temp); break;
            default: group_softmax16<16>(f, e, p.temp); break;
          }
          float* dst = p.probs + row * p.P + col;
          if (valid == 16 && (p.P & 3) == 0) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(e[i], e[i + 1], e[i + 2], e[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i < valid) dst[i] = e[i];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

// clsT: fp32 [P, D] (classifier transposed to K-major). Returns KB_ERR_ARG if the shape is not supported
// (caller falls back to the FMA kernel); *fused_probs tells whether probs were produced here.
int launch_similarity_tc(const float* feats, int64_t N, int D, const float* clsT, int P, int group, float temp,
                         float* logits, float* probs, bool* fused_probs, cudaStream_t stream, float* score_part) {
  *fused_probs = false;
  if (score_part != nullptr && (group < 2 || group > 16 || 16 % group != 0 || P % group != 0))
    return set_error(KB_ERR_ARG, "similarity_tc: fused screening needs a group size of 2, 4, 8 or 16 (got %d)", group);
  if (D % SBK != 0) return set_error(KB_ERR_ARG, "similarity_tc: D=%d not a multiple of 32", D);
  int BN = P >= 256 ? 256 : (P + 15) / 16 * 16;
  const int b_bytes = BN * 128;
  int stages = (227 * 1024 - 4096) / (A_BYTES + b_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) return set_error(KB_ERR_ARG, "similarity_tc: tile too large");
  CUtensorMap ta, tb;
  int rc = get_tmap_2d(feats, KB_F32, N, D, D, SBM, &ta);
  if (rc) return rc;
  rc = get_tmap_2d(clsT, KB_F32, P, D, D, BN, &tb);
  if (rc) return rc;
  const int smem = stages * (A_BYTES + b_bytes) + 2 * SBM * 4 + 512 + 1024;
  KB_TRY_ATTR(sim_tc_kernel, smem);
  SimParams p;
  p.N = N; p.D = D; p.P = P; p.BN = BN; p.stages = stages; p.group = group; p.temp = temp;
  p.logits = logits; p.probs = probs; p.score_part = score_part;
  p.idesc = make_idesc(kFmtTF32, SBM, BN);
  *fused_probs = probs != nullptr && (16 % group) == 0;
  if (logits == nullptr && probs != nullptr && !*fused_probs)
    return set_error(KB_ERR_ARG, "similarity_tc: probabilities without logits need a group size dividing 16 (got %d)", group);
  if (probs != nullptr && !*fused_probs) p.probs = nullptr;
  const long long tiles = ((N + SBM - 1) / SBM) * ((P + BN - 1) / BN);
  int grid = num_sms();
  if (tiles < grid) grid = (int)tiles;
  sim_tc_kernel<<<grid, kSimThreads, smem, stream>>>(ta, tb, p);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
