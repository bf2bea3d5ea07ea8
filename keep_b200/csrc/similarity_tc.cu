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

constexpr int SBM = 128;   // rows per CTA tile (TMEM lanes); fewer may be live (tile_rows)
constexpr int SBK = 32;    // fp32 elements per 128-byte swizzle row
constexpr int kSimEpiWarps = 8;
constexpr int kSimThreads = 512;  // warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 -, 4-7 norm, 8-15 epilogue (quadrant x column half)
constexpr int A_BYTES = SBM * 128;
constexpr int kOutTileBytes = 32 * 128;  // per epilogue warp and output: 32 rows x 32 fp32, SWIZZLE_128B

struct SimParams {
  long long N;
  int D, P, BN, stages, group;
  int tile_rows;      // feature rows per CTA tile (<= 128, multiple of 8): N is spread over every SM when it is small
  int a_slot;         // bytes between the feature slots of the ring (tile_rows * 128 rounded up to the 1 KB swizzle atom)
  float temp;
  float* logits;
  float* probs;
  float* score_part;  // prompt screening: [ceil(N/128)*4, P/group] partial sums of the top-2 margin term, or null
  int store_tma;      // outputs leave through per-warp 32 x 32 staging tiles and TMA stores (whole 128-byte row segments)
  uint32_t idesc;
};

// softmax(temp * f) over aligned groups of G consecutive entries of a 16-entry register block. The epilogue warps pace
// the wide shapes (ncu source page at 50k x 256: the MMA of a tile takes 7.2 us, the epilogue took 11), so the exponent is
// one FFMA + MUFU.EX2 (exp(x) = exp2(x * log2 e), x <= 0) and the normalisation one MUFU.RCP per group: ~2 ulp, against
// probabilities that are compared at 1e-3 (TF32 logits) - the fp32-exact path is the CUDA-core kernel of similarity.cu.
template <int G>
__device__ __forceinline__ void group_softmax16(const float (&f)[16], float (&e)[16], float temp) {
  const float tl = temp * 1.4426950408889634f;
#pragma unroll
  for (int b = 0; b < 16; b += G) {
    float t[G];
#pragma unroll
    for (int j = 0; j < G; ++j) t[j] = f[b + j] * tl;  // (any sign of temp)
    float mx = t[0];
#pragma unroll
    for (int j = 1; j < G; ++j) mx = fmaxf(mx, t[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < G; ++j) {
      const float d = t[j] - mx;
      asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[b + j]) : "f"(d));
      sum += e[b + j];
    }
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(sum));  // sum is in [1, G]
#pragma unroll
    for (int j = 0; j < G; ++j) e[b + j] *= inv;
  }
}

// prompt screening: top1 - top2 - |top1 + top2 - 1| of every group of G columns of a 16-column block (0 for dead rows / columns)
template <int G>
__device__ __forceinline__ void screen_terms(const uint32_t (&v)[16], float inv, bool row_live, int col, int P, float (&term)[8]) {
  constexpr int ng = 16 / G;
#pragma unroll
  for (int gi = 0; gi < 8; ++gi) {
    float t1 = -INFINITY, t2 = -INFINITY;
    if (gi < ng) {
#pragma unroll
      for (int i = 0; i < G; ++i) {
        const float x = __uint_as_float(v[(gi < ng ? gi : 0) * G + i]) * inv;
        t2 = fmaxf(t2, fminf(t1, x));
        t1 = fmaxf(t1, x);
      }
    }
    term[gi] = (gi < ng && row_live && col + gi * G < P) ? (t1 - t2) - fabsf(t1 + t2 - 1.0f) : 0.f;
  }
}

// PAIR = false: one CTA per tile of tile_rows x BN.
// PAIR = true : a CTA pair (cluster of 2, tcgen05 cta_group::2) per tile of 256 x BN: each CTA stages its own 128 feature
//   rows and HALF of the classifier block, so the classifier (re-read from L2 for every row tile: 786 KB per tile at
//   P = 256 against 393 KB of features) costs each SM half the L2 -> shared-memory traffic, and a stage is 32 KB instead
//   of 48 KB (6 stages instead of 4). The leader CTA issues the MMAs; TMA bytes of both CTAs are credited to the leader's
//   full barrier, so the norm warps (which read the feature blocks out of shared memory in BOTH CTAs) wait for the
//   multicast commit of the block's MMAs instead (mma_done): they read a block after the tensor core has, and only their
//   arrivals hand the slot back to the producer.
template <bool PAIR>
__global__ void __launch_bounds__(kSimThreads, 1)
sim_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
              const __grid_constant__ CUtensorMap tmap_l, const __grid_constant__ CUtensorMap tmap_p, const SimParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int bn_cta = PAIR ? p.BN / 2 : p.BN;     // classifier rows staged by this CTA
  const int b_bytes = bn_cta * 128;
  const int a_bytes = p.tile_rows * 128;         // bytes the TMA box of the features delivers (slot pitch stays 16 KB)
  const int stage_bytes = p.a_slot + b_bytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + p.stages * p.a_slot;
  uint8_t* smem_out = smem + p.stages * stage_bytes;  // [8 warps][2 outputs] 4 KB staging tiles (store_tma only)
  float* s_inv = reinterpret_cast<float*>(smem_out + (p.store_tma ? kSimEpiWarps * 2 * kOutTileBytes : 0));  // [2][128] 1/||row||
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_inv + 2 * SBM);
  uint64_t* full_bar = bars;                      // [stages] TMA -> MMA (PAIR: the leader's collects both CTAs' bytes)
  uint64_t* empty_bar = bars + p.stages;          // [stages] -> TMA: MMA commit + 4 norm warps (PAIR: the norm warps only)
  uint64_t* done_bar = bars + 2 * p.stages;       // [stages] PAIR: multicast commit of the block's MMAs -> norm warps
  uint64_t* tfull_bar = bars + 3 * p.stages;      // [2] MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] epilogue -> MMA (PAIR: in the leader, both CTAs' warps arrive)
  uint64_t* nready_bar = tempty_bar + 2;          // [2] norm warps -> epilogue
  uint64_t* sfree_bar = nready_bar + 2;           // [2] epilogue -> norm warps: s_inv[as] has been read
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // the persistent schedule's worker index
  const int num_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int tile_rows = PAIR ? 2 * SBM : p.tile_rows;
  const int m_tiles = (int)((p.N + tile_rows - 1) / tile_rows);
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
      mbar_init(&empty_bar[i], PAIR ? 4 : 1 + 4);
      mbar_init(&done_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], (PAIR ? 2 : 1) * kSimEpiWarps);
      mbar_init(&nready_bar[i], 4);
      mbar_init(&sfree_bar[i], kSimEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_cg2(tmem_slot, 512);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_slot, 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (separate rings for the feature and the classifier blocks - 8 x 16 KB of features in flight, a second producer
    // thread - were measured: no gain at 50k x 256, and 200k x 2 fell from 0.91 to 0.76 of the HBM peak because the MMA
    // thread then pays two barrier waits and two commits per 16 KB block)
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int m0 = m_blk * tile_rows + (int)rank * SBM;
        const int n0 = n_blk * p.BN + (int)rank * bn_cta;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 41);
          if constexpr (PAIR) {
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * (A_BYTES + b_bytes));
            tma_load_2d_cg2(&tmap_a, leader_full, smem_a + s * p.a_slot, kb * SBK, m0);
            tma_load_2d_cg2(&tmap_b, leader_full, smem_b + s * b_bytes, kb * SBK, n0);
          } else {
            mbar_arrive_expect_tx(&full_bar[s], a_bytes + b_bytes);
            tma_load_2d(&tmap_a, &full_bar[s], smem_a + s * p.a_slot, kb * SBK, m0);
            tma_load_2d(&tmap_b, &full_bar[s], smem_b + s * b_bytes, kb * SBK, n0);
          }
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (PAIR: leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units, ++lt) {
        const int as = lt & 1;
        mbar_wait(&tempty_bar[as], ((lt >> 1) & 1) ^ 1, 42);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 256;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph, 43);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * p.a_slot));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * b_bytes));
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // 8 tf32 elements = 32 bytes per MMA: +2 in the >>4 address field
            if constexpr (PAIR) umma_tf32_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_tf32_ss(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (PAIR) umma_commit_cg2_mc(&done_bar[s], (uint16_t)0b11);
          else umma_commit(&empty_bar[s]);
          if (++s == p.stages) { s = 0; ph ^= 1; }
        }
        if constexpr (PAIR) umma_commit_cg2_mc(&tfull_bar[as], (uint16_t)0b11);
        else umma_commit(&tfull_bar[as]);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== norm warps: sum of squares of each feature row, from the staged blocks ==========
    const int r = (warp - 4) * 32 + lane;  // row of the tile owned by this thread
    int s = 0;
    uint32_t ph = 0;
    int lt = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++lt) {
      const int as = lt & 1;
      float sq = 0.f;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(PAIR ? &done_bar[s] : &full_bar[s], ph, 44);
        const uint32_t row_addr = smem_u32(smem_a + s * p.a_slot) + r * 128;
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
      mbar_wait(&sfree_bar[as], ((lt >> 1) & 1) ^ 1, 45);  // the epilogue two tiles back has read s_inv[as]
      s_inv[as * SBM + r] = 1.0f / fmaxf(sqrtf(sq), 1e-12f);
      __syncwarp();
      if (lane == 0) mbar_arrive(&nready_bar[as]);
    }
  } else if (warp >= 8) {
    // ===================== epilogue: scale, logits, grouped softmax =====================
    const int q = warp & 3;
    const int half = (warp - 8) >> 2;  // which half of the BN columns (in 16-column blocks)
    const int r = q * 32 + lane;
    const bool fused_softmax = p.probs != nullptr && (16 % p.group) == 0;
    const bool tma_out = p.store_tma != 0 && p.score_part == nullptr;
    const uint32_t stage_l = smem_u32(smem_out + (warp - 8) * 2 * kOutTileBytes), stage_p = stage_l + kOutTileBytes;
    const int nblk = p.BN / 16;
    const int c_begin = (half == 0 ? 0 : (nblk + 1) / 2) * 16, c_end = (half == 0 ? (nblk + 1) / 2 : nblk) * 16;
    int lt = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++lt) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(&nready_bar[as], aph, 46);
      mbar_wait(&tfull_bar[as], aph, 47);
      tc_fence_after();
      const float inv = s_inv[as * SBM + r];
      const long long row = (long long)m_blk * tile_rows + (long long)rank * SBM + r;
      const bool row_live = r < (PAIR ? SBM : p.tile_rows) && row < p.N;
      const long long part_row = (PAIR ? (long long)m_blk * 2 + rank : (long long)m_blk) * 4 + q;  // screening: 128-row blocks
      const uint32_t t_row = tmem_base + as * 256 + (uint32_t(q * 32) << 16);
      // the TMEM load of the next 16-column block is in flight while this one is scaled, softmaxed and stored
      auto process = [&](const uint32_t (&v)[16], int c0, int hb) {
        const int col = n_blk * p.BN + c0;
        if (p.score_part != nullptr) {
          // prompt screening (WSI_evaluation/utils.py:107-117, rank_cls_score): per classifier (group of C columns) the
          // term top1 - top2 - |top1 + top2 - 1| of every tile, summed over the 32 rows of this warp; the [N, K*C] logits
          // never leave the SM. Whole warp participates (rows beyond N contribute 0); groups never straddle a block.
          if (col >= p.P) return;  // warp-uniform
          const int G = p.group, ng = 16 / G;
          float term[8];
          switch (G) {  // compile-time group size: the score block stays in registers
            case 2: screen_terms<2>(v, inv, row_live, col, p.P, term); break;
            case 4: screen_terms<4>(v, inv, row_live, col, p.P, term); break;
            case 8: screen_terms<8>(v, inv, row_live, col, p.P, term); break;
            default: screen_terms<16>(v, inv, row_live, col, p.P, term); break;
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
              p.score_part[part_row * (p.P / G) + (col / G + gi)] = t;
          }
          return;
        }
        if (col >= p.P) return;  // warp-uniform
        if (!tma_out && !row_live) return;
        float f[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) * inv;
        if (tma_out) {
          // this thread's 64 bytes of its row of the 32 x 32 staging tile (16-byte chunk ^ row: the SWIZZLE_128B pattern
          // the TMA store expects, and conflict-free); rows beyond N and columns beyond P are clipped by the store
          float e[16];
          if (fused_softmax) {
            switch (p.group) {
              case 1: group_softmax16<1>(f, e, p.temp); break;
              case 2: group_softmax16<2>(f, e, p.temp); break;
              case 4: group_softmax16<4>(f, e, p.temp); break;
              case 8: group_softmax16<8>(f, e, p.temp); break;
              default: group_softmax16<16>(f, e, p.temp); break;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            const uint32_t off = lane * 128 + (((hb * 4 + (i >> 2)) ^ (lane & 7)) << 4);
            if (p.logits != nullptr)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_l + off), "f"(f[i]), "f"(f[i + 1]), "f"(f[i + 2]), "f"(f[i + 3]) : "memory");
            if (fused_softmax)
              asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stage_p + off), "f"(e[i]), "f"(e[i + 1]), "f"(e[i + 2]), "f"(e[i + 3]) : "memory");
          }
          return;
        }
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
      };
      uint32_t va[16], vb[16];
      if (c_begin < c_end) tmem_ld_32x16(t_row + c_begin, va);
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        tmem_ld_wait_dep16(va);
        if (c0 + 16 < c_end) tmem_ld_32x16(t_row + c0 + 16, vb);
        if (tma_out) {  // the previous step's stores have read the staging tiles
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
        }
        process(va, c0, 0);
        if (c0 + 16 < c_end) {
          tmem_ld_wait_dep16(vb);
          if (c0 + 32 < c_end) tmem_ld_32x16(t_row + c0 + 32, va);
          process(vb, c0 + 16, 1);
        }
        if (tma_out && n_blk * p.BN + c0 < p.P) {
          fence_proxy_async_smem();  // every lane: its staging writes become visible to the TMA engine
          __syncwarp();
          if (lane == 0) {
            const int row_w = m_blk * tile_rows + (int)rank * SBM + q * 32, col_s = n_blk * p.BN + c0;
            if (p.logits != nullptr) tma_store_2d(&tmap_l, smem_out + (warp - 8) * 2 * kOutTileBytes, col_s, row_w);
            if (fused_softmax) tma_store_2d(&tmap_p, smem_out + ((warp - 8) * 2 + 1) * kOutTileBytes, col_s, row_w);
            tma_store_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&sfree_bar[as]);
        if constexpr (PAIR) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty_bar[as]), 0));
        else mbar_arrive(&tempty_bar[as]);
      }
    }
    if (tma_out && lane == 0) tma_store_wait_all();  // the staging tiles must outlive the last stores
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_cg2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
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
  // CTA pairs (256-row tiles, half of the classifier block per CTA) when the classifier block is wide enough for its
  // L2 -> shared-memory traffic to matter and there are row tiles for every pair; otherwise one CTA per tile, and when
  // the rows would not even give every SM a 128-row tile they are spread in smaller tiles (10k tiles x 32 prompts: 79
  // tiles of 128 rows leave 69 SMs idle, 139 tiles of 72 rows do not). The choice is a function of the shape only.
  const int sms = num_sms();
  const bool pair = BN >= 128 && BN % 32 == 0 && (N + 255) / 256 >= sms / 2;
  int tile_rows = SBM;
  if (!pair && score_part == nullptr) {
    const long long n_tiles = (P + BN - 1) / BN;
    if (((N + SBM - 1) / SBM) * n_tiles < sms) {
      const long long want = (N * n_tiles + sms - 1) / sms;  // rows per tile that give every SM one tile
      tile_rows = (int)((want + 7) / 8 * 8);
      if (tile_rows < 8) tile_rows = 8;
      if (tile_rows > SBM) tile_rows = SBM;
    }
  }
  const int b_bytes = (pair ? BN / 2 : BN) * 128;
  // bytes in flight per SM are what bounds the wide shapes (Little's law with ~3.5 us of loaded HBM latency: 200k x 2 keeps
  // 144 KB in flight and reaches 0.91 of the HBM peak, 50k x 256 has half of its ring taken by classifier blocks), so
  // every stage that fits is used: 7 x 32 KB for CTA pairs at P = 256
  // outputs through staging tiles + TMA stores (whole 128-byte row segments instead of 16 bytes per row and instruction)
  // when every epilogue warp owns whole 32-column steps and the rows of a warp are one contiguous block
  const bool store_tma = score_part == nullptr && BN % 64 == 0 && P % 4 == 0 && (pair || tile_rows == SBM) &&
                         (logits == nullptr || (reinterpret_cast<uintptr_t>(logits) & 15) == 0) &&
                         (probs == nullptr || (reinterpret_cast<uintptr_t>(probs) & 15) == 0);
  const int fixed = 2 * SBM * 4 + 512 + 1024 + (store_tma ? kSimEpiWarps * 2 * kOutTileBytes : 0);  // s_inv, barriers, slack, staging
  const int a_slot = pair ? A_BYTES : (tile_rows * 128 + 1023) / 1024 * 1024;
  int stages = (227 * 1024 - fixed) / (a_slot + b_bytes);
  if (stages > 16) stages = 16;  // 3 * stages + 8 mbarriers in the 512-byte barrier block
  if (stages < 2) return set_error(KB_ERR_ARG, "similarity_tc: tile too large");
  CUtensorMap ta, tb;
  int rc = get_tmap_2d(feats, KB_F32, N, D, D, pair ? SBM : tile_rows, &ta);
  if (rc) return rc;
  rc = get_tmap_2d(clsT, KB_F32, P, D, D, pair ? BN / 2 : BN, &tb);
  if (rc) return rc;
  const int smem = stages * (a_slot + b_bytes) + fixed;
  CUtensorMap tl = ta, tp = ta;  // output maps (32 x 32 fp32 boxes); placeholders when an output is absent
  if (store_tma && logits != nullptr) {
    rc = get_tmap_2d(logits, KB_F32, N, P, P, 32, &tl);
    if (rc) return rc;
  }
  SimParams p;
  p.N = N; p.D = D; p.P = P; p.BN = BN; p.stages = stages; p.group = group; p.temp = temp; p.tile_rows = tile_rows; p.a_slot = a_slot;
  p.logits = logits; p.probs = probs; p.score_part = score_part; p.store_tma = store_tma ? 1 : 0;
  p.idesc = make_idesc(kFmtTF32, pair ? 2 * SBM : SBM, BN);
  *fused_probs = probs != nullptr && (16 % group) == 0;
  if (logits == nullptr && probs != nullptr && !*fused_probs)
    return set_error(KB_ERR_ARG, "similarity_tc: probabilities without logits need a group size dividing 16 (got %d)", group);
  if (probs != nullptr && !*fused_probs) p.probs = nullptr;
  if (store_tma && p.probs != nullptr) {
    rc = get_tmap_2d(p.probs, KB_F32, N, P, P, 32, &tp);
    if (rc) return rc;
  }
  if (pair) {
    KB_TRY_ATTR(sim_tc_kernel<true>, smem);
    const long long tiles = ((N + 2 * SBM - 1) / (2 * SBM)) * ((P + BN - 1) / BN);
    int clusters = sms / 2;
    if (tiles < clusters) clusters = (int)tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(kSimThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, sim_tc_kernel<true>, ta, tb, tl, tp, p));
  } else {
    KB_TRY_ATTR(sim_tc_kernel<false>, smem);
    const long long tiles = ((N + tile_rows - 1) / tile_rows) * ((P + BN - 1) / BN);
    int grid = sms;
    if (tiles < grid) grid = (int)tiles;
    sim_tc_kernel<false><<<grid, kSimThreads, smem, stream>>>(ta, tb, tl, tp, p);
  }
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
