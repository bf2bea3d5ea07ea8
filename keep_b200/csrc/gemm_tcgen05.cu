// Persistent warp-specialised GEMMs for sm_100a:  out = epilogue(A[M,K] . W[N,K]^T)
//
//   * operands fp16 or bf16 (K-major: activations [M,K] row-major, torch Linear weights [N,K]), fp32
//     accumulation in TMEM via tcgen05.mma.kind::f16;
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages the operand tiles through an mbarrier ring; one
//     producer thread and one MMA-issuing thread per CTA (pair);
//   * the accumulator is double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
//     main loop of tile i+1;
//   * 8 epilogue warps drain TMEM with tcgen05.ld, transpose 32x32 blocks through warp-private XOR-swizzled
//     shared memory so that every global access (fp32 residual read-modify-write, 16-bit stores) is a
//     row-contiguous 64/128-byte segment, and apply the fused epilogue (bias / exact-erf GELU /
//     LayerScale + fp32 residual / patch-embed scatter + pos_embed).
//
// Two main-loop variants:
//   gemm_kernel<BN,EPI>   one CTA per tile, UMMA 128 x BN x 16 (cta_group::1) — small problems
//   gemm2_kernel<EPI>     a CTA PAIR (cluster of 2) per 256x256 tile, UMMA 256 x 256 x 16 (cta_group::2):
//                         each CTA stages its 128 rows of A and its 128 rows of W, so per-SM operand traffic
//                         (L2->SMEM and SMEM->tensor core) drops by a third against the 128x256 single-CTA tile
//
// These kernels serve every dense layer on the path: ViT patch-embed, qkv, proj, fc1, fc2, the visual_head,
// and the BERT q|k|v, attention-output, intermediate, output and pooler projections (reference call sites:
// quick_start/keep_inference.py:32-46,49-50; SURVEY.md §2.3 K1,K3,K5-K8,K10,K11).
#include "common.h"
#include "ptx.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

namespace kb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = (kFirstEpiWarp + kNumEpiWarps) * 32;  // 384
constexpr int kStageTileBytes = 32 * 32 * 4;                   // per-warp 32x32 fp32 transpose tile
constexpr int kEpiSmemBytes = kNumEpiWarps * kStageTileBytes;  // 32 KB

struct KParams {
  int M, N, K;
  const float* bias;
  const float* gamma;
  const float* resid;
  long long ldr;
  void* out;
  long long ldo;
  const float* pos;
  int patches;
  uint32_t idesc;
  int bf16;
  uint16_t* out16;       // EPI_RESID_F32_STATS: 16-bit copy of the output
  long long ldo16;
  float* stats_out;      // EPI_RESID_F32_STATS: [M, N/64] (sum, sum of squares)
  const float* ln_stats; // EPI_LN_*: [M, ln_slices] (sum, sum of squares) of the A rows
  const float* ln_s;     // EPI_LN_*: [N] column sums of the folded weight
  int ln_slices;
  float ln_inv_width, ln_eps;
  int prefetch;  // residual epilogues: L2-prefetch the next tile's residual block (only pays when a tile is short)
};

// Exact-erf GELU (torch.nn.GELU() default), gelu(x) = x * Phi(x), written for the epilogue's instruction budget
// (the epilogue of fc1 is what bounds that GEMM: every instruction per element counts):
//   Phi(x) = 1 - q(|x|) for x >= 0 and q(|x|) for x < 0, q(a) = 0.5 * erfc(a / sqrt(2)), hence
//   gelu(x) = relu(x) - |x| * q(|x|),   q(a) = exp2(P(a)) on [0, 6]  (q(6) = 1e-9: clamped beyond).
// P is a weighted minimax fit of log2(0.5 erfc(a/sqrt2)), the weight being the error it causes in gelu
// (tools/fit_gelu.py). The result is stored as a 16-bit float, whose rounding is >= 2.4e-5 for |gelu| >= 0.05:
//   KB_GELU_DEG 4 (default): |gelu error| <= 6.6e-6 in fp32 evaluation,  7 FMA-pipe instructions + 1 MUFU
//   KB_GELU_DEG 6          : |gelu error| <= 3.3e-7,                      9 FMA-pipe instructions + 1 MUFU
#ifndef KB_GELU_DEG
#define KB_GELU_DEG 4
#endif
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float m = fminf(ax, 6.0f);
#if KB_GELU_DEG == 6
  float p = fmaf(m, 2.904253473e-05f, -7.323236443e-04f);
  p = fmaf(m, p, 7.953787372e-03f);
  p = fmaf(m, p, -5.320511315e-02f);
  p = fmaf(m, p, -4.589348205e-01f);
  p = fmaf(m, p, -1.151144948e+00f);
  p = fmaf(m, p, -9.999990962e-01f);
#else
  float p = fmaf(m, 3.920550193e-03f, -4.439129536e-02f);
  p = fmaf(m, p, -4.674139173e-01f);
  p = fmaf(m, p, -1.147820817e+00f);
  p = fmaf(m, p, -1.000374045e+00f);
#endif
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));
  return fmaf(-ax, e, fmaxf(x, 0.0f));
}

// Two elements at once on the packed fp32 pipe (FFMA2, sm_100): same arithmetic, same rounding per element, half the
// FMA-pipe instructions for the polynomial and the final multiply-add.
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 m = make_float2(fminf(ax.x, 6.0f), fminf(ax.y, 6.0f));
#if KB_GELU_DEG == 6
  float2 p = __ffma2_rn(m, make_float2(2.904253473e-05f, 2.904253473e-05f), make_float2(-7.323236443e-04f, -7.323236443e-04f));
  p = __ffma2_rn(m, p, make_float2(7.953787372e-03f, 7.953787372e-03f));
  p = __ffma2_rn(m, p, make_float2(-5.320511315e-02f, -5.320511315e-02f));
  p = __ffma2_rn(m, p, make_float2(-4.589348205e-01f, -4.589348205e-01f));
  p = __ffma2_rn(m, p, make_float2(-1.151144948e+00f, -1.151144948e+00f));
  p = __ffma2_rn(m, p, make_float2(-9.999990962e-01f, -9.999990962e-01f));
#else
  float2 p = __ffma2_rn(m, make_float2(3.920550193e-03f, 3.920550193e-03f), make_float2(-4.439129536e-02f, -4.439129536e-02f));
  p = __ffma2_rn(m, p, make_float2(-4.674139173e-01f, -4.674139173e-01f));
  p = __ffma2_rn(m, p, make_float2(-1.147820817e+00f, -1.147820817e+00f));
  p = __ffma2_rn(m, p, make_float2(-1.000374045e+00f, -1.000374045e+00f));
#endif
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(p.y));
  return __ffma2_rn(make_float2(-ax.x, -ax.y), e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int EPI> struct EpiTraits {
  static constexpr bool kResid = (EPI == EPI_RESID_F32 || EPI == EPI_RESID_F32_STATS);
  static constexpr bool kStats = (EPI == EPI_RESID_F32_STATS);
  static constexpr bool kLn = (EPI == EPI_LN_BIAS_HALF || EPI == EPI_LN_BIAS_GELU_HALF);
  static constexpr bool kGelu = (EPI == EPI_BIAS_GELU_HALF || EPI == EPI_LN_BIAS_GELU_HALF);
  static constexpr bool kHalfOut = (EPI == EPI_BIAS_HALF || EPI == EPI_BIAS_GELU_HALF || kLn);
};

// Drain one warp's share of an accumulator tile: TMEM lanes [32q, 32q+32) x columns [c_begin, c_end) of the
// accumulator at `tmem_acc`; rows row0.. of the output, tile column base n0. The warp waits for the accumulator
// (`tfull`, `parity`) itself, AFTER it has issued the global loads that do not depend on it (bias, LayerScale, the
// first residual block, the LayerNorm statistics of its rows), so their DRAM latency hides behind the wait.
template <int EPI>
__device__ __forceinline__ void epilogue_warp(const KParams& p, uint32_t tmem_acc, int q, int lane, int row0, int n0,
                                              int c_begin, int c_end, uint8_t* stage, uint64_t* tfull, uint32_t parity,
                                              int tag) {
  using T = EpiTraits<EPI>;
  const uint32_t stage_addr = smem_u32(stage);
  // transposed role of this lane: rows 4i + (lane >> 3), columns 4*(lane & 7) .. +3 of the 32x32 block
  const int tr = lane >> 3, tc = (lane & 7) * 4;
  const uint32_t t_lane = tmem_acc + (uint32_t(q * 32) << 16);
  int c_stop = c_end;
  if (n0 + c_stop > p.N) c_stop = p.N - n0;  // N is a multiple of 32 (warp-uniform)

  auto load_vec = [&](const float* base, int col, float fill) {
    return base != nullptr ? __ldg(reinterpret_cast<const float4*>(base + col + tc)) : make_float4(fill, fill, fill, fill);
  };
  auto load_res = [&](float4 (&res)[8], int col) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + 4 * i + tr;
      res[i] = (r < p.M) ? *reinterpret_cast<const float4*>(p.resid + (long long)r * p.ldr + col + tc)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  // EPI_LN_*: LayerNorm statistics of row (row0 + lane) of A, from the partial sums the producing GEMM left behind;
  // the transposed role fetches the pair of the row it is working on with two shuffles
  float ln_nmu = 0.f, ln_rstd = 0.f;
  if constexpr (T::kLn) {
    const int r = row0 + lane;
    if (r < p.M) {
      const float4* st = reinterpret_cast<const float4*>(p.ln_stats + (long long)r * p.ln_slices * 2);
      float sum = 0.f, sq = 0.f;
      for (int j = 0; j < p.ln_slices / 2; ++j) {  // fixed order: deterministic
        const float4 t = st[j];
        sum += t.x + t.z;
        sq += t.y + t.w;
      }
      const float mean = sum * p.ln_inv_width;
      const float var = fmaxf(fmaf(sq, p.ln_inv_width, -mean * mean), 0.f);
      ln_rstd = 1.0f / sqrtf(var + p.ln_eps);
      ln_nmu = -mean;
    }
  }
  // EPI_RESID_F32_STATS: per-lane partial sums of the new residual rows over the current 64-column slice
  float st_s[8], st_q[8];

  // One 32x32 block: accumulators of this lane's row in v[] -> warp-private staging tile -> row-contiguous role.
  // g4 = LayerScale (residual epilogues) or the folded column sums ln_s (EPI_LN_*).
  auto process = [&](const uint32_t (&v)[32], const float4 (&res)[8], int col, float4 b4, float4 g4) {
    // own row `lane` -> staging, 16-byte chunk j at (j ^ (lane & 7)): conflict-free for both access patterns
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t a = stage_addr + lane * 128 + ((j ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v[4 * j]), "r"(v[4 * j + 1]),
                   "r"(v[4 * j + 2]), "r"(v[4 * j + 3])
                   : "memory");
    }
    __syncwarp();
    if constexpr (T::kGelu) {
      // row group by row group (load, GELU, store): measured 3% faster for this epilogue than the straight-line
      // form below (A/B on one B200: fc1 1013 vs 983 TFLOP/s in the step), the stores drain under the next group's math
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int rl = 4 * i + tr;
        const int r = row0 + rl;
        float4 a;
        const uint32_t sa = stage_addr + rl * 128 + ((((lane & 7)) ^ (rl & 7)) << 4);
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "r"(sa));
        if constexpr (T::kLn) {
          const float nm = __shfl_sync(0xffffffffu, ln_nmu, rl), rs = __shfl_sync(0xffffffffu, ln_rstd, rl);
          const float2 nm2 = make_float2(nm, nm), rs2 = make_float2(rs, rs);  // packed fp32 (FFMA2): two columns per instruction
          const float2 lo = __ffma2_rn(rs2, __ffma2_rn(nm2, make_float2(g4.x, g4.y), make_float2(a.x, a.y)), make_float2(b4.x, b4.y));
          const float2 hi = __ffma2_rn(rs2, __ffma2_rn(nm2, make_float2(g4.z, g4.w), make_float2(a.z, a.w)), make_float2(b4.z, b4.w));
          a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
        } else {
          const float2 one2 = make_float2(1.f, 1.f);  // a + b as one packed FFMA2 per two columns (exact: a * 1 + b)
          const float2 lo = __ffma2_rn(make_float2(a.x, a.y), one2, make_float2(b4.x, b4.y));
          const float2 hi = __ffma2_rn(make_float2(a.z, a.w), one2, make_float2(b4.z, b4.w));
          a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
        }
#ifdef KB_GELU_SCALAR
        a.x = gelu_erf(a.x); a.y = gelu_erf(a.y); a.z = gelu_erf(a.z); a.w = gelu_erf(a.w);
#else
        {
          const float2 lo = gelu_erf2(make_float2(a.x, a.y)), hi = gelu_erf2(make_float2(a.z, a.w));
          a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
        }
#endif
        if (r >= p.M) continue;
        uint2 w;
        w.x = pack16(a.x, a.y, p.bf16);
        w.y = pack16(a.z, a.w, p.bf16);
        *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + (long long)r * p.ldo + col + tc) = w;
      }
      __syncwarp();  // staging tile is rewritten by the next block
    } else {
    // straight-line form: the arithmetic of all 8 row groups is one basic block, the guarded stores follow
    float4 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int rl = 4 * i + tr;
      const uint32_t sa = stage_addr + rl * 128 + ((((lane & 7)) ^ (rl & 7)) << 4);
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a[i].x), "=f"(a[i].y), "=f"(a[i].z), "=f"(a[i].w) : "r"(sa));
    }
    __syncwarp();  // staging tile may be rewritten by the next block
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if constexpr (T::kLn) {
        const float nm = __shfl_sync(0xffffffffu, ln_nmu, 4 * i + tr), rs = __shfl_sync(0xffffffffu, ln_rstd, 4 * i + tr);
        const float2 nm2 = make_float2(nm, nm), rs2 = make_float2(rs, rs);  // packed fp32 (FFMA2): two columns per instruction
        const float2 lo = __ffma2_rn(rs2, __ffma2_rn(nm2, make_float2(g4.x, g4.y), make_float2(a[i].x, a[i].y)), make_float2(b4.x, b4.y));
        const float2 hi = __ffma2_rn(rs2, __ffma2_rn(nm2, make_float2(g4.z, g4.w), make_float2(a[i].z, a[i].w)), make_float2(b4.z, b4.w));
        a[i].x = lo.x; a[i].y = lo.y; a[i].z = hi.x; a[i].w = hi.y;
      } else if constexpr (T::kResid) {
        // resid + gamma * (acc + bias), two columns per packed fp32 instruction (FFMA2; a * 1 + b is an exact add)
        const float2 one2 = make_float2(1.f, 1.f);
        const float2 lo = __ffma2_rn(make_float2(g4.x, g4.y), __ffma2_rn(make_float2(a[i].x, a[i].y), one2, make_float2(b4.x, b4.y)),
                                     make_float2(res[i].x, res[i].y));
        const float2 hi = __ffma2_rn(make_float2(g4.z, g4.w), __ffma2_rn(make_float2(a[i].z, a[i].w), one2, make_float2(b4.z, b4.w)),
                                     make_float2(res[i].z, res[i].w));
        a[i].x = lo.x; a[i].y = lo.y; a[i].z = hi.x; a[i].w = hi.y;
      } else {
        a[i].x += b4.x; a[i].y += b4.y; a[i].z += b4.z; a[i].w += b4.w;
      }
      if constexpr (T::kStats) {
        st_s[i] += (a[i].x + a[i].y) + (a[i].z + a[i].w);
        st_q[i] = fmaf(a[i].x, a[i].x, fmaf(a[i].y, a[i].y, fmaf(a[i].z, a[i].z, fmaf(a[i].w, a[i].w, st_q[i]))));
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + 4 * i + tr;
      if constexpr (T::kHalfOut) {
        uint2 w;
        w.x = pack16(a[i].x, a[i].y, p.bf16);
        w.y = pack16(a[i].z, a[i].w, p.bf16);
        if (r < p.M) *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p.out) + (long long)r * p.ldo + col + tc) = w;
      } else if constexpr (EPI == EPI_PATCH_F32) {
        if (r < p.M) {
          const int img = r / p.patches, pi = r % p.patches;
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(1 + pi) * p.N + col + tc));
          float4 o = a[i];
          o.x += p4.x; o.y += p4.y; o.z += p4.z; o.w += p4.w;
          const long long orow = (long long)img * (p.patches + 1) + 1 + pi;
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col + tc) = o;
        }
      } else {
        if (r < p.M) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)r * p.ldo + col + tc) = a[i];
        if constexpr (T::kStats) {
          uint2 w;
          w.x = pack16(a[i].x, a[i].y, p.bf16);
          w.y = pack16(a[i].z, a[i].w, p.bf16);
          if (r < p.M) *reinterpret_cast<uint2*>(p.out16 + (long long)r * p.ldo16 + col + tc) = w;
        }
      }
    }
    }  // !kGelu
  };
  // EPI_RESID_F32_STATS: fold the per-lane partials of one 64-column slice across the 8 lanes that share a row
  // (exchange-and-add butterfly: 14 shuffles; lane j of a row group ends up with the totals of row 4j + tr) and
  // store them as stats[row, slice] = (sum, sum of squares)
  auto flush_stats = [&](int col64) {
    if constexpr (T::kStats) {
      const bool b4 = (lane & 4) != 0, b2 = (lane & 2) != 0, b1 = (lane & 1) != 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float ks = b4 ? st_s[k + 4] : st_s[k], ss = b4 ? st_s[k] : st_s[k + 4];
        const float kq = b4 ? st_q[k + 4] : st_q[k], sq = b4 ? st_q[k] : st_q[k + 4];
        st_s[k] = ks + __shfl_xor_sync(0xffffffffu, ss, 4);
        st_q[k] = kq + __shfl_xor_sync(0xffffffffu, sq, 4);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float ks = b2 ? st_s[k + 2] : st_s[k], ss = b2 ? st_s[k] : st_s[k + 2];
        const float kq = b2 ? st_q[k + 2] : st_q[k], sq = b2 ? st_q[k] : st_q[k + 2];
        st_s[k] = ks + __shfl_xor_sync(0xffffffffu, ss, 2);
        st_q[k] = kq + __shfl_xor_sync(0xffffffffu, sq, 2);
      }
      const float ks = b1 ? st_s[1] : st_s[0], ss = b1 ? st_s[0] : st_s[1];
      const float kq = b1 ? st_q[1] : st_q[0], sq = b1 ? st_q[0] : st_q[1];
      const float tot_s = ks + __shfl_xor_sync(0xffffffffu, ss, 1);
      const float tot_q = kq + __shfl_xor_sync(0xffffffffu, sq, 1);
      const int r = row0 + 4 * (lane & 7) + tr;
      if (r < p.M)
        *reinterpret_cast<float2*>(p.stats_out + ((long long)r * (p.N / kLnSliceCols) + col64 / kLnSliceCols) * 2) =
            make_float2(tot_s, tot_q);
    }
  };

  if constexpr (T::kResid) {
    // residual blocks are double-buffered: block c+1 is in flight (DRAM/L2 latency) while block c is transformed
    float4 ra[8], rb[8];
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba, ga = make_float4(1.f, 1.f, 1.f, 1.f), gb = ga;
    if (c_begin < c_stop) {
      load_res(ra, n0 + c_begin);
      ba = load_vec(p.bias, n0 + c_begin, 0.f);
      ga = load_vec(p.gamma, n0 + c_begin, 1.f);
    }
    mbar_wait(tfull, parity, tag);
    tc_fence_after();
    uint32_t v[32];
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_stop; c0 += 64) {
      const bool has_b = c0 + 32 < c_stop;
      if constexpr (T::kStats) {
#pragma unroll
        for (int i = 0; i < 8; ++i) st_s[i] = st_q[i] = 0.f;
      }
      if (has_b) {
        load_res(rb, n0 + c0 + 32);
        bb = load_vec(p.bias, n0 + c0 + 32, 0.f);
        gb = load_vec(p.gamma, n0 + c0 + 32, 1.f);
      }
      tmem_ld_32x32(t_lane + uint32_t(c0), v);
      tmem_ld_wait_dep(v);
      process(v, ra, n0 + c0, ba, ga);
      if (has_b) {
        if (c0 + 64 < c_stop) {
          load_res(ra, n0 + c0 + 64);
          ba = load_vec(p.bias, n0 + c0 + 64, 0.f);
          ga = load_vec(p.gamma, n0 + c0 + 64, 1.f);
        }
        tmem_ld_32x32(t_lane + uint32_t(c0 + 32), v);
        tmem_ld_wait_dep(v);
        process(v, rb, n0 + c0 + 32, bb, gb);
      }
      flush_stats(n0 + c0);  // N % 64 == 0 is checked at launch for this epilogue
    }
  } else {
    // software-pipelined: the TMEM load of the next block is in flight while this one is transformed and stored
    const float4 none[8] = {};
    const float* gvec = T::kLn ? p.ln_s : nullptr;
    uint32_t va[32], vb[32];
    float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bb = ba;  // bias of the block in flight, fetched with it
    float4 ga = make_float4(1.f, 1.f, 1.f, 1.f), gb = ga;
    if (c_begin < c_stop) {
      ba = load_vec(p.bias, n0 + c_begin, 0.f);
      if constexpr (T::kLn) ga = load_vec(gvec, n0 + c_begin, 0.f);
    }
    mbar_wait(tfull, parity, tag);
    tc_fence_after();
    if (c_begin < c_stop) tmem_ld_32x32(t_lane + uint32_t(c_begin), va);
#pragma unroll 1
    for (int c0 = c_begin; c0 < c_stop; c0 += 64) {
      tmem_ld_wait_dep(va);
      if (c0 + 32 < c_stop) {
        tmem_ld_32x32(t_lane + uint32_t(c0 + 32), vb);
        bb = load_vec(p.bias, n0 + c0 + 32, 0.f);
        if constexpr (T::kLn) gb = load_vec(gvec, n0 + c0 + 32, 0.f);
      }
      process(va, none, n0 + c0, ba, ga);
      if (c0 + 32 < c_stop) {
        tmem_ld_wait_dep(vb);
        if (c0 + 64 < c_stop) {
          tmem_ld_32x32(t_lane + uint32_t(c0 + 64), va);
          ba = load_vec(p.bias, n0 + c0 + 64, 0.f);
          if constexpr (T::kLn) ga = load_vec(gvec, n0 + c0 + 64, 0.f);
        }
        process(vb, none, n0 + c0 + 32, bb, gb);
      }
    }
  }
}

// Residual epilogues stream 4 B/element from HBM; the read of tile i+1's residual block is started (into L2) while
// tile i is being drained, so that the epilogue's loads hit L2 instead of paying the DRAM latency per 32x32 block.
template <int EPI>
__device__ __forceinline__ void prefetch_residual(const KParams& p, int lane, int row0, int col0, int ncols) {
  if constexpr (EpiTraits<EPI>::kResid) {
    if (!p.prefetch) return;
    // this warp's block: 32 rows x ncols fp32 = ncols/32 lines of 128 B per row
    const int lines_per_row = ncols >> 5;
    for (int i = lane; i < 32 * lines_per_row; i += 32) {
      const int r = row0 + i / lines_per_row, c = col0 + (i % lines_per_row) * 32;
      if (r < p.M && c < p.N)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.resid + (long long)r * p.ldr + c));
    }
  }
}

// ============================================================================================================
// single-CTA tiles
// ============================================================================================================
template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BLOCK_K * 2;       // 32 KB @ BN=256
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // 512 / 256: power of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + kEpiSmemBytes + BAR_BYTES + 1024;  // +1024: alignment
};

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const KParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + kEpiSmemBytes);
  uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * C::STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 1);
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          tma_load_2d(&tmap_a, &full_bar[s], smem_a + s * C::A_BYTES, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(&tmap_b, &full_bar[s], smem_b + s * C::B_BYTES, kb * BLOCK_K, n_blk * BN);
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1, 2);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph, 3);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * C::A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * C::B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128-byte swizzle row: +2 in the >>4 address field
            umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);   // accumulator complete
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue =====================
    const int q = warp & 3;                         // TMEM lane quadrant this warp may access
    const int half = (warp - kFirstEpiWarp) >> 2;   // which half of the BN columns
    uint8_t* stage = smem_epi + (warp - kFirstEpiWarp) * kStageTileBytes;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      {
        const int nt = tile + gridDim.x;  // residual block of this warp in the CTA's next tile
        if (nt < num_tiles)
          prefetch_residual<EPI>(p, lane, (nt / n_tiles) * BLOCK_M + q * 32, (nt % n_tiles) * BN + half * (BN / 2), BN / 2);
      }
      epilogue_warp<EPI>(p, tmem_base + as * BN, q, lane, m_blk * BLOCK_M + q * 32, n_blk * BN, half * (BN / 2),
                         (half + 1) * (BN / 2), stage, &tfull_bar[as], aph, 4);
      // all TMEM reads of this accumulator are complete (wait::ld): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ============================================================================================================
// CTA-pair tiles (cta_group::2): 256 x 256 output tile per cluster of two CTAs
// ============================================================================================================
template <int EW>  // EW = number of epilogue warps
struct Cfg2 {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;    // this CTA's 128 rows of A: 16 KB
  static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;   // this CTA's 128 rows of W: 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB per CTA per stage
  static constexpr int STAGES = EW == 16 ? 5 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_BYTES = EW * kStageTileBytes;
  static constexpr int THREADS = (kFirstEpiWarp + EW) * 32;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
};

// CL = cluster size: 2 = one CTA pair per cluster; 4 = two pairs that work on vertically adjacent 256x256 tiles (same
// columns of W): every CTA then loads only a 64-row quarter of the W tile and TMA-multicasts it to the CTA holding the
// same half in the other pair, so L2 -> SMEM operand traffic per CTA and k-block drops from 32 KB to 24 KB.
template <int EPI, int EW, int CL>
__global__ void __launch_bounds__(Cfg2<EW>::THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const KParams p) {
  using C = Cfg2<EW>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_BYTES);
  uint64_t* full_bar = bars;                   // [STAGES] used in the leader CTA: bytes of BOTH CTAs land here
  uint64_t* empty_bar = bars + C::STAGES;      // [STAGES] per CTA, signalled by the leader's multicast commit
  uint64_t* tfull_bar = bars + 2 * C::STAGES;  // [2]      per CTA, multicast commit
  uint64_t* tempty_bar = tfull_bar + 2;        // [2]      used in the leader: epilogue warps of both CTAs arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int PAIRS = CL / 2;                  // pairs per cluster
  const uint32_t crank = cluster_ctarank();      // rank in the cluster
  const uint32_t rank = crank & 1;               // rank in the pair: 0 = leader (issues the MMAs)
  const uint32_t pr = crank >> 1;                // pair index inside the cluster
  const uint32_t leader = crank & ~1u;           // cluster rank of this pair's leader
  const int pair = blockIdx.x / CL;              // cluster index: the unit of the persistent schedule
  const int num_pairs = gridDim.x / CL;
  // a cluster owns PAIRS vertically adjacent 256-row tiles ("super-tile"); the pair `pr` takes the pr-th of them
  const int m_tiles = (p.M + PAIRS * 2 * BLOCK_M - 1) / (PAIRS * 2 * BLOCK_M);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], PAIRS);  // a slot is multicast-written by every pair: all their MMAs must have read it
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EW);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; bytes are credited to the leader's full barrier) ========
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs) {
        const int m_blk = (tile / n_tiles) * PAIRS + (int)pr, n_blk = tile % n_tiles;
        const int m0 = m_blk * 2 * BLOCK_M + rank * BLOCK_M;
        const int n0 = n_blk * BN + rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 11);
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), leader);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
          tma_load_2d_cg2(&tmap_a, leader_full, smem_a + s * C::A_BYTES, kb * BLOCK_K, m0);
          if constexpr (CL == 2) {
            tma_load_2d_cg2(&tmap_b, leader_full, smem_b + s * C::B_BYTES, kb * BLOCK_K, n0);
          } else {
            // this CTA's quarter of the W tile (rows n0 + pr*QB .. +QB) goes to the CTAs of every pair that hold the
            // same half (pair-rank `rank`): cluster ranks rank, rank + 2, ...; tmap_b has a QB-row box
            constexpr int QB = (BN / 2) / PAIRS;
            constexpr uint16_t kSameHalf = PAIRS == 2 ? 0b0101 : 0b01010101;
            tma_load_2d_cg2_mc(&tmap_b, smem_u32(&full_bar[s]) & kPeerBitMask,
                               smem_b + s * C::B_BYTES + pr * (QB * BLOCK_K * 2), kb * BLOCK_K, n0 + (int)pr * QB,
                               (uint16_t)(kSameHalf << rank));
          }
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1, 12);  // both CTAs' epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph, 13);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * C::A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * C::B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_f16_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_cg2_mc(&empty_bar[s], (uint16_t)((1u << CL) - 1));  // one arrival on the slot of every CTA of the cluster
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_cg2_mc(&tfull_bar[as], (uint16_t)(0b11u << leader));  // accumulator halves complete in both CTAs of the pair
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue (each CTA drains its own 128 rows) =====================
    const int q = warp & 3;
    constexpr int PARTS = EW / 4, PCOLS = BN / PARTS;  // column slices per lane quadrant
    const int part = (warp - kFirstEpiWarp) >> 2;
    uint8_t* stage = smem_epi + (warp - kFirstEpiWarp) * kStageTileBytes;
    int lt = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
      const int m_blk = (tile / n_tiles) * PAIRS + (int)pr, n_blk = tile % n_tiles;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      {
        const int nt = tile + num_pairs;
        if (nt < num_tiles)
          prefetch_residual<EPI>(p, lane, ((nt / n_tiles) * PAIRS + (int)pr) * 2 * BLOCK_M + rank * BLOCK_M + q * 32,
                                 (nt % n_tiles) * BN + part * PCOLS, PCOLS);
      }
      epilogue_warp<EPI>(p, tmem_base + as * BN, q, lane, m_blk * 2 * BLOCK_M + rank * BLOCK_M + q * 32, n_blk * BN,
                         part * PCOLS, (part + 1) * PCOLS, stage, &tfull_bar[as], aph, 14);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty_bar[as]), leader));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while its peer can still signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
  }
}

// ============================================================================================================
// CTA-pair tiles on MIXED clusters: launched with a regular cluster size of 2 and a PREFERRED size of 4
// (cudaLaunchAttributePreferredClusterDimension), so the device forms 4-CTA clusters where a GPC has room (33 on a
// B200 = 132 SMs) and 2-CTA clusters on the SMs left over (one TPC per 9-TPC GPC). A 4-CTA cluster shares the W tile
// between its two pairs by TMA multicast (as gemm2_kernel<.., 4>); a 2-CTA cluster is a plain pair. Because the mix is
// only known at run time the schedule is dynamic: warp 3 of every cluster's rank-0 CTA draws 512-row "super-tiles"
// from a global atomic counter and publishes them to all CTAs of its cluster through a small shared-memory ring
// (st.shared::cluster + remote mbarrier arrive); a 4-CTA cluster works on both 256-row halves at once, a pair does them
// one after the other.
// ============================================================================================================
constexpr int kSchedSlots = 4;

template <int EPI, int EW>
__global__ void __launch_bounds__(Cfg2<EW>::THREADS, 1)
gemm2d_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b128,
              const __grid_constant__ CUtensorMap tmap_b64, const KParams p, int* sched) {
  using C = Cfg2<EW>;
  constexpr int BN = C::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint8_t* smem_epi = smem + C::STAGES * C::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + C::EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::STAGES;
  uint64_t* tfull_bar = bars + 2 * C::STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* sfull_bar = tempty_bar + 2;               // [kSchedSlots] per CTA: the scheduler published a super-tile
  uint64_t* sempty_bar = sfull_bar + kSchedSlots;     // [kSchedSlots] in the rank-0 CTA: every consumer has read it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty_bar + kSchedSlots);
  volatile int* sched_tile = reinterpret_cast<volatile int*>(tmem_slot + 1);  // [kSchedSlots]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t csize = cluster_nctaid_x();     // 4 (preferred) or 2 (regular)
  const int PAIRS = (int)csize / 2;
  const int SUBS = 2 / PAIRS;                    // 256-row halves of a super-tile this pair does one after the other
  const uint32_t crank = cluster_ctarank();
  const uint32_t rank = crank & 1;
  const uint32_t pr = crank >> 1;
  const uint32_t leader = crank & ~1u;
  const int m_super = (p.M + 4 * BLOCK_M - 1) / (4 * BLOCK_M);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_st = m_super * n_tiles;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(PAIRS == 2 ? &tmap_b64 : &tmap_b128);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], PAIRS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EW);
    }
    for (int i = 0; i < kSchedSlots; ++i) {
      mbar_init(&sfull_bar[i], 1);
      // consumers of a published super-tile: per CTA the producer thread and EW epilogue warps, per pair the MMA thread
      mbar_init(&sempty_bar[i], csize * (1 + EW) + PAIRS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cg2(tmem_slot, C::TMEM_COLS);
    tmem_relinquish_cg2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // consumer side of the schedule ring: returns the next super-tile (>= num_st: no more work)
  int sslot = 0;
  uint32_t sph = 0;
  auto next_super_tile = [&](bool whole_warp) {
    mbar_wait_cluster(&sfull_bar[sslot], sph, 15);
    const int st = sched_tile[sslot];
    if (whole_warp) __syncwarp();
    // Hand the slot back with a RELAXED remote arrive: a release at cluster scope would first drain this warp's global
    // stores (the epilogue's output) - nothing is published here, the scheduler only needs to know the slot was read.
    // The target rank is computed from `st` (always 0) so that the arrive cannot issue before the read has returned.
    if (!whole_warp || lane == 0)
      mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&sempty_bar[sslot]), st < 0 ? 1u : 0u));
    if (++sslot == kSchedSlots) { sslot = 0; sph ^= 1; }
    return st;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int st = next_super_tile(false); st < num_st; st = next_super_tile(false)) {
        const int n_blk = st % n_tiles;
        const int n0 = n_blk * BN + rank * (BN / 2);
        for (int sub = 0; sub < SUBS; ++sub) {
          const int m_blk = (st / n_tiles) * 2 + (PAIRS == 2 ? (int)pr : sub);
          const int m0 = m_blk * 2 * BLOCK_M + rank * BLOCK_M;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1, 11);
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), leader);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
            tma_load_2d_cg2(&tmap_a, leader_full, smem_a + s * C::A_BYTES, kb * BLOCK_K, m0);
            if (PAIRS == 1) {
              tma_load_2d_cg2(&tmap_b128, leader_full, smem_b + s * C::B_BYTES, kb * BLOCK_K, n0);
            } else {
              constexpr int QB = BN / 4;  // 64-row quarter of the W tile, multicast to the same half of the other pair
              tma_load_2d_cg2_mc(&tmap_b64, smem_u32(&full_bar[s]) & kPeerBitMask,
                                 smem_b + s * C::B_BYTES + pr * (QB * BLOCK_K * 2), kb * BLOCK_K, n0 + (int)pr * QB,
                                 (uint16_t)(0b0101u << rank));
            }
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of each pair) =====================
    if (rank == 0 && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      const uint16_t all_ctas = (uint16_t)((1u << csize) - 1);
      for (int st = next_super_tile(false); st < num_st; st = next_super_tile(false)) {
        for (int sub = 0; sub < SUBS; ++sub, ++lt) {
          const int as = lt & 1;
          const uint32_t aph = (lt >> 1) & 1;
          mbar_wait(&tempty_bar[as], aph ^ 1, 12);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + as * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[s], ph, 13);
            tc_fence_after();
            const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * C::A_BYTES));
            const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * C::B_BYTES));
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_f16_ss_cg2(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2_mc(&empty_bar[s], all_ctas);
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
          umma_commit_cg2_mc(&tfull_bar[as], (uint16_t)(0b11u << leader));
        }
      }
    }
  } else if (warp == 3) {
    // ===================== scheduler (rank-0 CTA of the cluster) =====================
    if (crank == 0 && lane == 0) {
      int slot = 0;
      uint32_t ph = 0;
      while (true) {
        mbar_wait_cluster(&sempty_bar[slot], ph ^ 1, 16);
        // A pair needs two tile times per super-tile, a 4-CTA cluster one: near the end the pairs stop drawing, so that
        // the last super-tiles go to the clusters that finish them in half the time (tail of one tile, not two).
        int st;
        if (csize == 2 && *reinterpret_cast<volatile int*>(&sched[0]) >= num_st - 2 * (int)(gridDim.x / 4)) st = num_st;
        else st = atomicAdd(&sched[0], 1);
        for (uint32_t c = 0; c < csize; ++c) st_shared_cluster_u32(mapa_shared(smem_u32((const void*)&sched_tile[slot]), c), (uint32_t)st);
        for (uint32_t c = 0; c < csize; ++c) mbar_arrive_cluster(mapa_shared(smem_u32(&sfull_bar[slot]), c));  // release.cluster
        if (st >= num_st) break;
        if (++slot == kSchedSlots) { slot = 0; ph ^= 1; }
      }
      atomicAdd(&sched[csize == 4 ? 3 : 2], 1);  // statistics: clusters of each size seen so far (KEEPB200_VERBOSE)
      // the last cluster to run dry re-arms the counters for the next launch on this stream
      __threadfence();
      const int done = atomicAdd(&sched[1], (int)csize) + (int)csize;
      if (done == (int)gridDim.x) {
        sched[0] = 0;
        sched[1] = 0;
        __threadfence();
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    constexpr int PARTS = EW / 4, PCOLS = BN / PARTS;
    const int part = (warp - kFirstEpiWarp) >> 2;
    uint8_t* stage = smem_epi + (warp - kFirstEpiWarp) * kStageTileBytes;
    int lt = 0;
    for (int st = next_super_tile(true); st < num_st; st = next_super_tile(true)) {
      const int n_blk = st % n_tiles;
      for (int sub = 0; sub < SUBS; ++sub, ++lt) {
        const int m_blk = (st / n_tiles) * 2 + (PAIRS == 2 ? (int)pr : sub);
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        epilogue_warp<EPI>(p, tmem_base + as * BN, q, lane, m_blk * 2 * BLOCK_M + rank * BLOCK_M + q * 32, n_blk * BN,
                           part * PCOLS, (part + 1) * PCOLS, stage, &tfull_bar[as], aph, 14);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty_bar[as]), leader));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg2(tmem_base, C::TMEM_COLS);
  }
}

// ============================================================================================================
// host side
// ============================================================================================================
KParams make_params(const GemmArgs& a, int umma_m, int umma_n) {
  KParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.gamma = a.gamma; p.resid = a.resid; p.ldr = a.ldr;
  p.out = a.out; p.ldo = a.ldo; p.pos = a.pos; p.patches = a.patches;
  p.idesc = make_idesc(a.bf16 ? kFmtBF16 : kFmtF16, umma_m, umma_n);
  p.bf16 = a.bf16;
  p.out16 = static_cast<uint16_t*>(a.out16); p.ldo16 = a.ldo16; p.stats_out = a.stats_out;
  p.ln_stats = a.ln_stats; p.ln_s = a.ln_s; p.ln_slices = a.ln_slices;
  p.ln_inv_width = a.ln_width > 0 ? 1.0f / (float)a.ln_width : 0.f; p.ln_eps = a.ln_eps;
  // L2 prefetch of the next tile's residual block: off. With the first residual block requested before the accumulator
  // wait and the blocks double-buffered it no longer pays (A/B on one B200: proj 906 vs 887 TFLOP/s without it), and
  // a K=4096 tile streams ~4 MB per CTA pair through L2 first, so the prefetched lines were evicted again (ncu: +0.4 GB
  // of DRAM reads per fc2 launch). KEEPB200_RESID_PREFETCH=1 re-enables it for measurements.
  static int forced = -2;
  if (forced == -2) {
    const char* e = std::getenv("KEEPB200_RESID_PREFETCH");
    forced = e ? std::atoi(e) : -1;
  }
  p.prefetch = forced >= 0 ? forced : 0;
  return p;
}

template <int BN, int EPI>
int launch_one(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const KParams p = make_params(a, BLOCK_M, BN);
  const int tiles = ((a.M + BLOCK_M - 1) / BLOCK_M) * ((a.N + BN - 1) / BN);
  int grid = num_sms();
  if (tiles < grid) grid = tiles;
  profile_gemm_tag(a.M, a.N, a.K, a.epi);
  profile_gemm_begin(stream);
  gemm_kernel<BN, EPI><<<grid, kThreads, C::SMEM_BYTES, stream>>>(ta, tb, p);
  profile_gemm_end(stream, 2.0 * a.M * a.N * a.K);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

// CTAs per cluster of the pair kernel. 4 = the W tile is shared by two pairs through TMA multicast: +8-10 % throughput
// per SM, but only 33 clusters of 4 are co-resident on a B200 (132 of 148 SMs: one TPC per 9-TPC GPC is left over), so
// the whole GEMM is 1-3 % slower than with pairs (A/B in the step: 7,152 vs 7,230 tiles/s). Default 2;
// KEEPB200_GEMM_CLUSTER=4 selects the multicast variant (read per call: tests exercise both in one process).
int pair_cluster_size() {
  const char* e = std::getenv("KEEPB200_GEMM_CLUSTER");
  if (e && !std::strcmp(e, "mixed")) return 6;  // 4-CTA clusters where they fit + pairs on the rest, dynamic schedule
  return (e && std::atoi(e) == 4) ? 4 : 2;
}

template <int EPI, int EW, int CL>
int launch_pair_ew(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using C = Cfg2<EW>;
  static int max_clusters = 0;  // clusters of CL CTAs (one per SM, 227 KB of shared memory each) that can be co-resident
  if (max_clusters == 0) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(gemm2_kernel<EPI, EW, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    cudaLaunchConfig_t q = {};
    q.gridDim = dim3((unsigned)(num_sms() / CL * CL));
    q.blockDim = dim3(C::THREADS);
    q.dynamicSmemBytes = C::SMEM_BYTES;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    q.attrs = &at; q.numAttrs = 1;
    int n = 0;
    KB_CUDA_CHECK(cudaOccupancyMaxActiveClusters(&n, gemm2_kernel<EPI, EW, CL>, &q));
    if (n <= 0) return set_error(KB_ERR_CUDA, "gemm: no cluster of %d CTAs fits on this device", CL);
    max_clusters = n;
    if (std::getenv("KEEPB200_VERBOSE")) fprintf(stderr, "keep_b200: gemm2<epi %d> clusters of %d: %d co-resident (%d SMs)\n", EPI, CL, n, n * CL);
  }
  const KParams p = make_params(a, 2 * BLOCK_M, C::BN);
  constexpr int PAIRS = CL / 2;
  const int tiles = ((a.M + PAIRS * 2 * BLOCK_M - 1) / (PAIRS * 2 * BLOCK_M)) * ((a.N + C::BN - 1) / C::BN);
  int clusters = num_sms() / CL;
  if (max_clusters < clusters) clusters = max_clusters;
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(CL * clusters));
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  profile_gemm_tag(a.M, a.N, a.K, a.epi);
  profile_gemm_begin(stream);
  KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm2_kernel<EPI, EW, CL>, ta, tb, p));
  profile_gemm_end(stream, 2.0 * a.M * a.N * a.K);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

// one {next, done} counter pair per stream (device memory, zero at rest: the kernel re-arms it)
int* sched_counters(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<cudaStream_t, int*> bufs;
  std::lock_guard<std::mutex> g(mu);
  auto it = bufs.find(stream);
  if (it != bufs.end()) return it->second;
  int* p = nullptr;
  if (cudaMalloc(&p, 4 * sizeof(int)) != cudaSuccess) return nullptr;  // next, done, #pair clusters, #quad clusters
  cudaMemset(p, 0, 4 * sizeof(int));
  bufs[stream] = p;
  return p;
}

template <int EPI, int EW>
int launch_pair_dynamic(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb128, const CUtensorMap& tb64,
                        cudaStream_t stream) {
  using C = Cfg2<EW>;
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(gemm2d_kernel<EPI, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    KB_CUDA_CHECK(cudaFuncSetAttribute(gemm2d_kernel<EPI, EW>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr_set = true;
  }
  int* sched = sched_counters(stream);
  if (sched == nullptr) return set_error(KB_ERR_CUDA, "gemm: cannot allocate the tile-scheduler counters");
  const KParams p = make_params(a, 2 * BLOCK_M, C::BN);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(num_sms() / 4 * 4));  // every SM; a multiple of the preferred cluster size
  cfg.blockDim = dim3(C::THREADS);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributePreferredClusterDimension;
  at[1].val.preferredClusterDim.x = 4; at[1].val.preferredClusterDim.y = 1; at[1].val.preferredClusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  profile_gemm_tag(a.M, a.N, a.K, a.epi);
  profile_gemm_begin(stream);
  KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm2d_kernel<EPI, EW>, ta, tb128, tb64, p, sched));
  profile_gemm_end(stream, 2.0 * a.M * a.N * a.K);
  if (std::getenv("KEEPB200_VERBOSE")) {
    int h[4] = {0, 0, 0, 0};
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, sched, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "keep_b200: gemm2d<epi %d>: clusters so far: %d of 2 CTAs, %d of 4 CTAs\n", EPI, h[2], h[3]);
  }
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

template <int EPI>
int launch_pair(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  // 8 epilogue warps: a 16-warp variant (102 registers/thread) was measured no faster on fc1 and slower on proj
  return a.cluster == 4 ? launch_pair_ew<EPI, 8, 4>(a, ta, tb, stream) : launch_pair_ew<EPI, 8, 2>(a, ta, tb, stream);
}

#define KB_DISPATCH_EPI(FN, ...)                                                             \
  switch (a.epi) {                                                                           \
    case EPI_BIAS_HALF: return FN<__VA_ARGS__ EPI_BIAS_HALF>(a, ta, tb, stream);             \
    case EPI_BIAS_GELU_HALF: return FN<__VA_ARGS__ EPI_BIAS_GELU_HALF>(a, ta, tb, stream);   \
    case EPI_RESID_F32: return FN<__VA_ARGS__ EPI_RESID_F32>(a, ta, tb, stream);             \
    case EPI_BIAS_F32: return FN<__VA_ARGS__ EPI_BIAS_F32>(a, ta, tb, stream);               \
    case EPI_PATCH_F32: return FN<__VA_ARGS__ EPI_PATCH_F32>(a, ta, tb, stream);             \
    case EPI_RESID_F32_STATS: return FN<__VA_ARGS__ EPI_RESID_F32_STATS>(a, ta, tb, stream); \
    case EPI_LN_BIAS_HALF: return FN<__VA_ARGS__ EPI_LN_BIAS_HALF>(a, ta, tb, stream);       \
    case EPI_LN_BIAS_GELU_HALF: return FN<__VA_ARGS__ EPI_LN_BIAS_GELU_HALF>(a, ta, tb, stream); \
    default: return set_error(KB_ERR_ARG, "gemm: unknown epilogue %d", a.epi);               \
  }

int dispatch_256(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  KB_DISPATCH_EPI(launch_one, 256, )
}
int dispatch_128(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  KB_DISPATCH_EPI(launch_one, 128, )
}
int dispatch_pair(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  KB_DISPATCH_EPI(launch_pair, )
}
template <int EPI>
int launch_pair_dyn8(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  CUtensorMap tb64;
  int rc = get_tmap_2d(a.W, a.bf16 ? KB_BF16 : KB_F16, a.N, a.K, a.ldw, 64, &tb64);
  if (rc) return rc;
  return launch_pair_dynamic<EPI, 8>(a, ta, tb, tb64, stream);
}
int dispatch_pair_dynamic(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  KB_DISPATCH_EPI(launch_pair_dyn8, )
}

// KEEPB200_GEMM = auto (default) | pair | wide | narrow : force one main-loop variant (A/B measurements)
int forced_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = std::getenv("KEEPB200_GEMM");
    mode = 0;
    if (e && !std::strcmp(e, "pair")) mode = 1;
    if (e && !std::strcmp(e, "wide")) mode = 2;
    if (e && !std::strcmp(e, "narrow")) mode = 3;
  }
  return mode;
}

}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return set_error(KB_ERR_ARG, "gemm: empty problem %dx%dx%d", a.M, a.N, a.K);
  if (a.K % BLOCK_K != 0) return set_error(KB_ERR_ARG, "gemm: K=%d must be a multiple of %d", a.K, BLOCK_K);
  if (a.N % 32 != 0) return set_error(KB_ERR_ARG, "gemm: N=%d must be a multiple of 32", a.N);
  if ((a.epi == EPI_RESID_F32 || a.epi == EPI_RESID_F32_STATS) && a.resid == nullptr)
    return set_error(KB_ERR_ARG, "gemm: residual epilogue without resid");
  if (a.epi == EPI_RESID_F32_STATS && (a.out16 == nullptr || a.stats_out == nullptr || a.N % kLnSliceCols != 0))
    return set_error(KB_ERR_ARG, "gemm: stats epilogue needs out16, stats_out and N %% %d == 0 (N=%d)", kLnSliceCols, a.N);
  if ((a.epi == EPI_LN_BIAS_HALF || a.epi == EPI_LN_BIAS_GELU_HALF) &&
      (a.ln_stats == nullptr || a.ln_s == nullptr || a.ln_slices <= 0 || a.ln_slices % 2 != 0 || a.ln_width <= 0))
    return set_error(KB_ERR_ARG, "gemm: LayerNorm epilogue needs ln_stats, ln_s, an even ln_slices and ln_width");
  if (a.epi == EPI_PATCH_F32 && (a.pos == nullptr || a.patches <= 0))
    return set_error(KB_ERR_ARG, "gemm: patch epilogue without pos/patches");
  // Variant choice: CTA-pair 256x256 tiles when they fill the machine; otherwise 128-row tiles, 256 wide when
  // that still gives every SM a tile, else 128 wide (twice as many work units for small problems).
  const long long m128 = (a.M + BLOCK_M - 1) / BLOCK_M, m256 = (a.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const bool n256 = (a.N % 256 == 0);
  int mode = forced_mode();
  if (mode == 0) mode = (n256 && m256 * (a.N / 256) >= num_sms() / 2) ? 1 : (n256 && m128 * (a.N / 256) >= num_sms()) ? 2 : 3;
  if ((mode == 1 || mode == 2) && !n256) mode = 3;
  const int dt = a.bf16 ? KB_BF16 : KB_F16;
  CUtensorMap ta, tb;
  int rc = get_tmap_2d(a.A, dt, a.M, a.K, a.lda, BLOCK_M, &ta);
  if (rc) return rc;
  // pair kernel: clusters of 4 (W tile multicast between two pairs) when there are enough 512-row super-tiles
  GemmArgs a2 = a;
  const bool many = ((a.M + 511) / 512) * (a.N / 256) >= num_sms() / 4;
  a2.cluster = (mode == 1 && pair_cluster_size() == 4 && many) ? 4 : 2;
  if (mode == 1 && pair_cluster_size() == 6 && many) {
    rc = get_tmap_2d(a.W, dt, a.N, a.K, a.ldw, 128, &tb);
    if (rc) return rc;
    return dispatch_pair_dynamic(a2, ta, tb, stream);
  }
  rc = get_tmap_2d(a.W, dt, a.N, a.K, a.ldw, mode == 2 ? 256 : (mode == 1 && a2.cluster == 4) ? 64 : 128, &tb);
  if (rc) return rc;
  return mode == 1 ? dispatch_pair(a2, ta, tb, stream) : mode == 2 ? dispatch_256(a, ta, tb, stream) : dispatch_128(a, ta, tb, stream);
}

}  // namespace kb
