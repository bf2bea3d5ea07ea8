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
//   gemm_kernel<BN,EPI>   one CTA per tile, UMMA 128 x BN x 16 (cta_group::1), BN = 128 — small problems
//   gemm2_kernel<EPI>     a CTA PAIR (cluster of 2) per 256x256 tile, UMMA 256 x 256 x 16 (cta_group::2):
//                         each CTA stages its 128 rows of A and its 128 rows of W, so per-SM operand traffic
//                         (L2->SMEM and SMEM->tensor core) drops by a third against the 128x256 single-CTA tile
//
// These kernels serve every dense layer on the path: ViT patch-embed, qkv, proj, fc1, fc2, the visual_head,
// and the BERT q|k|v, attention-output, intermediate, output and pooler projections (reference call sites:
// quick_start/keep_inference.py:32-46,49-50; SURVEY.md §2.3 K1,K3,K5-K8,K10,K11).
#include "common.h"
#include "ptx.cuh"


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
  // split-operand mode (common.h GemmArgs::split): the K loop runs `nseg` passes over the K/64 k-blocks; pass s reads the
  // A tile at column offset a_off[s] and the W tile at w_off[s] (0 = hi half, K = lo half of a [rows, 2K] hi|lo operand)
  int nseg;
  int a_off[3], w_off[3];
  long long lo_off;  // EPI_BIAS_GELU_HILO: element offset of the lo half inside an output row
};

// Exact-erf GELU (torch.nn.GELU() default), gelu(x) = x * Phi(x), written for the epilogue's instruction budget
// (the epilogue of fc1 is what bounds that GEMM: every instruction per element counts):
//   Phi(x) = 1 - q(|x|) for x >= 0 and q(|x|) for x < 0, q(a) = 0.5 * erfc(a / sqrt(2)), hence
//   gelu(x) = relu(x) - |x| * q(|x|),   q(a) = exp2(P(a)) on [0, 6]  (q(6) = 1e-9: clamped beyond).
// P is a weighted minimax fit of log2(0.5 erfc(a/sqrt2)), the weight being the error it causes in gelu
// (tools/fit_gelu.py). The result is stored as a 16-bit float, whose rounding is >= 2.4e-5 for |gelu| >= 0.05:
//   DEG 4: |gelu error| <= 6.6e-6 in fp32 evaluation,  7 FMA-pipe instructions + 1 MUFU   (16-bit outputs)
//   DEG 6: |gelu error| <= 3.3e-7,                      9 FMA-pipe instructions + 1 MUFU   (hi|lo outputs, ~22 bits)
// Evaluated for two elements at once on the packed fp32 pipe (FFMA2, sm_100): half the FMA-pipe instructions for the
// polynomial and the final multiply-add.
template <int DEG>
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
  const float2 m = make_float2(fminf(ax.x, 6.0f), fminf(ax.y, 6.0f));
  float2 p;
  if constexpr (DEG == 6) {
    p = __ffma2_rn(m, make_float2(2.904253473e-05f, 2.904253473e-05f), make_float2(-7.323236443e-04f, -7.323236443e-04f));
    p = __ffma2_rn(m, p, make_float2(7.953787372e-03f, 7.953787372e-03f));
    p = __ffma2_rn(m, p, make_float2(-5.320511315e-02f, -5.320511315e-02f));
    p = __ffma2_rn(m, p, make_float2(-4.589348205e-01f, -4.589348205e-01f));
    p = __ffma2_rn(m, p, make_float2(-1.151144948e+00f, -1.151144948e+00f));
    p = __ffma2_rn(m, p, make_float2(-9.999990962e-01f, -9.999990962e-01f));
  } else {
    p = __ffma2_rn(m, make_float2(3.920550193e-03f, 3.920550193e-03f), make_float2(-4.439129536e-02f, -4.439129536e-02f));
    p = __ffma2_rn(m, p, make_float2(-4.674139173e-01f, -4.674139173e-01f));
    p = __ffma2_rn(m, p, make_float2(-1.147820817e+00f, -1.147820817e+00f));
    p = __ffma2_rn(m, p, make_float2(-1.000374045e+00f, -1.000374045e+00f));
  }
  float2 e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(p.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(p.y));
  return __ffma2_rn(make_float2(-ax.x, -ax.y), e, make_float2(fmaxf(x.x, 0.0f), fmaxf(x.y, 0.0f)));
}

// Two independent pairs step by step (four elements of a row): the dependent FFMA2 chain of one pair fills the
// issue slots the other one waits in — the epilogue warps are only two per SM sub-partition.
template <int DEG>
__device__ __forceinline__ void gelu_erf2x2(float2& x0, float2& x1) {
  static_assert(DEG == 4, "interleaved form is written for the degree-4 polynomial");
  const float2 a0 = make_float2(fabsf(x0.x), fabsf(x0.y)), a1 = make_float2(fabsf(x1.x), fabsf(x1.y));
  const float2 m0 = make_float2(fminf(a0.x, 6.0f), fminf(a0.y, 6.0f)), m1 = make_float2(fminf(a1.x, 6.0f), fminf(a1.y, 6.0f));
  const float2 k4 = make_float2(3.920550193e-03f, 3.920550193e-03f), k3 = make_float2(-4.439129536e-02f, -4.439129536e-02f);
  const float2 k2 = make_float2(-4.674139173e-01f, -4.674139173e-01f), k1 = make_float2(-1.147820817e+00f, -1.147820817e+00f);
  const float2 k0 = make_float2(-1.000374045e+00f, -1.000374045e+00f);
  float2 p0 = __ffma2_rn(m0, k4, k3), p1 = __ffma2_rn(m1, k4, k3);
  p0 = __ffma2_rn(m0, p0, k2); p1 = __ffma2_rn(m1, p1, k2);
  p0 = __ffma2_rn(m0, p0, k1); p1 = __ffma2_rn(m1, p1, k1);
  p0 = __ffma2_rn(m0, p0, k0); p1 = __ffma2_rn(m1, p1, k0);
  float2 e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.x) : "f"(p0.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.x) : "f"(p1.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0.y) : "f"(p0.y));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1.y) : "f"(p1.y));
  x0 = __ffma2_rn(make_float2(-a0.x, -a0.y), e0, make_float2(fmaxf(x0.x, 0.0f), fmaxf(x0.y, 0.0f)));
  x1 = __ffma2_rn(make_float2(-a1.x, -a1.y), e1, make_float2(fmaxf(x1.x, 0.0f), fmaxf(x1.y, 0.0f)));
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float2 unpack16(uint32_t v, int bf16) {
  if (bf16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}

template <int EPI> struct EpiTraits {
  static constexpr bool kResid = (EPI == EPI_RESID_F32 || EPI == EPI_RESID_F32_STATS);
  static constexpr bool kStats = (EPI == EPI_RESID_F32_STATS);
  static constexpr bool kLn = (EPI == EPI_LN_BIAS_HALF || EPI == EPI_LN_BIAS_GELU_HALF);
  static constexpr bool kHiLo = (EPI == EPI_BIAS_GELU_HILO);
  static constexpr bool kGelu = (EPI == EPI_BIAS_GELU_HALF || EPI == EPI_LN_BIAS_GELU_HALF || kHiLo);
  static constexpr bool kHalfOut = (EPI == EPI_BIAS_HALF || EPI == EPI_BIAS_GELU_HALF || kLn || kHiLo);
};

// Drain one warp's share of an accumulator tile: TMEM lanes [32q, 32q+32) x columns [c_begin, c_end) of the
// accumulator at `tmem_acc`; rows row0.. of the output, tile column base n0. The warp waits for the accumulator
// (`tfull`, `parity`) itself, AFTER it has issued the global loads that do not depend on it (bias, LayerScale, the
// first residual block, the LayerNorm statistics of its rows), so their DRAM latency hides behind the wait.
//
// BF16 (16-bit output format) and FULL (all 32 rows of this warp lie inside M: no per-row guards, so the eight row groups
// of a block are one basic block the compiler can interleave) are compile-time: the instruction count of the epilogue is
// what paces the GEMMs with short K (profiles/r02_src_gemm notes in DESIGN.md), so nothing is re-decided per element.
template <int EPI, bool BF16, bool FULL_ROWS>
__device__ __forceinline__ void epilogue_warp(const KParams& p, uint32_t tmem_acc, int q, int lane, int row0, int n0,
                                              int c_begin, int c_end, uint8_t* stage, uint64_t* tfull, uint32_t parity,
                                              int tag) {
  using T = EpiTraits<EPI>;
  constexpr int kBf = BF16 ? 1 : 0;
  // the residual epilogues keep their row guards even on full tiles: without them ptxas hoists the address arithmetic of
  // all eight row groups of three arrays and spills (measured: fc2 1,209 -> 1,165 TFLOP/s in the step)
  constexpr bool FULL = FULL_ROWS && !T::kResid;
  const uint32_t stage_addr = smem_u32(stage);
  // transposed role of this lane: rows 4i + (lane >> 3), columns 4*(lane & 7) .. +3 of the 32x32 block
  const int tr = lane >> 3, tc = (lane & 7) * 4;
  const uint32_t t_lane = tmem_acc + (uint32_t(q * 32) << 16);
  int c_stop = c_end;
  if (n0 + c_stop > p.N) c_stop = p.N - n0;  // N is a multiple of 32 (warp-uniform)
  // element pointers of this lane's first row (row0 + tr) at column tc, and the distance between its row groups (4 rows):
  // a block adds its column, a row group i adds i * step, nothing else is recomputed per element
  // (the 16-bit-output epilogues only: the residual epilogues already hold two residual blocks in registers and
  // recompute their addresses from the row index instead of keeping 8 row pointers per array alive)
  const long long first = row0 + tr;
  uint16_t* out16_base = T::kHalfOut ? reinterpret_cast<uint16_t*>(p.out) + first * p.ldo + tc : nullptr;
  const long long out16_step = 4 * p.ldo;

  auto load_vec = [&](const float* base, int col, float fill) {
    return base != nullptr ? __ldg(reinterpret_cast<const float4*>(base + col + tc)) : make_float4(fill, fill, fill, fill);
  };
  auto load_res = [&](float4 (&res)[8], int col) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = row0 + 4 * i + tr;
      res[i] = (FULL || r < p.M) ? *reinterpret_cast<const float4*>(p.resid + (long long)r * p.ldr + col + tc)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  // EPI_LN_*: LayerNorm statistics of row (row0 + lane) of A, from the partial sums the producing GEMM left behind;
  // the transposed role fetches the pair of the row it is working on with two shuffles
  float ln_nmu = 0.f, ln_rstd = 0.f;
  if constexpr (T::kLn) {
    const int r = row0 + lane;
    if (FULL || r < p.M) {
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
        if constexpr (T::kHiLo) {
          const float2 lo = gelu_erf2<6>(make_float2(a.x, a.y)), hi = gelu_erf2<6>(make_float2(a.z, a.w));
          a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
        } else {
          float2 lo = make_float2(a.x, a.y), hi = make_float2(a.z, a.w);
          gelu_erf2x2<4>(lo, hi);
          a.x = lo.x; a.y = lo.y; a.z = hi.x; a.w = hi.y;
        }
        if (!FULL && r >= p.M) continue;
        uint2 w;
        w.x = pack16(a.x, a.y, kBf);
        w.y = pack16(a.z, a.w, kBf);
        uint16_t* orow = out16_base + (long long)i * out16_step + col;
        *reinterpret_cast<uint2*>(orow) = w;
        if constexpr (T::kHiLo) {  // lo = 16-bit(v - hi): hi + lo carries ~22 mantissa bits to the next split GEMM
          const float2 h0 = unpack16(w.x, kBf), h1 = unpack16(w.y, kBf);
          uint2 l;
          l.x = pack16(a.x - h0.x, a.y - h0.y, kBf);
          l.y = pack16(a.z - h1.x, a.w - h1.y, kBf);
          *reinterpret_cast<uint2*>(orow + p.lo_off) = l;
        }
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
        w.x = pack16(a[i].x, a[i].y, kBf);
        w.y = pack16(a[i].z, a[i].w, kBf);
        if (FULL || r < p.M) *reinterpret_cast<uint2*>(out16_base + (long long)i * out16_step + col) = w;
      } else if constexpr (EPI == EPI_PATCH_F32) {
        if (FULL || r < p.M) {
          const int img = r / p.patches, pi = r % p.patches;
          const float4 p4 = __ldg(reinterpret_cast<const float4*>(p.pos + (long long)(1 + pi) * p.N + col + tc));
          float4 o = a[i];
          o.x += p4.x; o.y += p4.y; o.z += p4.z; o.w += p4.w;
          const long long orow = (long long)img * (p.patches + 1) + 1 + pi;
          *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + orow * p.ldo + col + tc) = o;
        }
      } else {
        if (FULL || r < p.M) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (long long)r * p.ldo + col + tc) = a[i];
        if constexpr (T::kStats) {
          uint2 w;
          w.x = pack16(a[i].x, a[i].y, kBf);
          w.y = pack16(a[i].z, a[i].w, kBf);
          if (FULL || r < p.M) *reinterpret_cast<uint2*>(p.out16 + (long long)r * p.ldo16 + col + tc) = w;
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
      if (FULL || r < p.M)
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

template <int BN, int EPI, bool BF16, bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ KParams p) {
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
  const int seg_kb = p.K / BLOCK_K;  // k-blocks per pass over K
  const int num_kb = p.nseg * seg_kb;

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
        for (int sg = 0; sg < p.nseg; ++sg) {
          const int a0 = p.a_off[sg], w0 = p.w_off[sg];
          for (int kb = 0; kb < seg_kb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1, 1);
            mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
            tma_load_2d(&tmap_a, &full_bar[s], smem_a + s * C::A_BYTES, a0 + kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(&tmap_b, &full_bar[s], smem_b + s * C::B_BYTES, w0 + kb * BLOCK_K, n_blk * BN);
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
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
      epilogue_warp<EPI, BF16, FULL>(p, tmem_base + as * BN, q, lane, m_blk * BLOCK_M + q * 32, n_blk * BN, half * (BN / 2),
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
struct Cfg2 {
  static constexpr int BN = 256;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;    // this CTA's 128 rows of A: 16 KB
  static constexpr int B_BYTES = (BN / 2) * BLOCK_K * 2;   // this CTA's 128 rows of W: 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 32 KB per CTA per stage
  static constexpr int STAGES = 6;  // (4 and 5 stages measure the same on every layer shape: tools/bench_gemm_shapes.py)
  static constexpr int TMEM_COLS = 2 * BN;
  static constexpr int BAR_BYTES = 256;
  static constexpr int EPI_BYTES = kNumEpiWarps * kStageTileBytes;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES + 1024;
};

// Variants that were measured and dropped (clusters of 4 with the W tile multicast between two pairs; mixed 4+2 clusters
// with a dynamic scheduler; 16 epilogue warps) are described with their numbers in DESIGN.md section 8; their code is in
// the history (commit 87a3dc2), not in the product library.
template <int EPI, bool BF16, bool FULL>
__global__ void __launch_bounds__(kThreads, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
             const __grid_constant__ KParams p) {
  using C = Cfg2;
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
  const uint32_t rank = cluster_ctarank();       // rank in the pair: 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1;              // cluster index: the unit of the persistent schedule
  const int num_pairs = gridDim.x >> 1;
  const int m_tiles = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int seg_kb = p.K / BLOCK_K;              // k-blocks per pass over K

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
      mbar_init(&tempty_bar[i], 2 * kNumEpiWarps);
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
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        const int m0 = m_blk * 2 * BLOCK_M + rank * BLOCK_M;
        const int n0 = n_blk * BN + rank * (BN / 2);
        for (int sg = 0; sg < p.nseg; ++sg) {
          const int a0 = p.a_off[sg], w0 = p.w_off[sg];
          for (int kb = 0; kb < seg_kb; ++kb) {
            mbar_wait(&empty_bar[s], ph ^ 1, 11);
            const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[s]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * C::STAGE_BYTES);
            tma_load_2d_cg2(&tmap_a, leader_full, smem_a + s * C::A_BYTES, a0 + kb * BLOCK_K, m0);
            tma_load_2d_cg2(&tmap_b, leader_full, smem_b + s * C::B_BYTES, w0 + kb * BLOCK_K, n0);
            if (++s == C::STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      const int num_kb = p.nseg * seg_kb;
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
          umma_commit_cg2_mc(&empty_bar[s], (uint16_t)0b11);  // one arrival on the slot of both CTAs
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_cg2_mc(&tfull_bar[as], (uint16_t)0b11);   // accumulator halves complete in both CTAs of the pair
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue (each CTA drains its own 128 rows) =====================
    const int q = warp & 3;
    constexpr int PARTS = kNumEpiWarps / 4, PCOLS = BN / PARTS;  // column slices per lane quadrant
    const int part = (warp - kFirstEpiWarp) >> 2;
    uint8_t* stage = smem_epi + (warp - kFirstEpiWarp) * kStageTileBytes;
    int lt = 0;
    for (int tile = pair; tile < num_tiles; tile += num_pairs, ++lt) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      epilogue_warp<EPI, BF16, FULL>(p, tmem_base + as * BN, q, lane, m_blk * 2 * BLOCK_M + rank * BLOCK_M + q * 32,
                                     n_blk * BN, part * PCOLS, (part + 1) * PCOLS, stage, &tfull_bar[as], aph, 14);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_relaxed(mapa_shared(smem_u32(&tempty_bar[as]), 0));
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
  // split-operand passes: (Ah,Wh) | (Ah,Wh),(Ah,Wl) | (Ah,Wh),(Al,Wh),(Ah,Wl); the lo half sits K columns after the hi half
  p.nseg = a.split == GEMM_SPLIT_AW ? 3 : a.split == GEMM_SPLIT_W ? 2 : 1;
  for (int i = 0; i < 3; ++i) p.a_off[i] = p.w_off[i] = 0;
  if (a.split == GEMM_SPLIT_AW) { p.a_off[1] = a.K; p.w_off[2] = a.K; }
  if (a.split == GEMM_SPLIT_W) p.w_off[1] = a.K;
  p.lo_off = a.lo_off;
  return p;
}

template <int BN, int EPI, bool BF16, bool FULL>
int launch_one(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using C = Cfg<BN>;
  KB_TRY_ATTR((gemm_kernel<BN, EPI, BF16, FULL>), C::SMEM_BYTES);
  const KParams p = make_params(a, BLOCK_M, BN);
  const int tiles = ((a.M + BLOCK_M - 1) / BLOCK_M) * ((a.N + BN - 1) / BN);
  int grid = num_sms();
  if (tiles < grid) grid = tiles;
  profile_gemm_tag(a.M, a.N, a.K, a.epi);
  profile_gemm_begin(stream);
  gemm_kernel<BN, EPI, BF16, FULL><<<grid, kThreads, C::SMEM_BYTES, stream>>>(ta, tb, p);
  profile_gemm_end(stream, 2.0 * a.M * a.N * a.K);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

template <int EPI, bool BF16, bool FULL>
int launch_pair(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using C = Cfg2;
  KB_TRY_ATTR((gemm2_kernel<EPI, BF16, FULL>), C::SMEM_BYTES);
  const KParams p = make_params(a, 2 * BLOCK_M, 256);
  const int tiles = ((a.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M)) * ((a.N + 256 - 1) / 256);
  int clusters = num_sms() / 2;  // one CTA per SM (227 KB of shared memory each); 148 SMs = 74 pairs, all co-resident
  if (tiles < clusters) clusters = tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * clusters));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  profile_gemm_tag(a.M, a.N, a.K, a.epi);
  profile_gemm_begin(stream);
  KB_CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm2_kernel<EPI, BF16, FULL>, ta, tb, p));
  profile_gemm_end(stream, 2.0 * a.M * a.N * a.K);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

// One kernel per (epilogue, 16-bit format, full / ragged rows): the same problem always runs the same kernel. FULL = every
// row tile of the problem is complete (M a multiple of the tile height), so the epilogue carries no per-row guards.
#define KB_DISPATCH_EPI(FN, ...)                                                                       \
  switch (a.epi) {                                                                                     \
    case EPI_BIAS_HALF: return FN<__VA_ARGS__ EPI_BIAS_HALF, BF16, FULL>(a, ta, tb, stream);           \
    case EPI_BIAS_GELU_HALF: return FN<__VA_ARGS__ EPI_BIAS_GELU_HALF, BF16, FULL>(a, ta, tb, stream); \
    case EPI_RESID_F32: return FN<__VA_ARGS__ EPI_RESID_F32, BF16, FULL>(a, ta, tb, stream);           \
    case EPI_BIAS_F32: return FN<__VA_ARGS__ EPI_BIAS_F32, BF16, FULL>(a, ta, tb, stream);             \
    case EPI_PATCH_F32: return FN<__VA_ARGS__ EPI_PATCH_F32, BF16, FULL>(a, ta, tb, stream);           \
    case EPI_RESID_F32_STATS: return FN<__VA_ARGS__ EPI_RESID_F32_STATS, BF16, FULL>(a, ta, tb, stream); \
    case EPI_LN_BIAS_HALF: return FN<__VA_ARGS__ EPI_LN_BIAS_HALF, BF16, FULL>(a, ta, tb, stream);     \
    case EPI_LN_BIAS_GELU_HALF: return FN<__VA_ARGS__ EPI_LN_BIAS_GELU_HALF, BF16, FULL>(a, ta, tb, stream); \
    case EPI_BIAS_GELU_HILO: return FN<__VA_ARGS__ EPI_BIAS_GELU_HILO, BF16, FULL>(a, ta, tb, stream); \
    default: return set_error(KB_ERR_ARG, "gemm: unknown epilogue %d", a.epi);                         \
  }

template <bool BF16, bool FULL>
int dispatch_128(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  KB_DISPATCH_EPI(launch_one, 128, )
}
template <bool BF16, bool FULL>
int dispatch_pair(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  KB_DISPATCH_EPI(launch_pair, )
}
template <bool PAIR>
int dispatch_variant(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  const bool full = a.M % (PAIR ? 2 * BLOCK_M : BLOCK_M) == 0;
  if constexpr (PAIR) {
    if (a.bf16) return full ? dispatch_pair<true, true>(a, ta, tb, stream) : dispatch_pair<true, false>(a, ta, tb, stream);
    return full ? dispatch_pair<false, true>(a, ta, tb, stream) : dispatch_pair<false, false>(a, ta, tb, stream);
  } else {
    if (a.bf16) return full ? dispatch_128<true, true>(a, ta, tb, stream) : dispatch_128<true, false>(a, ta, tb, stream);
    return full ? dispatch_128<false, true>(a, ta, tb, stream) : dispatch_128<false, false>(a, ta, tb, stream);
  }
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
  if (a.split != GEMM_SPLIT_NONE && a.split != GEMM_SPLIT_W && a.split != GEMM_SPLIT_AW)
    return set_error(KB_ERR_ARG, "gemm: unknown split mode %d", a.split);
  if (a.epi == EPI_BIAS_GELU_HILO && a.lo_off <= 0) return set_error(KB_ERR_ARG, "gemm: hi|lo epilogue without lo_off");
  // Variant choice (a function of the shape only: the same problem always runs the same kernel): CTA-pair 256x256
  // tiles when they give every pair of SMs a tile; otherwise single-CTA 128x128 tiles (four times as many work units
  // for the small problems: CLS-row tails, small batches, the text tower of a WSI prompt set).
  const long long m256 = (a.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const bool pair = (a.N % 256 == 0) && m256 * (a.N / 256) >= num_sms() / 2;
  const int dt = a.bf16 ? KB_BF16 : KB_F16;
  // a split operand is a [rows, 2K] hi|lo matrix (row pitch >= 2K): the maps span both halves
  const int64_t a_cols = a.split == GEMM_SPLIT_AW ? 2LL * a.K : a.K, w_cols = a.split != GEMM_SPLIT_NONE ? 2LL * a.K : a.K;
  if (a.lda < a_cols || a.ldw < w_cols) return set_error(KB_ERR_ARG, "gemm: row pitch smaller than the (split) operand width");
  CUtensorMap ta, tb;
  int rc = get_tmap_2d(a.A, dt, a.M, a_cols, a.lda, BLOCK_M, &ta);
  if (rc) return rc;
  rc = get_tmap_2d(a.W, dt, a.N, w_cols, a.ldw, 128, &tb);
  if (rc) return rc;
  return pair ? dispatch_variant<true>(a, ta, tb, stream) : dispatch_variant<false>(a, ta, tb, stream);
}

}  // namespace kb
