// tcgen05 attention for sequences of 65..224 tokens whose keys fit ONE tile (the ViT's 197 tokens), head dim 64:
//     out = softmax(Q K^T * scale + key_mask) V      per (batch, head)
// The general kernel (attention_tc.cu: KV blocks of 112 keys, packed short sequences, up to 512 tokens) was measured
// slower on exactly this shape (481 vs 266 us per ViT layer of 512 tiles: a unit then costs 21 MMA instructions and twice
// the barrier round trips instead of 17, and the issue rate of the MMA thread is what bounds both kernels,
// profiles/r02_attention_experiments.txt), so the single-tile form stays for it.
//
// One persistent CTA per SM; everything between the two MMAs stays on-chip (FlashAttention-4 style roles):
//   * unit of work = 128 query rows of one (batch, head). The keys of a head fit ONE tile (S_pad <= 224), so the
//     whole score row is in TMEM at once: one pass over S with a lazily raised power-of-two reference (below), no
//     rescaling of O;
//   * warp 13     : TMA producer. Q tile(s) + K go through a 2-slot ring, V through a 3-slot ring (Q/K are dead as
//                   soon as the S MMAs of the head are done, V only after its last PV MMA), all read straight out of
//                   the fused q|k|v projection buffer [B*S, 3*H*64] with SWIZZLE_128B boxes;
//   * warp 12     : MMA issuer. S = Q K^T (tcgen05.mma SS, 128 x S_pad x 16, fp32 S in TMEM region u&1), and
//                   O = P V (tcgen05.mma TS: P is read from TMEM where it overwrote S; V is the MN-major B operand);
//                   issue order S0 S1 | PV0 S2 | PV1 S3 | ... keeps the tensor pipe busy under the softmax;
//   * warps 4-7 / 8-11 : two softmax groups, one per S region, each working on its own unit; a thread owns one query
//                   row: tcgen05.ld S (one block ahead of its wait) -> exp2 against the running reference (packed fp32
//                   FFMA2 for the scale-and-shift and the row sums, MUFU for exp2) -> 16-bit P written back over S with
//                   tcgen05.st. What bounds the kernel is the chain S(u) -> softmax(u) -> PV(u) -> S(u+2) on a region
//                   (DESIGN.md section 4): the other group's unit fills the gaps;
//   * warps 0-3   : output group: tcgen05.ld O (single 64-column accumulator), divide by the row sum, store rows
//                   (optionally followed by their 16-bit rounding remainder: the hi|lo operand of a split-operand GEMM).
// The single-thread roles sit on the highest warp ids (the SM's warp arbiter favours them) and walk the units with
// counters instead of integer divisions: nothing hides a division's latency behind one thread.
//
// Reference semantics: timm Attention -> F.scaled_dot_product_attention and BertSelfAttention with
// the additive key mask (transformers modeling_bert.py:115-140).
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

constexpr int kAtcThreads = 512;
constexpr int kMaxSpadShared = 224;           // separate O accumulator: 2 * S_pad + 64 (O) <= 512 TMEM columns
constexpr int kMaxSpad = 256;                 // O inside the unit's own region (below): 2 * S_pad <= 512
constexpr int Q_TILE_BYTES = 128 * 128;       // 128 rows x 64 x 16-bit
constexpr int kQkSlots = 2;

struct Atc1Params {
  int B, S, H, S_pad, n_qt;
  int items;       // B * H
  const long long* key_mask;
  long long mask_stride;
  uint16_t* out;
  long long out_pitch, lo_off;
  float scale_log2;
  uint32_t idesc_s;   // 128 x S_pad, both operands K-major
  uint32_t idesc_pv;  // 128 x 64, B (V) MN-major
  int bf16;
  long long* trace;   // optional [64 units][16 events] clock64 stamps of CTA 0 (debug/profiling aid), or null
};

// Shared-memory layout (offsets inside the 1024-aligned dynamic block), sized by S_pad: Q/K slots of 2 x 16 KB + S_pad
// key rows, V slots of S_pad rows (3 of them up to 224 keys, 2 beyond: 256 keys take 64 + 32 KB per slot)
struct Smem {
  int qk_slot, v_slot, v_slots;
  int qk, v, bias, meta, rowsum, bars, total;
  __host__ __device__ explicit Smem(int S_pad) {
    qk_slot = 2 * Q_TILE_BYTES + S_pad * 128;   // multiple of 1024 (S_pad is a multiple of 16)
    v_slot = S_pad * 128;
    v_slots = S_pad > kMaxSpadShared ? 2 : 3;
    qk = 0;
    v = kQkSlots * qk_slot;
    bias = v + v_slots * v_slot;              // [v_slots][256] float: 0 / -inf per key
    meta = bias + v_slots * 256 * 4;          // [v_slots] int: index of the first masked key
    rowsum = meta + 64;                       // [2 regions][2 parities][128] float
    bars = rowsum + 2 * 2 * 128 * 4;
    total = bars + 256;
  }
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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

// event slots of the optional trace (per unit)
enum { EV_S_ISSUE = 0, EV_PV_WAITED = 1, EV_PV_ISSUED = 2, EV_SM_START = 3, EV_SM_P1 = 4, EV_SM_BATON = 5, EV_SM_P2 = 6,
       EV_OUT_START = 7, EV_OUT_DONE = 8 };
#define ATC_TRACE(u, ev)                                                                         \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && (u) < 64 && (threadIdx.x & 127) == 0)           \
      p.trace[(u) * 16 + (ev)] = clock64();                                                      \
  } while (0)

template <bool BF16>
__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tc1_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const Atc1Params p) {
  constexpr int kBf = BF16 ? 1 : 0;  // compile-time 16-bit format: one F2FP per pair in the softmax, no predicated twin
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Smem L(p.S_pad);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  int* s_meta = reinterpret_cast<int*>(smem + L.meta);
  float* s_rowsum = reinterpret_cast<float*>(smem + L.rowsum);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* qk_full = bars;        // [2] TMA -> MMA
  uint64_t* qk_empty = bars + 2;   // [2] MMA (last S of the head committed) -> TMA
  uint64_t* v_full = bars + 4;     // [3] TMA + key-bias writer -> MMA, softmax
  uint64_t* v_empty = bars + 7;    // [3] MMA (last PV of the head committed) -> TMA
  uint64_t* s_ready = bars + 10;   // [2] MMA -> softmax group
  uint64_t* p_ready = bars + 12;   // [2] softmax group -> MMA, output group
  uint64_t* o_ready = bars + 14;   // [1] MMA -> output group
  uint64_t* o_free = bars + 15;    // [1] output group -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // heads of this CTA
  const int U = n_my * p.n_qt;                                                          // units of this CTA
  // O accumulator (64 columns): up to 224 keys one shared accumulator after the two S regions; beyond (<= 256 keys: the
  // padded prompt bank) the regions fill TMEM and O(u) lives in the upper half of ITS OWN region, which is free once the
  // softmax has replaced the fp32 scores by the packed 16-bit P in the lower half. S(u+2) then also waits for O(u) to
  // be drained (o_free) before it overwrites the region.
  const bool o_in_region = p.S_pad > kMaxSpadShared;
  const uint32_t o_col0 = o_in_region ? (uint32_t)(p.S_pad / 2) : 2u * p.S_pad;
  const uint32_t o_col1 = o_in_region ? (uint32_t)(p.S_pad + p.S_pad / 2) : 2u * p.S_pad;

  if (warp == 13 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == 12 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);  // one elected lane per softmax warp
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&v_full[i], 2);   // expect_tx arrive + bias-written arrive
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(o_ready, 1);
    mbar_init(o_free, 4);
    fence_mbar_init();
  }
  if (warp == 15) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kv_bytes = p.S_pad * 128;

  if (warp == 13) {
    // ===================== producer: TMA tiles + key bias =====================
    int b = (int)blockIdx.x / p.H, h = (int)blockIdx.x - b * p.H;
    const int db = (int)gridDim.x / p.H, dh = (int)gridDim.x - db * p.H;
    int vs = 0;
    uint32_t vph = 0;  // V slot of head j and its use parity, advanced with the head (no divisions)
    for (int j = 0; j < n_my; ++j, b += db, h += dh) {
      if (h >= p.H) { h -= p.H; ++b; }
      const int row0 = b * p.S;
      const int qs = j & 1;
      if (lane == 0) {
        mbar_wait(&qk_empty[qs], ((j >> 1) & 1) ^ 1, 21);
        uint8_t* base = smem + L.qk + qs * L.qk_slot;
        mbar_arrive_expect_tx(&qk_full[qs], p.n_qt * Q_TILE_BYTES + kv_bytes);
        for (int t = 0; t < p.n_qt; ++t)
          tma_load_2d(&tmap_q, &qk_full[qs], base + t * Q_TILE_BYTES, h * 64, row0 + t * 128);
        tma_load_2d(&tmap_kv, &qk_full[qs], base + 2 * Q_TILE_BYTES, (p.H + h) * 64, row0);
        mbar_wait(&v_empty[vs], vph ^ 1, 22);
        mbar_arrive_expect_tx(&v_full[vs], kv_bytes);
        tma_load_2d(&tmap_kv, &v_full[vs], smem + L.v + vs * L.v_slot, (2 * p.H + h) * 64, row0);
      }
      __syncwarp();
      // additive key bias: 0 for attended keys, -inf for masked keys and the padding up to S_pad
      int first_bad = p.S_pad, last_ok = 0;
      for (int k = lane; k < p.S_pad; k += 32) {
        bool ok = k < p.S;
        if (ok && p.key_mask != nullptr) ok = p.key_mask[(long long)b * p.mask_stride + k] != 0;
        s_bias[vs * 256 + k] = ok ? 0.f : -INFINITY;
        if (!ok && k < first_bad) first_bad = k;
        if (ok) last_ok = k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, o));
        last_ok = max(last_ok, __shfl_xor_sync(0xffffffffu, last_ok, o));
      }
      if (lane == 0) {
        s_meta[vs] = first_bad;                       // 32-key chunks entirely below it need no bias at all
        s_meta[4 + vs] = (last_ok + 16) / 16 * 16;    // keys from here on are all masked: exp(-inf) = 0 exactly, so neither
      }                                               // the softmax nor the PV MMAs visit them (padded prompts: 16-32 of 256)
      __syncwarp();
      if (lane == 0) mbar_arrive(&v_full[vs]);
      if (++vs == L.v_slots) { vs = 0; vph ^= 1; }
    }
  } else if (warp == 12) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // (head j, query tile t) of the next S / PV unit, advanced with counters: no divisions on this thread
      int sj = 0, st = 0, pj = 0, pt = 0;
      int pvs = 0;
      uint32_t pvph = 0;  // V slot / parity of head pj
      auto issue_s = [&](int u) {
        const int j = sj, t = st, qs = j & 1, r = u & 1;
        if (++st == p.n_qt) { st = 0; ++sj; }
        if (t == 0) mbar_wait(&qk_full[qs], (j >> 1) & 1, 23);
        if (o_in_region && u >= 2) mbar_wait(o_free, (u - 2) & 1, 26);  // O(u-2) sits in this region: drained first
        tc_fence_after();
        // region r is free: PV(u-2) was issued before this in program order (the tensor pipe executes in order) and
        // softmax(u-2) finished reading S before p_ready(u-2), which PV(u-2) waited for
        const uint8_t* base = smem + L.qk + qs * L.qk_slot;
        const uint64_t dq = make_smem_desc_sw128(smem_u32(base + t * Q_TILE_BYTES));
        const uint64_t dk = make_smem_desc_sw128(smem_u32(base + 2 * Q_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)  // head dim 64 = 4 x 16
          umma_f16_ss(tmem_base + r * p.S_pad, dq + 2 * k, dk + 2 * k, p.idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_ready[r]);
        if (p.trace != nullptr && blockIdx.x == 0 && u < 64) p.trace[u * 16 + EV_S_ISSUE] = clock64();
        if (t == p.n_qt - 1) umma_commit(&qk_empty[qs]);  // Q/K of this head are dead once these MMAs complete
      };
      auto issue_pv = [&](int u) {
        const int t = pt, vs = pvs, r = u & 1;
        const uint32_t vph = pvph;
        mbar_wait(&p_ready[r], (u >> 1) & 1, 24);
        if (t == 0) mbar_wait(&v_full[vs], vph, 25);
        // the output group has drained O of the previous unit (shared accumulator). With O in the regions this wait is not
        // needed for the data, but it keeps this thread at most one o_free phase ahead, so the parity wait in issue_s
        // (phase u-2) cannot be satisfied by a stale phase u-4
        if (u > 0) mbar_wait(o_free, (u - 1) & 1, 26);
        tc_fence_after();
        if (p.trace != nullptr && blockIdx.x == 0 && u < 64) p.trace[u * 16 + EV_PV_WAITED] = clock64();
        const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + L.v + vs * L.v_slot));
        const int ksteps = s_meta[4 + vs] / 16;  // (v_full of this head was waited for at its first tile: the value is visible)
        const uint32_t o_col = r ? o_col1 : o_col0;
        for (int k = 0; k < ksteps; ++k)  // 16 keys per MMA: P advances 8 TMEM columns, V advances 16 rows = 2048 B
          umma_f16_ts(tmem_base + o_col, tmem_base + r * p.S_pad + 8 * k, dv + 128 * k, p.idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(o_ready);
        if (p.trace != nullptr && blockIdx.x == 0 && u < 64) p.trace[u * 16 + EV_PV_ISSUED] = clock64();
        if (t == p.n_qt - 1) umma_commit(&v_empty[vs]);
        if (++pt == p.n_qt) {
          pt = 0;
          ++pj;
          if (++pvs == L.v_slots) { pvs = 0; pvph ^= 1; }
        }
      };
      if (U > 0) issue_s(0);
      if (U > 1) issue_s(1);
      for (int u = 0; u < U; ++u) {
        issue_pv(u);
        if (u + 2 < U) issue_s(u + 2);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;   // group = S region
    const int q = warp & 3;          // TMEM lane quadrant
    const uint32_t t_row = tmem_base + g * p.S_pad + (uint32_t(q * 32) << 16);
    const int row_in_tile = q * 32 + lane;
    int uj = 0, ut = g;  // (head, tile) of unit u, advanced by two units per iteration
    int uvs = 0;
    uint32_t uvph = 0;   // V slot / parity of head uj
    while (ut >= p.n_qt) {
      ut -= p.n_qt; ++uj;
      if (++uvs == L.v_slots) { uvs = 0; uvph ^= 1; }
    }
    for (int u = g; u < U; u += 2) {
      const int vs = uvs, t_here = ut;
      const uint32_t vph = uvph;
      ut += 2;
      while (ut >= p.n_qt) {
        ut -= p.n_qt; ++uj;
        if (++uvs == L.v_slots) { uvs = 0; uvph ^= 1; }
      }
      const int n = u >> 1;
      mbar_wait(&v_full[vs], vph, 27);  // key bias / meta visible
      const float* bias = s_bias + vs * 256;
      // warps whose 32 query rows all lie beyond S (tail of the last tile) skip the TMEM traffic: their P rows
      // stay whatever they were, the corresponding O rows are never stored
      const bool live = t_here * 128 + q * 32 < p.S;
      const int fast_end = live ? (s_meta[vs] & ~31) : 0;    // keys [0, fast_end) are all attended: no bias needed
      const int slow_end = live ? s_meta[4 + vs] : 0;        // keys beyond are all masked
      mbar_wait(&s_ready[g], n & 1, 28);
      tc_fence_after();
      ATC_TRACE(u, EV_SM_START);
      // ---- single pass over S: TMEM read bandwidth (~64 B/clk/SM) is the scarce resource of this kernel, so S is
      // read exactly once. p = exp2(scale*s + bias - m_ref) with a LAZY reference: m_ref starts as ceil(max of the
      // first block) and is raised (to an integer, so the rescale factor is an exact power of two) only when a later
      // block exceeds it by more than 2^10; the P blocks already written are then rescaled in place. softmax is
      // shift-invariant, so O / sum is unchanged; P <= 2^10 stays far inside the fp16 range.
      float m_ref = -INFINITY, sum = 0.f;
      auto raise_ref = [&](float cm, int c_done) {  // warp-uniform call; cm = this lane's block max (log2 domain)
        const bool need = cm > m_ref + 10.0f;
        if (__any_sync(0xffffffffu, need)) {
          const float m_new = need ? ceilf(cm) : m_ref;
          const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);  // exact power of two, or 1
          sum *= f;
          uint32_t f2;
          if (kBf) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f, f);
            f2 = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __half2 h = __floats2half2_rn(f, f);
            f2 = *reinterpret_cast<uint32_t*>(&h);
          }
          if (c_done > 0) tmem_st_wait();  // the P blocks stored so far must have landed before they are re-read
          for (int cc = 0; cc < c_done; cc += 16) {  // P blocks written so far: 16 keys = 8 packed columns each
            uint32_t w[8];
            tmem_ld_32x8(t_row + (cc >> 1), w);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (kBf) {
                __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&w[i]), *reinterpret_cast<__nv_bfloat162*>(&f2));
                w[i] = *reinterpret_cast<uint32_t*>(&r);
              } else {
                __half2 r = __hmul2(*reinterpret_cast<__half2*>(&w[i]), *reinterpret_cast<__half2*>(&f2));
                w[i] = *reinterpret_cast<uint32_t*>(&r);
              }
            }
            tmem_st_32x8(t_row + (cc >> 1), w);
          }
          m_ref = m_new;
        }
      };
      auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
        float c0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), c1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
        float c2 = fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5])), c3 = fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
        for (int i = 8; i < 32; i += 4) {  // four independent chains: short dependency depth
          c0 = fmaxf(c0, __uint_as_float(v[i]));
          c1 = fmaxf(c1, __uint_as_float(v[i + 1]));
          c2 = fmaxf(c2, __uint_as_float(v[i + 2]));
          c3 = fmaxf(c3, __uint_as_float(v[i + 3]));
        }
        raise_ref(fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * p.scale_log2, c);
        const float neg_m = -m_ref;
        float s0 = 0.f, s1 = 0.f;
        uint32_t pk[16];
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_m, neg_m), one2 = make_float2(1.f, 1.f);
        float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // packed fp32 (FFMA2): the scale-and-shift and the running sums of two keys per instruction
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
          // (exp2 of part of the pairs from an FMA-pipe polynomial, FlashAttention-4 style, was measured here: 246 us with
          // none, 248 us with 2 of 8 pairs, 272 us with 3 of 8 - tools/microbench/softmax_exp_mix.cu)
          const float2 e = make_float2(ex2f(t.x), ex2f(t.y));
          acc2 = __ffma2_rn(e, one2, acc2);
          pk[i] = pack16(e.x, e.y, kBf);
        }
        s0 += acc2.x;
        s1 += acc2.y;
        sum += s0 + s1;
        tmem_st_32x16(t_row + (c >> 1), pk);
      };
      ATC_TRACE(u, EV_SM_P1);
      ATC_TRACE(u, EV_SM_BATON);
      // TMEM loads are pipelined so that no tcgen05.wait::ld directly follows the load it would expose: wait::ld covers
      // every outstanding load, and a tcgen05.ld takes ~250 clk in this kernel (MMAs in flight; ~35 clk on an idle SM).
      // Invariant at the top of each step: `va` valid, `vb` (the next block) in flight since one block of work.
      uint32_t va[32], vb[32];
      if (fast_end > 0) {
        tmem_ld_32x32(t_row, va);
        tmem_ld_wait_dep(va);
        if (fast_end > 32) tmem_ld_32x32(t_row + 32, vb);
        for (int c = 0; c < fast_end; c += 64) {
          exp_chunk(va, c);
          if (c + 32 >= fast_end) break;
          tmem_ld_wait_dep(vb);
          if (c + 64 < fast_end) tmem_ld_32x32(t_row + c + 64, va);
          exp_chunk(vb, c + 32);
          if (c + 64 >= fast_end) break;
          tmem_ld_wait_dep(va);
          if (c + 96 < fast_end) tmem_ld_32x32(t_row + c + 96, vb);
        }
      }
      for (int c = fast_end; c < slow_end; c += 16) {  // blocks that contain masked keys: additive 0 / -inf bias
        uint32_t v[16];
        tmem_ld_32x16(t_row + c, v);
        tmem_ld_wait();
        float t[16];
        float cm = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + c + i);
          t[i] = fmaf(__uint_as_float(v[i]), p.scale_log2, b4.x);
          t[i + 1] = fmaf(__uint_as_float(v[i + 1]), p.scale_log2, b4.y);
          t[i + 2] = fmaf(__uint_as_float(v[i + 2]), p.scale_log2, b4.z);
          t[i + 3] = fmaf(__uint_as_float(v[i + 3]), p.scale_log2, b4.w);
          cm = fmaxf(fmaxf(cm, fmaxf(t[i], t[i + 1])), fmaxf(t[i + 2], t[i + 3]));
        }
        raise_ref(cm, c);
        const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;  // everything masked so far: exp2(-inf) = 0
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e0 = ex2f(t[2 * i] + neg_m), e1 = ex2f(t[2 * i + 1] + neg_m);
          pk[i] = pack16(e0, e1, kBf);
          if (p.key_mask != nullptr) {  // masked (text) rows: sum the ROUNDED probabilities, as attention_tc.cu does
            const float2 r = unpack16(pk[i], kBf);
            sum += r.x + r.y;
          } else {
            sum += e0 + e1;
          }
        }
        tmem_st_32x8(t_row + (c >> 1), pk);
      }
      ATC_TRACE(u, EV_SM_P2);
      s_rowsum[(g * 2 + (n & 1)) * 128 + row_in_tile] = sum;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);
    }
  } else if (warp < 4) {
    // ===================== output group: O / row sum -> context rows =====================
    const int q = warp & 3;
    const uint32_t t_o0 = tmem_base + o_col0 + (uint32_t(q * 32) << 16), t_o1 = tmem_base + o_col1 + (uint32_t(q * 32) << 16);
    const int row_in_tile = q * 32 + lane;
    int b = (int)blockIdx.x / p.H, h = (int)blockIdx.x - b * p.H, t = -1;
    const int db = (int)gridDim.x / p.H, dh = (int)gridDim.x - db * p.H;
    for (int u = 0; u < U; ++u) {
      if (++t == p.n_qt) {
        t = 0;
        b += db;
        h += dh;
        if (h >= p.H) { h -= p.H; ++b; }
      }
      const int r = u & 1, n = u >> 1;
      mbar_wait(&p_ready[r], n & 1, 31);  // row sums of this unit are visible
      const float sum = s_rowsum[(r * 2 + (n & 1)) * 128 + row_in_tile];
      mbar_wait(o_ready, u & 1, 32);
      tc_fence_after();
      ATC_TRACE(u, EV_OUT_START);
      uint32_t va[32], vb[32];
      if (t * 128 + q * 32 < p.S) {  // warp-uniform: skip the tail warps of the last tile
        const uint32_t t_o = r ? t_o1 : t_o0;
        tmem_ld_32x32(t_o, va);
        tmem_ld_32x32(t_o + 32, vb);
        tmem_ld_wait_dep(va);
        tmem_ld_wait_dep(vb);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);  // O may be overwritten by the next PV
      ATC_TRACE(u, EV_OUT_DONE);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      const int srow = t * 128 + row_in_tile;
      if (srow < p.S) {
        uint16_t* orow = p.out + ((long long)b * p.S + srow) * p.out_pitch + h * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(half == 0 ? va[i + e] : vb[i + e]) * inv;
            uint4 w;
            w.x = pack16(f[0], f[1], kBf);
            w.y = pack16(f[2], f[3], kBf);
            w.z = pack16(f[4], f[5], kBf);
            w.w = pack16(f[6], f[7], kBf);
            *reinterpret_cast<uint4*>(orow + 32 * half + i) = w;
            if (p.lo_off > 0) {  // rounding remainder of the context: the output projection then runs split-operand
              const float2 h0 = unpack16(w.x, kBf), h1 = unpack16(w.y, kBf), h2 = unpack16(w.z, kBf), h3 = unpack16(w.w, kBf);
              uint4 l;
              l.x = pack16(f[0] - h0.x, f[1] - h0.y, kBf);
              l.y = pack16(f[2] - h1.x, f[3] - h1.y, kBf);
              l.z = pack16(f[4] - h2.x, f[5] - h2.y, kBf);
              l.w = pack16(f[6] - h3.x, f[7] - h3.y, kBf);
              *reinterpret_cast<uint4*>(orow + p.lo_off + 32 * half + i) = l;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 15) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

bool attention_tc1_supports(int S) { return S > 64 && (S + 15) / 16 * 16 <= kMaxSpad; }

int launch_attention_tc1(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                         int64_t mask_stride, float scale, cudaStream_t stream, int64_t out_pitch, int64_t lo_off,
                         long long* trace) {
  const int S_pad = (S + 15) / 16 * 16;
  if (S_pad > kMaxSpad) return set_error(KB_ERR_ARG, "attention_tc1: S=%d > %d", S, kMaxSpad);
  const int dt = bf16 ? KB_BF16 : KB_F16;
  const int64_t rows = (int64_t)B * S, cols = 3LL * H * 64;
  CUtensorMap tq, tkv;
  int rc = get_tmap_2d(qkv, dt, rows, cols, cols, 128, &tq);
  if (rc) return rc;
  rc = get_tmap_2d(qkv, dt, rows, cols, cols, S_pad, &tkv);
  if (rc) return rc;
  const int smem = Smem(S_pad).total + 1024;
  KB_TRY_ATTR(attention_tc1_kernel<true>, smem);
  KB_TRY_ATTR(attention_tc1_kernel<false>, smem);
  Atc1Params p;
  p.B = B; p.S = S; p.H = H; p.S_pad = S_pad; p.n_qt = (S + 127) / 128; p.items = B * H;
  p.key_mask = reinterpret_cast<const long long*>(key_mask);
  p.mask_stride = mask_stride;
  p.out = static_cast<uint16_t*>(out);
  p.out_pitch = out_pitch; p.lo_off = lo_off;
  p.scale_log2 = scale * 1.4426950408889634f;
  const uint32_t fmt = bf16 ? kFmtBF16 : kFmtF16;
  p.idesc_s = make_idesc(fmt, 128, S_pad, 0, 0);
  p.idesc_pv = make_idesc(fmt, 128, 64, 0, 1);  // B operand (V) is MN-major: rows are keys, 64 head-dim values contiguous
  p.bf16 = bf16;
  p.trace = trace;
  int grid = num_sms();
  if (p.items < grid) grid = p.items;
  if (bf16) attention_tc1_kernel<true><<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
  else attention_tc1_kernel<false><<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
