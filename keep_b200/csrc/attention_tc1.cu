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
//   * warp 14     : TMA producer. Q tile(s) + K go through a 2-slot ring, V through a 3-slot ring (Q/K are dead as
//                   soon as the S MMAs of the head are done, V only after its last PV MMA), all read straight out of
//                   the fused q|k|v projection buffer [B*S, 3*H*64] with SWIZZLE_128B boxes;
//   * warp 15     : MMA issuer. S = Q K^T (tcgen05.mma SS, 128 x S_pad x 16, fp32 S in TMEM region u&1), and
//                   O = P V (tcgen05.mma TS: P is read from TMEM where it overwrote S; V is the MN-major B operand).
//                   The PV MMAs of a unit are issued in TWO parts: the first 128 keys as soon as the softmax has packed
//                   them (p_half), i.e. UNDER the rest of the softmax, the remaining keys when the softmax is done
//                   (p_ready) - the tail between the end of a softmax and the next S on its region is 5 MMAs instead of
//                   13 for the ViT's 208 padded keys. The thread polls its barriers (mbarrier.test_wait) and serves, in
//                   this priority, PV tails, S of the next unit, one early PV MMA at a time: it never blocks on one
//                   group while the other has work for it. Everything is issued in unit order;
//   * warps 4-7 / 8-11 : two softmax groups, one per S region, each working on its own unit; a thread owns one query
//                   row: tcgen05.ld S (one block ahead of its wait) -> exp2 against the running reference (packed fp32
//                   FFMA2 for the scale-and-shift and the row sums, MUFU for exp2) -> 16-bit P written back over S with
//                   tcgen05.st. What bounds the kernel is the chain S(u) -> softmax(u) -> PV tail(u) -> S(u+2) on a
//                   region: the other group's unit fills the gaps;
//   * warps 0-3   : output group: tcgen05.ld O, divide by the row sum, store rows (optionally followed by their 16-bit
//                   rounding remainder: the hi|lo operand of a split-operand GEMM).
// TMEM layout of a region (stride max(S_pad, 128) columns, two regions): fp32 S in [0, S_pad); the softmax packs the
// 16-bit P of keys 0..127 into [0, 64) and of keys 128.. in place at [128, 128 + (S_pad - 128) / 2); the unit's O
// accumulator is [64, 128) - columns the softmax has consumed by the time it signals p_half, so the early PV MMAs can
// write there while the softmax is still reading the scores of keys >= 128. S(u+2) waits for O(u) to be drained.
// The single-thread roles sit on the highest warp ids (the SM's warp arbiter favours them; warp 15 shares its
// sub-partition with the softmax warps of TMEM lane quadrant 3, which idle on the ViT's 69-row second tile) and walk the
// units with counters instead of integer divisions: nothing hides a division's latency behind one thread.
//
// Reference semantics: timm Attention -> F.scaled_dot_product_attention and BertSelfAttention with
// the additive key mask (transformers modeling_bert.py:115-140).
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

constexpr int kAtcThreads = 512;
constexpr int kMaxSpad = 256;                 // two regions of max(S_pad, 128) columns fill the 512 TMEM columns
constexpr int kHalfKeys = 128;                // keys whose PV MMAs are issued under the rest of the softmax
constexpr int kOCol = 64;                     // O accumulator of a unit: columns [64, 128) of its own region
constexpr int Q_TILE_BYTES = 128 * 128;       // 128 rows x 64 x 16-bit
constexpr int kQkSlots = 2;
#ifndef KB_ATC_MMA_WARP
#define KB_ATC_MMA_WARP 15
#endif
constexpr int kMmaWarp = KB_ATC_MMA_WARP, kTmaWarp = KB_ATC_MMA_WARP == 15 ? 14 : 13, kAllocWarp = KB_ATC_MMA_WARP == 15 ? 13 : 15;

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
    v_slots = S_pad > 224 ? 2 : 3;
    qk = 0;
    v = kQkSlots * qk_slot;
    bias = v + v_slots * v_slot;              // [v_slots][256] float: 0 / -inf per key
    meta = bias + v_slots * 256 * 4;          // [v_slots] int: index of the first masked key
    rowsum = meta + 64;                       // 2 x [2 regions][2 parities][128] float (row sums, Oa factors)
    bars = rowsum + 2 * 2 * 2 * 128 * 4;
    total = bars + 256;
  }
};

// KB_ATC_EXP (never defined in the product build; tools/build_variant.sh): timing-only variants that remove one resource
// the softmax warps use, to see what the MMA thread competes for. bit 0: no MUFU (exp2 replaced by an FMA), bit 1: no
// tcgen05.st of P, bit 2: no tcgen05.ld of S after the first block, bit 3: no per-pair pack instruction, bit 4: p_half only at the end of the softmax (no early PV), bit 5: block maximum first
#ifndef KB_ATC_EXP
#define KB_ATC_EXP 32
#endif
__device__ __forceinline__ float ex2f(float x) {
  float y;
  if (KB_ATC_EXP & 1) return fmaf(x, 0.001f, 0.5f);
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
// TMEM column (inside the region) of the packed P of key block starting at key c (c a multiple of 16)
__device__ __forceinline__ uint32_t p_col(int c) { return (uint32_t)(c >> 1) + (c >= kHalfKeys ? 64u : 0u); }

// event slots of the optional trace (per unit)
enum { EV_S_ISSUE = 0, EV_SM_START = 1, EV_SM_HALF = 2, EV_SM_END = 3, EV_PVA_START = 4, EV_PVA_END = 5, EV_PVB_START = 6,
       EV_PVB_END = 7, EV_OUT_START = 8, EV_OUT_DONE = 9, EV_S_POLL = 10 };
#define ATC_TRACE(u, ev)                                                                         \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && (u) < 64 && (threadIdx.x & 127) == 0)           \
      p.trace[(u) * 16 + (ev)] = clock64();                                                      \
  } while (0)
#define ATC_TRACE1(u, ev)                                                                        \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && (u) < 64) p.trace[(u) * 16 + (ev)] = clock64(); \
  } while (0)

// SPLIT (129..224 padded keys: two regions + 64 spare TMEM columns): the PV tail accumulates into its own 64-column
// accumulator Ob in the spare columns (shared by both regions), the output group drains the early accumulator Oa as soon
// as its MMAs are done and adds the two: S(u+2) then follows PV-tail(u) on the tensor pipe without waiting for any drain.
template <bool BF16, bool SPLIT>
__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tc1_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const Atc1Params p) {
  constexpr int kBf = BF16 ? 1 : 0;  // compile-time 16-bit format: one F2FP per pair in the softmax, no predicated twin
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const Smem L(p.S_pad);
  float* s_bias = reinterpret_cast<float*>(smem + L.bias);
  int* s_meta = reinterpret_cast<int*>(smem + L.meta);
  float* s_rowsum = reinterpret_cast<float*>(smem + L.rowsum);   // [2 regions][2 parities][128] row sums ...
  float* s_oascale = s_rowsum + 2 * 2 * 128;                     // ... and (SPLIT) the power-of-two factor Oa is still owed
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* qk_full = bars;        // [2] TMA -> MMA
  uint64_t* qk_empty = bars + 2;   // [2] MMA (last S of the head committed) -> TMA
  uint64_t* v_full = bars + 4;     // [3] TMA + key-bias writer -> MMA, softmax
  uint64_t* v_empty = bars + 7;    // [3] MMA (last PV of the head committed) -> TMA
  uint64_t* s_ready = bars + 10;   // [2] MMA -> softmax group
  uint64_t* p_half = bars + 12;    // [2] softmax group -> MMA: P of the keys below kHalfKeys is in TMEM, their S is consumed
  uint64_t* p_ready = bars + 14;   // [2] softmax group -> MMA, output group: all of P and the row sums
  uint64_t* pva_done = bars + 16;  // [2] MMA (early PV MMAs of the unit completed) -> output group (SPLIT) / softmax (rare O rescale)
  uint64_t* o_ready = bars + 18;   // [2] MMA -> output group (SPLIT: [0] only, the shared tail accumulator, one phase per unit)
  uint64_t* o_free = bars + 20;    // [2] output group -> MMA: O drained (SPLIT: [0] only, Ob: the next PV tail may start)
  uint64_t* oa_free = bars + 22;   // [2] SPLIT: output group -> MMA: Oa(u) drained, S(u+2) may overwrite the region
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // heads of this CTA
  const int U = n_my * p.n_qt;                                                          // units of this CTA
  const uint32_t region = p.S_pad > kHalfKeys ? (uint32_t)p.S_pad : (uint32_t)kHalfKeys;  // TMEM columns per region
  const int c_half = p.S_pad < kHalfKeys ? p.S_pad : kHalfKeys;                          // keys of the early PV part
  const uint32_t ob_col = 2u * region;                                                   // SPLIT: the shared tail accumulator

  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&qk_full[i], 1);
      mbar_init(&qk_empty[i], 1);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_half[i], 4);   // one elected lane per softmax warp
      mbar_init(&p_ready[i], 4);
      mbar_init(&pva_done[i], 1);
      mbar_init(&o_ready[i], 1);
      mbar_init(&o_free[i], 4);   // one elected lane per output warp
      mbar_init(&oa_free[i], 4);
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&v_full[i], 2);   // expect_tx arrive + bias-written arrive
      mbar_init(&v_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == kAllocWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kv_bytes = p.S_pad * 128;

  if (warp == kTmaWarp) {
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
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    // Fixed issue order PVa(u) | PVtail(u) | S(u+2), every wait a suspending try_wait: with the two softmax groups half a
    // period apart that is also the order in which the barriers complete. (A polling scheduler over the three streams was
    // measured at twice the kernel time: this thread gets an issue slot every ~10 clk next to two softmax warps, so
    // instructions, not barriers, are what it must save - tools/microbench/mma_under_load.cu.)
    if (lane == 0) {
      int s_j = 0, s_t = 0;                    // S stream: (head, query tile) of the next unit
      int a_t = 0, a_vs = 0;                   // early-PV stream: tile / V slot / parity of the next unit
      uint32_t a_vph = 0;
      int b_t = 0, b_vs = 0;                   // PV-tail stream
      const int k_half = c_half / 16;
      auto issue_s = [&](int u) {
        const int qs = s_j & 1, r = u & 1;
        if (s_t == 0) mbar_wait(&qk_full[qs], (s_j >> 1) & 1, 23);
        // region r: softmax(u-2) has read all of S (p_ready, waited for by PVtail(u-2)), the PV MMAs that read P(u-2) were
        // issued before this (the tensor pipe executes in order), and the accumulator inside the region has been drained
        if (u >= 2) mbar_wait(SPLIT ? &oa_free[r] : &o_free[r], ((u >> 1) - 1) & 1, 26);
        tc_fence_after();
        const uint8_t* base = smem + L.qk + qs * L.qk_slot;
        const uint64_t dq = make_smem_desc_sw128(smem_u32(base + s_t * Q_TILE_BYTES));
        const uint64_t dk = make_smem_desc_sw128(smem_u32(base + 2 * Q_TILE_BYTES));
#pragma unroll
        for (int k = 0; k < 4; ++k)  // head dim 64 = 4 x 16
          umma_f16_ss(tmem_base + r * region, dq + 2 * k, dk + 2 * k, p.idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_ready[r]);
        ATC_TRACE1(u, EV_S_ISSUE);
        if (s_t == p.n_qt - 1) umma_commit(&qk_empty[qs]);  // Q/K of this head are dead once these MMAs complete
        if (++s_t == p.n_qt) { s_t = 0; ++s_j; }
      };
      auto issue_pva = [&](int u) {
        const int r = u & 1;
        mbar_wait(&p_half[r], (u >> 1) & 1, 24);
        if (a_t == 0) mbar_wait(&v_full[a_vs], a_vph, 25);
        tc_fence_after();
        ATC_TRACE1(u, EV_PVA_START);
        const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + L.v + a_vs * L.v_slot));
        const int ksteps = s_meta[4 + a_vs] / 16;
        const int ka = ksteps < k_half ? ksteps : k_half;
        const uint32_t t_reg = tmem_base + r * region;
        for (int k = 0; k < ka; ++k)  // 16 keys per MMA: P advances 8 TMEM columns, V advances 16 rows = 2048 B
          umma_f16_ts(t_reg + kOCol, t_reg + 8 * k, dv + 128 * k, p.idesc_pv, k != 0 ? 1u : 0u);
        umma_commit(&pva_done[r]);
        ATC_TRACE1(u, EV_PVA_END);
        if (++a_t == p.n_qt) {
          a_t = 0;
          if (++a_vs == L.v_slots) { a_vs = 0; a_vph ^= 1; }
        }
      };
      auto issue_pvb = [&](int u) {
        const int r = u & 1;
        mbar_wait(&p_ready[r], (u >> 1) & 1, 24);
        if (SPLIT && u > 0) mbar_wait(&o_free[0], (u - 1) & 1, 26);  // the shared tail accumulator has been drained
        tc_fence_after();
        ATC_TRACE1(u, EV_PVB_START);
        const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + L.v + b_vs * L.v_slot));
        const int ksteps = s_meta[4 + b_vs] / 16;  // (v_full of this head was waited for by the early part)
        const uint32_t t_reg = tmem_base + r * region;
        const uint32_t t_o = SPLIT ? tmem_base + ob_col : t_reg + kOCol;
        for (int k = k_half; k < ksteps; ++k)
          umma_f16_ts(t_o, t_reg + 64 + 8 * k, dv + 128 * k, p.idesc_pv, (!SPLIT || k != k_half) ? 1u : 0u);
        umma_commit(SPLIT ? &o_ready[0] : &o_ready[r]);
        ATC_TRACE1(u, EV_PVB_END);
        if (b_t == p.n_qt - 1) umma_commit(&v_empty[b_vs]);
        if (++b_t == p.n_qt) {
          b_t = 0;
          if (++b_vs == L.v_slots) b_vs = 0;
        }
      };
      if (U > 0) issue_s(0);
      if (U > 1) issue_s(1);
      for (int u = 0; u < U; ++u) {
        issue_pva(u);
        issue_pvb(u);
        if (u + 2 < U) issue_s(u + 2);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== softmax groups =====================
    const int g = (warp - 4) >> 2;   // group = S region
    const int q = warp & 3;          // TMEM lane quadrant
    const uint32_t t_row = tmem_base + g * region + (uint32_t(q * 32) << 16);
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
      // ---- single pass over S. p = exp2(scale*s + bias - m_ref) with a LAZY reference: m_ref starts as ceil(max of the
      // first block) and is raised (to an integer, so the rescale factor is an exact power of two) only when a later
      // block exceeds it by more than 2^10; the P blocks already written are then rescaled in place - and, when the
      // early PV MMAs have already consumed the first kHalfKeys keys, the O accumulator rows instead (below). softmax is
      // shift-invariant, so O / sum is unchanged; P <= 2^10 stays far inside the fp16 range.
      float m_ref = -INFINITY, sum = 0.f;
      float oa_scale = 1.f;     // SPLIT: factor the early accumulator Oa is still owed (raises after p_half), applied by the output group
      bool half_sent = false;
      auto send_half = [&]() {  // warp-uniform
        tmem_st_wait();         // P of the keys below c_half has landed
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_half[g]);
        half_sent = true;
        ATC_TRACE(u, EV_SM_HALF);
      };
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
          int cc0 = 0;
          if (half_sent) {
            // the early PV MMAs own P[0, c_half): what they accumulate has to be rescaled instead
            if (SPLIT) {
              oa_scale *= f;  // Oa is drained on its own: the output group multiplies it
            } else {
              // they were issued after p_half and are complete at pva_done; the PV tail, which accumulates on top, is not
              // issued before this warp arrives on p_ready
              mbar_wait(&pva_done[g], n & 1, 29);
              tc_fence_after();
              for (int oc = 0; oc < 64; oc += 16) {
                uint32_t w[16];
                tmem_ld_32x16(t_row + kOCol + oc, w);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = __float_as_uint(__uint_as_float(w[i]) * f);
                tmem_st_32x16(t_row + kOCol + oc, w);
              }
            }
            cc0 = c_half;
          }
          if (c_done > cc0) tmem_st_wait();  // the P blocks stored so far must have landed before they are re-read
          for (int cc = cc0; cc < c_done; cc += 16) {  // P blocks written so far: 16 keys = 8 packed columns each
            uint32_t w[8];
            tmem_ld_32x8(t_row + p_col(cc), w);
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
            tmem_st_32x8(t_row + p_col(cc), w);
          }
          m_ref = m_new;
        }
      };
      auto chunk_max = [&](const uint32_t (&v)[32]) {
        float c0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), c1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
        float c2 = fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5])), c3 = fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
        for (int i = 8; i < 32; i += 4) {  // four independent chains: short dependency depth
          c0 = fmaxf(c0, __uint_as_float(v[i]));
          c1 = fmaxf(c1, __uint_as_float(v[i + 1]));
          c2 = fmaxf(c2, __uint_as_float(v[i + 2]));
          c3 = fmaxf(c3, __uint_as_float(v[i + 3]));
        }
        return fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * p.scale_log2;
      };
      // The sub-partition's issue slots are what the softmax warps, the MMA thread and the output warps compete for
      // (~7 instructions per pair of scores against 16 clk of MUFU for the two softmax warps), so the block maximum is NOT
      // part of the steady state: only the first block takes a real maximum for the reference; a later block is
      // exponentiated against the current reference straight away and its SUM decides: every P of the block is <= the
      // block sum, so a sum <= 2^15 proves that all of them fit the 16-bit format, and only a block whose sum exceeds that
      // (an element more than 2^10 above the reference: rare) takes the maximum, raises the reference and is redone.
      auto exp_block = [&](const uint32_t (&v)[32], uint32_t (&pk)[16]) {
        const float neg_m = -m_ref;
        const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_m, neg_m), one2 = make_float2(1.f, 1.f);
        float2 acc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          // packed fp32 (FFMA2): the scale-and-shift and the running sums of two keys per instruction; four independent
          // sum chains, so that the block sum (which the overflow vote below waits for) is 4 + 3 additions deep, not 16
          const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
          // (exp2 of part of the pairs from an FMA-pipe polynomial, FlashAttention-4 style, was measured here: 246 us with
          // none, 248 us with 2 of 8 pairs, 272 us with 3 of 8 - tools/microbench/softmax_exp_mix.cu)
          const float2 e = make_float2(ex2f(t.x), ex2f(t.y));
          acc[i & 3] = __ffma2_rn(e, one2, acc[i & 3]);
          pk[i] = (KB_ATC_EXP & 8) ? __float_as_uint(e.x) : pack16(e.x, e.y, kBf);
        }
        return (acc[0].x + acc[0].y + acc[1].x + acc[1].y) + (acc[2].x + acc[2].y + acc[3].x + acc[3].y);
      };
      auto exp_chunk = [&](const uint32_t (&v)[32], int c) {
        uint32_t pk[16];
        float bs;
        if (KB_ATC_EXP & 32) {  // experiment: the round-1/2 form, maximum of every block first
          raise_ref(chunk_max(v), c);
          bs = exp_block(v, pk);
        } else {
          if (c == 0) raise_ref(chunk_max(v), 0);
          bs = exp_block(v, pk);
          if (__any_sync(0xffffffffu, !(bs <= 32768.f))) {  // (also taken by inf / NaN sums)
            raise_ref(chunk_max(v), c);
            bs = exp_block(v, pk);
          }
        }
        sum += bs;
        if (!(KB_ATC_EXP & 2)) tmem_st_32x16(t_row + p_col(c), pk);
        if (c + 32 == c_half && !(KB_ATC_EXP & 16)) send_half();
      };
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
          if (c + 64 < fast_end && !(KB_ATC_EXP & 4)) tmem_ld_32x32(t_row + c + 64, va);
          exp_chunk(vb, c + 32);
          if (c + 64 >= fast_end) break;
          tmem_ld_wait_dep(va);
          if (c + 96 < fast_end && !(KB_ATC_EXP & 4)) tmem_ld_32x32(t_row + c + 96, vb);
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
        tmem_st_32x8(t_row + p_col(c), pk);
        if (c + 16 == c_half && !(KB_ATC_EXP & 16)) send_half();
      }
      // (keys in [slow_end, c_half) are all masked and visited by no MMA: nothing left to write for the early part)
      if (!half_sent) send_half();
      s_rowsum[(g * 2 + (n & 1)) * 128 + row_in_tile] = sum;
      if (SPLIT) s_oascale[(g * 2 + (n & 1)) * 128 + row_in_tile] = oa_scale;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[g]);
      ATC_TRACE(u, EV_SM_END);
    }
  } else if (warp < 4) {
    // ===================== output group: O / row sum -> context rows =====================
    const int q = warp & 3;
    const uint32_t lane_off = uint32_t(q * 32) << 16;
    const uint32_t t_o0 = tmem_base + kOCol + lane_off, t_o1 = t_o0 + region, t_ob = tmem_base + ob_col + lane_off;
    const int row_in_tile = q * 32 + lane;
    int b = (int)blockIdx.x / p.H, h = (int)blockIdx.x - b * p.H, t = -1;
    const int db = (int)gridDim.x / p.H, dh = (int)gridDim.x - db * p.H;
    int ovs = -1;  // V slot of the unit's head (SPLIT: whether the unit has a PV tail at all is read from its key meta)
    for (int u = 0; u < U; ++u) {
      if (++t == p.n_qt) {
        t = 0;
        b += db;
        h += dh;
        if (h >= p.H) { h -= p.H; ++b; }
      }
      if (t == 0 && ++ovs == L.v_slots) ovs = 0;
      const int r = u & 1, n = u >> 1;
      const bool live = t * 128 + q * 32 < p.S;  // warp-uniform: the tail warps of the last tile have nothing to store
      const uint32_t t_o = r ? t_o1 : t_o0;
      uint32_t va[32], vb[32];
      float sum;
      if (SPLIT) {
        // early accumulator: drained as soon as its MMAs are done (under the rest of the softmax), which frees the region
        mbar_wait(&pva_done[r], n & 1, 30);
        tc_fence_after();
        if (live) {
          tmem_ld_32x32(t_o, va);
          tmem_ld_32x32(t_o + 32, vb);
          tmem_ld_wait_dep(va);
          tmem_ld_wait_dep(vb);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&oa_free[r]);
        mbar_wait(&p_ready[r], n & 1, 31);  // row sums, Oa factors and (transitively) the key meta of this unit are visible
        sum = s_rowsum[(r * 2 + (n & 1)) * 128 + row_in_tile];
        const float fa = s_oascale[(r * 2 + (n & 1)) * 128 + row_in_tile];
        const bool has_tail = s_meta[4 + ovs] > c_half;
        mbar_wait(&o_ready[0], u & 1, 32);
        tc_fence_after();
        ATC_TRACE(u, EV_OUT_START);
        if (live) {
          if (has_tail) {
            uint32_t vc[32];
            tmem_ld_32x32(t_ob, vc);
            tmem_ld_wait_dep(vc);
#pragma unroll
            for (int i = 0; i < 32; ++i) va[i] = __float_as_uint(fmaf(__uint_as_float(va[i]), fa, __uint_as_float(vc[i])));
            tmem_ld_32x32(t_ob + 32, vc);
            tmem_ld_wait_dep(vc);
#pragma unroll
            for (int i = 0; i < 32; ++i) vb[i] = __float_as_uint(fmaf(__uint_as_float(vb[i]), fa, __uint_as_float(vc[i])));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              va[i] = __float_as_uint(__uint_as_float(va[i]) * fa);
              vb[i] = __float_as_uint(__uint_as_float(vb[i]) * fa);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[0]);  // the shared tail accumulator may be overwritten by the next unit's tail
      } else {
        mbar_wait(&p_ready[r], n & 1, 31);  // row sums of this unit are visible
        sum = s_rowsum[(r * 2 + (n & 1)) * 128 + row_in_tile];
        mbar_wait(&o_ready[r], n & 1, 32);
        tc_fence_after();
        ATC_TRACE(u, EV_OUT_START);
        if (live) {
          tmem_ld_32x32(t_o, va);
          tmem_ld_32x32(t_o + 32, vb);
          tmem_ld_wait_dep(va);
          tmem_ld_wait_dep(vb);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&o_free[r]);  // the region may be overwritten by S(u+2)
      }
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
  if (warp == kAllocWarp) {
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
  KB_TRY_ATTR((attention_tc1_kernel<true, true>), smem);
  KB_TRY_ATTR((attention_tc1_kernel<false, true>), smem);
  KB_TRY_ATTR((attention_tc1_kernel<true, false>), smem);
  KB_TRY_ATTR((attention_tc1_kernel<false, false>), smem);
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
  // separate tail accumulator whenever the two regions leave 64 TMEM columns and the unit has a tail at all
  const bool split = S_pad > kHalfKeys && 2 * S_pad + 64 <= 512;
  if (split) {
    if (bf16) attention_tc1_kernel<true, true><<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
    else attention_tc1_kernel<false, true><<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
  } else {
    if (bf16) attention_tc1_kernel<true, false><<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
    else attention_tc1_kernel<false, false><<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
  }
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
