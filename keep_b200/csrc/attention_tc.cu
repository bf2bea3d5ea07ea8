// tcgen05 attention, head dim 64, for every sequence length the towers use (1 .. 512 tokens):
//     out = softmax(Q K^T * scale + key_mask) V      per (sequence, head)
//
// One persistent CTA per SM; everything between the two MMAs stays on-chip (FlashAttention-4 style roles), and the keys
// of a unit are processed in KV BLOCKS of at most 112 keys so that the three stages of a unit overlap each other:
//
//   unit   = 128 query rows of one (sequence, head)                       [S > 56: ViT 197 tokens = 2 units per head]
//            or G = floor(112 / S) whole sequences of one head packed into one tile, block-diagonal   [S <= 56: prompts]
//   block  = up to 112 keys of the unit (ViT: 112 + 96; padded BERT S = 256: 112 + 112 + 32)
//   TMEM   = four S/P buffers of 112 columns, used round-robin by the blocks in issue order (block g -> buffer g & 3),
//            + O (64 columns)
//
//   warps 13, 14 : TMA producers, read straight out of the fused q|k|v projection buffer [rows, 3*H*64] with SWIZZLE_128B
//                  boxes. Warp 13: Q tile per unit (2 slots) and K block per block (4-slot ring); warp 14: V block (8-slot
//                  ring) and the additive key bias (0 / -inf) of the block next to it. Two producers, because a K block is
//                  wanted four blocks before its V block (S-MMAs run that far ahead of the PV-MMAs): behind one thread the
//                  K loads would queue up after V slots that are not free yet;
//   warps 15, 12 : MMA issuers (one thread each). Warp 15: S_g = Q K_g^T (SS-MMA 128 x NB x 16 into buffer g & 3), up to four
//                  blocks ahead: S(g + 4) re-uses the buffer of block g as soon as PV(g) has completed. Warp 12:
//                  O += P_g V_g (TS-MMA: P is read from TMEM where it overwrote S_g; V is the MN-major B operand). So the
//                  softmax of a block overlaps the S-MMAs of the blocks ahead and the PV-MMAs of the block before;
//   warps 4-7 / 8-11 : two softmax groups (units alternate between them), a thread owns one query row. Single pass over S with a LAZY
//                  reference: p = exp2(scale*s + bias - m_ref), m_ref starts as ceil(max of the first chunk) and is raised
//                  (to an integer: the rescale factor is an exact power of two) only when a later chunk exceeds it by
//                  more than 2^10. A raise rescales the P chunks of the current block in place and - when earlier blocks
//                  of the unit were already multiplied into O - the O rows (after waiting for those PV MMAs): rare, exact;
//   warps 0-3    : output group: tcgen05.ld O (one 64-column accumulator), divide by the row sum, store the context rows
//                  (optionally followed by their 16-bit rounding remainder: the hi|lo operand of a split-operand GEMM).
//
// Reference semantics: timm Attention -> F.scaled_dot_product_attention and BertSelfAttention
// with the additive key mask (transformers modeling_bert.py:115-140).
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

constexpr int kAtcThreads = 512;
constexpr int kBlk = 112;                  // keys per KV block = TMEM columns of one S/P buffer
constexpr int kOCol = 4 * kBlk;            // 448: the O accumulator (64 columns) sits after the four buffers
constexpr int Q_TILE_BYTES = 128 * 128;    // 128 rows x 64 x 16-bit
constexpr int KV_BLK_BYTES = kBlk * 128;   // 14 KB (a multiple of 1024: every slot keeps the swizzle alignment)
constexpr int kQSlots = 2, kKSlots = 4, kVSlots = 8;
constexpr int kPackMaxS = kBlk / 2;        // sequences of up to 56 tokens are packed two or more per tile
// Warp roles. The SM's warp arbiter favours high warp ids (B300_MICROARCH.md: "hi-wid-first"), so the three single-thread
// roles whose instruction streams sit on everybody's critical path get the highest ids; the output group, which only
// drains O once per unit, the lowest.
constexpr int kMmaWarp = 12, kProdKWarp = 13, kProdVWarp = 14, kSmmaWarp = 15, kAllocWarp = 15;  // warps 0-3 output, 4-11 softmax

struct AtcParams {
  int S, H;
  int n_seq;       // sequences (batch)
  int items;       // n_seq * H (one sequence per unit) or groups * H (packed)
  int n_qt;        // units per item: ceil(S / 128), or 1 when packed
  int nb;          // KV blocks per unit
  int G;           // sequences per unit when packed, 0 otherwise
  int kbox;        // rows of the K/V TMA box
  const long long* key_mask;
  long long mask_stride;
  uint16_t* out;
  long long out_pitch, lo_off;
  float scale_log2;
  uint32_t fmt;       // kFmtF16 / kFmtBF16
  uint32_t idesc_pv;  // 128 x 64, B (V) MN-major
  int bf16;
  long long* trace;   // optional [64 units][16 events] clock64 stamps of CTA 0 (profiling aid), or null
};

struct Smem {  // offsets inside the 1024-aligned dynamic smem block
  static constexpr int q = 0;
  static constexpr int k = kQSlots * Q_TILE_BYTES;
  static constexpr int v = k + kKSlots * KV_BLK_BYTES;
  static constexpr int bias = v + kVSlots * KV_BLK_BYTES;   // [kVSlots][kBlk] float: 0 / -inf per key of the block
  static constexpr int meta = bias + kVSlots * kBlk * 4;    // [kVSlots] int: first masked key of the block
  static constexpr int rowsum = meta + 64;                  // [8][128] float: row sums of unit u in slot u & 7 (up to 4 units
                                                            // are ahead of the output group, see the MMA issue order)
  static constexpr int bars = rowsum + 8 * 128 * 4;
  static constexpr int total = bars + 512;  // 62 mbarriers + the TMEM base address
};

struct Unit {
  int q_row0;   // first query row (row of the q|k|v matrix)
  int q_rows;   // live query rows of the tile (<= 128)
  int kv_row0;  // first key row
  int n_keys;   // keys of the unit (before padding to 16)
  int h;        // head
  int seq0;     // first sequence of the unit
  int ns;       // sequences in the unit (1 unless packed)
};

// Every role walks the CTA's units in the same order: item = blockIdx.x + i * gridDim.x, units t = 0 .. n_qt - 1 of each
// item. The walk is incremental (one division at the start): the single-thread roles have nobody to hide the latency of
// an integer division behind, and the first version of this kernel lost 3,000 cycles per unit to them.
struct Walker {
  int t, b, h;   // query tile inside the item; item / H (sequence, or group of packed sequences); item % H (head)
  int db, dh;    // gridDim.x / H, gridDim.x % H
  __device__ __forceinline__ Walker(const AtcParams& p) {
    t = 0;
    b = (int)blockIdx.x / p.H;
    h = (int)blockIdx.x - b * p.H;
    db = (int)gridDim.x / p.H;
    dh = (int)gridDim.x - db * p.H;
  }
  __device__ __forceinline__ void next(const AtcParams& p) {
    if (++t == p.n_qt) {
      t = 0;
      b += db;
      h += dh;
      if (h >= p.H) { h -= p.H; ++b; }
    }
  }
  __device__ __forceinline__ Unit unit(const AtcParams& p) const {
    Unit u;
    u.h = h;
    if (p.G == 0) {
      u.seq0 = b;
      u.kv_row0 = b * p.S;
      u.q_row0 = u.kv_row0 + t * 128;
      u.q_rows = min(128, p.S - t * 128);
      u.n_keys = p.S;
      u.ns = 1;
    } else {
      u.seq0 = b * p.G;
      u.ns = min(p.G, p.n_seq - u.seq0);
      u.q_row0 = u.kv_row0 = u.seq0 * p.S;
      u.q_rows = u.n_keys = u.ns * p.S;
    }
    return u;
  }
};
// keys of block j of a unit, padded to the MMA granularity of 16
__device__ __forceinline__ int block_keys(const Unit& u, int j) { return min(kBlk, ((u.n_keys + 15) & ~15) - j * kBlk); }

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
__device__ __forceinline__ uint32_t scale16(uint32_t w, uint32_t f2, int bf16) {  // packed 16-bit pair times a packed factor
  if (bf16) {
    __nv_bfloat162 r = __hmul2(*reinterpret_cast<__nv_bfloat162*>(&w), *reinterpret_cast<__nv_bfloat162*>(&f2));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __half2 r = __hmul2(*reinterpret_cast<__half2*>(&w), *reinterpret_cast<__half2*>(&f2));
  return *reinterpret_cast<uint32_t*>(&r);
}

// Rare path of the lazy softmax reference: a later KV block raised the reference after earlier blocks of the unit were
// already multiplied into O. Wait for those PV MMAs, then scale this warp's 32 O rows by f (a power of two). Kept out of
// line so that its registers do not count against the softmax loop.
static __device__ __noinline__ void rescale_o_rows(uint32_t t_o, float f, uint64_t* pv_done, uint32_t parity) {
  mbar_wait(pv_done, parity, 31);
  tc_fence_after();
#pragma unroll 1
  for (int c = 0; c < 64; c += 16) {
    uint32_t o[16];
    tmem_ld_32x16(t_o + c, o);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
    tmem_st_32x16(t_o + c, o);
  }
  tmem_st_wait();
}

// event slots of the optional trace (per unit)
enum { EV_S_ISSUE = 0, EV_PV0_WAITED = 1, EV_PV_ISSUED = 2, EV_SM_START = 3, EV_SM_B0 = 4, EV_SM_END = 5, EV_OUT_START = 6,
       EV_OUT_DONE = 7, EV_PVL_WAITED = 8, EV_SM_VFULL = 9, EV_SM_LD0 = 10, EV_SM_FAST = 11, EV_SM_STW = 12, EV_S_BEGIN = 13,
       EV_S_WAITED = 14 };
#define ATC_TRACE(u, ev)                                                                         \
  do {                                                                                           \
    if (p.trace != nullptr && blockIdx.x == 0 && (u) < 64 && (threadIdx.x & 127) == 0)           \
      p.trace[(u) * 16 + (ev)] = clock64();                                                      \
  } while (0)

__global__ void __launch_bounds__(kAtcThreads, 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_kv,
                    const AtcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* s_bias = reinterpret_cast<float*>(smem + Smem::bias);
  int* s_meta = reinterpret_cast<int*>(smem + Smem::meta);
  float* s_rowsum = reinterpret_cast<float*>(smem + Smem::rowsum);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem::bars);
  uint64_t* q_full = bars;          // [2] TMA -> MMA
  uint64_t* q_empty = bars + 2;     // [2] MMA (last S block of the unit committed) -> TMA
  uint64_t* k_full = bars + 4;      // [4] TMA -> MMA
  uint64_t* k_empty = bars + 8;     // [4] MMA (S block committed) -> TMA
  uint64_t* v_full = bars + 12;     // [8] TMA + key-bias writer -> MMA, softmax
  uint64_t* v_empty = bars + 20;    // [8] MMA (PV block committed) -> TMA
  // The next three exist once per (softmax group, buffer): a barrier is only ever waited on by threads that see EVERY one
  // of its phases in order (parity waits cannot tell phase k from phase k + 2), and a buffer serves both groups in turn
  // whenever the number of blocks per unit is odd.
  uint64_t* s_ready = bars + 28;    // [2][4] MMA -> softmax group
  uint64_t* p_ready = bars + 36;    // [2][4] softmax group -> MMA
  uint64_t* pv_done = bars + 44;    // [2][4] MMA: PV of a non-final block complete (O-correction path of the softmax)
  uint64_t* o_ready = bars + 52;    // [1] MMA -> output group
  uint64_t* o_free = bars + 53;     // [1] output group -> MMA
  uint64_t* sum_ready = bars + 54;  // [8] softmax group -> output group: row sums of unit u written (slot u & 7: the
                                    // softmax may run up to four units ahead of the output group)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 62);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_my = (p.items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // items of this CTA
  const int U = n_my * p.n_qt;                                                          // units of this CTA
  const int nb = p.nb;
  // block g = u * nb + j (unit-major) uses S/P buffer g & 3, K slot g & 3, V slot g & 3; its softmax group is u & 1.
  // Every role walks the blocks in the same order and keeps one parity bit per (group, buffer) barrier: bit 4 * group + buffer.

  if (warp == kProdKWarp && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_kv);
  }
  if (warp == kMmaWarp && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    for (int i = 0; i < kKSlots; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < kVSlots; ++i) {
      mbar_init(&v_full[i], 2);   // expect_tx arrive + bias-written arrive
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 4);  // one elected lane per softmax warp
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(o_ready, 1);
    mbar_init(o_free, 4);
    for (int i = 0; i < 8; ++i) mbar_init(&sum_ready[i], 4);
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
  const int kv_bytes = p.kbox * 128;

  if (warp == kProdKWarp) {
    // ===================== producer 1: Q tiles and K blocks =====================
    if (lane == 0) {
      Walker w(p);
      for (int u = 0; u < U; ++u, w.next(p)) {
        const Unit un = w.unit(p);
        const int qs = u & 1;
        mbar_wait(&q_empty[qs], ((u >> 1) & 1) ^ 1, 21);
        mbar_arrive_expect_tx(&q_full[qs], Q_TILE_BYTES);
        tma_load_2d(&tmap_q, &q_full[qs], smem + Smem::q + qs * Q_TILE_BYTES, un.h * 64, un.q_row0);
        for (int j = 0; j < nb; ++j) {
          const int g = u * nb + j, ks = g & (kKSlots - 1);
          mbar_wait(&k_empty[ks], ((g / kKSlots) & 1) ^ 1, 22);
          mbar_arrive_expect_tx(&k_full[ks], kv_bytes);
          tma_load_2d(&tmap_kv, &k_full[ks], smem + Smem::k + ks * KV_BLK_BYTES, (p.H + un.h) * 64, un.kv_row0 + j * kBlk);
        }
      }
    }
  } else if (warp == kProdVWarp) {
    // ===================== producer 2: V blocks + key bias =====================
    Walker w(p);
    for (int u = 0; u < U; ++u, w.next(p)) {
      const Unit un = w.unit(p);
      for (int j = 0; j < nb; ++j) {
        const int g = u * nb + j, vs = g & (kVSlots - 1);
        if (lane == 0) {
          mbar_wait(&v_empty[vs], ((g / kVSlots) & 1) ^ 1, 23);
          mbar_arrive_expect_tx(&v_full[vs], kv_bytes);
          tma_load_2d(&tmap_kv, &v_full[vs], smem + Smem::v + vs * KV_BLK_BYTES, (2 * p.H + un.h) * 64, un.kv_row0 + j * kBlk);
        }
        __syncwarp();  // the V slot (and with it its bias row) is free: the PV MMAs that used it have completed
        // additive key bias: 0 for attended keys, -inf for masked keys and the padding up to a multiple of 16
        const int nk = block_keys(un, j);
        int first_bad = nk;
        for (int k = lane; k < nk; k += 32) {
          const int kk = j * kBlk + k;  // key index inside the unit
          bool ok = kk < un.n_keys;
          if (ok && p.key_mask != nullptr) {
            const int seq = p.G == 0 ? 0 : kk / p.S;
            ok = p.key_mask[(long long)(un.seq0 + seq) * p.mask_stride + (kk - seq * p.S)] != 0;
          }
          s_bias[vs * kBlk + k] = ok ? 0.f : -INFINITY;
          if (!ok && k < first_bad) first_bad = k;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) first_bad = min(first_bad, __shfl_xor_sync(0xffffffffu, first_bad, o));
        if (lane == 0) s_meta[vs] = first_bad;  // 32-key chunks entirely below it need no bias at all
        __syncwarp();
        if (lane == 0) mbar_arrive(&v_full[vs]);
      }
    }
  } else if (warp == kSmmaWarp) {
    // ===================== MMA issuer 1: S_g = Q K_g^T =====================
    // Two issuing threads (on different SM sub-partitions) because the issue rate of ONE thread is what bounded the
    // first version: a tcgen05.mma costs its issuing thread >= 70-80 clk whatever its size (tools/microbench/mma_rate.cu),
    // a unit needs 8 + 13 of them plus a dozen mbarrier round trips. The tensor pipe no longer orders S(g + 4) behind
    // PV(g) by itself, so this thread waits for PV(g)'s completion (v_empty) before it re-uses buffer g & 3.
    if (lane == 0) {
      Walker w(p);
      int u = 0, j = 0, nk = block_keys(w.unit(p), 0);
      const int total = U * nb;
      const uint32_t q_desc0 = smem_u32(smem + Smem::q), k_desc0 = smem_u32(smem + Smem::k);
      for (int g = 0; g < total; ++g) {
        const int ks = g & (kKSlots - 1), qs = u & 1, buf = g & 3, gb = 4 * (u & 1) + buf;
        if (g >= 4) mbar_wait(&v_empty[(g - 4) & (kVSlots - 1)], ((g - 4) / kVSlots) & 1, 34);  // PV(g - 4) has read P
        if (j == 0) mbar_wait(&q_full[qs], (u >> 1) & 1, 24);
        mbar_wait(&k_full[ks], (g / kKSlots) & 1, 25);
        tc_fence_after();
        const uint32_t idesc = make_idesc(p.fmt, 128, (uint32_t)nk, 0, 0);
        const uint64_t dq = make_smem_desc_sw128(q_desc0 + qs * Q_TILE_BYTES);
        const uint64_t dk = make_smem_desc_sw128(k_desc0 + ks * KV_BLK_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // head dim 64 = 4 x 16
          umma_f16_ss(tmem_base + buf * kBlk, dq + 2 * k, dk + 2 * k, idesc, k != 0 ? 1u : 0u);
        umma_commit(&s_ready[gb]);
        umma_commit(&k_empty[ks]);                       // K block dead once these MMAs complete
        if (j == nb - 1) umma_commit(&q_empty[qs]);      // ... and the Q tile after the unit's last S block
        if (j == 0 && p.trace != nullptr && blockIdx.x == 0 && u < 64) p.trace[u * 16 + EV_S_ISSUE] = clock64();
        if (++j == nb) { j = 0; ++u; w.next(p); }
        nk = block_keys(w.unit(p), j);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer 2: O += P_g V_g =====================
    if (lane == 0) {
      uint32_t p_par = 0;  // parity of the next p_ready phase per (group, buffer)
      Walker w(p);
      int u = 0, j = 0, nk = block_keys(w.unit(p), 0);
      const int total = U * nb;
      const uint32_t v_desc0 = smem_u32(smem + Smem::v);
      for (int g = 0; g < total; ++g) {
        const int vs = g & (kVSlots - 1), buf = g & 3, gb = 4 * (u & 1) + buf;
        mbar_wait(&p_ready[gb], (p_par >> gb) & 1u, 26);
        p_par ^= 1u << gb;
        mbar_wait(&v_full[vs], (g / kVSlots) & 1, 27);
        if (j == 0 && u > 0) mbar_wait(o_free, (u - 1) & 1, 28);  // the output group has drained O of the previous unit
        tc_fence_after();
        if (p.trace != nullptr && blockIdx.x == 0 && u < 64) p.trace[u * 16 + (j == 0 ? EV_PV0_WAITED : EV_PVL_WAITED)] = clock64();
        const uint64_t dv = make_smem_desc_sw128(v_desc0 + vs * KV_BLK_BYTES);
        const uint32_t a0 = tmem_base + buf * kBlk;
        const int ksteps = nk >> 4;
        // 16 keys per MMA: P advances 8 TMEM columns, V advances 16 rows = 2048 B
        umma_f16_ts(tmem_base + kOCol, a0, dv, p.idesc_pv, j != 0 ? 1u : 0u);
        for (int k = 1; k < ksteps; ++k) umma_f16_ts(tmem_base + kOCol, a0 + 8 * k, dv + 128 * k, p.idesc_pv, 1u);
        umma_commit(&v_empty[vs]);   // V block dead, P read: the S issuer and the V producer both wait on it
        if (j == nb - 1) umma_commit(o_ready);
        else umma_commit(&pv_done[gb]);
        if (j == nb - 1 && p.trace != nullptr && blockIdx.x == 0 && u < 64) p.trace[u * 16 + EV_PV_ISSUED] = clock64();
        if (++j == nb) { j = 0; ++u; w.next(p); }
        nk = block_keys(w.unit(p), j);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ===================== softmax groups =====================
    const int grp = (warp - 4) >> 2;  // units alternate between the two groups
    const int q = warp & 3;           // TMEM lane quadrant
    const uint32_t lane_bits = uint32_t(q * 32) << 16;
    const uint32_t t_o = tmem_base + kOCol + lane_bits;
    const int row_in_tile = q * 32 + lane;
    uint32_t s_par = 0, pv_par = 0;  // parities of this group's s_ready / pv_done barriers, one bit per buffer
    // packed units: which sequences of the tile this warp's rows / this thread's row belong to (divisions hoisted)
    const int pk_first = p.G != 0 ? (q * 32) / p.S : 0, pk_last = p.G != 0 ? (q * 32 + 31) / p.S : 0;
    const int pk_mine = p.G != 0 ? row_in_tile / p.S : 0;
    Walker w(p);
    if (grp == 1 && U > 1) w.next(p);
    for (int u = grp; u < U; u += 2, w.next(p), w.next(p)) {
      const Unit un = w.unit(p);
      // warps whose 32 query rows all lie beyond the unit's rows (tail of the last tile) skip the TMEM traffic: their P
      // rows stay whatever they were, the corresponding O rows are never stored
      const bool live = q * 32 < un.q_rows;
      // packed units: a row sees only the keys of its own sequence; a warp walks the columns of the sequences its rows
      // belong to and writes zeros elsewhere
      int my_lo = 0, my_hi = kBlk, w_lo = 0, w_hi = kBlk;
      if (p.G != 0) {
        const int s_last = min(pk_last, un.ns - 1);
        w_lo = (pk_first * p.S) & ~15;
        w_hi = ((s_last + 1) * p.S + 15) & ~15;
        my_lo = row_in_tile < un.q_rows ? pk_mine * p.S : 0;
        my_hi = row_in_tile < un.q_rows ? my_lo + p.S : 0;
      }
      float m_ref = -INFINITY, sum = 0.f;
      for (int j = 0; j < nb; ++j) {
        const int g = u * nb + j, vs = g & (kVSlots - 1), buf = g & 3;
        const uint32_t t_row = tmem_base + buf * kBlk + lane_bits;
        const int nk = block_keys(un, j);
        mbar_wait(&v_full[vs], (g / kVSlots) & 1, 29);  // key bias / meta of the block visible
        if (j == 0) ATC_TRACE(u, EV_SM_VFULL);
        const float* bias = s_bias + vs * kBlk;
        int c_lo = 0, c_hi = nk, fast_end = s_meta[vs] & ~31;
        if (p.G != 0) { c_lo = w_lo; c_hi = min(w_hi, nk); fast_end = 0; }  // packed: every chunk takes the masked path
        if (!live) c_lo = c_hi = fast_end = 0;
        mbar_wait(&s_ready[4 * grp + buf], (s_par >> buf) & 1u, 30);
        s_par ^= 1u << buf;
        // PV(g - 1), a non-final block of this unit, completes pv_done[group][buffer of g - 1] (waited for only when a
        // raise of the reference has to rescale O); the parity advances with every non-final block this group hands over
        const int pbuf = (g - 1) & 3;
        const uint32_t pv_wait_par = (pv_par >> pbuf) & 1u;
        if (j > 0) pv_par ^= 1u << pbuf;
        tc_fence_after();
        if (j == 0) ATC_TRACE(u, EV_SM_START);
        // p = exp2(scale*s + bias - m_ref) with the lazy reference; softmax is shift-invariant, so O / sum is unchanged,
        // and P <= 2^10 stays far inside the fp16 range. [c_first, c_done) = keys of THIS block already written as P.
        auto raise_ref = [&](float cm, int c_first, int c_done) {  // warp-uniform call; cm = this lane's chunk max (log2 domain)
          const bool need = cm > m_ref + 10.0f;
          if (__any_sync(0xffffffffu, need)) {
            const float m_new = need ? ceilf(cm) : m_ref;
            const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);  // exact power of two, or 1
            sum *= f;
            const uint32_t f2 = pack16(f, f, p.bf16);
            if (c_done > c_first) tmem_st_wait();  // the P chunks stored so far must have landed before they are re-read
            for (int cc = c_first; cc < c_done; cc += 16) {  // P chunks of this block: 16 keys = 8 packed columns each
              uint32_t w[8];
              tmem_ld_32x8(t_row + (cc >> 1), w);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 8; ++i) w[i] = scale16(w[i], f2, p.bf16);
              tmem_st_32x8(t_row + (cc >> 1), w);
            }
            if (j > 0) {
              // earlier blocks of this unit are already in O under the old reference: wait for their PV MMAs (PV(u, j)
              // itself cannot have been issued: it waits for this warp's p_ready), then rescale this warp's O rows
              rescale_o_rows(t_o, f, &pv_done[4 * grp + pbuf], pv_wait_par);
            }
            m_ref = m_new;
          }
        };
        auto exp_chunk = [&](const uint32_t (&v)[32], int c) {  // 32 keys, all attended
          float c0 = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1])), c1 = fmaxf(__uint_as_float(v[2]), __uint_as_float(v[3]));
          float c2 = fmaxf(__uint_as_float(v[4]), __uint_as_float(v[5])), c3 = fmaxf(__uint_as_float(v[6]), __uint_as_float(v[7]));
#pragma unroll
          for (int i = 8; i < 32; i += 4) {  // four independent chains: short dependency depth
            c0 = fmaxf(c0, __uint_as_float(v[i]));
            c1 = fmaxf(c1, __uint_as_float(v[i + 1]));
            c2 = fmaxf(c2, __uint_as_float(v[i + 2]));
            c3 = fmaxf(c3, __uint_as_float(v[i + 3]));
          }
          raise_ref(fmaxf(fmaxf(c0, c1), fmaxf(c2, c3)) * p.scale_log2, c_lo, c);
          const float neg_m = -m_ref;
          uint32_t pk[16];
          const float2 sc2 = make_float2(p.scale_log2, p.scale_log2), nm2 = make_float2(neg_m, neg_m), one2 = make_float2(1.f, 1.f);
          float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            // packed fp32 (FFMA2): the scale-and-shift and the running sums of two keys per instruction
            const float2 t = __ffma2_rn(make_float2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
            const float2 e = make_float2(ex2f(t.x), ex2f(t.y));
            acc2 = __ffma2_rn(e, one2, acc2);
            pk[i] = pack16(e.x, e.y, p.bf16);
          }
          sum += acc2.x + acc2.y;
          tmem_st_32x16(t_row + (c >> 1), pk);
        };
        // TMEM loads are pipelined so that no tcgen05.wait::ld directly follows the load it would expose.
        // Invariant at the top of each step: `va` valid, `vb` (the next chunk) in flight since one chunk of work.
        uint32_t va[32], vb[32];
        if (fast_end > 0) {
          tmem_ld_32x32(t_row, va);
          tmem_ld_wait_dep(va);
          if (j == 0) ATC_TRACE(u, EV_SM_LD0);
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
        if (j == 0) ATC_TRACE(u, EV_SM_FAST);
        // Packed units: the PV MMA reads every key column of the block, so the columns of other warps' sequences get
        // P = 0. P (16-bit) lives in the lower half of the columns S occupied: the chunks below the window are zeroed
        // before it is read, the chunks above it only afterwards (their P columns may alias S columns of the window).
        const uint32_t zeros[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        if (p.G != 0 && live)
          for (int c = 0; c < c_lo; c += 16) tmem_st_32x8(t_row + (c >> 1), zeros);
        const int slow_begin = max(c_lo, fast_end);
        if (p.G != 0 && live) {
          // Packed sequences: the reference is the exact row maximum (rounded up to an integer), taken in a pre-pass over
          // the few columns of the window. A prompt's probabilities - and with them their 16-bit rounding - then do not
          // depend on which other prompts share its tile or on how far the batch was trimmed: encode_text of one prompt
          // equals its row in a batched call up to fp32 summation order.
          float mx = -INFINITY;
          for (int c = c_lo; c < c_hi; c += 16) {
            uint32_t v[16];
            tmem_ld_32x16(t_row + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const bool mine = (c + i >= my_lo) && (c + i < my_hi);
              mx = fmaxf(mx, mine ? fmaf(__uint_as_float(v[i]), p.scale_log2, bias[c + i]) : -INFINITY);
            }
          }
          m_ref = mx == -INFINITY ? -INFINITY : ceilf(mx);
        }
        for (int c = slow_begin; c < c_hi; c += 16) {  // chunks with masked keys (or foreign sequences): additive 0 / -inf bias
          uint32_t v[16];
          tmem_ld_32x16(t_row + c, v);
          tmem_ld_wait();
          float t[16];
          float cm = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool mine = (c + i >= my_lo) && (c + i < my_hi);
            t[i] = mine ? fmaf(__uint_as_float(v[i]), p.scale_log2, bias[c + i]) : -INFINITY;
            cm = fmaxf(cm, t[i]);
          }
          raise_ref(cm, c_lo, c);
          const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;  // everything masked so far: exp2(-inf) = 0
          uint32_t pk[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float e0 = ex2f(t[2 * i] + neg_m), e1 = ex2f(t[2 * i + 1] + neg_m);
            pk[i] = pack16(e0, e1, p.bf16);
            // the row sum is taken over the ROUNDED probabilities the PV MMA will see: the weights then sum to exactly one
            // (with a handful of keys the rounding of P would otherwise scale the whole context row by up to 2^-11)
            const float2 r = unpack16(pk[i], p.bf16);
            sum += r.x + r.y;
          }
          tmem_st_32x8(t_row + (c >> 1), pk);
        }
        if (p.G != 0 && live)
          for (int c = c_hi; c < nk; c += 16) tmem_st_32x8(t_row + (c >> 1), zeros);
        if (j == 0) ATC_TRACE(u, EV_SM_B0);
        if (j == nb - 1) s_rowsum[(u & 7) * 128 + row_in_tile] = sum;
        tmem_st_wait();
        if (j == 0) ATC_TRACE(u, EV_SM_STW);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (j == nb - 1) mbar_arrive(&sum_ready[u & 7]);
          mbar_arrive(&p_ready[4 * grp + buf]);
        }
      }
      ATC_TRACE(u, EV_SM_END);
    }
  } else if (warp < 4) {
    // ===================== output group: O / row sum -> context rows =====================
    const int q = warp & 3;
    const uint32_t t_o = tmem_base + kOCol + (uint32_t(q * 32) << 16);
    const int row_in_tile = q * 32 + lane;
    Walker w(p);
    for (int u = 0; u < U; ++u, w.next(p)) {
      const Unit un = w.unit(p);
      mbar_wait(&sum_ready[u & 7], (u >> 3) & 1, 32);  // row sums of this unit are visible
      const float sum = s_rowsum[(u & 7) * 128 + row_in_tile];
      mbar_wait(o_ready, u & 1, 33);
      tc_fence_after();
      ATC_TRACE(u, EV_OUT_START);
      uint32_t va[32], vb[32];
      if (q * 32 < un.q_rows) {  // warp-uniform: skip the tail warps of the last tile
        tmem_ld_32x32(t_o, va);
        tmem_ld_32x32(t_o + 32, vb);
        tmem_ld_wait_dep(va);
        tmem_ld_wait_dep(vb);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);  // O may be overwritten by the next unit's PV
      ATC_TRACE(u, EV_OUT_DONE);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      if (row_in_tile < un.q_rows) {
        uint16_t* orow = p.out + (long long)(un.q_row0 + row_in_tile) * p.out_pitch + un.h * 64;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(half == 0 ? va[i + e] : vb[i + e]) * inv;
            uint4 w;
            w.x = pack16(f[0], f[1], p.bf16);
            w.y = pack16(f[2], f[3], p.bf16);
            w.z = pack16(f[4], f[5], p.bf16);
            w.w = pack16(f[6], f[7], p.bf16);
            *reinterpret_cast<uint4*>(orow + 32 * half + i) = w;
            if (p.lo_off > 0) {  // rounding remainder of the context: the output projection then runs split-operand
              const float2 h0 = unpack16(w.x, p.bf16), h1 = unpack16(w.y, p.bf16), h2 = unpack16(w.z, p.bf16), h3 = unpack16(w.w, p.bf16);
              uint4 l;
              l.x = pack16(f[0] - h0.x, f[1] - h0.y, p.bf16);
              l.y = pack16(f[2] - h1.x, f[3] - h1.y, p.bf16);
              l.z = pack16(f[4] - h2.x, f[5] - h2.y, p.bf16);
              l.w = pack16(f[6] - h3.x, f[7] - h3.y, p.bf16);
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

static long long* g_atc_trace = nullptr;
void attention_tc_set_trace(long long* dev_buf) { g_atc_trace = dev_buf; }

int launch_attention(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                     int64_t mask_stride, float scale, cudaStream_t stream, int64_t out_pitch, int64_t lo_off) {
  if (B <= 0 || S <= 0 || H <= 0) return KB_OK;
  if (S > 512) return set_error(KB_ERR_ARG, "attention: S=%d > 512 unsupported", S);
  if (out_pitch <= 0) out_pitch = (int64_t)H * 64;
  if (out_pitch % 8 != 0 || lo_off % 8 != 0 || (lo_off > 0 && lo_off + (int64_t)H * 64 > out_pitch))
    return set_error(KB_ERR_ARG, "attention: output pitch %lld / lo offset %lld invalid", (long long)out_pitch, (long long)lo_off);
  // one kernel per shape class: sequences whose keys fit one 224-column tile (the ViT) run the single-tile kernel
  if (attention_tc1_supports(S))
    return launch_attention_tc1(qkv, out, B, S, H, bf16, key_mask, mask_stride, scale, stream, out_pitch, lo_off, g_atc_trace);
  AtcParams p;
  p.S = S; p.H = H; p.n_seq = B;
  const int S_pad = (S + 15) / 16 * 16;
  if (S <= kPackMaxS) {  // G >= 2 sequences per tile
    p.G = kBlk / S;
    const int groups = (B + p.G - 1) / p.G;
    p.items = groups * H; p.n_qt = 1; p.nb = 1;
    p.kbox = (p.G * S + 15) / 16 * 16;
  } else {
    p.G = 0;
    p.items = B * H; p.n_qt = (S + 127) / 128; p.nb = (S_pad + kBlk - 1) / kBlk;
    p.kbox = S_pad < kBlk ? S_pad : kBlk;
  }
  const int dt = bf16 ? KB_BF16 : KB_F16;
  const int64_t rows = (int64_t)B * S, cols = 3LL * H * 64;
  CUtensorMap tq, tkv;
  int rc = get_tmap_2d(qkv, dt, rows, cols, cols, 128, &tq);
  if (rc) return rc;
  rc = get_tmap_2d(qkv, dt, rows, cols, cols, p.kbox, &tkv);
  if (rc) return rc;
  const int smem = Smem::total + 1024;
  KB_TRY_ATTR(attention_tc_kernel, smem);
  p.key_mask = reinterpret_cast<const long long*>(key_mask);
  p.mask_stride = mask_stride;
  p.out = static_cast<uint16_t*>(out);
  p.out_pitch = out_pitch; p.lo_off = lo_off;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.fmt = bf16 ? kFmtBF16 : kFmtF16;
  p.idesc_pv = make_idesc(p.fmt, 128, 64, 0, 1);  // B operand (V) is MN-major: rows are keys, 64 head-dim values contiguous
  p.bf16 = bf16;
  p.trace = g_atc_trace;
  int grid = num_sms();
  if (p.items < grid) grid = p.items;
  attention_tc_kernel<<<grid, kAtcThreads, smem, stream>>>(tq, tkv, p);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
