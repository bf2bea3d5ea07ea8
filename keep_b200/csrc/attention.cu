// Multi-head attention for short sequences (ViT: 197 tokens, BERT: <= 512), head dim 64, non-causal,
// optional key-padding mask:  out = softmax(Q K^T * scale + mask) V   per (batch, head).
//
// Reference semantics: timm Attention -> F.scaled_dot_product_attention(q, k, v) with scale 1/8
//  and BertSelfAttention with the additive key mask built from attention_mask
// (transformers modeling_bert.py:115-140).
//
// Layout: q|k|v are read in place from the fused projection output [B*S, 3*H*64] (no head-split copy);
// the context is written as [B*S, H*64], i.e. already in the layout the output projection consumes.
//
// v1 structure (register-resident flash attention on the warp-level tensor-core path): a CTA owns one
// (batch, head) and 128 query rows; all K/V rows of that head are staged once in XOR-swizzled shared
// memory with cp.async; each warp owns 16 query rows, walks the keys in blocks of 64 with an fp32 online
// softmax, and keeps P in registers between the two MMAs.
#include "common.h"
#include "ptx.cuh"


namespace kb {
namespace {

constexpr int DH = 64;
constexpr int QB = 128;  // query rows per CTA
constexpr int KB_ = 64;  // keys per inner block
constexpr int kAttThreads = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const uint32_t s = smem_u32(smem);
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (BF16) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  } else {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
}
template <bool BF16>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  if constexpr (BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  } else {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
}

template <bool BF16>
__device__ __forceinline__ float2 unpack2(uint32_t v) {
  if constexpr (BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  else return __half22float2(*reinterpret_cast<__half2*>(&v));
}

// row r, 16-byte chunk c (0..7) of a [rows][64 x 16-bit] tile, XOR-swizzled against bank conflicts
__device__ __forceinline__ uint32_t sw_off(int r, int c) { return uint32_t(r) * 128u + uint32_t((c ^ (r & 7)) << 4); }

template <bool BF16>
__global__ void __launch_bounds__(kAttThreads)
attention_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int S, int H,
                 const long long* __restrict__ key_mask, long long mask_stride, float scale_log2, long long ldo,
                 long long lo_off) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int spad = (S + KB_ - 1) / KB_ * KB_;
  uint8_t* sK = smem;
  uint8_t* sV = sK + spad * 128;
  uint8_t* sQ = sV + spad * 128;
  float* sBias = reinterpret_cast<float*>(sQ + QB * 128);  // [spad]: 0 or -inf per key

  const int nqb = (S + QB - 1) / QB;
  const int qb = blockIdx.x % nqb, h = (blockIdx.x / nqb) % H, b = blockIdx.x / (nqb * H);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long ld = 3LL * H * DH;
  const uint16_t* base = qkv + (long long)b * S * ld;
  const int q0 = qb * QB;

  // ---- stage K, V (all keys) and this CTA's Q rows ----
  for (int i = tid; i < spad * 8; i += kAttThreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = r < S;
    const uint16_t* src = base + (long long)(ok ? r : 0) * ld + h * DH + c * 8;
    cp_async16(sK + sw_off(r, c), src + (long long)H * DH, ok);
    cp_async16(sV + sw_off(r, c), src + 2LL * H * DH, ok);
  }
  for (int i = tid; i < QB * 8; i += kAttThreads) {
    const int r = i >> 3, c = i & 7;
    const bool ok = (q0 + r) < S;
    const uint16_t* src = base + (long long)(ok ? q0 + r : 0) * ld + h * DH + c * 8;
    cp_async16(sQ + sw_off(r, c), src, ok);
  }
  for (int i = tid; i < spad; i += kAttThreads) {
    bool ok = i < S;
    if (ok && key_mask != nullptr) ok = key_mask[(long long)b * mask_stride + i] != 0;
    sBias[i] = ok ? 0.f : -INFINITY;
  }
  cp_async_wait_all();
  __syncthreads();

  const int r0 = q0 + warp * 16;
  if (r0 >= S) return;  // no further block-level sync below

  // ---- Q fragments: 4 k-steps of m16k16 ----
  uint32_t qf[4][4];
  {
    const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int c = ks * 2 + (lane >> 4);
      ldsm_x4(smem_u32(sQ + sw_off(r, c)), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
  }

  float o[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int k0 = 0; k0 < spad; k0 += KB_) {
    // S = Q K^T for 64 keys
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of n8 tiles (16 keys)
        const int key = k0 + np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int c = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(smem_u32(sK + sw_off(key, c)), b0, b1, b2, b3);
        mma16816<BF16>(s[np * 2], qf[ks], b0, b1);
        mma16816<BF16>(s[np * 2 + 1], qf[ks], b2, b3);
      }
    }
    // scale (log2 domain) + key bias, block row max
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int key = k0 + nt * 8 + (lane & 3) * 2;
      const float b0 = sBias[key], b1 = sBias[key + 1];
      s[nt][0] = s[nt][0] * scale_log2 + b0;
      s[nt][1] = s[nt][1] * scale_log2 + b1;
      s[nt][2] = s[nt][2] * scale_log2 + b0;
      s[nt][3] = s[nt][3] * scale_log2 + b1;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2], m_use[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 1));
      mx[i] = fmaxf(mx[i], __shfl_xor_sync(0xffffffffu, mx[i], 2));
      const float m_new = fmaxf(m_run[i], mx[i]);
      m_use[i] = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: avoid (-inf) - (-inf)
      corr[i] = exp2f(m_run[i] - m_use[i]);
      m_run[i] = m_new;
      l_run[i] *= corr[i];
    }
    float ls[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = exp2f(s[nt][0] - m_use[0]);
      s[nt][1] = exp2f(s[nt][1] - m_use[0]);
      s[nt][2] = exp2f(s[nt][2] - m_use[1]);
      s[nt][3] = exp2f(s[nt][3] - m_use[1]);
      ls[0] += s[nt][0] + s[nt][1];
      ls[1] += s[nt][2] + s[nt][3];
    }
    l_run[0] += ls[0];
    l_run[1] += ls[1];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      o[nt][0] *= corr[0]; o[nt][1] *= corr[0];
      o[nt][2] *= corr[1]; o[nt][3] *= corr[1];
    }
    // O += P V
#pragma unroll
    for (int j = 0; j < 4; ++j) {  // 16 keys per step
      uint32_t pa[4];
      pa[0] = pack2<BF16>(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack2<BF16>(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack2<BF16>(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack2<BF16>(s[2 * j + 1][2], s[2 * j + 1][3]);
      const int key = k0 + j * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of dh n8 tiles
        const int c = np * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(smem_u32(sV + sw_off(key, c)), b0, b1, b2, b3);
        mma16816<BF16>(o[np * 2], pa, b0, b1);
        mma16816<BF16>(o[np * 2 + 1], pa, b2, b3);
      }
    }
  }

  // ---- finalise: O / l, write [B*S, H*64] ----
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 1);
    l_run[i] += __shfl_xor_sync(0xffffffffu, l_run[i], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  const int row_a = r0 + (lane >> 2), row_b = row_a + 8;
  uint16_t* ob = out + (long long)b * S * ldo + h * DH + (lane & 3) * 2;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const float a0 = o[nt][0] * inv0, a1 = o[nt][1] * inv0, b0 = o[nt][2] * inv1, b1 = o[nt][3] * inv1;
    const uint32_t ha = pack2<BF16>(a0, a1), hb = pack2<BF16>(b0, b1);
    if (row_a < S) *reinterpret_cast<uint32_t*>(ob + row_a * ldo + nt * 8) = ha;
    if (row_b < S) *reinterpret_cast<uint32_t*>(ob + row_b * ldo + nt * 8) = hb;
    if (lo_off > 0) {  // rounding remainder of the context: the output projection then runs as a split-operand GEMM
      const float2 fa = unpack2<BF16>(ha), fb = unpack2<BF16>(hb);
      if (row_a < S) *reinterpret_cast<uint32_t*>(ob + row_a * ldo + lo_off + nt * 8) = pack2<BF16>(a0 - fa.x, a1 - fa.y);
      if (row_b < S) *reinterpret_cast<uint32_t*>(ob + row_b * ldo + lo_off + nt * 8) = pack2<BF16>(b0 - fb.x, b1 - fb.y);
    }
  }
}

}  // namespace

int launch_attention(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                     int64_t mask_stride, float scale, cudaStream_t stream, int64_t out_pitch, int64_t lo_off) {
  if (B <= 0 || S <= 0 || H <= 0) return KB_OK;
  if (out_pitch <= 0) out_pitch = (int64_t)H * DH;
  if (out_pitch % 8 != 0 || lo_off % 8 != 0 || (lo_off > 0 && lo_off + (int64_t)H * DH > out_pitch))
    return set_error(KB_ERR_ARG, "attention: output pitch %lld / lo offset %lld invalid", (long long)out_pitch, (long long)lo_off);
  if (attention_tc_supports(S) && lo_off == 0 && out_pitch == (int64_t)H * DH)
    return launch_attention_tc(qkv, out, B, S, H, bf16, key_mask, mask_stride, scale, stream);
  if (S > 512) return set_error(KB_ERR_ARG, "attention: S=%d > 512 unsupported", S);
  const int spad = (S + KB_ - 1) / KB_ * KB_;
  const int smem = (2 * spad + QB) * 128 + spad * 4;
  const unsigned grid = (unsigned)((S + QB - 1) / QB) * H * B;
  const float scale_log2 = scale * 1.4426950408889634f;
  if (bf16) {
    KB_TRY_ATTR(attention_kernel<true>, smem);
    attention_kernel<true><<<grid, kAttThreads, smem, stream>>>((const uint16_t*)qkv, (uint16_t*)out, S, H,
                                                                (const long long*)key_mask, mask_stride, scale_log2,
                                                                out_pitch, lo_off);
  } else {
    KB_TRY_ATTR(attention_kernel<false>, smem);
    attention_kernel<false><<<grid, kAttThreads, smem, stream>>>((const uint16_t*)qkv, (uint16_t*)out, S, H,
                                                                 (const long long*)key_mask, mask_stride, scale_log2,
                                                                 out_pitch, lo_off);
  }
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
