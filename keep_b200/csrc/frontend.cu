// Input-side kernels (HBM-bound gathers):
//   * ViT patch gather: tiles -> 16-bit patch matrix [B*G*G, 3*16*16] whose columns follow the flattened
//     Conv2d weight [D,3,16,16] (timm PatchEmbed), so patch-embed becomes one GEMM;
//     the CLS rows x[b,0,:] = cls_token + pos_embed[0] are written here as well (timm _pos_embed).
//     A uint8 NHWC variant fuses ToTensor + Normalize(mean,std) (quick_start/keep_inference.py:91-92).
//   * BERT embeddings: word[ids] + token_type[tt] + position[s] -> LayerNorm (BertEmbeddings,
//     transformers modeling_bert.py:53-112).
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

__device__ __forceinline__ float2 unpk(uint32_t v, int bf16) {
  if (bf16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
  return __half22float2(*reinterpret_cast<__half2*>(&v));
}
__device__ __forceinline__ uint32_t pk(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// one thread = 8 consecutive pixels of one image row of one channel
__global__ void im2col_f32_kernel(const float* __restrict__ tiles, long long total, int Gh, int G, uint16_t* __restrict__ patches,
                                  int bf16, int pitch, int lo_off) {
  const int W = G * 16, H = Gh * 16, X8 = W / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int x8 = (int)(i % X8);
    long long t = i / X8;
    const int y = (int)(t % H); t /= H;
    const int c = (int)(t % 3);
    const long long b = t / 3;
    const float4* src = reinterpret_cast<const float4*>(tiles + ((b * 3 + c) * H + y) * W + x8 * 8);
    const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1);
    const int x = x8 * 8, px = x >> 4, kx = x & 15, py = y >> 4, ky = y & 15;
    const long long row = b * Gh * G + py * G + px;
    uint4 w;
    w.x = pk(v0.x, v0.y, bf16); w.y = pk(v0.z, v0.w, bf16);
    w.z = pk(v1.x, v1.y, bf16); w.w = pk(v1.z, v1.w, bf16);
    uint16_t* dst = patches + row * pitch + c * 256 + ky * 16 + kx;
    *reinterpret_cast<uint4*>(dst) = w;
    if (lo_off > 0) {  // rounding remainders: the patch matrix becomes the hi|lo operand of a split-operand patch embedding
      const float2 h0 = unpk(w.x, bf16), h1 = unpk(w.y, bf16), h2 = unpk(w.z, bf16), h3 = unpk(w.w, bf16);
      uint4 l;
      l.x = pk(v0.x - h0.x, v0.y - h0.y, bf16); l.y = pk(v0.z - h1.x, v0.w - h1.y, bf16);
      l.z = pk(v1.x - h2.x, v1.y - h2.y, bf16); l.w = pk(v1.z - h3.x, v1.w - h3.y, bf16);
      *reinterpret_cast<uint4*>(dst + lo_off) = l;
    }
  }
}

// one thread = 8 consecutive pixels (24 bytes, all 3 channels) of one image row
__global__ void im2col_u8_kernel(const uint8_t* __restrict__ tiles, long long total, int Gh, int G, uint16_t* __restrict__ patches,
                                 int bf16, int pitch, int lo_off) {
  const int W = G * 16, H = Gh * 16, X8 = W / 8;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float istd[3] = {1.f / 0.229f, 1.f / 0.224f, 1.f / 0.225f};
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int x8 = (int)(i % X8);
    long long t = i / X8;
    const int y = (int)(t % H);
    const long long b = t / H;
    const uint2* src = reinterpret_cast<const uint2*>(tiles + ((b * H + y) * W + x8 * 8) * 3);  // 24 B, 8-aligned
    uint2 raw[3] = {src[0], src[1], src[2]};
    const uint8_t* px8 = reinterpret_cast<const uint8_t*>(raw);
    const int x = x8 * 8, px = x >> 4, kx = x & 15, py = y >> 4, ky = y & 15;
    const long long row = b * Gh * G + py * G + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (px8[j * 3 + c] * (1.f / 255.f) - mean[c]) * istd[c];
      uint4 w;
      w.x = pk(f[0], f[1], bf16); w.y = pk(f[2], f[3], bf16);
      w.z = pk(f[4], f[5], bf16); w.w = pk(f[6], f[7], bf16);
      uint16_t* dst = patches + row * pitch + c * 256 + ky * 16 + kx;
      *reinterpret_cast<uint4*>(dst) = w;
      if (lo_off > 0) {
        const float2 h0 = unpk(w.x, bf16), h1 = unpk(w.y, bf16), h2 = unpk(w.z, bf16), h3 = unpk(w.w, bf16);
        uint4 l;
        l.x = pk(f[0] - h0.x, f[1] - h0.y, bf16); l.y = pk(f[2] - h1.x, f[3] - h1.y, bf16);
        l.z = pk(f[4] - h2.x, f[5] - h2.y, bf16); l.w = pk(f[6] - h3.x, f[7] - h3.y, bf16);
        *reinterpret_cast<uint4*>(dst + lo_off) = l;
      }
    }
  }
}

// dynamic_img_size (quick_start/keep_inference.py:39 -> timm resample_abs_pos_embed): the G0 x G0 grid part of
// pos_embed is resampled to Gh x Gw with F.interpolate(mode="bicubic", antialias=True) semantics, the prefix (CLS) row is
// copied. Same weights as ATen's anti-aliased kernels: cubic a = -0.5, support = 2 * max(scale, 1), window
// [int(center - support + 0.5), int(center + support + 0.5)) clipped to the input, weights normalised to sum 1;
// separable, horizontal pass inside the vertical one. One thread = one output token x 4 channels.
__device__ __forceinline__ float cubic_aa(float x) {
  const float a = -0.5f;
  x = fabsf(x);
  if (x < 1.0f) return ((a + 2.0f) * x - (a + 3.0f)) * x * x + 1.0f;
  if (x < 2.0f) return (((x - 5.0f) * x + 8.0f) * x - 4.0f) * a;
  return 0.0f;
}
__device__ __forceinline__ void aa_span(int o, int in_size, float scale, float support, int* xmin, int* xsize, float* center) {
  *center = scale * (o + 0.5f);
  int lo = (int)(*center - support + 0.5f);
  if (lo < 0) lo = 0;
  int hi = (int)(*center + support + 0.5f);
  if (hi > in_size) hi = in_size;
  *xmin = lo;
  *xsize = hi - lo;
}
__global__ void pos_resample_kernel(const float* __restrict__ pos, int G0, int Gh, int Gw, int D, float* __restrict__ out) {
  const int d4 = D / 4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)(1 + Gh * Gw) * d4) return;
  const int c = (int)(i % d4);
  const int tok = (int)(i / d4);
  const float4* src = reinterpret_cast<const float4*>(pos);
  float4* dst = reinterpret_cast<float4*>(out);
  if (tok == 0) { dst[c] = src[c]; return; }
  const int oy = (tok - 1) / Gw, ox = (tok - 1) % Gw;
  const float sy = (float)G0 / (float)Gh, sx = (float)G0 / (float)Gw;
  const float sup_y = sy >= 1.0f ? 2.0f * sy : 2.0f, sup_x = sx >= 1.0f ? 2.0f * sx : 2.0f;
  const float inv_y = sy >= 1.0f ? 1.0f / sy : 1.0f, inv_x = sx >= 1.0f ? 1.0f / sx : 1.0f;
  int ymin, ysize, xmin, xsize;
  float cy, cx;
  aa_span(oy, G0, sy, sup_y, &ymin, &ysize, &cy);
  aa_span(ox, G0, sx, sup_x, &xmin, &xsize, &cx);
  float tot_y = 0.f, tot_x = 0.f;
  for (int j = 0; j < ysize; ++j) tot_y += cubic_aa((j + ymin - cy + 0.5f) * inv_y);
  for (int j = 0; j < xsize; ++j) tot_x += cubic_aa((j + xmin - cx + 0.5f) * inv_x);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int jy = 0; jy < ysize; ++jy) {
    float wy = cubic_aa((jy + ymin - cy + 0.5f) * inv_y);
    if (tot_y != 0.f) wy /= tot_y;
    float4 row = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int jx = 0; jx < xsize; ++jx) {
      float wx = cubic_aa((jx + xmin - cx + 0.5f) * inv_x);
      if (tot_x != 0.f) wx /= tot_x;
      const float4 v = __ldg(src + (long long)(1 + (ymin + jy) * G0 + (xmin + jx)) * d4 + c);
      row.x = fmaf(wx, v.x, row.x); row.y = fmaf(wx, v.y, row.y); row.z = fmaf(wx, v.z, row.z); row.w = fmaf(wx, v.w, row.w);
    }
    acc.x = fmaf(wy, row.x, acc.x); acc.y = fmaf(wy, row.y, acc.y); acc.z = fmaf(wy, row.z, acc.z); acc.w = fmaf(wy, row.w, acc.w);
  }
  dst[(long long)tok * d4 + c] = acc;
}

__global__ void cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                                long long B, int T, int D) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int d = (int)(i % D);
  const long long b = i / D;
  x[b * T * D + d] = cls[d] + pos[d];
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int NV>
__global__ void __launch_bounds__(256)
bert_embed_kernel(const long long* __restrict__ ids, const long long* __restrict__ tts, long long id_stride,
                  long long rows, int S, const float* __restrict__ word, const float* __restrict__ type,
                  const float* __restrict__ pos, const float* __restrict__ lnw, const float* __restrict__ lnb, float eps,
                  float* __restrict__ x32, uint16_t* __restrict__ x16, int bf16, int vocab, int type_vocab,
                  long long x16_pitch, long long lo_off) {
  constexpr int D = NV * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long p = row / S;
  const int s = (int)(row % S);
  long long id = ids[p * id_stride + s];
  long long tt = tts ? tts[p * id_stride + s] : 0;
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  tt = tt < 0 ? 0 : (tt >= type_vocab ? type_vocab - 1 : tt);
  const float4* w4 = reinterpret_cast<const float4*>(word + id * D);
  const float4* t4 = reinterpret_cast<const float4*>(type + tt * D);
  const float4* p4 = reinterpret_cast<const float4*>(pos + (long long)s * D);
  float4 v[NV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 a = __ldg(w4 + lane + 32 * i), b = __ldg(t4 + lane + 32 * i), c = __ldg(p4 + lane + 32 * i);
    // same association as BertEmbeddings: (word + token_type) + position
    v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(sum) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 ww = __ldg(reinterpret_cast<const float4*>(lnw) + lane + 32 * i);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(lnb) + lane + 32 * i);
    float4 o;
    o.x = v[i].x * rstd * ww.x + bb.x;
    o.y = v[i].y * rstd * ww.y + bb.y;
    o.z = v[i].z * rstd * ww.z + bb.z;
    o.w = v[i].w * rstd * ww.w + bb.w;
    *reinterpret_cast<float4*>(x32 + row * D + (lane + 32 * i) * 4) = o;
    uint2 h;
    h.x = pk(o.x, o.y, bf16);
    h.y = pk(o.z, o.w, bf16);
    *reinterpret_cast<uint2*>(x16 + row * x16_pitch + (lane + 32 * i) * 4) = h;
    if (lo_off > 0) {  // rounding remainder for the split-operand GEMMs
      const float2 a = unpk(h.x, bf16), b = unpk(h.y, bf16);
      uint2 l;
      l.x = pk(o.x - a.x, o.y - a.y, bf16);
      l.y = pk(o.z - b.x, o.w - b.y, bf16);
      *reinterpret_cast<uint2*>(x16 + row * x16_pitch + lo_off + (lane + 32 * i) * 4) = l;
    }
  }
}

}  // namespace

static int launch_cls_rows(const float* cls, const float* pos, float* x, int64_t B, int T, int D, cudaStream_t stream) {
  const long long n = (long long)B * D;
  cls_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(cls, pos, x, B, T, D);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_pos_resample(const float* pos, int G0, int Gh, int Gw, int D, float* out, cudaStream_t stream) {
  if (G0 <= 0 || Gh <= 0 || Gw <= 0 || D % 4 != 0) return set_error(KB_ERR_ARG, "pos_resample: bad shape");
  const long long n = (long long)(1 + Gh * Gw) * (D / 4);
  pos_resample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(pos, G0, Gh, Gw, D, out);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_im2col(const float* tiles, int64_t B, int Gh, int G, void* patches16, int bf16, const float* cls,
                  const float* pos, float* x, int D, cudaStream_t stream, int hilo) {
  if (B <= 0) return KB_OK;
  const int W = G * 16;
  const long long total = (long long)B * 3 * (Gh * 16) * (W / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  im2col_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tiles, total, Gh, G, (uint16_t*)patches16, bf16, hilo ? 1536 : 768,
                                                          hilo ? 768 : 0);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return launch_cls_rows(cls, pos, x, B, Gh * G + 1, D, stream);
}

int launch_im2col_u8(const uint8_t* tiles, int64_t B, int Gh, int G, void* patches16, int bf16, const float* cls,
                     const float* pos, float* x, int D, cudaStream_t stream, int hilo) {
  if (B <= 0) return KB_OK;
  const int W = G * 16;
  const long long total = (long long)B * (Gh * 16) * (W / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  im2col_u8_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tiles, total, Gh, G, (uint16_t*)patches16, bf16, hilo ? 1536 : 768,
                                                         hilo ? 768 : 0);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return launch_cls_rows(cls, pos, x, B, Gh * G + 1, D, stream);
}

int launch_bert_embed(const int64_t* ids, const int64_t* tts, int64_t id_stride, int64_t P, int S, int D,
                      const float* word, const float* type, const float* pos, const float* lnw, const float* lnb,
                      float eps, float* x32, void* x16, int bf16, int vocab, int type_vocab, cudaStream_t stream,
                      int64_t x16_pitch, int64_t lo_off) {
  const long long rows = (long long)P * S;
  if (rows <= 0) return KB_OK;
  if (D % 128 != 0 || D > 1024) return set_error(KB_ERR_ARG, "bert_embed: hidden=%d unsupported", D);
  if (x16_pitch <= 0) x16_pitch = D;
  if (x16_pitch % 4 != 0 || lo_off % 4 != 0 || (lo_off > 0 && lo_off + D > x16_pitch))
    return set_error(KB_ERR_ARG, "bert_embed: 16-bit output pitch / lo offset invalid");
  const unsigned grid = (unsigned)((rows + 7) / 8);
#define KB_EMB(NVV)                                                                                              \
  bert_embed_kernel<NVV><<<grid, 256, 0, stream>>>((const long long*)ids, (const long long*)tts, id_stride, rows, S, \
                                                   word, type, pos, lnw, lnb, eps, x32, (uint16_t*)x16, bf16, vocab,  \
                                                   type_vocab, x16_pitch, lo_off)
  switch (D / 128) {
    case 1: KB_EMB(1); break;
    case 2: KB_EMB(2); break;
    case 3: KB_EMB(3); break;
    case 4: KB_EMB(4); break;
    case 5: KB_EMB(5); break;
    case 6: KB_EMB(6); break;
    case 7: KB_EMB(7); break;
    default: KB_EMB(8); break;
  }
#undef KB_EMB
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
