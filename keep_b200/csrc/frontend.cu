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

__device__ __forceinline__ uint32_t pk(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// one thread = 8 consecutive pixels of one image row of one channel
__global__ void im2col_f32_kernel(const float* __restrict__ tiles, long long total, int G, uint16_t* __restrict__ patches,
                                  int bf16) {
  const int W = G * 16, X8 = W / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int x8 = (int)(i % X8);
    long long t = i / X8;
    const int y = (int)(t % W); t /= W;
    const int c = (int)(t % 3);
    const long long b = t / 3;
    const float4* src = reinterpret_cast<const float4*>(tiles + ((b * 3 + c) * W + y) * W + x8 * 8);
    const float4 v0 = __ldcs(src), v1 = __ldcs(src + 1);
    const int x = x8 * 8, px = x >> 4, kx = x & 15, py = y >> 4, ky = y & 15;
    const long long row = b * G * G + py * G + px;
    uint4 w;
    w.x = pk(v0.x, v0.y, bf16); w.y = pk(v0.z, v0.w, bf16);
    w.z = pk(v1.x, v1.y, bf16); w.w = pk(v1.z, v1.w, bf16);
    *reinterpret_cast<uint4*>(patches + row * 768 + c * 256 + ky * 16 + kx) = w;
  }
}

// one thread = 8 consecutive pixels (24 bytes, all 3 channels) of one image row
__global__ void im2col_u8_kernel(const uint8_t* __restrict__ tiles, long long total, int G, uint16_t* __restrict__ patches,
                                 int bf16) {
  const int W = G * 16, X8 = W / 8;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float istd[3] = {1.f / 0.229f, 1.f / 0.224f, 1.f / 0.225f};
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < total; i += stride) {
    const int x8 = (int)(i % X8);
    long long t = i / X8;
    const int y = (int)(t % W);
    const long long b = t / W;
    const uint2* src = reinterpret_cast<const uint2*>(tiles + ((b * W + y) * W + x8 * 8) * 3);  // 24 B, 8-aligned
    uint2 raw[3] = {src[0], src[1], src[2]};
    const uint8_t* px8 = reinterpret_cast<const uint8_t*>(raw);
    const int x = x8 * 8, px = x >> 4, kx = x & 15, py = y >> 4, ky = y & 15;
    const long long row = b * G * G + py * G + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = (px8[j * 3 + c] * (1.f / 255.f) - mean[c]) * istd[c];
      uint4 w;
      w.x = pk(f[0], f[1], bf16); w.y = pk(f[2], f[3], bf16);
      w.z = pk(f[4], f[5], bf16); w.w = pk(f[6], f[7], bf16);
      *reinterpret_cast<uint4*>(patches + row * 768 + c * 256 + ky * 16 + kx) = w;
    }
  }
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
                  float* __restrict__ x32, uint16_t* __restrict__ x16, int bf16, int vocab, int type_vocab) {
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
    *reinterpret_cast<uint2*>(x16 + row * D + (lane + 32 * i) * 4) = h;
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

int launch_im2col(const float* tiles, int64_t B, int G, void* patches16, int bf16, const float* cls,
                  const float* pos, float* x, int D, cudaStream_t stream) {
  if (B <= 0) return KB_OK;
  const int W = G * 16;
  const long long total = (long long)B * 3 * W * (W / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  im2col_f32_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tiles, total, G, (uint16_t*)patches16, bf16);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return launch_cls_rows(cls, pos, x, B, G * G + 1, D, stream);
}

int launch_im2col_u8(const uint8_t* tiles, int64_t B, int G, void* patches16, int bf16, const float* cls,
                     const float* pos, float* x, int D, cudaStream_t stream) {
  if (B <= 0) return KB_OK;
  const int W = G * 16;
  const long long total = (long long)B * W * (W / 8);
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 32;
  if (blocks > cap) blocks = cap;
  im2col_u8_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tiles, total, G, (uint16_t*)patches16, bf16);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return launch_cls_rows(cls, pos, x, B, G * G + 1, D, stream);
}

int launch_bert_embed(const int64_t* ids, const int64_t* tts, int64_t id_stride, int64_t P, int S, int D,
                      const float* word, const float* type, const float* pos, const float* lnw, const float* lnb,
                      float eps, float* x32, void* x16, int bf16, int vocab, int type_vocab, cudaStream_t stream) {
  const long long rows = (long long)P * S;
  if (rows <= 0) return KB_OK;
  if (D % 128 != 0 || D > 1024) return set_error(KB_ERR_ARG, "bert_embed: hidden=%d unsupported", D);
  const unsigned grid = (unsigned)((rows + 7) / 8);
#define KB_EMB(NVV)                                                                                              \
  bert_embed_kernel<NVV><<<grid, 256, 0, stream>>>((const long long*)ids, (const long long*)tts, id_stride, rows, S, \
                                                   word, type, pos, lnw, lnb, eps, x32, (uint16_t*)x16, bf16, vocab,  \
                                                   type_vocab)
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
