// HBM-bound row kernels: LayerNorm (fp32 in, 16-bit and/or fp32 out), activation + L2-normalise,
// fp32 -> 16-bit cast, fp32 transpose. One warp owns one row; the row lives in registers between the
// statistics pass and the write, so every element is read once and written once.
//
// Reference semantics: torch.nn.LayerNorm (biased variance, eps inside the sqrt) as used by timm's ViT
// blocks (eps 1e-6) and BertSelfOutput/BertOutput/BertEmbeddings (eps 1e-12; §3.4);
// F.normalize(dim=-1) = x / max(||x||_2, 1e-12) (quick_start/keep_inference.py:55,61).
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint2 pack4(float4 v, int bf16) {
  uint2 r;
  if (bf16) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
  } else {
    __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    r.x = *reinterpret_cast<uint32_t*>(&a);
    r.y = *reinterpret_cast<uint32_t*>(&b);
  }
  return r;
}
// rounding remainder of pack4: lo = 16-bit(v - hi)
__device__ __forceinline__ uint2 pack4_lo(float4 v, uint2 hi, int bf16) {
  float4 h;
  if (bf16) {
    const float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&hi.x)), b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&hi.y));
    h = make_float4(a.x, a.y, b.x, b.y);
  } else {
    const float2 a = __half22float2(*reinterpret_cast<__half2*>(&hi.x)), b = __half22float2(*reinterpret_cast<__half2*>(&hi.y));
    h = make_float4(a.x, a.y, b.x, b.y);
  }
  return pack4(make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w), bf16);
}

// NV = D / 128 float4 per lane
template <int NV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* x, long long row_stride,
                                                        long long rows, const float* __restrict__ w,
                                                        const float* __restrict__ b, float eps, uint16_t* y16, int bf16,
                                                        float* y32, long long y16_pitch, long long lo_off) {
  constexpr int D = NV * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * row_stride);
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 ww = __ldg(w4 + lane + 32 * i), bb = __ldg(b4 + lane + 32 * i);
    float4 o;
    o.x = v[i].x * rstd * ww.x + bb.x;
    o.y = v[i].y * rstd * ww.y + bb.y;
    o.z = v[i].z * rstd * ww.z + bb.z;
    o.w = v[i].w * rstd * ww.w + bb.w;
    if (y16) {
      const uint2 hi = pack4(o, bf16);
      *reinterpret_cast<uint2*>(y16 + row * y16_pitch + (lane + 32 * i) * 4) = hi;
      if (lo_off > 0) *reinterpret_cast<uint2*>(y16 + row * y16_pitch + lo_off + (lane + 32 * i) * 4) = pack4_lo(o, hi, bf16);
    }
    if (y32) *reinterpret_cast<float4*>(y32 + row * D + (lane + 32 * i) * 4) = o;
  }
}

template <int NV>
__global__ void __launch_bounds__(256) act_l2norm_kernel(const float* __restrict__ x, long long rows, int act,
                                                         float* __restrict__ y) {
  constexpr int D = NV * 128;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float4* xr = reinterpret_cast<const float4*>(x + row * D);
  float4 v[NV];
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = xr[lane + 32 * i];
    if (act == 1) { v[i].x = tanhf(v[i].x); v[i].y = tanhf(v[i].y); v[i].z = tanhf(v[i].z); v[i].w = tanhf(v[i].w); }
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float inv = 1.0f / fmaxf(sqrtf(warp_sum(q)), 1e-12f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float4 o = make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv);
    *reinterpret_cast<float4*>(y + row * D + (lane + 32 * i) * 4) = o;
  }
}

// dst[r, 0:K] = 16-bit(src[r, :]), dst[r, K:2K] = 16-bit(src - hi): the [rows, 2K] hi|lo operand of a split GEMM
__global__ void cast_hilo_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long rows, int K4, int bf16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x, n4 = rows * K4;
  for (; i < n4; i += stride) {
    const long long r = i / K4;
    const int c = (int)(i % K4);
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    const uint2 hi = pack4(v, bf16);
    uint2* drow = reinterpret_cast<uint2*>(dst + r * (8LL * K4));
    drow[c] = hi;
    drow[K4 + c] = pack4_lo(v, hi, bf16);
  }
}

__global__ void cast_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, long long n4, int bf16) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    reinterpret_cast<uint2*>(dst)[i] = pack4(v, bf16);
  }
}

// One warp per output row n of W[N,K]: W'[n,:] = 16-bit(W[n,:] * lnw), s[n] = sum of the rounded W'[n,:],
// c[n] = bias[n] + <lnb, W[n,:]>  (LayerNorm folded into the following Linear; see EPI_LN_* in common.h)
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W, int N, int K,
                                                      const float* __restrict__ lnw, const float* __restrict__ lnb,
                                                      const float* __restrict__ bias, uint16_t* __restrict__ W16,
                                                      int bf16, float* __restrict__ s, float* __restrict__ c) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n >= N) return;
  const int lane = threadIdx.x & 31;
  const float4* wr = reinterpret_cast<const float4*>(W + (long long)n * K);
  float ssum = 0.f, csum = 0.f;
  for (int k4 = lane; k4 < K / 4; k4 += 32) {
    const float4 w = wr[k4];
    const float4 g = __ldg(reinterpret_cast<const float4*>(lnw) + k4), b = __ldg(reinterpret_cast<const float4*>(lnb) + k4);
    const float4 f = make_float4(w.x * g.x, w.y * g.y, w.z * g.z, w.w * g.w);
    const uint2 pk = pack4(f, bf16);
    *reinterpret_cast<uint2*>(W16 + (long long)n * K + k4 * 4) = pk;
    float r0, r1, r2, r3;
    if (bf16) {
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&pk.x), d = *reinterpret_cast<const __nv_bfloat162*>(&pk.y);
      r0 = __low2float(a); r1 = __high2float(a); r2 = __low2float(d); r3 = __high2float(d);
    } else {
      const __half2 a = *reinterpret_cast<const __half2*>(&pk.x), d = *reinterpret_cast<const __half2*>(&pk.y);
      r0 = __low2float(a); r1 = __high2float(a); r2 = __low2float(d); r3 = __high2float(d);
    }
    ssum += (r0 + r1) + (r2 + r3);
    csum += (w.x * b.x + w.y * b.y) + (w.z * b.z + w.w * b.w);
  }
  ssum = warp_sum(ssum);
  csum = warp_sum(csum);
  if (lane == 0) {
    s[n] = ssum;
    c[n] = csum + (bias != nullptr ? bias[n] : 0.f);
  }
}

__global__ void transpose_kernel(const float* __restrict__ src, float* __restrict__ dst, int rows, int cols) {
  __shared__ float t[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = blockIdx.y * 32 + j;
    if (r < rows && c < cols) t[j][threadIdx.x] = src[(long long)r * cols + c];
  }
  __syncthreads();
  const int r2 = blockIdx.y * 32 + threadIdx.x;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c2 = blockIdx.x * 32 + j;
    if (r2 < rows && c2 < cols) dst[(long long)c2 * rows + r2] = t[threadIdx.x][j];
  }
}

}  // namespace

#define KB_DISPATCH_NV(NVAL, CALL)                                                                  \
  switch (NVAL) {                                                                                    \
    case 1: { constexpr int NV = 1; CALL; break; }                                                   \
    case 2: { constexpr int NV = 2; CALL; break; }                                                   \
    case 3: { constexpr int NV = 3; CALL; break; }                                                   \
    case 4: { constexpr int NV = 4; CALL; break; }                                                   \
    case 5: { constexpr int NV = 5; CALL; break; }                                                   \
    case 6: { constexpr int NV = 6; CALL; break; }                                                   \
    case 7: { constexpr int NV = 7; CALL; break; }                                                   \
    case 8: { constexpr int NV = 8; CALL; break; }                                                   \
    default: return set_error(KB_ERR_ARG, "row kernel: width %d unsupported (multiple of 128, <= 1024)", (NVAL) * 128); \
  }

int launch_layernorm(const float* x, int64_t x_row_stride, int64_t rows, int D, const float* w, const float* b,
                     float eps, void* y16, int bf16, float* y32, cudaStream_t stream, int64_t y16_pitch, int64_t lo_off) {
  if (rows <= 0) return KB_OK;
  if (D % 128 != 0 || x_row_stride % 4 != 0) return set_error(KB_ERR_ARG, "layernorm: D=%d / stride not vectorisable", D);
  if (y16_pitch <= 0) y16_pitch = D;
  if (y16_pitch % 4 != 0 || lo_off % 4 != 0 || (lo_off > 0 && lo_off + D > y16_pitch))
    return set_error(KB_ERR_ARG, "layernorm: 16-bit output pitch %lld / lo offset %lld invalid for D=%d", (long long)y16_pitch, (long long)lo_off, D);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  KB_DISPATCH_NV(D / 128, (layernorm_kernel<NV><<<grid, 256, 0, stream>>>(x, x_row_stride, rows, w, b, eps,
                                                                          (uint16_t*)y16, bf16, y32, y16_pitch, lo_off)));
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_act_l2norm(const float* x, int64_t rows, int D, int act, float* y, cudaStream_t stream) {
  if (rows <= 0) return KB_OK;
  if (D % 128 != 0) return set_error(KB_ERR_ARG, "l2norm: D=%d must be a multiple of 128", D);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  KB_DISPATCH_NV(D / 128, (act_l2norm_kernel<NV><<<grid, 256, 0, stream>>>(x, rows, act, y)));
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_cast_f32_to_16(const float* src, void* dst, int64_t n, int bf16, cudaStream_t stream) {
  if (n <= 0) return KB_OK;
  if (n % 4 != 0) return set_error(KB_ERR_ARG, "cast: n=%lld must be a multiple of 4", (long long)n);
  const long long n4 = n / 4;
  unsigned grid = (unsigned)((n4 + 255) / 256);
  if (grid > (unsigned)num_sms() * 16) grid = num_sms() * 16;
  cast_kernel<<<grid, 256, 0, stream>>>(src, (uint16_t*)dst, n4, bf16);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_cast_f32_to_hilo(const float* src, void* dst, int64_t rows, int K, int bf16, cudaStream_t stream) {
  if (rows <= 0 || K <= 0) return KB_OK;
  if (K % 4 != 0) return set_error(KB_ERR_ARG, "cast hi|lo: K=%d must be a multiple of 4", K);
  const long long n4 = rows * (K / 4);
  unsigned grid = (unsigned)((n4 + 255) / 256);
  if (grid > (unsigned)num_sms() * 16) grid = num_sms() * 16;
  cast_hilo_kernel<<<grid, 256, 0, stream>>>(src, (uint16_t*)dst, rows, K / 4, bf16);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_fold_ln(const float* W, int N, int K, const float* lnw, const float* lnb, const float* bias, void* W16,
                   int bf16, float* s, float* c, cudaStream_t stream) {
  if (N <= 0 || K <= 0 || K % 4 != 0) return set_error(KB_ERR_ARG, "fold_ln: bad shape %dx%d", N, K);
  fold_ln_kernel<<<(N + 7) / 8, 256, 0, stream>>>(W, N, K, lnw, lnb, bias, (uint16_t*)W16, bf16, s, c);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

int launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return KB_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  transpose_kernel<<<grid, block, 0, stream>>>(src, dst, rows, cols);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
