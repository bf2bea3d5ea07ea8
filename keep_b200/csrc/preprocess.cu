// The reference's input transform on the device, for raw uint8 RGB tiles of any size
// (quick_start/keep_inference.py:88-93, repeated in WSI_evaluation/zeroshot_*_WSI.py:38-43):
//     Resize(224, BICUBIC)  ->  CenterCrop(224)  [ -> ToTensor -> Normalize: fused into the patch gather, frontend.cu ]
//
// torchvision applies Resize to a PIL image, i.e. Pillow's two-pass 8-bit resampler (Pillow is an un-vendored dependency
// of the reference, pinned Pillow==10.0.0 in training/requirements.txt:10; algorithm: libImaging/Resample.c). Restated
// here so that the result is BIT-IDENTICAL to the PIL path:
//   * per output coordinate a window [xmin, xmin + n) and n double-precision weights: center = (xx + 0.5) * scale,
//     support = 2 * max(scale, 1), w = cubic_{a=-0.5}((x + xmin - center + 0.5) / max(scale, 1)), normalised to sum 1,
//     then converted to 22-bit fixed point with round-half-away (precompute_coeffs / normalize_coeffs_8bpc);
//   * horizontal pass first, each pixel = clip8((2^21 + sum pixel * k) >> 22), rounded back to uint8; then the vertical
//     pass on that uint8 image (ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc);
//   * Resize(int) keeps the aspect ratio: short side -> size, long side -> int(size * long / short); CenterCrop offsets are
//     Python round() (half to even) of (extent - size) / 2.
// The coefficient tables are computed on the host in double precision with exactly Pillow's operation order (this file is
// compiled without floating-point contraction on the host side) and cached per (input extent, output extent).
#include "common.h"

#include <cmath>
#include <map>
#include <mutex>
#include <vector>

namespace kb {
namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

struct Coeffs {
  int ksize = 0;
  std::vector<int> table;  // per output coordinate: xmin, n, k[ksize]
};

// Pillow precompute_coeffs + normalize_coeffs_8bpc for the whole-image box
const Coeffs& coeffs_for(int in_size, int out_size) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, Coeffs> cache;
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find({in_size, out_size});
  if (it != cache.end()) return it->second;
  Coeffs c;
  double scale, filterscale;
  filterscale = scale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  c.ksize = (int)std::ceil(support) * 2 + 1;
  c.table.assign((size_t)out_size * (2 + c.ksize), 0);
  std::vector<double> k(c.ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    const double ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    int* row = &c.table[(size_t)xx * (2 + c.ksize)];
    row[0] = xmin;
    row[1] = xmax;
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      row[2 + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << kPrecisionBits)) : (int)(0.5 + k[x] * (1 << kPrecisionBits));
    }
  }
  return cache.emplace(std::make_pair(in_size, out_size), std::move(c)).first->second;
}

__device__ __forceinline__ uint8_t clip8(int v) {
  v >>= kPrecisionBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// out[b, y, xo - x0, c] for xo in [x0, x0 + Wc): resampled along x. in: [B, H, W, 3]; out: [B, H, Wc, 3]
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, long long B, int H, int W, const int* __restrict__ tab,
                                  int ksize, int x0, int Wc, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * H * Wc) return;
  const int xo = (int)(i % Wc);
  const long long by = i / Wc;  // b * H + y
  const int* row = tab + (long long)(x0 + xo) * (2 + ksize);
  const int xmin = row[0], n = row[1];
  const uint8_t* src = in + (by * W + xmin) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < n; ++x) {
    const int k = row[2 + x];
    s0 += src[3 * x] * k;
    s1 += src[3 * x + 1] * k;
    s2 += src[3 * x + 2] * k;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// out[b, yo - y0, x, c] for yo in [y0, y0 + Hc): resampled along y. in: [B, H, Wc, 3]; out: [B, Hc, Wc, 3]
__global__ void resample_v_kernel(const uint8_t* __restrict__ in, long long B, int H, int Wc, const int* __restrict__ tab,
                                  int ksize, int y0, int Hc, uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Hc * Wc) return;
  const int x = (int)(i % Wc);
  const int yo = (int)((i / Wc) % Hc);
  const long long b = i / ((long long)Wc * Hc);
  const int* row = tab + (long long)(y0 + yo) * (2 + ksize);
  const int ymin = row[0], n = row[1];
  const uint8_t* src = in + ((b * H + ymin) * Wc + x) * 3;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int y = 0; y < n; ++y) {
    const int k = row[2 + y];
    const uint8_t* p = src + (long long)y * Wc * 3;
    s0 += p[0] * k;
    s1 += p[1] * k;
    s2 += p[2] * k;
  }
  uint8_t* dst = out + i * 3;
  dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// plain crop: out[b, y, x, c] = in[b, y0 + y, x0 + x, c]
__global__ void crop_kernel(const uint8_t* __restrict__ in, long long B, int H, int W, int y0, int x0, int Hc, int Wc,
                            uint8_t* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Hc * Wc) return;
  const int x = (int)(i % Wc);
  const int y = (int)((i / Wc) % Hc);
  const long long b = i / ((long long)Wc * Hc);
  const uint8_t* src = in + ((b * H + y0 + y) * W + x0 + x) * 3;
  uint8_t* dst = out + i * 3;
  dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
}

// Python round(): half to even
int py_round(double v) {
  const double f = std::floor(v);
  const double d = v - f;
  if (d > 0.5) return (int)f + 1;
  if (d < 0.5) return (int)f;
  return ((long long)f % 2 == 0) ? (int)f : (int)f + 1;
}

struct Plan {
  int Hn, Wn, top, left;  // resized extent and crop offsets
  size_t tab_h, tab_v, tmp, total;  // workspace offsets
};
Plan make_plan(int64_t B, int H, int W, int size) {
  Plan p;
  // torchvision _compute_resized_output_size(size=int): short side -> size, long side -> int(size * long / short)
  const int sh = W <= H ? W : H, lg = W <= H ? H : W;
  const int ns = size, nl = (int)((long long)size * lg / sh);
  p.Wn = W <= H ? ns : nl;
  p.Hn = W <= H ? nl : ns;
  p.top = py_round((p.Hn - size) / 2.0);
  p.left = py_round((p.Wn - size) / 2.0);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) / 256 * 256; return o; };
  const int kh = (int)std::ceil(2.0 * std::max((double)W / p.Wn, 1.0)) * 2 + 1;
  const int kv = (int)std::ceil(2.0 * std::max((double)H / p.Hn, 1.0)) * 2 + 1;
  p.tab_h = take((size_t)p.Wn * (2 + kh) * 4);
  p.tab_v = take((size_t)p.Hn * (2 + kv) * 4);
  p.tmp = take((size_t)B * H * size * 3);  // horizontally resampled, already cropped in x
  p.total = off;
  return p;
}

}  // namespace

size_t preprocess_workspace_bytes(int64_t B, int64_t H, int64_t W, int size) {
  if (B <= 0 || H <= 0 || W <= 0 || size <= 0) return 0;
  return make_plan(B, (int)H, (int)W, size).total;
}

int launch_preprocess_u8(const uint8_t* tiles, int64_t B, int64_t H, int64_t W, int size, uint8_t* out, void* ws,
                         size_t ws_bytes, cudaStream_t stream) {
  if (B == 0) return KB_OK;
  if (!tiles || !out || B < 0 || H <= 0 || W <= 0 || size <= 0 || H > 16384 || W > 16384)
    return set_error(KB_ERR_ARG, "preprocess: bad arguments (B=%lld, %lldx%lld -> %d)", (long long)B, (long long)H,
                     (long long)W, size);
  const Plan p = make_plan(B, (int)H, (int)W, size);
  const long long n_out = (long long)B * size * size;
  const unsigned g_out = (unsigned)((n_out + 255) / 256);
  if (p.Hn == H && p.Wn == W) {  // short side already `size`: CenterCrop only (PIL's resize returns a copy)
    crop_kernel<<<g_out, 256, 0, stream>>>(tiles, B, (int)H, (int)W, p.top, p.left, size, size, out);
    note_launch();
    KB_CUDA_CHECK(cudaGetLastError());
    return KB_OK;
  }
  if (!ws || ws_bytes < p.total || (reinterpret_cast<uintptr_t>(ws) & 255) != 0)
    return set_error(KB_ERR_WORKSPACE, "preprocess: workspace %zu B < %zu B (or not 256-byte aligned)", ws_bytes, p.total);
  char* w8 = static_cast<char*>(ws);
  const Coeffs& ch = coeffs_for((int)W, p.Wn);
  const Coeffs& cv = coeffs_for((int)H, p.Hn);
  // the tables live in a process-lifetime cache, so the asynchronous copies read stable host memory
  KB_CUDA_CHECK(cudaMemcpyAsync(w8 + p.tab_h, ch.table.data(), ch.table.size() * 4, cudaMemcpyHostToDevice, stream));
  KB_CUDA_CHECK(cudaMemcpyAsync(w8 + p.tab_v, cv.table.data(), cv.table.size() * 4, cudaMemcpyHostToDevice, stream));
  uint8_t* tmp = reinterpret_cast<uint8_t*>(w8 + p.tmp);
  // Pillow resamples horizontally first (only when the width changes), then vertically (only when the height changes)
  const uint8_t* vsrc = tiles;
  int vsrc_w = (int)W, vx0 = p.left;
  if (p.Wn != W) {
    const long long n_h = (long long)B * H * size;
    resample_h_kernel<<<(unsigned)((n_h + 255) / 256), 256, 0, stream>>>(tiles, B, (int)H, (int)W,
                                                                        reinterpret_cast<const int*>(w8 + p.tab_h), ch.ksize,
                                                                        p.left, size, tmp);
    note_launch();
    KB_CUDA_CHECK(cudaGetLastError());
    vsrc = tmp;
    vsrc_w = size;
    vx0 = 0;
  }
  if (p.Hn != H) {
    if (vsrc == tiles) {  // width unchanged: crop the columns first so the vertical pass sees [B, H, size, 3]
      crop_kernel<<<(unsigned)(((long long)B * H * size + 255) / 256), 256, 0, stream>>>(tiles, B, (int)H, (int)W, 0, vx0,
                                                                                        (int)H, size, tmp);
      note_launch();
      KB_CUDA_CHECK(cudaGetLastError());
      vsrc = tmp;
    }
    resample_v_kernel<<<g_out, 256, 0, stream>>>(vsrc, B, (int)H, size, reinterpret_cast<const int*>(w8 + p.tab_v),
                                                 cv.ksize, p.top, size, out);
    note_launch();
    KB_CUDA_CHECK(cudaGetLastError());
  } else {  // height unchanged: the horizontally resampled rows only need the vertical crop
    crop_kernel<<<g_out, 256, 0, stream>>>(vsrc, B, (int)H, vsrc_w, p.top, vx0, size, size, out);
    note_launch();
    KB_CUDA_CHECK(cudaGetLastError());
  }
  return KB_OK;
}

}  // namespace kb
