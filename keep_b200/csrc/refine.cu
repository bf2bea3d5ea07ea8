// refine_seg on the device (WSI_evaluation/detection_utils.py:39-74, subtyping_utils.py:38-65,
// segment_utils.py:63-89).
//
// The reference walks the tiles in order, keeps the FIRST tile seen at each (x,y) in a dict keyed by the
// string "x_y", then (overlap=True) replaces each kept tile's class probabilities by the float32 mean over
// the kept tiles present among (x-ps,y-ps), (x,y-ps), (x-ps,y), (x,y), in that order.
//
// Here: an open-addressing hash table over the packed 64-bit coordinate holds, per distinct coordinate, the
// minimum tile index (atomicMin = "first occurrence wins"); a second pass gathers the neighbours. Integer
// work is exact; the mean is accumulated in the reference's order ((lt+rt)+lb)+rb and divided by the count,
// which reproduces numpy's float32 result bit for bit.
#include "common.h"
#include "ptx.cuh"

namespace kb {
namespace {

constexpr unsigned long long kEmpty = 0xFFFFFFFFFFFFFFFFull;

// Key of a coordinate: both components biased by 2^31 into 32 bits each. Coordinates in [-2^31, 2^31 - 2] (negative ones
// included: the reference's dict keeps a tile at (-1,-1) like any other) map to distinct keys, none of which is kEmpty;
// anything outside that range has no key (kEmpty: "absent" for a lookup, e.g. a neighbour x - ps below the range) - the
// Python front end refuses such tile coordinates before the call (keep_b200/ops.py::refine), the C entry point documents it.
__device__ __forceinline__ unsigned long long pack_xy(long long x, long long y) {
  const long long lo = -2147483648ll, hi = 2147483646ll;
  if (x < lo || x > hi || y < lo || y > hi) return kEmpty;
  return (static_cast<unsigned long long>(x - lo) << 32) | static_cast<unsigned long long>(y - lo);
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z ^= z >> 33; z *= 0xff51afd7ed558ccdull;
  z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ull;
  z ^= z >> 33;
  return z;
}

__global__ void table_clear_kernel(unsigned long long* keys, long long* vals, long long cap) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < cap) { keys[i] = kEmpty; vals[i] = 0x7FFFFFFFFFFFFFFFll; }
}

__global__ void table_insert_kernel(const long long* __restrict__ coords, long long N, unsigned long long* keys,
                                    long long* vals, long long mask) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const unsigned long long key = pack_xy(coords[2 * i], coords[2 * i + 1]);
  if (key == kEmpty) return;  // outside the supported range: never stored (keep = 0 for it; refused by the front end)
  long long slot = (long long)(mix64(key) & (unsigned long long)mask);
  while (true) {
    const unsigned long long prev = atomicCAS(&keys[slot], kEmpty, key);
    if (prev == kEmpty || prev == key) {
      atomicMin(&vals[slot], i);
      return;
    }
    slot = (slot + 1) & mask;
  }
}

__device__ __forceinline__ long long table_find(const unsigned long long* keys, const long long* vals, long long mask,
                                                unsigned long long key) {
  if (key == kEmpty) return -1;  // no key: outside the coordinate range
  long long slot = (long long)(mix64(key) & (unsigned long long)mask);
  while (true) {
    const unsigned long long k = keys[slot];
    if (k == key) return vals[slot];
    if (k == kEmpty) return -1;
    slot = (slot + 1) & mask;
  }
}

__global__ void refine_kernel(const long long* __restrict__ coords, const float* __restrict__ probs, long long N, int C,
                              long long ps, int overlap, const unsigned long long* __restrict__ keys,
                              const long long* __restrict__ vals, long long mask, uint8_t* __restrict__ keep,
                              float* __restrict__ refined) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const long long x = coords[2 * i], y = coords[2 * i + 1];
  const long long first = table_find(keys, vals, mask, pack_xy(x, y));
  const bool kept = (first == i);
  keep[i] = kept ? 1 : 0;
  float* dst = refined + i * C;
  if (!kept) {
    for (int c = 0; c < C; ++c) dst[c] = 0.f;
    return;
  }
  if (!overlap) {
    for (int c = 0; c < C; ++c) dst[c] = probs[i * C + c];
    return;
  }
  long long nb[4];
  nb[0] = table_find(keys, vals, mask, pack_xy(x - ps, y - ps));
  nb[1] = table_find(keys, vals, mask, pack_xy(x, y - ps));
  nb[2] = table_find(keys, vals, mask, pack_xy(x - ps, y));
  nb[3] = i;
  int cnt = 0;
  for (int j = 0; j < 4; ++j) cnt += nb[j] >= 0;
  const float fc = (float)cnt;
  for (int c = 0; c < C; ++c) {
    float s = 0.f;
    bool any = false;
    for (int j = 0; j < 4; ++j) {
      if (nb[j] < 0) continue;
      const float v = probs[nb[j] * C + c];
      s = any ? __fadd_rn(s, v) : v;
      any = true;
    }
    dst[c] = __fdiv_rn(s, fc);
  }
}

long long table_capacity(long long N) {
  long long cap = 64;
  while (cap < 2 * N) cap <<= 1;
  return cap;
}

}  // namespace

size_t refine_workspace_bytes(int64_t N) { return (size_t)table_capacity(N) * 16; }

int launch_refine(const int64_t* coords, const float* probs, int64_t N, int C, int64_t ps, int overlap, uint8_t* keep,
                  float* refined, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (N <= 0) return KB_OK;
  const long long cap = table_capacity(N);
  if (ws == nullptr || ws_bytes < (size_t)cap * 16)
    return set_error(KB_ERR_WORKSPACE, "refine: workspace %zu B < %zu B", ws_bytes, (size_t)cap * 16);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws);
  long long* vals = reinterpret_cast<long long*>(keys + cap);
  table_clear_kernel<<<(unsigned)((cap + 255) / 256), 256, 0, stream>>>(keys, vals, cap);
  table_insert_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>((const long long*)coords, N, keys, vals, cap - 1);
  refine_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>((const long long*)coords, probs, N, C, ps, overlap, keys,
                                                                 vals, cap - 1, keep, refined);
  note_launch(3);
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

}  // namespace kb
