// Host runtime shared by the launchers: error slot, SM count, TMA tensor-map cache.
#include "common.h"
#include "ptx.cuh"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

namespace kb {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

int num_sms() {  // of the CURRENT device (a process may drive several GPUs: cached per device index)
  static int cache[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (cache[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev] = n > 0 ? n : 148;
  }
  return cache[dev];
}

// ---------------------------------------------------------------------------------------------------
// launch accounting and live GEMM timing
// ---------------------------------------------------------------------------------------------------
static long long g_launches = 0;
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_ev_pool;  // pairs: [2i] start, [2i+1] stop
static size_t g_ev_used = 0;
static double g_prof_flops = 0.0;
static long long g_prof_launch_base = 0;
struct ProfTag { int M, N, K, epi; double flops; };
static std::vector<ProfTag> g_prof_tags;  // one per GEMM event pair
static std::string g_prof_table;

void note_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }
bool profiling() { return g_prof_on; }

void profile_gemm_begin(cudaStream_t s) {
  if (!g_prof_on) return;
  if (g_ev_used + 2 > g_ev_pool.size()) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    g_ev_pool.push_back(a);
    g_ev_pool.push_back(b);
  }
  cudaEventRecord(g_ev_pool[g_ev_used], s);
}
void profile_gemm_end(cudaStream_t s, double flops) {
  if (!g_prof_on) return;
  cudaEventRecord(g_ev_pool[g_ev_used + 1], s);
  g_ev_used += 2;
  g_prof_flops += flops;
  if (g_prof_tags.size() < g_ev_used / 2) g_prof_tags.push_back(ProfTag{0, 0, 0, -1, flops});
}
void profile_gemm_tag(int M, int N, int K, int epi) {
  if (!g_prof_on) return;
  g_prof_tags.push_back(ProfTag{M, N, K, epi, 2.0 * M * N * K});
}
const char* profile_table() { return g_prof_table.c_str(); }
int profile_begin() {
  g_prof_on = true;
  g_ev_used = 0;
  g_prof_flops = 0.0;
  g_prof_launch_base = g_launches;
  g_prof_tags.clear();
  return KB_OK;
}
int profile_end(double* gemm_ms, double* gemm_flops, long long* gemm_launches, long long* all_launches) {
  g_prof_on = false;
  KB_CUDA_CHECK(cudaDeviceSynchronize());
  double ms = 0.0;
  struct Agg { long long n = 0; double ms = 0, flops = 0; };
  std::map<std::tuple<int, int, int, int>, Agg> agg;
  for (size_t i = 0; i + 1 < g_ev_used; i += 2) {
    float t = 0.f;
    KB_CUDA_CHECK(cudaEventElapsedTime(&t, g_ev_pool[i], g_ev_pool[i + 1]));
    ms += t;
    if (i / 2 < g_prof_tags.size()) {
      const ProfTag& tg = g_prof_tags[i / 2];
      Agg& a = agg[std::make_tuple(tg.N, tg.K, tg.epi, tg.M)];
      a.n++; a.ms += t; a.flops += tg.flops;
    }
  }
  g_prof_table.clear();
  for (auto& kv : agg) {  // "M,N,K,epi,launches,ms,TFLOP/s;"
    char buf[160];
    snprintf(buf, sizeof(buf), "%d,%d,%d,%d,%lld,%.4f,%.1f;", std::get<3>(kv.first), std::get<0>(kv.first),
             std::get<1>(kv.first), std::get<2>(kv.first), kv.second.n, kv.second.ms,
             kv.second.ms > 0 ? kv.second.flops / kv.second.ms / 1e9 : 0.0);
    g_prof_table += buf;
  }
  if (gemm_ms) *gemm_ms = ms;
  if (gemm_flops) *gemm_flops = g_prof_flops;
  if (gemm_launches) *gemm_launches = (long long)(g_ev_used / 2);
  if (all_launches) *all_launches = g_launches - g_prof_launch_base;
  g_ev_used = 0;
  return KB_OK;
}

// ---------------------------------------------------------------------------------------------------
// tensor maps
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    // resolved through the runtime so the library has no link-time dependency on libcuda
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t rows, cols, ld;
  int dtype, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && dtype == o.dtype &&
           box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h = h * 1000003u ^ static_cast<size_t>(k.rows);
    h = h * 1000003u ^ static_cast<size_t>(k.cols);
    h = h * 1000003u ^ static_cast<size_t>(k.ld);
    h = h * 1000003u ^ static_cast<size_t>(k.dtype * 1024 + k.box_rows);
    return h;
  }
};

int get_tmap_2d(const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                CUtensorMap* out) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, rows, cols, ld, dtype, box_rows};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return KB_OK;
    }
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return set_error(KB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  const int esz = (dtype == KB_F32) ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld * esz) % 16 != 0)
    return set_error(KB_ERR_ARG, "tensor map: base %p / pitch %lld B not 16-byte aligned", ptr, (long long)(ld * esz));
  if (box_rows < 1 || box_rows > 256) return set_error(KB_ERR_ARG, "tensor map: box_rows=%d out of range", box_rows);
  CUtensorMapDataType dt = dtype == KB_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                           : dtype == KB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                              : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld * esz)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / esz), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return set_error(KB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for [%lld x %lld] ld=%lld box_rows=%d", (int)r,
                     (long long)rows, (long long)cols, (long long)ld, box_rows);
  {
    std::lock_guard<std::mutex> g(mu);
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, m);
  }
  *out = m;
  return KB_OK;
}

}  // namespace kb
