// C ABI of libkeep_b200.so (include/keep_b200.h): model handle, strict state-dict ingestion with
// repacking into kernel layouts, and the encode_image / encode_text / similarity / screening / refine
// pipelines expressed as sequences of the kernels in this directory on one CUDA stream.
#include "../../include/keep_b200.h"
#include "common.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

using namespace kb;

namespace {

// One tensor of the reference state-dict and where it lands.
struct WeightSlot {
  std::string name;
  std::vector<int64_t> shape;  // reference shape
  bool as16 = false;           // converted to the 16-bit operand dtype (GEMM weights)
  bool hilo = false;           // ... as a [N, 2K] hi|lo operand (split-operand GEMMs, common.h GEMM_SPLIT_*)
  void* dst_hl = nullptr;      // a second, hi|lo copy next to the plain 16-bit one (ViT weights: fast and high-precision paths)
  void* dst = nullptr;         // device destination (base of the owning allocation + offset)
  bool loaded = false;
  bool ignored = false;        // accepted but unused (logit_scale: dead at inference)
  float* master = nullptr;     // fp32 copy kept for weights that are re-derived at finalize (LayerNorm folding)
};

struct VitBlock {
  float *n1w, *n1b, *qkv_b, *proj_b, *ls1, *n2w, *n2b, *fc1_b, *fc2_b, *ls2;
  void *qkv_w, *proj_w, *fc1_w, *fc2_w;      // plain 16-bit [N, K]: the one-pass path
  void *qkv_hl, *proj_hl, *fc1_hl, *fc2_hl;  // [N, 2K] hi|lo: the split-operand path (image_precision high; CLS-row tail)
  // LayerNorm folded into the following Linear (EPI_LN_*): 16-bit W * ln.weight, column sums, b.W^T + bias
  void *qkv_wf = nullptr, *fc1_wf = nullptr;
  float *qkv_s = nullptr, *qkv_c = nullptr, *fc1_s = nullptr, *fc1_c = nullptr;
};
struct BertLayer {
  void *qkv_w, *ao_w, *in_w, *out_w;  // [N, 2K] hi|lo operands: the hi half alone (row pitch 2K) is the plain 16-bit weight
  float *qkv_b, *ao_b, *ao_lnw, *ao_lnb, *in_b, *out_b, *out_lnw, *out_lnb;
};

struct Model {
  KeepB200Config cfg;
  int device = 0;
  bool finalized = false;
  std::vector<WeightSlot> slots;
  std::unordered_map<std::string, int> index;
  std::vector<void*> allocs;
  // vision
  float *cls = nullptr, *pos = nullptr, *pe_b = nullptr, *norm_w = nullptr, *norm_b = nullptr;
  void *pe_w = nullptr, *pe_hl = nullptr;
  std::vector<VitBlock> blocks;
  float *h0_w = nullptr, *h2_w = nullptr, *h0_b = nullptr, *h2_b = nullptr;  // visual_head in fp32 (head.cu)
  float *h0_wt = nullptr, *h2_wt = nullptr;                                  // ... transposed [K, N] at finalize
  // text
  float *word = nullptr, *tpos = nullptr, *ttype = nullptr, *emb_lnw = nullptr, *emb_lnb = nullptr;
  std::vector<BertLayer> layers;
  float *pool_w = nullptr, *pool_b = nullptr, *pool_wt = nullptr;  // pooler in fp32 (head.cu), transposed at finalize
  // test / analysis hooks (include/keep_b200.h "debug")
  int ln_fuse = 1;                 // 0: stand-alone LayerNorm kernels, 1: norm1 folded into qkv (default), 2: norm2 -> fc1 too
  float* dump = nullptr;           // per-layer residual-stream dump of the next encode call (single chunk), or null
  size_t dump_bytes = 0;

  int grid() const { return cfg.img_size / cfg.patch_size; }
  int tokens() const { return grid() * grid() + 1; }
};

int alloc_dev(Model* m, size_t bytes, void** p) {
  KB_CUDA_CHECK(cudaMalloc(p, bytes));
  m->allocs.push_back(*p);
  return KB_OK;
}

// register a slot whose destination is (base + elem_offset) of an allocation of `as16 ? 2 : 4`-byte elements
void add_slot(Model* m, const std::string& name, std::vector<int64_t> shape, bool as16, void* base, size_t elem_off) {
  WeightSlot s;
  s.name = name;
  s.shape = std::move(shape);
  s.as16 = as16;
  s.dst = static_cast<char*>(base) + elem_off * (as16 ? 2 : 4);
  m->index[name] = (int)m->slots.size();
  m->slots.push_back(std::move(s));
}

#define KB_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != KB_OK) return _rc; \
  } while (0)

int new_f32(Model* m, const std::string& name, std::vector<int64_t> shape, float** out) {
  size_t n = 1;
  for (auto d : shape) n *= (size_t)d;
  void* p;
  KB_TRY(alloc_dev(m, n * 4, &p));
  *out = static_cast<float*>(p);
  add_slot(m, name, std::move(shape), false, p, 0);
  return KB_OK;
}
int new_w16(Model* m, const std::string& name, std::vector<int64_t> shape, void** out, bool hilo = false) {
  size_t n = 1;
  for (auto d : shape) n *= (size_t)d;
  void* p;
  KB_TRY(alloc_dev(m, n * (hilo ? 4 : 2), &p));
  *out = p;
  add_slot(m, name, std::move(shape), true, p, 0);
  m->slots.back().hilo = hilo;
  return KB_OK;
}
// plain 16-bit [N, K] and hi|lo [N, 2K] copies of the same weight
int new_w16_both(Model* m, const std::string& name, std::vector<int64_t> shape, void** plain, void** hl) {
  size_t n = 1;
  for (auto d : shape) n *= (size_t)d;
  KB_TRY(new_w16(m, name, std::move(shape), plain, false));
  KB_TRY(alloc_dev(m, n * 4, hl));
  m->slots.back().dst_hl = *hl;
  return KB_OK;
}

// the slot registered last (a GEMM weight) also keeps an fp32 master and gets a LayerNorm-folded twin
int add_fold(Model* m, size_t N, size_t K, void** wf, float** s_vec, float** c_vec) {
  void* p;
  KB_TRY(alloc_dev(m, N * K * 4, &p));
  m->slots.back().master = static_cast<float*>(p);
  KB_TRY(alloc_dev(m, N * K * 2, wf));
  KB_TRY(alloc_dev(m, N * 4, &p));
  *s_vec = static_cast<float*>(p);
  KB_TRY(alloc_dev(m, N * 4, &p));
  *c_vec = static_cast<float*>(p);
  return KB_OK;
}

int build_tables(Model* m) {
  const KeepB200Config& c = m->cfg;
  const int D = c.vit_width, F = c.vit_mlp, T = m->tokens(), ps = c.patch_size;
  // ---- vision tower: timm vit_large_patch16_224 state-dict ----
  KB_TRY(new_f32(m, "visual.cls_token", {1, 1, D}, &m->cls));
  KB_TRY(new_f32(m, "visual.pos_embed", {1, T, D}, &m->pos));
  KB_TRY(new_w16_both(m, "visual.patch_embed.proj.weight", {D, 3, ps, ps}, &m->pe_w, &m->pe_hl));
  KB_TRY(new_f32(m, "visual.patch_embed.proj.bias", {D}, &m->pe_b));
  m->blocks.resize(c.vit_depth);
  for (int i = 0; i < c.vit_depth; ++i) {
    VitBlock& b = m->blocks[i];
    const std::string p = "visual.blocks." + std::to_string(i) + ".";
    KB_TRY(new_f32(m, p + "norm1.weight", {D}, &b.n1w));
    KB_TRY(new_f32(m, p + "norm1.bias", {D}, &b.n1b));
    KB_TRY(new_w16_both(m, p + "attn.qkv.weight", {3 * D, D}, &b.qkv_w, &b.qkv_hl));
    KB_TRY(add_fold(m, (size_t)3 * D, D, &b.qkv_wf, &b.qkv_s, &b.qkv_c));
    KB_TRY(new_f32(m, p + "attn.qkv.bias", {3 * D}, &b.qkv_b));
    KB_TRY(new_w16_both(m, p + "attn.proj.weight", {D, D}, &b.proj_w, &b.proj_hl));
    KB_TRY(new_f32(m, p + "attn.proj.bias", {D}, &b.proj_b));
    KB_TRY(new_f32(m, p + "ls1.gamma", {D}, &b.ls1));
    KB_TRY(new_f32(m, p + "norm2.weight", {D}, &b.n2w));
    KB_TRY(new_f32(m, p + "norm2.bias", {D}, &b.n2b));
    KB_TRY(new_w16_both(m, p + "mlp.fc1.weight", {F, D}, &b.fc1_w, &b.fc1_hl));
    KB_TRY(add_fold(m, (size_t)F, D, &b.fc1_wf, &b.fc1_s, &b.fc1_c));
    KB_TRY(new_f32(m, p + "mlp.fc1.bias", {F}, &b.fc1_b));
    KB_TRY(new_w16_both(m, p + "mlp.fc2.weight", {D, F}, &b.fc2_w, &b.fc2_hl));
    KB_TRY(new_f32(m, p + "mlp.fc2.bias", {D}, &b.fc2_b));
    KB_TRY(new_f32(m, p + "ls2.gamma", {D}, &b.ls2));
  }
  KB_TRY(new_f32(m, "visual.norm.weight", {D}, &m->norm_w));
  KB_TRY(new_f32(m, "visual.norm.bias", {D}, &m->norm_b));
  // ---- visual_head (keep_inference.py:42-46) ----
  KB_TRY(new_f32(m, "visual_head.0.weight", {c.proj_dim, D}, &m->h0_w));
  KB_TRY(new_f32(m, "visual_head.0.bias", {c.proj_dim}, &m->h0_b));
  KB_TRY(new_f32(m, "visual_head.2.weight", {c.proj_dim, c.proj_dim}, &m->h2_w));
  KB_TRY(new_f32(m, "visual_head.2.bias", {c.proj_dim}, &m->h2_b));
  {
    void* t;
    KB_TRY(alloc_dev(m, (size_t)c.proj_dim * D * 4, &t));
    m->h0_wt = static_cast<float*>(t);
    KB_TRY(alloc_dev(m, (size_t)c.proj_dim * c.proj_dim * 4, &t));
    m->h2_wt = static_cast<float*>(t);
  }
  // ---- logit_scale (keep_inference.py:52): present in the state-dict, unused at inference ----
  {
    WeightSlot s;
    s.name = "logit_scale";
    s.ignored = true;
    m->index[s.name] = (int)m->slots.size();
    m->slots.push_back(s);
  }
  // ---- text tower: transformers BertModel state-dict ----
  const int d = c.hidden, I = c.intermediate;
  KB_TRY(new_f32(m, "text.embeddings.word_embeddings.weight", {c.vocab_size, d}, &m->word));
  KB_TRY(new_f32(m, "text.embeddings.position_embeddings.weight", {c.max_pos, d}, &m->tpos));
  KB_TRY(new_f32(m, "text.embeddings.token_type_embeddings.weight", {c.type_vocab, d}, &m->ttype));
  KB_TRY(new_f32(m, "text.embeddings.LayerNorm.weight", {d}, &m->emb_lnw));
  KB_TRY(new_f32(m, "text.embeddings.LayerNorm.bias", {d}, &m->emb_lnb));
  m->layers.resize(c.layers);
  for (int i = 0; i < c.layers; ++i) {
    BertLayer& L = m->layers[i];
    const std::string p = "text.encoder.layer." + std::to_string(i) + ".";
    // query / key / value are fused into one [3d, d] operand and one [3d] bias (rows q | k | v)
    void* qkv_w;
    void* qkv_b;
    KB_TRY(alloc_dev(m, (size_t)3 * d * d * 4, &qkv_w));  // [3d, 2d] hi|lo
    KB_TRY(alloc_dev(m, (size_t)3 * d * 4, &qkv_b));
    L.qkv_w = qkv_w;
    L.qkv_b = static_cast<float*>(qkv_b);
    const char* nm[3] = {"query", "key", "value"};
    for (int j = 0; j < 3; ++j) {
      add_slot(m, p + "attention.self." + nm[j] + ".weight", {d, d}, true, qkv_w, (size_t)j * d * 2 * d);
      m->slots.back().hilo = true;
      add_slot(m, p + "attention.self." + nm[j] + ".bias", {d}, false, qkv_b, (size_t)j * d);
    }
    KB_TRY(new_w16(m, p + "attention.output.dense.weight", {d, d}, &L.ao_w, true));
    KB_TRY(new_f32(m, p + "attention.output.dense.bias", {d}, &L.ao_b));
    KB_TRY(new_f32(m, p + "attention.output.LayerNorm.weight", {d}, &L.ao_lnw));
    KB_TRY(new_f32(m, p + "attention.output.LayerNorm.bias", {d}, &L.ao_lnb));
    KB_TRY(new_w16(m, p + "intermediate.dense.weight", {I, d}, &L.in_w, true));
    KB_TRY(new_f32(m, p + "intermediate.dense.bias", {I}, &L.in_b));
    KB_TRY(new_w16(m, p + "output.dense.weight", {d, I}, &L.out_w, true));
    KB_TRY(new_f32(m, p + "output.dense.bias", {d}, &L.out_b));
    KB_TRY(new_f32(m, p + "output.LayerNorm.weight", {d}, &L.out_lnw));
    KB_TRY(new_f32(m, p + "output.LayerNorm.bias", {d}, &L.out_lnb));
  }
  KB_TRY(new_f32(m, "text.pooler.dense.weight", {d, d}, &m->pool_w));
  KB_TRY(new_f32(m, "text.pooler.dense.bias", {d}, &m->pool_b));
  {
    void* t;
    KB_TRY(alloc_dev(m, (size_t)d * d * 4, &t));
    m->pool_wt = static_cast<float*>(t);
  }
  return KB_OK;
}

int validate_cfg(const KeepB200Config& c) {
  if (c.struct_size != (int32_t)sizeof(KeepB200Config))
    return set_error(KB_ERR_ARG, "config: struct_size %d != %zu", c.struct_size, sizeof(KeepB200Config));
  if (c.patch_size != 16) return set_error(KB_ERR_ARG, "config: patch_size %d unsupported (16 only)", c.patch_size);
  if (c.img_size <= 0 || c.img_size % 16 != 0) return set_error(KB_ERR_ARG, "config: img_size %d", c.img_size);
  if (c.vit_width != c.vit_heads * 64 || c.hidden != c.heads * 64)
    return set_error(KB_ERR_ARG, "config: head dim must be 64 (width %d / heads %d, hidden %d / heads %d)", c.vit_width,
                     c.vit_heads, c.hidden, c.heads);
  if (c.vit_width % 128 || c.vit_width > 1024 || c.hidden % 128 || c.hidden > 1024 || c.proj_dim % 128 ||
      c.proj_dim > 1024)
    return set_error(KB_ERR_ARG, "config: widths must be multiples of 128 and <= 1024");
  if (c.vit_mlp % 64 || c.intermediate % 64) return set_error(KB_ERR_ARG, "config: mlp widths must be multiples of 64");
  if (c.vit_depth < 1 || c.layers < 1) return set_error(KB_ERR_ARG, "config: depth/layers must be >= 1");
  if (c.operand_dtype != KEEPB200_FP16 && c.operand_dtype != KEEPB200_BF16)
    return set_error(KB_ERR_ARG, "config: operand_dtype %d", c.operand_dtype);
  return KB_OK;
}

int check_device(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return set_error(KB_ERR_CUDA, "no CUDA device available (%s); keep_b200 has no CPU path", cudaGetErrorString(e));
  if (device < 0 || device >= count) return set_error(KB_ERR_ARG, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  KB_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return set_error(KB_ERR_CUDA, "device %d is sm_%d%d; keep_b200 kernels are built for sm_100a only", device, prop.major,
                     prop.minor);
  return KB_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- workspace layouts -------------------------------------------------------------------------------------
struct ImageWs {
  size_t x, xn, qkv, att, hid, xc, cls16, stats, pos, total;
};
// gh x gw = patch grid of the tiles (the model's own grid unless dynamic_img_size is exercised)
// high: every 16-bit activation that feeds a GEMM is a [rows, 2*width] hi|lo operand (split-operand path)
ImageWs image_ws(const Model* m, int64_t n, int gh, int gw, bool high) {
  const KeepB200Config& c = m->cfg;
  const size_t T = (size_t)gh * gw + 1;
  const size_t M = (size_t)n * T;
  const size_t D = c.vit_width, F = c.vit_mlp, hl = high ? 2 : 1;
  const size_t patch_bytes = (size_t)n * (T - 1) * 768 * 2 * hl;
  ImageWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  w.x = take(M * D * 4);
  w.xn = take(M * D * 2 * hl);
  w.qkv = take(M * 3 * D * 2);
  w.att = take(M * D * 2 * hl);  // high: the attention context as a [M, 2D] hi|lo operand of the output projection
  size_t hid_bytes = M * F * 2 * hl;  // also holds the [n, 2F] hi|lo hidden of the CLS-row tail (T >= 2)
  if (patch_bytes > hid_bytes) hid_bytes = patch_bytes;  // the patch matrix aliases the MLP hidden buffer
  w.hid = take(hid_bytes);
  w.xc = take((size_t)n * D * 4);  // CLS rows of the residual stream (last block onwards)
  w.cls16 = take((size_t)n * 2 * D * 2);  // their LayerNorm output as a [n, 2D] hi|lo operand
  w.stats = take(M * (D / kLnSliceCols) * 8);  // LayerNorm partial sums of the residual rows (fused-LN path)
  w.pos = take((gh == m->grid() && gw == m->grid()) ? 0 : T * D * 4);  // resampled pos_embed (dynamic_img_size)
  w.total = off;
  return w;
}
struct TextWs {
  size_t x32, x16, qkv, att, hid, xc32, xc16, total;
};
// 16-bit activations are laid out as [rows, 2*width] hi|lo operands in both precision modes (fast mode leaves the lo
// halves unused), so one workspace size serves both
TextWs text_ws(const Model* m, int64_t n, int64_t s) {
  const KeepB200Config& c = m->cfg;
  const size_t M = (size_t)n * s, d = c.hidden, I = c.intermediate;
  TextWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 1024); return o; };
  w.x32 = take(M * d * 4);
  w.x16 = take(M * 2 * d * 2);
  w.qkv = take(M * 3 * d * 2);
  w.att = take(M * 2 * d * 2);
  w.hid = take(M * 2 * I * 2);
  w.xc32 = take((size_t)n * d * 4);  // [CLS] rows (last layer onwards)
  w.xc16 = take((size_t)n * 2 * d * 2);
  w.total = off;
  return w;
}

// out = epilogue(A . W^T); lda / ldw default to K (plain operands)
struct G {
  GemmArgs a;
  G(const void* A, const void* W, int M, int N, int K, int epi, int bf16, void* out, int64_t ldo) {
    a.A = A; a.lda = K; a.W = W; a.ldw = K; a.M = M; a.N = N; a.K = K; a.epi = epi; a.bf16 = bf16;
    a.bias = nullptr; a.gamma = nullptr; a.resid = nullptr; a.ldr = ldo; a.out = out; a.ldo = ldo; a.pos = nullptr; a.patches = 0;
  }
  G& bias(const float* b) { a.bias = b; return *this; }
  G& gamma(const float* g) { a.gamma = g; return *this; }
  G& resid(const float* r, int64_t ldr = -1) { a.resid = r; if (ldr >= 0) a.ldr = ldr; return *this; }
  G& pitch(int64_t lda, int64_t ldw) { if (lda > 0) a.lda = lda; if (ldw > 0) a.ldw = ldw; return *this; }
  G& split(int mode) { a.split = mode; return *this; }
  G& lo(int64_t off) { a.lo_off = off; return *this; }
  G& patch(const float* pos, int patches) { a.pos = pos; a.patches = patches; return *this; }
  int run(cudaStream_t st) const { return launch_gemm(a, st); }
};

// residual GEMM that also leaves the 16-bit copy of the new residual rows and their LayerNorm partial sums behind
int gemm_resid_stats(const void* A, const void* W, int M, int N, int K, int bf16, const float* bias, const float* gamma,
                     float* x, void* x16, float* stats, cudaStream_t st) {
  GemmArgs a;
  a.A = A; a.lda = K; a.W = W; a.ldw = K; a.M = M; a.N = N; a.K = K; a.epi = EPI_RESID_F32_STATS; a.bf16 = bf16;
  a.bias = bias; a.gamma = gamma; a.resid = x; a.ldr = N; a.out = x; a.ldo = N; a.pos = nullptr; a.patches = 0;
  a.out16 = x16; a.ldo16 = N; a.stats_out = stats;
  return launch_gemm(a, st);
}
// GEMM on the un-normalised 16-bit residual copy with the LayerNorm folded into W / finished in the epilogue
int gemm_ln(const void* x16, int D, const void* Wf, int M, int N, int epi, int bf16, const float* c_vec, const float* s_vec,
            const float* stats, float eps, void* out, cudaStream_t st) {
  GemmArgs a;
  a.A = x16; a.lda = D; a.W = Wf; a.ldw = D; a.M = M; a.N = N; a.K = D; a.epi = epi; a.bf16 = bf16;
  a.bias = c_vec; a.gamma = nullptr; a.resid = nullptr; a.ldr = 0; a.out = out; a.ldo = N; a.pos = nullptr; a.patches = 0;
  a.ln_stats = stats; a.ln_slices = D / kLnSliceCols; a.ln_width = D; a.ln_eps = eps; a.ln_s = s_vec;
  return launch_gemm(a, st);
}

// debug hook: copy `floats` fp32 values into slot `index` (slots of `slot_floats`) of the caller's layer dump
int dump_layer(Model* m, int index, size_t slot_floats, const float* src, size_t floats, cudaStream_t st) {
  if (m->dump == nullptr) return KB_OK;
  if (((size_t)index * slot_floats + floats) * 4 > m->dump_bytes)
    return set_error(KB_ERR_WORKSPACE, "layer dump: buffer of %zu B too small for layer %d", m->dump_bytes, index);
  KB_CUDA_CHECK(cudaMemcpyAsync(m->dump + (size_t)index * slot_floats, src, floats * 4, cudaMemcpyDeviceToDevice, st));
  return KB_OK;
}

// precision level of a call: LEVEL_FAST one MMA pass per GEMM, LEVEL_BALANCED two (hi|lo weights, one 16-bit value per
// activation: Ah.Wh + Ah.Wl - the weights' share of the operand rounding is gone), LEVEL_HIGH three (hi|lo both sides)
enum { LEVEL_FAST = 0, LEVEL_HIGH = 1, LEVEL_BALANCED = 2 };

int encode_image_chunk(Model* m, const void* tiles, int layout, int64_t n, int gh, int gw, int level, float* out, char* ws,
                       cudaStream_t st) {
  const bool high = level == LEVEL_HIGH, wsplit = level == LEVEL_BALANCED;
  const KeepB200Config& c = m->cfg;
  const int bf = c.operand_dtype, D = c.vit_width, F = c.vit_mlp, T = gh * gw + 1;
  const int M = (int)(n * T);
  const ImageWs w = image_ws(m, n, gh, gw, high);
  float* x = reinterpret_cast<float*>(ws + w.x);
  void* xn = ws + w.xn;
  void* qkv = ws + w.qkv;
  void* att = ws + w.att;
  void* hid = ws + w.hid;
  float* xc = reinterpret_cast<float*>(ws + w.xc);
  void* cls16 = ws + w.cls16;
  float* stats = reinterpret_cast<float*>(ws + w.stats);
  // Fused LayerNorm (default): proj / fc2 leave 16-bit(x) in `xn` plus per-row partial sums, and fc1 / the next
  // block's qkv run on it with the LayerNorm folded in (EPI_LN_*). Only norm1 of block 0 (x comes from the patch
  // embedding) and the CLS-row tail of the last block use the stand-alone kernel.
  const int fuse = m->ln_fuse;

  // dynamic_img_size (keep_inference.py:39): other grids use pos_embed resampled like timm's resample_abs_pos_embed
  const float* pos = m->pos;
  if (gh != m->grid() || gw != m->grid()) {
    float* rp = reinterpret_cast<float*>(ws + w.pos);
    KB_TRY(launch_pos_resample(m->pos, m->grid(), gh, gw, D, rp, st));
    pos = rp;
  }
  // patch gather (+ CLS rows), then patch-embed GEMM scattering into x[b, 1+p, :] with +bias +pos
  if (layout == KEEPB200_TILES_F32_NCHW)
    KB_TRY(launch_im2col(static_cast<const float*>(tiles), n, gh, gw, hid, bf, m->cls, pos, x, D, st, high));
  else
    KB_TRY(launch_im2col_u8(static_cast<const uint8_t*>(tiles), n, gh, gw, hid, bf, m->cls, pos, x, D, st, high));
  if (high)
    KB_TRY(G(hid, m->pe_hl, (int)(n * (T - 1)), D, 768, EPI_PATCH_F32, bf, x, D).bias(m->pe_b).patch(pos, T - 1)
               .pitch(1536, 1536).split(GEMM_SPLIT_AW).run(st));
  else if (wsplit)
    KB_TRY(G(hid, m->pe_hl, (int)(n * (T - 1)), D, 768, EPI_PATCH_F32, bf, x, D).bias(m->pe_b).patch(pos, T - 1)
               .pitch(768, 1536).split(GEMM_SPLIT_W).run(st));
  else
    KB_TRY(G(hid, m->pe_w, (int)(n * (T - 1)), D, 768, EPI_PATCH_F32, bf, x, D).bias(m->pe_b).patch(pos, T - 1).run(st));
  for (int i = 0; i < c.vit_depth; ++i) {
    const VitBlock& b = m->blocks[i];
    if (high) {
      // split-operand path: stand-alone LayerNorm kernels emit hi|lo rows, every GEMM makes three passes over hi|lo
      // operands; what stays 16-bit is the attention (q, k, v, P and the context) - oracle/precision_model.py
      KB_TRY(launch_layernorm(x, D, M, D, b.n1w, b.n1b, c.vit_ln_eps, xn, bf, nullptr, st, 2 * D, D));
      KB_TRY(G(xn, b.qkv_hl, M, 3 * D, D, EPI_BIAS_HALF, bf, qkv, 3 * D).bias(b.qkv_b).pitch(2 * D, 2 * D).split(GEMM_SPLIT_AW).run(st));
    } else if (wsplit) {
      // two-pass path: the un-folded hi|lo weights against LayerNorm rows rounded once (stand-alone LayerNorm kernels)
      KB_TRY(launch_layernorm(x, D, M, D, b.n1w, b.n1b, c.vit_ln_eps, xn, bf, nullptr, st));
      KB_TRY(G(xn, b.qkv_hl, M, 3 * D, D, EPI_BIAS_HALF, bf, qkv, 3 * D).bias(b.qkv_b).pitch(D, 2 * D).split(GEMM_SPLIT_W).run(st));
    } else if (fuse >= 1 && i > 0) {
      KB_TRY(gemm_ln(xn, D, b.qkv_wf, M, 3 * D, EPI_LN_BIAS_HALF, bf, b.qkv_c, b.qkv_s, stats, c.vit_ln_eps, qkv, st));
    } else {
      KB_TRY(launch_layernorm(x, D, M, D, b.n1w, b.n1b, c.vit_ln_eps, xn, bf, nullptr, st));
      KB_TRY(G(xn, b.qkv_w, M, 3 * D, D, EPI_BIAS_HALF, bf, qkv, 3 * D).bias(b.qkv_b).run(st));
    }
    if (high) KB_TRY(launch_attention(qkv, att, (int)n, T, c.vit_heads, bf, nullptr, 0, 0.125f, st, 2 * D, D));  // context hi|lo
    else KB_TRY(launch_attention(qkv, att, (int)n, T, c.vit_heads, bf, nullptr, 0, 0.125f, st));
    if (i + 1 < c.vit_depth && high) {
      KB_TRY(G(att, b.proj_hl, M, D, D, EPI_RESID_F32, bf, x, D).bias(b.proj_b).gamma(b.ls1).resid(x).pitch(2 * D, 2 * D)
                 .split(GEMM_SPLIT_AW).run(st));
      KB_TRY(launch_layernorm(x, D, M, D, b.n2w, b.n2b, c.vit_ln_eps, xn, bf, nullptr, st, 2 * D, D));
      KB_TRY(G(xn, b.fc1_hl, M, F, D, EPI_BIAS_GELU_HILO, bf, hid, 2 * F).bias(b.fc1_b).pitch(2 * D, 2 * D)
                 .split(GEMM_SPLIT_AW).lo(F).run(st));
      KB_TRY(G(hid, b.fc2_hl, M, D, F, EPI_RESID_F32, bf, x, D).bias(b.fc2_b).gamma(b.ls2).resid(x).pitch(2 * F, 2 * F)
                 .split(GEMM_SPLIT_AW).run(st));
    } else if (i + 1 < c.vit_depth && wsplit) {
      KB_TRY(G(att, b.proj_hl, M, D, D, EPI_RESID_F32, bf, x, D).bias(b.proj_b).gamma(b.ls1).resid(x).pitch(D, 2 * D)
                 .split(GEMM_SPLIT_W).run(st));
      KB_TRY(launch_layernorm(x, D, M, D, b.n2w, b.n2b, c.vit_ln_eps, xn, bf, nullptr, st));
      KB_TRY(G(xn, b.fc1_hl, M, F, D, EPI_BIAS_GELU_HALF, bf, hid, F).bias(b.fc1_b).pitch(D, 2 * D).split(GEMM_SPLIT_W).run(st));
      KB_TRY(G(hid, b.fc2_hl, M, D, F, EPI_RESID_F32, bf, x, D).bias(b.fc2_b).gamma(b.ls2).resid(x).pitch(F, 2 * F)
                 .split(GEMM_SPLIT_W).run(st));
    } else if (i + 1 < c.vit_depth && fuse == 2) {
      KB_TRY(gemm_resid_stats(att, b.proj_w, M, D, D, bf, b.proj_b, b.ls1, x, xn, stats, st));
      KB_TRY(gemm_ln(xn, D, b.fc1_wf, M, F, EPI_LN_BIAS_GELU_HALF, bf, b.fc1_c, b.fc1_s, stats, c.vit_ln_eps, hid, st));
      KB_TRY(gemm_resid_stats(hid, b.fc2_w, M, D, F, bf, b.fc2_b, b.ls2, x, xn, stats, st));
    } else if (i + 1 < c.vit_depth) {
      // proj is HBM-bound (fp32 residual read-modify-write): the extra 16-bit copy costs it what the LayerNorm kernel
      // costs, so norm2 stays a kernel; fc2 (K = 4096, compute-bound) emits the copy and the statistics for free
      KB_TRY(G(att, b.proj_w, M, D, D, EPI_RESID_F32, bf, x, D).bias(b.proj_b).gamma(b.ls1).resid(x).run(st));
      KB_TRY(launch_layernorm(x, D, M, D, b.n2w, b.n2b, c.vit_ln_eps, xn, bf, nullptr, st));
      KB_TRY(G(xn, b.fc1_w, M, F, D, EPI_BIAS_GELU_HALF, bf, hid, F).bias(b.fc1_b).run(st));
      if (fuse == 1) KB_TRY(gemm_resid_stats(hid, b.fc2_w, M, D, F, bf, b.fc2_b, b.ls2, x, xn, stats, st));
      else KB_TRY(G(hid, b.fc2_w, M, D, F, EPI_RESID_F32, bf, x, D).bias(b.fc2_b).gamma(b.ls2).resid(x).run(st));
    } else {
      // Last block: only the CLS token is consumed downstream (global_pool='token'), and everything after the
      // attention is row-wise, so proj / norm2 / fc1 / fc2 run on the n CLS rows (row pitch T*D) only. At n rows the
      // MMA time is nothing, so these GEMMs always run split-operand (hi|lo weights, hi|lo LayerNorm / GELU outputs):
      // the tail adds no 2^-11 operand rounding of its own to the embedding.
      const int64_t pitch = (int64_t)T * D;
      if (high)
        KB_TRY(G(att, b.proj_hl, (int)n, D, D, EPI_RESID_F32, bf, xc, D).bias(b.proj_b).gamma(b.ls1).resid(x, pitch)
                   .pitch(2 * pitch, 2 * D).split(GEMM_SPLIT_AW).run(st));
      else
        KB_TRY(G(att, b.proj_hl, (int)n, D, D, EPI_RESID_F32, bf, xc, D).bias(b.proj_b).gamma(b.ls1).resid(x, pitch)
                   .pitch(pitch, 2 * D).split(GEMM_SPLIT_W).run(st));
      KB_TRY(launch_layernorm(xc, D, n, D, b.n2w, b.n2b, c.vit_ln_eps, cls16, bf, nullptr, st, 2 * D, D));
      KB_TRY(G(cls16, b.fc1_hl, (int)n, F, D, EPI_BIAS_GELU_HILO, bf, hid, 2 * F).bias(b.fc1_b).pitch(2 * D, 2 * D)
                 .split(GEMM_SPLIT_AW).lo(F).run(st));
      KB_TRY(G(hid, b.fc2_hl, (int)n, D, F, EPI_RESID_F32, bf, xc, D).bias(b.fc2_b).gamma(b.ls2).resid(xc)
                 .pitch(2 * F, 2 * F).split(GEMM_SPLIT_AW).run(st));
    }
    if (i + 1 < c.vit_depth) KB_TRY(dump_layer(m, i, (size_t)M * D, x, (size_t)M * D, st));
    else KB_TRY(dump_layer(m, i, (size_t)M * D, xc, (size_t)n * D, st));
  }
  // final norm on the CLS rows (global_pool='token') -> visual_head -> L2-normalise: one fp32 kernel (head.cu)
  return launch_visual_head(xc, D, n, D, m->norm_w, m->norm_b, c.vit_ln_eps, m->h0_wt, m->h0_b, c.proj_dim, m->h2_wt,
                            m->h2_b, c.proj_dim, out, st);
}

// precise: every GEMM of the tower runs split-operand on hi|lo activations and weights (3 MMA passes, ~22-bit operands);
// what is left of the 16-bit rounding is the attention's q/k/v/P (see oracle/precision_model.py)
int encode_text_chunk(Model* m, const int64_t* ids, const int64_t* tts, const int64_t* mask, int64_t n, int64_t S,
                      int64_t se, int level, float* out, char* ws, cudaStream_t st) {
  const bool precise = level == LEVEL_HIGH;
  const KeepB200Config& c = m->cfg;
  const int bf = c.operand_dtype, d = c.hidden, I = c.intermediate;
  const int M = (int)(n * se);
  const TextWs w = text_ws(m, n, se);
  float* x32 = reinterpret_cast<float*>(ws + w.x32);
  void* x16 = ws + w.x16;
  void* qkv = ws + w.qkv;
  void* att = ws + w.att;
  void* hid = ws + w.hid;
  float* xc32 = reinterpret_cast<float*>(ws + w.xc32);
  void* xc16 = ws + w.xc16;
  const int sp = precise ? GEMM_SPLIT_AW : level == LEVEL_BALANCED ? GEMM_SPLIT_W : GEMM_SPLIT_NONE;
  const int gelu = precise ? EPI_BIAS_GELU_HILO : EPI_BIAS_GELU_HALF;
  const int64_t lo_d = precise ? d : 0, lo_I = precise ? I : 0;  // offsets of the lo halves (0 = not written)
  KB_TRY(launch_bert_embed(ids, tts, S, n, (int)se, d, m->word, m->ttype, m->tpos, m->emb_lnw, m->emb_lnb,
                           c.bert_ln_eps, x32, x16, bf, c.vocab_size, c.type_vocab, st, 2 * d, lo_d));
  for (int i = 0; i < c.layers; ++i) {
    const BertLayer& L = m->layers[i];
    KB_TRY(G(x16, L.qkv_w, M, 3 * d, d, EPI_BIAS_HALF, bf, qkv, 3 * d).bias(L.qkv_b).pitch(2 * d, 2 * d).split(sp).run(st));
    KB_TRY(launch_attention(qkv, att, (int)n, (int)se, c.heads, bf, mask, S, 0.125f, st, 2 * d, lo_d));
    // post-LN: x = LN(x + dense(ctx)) ; x = LN(x + dense(gelu(dense(x))))
    if (i + 1 < c.layers) {
      KB_TRY(G(att, L.ao_w, M, d, d, EPI_RESID_F32, bf, x32, d).bias(L.ao_b).resid(x32).pitch(2 * d, 2 * d).split(sp).run(st));
      KB_TRY(launch_layernorm(x32, d, M, d, L.ao_lnw, L.ao_lnb, c.bert_ln_eps, x16, bf, x32, st, 2 * d, lo_d));
      KB_TRY(G(x16, L.in_w, M, I, d, gelu, bf, hid, 2 * I).bias(L.in_b).pitch(2 * d, 2 * d).split(sp).lo(lo_I).run(st));
      KB_TRY(G(hid, L.out_w, M, d, I, EPI_RESID_F32, bf, x32, d).bias(L.out_b).resid(x32).pitch(2 * I, 2 * I).split(sp).run(st));
      KB_TRY(launch_layernorm(x32, d, M, d, L.out_lnw, L.out_lnb, c.bert_ln_eps, x16, bf, x32, st, 2 * d, lo_d));
      KB_TRY(dump_layer(m, i, (size_t)M * d, x32, (size_t)M * d, st));
    } else {
      // last layer: the pooler reads only the [CLS] row and everything after the attention is row-wise
      const int64_t pitch = (int64_t)se * d;
      KB_TRY(G(att, L.ao_w, (int)n, d, d, EPI_RESID_F32, bf, xc32, d).bias(L.ao_b).resid(x32, pitch)
                 .pitch(2 * pitch, 2 * d).split(sp).run(st));
      KB_TRY(launch_layernorm(xc32, d, n, d, L.ao_lnw, L.ao_lnb, c.bert_ln_eps, xc16, bf, xc32, st, 2 * d, lo_d));
      KB_TRY(G(xc16, L.in_w, (int)n, I, d, gelu, bf, hid, 2 * I).bias(L.in_b).pitch(2 * d, 2 * d).split(sp).lo(lo_I).run(st));
      KB_TRY(G(hid, L.out_w, (int)n, d, I, EPI_RESID_F32, bf, xc32, d).bias(L.out_b).resid(xc32).pitch(2 * I, 2 * I).split(sp).run(st));
      KB_TRY(launch_layernorm(xc32, d, n, d, L.out_lnw, L.out_lnb, c.bert_ln_eps, nullptr, bf, xc32, st));
      KB_TRY(dump_layer(m, i, (size_t)M * d, xc32, (size_t)n * d, st));
    }
  }
  // pooler on the [CLS] rows: tanh(dense(x)) -> L2-normalise, one fp32 kernel (head.cu)
  return launch_pooler(xc32, d, n, d, m->pool_wt, m->pool_b, out, st);
}

}  // namespace

// =============================================================================================================
// exported C ABI
// =============================================================================================================
extern "C" {

int keepb200_version(void) { return KEEPB200_ABI_VERSION; }
const char* keepb200_last_error(void) { return last_error(); }

int keepb200_create(const KeepB200Config* cfg, int device, void** handle) {
  if (!cfg || !handle) return set_error(KB_ERR_ARG, "create: null argument");
  *handle = nullptr;
  KB_TRY(validate_cfg(*cfg));
  KB_TRY(check_device(device));
  KB_CUDA_CHECK(cudaSetDevice(device));
  Model* m = new Model();
  m->cfg = *cfg;
  m->device = device;
  int rc = build_tables(m);
  if (rc != KB_OK) {
    for (void* p : m->allocs) cudaFree(p);
    delete m;
    return rc;
  }
  *handle = m;
  return KB_OK;
}

void keepb200_destroy(void* handle) {
  if (!handle) return;
  Model* m = static_cast<Model*>(handle);
  for (void* p : m->allocs) cudaFree(p);
  delete m;
}

int keepb200_num_weights(void* handle) {
  if (!handle) return set_error(KB_ERR_ARG, "null handle");
  return (int)static_cast<Model*>(handle)->slots.size();
}
const char* keepb200_weight_name(void* handle, int index) {
  if (!handle) return nullptr;
  Model* m = static_cast<Model*>(handle);
  if (index < 0 || index >= (int)m->slots.size()) return nullptr;
  return m->slots[index].name.c_str();
}

int keepb200_load_weight(void* handle, const char* name, const float* data, const int64_t* shape, int ndim,
                         void* stream) {
  if (!handle || !name) return set_error(KB_ERR_ARG, "load_weight: null argument");
  Model* m = static_cast<Model*>(handle);
  auto it = m->index.find(name);
  if (it == m->index.end()) {
    // buffers some transformers versions serialise; not parameters
    if (std::strcmp(name, "text.embeddings.position_ids") == 0 || std::strcmp(name, "text.embeddings.token_type_ids") == 0)
      return KB_OK;
    return set_error(KB_ERR_ARG, "load_weight: unexpected key \"%s\"", name);
  }
  WeightSlot& s = m->slots[it->second];
  if (s.ignored) {
    s.loaded = true;
    return KB_OK;
  }
  if (!data) return set_error(KB_ERR_ARG, "load_weight: null data for \"%s\"", name);
  bool same = (ndim == (int)s.shape.size());
  for (int i = 0; same && i < ndim; ++i) same = (shape[i] == s.shape[i]);
  if (!same) {
    std::string want, got;
    for (auto d : s.shape) want += std::to_string(d) + ",";
    for (int i = 0; i < ndim; ++i) got += std::to_string(shape[i]) + ",";
    return set_error(KB_ERR_ARG, "load_weight: size mismatch for \"%s\": expected [%s] got [%s]", name, want.c_str(),
                     got.c_str());
  }
  size_t n = 1;
  for (auto d : s.shape) n *= (size_t)d;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (s.master != nullptr) KB_CUDA_CHECK(cudaMemcpyAsync(s.master, data, n * 4, cudaMemcpyDeviceToDevice, st));
  if (s.dst_hl != nullptr)
    KB_TRY(launch_cast_f32_to_hilo(data, s.dst_hl, s.shape[0], (int)(n / (size_t)s.shape[0]), m->cfg.operand_dtype, st));
  if (s.as16 && s.hilo)
    KB_TRY(launch_cast_f32_to_hilo(data, s.dst, s.shape[0], (int)(n / (size_t)s.shape[0]), m->cfg.operand_dtype, st));
  else if (s.as16)
    KB_TRY(launch_cast_f32_to_16(data, s.dst, (int64_t)n, m->cfg.operand_dtype, st));
  else
    KB_CUDA_CHECK(cudaMemcpyAsync(s.dst, data, n * 4, cudaMemcpyDeviceToDevice, st));
  s.loaded = true;
  m->finalized = false;
  return KB_OK;
}

int keepb200_finalize(void* handle) {
  if (!handle) return set_error(KB_ERR_ARG, "null handle");
  Model* m = static_cast<Model*>(handle);
  std::string missing;
  int count = 0;
  for (const WeightSlot& s : m->slots)
    if (!s.loaded) {
      if (count < 4) missing += "\"" + s.name + "\" ";
      ++count;
    }
  if (count) return set_error(KB_ERR_STATE, "finalize: %d missing key(s): %s%s", count, missing.c_str(), count > 4 ? "..." : "");
  // derive the LayerNorm-folded operands (norm1 -> qkv, norm2 -> fc1) from the fp32 masters; the weights were copied
  // on caller streams, so fence the device on both sides
  KB_CUDA_CHECK(cudaDeviceSynchronize());
  {
    const KeepB200Config& c = m->cfg;
    const int D = c.vit_width, F = c.vit_mlp, bf = c.operand_dtype;
    for (int i = 0; i < c.vit_depth; ++i) {
      VitBlock& b = m->blocks[i];
      const std::string p = "visual.blocks." + std::to_string(i) + ".";
      const float* qkv32 = m->slots[m->index[p + "attn.qkv.weight"]].master;
      const float* fc132 = m->slots[m->index[p + "mlp.fc1.weight"]].master;
      KB_TRY(launch_fold_ln(qkv32, 3 * D, D, b.n1w, b.n1b, b.qkv_b, b.qkv_wf, bf, b.qkv_s, b.qkv_c, nullptr));
      KB_TRY(launch_fold_ln(fc132, F, D, b.n2w, b.n2b, b.fc1_b, b.fc1_wf, bf, b.fc1_s, b.fc1_c, nullptr));
    }
    // K-major-transposed fp32 copies for the fused fp32 tails (head.cu)
    KB_TRY(launch_transpose_f32(m->h0_w, m->h0_wt, c.proj_dim, D, nullptr));
    KB_TRY(launch_transpose_f32(m->h2_w, m->h2_wt, c.proj_dim, c.proj_dim, nullptr));
    KB_TRY(launch_transpose_f32(m->pool_w, m->pool_wt, c.hidden, c.hidden, nullptr));
  }
  KB_CUDA_CHECK(cudaDeviceSynchronize());
  m->finalized = true;
  return KB_OK;
}

size_t keepb200_workspace_bytes(void* handle, int op, int64_t n, int64_t seq_len) {
  if (!handle || n <= 0) return 0;
  Model* m = static_cast<Model*>(handle);
  if (op == KEEPB200_OP_ENCODE_IMAGE) return image_ws(m, n, m->grid(), m->grid(), false).total;
  if (op == KEEPB200_OP_ENCODE_IMAGE_HIGH) return image_ws(m, n, m->grid(), m->grid(), true).total;
  if (op == KEEPB200_OP_ENCODE_TEXT) return text_ws(m, n, seq_len > 0 ? seq_len : m->cfg.max_pos).total;
  return 0;
}

static int check_hw(const Model* m, int64_t H, int64_t W, int* gh, int* gw) {
  if (H <= 0 || W <= 0 || H % 16 != 0 || W % 16 != 0)
    return set_error(KB_ERR_ARG, "encode_image: tile size %lldx%lld must be a positive multiple of the 16-pixel patch",
                     (long long)H, (long long)W);
  *gh = (int)(H / 16);
  *gw = (int)(W / 16);
  if ((long long)*gh * *gw + 1 > 512)
    return set_error(KB_ERR_ARG, "encode_image: %lldx%lld gives %lld tokens; the attention kernels serve <= 512",
                     (long long)H, (long long)W, (long long)*gh * *gw + 1);
  (void)m;
  return KB_OK;
}

size_t keepb200_workspace_bytes_hw(void* handle, int64_t n, int64_t H, int64_t W, int high) {
  if (!handle || n <= 0) return 0;
  Model* m = static_cast<Model*>(handle);
  int gh, gw;
  if (check_hw(m, H, W, &gh, &gw) != KB_OK) return 0;
  return image_ws(m, n, gh, gw, high != 0).total;
}

int keepb200_image_precision_is_high(int precision, int64_t B) {
  return precision == KEEPB200_PRECISION_HIGH || (precision == KEEPB200_PRECISION_AUTO && B <= KEEPB200_IMAGE_AUTO_MAX_TILES);
}

int keepb200_encode_image_hw(void* handle, const void* tiles, int layout, int64_t B, int64_t H, int64_t W, int precision,
                             float* out, void* workspace, size_t workspace_bytes, void* stream) {
  if (!handle) return set_error(KB_ERR_ARG, "null handle");
  Model* m = static_cast<Model*>(handle);
  if (!m->finalized) return set_error(KB_ERR_STATE, "encode_image: handle not finalised");
  if (B == 0) return KB_OK;
  if (B < 0 || !tiles || !out) return set_error(KB_ERR_ARG, "encode_image: bad arguments");
  if (layout != KEEPB200_TILES_F32_NCHW && layout != KEEPB200_TILES_U8_NHWC)
    return set_error(KB_ERR_ARG, "encode_image: unknown tile layout %d", layout);
  if (precision < KEEPB200_PRECISION_AUTO || precision > KEEPB200_PRECISION_BALANCED)
    return set_error(KB_ERR_ARG, "encode_image: unknown precision %d", precision);
  // AUTO is a function of the call's tile count only (never of the workspace or the chunking)
  const bool high = keepb200_image_precision_is_high(precision, B) != 0;
  const int level = high ? LEVEL_HIGH : precision == KEEPB200_PRECISION_BALANCED ? LEVEL_BALANCED : LEVEL_FAST;
  int gh, gw;
  KB_TRY(check_hw(m, H, W, &gh, &gw));
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return set_error(KB_ERR_ARG, "encode_image: workspace must be 1024-byte aligned");
  const size_t per1 = image_ws(m, 1, gh, gw, high).total;
  if (!workspace || workspace_bytes < per1)
    return set_error(KB_ERR_WORKSPACE, "encode_image: workspace %zu B < %zu B needed for one tile", workspace_bytes, per1);
  // largest chunk that fits (the layout is monotone in n); rows are limited to int32 GEMM extents
  const int64_t T = (int64_t)gh * gw + 1;
  int64_t chunk = B;
  const int64_t max_rows = (int64_t)1 << 30;
  if (chunk * T > max_rows) chunk = max_rows / T;
  while (chunk > 1 && image_ws(m, chunk, gh, gw, high).total > workspace_bytes) {
    int64_t guess = (int64_t)(workspace_bytes / (image_ws(m, chunk, gh, gw, high).total / (double)chunk));
    chunk = guess < chunk ? (guess > 1 ? guess : 1) : chunk - 1;
  }
  const size_t tile_elems = (size_t)3 * H * W;
  const size_t tile_bytes = tile_elems * (layout == KEEPB200_TILES_F32_NCHW ? 4 : 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t b0 = 0; b0 < B; b0 += chunk) {
    const int64_t n = (B - b0 < chunk) ? (B - b0) : chunk;
    KB_TRY(encode_image_chunk(m, static_cast<const char*>(tiles) + (size_t)b0 * tile_bytes, layout, n, gh, gw, level,
                              out + (size_t)b0 * m->cfg.proj_dim, static_cast<char*>(workspace), st));
  }
  return KB_OK;
}

int keepb200_encode_image(void* handle, const void* tiles, int layout, int64_t B, int precision, float* out,
                          void* workspace, size_t workspace_bytes, void* stream) {
  if (!handle) return set_error(KB_ERR_ARG, "null handle");
  const Model* m = static_cast<Model*>(handle);
  return keepb200_encode_image_hw(handle, tiles, layout, B, m->cfg.img_size, m->cfg.img_size, precision, out, workspace,
                                  workspace_bytes, stream);
}

int keepb200_encode_text(void* handle, const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int64_t P,
                         int64_t S, int64_t s_eff, int precision, float* out, void* workspace, size_t workspace_bytes,
                         void* stream) {
  if (!handle) return set_error(KB_ERR_ARG, "null handle");
  Model* m = static_cast<Model*>(handle);
  if (!m->finalized) return set_error(KB_ERR_STATE, "encode_text: handle not finalised");
  if (P == 0) return KB_OK;
  if (P < 0 || !ids || !out) return set_error(KB_ERR_ARG, "encode_text: bad arguments");
  if (S < 1 || S > m->cfg.max_pos || S > 512)
    return set_error(KB_ERR_ARG, "encode_text: sequence length %lld outside [1, %d]", (long long)S,
                     m->cfg.max_pos < 512 ? m->cfg.max_pos : 512);
  if (s_eff < 1 || s_eff > S) return set_error(KB_ERR_ARG, "encode_text: s_eff %lld outside [1, %lld]", (long long)s_eff, (long long)S);
  if (precision < KEEPB200_PRECISION_AUTO || precision > KEEPB200_PRECISION_BALANCED)
    return set_error(KB_ERR_ARG, "encode_text: unknown precision %d", precision);
  // AUTO is a function of the call's prompt count only (never of the workspace or the chunking)
  const bool precise = precision == KEEPB200_PRECISION_HIGH || (precision == KEEPB200_PRECISION_AUTO && P <= KEEPB200_TEXT_AUTO_MAX_PROMPTS);
  const int level = precise ? LEVEL_HIGH : precision == KEEPB200_PRECISION_BALANCED ? LEVEL_BALANCED : LEVEL_FAST;
  if ((reinterpret_cast<uintptr_t>(workspace) & 1023) != 0) return set_error(KB_ERR_ARG, "encode_text: workspace must be 1024-byte aligned");
  const size_t per1 = text_ws(m, 1, s_eff).total;
  if (!workspace || workspace_bytes < per1)
    return set_error(KB_ERR_WORKSPACE, "encode_text: workspace %zu B < %zu B needed for one prompt", workspace_bytes, per1);
  int64_t chunk = P;
  const int64_t max_rows = (int64_t)1 << 30;
  if (chunk * s_eff > max_rows) chunk = max_rows / s_eff;
  while (chunk > 1 && text_ws(m, chunk, s_eff).total > workspace_bytes) {
    int64_t guess = (int64_t)(workspace_bytes / (text_ws(m, chunk, s_eff).total / (double)chunk));
    chunk = guess < chunk ? (guess > 1 ? guess : 1) : chunk - 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int64_t p0 = 0; p0 < P; p0 += chunk) {
    const int64_t n = (P - p0 < chunk) ? (P - p0) : chunk;
    KB_TRY(encode_text_chunk(m, ids + p0 * S, type_ids ? type_ids + p0 * S : nullptr, mask ? mask + p0 * S : nullptr, n, S,
                             s_eff, level, out + (size_t)p0 * m->cfg.hidden, static_cast<char*>(workspace), st));
  }
  return KB_OK;
}

size_t keepb200_preprocess_workspace_bytes(int64_t B, int64_t H, int64_t W, int size) {
  return preprocess_workspace_bytes(B, H, W, size);
}
int keepb200_preprocess_u8(const uint8_t* tiles, int64_t B, int64_t H, int64_t W, int size, uint8_t* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
  return launch_preprocess_u8(tiles, B, H, W, size, out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t keepb200_similarity_workspace_bytes(int64_t D, int64_t P) { return (D > 0 && P > 0) ? (size_t)D * P * 4 : 0; }

int keepb200_similarity(const float* feats, int64_t N, int64_t D, const float* cls, int64_t P, int group, float temp,
                        float* logits, float* probs, void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0 || P == 0) return KB_OK;
  if (!feats || !cls || N < 0 || P < 0 || D <= 0) return set_error(KB_ERR_ARG, "similarity: bad arguments");
  if (P > (1 << 30)) return set_error(KB_ERR_ARG, "similarity: P too large");
  float* clsT = nullptr;
  if (workspace != nullptr && workspace_bytes >= (size_t)D * P * 4 && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0)
    clsT = static_cast<float*>(workspace);
  return launch_similarity(feats, N, (int)D, cls, (int)P, group, temp, logits, probs, static_cast<cudaStream_t>(stream), clsT);
}

// workspace of the default path: K-major classifier copy | per-row-block score partials | rows x P logits
static size_t ps_cls_bytes(int64_t D, int64_t P) { return ((size_t)P * D * 4 + 1023) / 1024 * 1024; }
static size_t ps_part_bytes(int64_t N, int64_t K) { return ((size_t)((N + 63) / 64) * K * 4 + 1023) / 1024 * 1024; }

size_t keepb200_prompt_scores_workspace_bytes(int64_t N, int64_t D, int64_t K, int64_t C) {
  if (N <= 0 || D <= 0 || K <= 0 || C <= 0) return 0;
  const size_t P = (size_t)K * C;
  const size_t rows = (size_t)((N + 63) / 64 * 64);
  return ps_cls_bytes(D, P) + ps_part_bytes(N, K) + rows * P * 4;
}

int keepb200_prompt_scores(const float* feats, int64_t N, int64_t D, const float* cls, int64_t K, int64_t C, int fused,
                           float* scores, void* workspace, size_t workspace_bytes, void* stream) {
  if (K == 0) return KB_OK;
  if (!feats || !cls || !scores || N <= 0 || K < 0 || C < 2 || D <= 0) return set_error(KB_ERR_ARG, "prompt_scores: bad arguments");
  const int64_t P = K * C;
  if (P >= (1 << 30)) return set_error(KB_ERR_ARG, "prompt_scores: K*C too large");
  if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return set_error(KB_ERR_ARG, "prompt_scores: a 16-byte aligned workspace is required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // fused != 0: the top-2 margin is reduced inside the similarity epilogue and the [N, K*C] logits are never written
  // (workspace: classifier copy + 16 B per tile and classifier instead of 4 B per tile and column). Measured slower than
  // the logits round trip (50k x 1782 x 4: 1.69 ms vs 1.49 ms), so it is the caller's choice for memory-constrained slides.
  if (fused) {
    if (!(C == 2 || C == 4 || C == 8 || C == 16) || D % 32 != 0 || (reinterpret_cast<uintptr_t>(feats) & 15) != 0)
      return set_error(KB_ERR_ARG, "prompt_scores: the fused path needs C in {2,4,8,16}, D %% 32 == 0 and 16-byte aligned features");
    return launch_prompt_scores_fused(feats, N, (int)D, cls, (int)K, (int)C, scores, workspace, workspace_bytes, st);
  }
  const size_t row_bytes = (size_t)P * 4;
  const size_t head = ps_cls_bytes(D, P) + ps_part_bytes(N, K);
  if (workspace_bytes < head + row_bytes * 64)
    return set_error(KB_ERR_WORKSPACE, "prompt_scores: workspace %zu B < %zu B (classifier copy + partials + 64 rows of logits)",
                     workspace_bytes, head + row_bytes * 64);
  float* clsT = static_cast<float*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<char*>(workspace) + ps_cls_bytes(D, P));
  float* logits = reinterpret_cast<float*>(static_cast<char*>(workspace) + head);
  int64_t chunk = (int64_t)((workspace_bytes - head) / row_bytes);
  chunk = chunk / 64 * 64;
  if (chunk > N) chunk = N;
  // every 256-row block of every chunk writes its own partial row; they are summed in a fixed order afterwards, so the
  // scores (and the ranking that zero_shot_prompt_select derives from them) are identical from run to run
  int64_t nparts = 0;
  for (int64_t r0 = 0; r0 < N; r0 += chunk) {
    const int64_t n = (N - r0 < chunk) ? (N - r0) : chunk;
    KB_TRY(launch_similarity(feats + r0 * D, n, (int)D, cls, (int)P, (int)C, 1.0f, logits, nullptr, st, clsT));
    KB_TRY(launch_prompt_score_partials(logits, n, (int)K, (int)C, part + nparts * K, st));
    nparts += (n + 255) / 256;
  }
  return launch_score_reduce(part, nparts, (int)K, 1.0f / (float)N, scores, st);
}

size_t keepb200_refine_workspace_bytes(int64_t N) { return N > 0 ? refine_workspace_bytes(N) : 0; }

int keepb200_refine(const int64_t* coords, const float* probs, int64_t N, int64_t C, int64_t patch_size, int overlap,
                    uint8_t* keep, float* refined, void* workspace, size_t workspace_bytes, void* stream) {
  if (N == 0) return KB_OK;
  if (!coords || !probs || !keep || !refined || N < 0 || C <= 0) return set_error(KB_ERR_ARG, "refine: bad arguments");
  return launch_refine(coords, probs, N, (int)C, patch_size, overlap, keep, refined, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

int keepb200_debug_set_ln_fuse(void* handle, int mode) {
  if (!handle || mode < 0 || mode > 2) return set_error(KB_ERR_ARG, "debug_set_ln_fuse: bad arguments");
  static_cast<Model*>(handle)->ln_fuse = mode;
  return KB_OK;
}
int keepb200_debug_layer_dump(void* handle, float* buf, size_t bytes) {
  if (!handle) return set_error(KB_ERR_ARG, "null handle");
  Model* m = static_cast<Model*>(handle);
  m->dump = buf;
  m->dump_bytes = buf ? bytes : 0;
  return KB_OK;
}

int keepb200_profile_begin(void) { return profile_begin(); }
int keepb200_profile_end(double* gemm_ms, double* gemm_flops, int64_t* gemm_launches, int64_t* all_launches) {
  long long gl = 0, al = 0;
  int rc = profile_end(gemm_ms, gemm_flops, &gl, &al);
  if (gemm_launches) *gemm_launches = gl;
  if (all_launches) *all_launches = al;
  return rc;
}
int64_t keepb200_launch_count(void) { return launch_count(); }
const char* keepb200_profile_table(void) { return profile_table(); }

// ---- single-kernel entry points -----------------------------------------------------------------------------------
int keepb200_op_gemm_split(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epi, int bf16,
                           int split, const float* bias, const float* resid, int64_t ldr, void* out, int64_t ldo,
                           int64_t lo_off, void* stream) {
  GemmArgs a;
  a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.M = M; a.N = N; a.K = K; a.epi = epi; a.bf16 = bf16;
  a.bias = bias; a.gamma = nullptr; a.resid = resid; a.ldr = ldr; a.out = out; a.ldo = ldo; a.pos = nullptr; a.patches = 0;
  a.split = split; a.lo_off = lo_off;
  return launch_gemm(a, static_cast<cudaStream_t>(stream));
}
int keepb200_op_cast_hilo(const float* src, void* dst, int64_t rows, int K, int bf16, void* stream) {
  return launch_cast_f32_to_hilo(src, dst, rows, K, bf16, static_cast<cudaStream_t>(stream));
}
int keepb200_op_visual_head(const float* x, int64_t ldx, int64_t n, int D, const float* lnw, const float* lnb, float eps,
                            const float* w0t, const float* b0, int N0, const float* w1t, const float* b1, int N1, float* out,
                            void* stream) {
  return launch_visual_head(x, ldx, n, D, lnw, lnb, eps, w0t, b0, N0, w1t, b1, N1, out, static_cast<cudaStream_t>(stream));
}
int keepb200_op_pooler(const float* x, int64_t ldx, int64_t n, int D, const float* wt, const float* b, float* out,
                       void* stream) {
  return launch_pooler(x, ldx, n, D, wt, b, out, static_cast<cudaStream_t>(stream));
}
int keepb200_op_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epi, int bf16,
                     const float* bias, const float* gamma, const float* resid, int64_t ldr, void* out, int64_t ldo,
                     const float* pos, int patches, void* stream) {
  GemmArgs a;
  a.A = A; a.lda = lda; a.W = W; a.ldw = ldw; a.M = M; a.N = N; a.K = K; a.epi = epi; a.bf16 = bf16;
  a.bias = bias; a.gamma = gamma; a.resid = resid; a.ldr = ldr; a.out = out; a.ldo = ldo; a.pos = pos;
  a.patches = patches;
  return launch_gemm(a, static_cast<cudaStream_t>(stream));
}
int keepb200_op_gemm_resid_stats(const void* A, const void* W, int M, int N, int K, int bf16, const float* bias,
                                 const float* gamma, float* x, void* x16, float* stats, void* stream) {
  return gemm_resid_stats(A, W, M, N, K, bf16, bias, gamma, x, x16, stats, static_cast<cudaStream_t>(stream));
}
int keepb200_op_gemm_ln(const void* x16, const void* Wf, int M, int N, int K, int gelu, int bf16, const float* c_vec,
                        const float* s_vec, const float* stats, float eps, void* out16, void* stream) {
  if (K % (2 * kLnSliceCols) != 0) return set_error(KB_ERR_ARG, "op_gemm_ln: K=%d must be a multiple of %d", K, 2 * kLnSliceCols);
  return gemm_ln(x16, K, Wf, M, N, gelu ? EPI_LN_BIAS_GELU_HALF : EPI_LN_BIAS_HALF, bf16, c_vec, s_vec, stats, eps, out16,
                 static_cast<cudaStream_t>(stream));
}
int keepb200_op_fold_ln(const float* W, int N, int K, const float* lnw, const float* lnb, const float* bias, void* W16,
                        int bf16, float* s, float* c, void* stream) {
  return launch_fold_ln(W, N, K, lnw, lnb, bias, W16, bf16, s, c, static_cast<cudaStream_t>(stream));
}
int keepb200_op_pos_resample(const float* pos, int G0, int Gh, int Gw, int D, float* out, void* stream) {
  return launch_pos_resample(pos, G0, Gh, Gw, D, out, static_cast<cudaStream_t>(stream));
}
int keepb200_op_layernorm(const float* x, int64_t row_stride, int64_t rows, int D, const float* w, const float* b,
                          float eps, void* y16, int bf16, float* y32, int64_t y16_pitch, int64_t lo_off, void* stream) {
  return launch_layernorm(x, row_stride, rows, D, w, b, eps, y16, bf16, y32, static_cast<cudaStream_t>(stream), y16_pitch,
                          lo_off);
}
int keepb200_op_attention(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                          int64_t mask_stride, float scale, int64_t out_pitch, int64_t lo_off, void* stream) {
  return launch_attention(qkv, out, B, S, H, bf16, key_mask, mask_stride, scale, static_cast<cudaStream_t>(stream), out_pitch,
                          lo_off);
}
int keepb200_debug_attention_trace(int64_t* dev_buf) {
  attention_tc_set_trace(reinterpret_cast<long long*>(dev_buf));
  return KB_OK;
}
int keepb200_op_act_l2norm(const float* x, int64_t rows, int D, int act, float* y, void* stream) {
  return launch_act_l2norm(x, rows, D, act, y, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
