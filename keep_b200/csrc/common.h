// Host-side shared declarations for the library (error handling, tensor-map cache, kernel launchers).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

namespace kb {

// ---- error plumbing: every launcher returns 0 or a negative code and records a message ----------
enum : int {
  KB_OK = 0,
  KB_ERR_ARG = -1,      // bad argument / unsupported shape
  KB_ERR_CUDA = -2,     // CUDA runtime / driver error
  KB_ERR_STATE = -3,    // handle not finalised, missing weight, ...
  KB_ERR_WORKSPACE = -4 // workspace too small
};
int set_error(int code, const char* fmt, ...);
const char* last_error();
#define KB_CUDA_CHECK(expr)                                                                    \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ::kb::set_error(::kb::KB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                \
                             cudaGetErrorString(_e), __FILE__, __LINE__);                      \
  } while (0)

int num_sms();  // SM count of the current device

// cudaFuncSetAttribute(max dynamic shared memory) once per kernel and DEVICE (a process may drive several GPUs)
#define KB_TRY_ATTR(KERNEL, BYTES)                                                                              \
  do {                                                                                                          \
    static int done_[64] = {};                                                                                  \
    int dev_ = 0;                                                                                               \
    KB_CUDA_CHECK(cudaGetDevice(&dev_));                                                                        \
    if (dev_ < 0 || dev_ >= 64) return ::kb::set_error(::kb::KB_ERR_ARG, "device index %d out of range", dev_); \
    if (done_[dev_] < (int)(BYTES)) {                                                                           \
      KB_CUDA_CHECK(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BYTES)));   \
      done_[dev_] = (int)(BYTES);                                                                               \
    }                                                                                                           \
  } while (0)

// ---- launch accounting / live GEMM timing (bench.py: gpu_launches and roofline.achieved) ----------------
// Every launcher calls note_launch() once per kernel launch. While profiling is on, launch_gemm brackets each
// GEMM launch with CUDA events on the launching stream; profile_end() synchronises and sums them.
void note_launch(int n = 1);
long long launch_count();
bool profiling();
void profile_gemm_begin(cudaStream_t s);
void profile_gemm_end(cudaStream_t s, double flops);
void profile_gemm_tag(int M, int N, int K, int epi);  // call right before profile_gemm_begin
const char* profile_table();                          // per-shape summary of the last profile_end()
int profile_begin();
int profile_end(double* gemm_ms, double* gemm_flops, long long* gemm_launches, long long* all_launches);

// ---- TMA tensor maps ---------------------------------------------------------------------------
// 2-D row-major matrix [rows, cols] of 2- or 4-byte elements with row pitch `ld` elements, box =
// [box_rows, 128 bytes] and SWIZZLE_128B. Cached by value of all arguments.
enum : int { KB_F16 = 0, KB_BF16 = 1, KB_F32 = 2 };
int get_tmap_2d(const void* ptr, int dtype, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                CUtensorMap* out);

// ---- GEMM: out = epilogue(A[M,K] . W[N,K]^T) ------------------------------------------------------
enum : int {
  EPI_BIAS_HALF = 0,       // out16[r,c] = acc + bias[c]
  EPI_BIAS_GELU_HALF = 1,  // out16[r,c] = gelu_erf(acc + bias[c])
  EPI_RESID_F32 = 2,       // out32[r,c] = resid[r,c] + gamma[c] * (acc + bias[c])   (gamma may be null => 1)
  EPI_BIAS_F32 = 3,        // out32[r,c] = acc + bias[c]
  EPI_PATCH_F32 = 4,       // out32[(r/P)*(P+1)+1+r%P, c] = acc + bias[c] + pos[1+r%P, c]   (ViT patch embed)
  // LayerNorm fused across two GEMMs (the pre-LN ViT block): the residual GEMM also emits the 16-bit
  // copy of the new residual row and per-row partial sums; the next GEMM runs on that copy with the LayerNorm scale
  // folded into its weights (launch_fold_ln) and finishes the normalisation per output row in its epilogue.
  EPI_RESID_F32_STATS = 5,   // EPI_RESID_F32 + out16[r,c] = 16-bit(out32[r,c]); stats[r, c/64] = (sum, sum of squares)
  EPI_LN_BIAS_HALF = 6,      // out16[r,c] = rstd_r * (acc - mean_r * ln_s[c]) + bias[c]
  EPI_LN_BIAS_GELU_HALF = 7, // out16[r,c] = gelu_erf(rstd_r * (acc - mean_r * ln_s[c]) + bias[c])
  // EPI_BIAS_GELU_HALF that also stores the rounding remainder: out16[r, lo_off + c] = 16-bit(v - hi). The next GEMM reads the
  // row as a split operand [hi | lo] (GEMM_SPLIT_AW) and sees v to ~22 mantissa bits.
  EPI_BIAS_GELU_HILO = 8
};
// Split-operand GEMM: a 16-bit operand stored as [rows, 2K] = [hi | lo] with lo = 16-bit(v - hi) carries ~22 mantissa
// bits; the kernel makes extra passes over K into the same fp32 accumulator (Ah.Wh + Al.Wh + Ah.Wl; the lo.lo term is
// below fp32 round-off). Used where the MMA time is negligible and the 2^-11 operand rounding is what limits parity with
// the fp32 reference: the whole text tower for prompt sets of WSI size, the CLS-row tail of the last ViT block.
enum : int {
  GEMM_SPLIT_NONE = 0,  // A [M,K], W [N,K]
  GEMM_SPLIT_W = 1,     // A [M,K] (hi only), W [N,2K] hi|lo : 2 passes
  GEMM_SPLIT_AW = 2     // A [M,2K] hi|lo,    W [N,2K] hi|lo : 3 passes
};
constexpr int kLnSliceCols = 64;  // width of one partial-sum slice of the row statistics
struct GemmArgs {
  const void* A;  int64_t lda;   // [M,K] 16-bit, K contiguous
  const void* W;  int64_t ldw;   // [N,K] 16-bit, K contiguous (torch Linear weight)
  int M, N, K;
  int epi;
  int bf16;                      // 0: fp16 operands/outputs, 1: bf16
  const float* bias;             // [N] or null
  const float* gamma;            // [N] or null
  const float* resid; int64_t ldr;
  void* out; int64_t ldo;
  const float* pos; int patches; // EPI_PATCH_F32 only
  // EPI_RESID_F32_STATS (producer): 16-bit copy + row statistics [M, N/64] float2
  void* out16 = nullptr; int64_t ldo16 = 0;
  float* stats_out = nullptr;
  // EPI_LN_* (consumer): statistics of the A rows ([M, ln_slices] float2 over ln_width columns), folded column sums
  const float* ln_stats = nullptr; int ln_slices = 0; int ln_width = 0; float ln_eps = 0.f;
  const float* ln_s = nullptr;   // [N]  sum_k W'[n,k]   (bias carries b.W^T + bias)
  int split = GEMM_SPLIT_NONE;   // GEMM_SPLIT_*: lda / ldw are then the pitches of the [rows, 2K] hi|lo operands
  int64_t lo_off = 0;            // EPI_BIAS_GELU_HILO: element offset of the lo half inside an output row (ldo >= 2N)
};
int launch_gemm(const GemmArgs& a, cudaStream_t stream);

// ---- row kernels ------------------------------------------------------------------------------------
// y16[i,:] (and optionally y32[i,:]) = LayerNorm(x[i*row_stride : +D]) * w + b
// y16 rows have pitch y16_pitch elements (0 = D); lo_off > 0 also stores the rounding remainder 16-bit(y - hi) at
// y16[i, lo_off + c] (the [hi | lo] operand of a split GEMM)
int launch_layernorm(const float* x, int64_t x_row_stride, int64_t rows, int D, const float* w, const float* b,
                     float eps, void* y16, int bf16, float* y32, cudaStream_t stream, int64_t y16_pitch = 0,
                     int64_t lo_off = 0);
// act: 0 none, 1 tanh ; then y = x / max(||x||, 1e-12)
int launch_act_l2norm(const float* x, int64_t rows, int D, int act, float* y, cudaStream_t stream);

// ---- fused fp32 tails (head.cu): one launch each ----------------------------------------------------------------
// out[i,:] = normalize(W1 . gelu_erf(W0 . LayerNorm(x[i*ldx : +D]) + b0) + b1); w0t [D, N0], w1t [N0, N1] fp32 (transposed)
int launch_visual_head(const float* x, int64_t ldx, int64_t n, int D, const float* lnw, const float* lnb, float eps,
                       const float* w0t, const float* b0, int N0, const float* w1t, const float* b1, int N1, float* out,
                       cudaStream_t stream);
// out[i,:] = normalize(tanh(W . x[i*ldx : +D] + b)); wt [D, D] fp32 (transposed)
int launch_pooler(const float* x, int64_t ldx, int64_t n, int D, const float* wt, const float* b, float* out,
                  cudaStream_t stream);

// ---- attention ----------------------------------------------------------------------------------------
// qkv: [B*S, 3*H*64] 16-bit (q | k | v, head-major 64-wide groups); out: [B*S, H*64]
// key_mask: optional int64 [B, mask_stride] (non-zero = attend, 0 = masked key), as BERT's attention_mask
// out rows have pitch out_pitch elements (0 = H*64); lo_off > 0 also stores the rounding remainder of the context at
// out[i, lo_off + c] (the [hi | lo] operand of a split-operand output projection)
int launch_attention(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                     int64_t mask_stride, float scale, cudaStream_t stream, int64_t out_pitch = 0, int64_t lo_off = 0);

void attention_tc_set_trace(long long* dev_buf);  // debug aid: clock64 stamps of CTA 0, [64 units][16 events]
// single-tile variant for 64 < S <= 224 (attention_tc1.cu); launch_attention dispatches to it
bool attention_tc1_supports(int S);
int launch_attention_tc1(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                         int64_t mask_stride, float scale, cudaStream_t stream, int64_t out_pitch, int64_t lo_off,
                         long long* trace);

// ---- ViT front end ----------------------------------------------------------------------------------------
// tiles fp32 NCHW [B,3,Gh*16,Gw*16] -> patches16 [B*Gh*Gw, 768] (col = c*256+ky*16+kx); also writes the CLS
// rows x[b*(Gh*Gw+1), :] = cls + pos[0].
// hilo != 0: patches16 is [B*Gh*Gw, 1536] = [hi | rounding remainder] (split-operand patch embedding)
int launch_im2col(const float* tiles, int64_t B, int Gh, int Gw, void* patches16, int bf16, const float* cls,
                  const float* pos, float* x, int D, cudaStream_t stream, int hilo = 0);
// uint8 NHWC tiles [B,H,W,3] with fused (x/255-mean)/std
int launch_im2col_u8(const uint8_t* tiles, int64_t B, int Gh, int Gw, void* patches16, int bf16, const float* cls,
                     const float* pos, float* x, int D, cudaStream_t stream, int hilo = 0);
// pos_embed [1 + G0*G0, D] -> out [1 + Gh*Gw, D]: prefix row copied, grid rows resampled (bicubic, antialias)
int launch_pos_resample(const float* pos, int G0, int Gh, int Gw, int D, float* out, cudaStream_t stream);

// Resize(size, BICUBIC) + CenterCrop(size) of uint8 RGB tiles [B,H,W,3] -> [B,size,size,3], bit-identical to the
// torchvision-on-PIL transform of the reference (keep_inference.py:88-90); preprocess.cu
size_t preprocess_workspace_bytes(int64_t B, int64_t H, int64_t W, int size);
int launch_preprocess_u8(const uint8_t* tiles, int64_t B, int64_t H, int64_t W, int size, uint8_t* out, void* ws,
                         size_t ws_bytes, cudaStream_t stream);

// ---- BERT front end ------------------------------------------------------------------------------------------
// x32/x16[p*S+s,:] = LN(word[ids] + type[tt] + pos[s])
int launch_bert_embed(const int64_t* ids, const int64_t* tts, int64_t id_stride, int64_t P, int S, int D,
                      const float* word, const float* type, const float* pos, const float* lnw, const float* lnb,
                      float eps, float* x32, void* x16, int bf16, int vocab, int type_vocab, cudaStream_t stream,
                      int64_t x16_pitch = 0, int64_t lo_off = 0);  // as launch_layernorm: pitch of x16 rows, hi|lo remainder

// ---- similarity -------------------------------------------------------------------------------------------------
// logits[n,p] = <feats[n,:]/max(||feats[n]||,1e-12), cls[:,p]> ; probs = softmax over each consecutive
// group of `group` columns of (temp * logits). feats fp32 [N,D], cls fp32 [D,P] (torch layout, utils.py:83).
// clsT_scratch: optional device buffer of P*D floats; when given (and D % 32 == 0) the TF32 tcgen05 kernel is used.
int launch_similarity(const float* feats, int64_t N, int D, const float* cls, int P, int group, float temp,
                      float* logits, float* probs, cudaStream_t stream, float* clsT_scratch = nullptr);
int launch_similarity_tc(const float* feats, int64_t N, int D, const float* clsT, int P, int group, float temp,
                         float* logits, float* probs, bool* fused_probs, cudaStream_t stream, float* score_part = nullptr);
// fused prompt screening: scores[k] = mean_n(top1 - top2 - |top1 + top2 - 1|) over the C columns of classifier k, with the
// margin term reduced inside the similarity epilogue (C in {2,4,8,16}, D % 32 == 0). part: [ceil(N/128)*4, K] floats.
size_t prompt_scores_fused_workspace_bytes(int64_t N, int64_t D, int64_t K, int64_t C);
int launch_prompt_scores_fused(const float* feats, int64_t N, int D, const float* cls, int K, int C, float* scores,
                               void* ws, size_t ws_bytes, cudaStream_t stream);
// part[b, k] = sum over the b-th block of 256 rows of (top1 - top2 - |top1 + top2 - 1|) of logits[n, k*C:(k+1)*C];
// scores[k] = scale * sum_b part[b, k] in a fixed order (deterministic)
int launch_prompt_score_partials(const float* logits, int64_t rows, int K, int C, float* part, cudaStream_t stream);
int launch_score_reduce(const float* part, int64_t nparts, int K, float scale, float* scores, cudaStream_t stream);

// ---- refine_seg ------------------------------------------------------------------------------------------------
size_t refine_workspace_bytes(int64_t N);
int launch_refine(const int64_t* coords, const float* probs, int64_t N, int C, int64_t ps, int overlap, uint8_t* keep,
                  float* refined, void* ws, size_t ws_bytes, cudaStream_t stream);

// LayerNorm folding for EPI_LN_*: W'[n,k] = 16-bit(W[n,k] * lnw[k]); s[n] = sum_k W'[n,k] (of the ROUNDED values: it
// cancels against what the tensor core accumulates); c[n] = bias[n] + sum_k lnb[k] * W[n,k]
int launch_fold_ln(const float* W, int N, int K, const float* lnw, const float* lnb, const float* bias, void* W16,
                   int bf16, float* s, float* c, cudaStream_t stream);

// generic helpers
int launch_cast_f32_to_16(const float* src, void* dst, int64_t n, int bf16, cudaStream_t stream);
// src fp32 [rows, K] -> dst 16-bit [rows, 2K] = [hi | lo], lo = 16-bit(src - hi)
int launch_cast_f32_to_hilo(const float* src, void* dst, int64_t rows, int K, int bf16, cudaStream_t stream);
int launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t stream);  // dst[c,r]=src[r,c]

}  // namespace kb
