/*
 * keep_b200 C ABI — the drop-in boundary of the B200-native KEEP zero-shot WSI inference path.
 *
 * The reference (MAGIC-AI4Med/KEEP) is pure Python/PyTorch and has no FFI of its own; what this
 * library replaces is the arithmetic behind three Python-level call sites:
 *
 *   KEEPModel.encode_image   quick_start/keep_inference.py:54-58   -> keepb200_encode_image
 *   KEEPModel.encode_text    quick_start/keep_inference.py:60-62   -> keepb200_encode_text
 *   tile x prompt similarity WSI_evaluation/detection_utils.py:90-93,
 *                            subtyping_utils.py:69-72, segment_utils.py:46-49,
 *                            quick_start/keep_inference.py:104      -> keepb200_similarity
 *   prompt screening         WSI_evaluation/utils.py:107-146       -> keepb200_prompt_scores
 *   refine_seg               detection_utils.py:39-74, subtyping_utils.py:38-65,
 *                            segment_utils.py:63-89                 -> keepb200_refine
 *
 * The Python host mirror (keep_b200/modeling_keep.py, keep_b200/wsi.py) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer that carries tensor data is a DEVICE pointer owned by the caller (e.g. a torch
 *     allocation), unless the parameter is documented as a host pointer;
 *   - all calls are asynchronous on the CUDA stream that is passed (a cudaStream_t cast to void*;
 *     NULL = the legacy default stream);
 *   - return value 0 = success, negative = error (KEEPB200_ERR_*); nothing throws across the boundary;
 *     keepb200_last_error() returns a thread-local message for the last failing call;
 *   - one handle per device and per thread of control (a handle is not thread-safe);
 *   - after keepb200_finalize() the library performs no hidden allocations: activations live in the
 *     workspace the caller passes, sized by keepb200_workspace_bytes();
 *   - no environment variable changes what is computed: numerics are selected by arguments only.
 *   - there is no CPU fallback anywhere: without a CUDA device of compute capability 10.x every
 *     compute entry point fails.
 */
#ifndef KEEP_B200_H_
#define KEEP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KEEPB200_ABI_VERSION 2

#define KEEPB200_OK 0
#define KEEPB200_ERR_ARG (-1)
#define KEEPB200_ERR_CUDA (-2)
#define KEEPB200_ERR_STATE (-3)
#define KEEPB200_ERR_WORKSPACE (-4)

/* operand dtype of the tensor-core path (accumulation, residual stream, LayerNorm, softmax are fp32) */
#define KEEPB200_FP16 0
#define KEEPB200_BF16 1

/* tile layouts accepted by keepb200_encode_image */
#define KEEPB200_TILES_F32_NCHW 0 /* float [B,3,H,W], already ImageNet-normalised (keep_inference.py:88-93) */
#define KEEPB200_TILES_U8_NHWC 1  /* uint8 [B,H,W,3] raw RGB; ToTensor+Normalize fused into the patch gather */

/* Precision of keepb200_encode_image / keepb200_encode_text (operands are 16-bit either way; accumulation, residual
 * stream, LayerNorm and softmax are always fp32).
 *   FAST: every GEMM makes one tensor-core pass over 16-bit operands: the 2^-11 operand rounding shows as ~1.0-1.2e-3
 *         rel-L2 on the image embedding and ~1.4e-3 on the text embedding against the fp32 reference (torch's own fp16
 *         autocast of the same ViT-L: 1.2e-3). This is the throughput path (tiles/s, prompt banks).
 *   HIGH: split-operand GEMMs: activations and weights are carried as hi + lo 16-bit pairs and every GEMM makes three
 *         passes into the same fp32 accumulator (Ah.Wh + Al.Wh + Ah.Wl): ~2.6e-4 / ~3e-4 rel-L2, at three times the MMA
 *         work. What stays 16-bit is the attention's q, k, v and probabilities (its context output is hi + lo too).
 *   BALANCED: the weights as hi + lo pairs, every activation as ONE 16-bit value, two passes (Ah.Wh + Ah.Wl): the weights'
 *         share of the operand rounding is gone (~7e-4 image / ~8e-4 text, i.e. inside 1e-3 with margin) at twice the MMA
 *         work instead of three times; LayerNorms run as stand-alone kernels (the folded weights exist in one 16-bit copy).
 *   AUTO: HIGH when the call is small enough for the extra passes not to matter - at most
 *         KEEPB200_IMAGE_AUTO_MAX_TILES tiles (quick-start / interactive use: latency-bound either way) or
 *         KEEPB200_TEXT_AUTO_MAX_PROMPTS prompts (every WSI classifier bank) - FAST otherwise. A function of the call's
 *         item count only. The CLS-row tail of the last ViT block and both heads are always computed at HIGH / fp32. */
#define KEEPB200_PRECISION_AUTO 0
#define KEEPB200_PRECISION_HIGH 1
#define KEEPB200_PRECISION_FAST 2
#define KEEPB200_PRECISION_BALANCED 3
#define KEEPB200_IMAGE_AUTO_MAX_TILES 16
#define KEEPB200_TEXT_AUTO_MAX_PROMPTS 8192

/* ops for keepb200_workspace_bytes */
#define KEEPB200_OP_ENCODE_IMAGE 0      /* FAST */
#define KEEPB200_OP_ENCODE_TEXT 1
#define KEEPB200_OP_ENCODE_IMAGE_HIGH 2 /* HIGH: hi|lo activations double the 16-bit buffers */

/* Model geometry. Mirrors KEEPConfig (keep_inference.py:9-22): vision_config is fixed by the timm call at
 * keep_inference.py:32-40 (ViT-L/16), text_config is the BertConfig dict, projection_dim = 768. */
typedef struct KeepB200Config {
  int32_t struct_size; /* = sizeof(KeepB200Config), for ABI evolution */
  /* vision tower */
  int32_t img_size;   /* 224 */
  int32_t patch_size; /* 16 (only 16 is supported) */
  int32_t vit_width;  /* 1024 */
  int32_t vit_depth;  /* 24 */
  int32_t vit_heads;  /* 16 (head dim must be 64) */
  int32_t vit_mlp;    /* 4096 */
  float vit_ln_eps;   /* 1e-6 */
  int32_t proj_dim;   /* 768 */
  /* text tower (BertConfig) */
  int32_t vocab_size;   /* 30522 */
  int32_t hidden;       /* 768 */
  int32_t layers;       /* 12 */
  int32_t heads;        /* 12 (head dim must be 64) */
  int32_t intermediate; /* 3072 */
  int32_t max_pos;      /* 512 */
  int32_t type_vocab;   /* 2 */
  float bert_ln_eps;    /* 1e-12 */
  /* numerics */
  int32_t operand_dtype; /* KEEPB200_FP16 (default; ~1e-3 vs the fp32 reference) or KEEPB200_BF16 */
} KeepB200Config;

int keepb200_version(void);
const char* keepb200_last_error(void);

/* ---- model handle -------------------------------------------------------------------------------- */
int keepb200_create(const KeepB200Config* cfg, int device, void** handle);
void keepb200_destroy(void* handle);

/* Load one tensor of the reference state-dict (names exactly as in KEEPModel.state_dict(), e.g.
 * "visual.blocks.0.attn.qkv.weight", "text.encoder.layer.3.attention.self.query.weight", "logit_scale").
 * `data` is a contiguous fp32 DEVICE tensor of the reference shape; it is converted/repacked into the
 * kernel layout (16-bit K-major GEMM operands, fused BERT q|k|v) and the caller keeps ownership.
 * Unknown names and shape mismatches are errors, as with load_state_dict(strict=True) (keep_inference.py:83). */
int keepb200_load_weight(void* handle, const char* name, const float* data, const int64_t* shape, int ndim,
                         void* stream);
/* number of state-dict tensors the handle expects / name of the i-th one (host strings) */
int keepb200_num_weights(void* handle);
const char* keepb200_weight_name(void* handle, int index);
/* strict check that every expected tensor was loaded; must be called before encode_* */
int keepb200_finalize(void* handle);

size_t keepb200_workspace_bytes(void* handle, int op, int64_t n, int64_t seq_len);

/* out[B, proj_dim] fp32, unit L2 norm = normalize(visual_head(ViT(tiles)))  (keep_inference.py:54-58).
 * The batch is processed in chunks as large as the workspace allows (>= 1 tile). */
int keepb200_encode_image(void* handle, const void* tiles, int layout, int64_t B, int precision, float* out,
                          void* workspace, size_t workspace_bytes, void* stream);
/* how AUTO resolves for a call of B tiles (1 = HIGH): callers size the workspace with it */
int keepb200_image_precision_is_high(int precision, int64_t B);
/* Tiles of another size, H and W multiples of 16 with (H/16)*(W/16) + 1 <= 512 tokens: the reference builds its ViT with
 * dynamic_img_size=True (keep_inference.py:39), i.e. timm resamples pos_embed to the new patch grid (bicubic, antialias,
 * prefix token kept). keepb200_encode_image == this call with H = W = img_size. */
size_t keepb200_workspace_bytes_hw(void* handle, int64_t n, int64_t H, int64_t W, int high);
int keepb200_encode_image_hw(void* handle, const void* tiles, int layout, int64_t B, int64_t H, int64_t W, int precision,
                             float* out, void* workspace, size_t workspace_bytes, void* stream);

/* The reference's input transform for raw RGB tiles of any size (keep_inference.py:88-90; WSI scripts :38-41):
 * Resize(size, BICUBIC) on the PIL image (short side -> size, aspect kept) then CenterCrop(size). tiles: uint8 [B,H,W,3],
 * out: uint8 [B,size,size,3], both device memory; the result is bit-identical to torchvision + Pillow (two-pass 8-bit
 * resampler with 22-bit fixed-point weights). ToTensor + Normalize follow inside keepb200_encode_image(…U8_NHWC).
 * Workspace (256-byte aligned) is only needed when a resize actually happens. */
size_t keepb200_preprocess_workspace_bytes(int64_t B, int64_t H, int64_t W, int size);
int keepb200_preprocess_u8(const uint8_t* tiles, int64_t B, int64_t H, int64_t W, int size, uint8_t* out, void* workspace,
                           size_t workspace_bytes, void* stream);

/* out[P, hidden] fp32, unit L2 norm = normalize(BertModel(ids, type_ids, mask).pooler_output)
 * (keep_inference.py:60-62). ids/type_ids/mask are int64 [P,S] row-major (type_ids or mask may be NULL:
 * zeros / ones). `s_eff` (1..S) is the number of leading positions to compute: positions >= s_eff must be
 * masked in every row, in which case the result is identical to the padded computation; pass S to disable.
 * precision: KEEPB200_PRECISION_* above. */
int keepb200_encode_text(void* handle, const int64_t* ids, const int64_t* type_ids, const int64_t* mask, int64_t P,
                         int64_t S, int64_t s_eff, int precision, float* out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- similarity (no handle) ------------------------------------------------------------------------ */
/* logits[N,P] = normalize(feats)[N,D] @ cls[D,P];  probs[N,P] = softmax(temp * logits) over each
 * consecutive group of `group` columns (0 = one group of all P columns, the reference's single
 * classifier [D,C]; K stacked classifiers of C classes use group = C). feats and cls are fp32, cls in the
 * reference's [D,P] layout (utils.py:83). Either output may be NULL (probs alone: group must divide 16 and the
 * tensor-core path be usable - the task heads only ever read the probabilities, and the kernel is HBM-bound on its
 * outputs). D % 4 == 0.
 * workspace: P*D*4 bytes (keepb200_similarity_workspace_bytes) for the K-major copy of the classifier used by
 * the TF32 tcgen05 kernel (D % 32 == 0); with workspace == NULL the fp32 FMA kernel runs instead. */
int keepb200_similarity(const float* feats, int64_t N, int64_t D, const float* cls, int64_t P, int group,
                        float temp, float* logits, float* probs, void* workspace, size_t workspace_bytes,
                        void* stream);
size_t keepb200_similarity_workspace_bytes(int64_t D, int64_t P);

/* Prompt screening (utils.py:107-146): for K classifiers of C classes stacked as cls[D, K*C], with
 * logits_k = normalize(feats) @ cls_k: scores[k] = mean_n( top1 - top2 - |top1 + top2 - 1| ). The reduction over tiles
 * runs in a fixed order (per-256-row partials, then one ordered sum): scores are bit-identical from run to run.
 * fused != 0 (C in {2,4,8,16}, D % 32 == 0): the margin is reduced inside the similarity epilogue and the [N, K*C]
 * logits never reach memory - less workspace, measured slower; the caller's choice. */
int keepb200_prompt_scores(const float* feats, int64_t N, int64_t D, const float* cls, int64_t K, int64_t C, int fused,
                           float* scores, void* workspace, size_t workspace_bytes, void* stream);
/* workspace for all rows at once: K-major classifier copy + score partials + N rows of logits; the call also accepts
 * less (down to 64 rows of logits: keepb200_prompt_scores_workspace_bytes(64, ...) + partials for N) and then chunks */
size_t keepb200_prompt_scores_workspace_bytes(int64_t N, int64_t D, int64_t K, int64_t C);

/* refine_seg (detection_utils.py:39-74 et al.): coords int64 [N,2] (x,y), probs fp32 [N,C].
 * First occurrence of a coordinate wins; with overlap != 0 every kept tile's probabilities are replaced by
 * the mean over the present tiles among (x-ps,y-ps),(x,y-ps),(x-ps,y),(x,y).
 * keep[N] uint8 = 1 for first occurrences; refined[N,C] fp32 (rows with keep==0 are zero).
 * Coordinates must lie in [-2^31, 2^31 - 2] (negative ones are ordinary keys, as in the reference's dict); a tile outside
 * that range is never stored (keep = 0) - the caller checks the range (keep_b200/ops.py::refine does, one host read). */
int keepb200_refine(const int64_t* coords, const float* probs, int64_t N, int64_t C, int64_t patch_size, int overlap,
                    uint8_t* keep, float* refined, void* workspace, size_t workspace_bytes, void* stream);
size_t keepb200_refine_workspace_bytes(int64_t N);

/* ---- measurement hooks (bench.py) ------------------------------------------------------------------------ */
/* Number of kernels this library has launched in this process (every launcher counts its launches). */
int64_t keepb200_launch_count(void);
/* Between begin and end every tcgen05 GEMM launch is bracketed with CUDA events on its launching stream.
 * end() synchronises the device and returns the summed GEMM device time (ms), the algorithmic FLOPs
 * (2*M*N*K per launch) of those launches, their count, and the count of ALL kernels launched in between. */
int keepb200_profile_begin(void);
int keepb200_profile_end(double* gemm_ms, double* gemm_flops, int64_t* gemm_launches, int64_t* all_launches);
/* Per-shape summary of the last profile_end(): "M,N,K,epi,launches,ms,TFLOP/s;" records (host string). */
const char* keepb200_profile_table(void);

/* ---- test / analysis hooks ----------------------------------------------------------------------------------------- */
/* LayerNorm placement in the ViT blocks: 0 stand-alone LayerNorm kernels, 1 (default) norm1 folded into the qkv GEMM,
 * 2 norm2 folded into fc1 as well. The parity tests run all three against the oracle. */
int keepb200_debug_set_ln_fuse(void* handle, int mode);
/* Per-layer residual-stream dump of the following encode_image / encode_text calls (each call must fit ONE workspace
 * chunk): slot i of [rows*width] floats receives the stream after block/layer i; the last slot holds the CLS rows only
 * ([n*width] floats). buf = NULL switches it off. Used by the per-layer parity table (tests/test_gpu_model.py). */
int keepb200_debug_layer_dump(void* handle, float* buf, size_t bytes);

/* ---- single-kernel entry points (unit tests and profiling) ------------------------------------------- */
/* out = epilogue(A[M,K] . W[N,K]^T); epi: 0 bias->16-bit, 1 bias+GELU(erf)->16-bit,
 * 2 resid + gamma*(acc+bias) -> fp32, 3 bias -> fp32, 4 ViT patch-embed scatter (+pos) -> fp32 */
int keepb200_op_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epi, int bf16,
                     const float* bias, const float* gamma, const float* resid, int64_t ldr, void* out, int64_t ldo,
                     const float* pos, int patches, void* stream);
/* Fused LayerNorm across two GEMMs (the pre-LN ViT block, timm Block.forward: x + ls(attn(norm1(x))), x + ls(mlp(norm2(x)))).
 *   op_gemm_resid_stats: x[M,N] += gamma * (A.W^T + bias); also x16 = 16-bit(x) and stats[M, N/64] = per-row (sum, sum of
 *                        squares) over each 64-column slice of the NEW x
 *   op_fold_ln:          W16 = 16-bit(W * lnw), s[n] = sum_k W16[n,k], c[n] = bias[n] + sum_k lnb[k] W[n,k]
 *   op_gemm_ln:          out16 = act(rstd_r * (x16.W16^T - mean_r * s) + c) with mean/rstd from `stats` (K = width of x16,
 *                        a multiple of 128); act = GELU(erf) when gelu != 0.  Equals act(Linear(LayerNorm(x))). */
int keepb200_op_gemm_resid_stats(const void* A, const void* W, int M, int N, int K, int bf16, const float* bias,
                                 const float* gamma, float* x, void* x16, float* stats, void* stream);
int keepb200_op_gemm_ln(const void* x16, const void* Wf, int M, int N, int K, int gelu, int bf16, const float* c_vec,
                        const float* s_vec, const float* stats, float eps, void* out16, void* stream);
int keepb200_op_fold_ln(const float* W, int N, int K, const float* lnw, const float* lnb, const float* bias, void* W16,
                        int bf16, float* s, float* c, void* stream);
/* pos_embed [1 + G0*G0, D] -> out [1 + Gh*Gw, D] (timm resample_abs_pos_embed: bicubic, antialias=True, 1 prefix token) */
int keepb200_op_pos_resample(const float* pos, int G0, int Gh, int Gw, int D, float* out, void* stream);
/* y16 rows have pitch y16_pitch (0 = D); lo_off > 0 also stores 16-bit(y - hi) at y16[i, lo_off + c] (hi|lo operand) */
int keepb200_op_layernorm(const float* x, int64_t row_stride, int64_t rows, int D, const float* w, const float* b,
                          float eps, void* y16, int bf16, float* y32, int64_t y16_pitch, int64_t lo_off, void* stream);
/* out rows have pitch out_pitch (0 = H*64); lo_off > 0 also stores the rounding remainder of the context */
int keepb200_op_attention(const void* qkv, void* out, int B, int S, int H, int bf16, const int64_t* key_mask,
                          int64_t mask_stride, float scale, int64_t out_pitch, int64_t lo_off, void* stream);
/* Split-operand GEMM (csrc/common.h GEMM_SPLIT_*): split 0 plain, 1 W = [N,2K] hi|lo (2 passes), 2 A = [M,2K] and
 * W = [N,2K] hi|lo (3 passes: Ah.Wh + Al.Wh + Ah.Wl). epi 8 = bias + GELU(erf) -> [hi | lo] output, lo at lo_off.
 * op_cast_hilo builds a hi|lo operand from fp32: dst[r, 0:K] = 16-bit(src), dst[r, K:2K] = 16-bit(src - hi). */
int keepb200_op_gemm_split(const void* A, int64_t lda, const void* W, int64_t ldw, int M, int N, int K, int epi, int bf16,
                           int split, const float* bias, const float* resid, int64_t ldr, void* out, int64_t ldo,
                           int64_t lo_off, void* stream);
int keepb200_op_cast_hilo(const float* src, void* dst, int64_t rows, int K, int bf16, void* stream);
/* Fused fp32 tails (csrc/head.cu), weights TRANSPOSED ([K, N] fp32):
 *   op_visual_head: out = normalize(W1 . gelu(W0 . LayerNorm(x) + b0) + b1)      keep_inference.py:42-46, 54-58
 *   op_pooler:      out = normalize(tanh(W . x + b))                              keep_inference.py:60-62 */
int keepb200_op_visual_head(const float* x, int64_t ldx, int64_t n, int D, const float* lnw, const float* lnb, float eps,
                            const float* w0t, const float* b0, int N0, const float* w1t, const float* b1, int N1, float* out,
                            void* stream);
int keepb200_op_pooler(const float* x, int64_t ldx, int64_t n, int D, const float* wt, const float* b, float* out,
                       void* stream);
/* Debug/profiling aid: when dev_buf (device int64 [64*16]) is non-NULL the tcgen05 attention kernel's CTA 0 records
 * clock64() stamps of its pipeline events for its first 64 work units; NULL switches tracing off. */
int keepb200_debug_attention_trace(int64_t* dev_buf);
int keepb200_op_act_l2norm(const float* x, int64_t rows, int D, int act, float* y, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* KEEP_B200_H_ */
