// Persistent warp-specialised GEMM for sm_100a:  out = epilogue(A[M,K] . W[N,K]^T)
//
//   * operands fp16 or bf16 (K-major, i.e. activations [M,K] row-major and torch Linear weights [N,K]),
//     fp32 accumulation in TMEM via tcgen05.mma.kind::f16 (UMMA 128 x BLOCK_N x 16);
//   * TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) stages 128x64 / BLOCK_Nx64 tiles through a
//     kStages-deep mbarrier ring; one producer thread, one MMA-issuing thread;
//   * the accumulator is double-buffered in TMEM (2 x BLOCK_N columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1; 8 epilogue warps drain it with tcgen05.ld and apply the
//     fused epilogue (bias / exact-erf GELU / LayerScale + fp32 residual / patch-embed scatter).
//
// This one kernel serves every dense layer on the path: ViT patch-embed, qkv, proj, fc1, fc2, the
// visual_head, and the BERT q|k|v, attention-output, intermediate, output and pooler projections
// (reference call sites: quick_start/keep_inference.py:32-46,49-50; SURVEY.md §2.3 K1,K3,K5-K8,K10,K11).
#include "common.h"
#include "ptx.cuh"

namespace kb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumEpiWarps = 8;
constexpr int kFirstEpiWarp = 4;
constexpr int kThreads = (kFirstEpiWarp + kNumEpiWarps) * 32;  // 384

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KB
  static constexpr int B_BYTES = BN * BLOCK_K * 2;       // 32 KB @ BN=256
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;  // 512 / 256: power of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024: manual alignment
};

struct KParams {
  int M, N, K;
  const float* bias;
  const float* gamma;
  const float* resid;
  long long ldr;
  void* out;
  long long ldo;
  const float* pos;
  int patches;
  uint32_t idesc;
  int bf16;
};

__device__ __forceinline__ float gelu_erf(float x) {
  // exact (erf) GELU as torch.nn.GELU() default: 0.5 x (1 + erf(x / sqrt(2)))
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ uint32_t pack16(float a, float b, int bf16) {
  if (bf16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const KParams p) {
  using C = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::STAGES * C::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  TMA -> MMA
  uint64_t* empty_bar = bars + C::STAGES;       // [STAGES]  MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * C::STAGES;   // [2]       MMA -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;         // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (p.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = p.K / BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kNumEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          tma_load_2d(&tmap_a, &full_bar[s], smem_a + s * C::A_BYTES, kb * BLOCK_K, m_blk * BLOCK_M);
          tma_load_2d(&tmap_b, &full_bar[s], smem_b + s * C::B_BYTES, kb * BLOCK_K, n_blk * BN);
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      int lt = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
        const int as = lt & 1;
        const uint32_t aph = (lt >> 1) & 1;
        mbar_wait(&tempty_bar[as], aph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint64_t da = make_smem_desc_sw128(smem_u32(smem_a + s * C::A_BYTES));
          const uint64_t db = make_smem_desc_sw128(smem_u32(smem_b + s * C::B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128-byte swizzle row: +2 in the >>4 address field
            umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, p.idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
          if (++s == C::STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);   // accumulator complete
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ===================== epilogue =====================
    const int q = warp & 3;                      // TMEM lane quadrant this warp may access
    const int half = (warp - kFirstEpiWarp) >> 2;  // which half of the BN columns
    constexpr int COLS_PER_WARP = BN / 2;
    int lt = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++lt) {
      const int m_blk = tile / n_tiles, n_blk = tile % n_tiles;
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      mbar_wait(&tfull_bar[as], aph);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      long long out_row = row;
      const float* pos_row = nullptr;
      if constexpr (EPI == EPI_PATCH_F32) {
        const int img = row / p.patches, pi = row % p.patches;
        out_row = (long long)img * (p.patches + 1) + 1 + pi;
        pos_row = p.pos + (long long)(1 + pi) * p.N;
      }
#pragma unroll 1
      for (int c0 = 0; c0 < COLS_PER_WARP; c0 += 32) {
        const int col_in_tile = half * COLS_PER_WARP + c0;
        const int col = n_blk * BN + col_in_tile;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN + col_in_tile), v);
        tmem_ld_wait();
        if (col >= p.N) continue;  // warp-uniform (N is a multiple of 32 for every layer on the path)
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col + j));
            f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
          }
        }
        if constexpr (EPI == EPI_BIAS_GELU_HALF) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
        }
        if constexpr (EPI == EPI_BIAS_HALF || EPI == EPI_BIAS_GELU_HALF) {
          if (row_ok) {
            uint16_t* o = reinterpret_cast<uint16_t*>(p.out) + out_row * p.ldo + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 w;
              w.x = pack16(f[j], f[j + 1], p.bf16);
              w.y = pack16(f[j + 2], f[j + 3], p.bf16);
              w.z = pack16(f[j + 4], f[j + 5], p.bf16);
              w.w = pack16(f[j + 6], f[j + 7], p.bf16);
              *reinterpret_cast<uint4*>(o + j) = w;
            }
          }
        } else {
          if constexpr (EPI == EPI_RESID_F32) {
            if (p.gamma != nullptr) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gamma + col + j));
                f[j] *= g4.x; f[j + 1] *= g4.y; f[j + 2] *= g4.z; f[j + 3] *= g4.w;
              }
            }
            if (row_ok) {
              const float* r = p.resid + (long long)row * p.ldr + col;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 r4 = *reinterpret_cast<const float4*>(r + j);
                f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
              }
            }
          }
          if constexpr (EPI == EPI_PATCH_F32) {
            if (row_ok) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(pos_row + col + j));
                f[j] += p4.x; f[j + 1] += p4.y; f[j + 2] += p4.z; f[j + 3] += p4.w;
              }
            }
          }
          if (row_ok) {
            float* o = reinterpret_cast<float*>(p.out) + out_row * p.ldo + col;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
        }
      }
      // all TMEM reads of this accumulator are complete (wait::ld above): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

template <int BN, int EPI>
int launch_one(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  using C = Cfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    KB_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::SMEM_BYTES));
    attr_set = true;
  }
  KParams p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.bias = a.bias; p.gamma = a.gamma; p.resid = a.resid; p.ldr = a.ldr;
  p.out = a.out; p.ldo = a.ldo; p.pos = a.pos; p.patches = a.patches;
  p.idesc = make_idesc(a.bf16 ? kFmtBF16 : kFmtF16, BLOCK_M, BN);
  p.bf16 = a.bf16;
  const int m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const int n_tiles = (a.N + BN - 1) / BN;
  const int tiles = m_tiles * n_tiles;
  int grid = num_sms();
  if (tiles < grid) grid = tiles;
  profile_gemm_begin(stream);
  gemm_kernel<BN, EPI><<<grid, kThreads, C::SMEM_BYTES, stream>>>(ta, tb, p);
  profile_gemm_end(stream, 2.0 * a.M * a.N * a.K);
  note_launch();
  KB_CUDA_CHECK(cudaGetLastError());
  return KB_OK;
}

template <int BN>
int dispatch_epi(const GemmArgs& a, const CUtensorMap& ta, const CUtensorMap& tb, cudaStream_t stream) {
  switch (a.epi) {
    case EPI_BIAS_HALF: return launch_one<BN, EPI_BIAS_HALF>(a, ta, tb, stream);
    case EPI_BIAS_GELU_HALF: return launch_one<BN, EPI_BIAS_GELU_HALF>(a, ta, tb, stream);
    case EPI_RESID_F32: return launch_one<BN, EPI_RESID_F32>(a, ta, tb, stream);
    case EPI_BIAS_F32: return launch_one<BN, EPI_BIAS_F32>(a, ta, tb, stream);
    case EPI_PATCH_F32: return launch_one<BN, EPI_PATCH_F32>(a, ta, tb, stream);
    default: return set_error(KB_ERR_ARG, "gemm: unknown epilogue %d", a.epi);
  }
}

}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t stream) {
  if (a.M <= 0 || a.N <= 0 || a.K <= 0) return set_error(KB_ERR_ARG, "gemm: empty problem %dx%dx%d", a.M, a.N, a.K);
  if (a.K % BLOCK_K != 0) return set_error(KB_ERR_ARG, "gemm: K=%d must be a multiple of %d", a.K, BLOCK_K);
  if (a.N % 32 != 0) return set_error(KB_ERR_ARG, "gemm: N=%d must be a multiple of 32", a.N);
  if (a.epi == EPI_RESID_F32 && a.resid == nullptr) return set_error(KB_ERR_ARG, "gemm: residual epilogue without resid");
  if (a.epi == EPI_PATCH_F32 && (a.pos == nullptr || a.patches <= 0))
    return set_error(KB_ERR_ARG, "gemm: patch epilogue without pos/patches");
  // Tile-width choice: 256-wide tiles halve the A re-reads; 128-wide tiles give twice as many work units
  // when the problem is too small to fill the machine.
  const int m_tiles = (a.M + BLOCK_M - 1) / BLOCK_M;
  const bool wide = (a.N % 256 == 0) && ((long long)m_tiles * (a.N / 256) >= num_sms());
  const int dt = a.bf16 ? KB_BF16 : KB_F16;
  CUtensorMap ta, tb;
  int rc = get_tmap_2d(a.A, dt, a.M, a.K, a.lda, BLOCK_M, &ta);
  if (rc) return rc;
  rc = get_tmap_2d(a.W, dt, a.N, a.K, a.ldw, wide ? 256 : 128, &tb);
  if (rc) return rc;
  return wide ? dispatch_epi<256>(a, ta, tb, stream) : dispatch_epi<128>(a, ta, tb, stream);
}

}  // namespace kb
