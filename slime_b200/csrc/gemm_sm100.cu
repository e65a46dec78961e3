// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * both operands K-major (activations row-major, nn.Linear weights [out,in] row-major), moved
//     HBM -> shared memory by TMA (cp.async.bulk.tensor, 128-byte swizzle),
//   * tcgen05.mma (UMMA 128 x BLOCK_N x 16, bf16 in / fp32 accumulate) issued by ONE thread,
//   * accumulators live in TMEM, double-buffered (2 x BLOCK_N columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1,
//   * epilogue warps read TMEM with tcgen05.ld and apply bias / activation / residual / SwiGLU /
//     row scatter before writing bf16 (or fp32) straight to HBM with 16-byte vector stores.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (warp w owns TMEM lane quadrant w % 4; two warps per quadrant split the columns;
// see gemm_epilogue.cuh).
//
// This one kernel carries every dense contraction of the SliME prefill path: CLIP patch-embed /
// QKV / out-proj / MLP (HF clip/modeling_clip.py:209,310-312,334,348-350), the Resampler K/V and
// out projections (reference llava/model/multimodal_resampler/sampler.py:128), the mm_projector
// MLP (reference llava/model/multimodal_projector/builder.py:53-57), and the Llama QKV / o-proj /
// gate-up / down / lm_head projections (HF llama/modeling_llama.py:183,262-288,487).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "attention.h"
#include "elementwise.h"
#include "errors.h"
#include "gemm.h"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle-128B row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;  // TMA warp + MMA warp + 8 epilogue warps

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int STAGES = BLOCK_N == 256 ? 4 : 6;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES + ((BAR_BYTES + 127) / 128) * 128;  // epilogue staging tiles
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + EPI_OFF + NUM_EPI_WARPS * 4096;
};

struct TileCoord {
  int m_blk, n_blk;
};

SLIME_DEVINL TileCoord tile_coord(int t, int num_m, int num_n, int group_m_in) {
  // group_m_in < 0: serpentine - odd groups sweep the n-tiles downwards, so the W tiles still in L2 at a group boundary
  // are used again
  const bool snake = group_m_in < 0;
  const int group_m = snake ? -group_m_in : group_m_in;
  const int per_group = group_m * num_n;
  const int group = t / per_group;
  const int first_m = group * group_m;
  const int gsize = min(group_m, num_m - first_m);
  const int in = t - group * per_group;
  TileCoord c;
  c.m_blk = first_m + in % gsize;
  c.n_blk = in / gsize;
  if (snake && (group & 1)) c.n_blk = num_n - 1 - c.n_blk;
  return c;
}

#include "gemm_epilogue.cuh"

template <int BLOCK_N, int EPI, bool STAGED>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, const GemmParams p, const int group_m) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_stage = smem + Cfg::EPI_OFF;  // [NUM_EPI_WARPS][EPI_STAGE_BYTES]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 32 * NUM_EPI_WARPS);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_trigger();
  pdl_wait();  // the prologue above overlaps the previous kernel's tail when launched with the programmatic attribute

  if (warp_idx == 0) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const TileCoord tc = tile_coord(t, num_m, num_n, group_m);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(smem_a + stage * Cfg::A_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K,
                      tc.m_blk * BLOCK_M);
          tma_load_2d(smem_b + stage * Cfg::B_BYTES, &tmap_b, &full_bar[stage], kb * BLOCK_K,
                      tc.n_blk * BLOCK_N);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer ==============================
    // The whole warp runs the loop converged and ONE elected lane issues: every descriptor input is then known to be
    // warp-uniform and lives in uniform registers, so the four MMAs of a k-block are issued back to back (from inside
    // `if (lane == 0)` each tcgen05.mma costs ~14 instructions of R2UR / ELECT / branch, ~95 cycles - more than a
    // 128 x 128 x 16 MMA takes to execute).
    {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
      const uint32_t smem_a_u = static_cast<uint32_t>(__shfl_sync(0xffffffffu, static_cast<int>(smem_u32(smem_a)), 0));
      const uint32_t smem_b_u = smem_a_u + STAGES * Cfg::A_BYTES;
      const uint32_t tmem_u = static_cast<uint32_t>(__shfl_sync(0xffffffffu, static_cast<int>(tmem_base), 0));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_u + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t desc_a = make_umma_desc_sw128(smem_a_u + stage * Cfg::A_BYTES);
          const uint64_t desc_b = make_umma_desc_sw128(smem_b_u + stage * Cfg::B_BYTES);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in (addr >> 4) units
              umma_bf16_ss(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
            if (kb == num_kb - 1) umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else {
    // ============================ epilogue (8 warps) ======================
    const int quad = warp_idx & 3;          // TMEM lane quadrant this warp may access
    const int half = (warp_idx - 2) >> 2;   // warps 2..5 take the low half of the columns, 6..9 the high half
    int it = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
      const TileCoord tc = tile_coord(t, num_m, num_n, group_m);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      epilogue_tile<BLOCK_N, EPI, STAGED>(p, tmem_base + acc * BLOCK_N, tc.m_blk * BLOCK_M, tc.n_blk * BLOCK_N, quad, half,
                                  lane, epi_stage + (warp_idx - 2) * EPI_STAGE_BYTES);
      tcgen05_fence_before();
      mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched at run time, so the library
// itself has no link-time dependency on libcuda and still loads on a GPU-less build box)
// ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(f);
    }
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int rows, cols, ld, box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && ld == o.ld && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (static_cast<size_t>(k.rows) * 0x9E3779B97F4A7C15ull) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.cols) * 0xC2B2AE3D27D4EB4Full) + (h << 6) + (h >> 2);
    h ^= (static_cast<size_t>(k.ld) * 0x165667B19E3779F9ull) + (h << 6) + (h >> 2);
    h ^= static_cast<size_t>(k.box_rows) + (h << 6) + (h >> 2);
    return h;
  }
};

std::mutex g_tmap_mu;
std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;

int get_tmap(const bf16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  TmapKey key{ptr, rows, cols, ld, box_rows};
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *out = it->second;
      return SLIME_OK;
    }
  }
  PFN_encodeTiled enc = get_encode_fn();
  if (enc == nullptr) {
    slime_set_error("cuTensorMapEncodeTiled driver entry point unavailable (no CUDA driver?)");
    return SLIME_ECUDA;
  }
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * sizeof(bf16)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = enc(&m, SLIME_TMAP_ELEM, 2, const_cast<bf16*>(ptr), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    slime_set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%d cols=%d ld=%d box_rows=%d",
                    static_cast<int>(r), ptr, rows, cols, ld, box_rows);
    return SLIME_ECUDA;
  }
  {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (g_tmap_cache.size() > 8192) g_tmap_cache.clear();
    g_tmap_cache[key] = m;
  }
  *out = m;
  return SLIME_OK;
}

template <int BLOCK_N, int EPI, bool STAGED>
int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
               cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, EPI, STAGED>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = num_m * num_n;
  const int grid = tiles < num_sms ? tiles : num_sms;
  slime_prof_begin(0, 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
  const cudaError_t le = slime_launch_prefill(kern, dim3(grid), dim3(NUM_THREADS), Cfg::SMEM_BYTES, stream, ta, tb, p,
                                              slime_gemm_group_m(p.K, BLOCK_M));
  slime_prof_end(stream);
  SLIME_CHECK_CUDA(le);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

template <int BLOCK_N, bool STAGED>
int launch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int epi,
               int num_sms, cudaStream_t stream) {
  switch (epi) {
    case GEMM_EPI_NONE:
      return launch_cfg<BLOCK_N, GEMM_EPI_NONE, STAGED>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_QUICK_GELU:
      return launch_cfg<BLOCK_N, GEMM_EPI_QUICK_GELU, STAGED>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_GELU_ERF:
      return launch_cfg<BLOCK_N, GEMM_EPI_GELU_ERF, STAGED>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_SWIGLU:
      return launch_cfg<BLOCK_N, GEMM_EPI_SWIGLU, STAGED>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_ROPE:
      return launch_cfg<BLOCK_N, GEMM_EPI_ROPE, STAGED>(ta, tb, p, num_sms, stream);
    default:
      slime_set_error("unknown GEMM epilogue %d", epi);
      return SLIME_EINVAL;
  }
}

}  // namespace

// Rasterisation group: consecutive tiles sweep `group` m-tiles before moving to the next n-tile, so one group's A
// panel (group * tile_rows x K) stays L2-resident while the whole W matrix streams past it once.  The panel is
// sized to ~32 MB of the 126 MB L2; a bigger group means fewer DRAM passes over W (SLIME_GEMM_GROUP_ROWS overrides).
int slime_gemm_group_m(int K, int tile_rows) {
  static long long env_rows = -1;
  if (env_rows < 0) {
    const char* e = getenv("SLIME_GEMM_GROUP_ROWS");
    env_rows = e != nullptr ? atoll(e) : 0;
  }
  static int snake = -1;
  if (snake < 0) {
    const char* e = getenv("SLIME_GEMM_SNAKE");
    snake = (e != nullptr && e[0] == '0') ? 0 : 1;  // on by default (profiles/r02_gemm_experiments.txt, item 6)
  }
  static int g_min = -1;
  if (g_min < 0) {
    const char* e = getenv("SLIME_GEMM_GROUP_MIN");
    g_min = e != nullptr ? atoi(e) : 8;
    if (g_min < 1 || g_min > 64) g_min = 8;
  }
  long long rows = env_rows > 0 ? env_rows : (32ll << 20) / (2ll * K);
  long long g = rows / tile_rows;
  // long reductions (K = 14336: 7 MB per 256-row tile) fit no panel; what is left is the sharing between the ~74 clusters of
  // one wave, best for a near-square patch of tiles: at least 8 m-tiles per group (down-projection: 6.4 -> 5.3 GB of DRAM
  // reads per launch, +1.9 %; profiles/r02_gemm_experiments.txt item 10)
  if (env_rows <= 0 && g < g_min) g = g_min;
  if (g < 1) g = 1;
  if (g > 64) g = 64;
  return snake ? -static_cast<int>(g) : static_cast<int>(g);
}

static int g_mode_2cta = -1;  // -1: read SLIME_GEMM_2CTA / the compile-time default on first use
static int g_epi_mode = -1;   // -1: read SLIME_GEMM_EPI_MODE / the compile-time default on first use

extern "C" int slime_gemm_set_epi_mode(int mode) {
  if (mode < -1 || mode > 2) {
    slime_set_error("gemm epilogue mode %d not in {-1 (default), 0 direct, 1 staged, 2 by shape}", mode);
    return SLIME_EINVAL;
  }
  g_epi_mode = mode;
  return SLIME_OK;
}

extern "C" int slime_gemm_set_2cta_mode(int mode) {
  g_mode_2cta = mode;
  return SLIME_OK;
}

int slime_get_tmap(const bf16* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  return get_tmap(ptr, rows, cols, ld, box_rows, out);
}

int slime_launch_gemm(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p_in, int epi,
                      int num_sms, cudaStream_t stream) {
  if (g_epi_mode < 0) {
    const char* e = getenv("SLIME_GEMM_EPI_MODE");
    g_epi_mode = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : SLIME_GEMM_EPI_MODE_DEFAULT;
  }
  GemmParams p = p_in;
  // epilogue access pattern: 0 direct, 1 staged through shared memory (coalesced), 2 = by shape - staged where the A/B on
  // B200 shows it ahead: the ViT-sized problems (tens of thousands of rows with K <= 1024 or N <= 1024, whose epilogue
  // traffic is large next to their few k-blocks: qkv +12 %, fc2 +5 %, fc1 +2 %), direct for the decoder's K >= 4096
  // GEMMs (down-projection -17 % when staged) and for small M (profiles/r02_gemm_epilogue_ab.txt)
  int mode = g_epi_mode;
  if (mode == 2) mode = (p_in.M >= 16384 && (p_in.K <= 1024 || p_in.N <= 1024)) ? 1 : 0;
  p.epi_mode = p.out_f32 != nullptr ? 0 : mode;  // fp32 outputs always go direct
  {
    static int store_hint = -1;
    if (store_hint < 0) {
      const char* e = getenv("SLIME_GEMM_STORE_HINT");
      store_hint = (e != nullptr && e[0] == '0') ? 0 : 1;  // on by default (profiles/r02_gemm_experiments.txt, item 6)
    }
    p.store_hint = (p.out_f32 == nullptr && static_cast<double>(p.M) * p.N * 2 > 64e6) ? store_hint : 0;
  }
  SLIME_REQUIRE(A != nullptr && W != nullptr, "gemm: null operand");
  SLIME_REQUIRE(p.out != nullptr || p.out_f32 != nullptr, "gemm: no output pointer");
  SLIME_REQUIRE(p.M > 0 && p.N > 0 && p.K > 0, "gemm: empty problem M=%d N=%d K=%d", p.M, p.N, p.K);
  SLIME_REQUIRE(p.K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0,
                "gemm: K/lda/ldw must be multiples of 8 (TMA 16-byte strides): K=%d lda=%d ldw=%d",
                p.K, lda, ldw);
  SLIME_REQUIRE(p.N % (epi == GEMM_EPI_SWIGLU ? 16 : 8) == 0, "gemm: N=%d not a multiple of %d", p.N,
                epi == GEMM_EPI_SWIGLU ? 16 : 8);
  SLIME_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0,
                "gemm: operands must be 16-byte aligned");
  const int out_align = p.out_f32 != nullptr ? 4 : 8;
  SLIME_REQUIRE(p.out_ld % out_align == 0, "gemm: out_ld=%d must be a multiple of %d", p.out_ld,
                out_align);
  if (p.residual != nullptr)
    SLIME_REQUIRE(p.res_ld % 8 == 0 && epi != GEMM_EPI_SWIGLU, "gemm: bad residual configuration");
  if (epi == GEMM_EPI_SWIGLU)
    SLIME_REQUIRE(p.bias == nullptr && p.out != nullptr, "gemm: SwiGLU epilogue takes no bias");
  if (epi == GEMM_EPI_ROPE)
    SLIME_REQUIRE(p.rope_pos != nullptr && p.rope_table != nullptr && p.rope_half > 0 && p.rope_half % 16 == 0 &&
                      (p.rope_half & (p.rope_half - 1)) == 0 && p.rope_cols % (2 * p.rope_half) == 0 &&
                      p.rope_max_pos > 0 && p.residual == nullptr,
                  "gemm: bad RoPE epilogue configuration (half=%d cols=%d)", p.rope_half, p.rope_cols);

  if (p.norm_w != nullptr)
    SLIME_REQUIRE(epi == GEMM_EPI_NONE && p.out != nullptr && p.out_f32 == nullptr && p.norm_out != nullptr &&
                      p.row_map == nullptr && p.norm_ld % 8 == 0,
                  "gemm: bad fused-RMSNorm configuration");

  const bool norm_fold = p.row_scale != nullptr || p.sumsq_out != nullptr;
  if (norm_fold) {
    SLIME_REQUIRE(p.out != nullptr && p.out_f32 == nullptr && p.row_map == nullptr,
                  "gemm: norm folding needs a 16-bit output and no row map");
    SLIME_REQUIRE(p.sumsq_out == nullptr || (epi == GEMM_EPI_NONE && p.N % 64 == 0 && p.sumsq_parts == p.N / 64),
                  "gemm: sum-of-squares partials need EPI_NONE and N %% 64 == 0 (N=%d parts=%d)", p.N, p.sumsq_parts);
    p.epi_mode = 0;  // the direct epilogue carries the hooks
  }
  // Decode-step problems (M <= 32): HBM-bound weight streaming instead of 128-row tensor-core tiles.
  if (p.M <= 32 && !norm_fold && slime_gemm_skinny_applies(A, lda, W, ldw, p, epi, num_sms))
    return slime_launch_gemm_skinny(A, lda, W, ldw, p, epi, num_sms, stream);
  SLIME_REQUIRE(p.a_norm_w == nullptr, "gemm: the RMSNorm of the A rows exists on the weight-streaming path only "
                "(M=%d N=%d K=%d does not qualify)", p.M, p.N, p.K);
  if (p.norm_w != nullptr) {  // not fusable here: GEMM, then the RMSNorm of its output rows
    GemmParams q = p_in;
    q.norm_w = nullptr;
    q.norm_out = nullptr;
    SLIME_PROPAGATE(slime_launch_gemm(A, lda, W, ldw, q, epi, num_sms, stream));
    return slime_launch_rmsnorm(p.out, p.out_ld, p.norm_w, p.norm_out, p.norm_ld, p.M, p.N, p.norm_eps, nullptr, stream);
  }
  if (p.kv_k != nullptr) {  // same for the KV-cache append of the decode step's QKV projection
    SLIME_REQUIRE(p.out != nullptr && p.kv_v != nullptr && p.kv_lens != nullptr && p.kv_q_cols + 2 * p.kv_dim == p.N,
                  "gemm: bad KV append configuration");
    GemmParams q = p_in;
    q.kv_k = nullptr;
    SLIME_PROPAGATE(slime_launch_gemm(A, lda, W, ldw, q, epi, num_sms, stream));
    return slime_launch_kv_append(p.out + p.kv_q_cols, p.out + p.kv_q_cols + p.kv_dim, p.out_ld, p.kv_k, p.kv_v, p.kv_dim,
                                  p.kv_lens, p.M, p.kv_cache_len, stream);
  }

  // BLOCK_N = 256 halves the A re-reads; fall back to 128 when the 256-wide grid cannot fill the GPU.
  const int num_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  const int tiles256 = num_m * ((p.N + 255) / 256);
  const bool use256 = p.N >= 256 && tiles256 >= num_sms;
  const int block_n = use256 ? 256 : 128;

  // Large problems go to the 2-CTA kernel (gemm2_sm100.cu).  SLIME_GEMM_2CTA=0 disables it, =1 forces it.
  if (g_mode_2cta < 0) {
    const char* e = getenv("SLIME_GEMM_2CTA");
    g_mode_2cta = (e == nullptr) ? SLIME_GEMM_2CTA_DEFAULT : (e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2));
  }
  const int mode_2cta = g_mode_2cta;
  const int tiles2 = ((p.M + 255) / 256) * ((p.N + 255) / 256);
  if (mode_2cta == 1 || (mode_2cta == 2 && tiles2 >= num_sms / 2 && p.N >= 256))
    return slime_launch_gemm_2cta(A, lda, W, ldw, p, epi, num_sms, stream);

  CUtensorMap ta, tb;
  SLIME_PROPAGATE(get_tmap(A, p.M, p.K, lda, BLOCK_M, &ta));
  SLIME_PROPAGATE(get_tmap(W, p.N, p.K, ldw, block_n, &tb));
  if (p.epi_mode != 0) {
    if (use256) return launch_epi<256, true>(ta, tb, p, epi, num_sms, stream);
    return launch_epi<128, true>(ta, tb, p, epi, num_sms, stream);
  }
  if (use256) return launch_epi<256, false>(ta, tb, p, epi, num_sms, stream);
  return launch_epi<128, false>(ta, tb, p, epi, num_sms, stream);
}
