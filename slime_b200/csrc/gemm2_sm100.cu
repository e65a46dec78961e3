// 2-CTA (cta_group::2) variant of the persistent tcgen05 GEMM:  C[M,N] = A[M,K] * W[N,K]^T (+ epilogue).
//
// A cluster of two CTAs on one TPC executes UMMA 256 x 256 x 16: CTA r owns rows [r*128, r*128+128) of the
// 256-row tile (its own A tile and its own 128 TMEM lanes x 256 accumulator columns) and stages only HALF of
// the W tile (128 of the 256 output columns); the tensor cores of the pair read both halves.  Per SM and
// k-block that is 16 KB (A) + 16 KB (half of W) of shared-memory traffic instead of 16 + 32 KB, and 6 pipeline
// stages fit where the 1-CTA kernel has 4 - less shared-memory power per FLOP on a power-capped part.
//
//   warp 0 (both CTAs): TMA producer; loads signal the LEADER's (rank 0) full barrier (.cta_group::2 TMA)
//   warp 1 (leader)   : issues tcgen05.mma.cta_group::2; commits multicast to both CTAs' barriers
//   warps 2..9 (both) : epilogue of the CTA's own 128 rows (gemm_epilogue.cuh), arriving on the leader's
//                       tmem_empty barrier
#include <cstdlib>

#include "errors.h"
#include "gemm.h"

namespace {

#include "gemm2_common.cuh"

template <int BLOCK_N, int EPI, bool STAGED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmParams p, const int group_m) {
  using Cfg = Gemm2Cfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES, TMEM_COLS = Cfg::TMEM_COLS,
                EPI_OFF = Cfg::EPI_OFF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  uint8_t* epi_stage = smem + EPI_OFF;  // [NUM_EPI_WARPS][EPI_STAGE_BYTES]

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);  // 256-row tiles
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = num_m * num_n;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);   // leader's arrive.expect_tx + the peer producer's remote arrive
      mbar_init(&empty_bar[s], 1);  // one multicast commit from the leader's MMA thread
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * 32 * NUM_EPI_WARPS);  // epilogue threads of BOTH CTAs (leader's copy is used)
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc_2sm(tmem_holder, TMEM_COLS);
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_trigger();
  pdl_wait();  // everything above overlaps the previous kernel's tail when launched with the programmatic attribute

  if (warp_idx == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        const TileCoord tc = tile_coord(t, num_m, num_n, group_m);
        const int m_row = tc.m_blk * 2 * BLOCK_M + rank * BLOCK_M;          // this CTA's 128 rows of A
        const int n_row = tc.n_blk * BLOCK_N + rank * (BLOCK_N / 2);        // this CTA's half of the W tile
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);  // bytes of both CTAs land on this barrier
          } else {
            mbar_arrive_leader(&full_bar[stage]);
          }
          tma_load_2d_2sm(smem_a + stage * A_BYTES, &tmap_a, &full_bar[stage], kb * BLOCK_K, m_row);
          tma_load_2d_2sm(smem_b + stage * B_BYTES, &tmap_b, &full_bar[stage], kb * BLOCK_K, n_row);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ============================ MMA issuer (leader CTA only) ========================
    // Whole warp converged, one elected lane issues (descriptors stay in uniform registers: see gemm_sm100.cu).
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N);
      const uint32_t smem_a_u = static_cast<uint32_t>(__shfl_sync(0xffffffffu, static_cast<int>(smem_u32(smem_a)), 0));
      const uint32_t smem_b_u = smem_a_u + STAGES * A_BYTES;
      const uint32_t tmem_u = static_cast<uint32_t>(__shfl_sync(0xffffffffu, static_cast<int>(tmem_base), 0));
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_u + acc * BLOCK_N;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t desc_a = make_umma_desc_sw128(smem_a_u + stage * A_BYTES);
          const uint64_t desc_b = make_umma_desc_sw128(smem_b_u + stage * B_BYTES);
          if (elect_one_sync()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              umma_bf16_ss_2sm(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == num_kb - 1) umma_commit_2sm(&tmem_full_bar[acc]);
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
    // ============================ epilogue (both CTAs, own 128 rows) ==================
    const int quad = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;
    int it = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
      const TileCoord tc = tile_coord(t, num_m, num_n, group_m);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      epilogue_tile<BLOCK_N, EPI, STAGED>(p, tmem_base + acc * BLOCK_N, tc.m_blk * 2 * BLOCK_M + rank * BLOCK_M,
                                  tc.n_blk * BLOCK_N, quad, half, lane,
                                  epi_stage + (warp_idx - 2) * EPI_STAGE_BYTES);
      tcgen05_fence_before();
      mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();  // nobody leaves (or frees TMEM) while the peer may still touch this CTA's smem / TMEM
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

template <int BLOCK_N, int EPI, bool STAGED>
int launch2s(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms, cudaStream_t stream) {
  constexpr int SMEM_BYTES = Gemm2Cfg<BLOCK_N>::SMEM_BYTES;
  auto kern = gemm_bf16_tn_2cta_kernel<BLOCK_N, EPI, STAGED>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = num_m * num_n;
  const int max_clusters = num_sms / 2;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  slime_prof_begin(0, 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
  const cudaError_t le = slime_launch_prefill(kern, dim3(2 * clusters), dim3(NUM_THREADS), SMEM_BYTES, stream, ta, tb, p,
                                              slime_gemm_group_m(p.K, 2 * BLOCK_M));
  slime_prof_end(stream);
  SLIME_CHECK_CUDA(le);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

template <int BLOCK_N, int EPI>
int launch2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms, cudaStream_t stream) {
  if (p.epi_mode != 0 && p.out_f32 == nullptr) return launch2s<BLOCK_N, EPI, true>(ta, tb, p, num_sms, stream);
  return launch2s<BLOCK_N, EPI, false>(ta, tb, p, num_sms, stream);
}

template <int BLOCK_N>
int launch2n(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi, int num_sms, cudaStream_t stream) {
  CUtensorMap ta, tb;
  SLIME_PROPAGATE(slime_get_tmap(A, p.M, p.K, lda, BLOCK_M, &ta));
  SLIME_PROPAGATE(slime_get_tmap(W, p.N, p.K, ldw, BLOCK_N / 2, &tb));
  switch (epi) {
    case GEMM_EPI_NONE:
      return launch2<BLOCK_N, GEMM_EPI_NONE>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_QUICK_GELU:
      return launch2<BLOCK_N, GEMM_EPI_QUICK_GELU>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_GELU_ERF:
      return launch2<BLOCK_N, GEMM_EPI_GELU_ERF>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_SWIGLU:
      return launch2<BLOCK_N, GEMM_EPI_SWIGLU>(ta, tb, p, num_sms, stream);
    case GEMM_EPI_ROPE:
      return launch2<BLOCK_N, GEMM_EPI_ROPE>(ta, tb, p, num_sms, stream);
    default:
      slime_set_error("unknown GEMM epilogue %d", epi);
      return SLIME_EINVAL;
  }
}

}  // namespace

// Tile width: 256 columns unless 192-column tiles finish the problem in fewer (width-weighted) waves - only small problems
// qualify (batch-1 prefill: o- / down-projection), anything of many waves stays on the throughput shape.  SLIME_GEMM2_BN=256 /
// 192 forces one (A/B, tests).
static int g_force_bn = -1;  // -1: read SLIME_GEMM2_BN on first use; 0 = by shape; 256 / 192 = forced
extern "C" int slime_gemm_set_tile_n(int bn) {
  if (bn != -1 && bn != 0 && bn != 192 && bn != 256) {
    slime_set_error("gemm tile width %d not in {-1 (default), 0 (by shape), 192, 256}", bn);
    return SLIME_EINVAL;
  }
  g_force_bn = bn;
  return SLIME_OK;
}
int slime_gemm2_block_n(int M, int N, int num_sms) {
  if (g_force_bn < 0) {
    const char* e = getenv("SLIME_GEMM2_BN");
    g_force_bn = e != nullptr ? atoi(e) : 0;
    if (g_force_bn != 192 && g_force_bn != 256) g_force_bn = 0;
  }
  const int force_bn = g_force_bn;
  if (force_bn == 256 || force_bn == 192) return N >= 192 ? force_bn : 256;
  const int clusters = num_sms / 2;
  const long long m_tiles = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const long long t256 = m_tiles * ((N + 255) / 256), t192 = m_tiles * ((N + 191) / 192);
  const double w256 = static_cast<double>((t256 + clusters - 1) / clusters);              // in units of a 256-wide tile
  const double w192 = static_cast<double>((t192 + clusters - 1) / clusters) * 0.75 * 1.04;  // (slightly less efficient per FLOP)
  return (N >= 192 && w192 < w256) ? 192 : 256;
}

int slime_launch_gemm_2cta(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi, int num_sms,
                           cudaStream_t stream) {
  if (slime_gemm2_block_n(p.M, p.N, num_sms) == 192) return launch2n<192>(A, lda, W, ldw, p, epi, num_sms, stream);
  return launch2n<256>(A, lda, W, ldw, p, epi, num_sms, stream);
}
