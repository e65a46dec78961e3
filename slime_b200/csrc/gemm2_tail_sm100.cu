// OPT-IN, NOT YET VALIDATED ON HARDWARE (written after round 1's GPU budget was spent; SLIME_GEMM_TAIL_SPLIT=1 enables it,
// tests/test_gemm_tail_split_gpu.py is the validation to run first).  Nothing on the default path reaches this file.
//
// Tail-split ("stream-K tail") variant of the 2-CTA tcgen05 GEMM (gemm2_sm100.cu).  The persistent kernel there gives
// every cluster ceil(tiles / clusters) whole tiles; when the tile count is a little more than a multiple of the 74
// clusters the last wave runs nearly empty - at batch 1 the o- and down-projections of the decoder have 96 tiles,
// i.e. two waves 65 % full (profiles/r01_bench_baseline_configs.txt: 0.63 of peak at batch 1 vs 0.84 batched).
// Here the R = tiles mod clusters tail tiles are cut along K into S = clusters / R parts each, so the last wave is one
// round of R * S <= clusters partial items of K / S instead of R whole tiles:
//   * parts 1 .. S-1 ("writers") dump their raw fp32 accumulator tile into a workspace and signal a counter;
//   * part 0 (the "finisher") waits for the counter, adds the partial tiles IN PART ORDER (deterministic) to its own
//     accumulator registers and runs the ordinary epilogue (bias / activation / SwiGLU / RoPE / residual / scatter).
// All clusters are co-resident (persistent grid <= one cluster per SM pair) and every cluster handles its tail item
// last, so the finisher's wait cannot deadlock.  The counters clean themselves (the last finisher warp resets them).
// Cost: (S - 1) * 256 KB of L2 traffic per tail tile.  Numerics: a tail tile's K sum is formed from S partial sums, so
// its rows can differ in the last bit from the same rows computed in a whole tile - the bit-exact batch-invariance of the
// default schedule (tests/test_fullsize_gpu.py) does not hold with the split on, which is why it is a switch.
#include <cstdlib>

#include "errors.h"
#include "gemm.h"

namespace {

#include "gemm2_common.cuh"

constexpr int TAIL_TILE_FLOATS = 2 * BLOCK_M * BLOCK_N;  // one 256 x 256 fp32 accumulator tile
constexpr int TAIL_ARRIVALS = 2 * NUM_EPI_WARPS;         // epilogue warps of both CTAs signal per part

struct WorkItem {
  int tile;       // linear tile index (rasterised by tile_coord)
  int kb0, kb1;   // k-block range
  int part, parts;
  int tail;       // index of the tail tile (workspace / counter slot), -1 for whole tiles
};

// Item i of cluster c: whole tiles c, c + C, ... for i < waves (= full / C), then - as item `waves` - at most one
// tail item.  Tail item of cluster c: tail tile r = c % R, part s = c / R, i.e. the parts of one tile sit on clusters
// r, r + R, r + 2R, ... and R * S <= C clusters get one.
SLIME_DEVINL bool get_item(int i, int cluster, int clusters, int num_kb, int full, int R, int S, WorkItem& w) {
  const int waves = full / clusters;
  if (i < waves) {
    w.tile = cluster + i * clusters;
    w.kb0 = 0; w.kb1 = num_kb; w.part = 0; w.parts = 1; w.tail = -1;
    return true;
  }
  if (i > waves || cluster >= R * S) return false;
  const int r = cluster % R, s = cluster / R;
  w.tile = full + r;
  w.kb0 = static_cast<int>(static_cast<long long>(s) * num_kb / S);
  w.kb1 = static_cast<int>(static_cast<long long>(s + 1) * num_kb / S);
  w.part = s; w.parts = S; w.tail = r;
  return true;
}

SLIME_DEVINL int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Writer: this warp's share of the raw accumulator tile -> workspace (row-major 256 x 256 fp32), then one arrival.
SLIME_DEVINL void tail_store_partial(uint32_t tmem_acc, float* __restrict__ ws_tile, int rank, int quad, int half, int lane,
                                     int* counter) {
  const int row = rank * BLOCK_M + quad * 32 + lane;
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + half * (BLOCK_N / 2);
  float* dst = ws_tile + static_cast<size_t>(row) * BLOCK_N + half * (BLOCK_N / 2);
#pragma unroll 1
  for (int i = 0; i < BLOCK_N / 64; ++i) {
    uint32_t acc[32];
    tmem_ld_32x32b_x32(taddr + i * 32, acc);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 8; ++g)
      __stcg(reinterpret_cast<float4*>(dst + i * 32 + g * 4),
             make_float4(__uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1]), __uint_as_float(acc[4 * g + 2]),
                         __uint_as_float(acc[4 * g + 3])));
  }
  __threadfence();
  __syncwarp();
  if (lane == 0) atomicAdd(counter, 1);
}

// Finisher: ordinary (direct-store) epilogue with the partial tiles of parts 1 .. parts-1 added to the accumulator.
template <int EPI>
SLIME_DEVINL void tail_finish_tile(const GemmParams& p, uint32_t tmem_acc, int m0, int n0, int rank, int quad, int half, int lane,
                                   const float* __restrict__ ws_tiles, int parts, int* counter, int* done) {
  if (lane == 0) {
    const int need = (parts - 1) * TAIL_ARRIVALS;
    while (ld_acquire_gpu(counter) < need) __nanosleep(64);
  }
  __syncwarp();
  EpiRow er;
  const int row = m0 + quad * 32 + lane;
  const bool row_ok = row < p.M;
  er.out_row = row;
  if (row_ok && p.row_map != nullptr) er.out_row = p.row_map[row];
  er.store_ok = row_ok && er.out_row >= 0;
  er.res_row = (p.res_period > 0) ? row % p.res_period : row;
  er.pos = 0;
  if constexpr (EPI == GEMM_EPI_ROPE) {
    if (row_ok) er.pos = min(max(p.rope_pos[row], 0), p.rope_max_pos - 1);
  }
  const EpiCoal ec = {};
  const int trow = rank * BLOCK_M + quad * 32 + lane;  // row inside the 256-row tile
  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + half * (BLOCK_N / 2);
  const int col_begin = n0 + half * (BLOCK_N / 2);
#pragma unroll 1
  for (int i = 0; i < BLOCK_N / 64; ++i) {
    uint32_t acc[32];
    uint4 res[4], bia[4];
    tmem_ld_32x32b_x32(taddr + i * 32, acc);
    if constexpr (EPI != GEMM_EPI_SWIGLU) epi_issue_bias(p, col_begin + i * 32, bia);
    epi_issue_residual<EPI, false>(p, er, ec, lane, col_begin + i * 32, res);
    tmem_ld_wait();
    for (int s = 1; s < parts; ++s) {  // fixed order: deterministic
      const float* src = ws_tiles + static_cast<size_t>(s - 1) * TAIL_TILE_FLOATS + static_cast<size_t>(trow) * BLOCK_N +
                         half * (BLOCK_N / 2) + i * 32;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + g * 4));
        acc[4 * g] = __float_as_uint(__uint_as_float(acc[4 * g]) + v.x);
        acc[4 * g + 1] = __float_as_uint(__uint_as_float(acc[4 * g + 1]) + v.y);
        acc[4 * g + 2] = __float_as_uint(__uint_as_float(acc[4 * g + 2]) + v.z);
        acc[4 * g + 3] = __float_as_uint(__uint_as_float(acc[4 * g + 3]) + v.w);
      }
    }
    epi_process_chunk_direct<EPI>(p, er, col_begin + i * 32, acc, res, bia);
  }
  __syncwarp();
  if (lane == 0) {  // the last of the 16 finisher warps leaves the slot clean for the next launch
    if (atomicAdd(done, 1) == TAIL_ARRIVALS - 1) {
      *counter = 0;
      *done = 0;
      __threadfence();
    }
  }
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_tn_2cta_tail_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                              const GemmParams p, const int group_m, const int full, const int R, const int S,
                              float* __restrict__ ws, int* __restrict__ counters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 2);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full_bar[s], 1);
      mbar_init(&tmem_empty_bar[s], 2 * 32 * NUM_EPI_WARPS);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc_2sm(tmem_holder, TMEM_COLS);
  }
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx == 0) {
    // ============================ TMA producer (both CTAs) ============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      WorkItem w;
      for (int i = 0; get_item(i, cluster_id, num_clusters, num_kb, full, R, S, w); ++i) {
        const TileCoord tc = tile_coord(w.tile, num_m, num_n, group_m);
        const int m_row = tc.m_blk * 2 * BLOCK_M + rank * BLOCK_M;
        const int n_row = tc.n_blk * BLOCK_N + rank * (BLOCK_N / 2);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) {
            mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
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
    if (leader && lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      WorkItem w;
      for (int it = 0; get_item(it, cluster_id, num_clusters, num_kb, full, R, S, w); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tcgen05_fence_after();
          const uint64_t desc_a = make_umma_desc_sw128(smem_u32(smem_a + stage * A_BYTES));
          const uint64_t desc_b = make_umma_desc_sw128(smem_u32(smem_b + stage * B_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            umma_bf16_ss_2sm(tmem_d, desc_a + 2 * k, desc_b + 2 * k, idesc, (kb != w.kb0 || k != 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tmem_full_bar[acc]);
      }
    }
  } else {
    // ============================ epilogue (both CTAs, own 128 rows) ==================
    const int quad = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;
    WorkItem w;
    for (int it = 0; get_item(it, cluster_id, num_clusters, num_kb, full, R, S, w); ++it) {
      const TileCoord tc = tile_coord(w.tile, num_m, num_n, group_m);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * BLOCK_N;
      const int m0 = tc.m_blk * 2 * BLOCK_M + rank * BLOCK_M, n0 = tc.n_blk * BLOCK_N;
      if (w.tail < 0) {
        epilogue_tile<BLOCK_N, EPI, false>(p, tmem_acc, m0, n0, quad, half, lane, nullptr);
      } else {
        float* tiles = ws + static_cast<size_t>(w.tail) * (w.parts - 1) * TAIL_TILE_FLOATS;
        if (w.part > 0) {
          tail_store_partial(tmem_acc, tiles + static_cast<size_t>(w.part - 1) * TAIL_TILE_FLOATS, rank, quad, half, lane,
                             counters + 2 * w.tail);
        } else {
          tail_finish_tile<EPI>(p, tmem_acc, m0, n0, rank, quad, half, lane, tiles, w.parts, counters + 2 * w.tail,
                                counters + 2 * w.tail + 1);
        }
      }
      tcgen05_fence_before();
      mbar_arrive_leader(&tmem_empty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  cluster_sync_all();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  }
}

int g_tail_mode = -1;  // -1 unset, 0 off (default), 1 on

template <int EPI>
int launch_tail(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int clusters, int full, int R, int S,
                cudaStream_t stream) {
  auto kern = gemm_bf16_tn_2cta_tail_kernel<EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  slime_prof_begin(0, 2.0 * p.M * static_cast<double>(p.N) * p.K, stream);
  kern<<<2 * clusters, NUM_THREADS, SMEM_BYTES, stream>>>(ta, tb, p, slime_gemm_group_m(p.K, 2 * BLOCK_M), full, R, S,
                                                          p.splitk_ws, p.tail_counters);
  slime_prof_end(stream);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

}  // namespace

extern "C" int slime_gemm_set_tail_split(int mode) {
  g_tail_mode = mode < 0 ? -1 : (mode != 0 ? 1 : 0);
  return SLIME_OK;
}

bool slime_gemm_tail_split_enabled() {
  if (g_tail_mode < 0) {
    const char* e = getenv("SLIME_GEMM_TAIL_SPLIT");
    g_tail_mode = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return g_tail_mode == 1;
}

size_t slime_gemm_tail_ws_floats(int num_sms) { return static_cast<size_t>(num_sms / 2) * TAIL_TILE_FLOATS; }
int slime_gemm_tail_counter_ints(int num_sms) { return 2 * (num_sms / 2); }

// Returns 1 when the problem was launched here, 0 when the tail split does not apply (the caller runs the ordinary kernel).
int slime_launch_gemm_2cta_tail(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi, int num_sms,
                                cudaStream_t stream, int* launched) {
  *launched = 0;
  if (!slime_gemm_tail_split_enabled() || p.splitk_ws == nullptr || p.tail_counters == nullptr || p.epi_mode != 0 ||
      p.out_f32 != nullptr)
    return SLIME_OK;
  const int num_m = (p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles = num_m * num_n;
  const int num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int clusters = num_sms / 2;
  if (tiles <= clusters) return SLIME_OK;  // a single (partial) wave: nothing to rebalance against
  const int R = tiles % clusters, full = tiles - R;
  if (R == 0) return SLIME_OK;
  int S = clusters / R;
  if (S > num_kb / 4) S = num_kb / 4;  // at least 4 k-blocks per part
  if (S < 2) return SLIME_OK;
  if (static_cast<size_t>(R) * (S - 1) * TAIL_TILE_FLOATS > p.splitk_ws_floats) return SLIME_OK;
  CUtensorMap ta, tb;
  SLIME_PROPAGATE(slime_get_tmap(A, p.M, p.K, lda, BLOCK_M, &ta));
  SLIME_PROPAGATE(slime_get_tmap(W, p.N, p.K, ldw, BLOCK_N / 2, &tb));
  int rc = SLIME_OK;
  switch (epi) {
    case GEMM_EPI_NONE: rc = launch_tail<GEMM_EPI_NONE>(ta, tb, p, clusters, full, R, S, stream); break;
    case GEMM_EPI_QUICK_GELU: rc = launch_tail<GEMM_EPI_QUICK_GELU>(ta, tb, p, clusters, full, R, S, stream); break;
    case GEMM_EPI_GELU_ERF: rc = launch_tail<GEMM_EPI_GELU_ERF>(ta, tb, p, clusters, full, R, S, stream); break;
    case GEMM_EPI_SWIGLU: rc = launch_tail<GEMM_EPI_SWIGLU>(ta, tb, p, clusters, full, R, S, stream); break;
    case GEMM_EPI_ROPE: rc = launch_tail<GEMM_EPI_ROPE>(ta, tb, p, clusters, full, R, S, stream); break;
    default: return SLIME_OK;
  }
  SLIME_PROPAGATE(rc);
  *launched = 1;
  return SLIME_OK;
}
