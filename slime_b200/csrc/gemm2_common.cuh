// Shared pieces of the 2-CTA (cta_group::2) tcgen05 GEMM kernels: tile constants, the cluster / TMA / UMMA / TMEM
// wrappers and the rasterisation.  Included INSIDE the anonymous namespace of gemm2_sm100.cu; needs gemm.h and
// common.cuh before it.
#pragma once

constexpr int BLOCK_M = 128;      // rows per CTA (256 per cluster tile)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;        // 16 KB
// Cluster tile 256 x BLOCK_N.  256 columns is the throughput shape (UMMA 256 x 256 x 16; 96 B/clk of shared-memory operand
// reads per SM).  192 columns (UMMA 256 x 192 x 16, 108 B/clk - still under the 128 B/clk limit that made a 128-wide tile
// slow) exists for problems whose 256-wide tiles leave the last wave nearly empty: a batch-1 prefill has M = 1379 rows, and
// N = 4096 is 96 tiles for 74 clusters - two waves, the second 30 % full; as 132 tiles of 192 columns it is two waves of
// three quarters the length.  Same k order per output element: results do not depend on the tile width.
template <int BLOCK_N>
struct Gemm2Cfg {
  static_assert(BLOCK_N == 256 || BLOCK_N == 192, "cluster tile width");
  static constexpr int STAGES = 6;
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // this CTA's half of the W tile (16 / 12 KB)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 512;                // two accumulators of BLOCK_N columns (allocation: a power of two)
  static constexpr int EPI_OFF = STAGES * STAGE_BYTES + 256;  // barriers + TMEM holder live in the 256 bytes before it
  static constexpr int SMEM_BYTES = 1024 + EPI_OFF + NUM_EPI_WARPS * 4096;  // + the epilogue staging tiles
  static_assert((2 * STAGES + 4) * 8 + 16 <= 256, "barrier block");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> rank 0

#include "gemm_epilogue.cuh"

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

SLIME_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
SLIME_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in the LEADER CTA (works from either CTA)
SLIME_DEVINL void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(smem_u32(bar) & PEER_MASK) : "memory");
}
SLIME_DEVINL void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
SLIME_DEVINL void umma_bf16_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                   uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit -> arrive on the barrier at this offset in BOTH CTAs of the pair
SLIME_DEVINL void umma_commit_2sm(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
SLIME_DEVINL void tmem_alloc_2sm(uint32_t* smem_holder, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_holder)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
SLIME_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(ncols) : "memory");
}

