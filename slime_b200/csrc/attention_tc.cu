// tcgen05 / TMEM flash attention forward for sm_100a (bf16 in, fp32 softmax + accumulation, bf16 out).
//
// One persistent CTA per SM walks a static list of work items (batch, head, 128-row query tile).
//   warp 0     : TMA producer - decodes the items (handed to the other warps through a small ring in shared
//                memory), Q tile per item (double-buffered, posted one item ahead), K and V tiles through 2/2-stage
//                (hd 128) or 4/3-stage (hd 64) rings (64-column slabs of 128 rows, 128-byte swizzle; heads are
//                column slices of packed rows)
//   warp 1     : tcgen05.mma issuer (one thread) + TMEM allocator
//                  S_g = Q K_g^T       SS-MMA 128 x 128 x HD  -> TMEM S buffer (double buffered per tile)
//                  O  += P_g V_g       TS-MMA 128 x HD x 128  -> TMEM O (double buffered per item), A = P_g read
//                                      from TMEM, B = V_g in shared memory as an MN-major operand
//                one continuous stream of tiles across items: S_{g+1} is issued before PV_g even when tile g+1
//                opens the next item, so the tensor pipe never waits for an epilogue
//   warps 2..9 : softmax + epilogue.  Two threads per query row (TMEM lane r is shared by warps w and w+4), each
//                covering 64 of the 128 score columns; half-row max / sum are exchanged through shared memory.
//                exp2 with the softmax scale folded in; P_g (bf16, two per 32-bit column) overwrites the first 64
//                columns of the S buffer it came from.  O is only rescaled when the running max grows by more than
//                2^8 (lazy rescale).  The epilogue of item i (O / l -> HBM, transposed through shared memory so the
//                stores are coalesced) runs after the first score tile of item i+1 has been handed over.
//
// Shapes on the SliME path: CLIP (16 heads x 64, S = 577, non-causal), Resampler cross-attention (8 x 128,
// 144/576 shared queries x 576 keys), Llama decoder (h x 128, causal, GQA, packed variable-length rows).
#include "attention.h"
#include "errors.h"
#include "gemm.h"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int SLAB_BYTES = 128 * 128;  // 128 rows x 64 bf16
constexpr int NT = 64 + 8 * 32;  // TMA warp + MMA warp + 8 softmax/epilogue warps
constexpr float RESCALE_THRESHOLD = 8.0f;  // in log2 units: P stays below 2^8

template <int HD>
struct TcCfg {
  static constexpr int SLABS = HD / 64;
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  // Q is double-buffered per ITEM so the first S MMA of the next item can be issued before the last PV of the
  // current one; K / V rings: 2/2 stages at hd 128 (7 tiles of 32 KB do not fit next to the epilogue staging),
  // 4/3 at hd 64.  K runs NK-1 tiles ahead of V.
  static constexpr int NK = HD == 128 ? 2 : 4;
  static constexpr int NV = HD == 128 ? 2 : 3;
  static constexpr int BAR_BYTES = 256;            // 28 mbarriers + the TMEM base address
  static constexpr int XCH_BYTES = 6 * 128 * 4;    // half-row max (2 slots x 2 halves) and item sums (2 halves)
  static constexpr int STAGE_BYTES = 8 * 32 * 64;  // epilogue: per softmax warp 32 rows x 64 B (one 32-column chunk)
  static constexpr int RING_BYTES = 8 * 64;        // decoded work items, written by the producer
  static constexpr int SMEM_RAW = 1024 + TILE_BYTES * (2 + NK + NV) + BAR_BYTES + XCH_BYTES + STAGE_BYTES + RING_BYTES;
  // >= 120 KB so that two CTAs can never share an SM (each allocates all 512 TMEM columns)
  static constexpr int SMEM_BYTES = SMEM_RAW > 120 * 1024 ? SMEM_RAW : 120 * 1024;
  static constexpr int TMEM_COLS = 512;
  // S double-buffered per kv tile, O double-buffered per ITEM (the epilogue of item i runs after the first score
  // tile of item i+1 has been handed to the tensor pipe, so PV of item i+1 must not touch O of item i)
  static constexpr int S_COL0 = 0, S_COL1 = 128, O_COL = 256, O_STRIDE = 128;
};

// MUFU.EX2 without the denormal/range fix-up code exp2f() adds (inputs here are <= 8, -inf -> 0)
SLIME_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// One unit of work: a 128-row query tile of one (sequence, head).  Decoded once by the producer warp and handed to
// the other warps through an 8-entry ring in shared memory (`more` says whether another item follows; a CTA
// without any work posts one entry with n_tiles == 0).
struct alignas(64) Item {
  int n_tiles;  // kv tiles this query tile attends
  int t;        // query tile index inside the sequence
  int head, kv_head;
  int len_q, len_k, causal_off;
  int q_row0, k_row0;  // first row of this sequence in the q / kv matrices
  int more;            // another item follows in this CTA's list
  long long o_row0;
};
static_assert(sizeof(Item) <= 64, "ring slot");

template <bool CAUSAL>
SLIME_DEVINL Item decode_item(const AttnParams& p, int w, int q_tiles) {
  Item it;
  int b;
  if (CAUSAL) {
    // heavy (late) query tiles first for load balance; the q heads of one GQA group are adjacent (shared K/V in L2)
    it.head = w % p.num_heads;
    const int rest = w / p.num_heads;
    b = rest % p.batch;
    it.t = q_tiles - 1 - rest / p.batch;
  } else {
    // all query tiles of one (sequence, head) are adjacent, so their common K/V stream is read from DRAM once
    it.t = w % q_tiles;
    const int rest = w / q_tiles;
    it.head = rest % p.num_heads;
    b = rest / p.num_heads;
  }
  it.kv_head = it.head / (p.num_heads / p.num_kv_heads);
  if (p.cu_q != nullptr) {
    it.q_row0 = p.cu_q[b];
    it.len_q = p.cu_q[b + 1] - it.q_row0;
    it.o_row0 = it.q_row0;
  } else {
    it.q_row0 = static_cast<int>(b * p.q_batch_rows);
    it.o_row0 = b * p.o_batch_rows;
    it.len_q = p.seqlen_q;
  }
  if (p.cu_k != nullptr) {
    it.k_row0 = p.cu_k[b];
    it.len_k = p.cu_k[b + 1] - it.k_row0;
  } else {
    it.k_row0 = static_cast<int>(b * p.k_batch_rows);
    it.len_k = p.seqlen_k;
  }
  it.causal_off = it.len_k - it.len_q;
  it.more = 0;
  const int m0 = it.t * BM;
  int n = 0;
  if (m0 < it.len_q && it.len_k > 0) {
    int last = it.len_k;
    if (CAUSAL) last = min(it.len_k, m0 + BM + it.causal_off);
    n = max(0, (last + BN - 1) / BN);
  }
  it.n_tiles = n;
  return it;
}

// Epilogue of one item for one softmax warp: O / l -> bf16 -> HBM.  Each thread owns HD/2 columns of one row in
// TMEM; writing those straight to HBM makes every store instruction touch 32 different rows (32 L1 wavefronts per
// instruction, ~2 k cycles per item with the tensor pipe idle - the gap between items in
// profiles/r01_attention_clock_trace.txt).  Instead each warp transposes one 32-column chunk at a time through its
// own 2 KB of shared memory (16-byte slots XOR-swizzled by row pair: conflict-free both ways) and writes 8 rows x 64
// contiguous bytes per instruction.
template <int HD>
SLIME_DEVINL void epilogue_tile(uint64_t* o_done_bar, uint32_t o_done_parity, uint64_t* o_free_bar, uint32_t o_addr,
                                uint8_t* stage, int lane, int quad, float inv_l, int rows_valid, bf16* out, int o_ld,
                                long long* tr) {
  if (tr) tr[6] = clock64();
  mbar_wait(o_done_bar, o_done_parity);
  tcgen05_fence_after();
  if (tr) tr[7] = clock64();
#pragma unroll
  for (int c = 0; c < HD / 64; ++c) {
    uint32_t orow[32];
    tmem_ld_32x32b_x32(o_addr + c * 32, orow);
    tmem_ld_wait();
    if (tr && c == 0) tr[12] = clock64();
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      uint4 pkv;
      pkv.x = pack_bf16x2(__uint_as_float(orow[q * 8 + 0]) * inv_l, __uint_as_float(orow[q * 8 + 1]) * inv_l);
      pkv.y = pack_bf16x2(__uint_as_float(orow[q * 8 + 2]) * inv_l, __uint_as_float(orow[q * 8 + 3]) * inv_l);
      pkv.z = pack_bf16x2(__uint_as_float(orow[q * 8 + 4]) * inv_l, __uint_as_float(orow[q * 8 + 5]) * inv_l);
      pkv.w = pack_bf16x2(__uint_as_float(orow[q * 8 + 6]) * inv_l, __uint_as_float(orow[q * 8 + 7]) * inv_l);
      *reinterpret_cast<uint4*>(stage + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = pkv;
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = i * 8 + (lane >> 2), q = lane & 3;  // 8 rows x 4 slots per instruction
      const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
      if (quad * 32 + r < rows_valid)
        *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * o_ld + c * 32 + q * 8) = v;
    }
    __syncwarp();
    if (tr && c == 0) tr[13] = clock64();
  }
  tcgen05_fence_before();
  mbar_arrive(o_free_bar);
  if (tr) tr[14] = clock64();
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NT, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, int q_tiles, int total_items) {
  using Cfg = TcCfg<HD>;
  constexpr int NK = Cfg::NK, NV = Cfg::NV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                        // [2]
  uint8_t* sK = sQ + 2 * Cfg::TILE_BYTES;    // [NK]
  uint8_t* sV = sK + NK * Cfg::TILE_BYTES;   // [NV]
  uint8_t* aux = sV + NV * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* q_full = bars + 0;    // [2] per Q buffer = item parity
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* k_full = bars + 4;    // [NK <= 4]
  uint64_t* k_empty = bars + 8;   // [NK]
  uint64_t* v_full = bars + 12;   // [NV <= 4]
  uint64_t* v_empty = bars + 16;  // [NV]
  uint64_t* s_full = bars + 20;   // [2]
  uint64_t* p_ready = bars + 22;  // [2]
  uint64_t* o_done = bars + 24;   // [2]: PV with global index g commits o_done[g & 1] (phase g >> 1), so a waiter is
                                  // never more than one phase behind and parity waits stay unambiguous
  uint64_t* o_free = bars + 26;   // [2] per O buffer = item parity
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 28);
  float* xch = reinterpret_cast<float*>(aux + Cfg::BAR_BYTES);
  uint8_t* stage_all = aux + Cfg::BAR_BYTES + Cfg::XCH_BYTES;  // [8 warps][32 rows][64 B]
  Item* ring = reinterpret_cast<Item*>(aux + Cfg::BAR_BYTES + Cfg::XCH_BYTES + Cfg::STAGE_BYTES);  // [8] x 64 B

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_ready[s], 256);
      mbar_init(&o_done[s], 1);
      mbar_init(&o_free[s], 256);
    }
    for (int s = 0; s < NK; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < NV; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    fence_barrier_init();
  } else if (warp_idx == 1) {
    tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp_idx == 0) {
    // ================================ TMA producer ================================
    // Item k uses Q buffer / barrier phase k & 1 and ring slot k & 7 (the producer runs at most two items ahead of
    // the last S MMA, the deferred epilogue at most two items behind it).  The Q tile of item k+1 is
    // posted BEFORE the K/V stream of item k, so the MMA warp can issue S(0) of item k+1 ahead of the last PV of k.
    if (lane == 0) {
      int w = blockIdx.x;
      auto next_valid = [&]() {
        Item it;
        it.n_tiles = 0;
        while (w < total_items) {
          it = decode_item<CAUSAL>(p, w, q_tiles);
          w += gridDim.x;
          if (it.n_tiles > 0) return it;
        }
        it.n_tiles = 0;
        return it;
      };
      auto post = [&](const Item& it, int k) {
        const int qb = k & 1;
        mbar_wait(&q_empty[qb], ((k >> 1) & 1) ^ 1);  // last S MMA of item k-2 done (its ring slot is long dead)
        ring[k & 7] = it;
        if (it.n_tiles > 0) {
          mbar_arrive_expect_tx(&q_full[qb], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sQ + qb * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_q, &q_full[qb], it.head * HD + s * 64,
                        it.q_row0 + it.t * BM);
        } else {
          mbar_arrive(&q_full[qb]);  // this CTA has no work at all
        }
      };
      Item cur = next_valid();
      Item nxt = cur;
      if (cur.n_tiles > 0) {
        nxt = next_valid();
        cur.more = nxt.n_tiles > 0;
      }
      post(cur, 0);
      int g = 0;
      for (int k = 0; cur.n_tiles > 0; ++k) {
        Item nn = nxt;
        if (nxt.n_tiles > 0) {
          nn = next_valid();
          nxt.more = nn.n_tiles > 0;
          post(nxt, k + 1);
        }
        auto load_k = [&](int j) {
          const int gi = g + j, st = gi % NK;
          mbar_wait(&k_empty[st], ((gi / NK) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sK + st * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_k, &k_full[st], cur.kv_head * HD + s * 64,
                        cur.k_row0 + j * BN);
        };
        auto load_v = [&](int j) {
          const int gi = g + j, st = gi % NV;
          mbar_wait(&v_empty[st], ((gi / NV) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sV + st * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_v, &v_full[st], cur.kv_head * HD + s * 64,
                        cur.k_row0 + j * BN);
        };
        constexpr int KA = NK - 1;  // K tiles in flight ahead of V
        for (int j = 0; j < KA && j < cur.n_tiles; ++j) load_k(j);
        for (int j = 0; j < cur.n_tiles; ++j) {
          if (j + KA < cur.n_tiles) load_k(j + KA);
          load_v(j);
        }
        g += cur.n_tiles;
        cur = nxt;
        nxt = nn;
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ==================================
    // One continuous stream of tiles g = 0, 1, 2, ... across items: S(g+1) is always issued before PV(g), also when
    // tile g+1 is the first tile of the next item.
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(BM, BN, 0, 0);   // Q, K both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(BM, HD, 0, 1);  // P from TMEM, V MN-major

      auto issue_s = [&](int k, int gi, bool last_of_item) {
        const int st = gi % NK;
        mbar_wait(&k_full[st], (gi / NK) & 1);
        tcgen05_fence_after();
        const uint32_t sQ_addr = smem_u32(sQ + (k & 1) * Cfg::TILE_BYTES);
        const uint32_t sK_addr = smem_u32(sK + st * Cfg::TILE_BYTES);
        const uint32_t tmem_s = tmem_base + ((gi & 1) ? Cfg::S_COL1 : Cfg::S_COL0);
#pragma unroll
        for (int s = 0; s < Cfg::SLABS; ++s) {
          const uint64_t dq = make_umma_desc_sw128(sQ_addr + s * SLAB_BYTES);
          const uint64_t dk = make_umma_desc_sw128(sK_addr + s * SLAB_BYTES);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_ss(tmem_s, dq + 2 * kk, dk + 2 * kk, idesc_s, (s | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[gi & 1]);
        if (last_of_item) umma_commit(&q_empty[k & 1]);
      };

      mbar_wait(&q_full[0], 0);
      int n = ring[0].n_tiles;
      int more = ring[0].more;
      int g = 0;
      if (n > 0) {
        tcgen05_fence_after();
        issue_s(0, 0, n == 1);
      }
      for (int k = 0; n > 0; ++k) {
        int n_next = 0, more_next = 0;
        for (int j = 0; j < n; ++j) {
          const int gj = g + j;
          const int sb = gj & 1;    // S / P buffer
          const int vs = gj % NV;   // V ring stage
          if (j + 1 < n) {
            issue_s(k, gj + 1, j + 2 == n);
          } else if (more) {
            // first tile of the next item (its Q / ring entry were posted before this item's K/V stream)
            mbar_wait(&q_full[(k + 1) & 1], ((k + 1) >> 1) & 1);
            n_next = ring[(k + 1) & 7].n_tiles;
            more_next = ring[(k + 1) & 7].more;
            tcgen05_fence_after();
            issue_s(k + 1, gj + 1, n_next == 1);
          }
          // the epilogue of item k-2 has drained this O buffer
          if (j == 0) mbar_wait(&o_free[k & 1], ((k >> 1) & 1) ^ 1);
          const bool tr = p.trace != nullptr && blockIdx.x == 0 && gj < 64;
          if (tr) p.trace[gj * 16 + 8] = clock64();
          mbar_wait(&p_ready[sb], (gj >> 1) & 1);
          if (tr) p.trace[gj * 16 + 9] = clock64();
          mbar_wait(&v_full[vs], (gj / NV) & 1);
          tcgen05_fence_after();
          if (tr) p.trace[gj * 16 + 10] = clock64();
          const uint32_t tmem_p = tmem_base + (sb ? Cfg::S_COL1 : Cfg::S_COL0);
          const uint64_t dv = make_umma_desc_mn_sw128(smem_u32(sV + vs * Cfg::TILE_BYTES), SLAB_BYTES);
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            // A: 16 kv positions = 8 TMEM columns of packed bf16 pairs;  B: 16 kv rows = 2048 bytes further down
            umma_bf16_ts(tmem_base + Cfg::O_COL + (k & 1) * Cfg::O_STRIDE, tmem_p + kk * 8,
                         dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv, (j | kk) != 0 ? 1u : 0u);
          }
          umma_commit(&v_empty[vs]);
          umma_commit(&o_done[gj & 1]);
          if (tr) p.trace[gj * 16 + 11] = clock64();
        }
        g += n;
        n = n_next;
        more = more_next;
      }
    }
  } else {
    // ================================ softmax + epilogue (8 warps) ================
    // Two threads per query row: warps w and w+4 share TMEM lane quadrant w % 4 and split the 128 score
    // columns (and the O columns) in halves; they exchange their half-row max (per tile) and sum (per item)
    // through shared memory with a 64-thread named barrier.
    const int quad = warp_idx & 3;
    const int half = (warp_idx - 2) >> 2;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float scale_log2 = p.scale * 1.4426950408889634f;
    const int pair_bar = 1 + quad;  // named barrier id (0 is __syncthreads)
    int g = 0;
    // deferred epilogue of the previous item (its description is still in the ring)
    float pend_inv_l = 0.f;
    int pend_g_last = -1;
    auto run_epilogue = [&](int k_prev, long long* tr) {
      const Item* pi = &ring[k_prev & 7];
      const int obuf = k_prev & 1;
      epilogue_tile<HD>(&o_done[pend_g_last & 1], (pend_g_last >> 1) & 1, &o_free[obuf],
                        tmem_base + lane_addr + Cfg::O_COL + obuf * Cfg::O_STRIDE + half * (HD / 2),
                        stage_all + (warp_idx - 2) * (32 * 64), lane, quad, pend_inv_l, pi->len_q - pi->t * BM,
                        p.o + (pi->o_row0 + pi->t * BM + quad * 32) * p.o_ld + pi->head * HD + half * (HD / 2), p.o_ld, tr);
    };
    // Item 0 is announced by q_full[0]; every later item by the completion of its first S MMA (the MMA warp waited
    // for the item's q_full before issuing it).  Waiting on q_full here for k > 0 would be wrong: with single-tile
    // items the producer can post item k+2 into the same barrier before these warps reach item k, and a parity wait
    // must never be two phases behind.
    mbar_wait(&q_full[0], 0);
    int k = 0;
    for (;; ++k) {
      if (k > 0) mbar_wait(&s_full[g & 1], (g >> 1) & 1);
      const int n_tiles = ring[k & 7].n_tiles;
      if (n_tiles == 0) break;
      const int more = ring[k & 7].more;
      const int it_t = ring[k & 7].t, len_k = ring[k & 7].len_k, causal_off = ring[k & 7].causal_off;
      const int row = it_t * BM + r_in_tile;  // query index inside the sequence
      float m_ref = -INFINITY;                // raw-score max the exponentials are taken against
      float l_sum = 0.f;                      // sum over THIS thread's 64 columns of every tile
      for (int j = 0; j < n_tiles; ++j) {
        const int gj = g + j;
        const int buf = gj & 1;
        const uint32_t s_base = tmem_base + lane_addr + (buf ? Cfg::S_COL1 : Cfg::S_COL0);
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && warp_idx == 2 && lane == 0 && gj < 64;
        if (tr) p.trace[gj * 16 + 0] = clock64();
        mbar_wait(&s_full[buf], (gj >> 1) & 1);
        tcgen05_fence_after();
        if (tr) p.trace[gj * 16 + 1] = clock64();
        uint32_t sr[64];
        tmem_ld_32x32b_x32(s_base + half * 64, sr);
        tmem_ld_32x32b_x32(s_base + half * 64 + 32, sr + 32);
        tmem_ld_wait();
        if (tr) p.trace[gj * 16 + 2] = clock64();

        const int col_base = j * BN + half * 64;
        const bool need_mask = (j * BN + BN > len_k) || (CAUSAL && (j * BN + BN - 1 > it_t * BM + causal_off));
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains (one warp per scheduler)
        if (need_mask) {
          const int limit = CAUSAL ? min(len_k - 1, row + causal_off) : len_k - 1;  // last visible column
#pragma unroll
          for (int c = 0; c < 64; ++c) {
            float v = __uint_as_float(sr[c]);
            if (col_base + c > limit) v = -INFINITY;
            sr[c] = __float_as_uint(v);
            mx4[c & 3] = fmaxf(mx4[c & 3], v);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 64; ++c) mx4[c & 3] = fmaxf(mx4[c & 3], __uint_as_float(sr[c]));
        }
        const float m_half = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        xch[((j & 1) * 2 + half) * BM + r_in_tile] = m_half;
        asm volatile("bar.sync %0, 64;\n" ::"r"(pair_bar) : "memory");
        const float m_tile = fmaxf(m_half, xch[((j & 1) * 2 + (half ^ 1)) * BM + r_in_tile]);
        if (tr) p.trace[gj * 16 + 3] = clock64();

        // lazy rescale: only move the reference max when it grew by more than 2^8 (or on the first tile).
        // Both threads of a row see the same m_tile / m_ref, so the two warps take the same branches.
        bool grow = (m_tile - m_ref) * scale_log2 > RESCALE_THRESHOLD;
        if (m_tile == -INFINITY) grow = false;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2((m_ref - m_tile) * scale_log2) : 1.0f;  // m_ref == -inf -> 0
          l_sum *= alpha;
          mbar_wait(&o_done[(gj - 1) & 1], ((gj - 1) >> 1) & 1);  // PV_{j-1} finished: O is stable until PV_j
          tcgen05_fence_after();
          const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + (k & 1) * Cfg::O_STRIDE + half * (HD / 2);
#pragma unroll
          for (int c = 0; c < HD / 64; ++c) {
            uint32_t orow[32];
            tmem_ld_32x32b_x32(o_addr + c * 32, orow);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * alpha);
            tmem_st_32x32b_x32(o_addr + c * 32, orow);
          }
          tmem_st_wait();
        }
        if (grow) m_ref = m_tile;
        const float m_scaled = (m_ref == -INFINITY) ? 0.f : m_ref * scale_log2;

        uint32_t pk[32];
        float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float p0 = fast_exp2(fmaf(__uint_as_float(sr[2 * c]), scale_log2, -m_scaled));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sr[2 * c + 1]), scale_log2, -m_scaled));
          ps4[c & 3] += p0 + p1;
          pk[c] = pack_bf16x2(p0, p1);
        }
        l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        if (tr) p.trace[gj * 16 + 4] = clock64();
        tmem_st_32x32b_x32(s_base + half * 32, pk);  // P: 64 packed columns per row, this thread's half
        tmem_st_wait();
        tcgen05_fence_before();
        if (tr) p.trace[gj * 16 + 5] = clock64();
        mbar_arrive(&p_ready[buf]);
        if (j == 0 && pend_g_last >= 0) {
          // previous item's O: its last PV finished long ago; S(1) / PV(0) of this item keep the tensor pipe busy
          run_epilogue(k - 1, tr ? p.trace + gj * 16 : nullptr);
          pend_g_last = -1;
        }
      }
      // ---- item end: total row sum now (the exchange slots are reused by the next item); the epilogue itself is
      //      deferred until the first score tile of the next item is with the tensor pipe ----
      xch[(4 + half) * BM + r_in_tile] = l_sum;
      asm volatile("bar.sync %0, 64;\n" ::"r"(pair_bar) : "memory");
      const float l_tot = l_sum + xch[(4 + (half ^ 1)) * BM + r_in_tile];
      pend_inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      pend_g_last = g + n_tiles - 1;
      g += n_tiles;
      if (!more) break;
    }
    if (pend_g_last >= 0) run_epilogue(k, nullptr);
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int HD, bool CAUSAL>
int launch_tc(const AttnParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = TcCfg<HD>;
  auto kern = attn_tc_kernel<HD, CAUSAL>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  // Tensor maps over the packed row matrices; the kernel addresses heads by column offset.  Base pointers are
  // rounded down to the start of their row so q / k / v views of one packed qkv buffer share a map shape.
  const long long q_rows = p.total_q_rows > 0 ? p.total_q_rows
                           : (p.q_batch_rows > 0 ? p.q_batch_rows * p.batch : p.seqlen_q);
  const long long k_rows = p.total_k_rows > 0 ? p.total_k_rows
                           : (p.k_batch_rows > 0 ? p.k_batch_rows * p.batch : p.seqlen_k);
  CUtensorMap tq, tk, tv;
  SLIME_PROPAGATE(slime_get_tmap(p.q, static_cast<int>(q_rows), p.num_heads * HD, p.q_ld, BM, &tq));
  SLIME_PROPAGATE(slime_get_tmap(p.k, static_cast<int>(k_rows), p.num_kv_heads * HD, p.k_ld, BN, &tk));
  SLIME_PROPAGATE(slime_get_tmap(p.v, static_cast<int>(k_rows), p.num_kv_heads * HD, p.v_ld, BN, &tv));
  const int q_tiles = (p.seqlen_q + BM - 1) / BM;
  const int total = q_tiles * p.num_heads * p.batch;
  const int grid = total < num_sms ? total : num_sms;
  double flops = 0.0;
  if (p.cu_q == nullptr)
    flops = 4.0 * p.seqlen_q * static_cast<double>(p.seqlen_k) * HD * p.num_heads * p.batch * (CAUSAL ? 0.5 : 1.0);
  slime_prof_begin(1, flops, stream);
  kern<<<grid, NT, Cfg::SMEM_BYTES, stream>>>(tq, tk, tv, p, q_tiles, total);
  slime_prof_end(stream);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

}  // namespace

int slime_launch_attention_tc(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (p.batch <= 0 || p.seqlen_q <= 0 || p.seqlen_k <= 0) return SLIME_OK;
  if (p.head_dim == 64) {
    return p.causal ? launch_tc<64, true>(p, num_sms, stream) : launch_tc<64, false>(p, num_sms, stream);
  }
  return p.causal ? launch_tc<128, true>(p, num_sms, stream) : launch_tc<128, false>(p, num_sms, stream);
}
