// tcgen05 / TMEM flash attention forward, "ping-pong" version: one persistent CTA per SM works on TWO 128-row
// query tiles (A, B) of the same (sequence, head) at a time, sharing one K/V stream.
//
// Why: the clock trace of the single-tile kernel (profiles/r01_attention_clock_trace.txt) shows a 128 x 128 score
// tile needs ~1240 cycles of exp2 on the SFUs (16 results / clock / SM) - as long as its two MMAs (2 x 512) - and
// with one tile in flight the tensor pipe idles while the softmax runs.  With two tiles in flight the MMAs of one
// tile overlap the softmax of the other:
//     tensor pipe :  PV_A(j)  S_A(j+1)  PV_B(j)  S_B(j+1)  PV_A(j+1) ...
//     SFU / ALU   :  softmax_B(j) .......  softmax_A(j+1) .......  softmax_B(j+1) ...
//
//   warp 0      : TMA producer (Q_A, Q_B once per item; K/V ring, 64-column slabs, 128-byte swizzle)
//   warp 1      : tcgen05.mma issuer + TMEM allocator
//   warps 2..5  : softmax + epilogue of tile A (thread r <-> query row r, TMEM lane r)
//   warps 6..9  : softmax + epilogue of tile B
// TMEM: S_A [0,128)  S_B [128,256)  O_A [256,256+HD)  O_B [384,384+HD).  P (bf16) overwrites the first 64 columns of
// its S; PV is a TS-MMA (A from TMEM, V MN-major in shared memory); lazy rescale of O as in attention_tc.cu.
#include "attention.h"
#include "errors.h"
#include "gemm.h"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int SLAB_BYTES = 128 * 128;
constexpr int NT = 320;
constexpr float RESCALE_THRESHOLD = 8.0f;

template <int HD>
struct Cfg2 {
  static constexpr int SLABS = HD / 64;
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int NKV = HD == 128 ? 2 : 3;
  static constexpr int SMEM_RAW = 1024 + TILE_BYTES * (2 + 2 * NKV) + 512;
  static constexpr int SMEM_BYTES = SMEM_RAW > 120 * 1024 ? SMEM_RAW : 120 * 1024;  // one CTA per SM (512 TMEM columns)
  static constexpr int TMEM_COLS = 512;
  static constexpr int S_COL = 0, O_COL = 256;  // + X * 128
};

SLIME_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct Item2 {
  int b, head, kv_head;
  int t[2];      // query tile index of A and B
  int n[2];      // kv tiles of A and B (0 = tile not valid)
  int n_max;
  int len_q, len_k, causal_off;
  int q_row0, k_row0;
  long long o_row0;
  bool valid;
};

template <bool CAUSAL>
SLIME_DEVINL Item2 decode_item2(const AttnParams& p, int w, int q_pairs) {
  Item2 it;
  int u;
  if (CAUSAL) {
    it.head = w % p.num_heads;
    const int rest = w / p.num_heads;
    it.b = rest % p.batch;
    u = q_pairs - 1 - rest / p.batch;  // heavy (late) pairs first
  } else {
    u = w % q_pairs;
    const int rest = w / q_pairs;
    it.head = rest % p.num_heads;
    it.b = rest / p.num_heads;
  }
  it.kv_head = it.head / (p.num_heads / p.num_kv_heads);
  if (p.cu_q != nullptr) {
    it.q_row0 = p.cu_q[it.b];
    it.len_q = p.cu_q[it.b + 1] - it.q_row0;
    it.o_row0 = it.q_row0;
  } else {
    it.q_row0 = static_cast<int>(it.b * p.q_batch_rows);
    it.o_row0 = it.b * p.o_batch_rows;
    it.len_q = p.seqlen_q;
  }
  if (p.cu_k != nullptr) {
    it.k_row0 = p.cu_k[it.b];
    it.len_k = p.cu_k[it.b + 1] - it.k_row0;
  } else {
    it.k_row0 = static_cast<int>(it.b * p.k_batch_rows);
    it.len_k = p.seqlen_k;
  }
  it.causal_off = it.len_k - it.len_q;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    it.t[x] = 2 * u + x;
    const int m0 = it.t[x] * BM;
    int nt = 0;
    if (m0 < it.len_q && it.len_k > 0) {
      int last = it.len_k;
      if (CAUSAL) last = min(it.len_k, m0 + BM + it.causal_off);
      nt = max(0, (last + BN - 1) / BN);
    }
    it.n[x] = nt;
  }
  it.n_max = max(it.n[0], it.n[1]);
  it.valid = it.n[0] > 0;  // B is only ever valid together with A (later rows, at least as many kv tiles)
  return it;
}

template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(NT, 1)
attn_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, int q_pairs, int total_items) {
  using Cfg = Cfg2<HD>;
  constexpr int NKV = Cfg::NKV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                                  // [2] tiles A, B
  uint8_t* sK = sQ + 2 * Cfg::TILE_BYTES;              // [NKV]
  uint8_t* sV = sK + NKV * Cfg::TILE_BYTES;            // [NKV]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + NKV * Cfg::TILE_BYTES);
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;    // [3]
  uint64_t* k_empty = bars + 5;   // [3]
  uint64_t* v_full = bars + 8;    // [3]
  uint64_t* v_empty = bars + 11;  // [3]
  uint64_t* s_full = bars + 14;   // [2] per query tile
  uint64_t* p_ready = bars + 16;  // [2]
  uint64_t* o_done = bars + 18;   // [2][2]: tile X, parity of the PV index (waiters stay within one phase)
  uint64_t* o_free = bars + 22;   // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 3; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    for (int x = 0; x < 2; ++x) {
      mbar_init(&s_full[x], 1);
      mbar_init(&p_ready[x], 128);
      mbar_init(&o_done[2 * x], 1);
      mbar_init(&o_done[2 * x + 1], 1);
      mbar_init(&o_free[x], 128);
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
    if (lane == 0) {
      int item_cnt = 0, g = 0;
      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const Item2 it = decode_item2<CAUSAL>(p, w, q_pairs);
        if (!it.valid) continue;
        const int ntiles_q = it.n[1] > 0 ? 2 : 1;
        mbar_wait(q_empty, (item_cnt & 1) ^ 1);
        mbar_arrive_expect_tx(q_full, ntiles_q * Cfg::TILE_BYTES);
        for (int x = 0; x < ntiles_q; ++x) {
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sQ + x * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_q, q_full, it.head * HD + s * 64,
                        it.q_row0 + it.t[x] * BM);
        }
        auto load_k = [&](int j) {
          const int gi = g + j, st = gi % NKV;
          mbar_wait(&k_empty[st], ((gi / NKV) & 1) ^ 1);
          mbar_arrive_expect_tx(&k_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sK + st * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_k, &k_full[st], it.kv_head * HD + s * 64,
                        it.k_row0 + j * BN);
        };
        auto load_v = [&](int j) {
          const int gi = g + j, st = gi % NKV;
          mbar_wait(&v_empty[st], ((gi / NKV) & 1) ^ 1);
          mbar_arrive_expect_tx(&v_full[st], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sV + st * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_v, &v_full[st], it.kv_head * HD + s * 64,
                        it.k_row0 + j * BN);
        };
        load_k(0);
        for (int j = 0; j < it.n_max; ++j) {
          if (j + 1 < it.n_max) load_k(j + 1);  // K one tile ahead of V: S(j+1) is issued before PV(j) completes
          load_v(j);
        }
        g += it.n_max;
        ++item_cnt;
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(BM, HD, 0, 1);
      int item_cnt = 0, g = 0;
      int h[2] = {0, 0};       // kv iterations issued so far per query tile (S / P / O_done phase counters)
      int items[2] = {0, 0};   // items in which the tile was valid (o_free phase counters)

      auto issue_s = [&](int x, int gi, int idx) {
        const int st = gi % NKV;
        const uint32_t sQ_addr = smem_u32(sQ + x * Cfg::TILE_BYTES);
        const uint32_t sK_addr = smem_u32(sK + st * Cfg::TILE_BYTES);
        const uint32_t tmem_s = tmem_base + Cfg::S_COL + x * 128;
#pragma unroll
        for (int s = 0; s < Cfg::SLABS; ++s) {
          const uint64_t dq = make_umma_desc_sw128(sQ_addr + s * SLAB_BYTES);
          const uint64_t dk = make_umma_desc_sw128(sK_addr + s * SLAB_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_s, dq + 2 * k, dk + 2 * k, idesc_s, (s | k) != 0 ? 1u : 0u);
        }
        (void)idx;
        umma_commit(&s_full[x]);
      };
      auto wait_k = [&](int gi) {
        mbar_wait(&k_full[gi % NKV], (gi / NKV) & 1);
        tcgen05_fence_after();
      };
      auto issue_pv = [&](int x, int gi, int j, int idx) {
        const int st = gi % NKV;
        const uint32_t tmem_p = tmem_base + Cfg::S_COL + x * 128;
        const uint32_t tmem_o = tmem_base + Cfg::O_COL + x * 128;
        const uint64_t dv = make_umma_desc_mn_sw128(smem_u32(sV + st * Cfg::TILE_BYTES), SLAB_BYTES);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)
          umma_bf16_ts(tmem_o, tmem_p + kk * 8, dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv,
                       (j | kk) != 0 ? 1u : 0u);
        umma_commit(&o_done[2 * x + (idx & 1)]);
      };

      for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
        const Item2 it = decode_item2<CAUSAL>(p, w, q_pairs);
        if (!it.valid) continue;
        const int n = it.n_max;
        mbar_wait(q_full, item_cnt & 1);
        tcgen05_fence_after();
        // S(0) of both tiles
        wait_k(g);
        issue_s(0, g, h[0]);
        if (it.n[1] > 0) issue_s(1, g, h[1]);
        umma_commit(&k_empty[g % NKV]);
        if (n == 1) umma_commit(q_empty);
        for (int j = 0; j < n; ++j) {
          const int gj = g + j;
          bool v_waited = false, k_waited = false;
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            if (j < it.n[x]) {
              const int idx = h[x] + j;
              if (j == 0) mbar_wait(&o_free[x], (items[x] & 1) ^ 1);  // the tile's previous epilogue has drained O_x
              mbar_wait(&p_ready[x], idx & 1);
              if (!v_waited) {
                mbar_wait(&v_full[gj % NKV], (gj / NKV) & 1);
                v_waited = true;
              }
              tcgen05_fence_after();
              issue_pv(x, gj, j, idx);
            }
            // V_j is free once the last PV that reads it has been issued (B's when B is still running, else A's)
            if (x == 1) umma_commit(&v_empty[gj % NKV]);
            if (j + 1 < it.n[x]) {
              if (!k_waited) {
                wait_k(gj + 1);
                k_waited = true;
              }
              issue_s(x, gj + 1, h[x] + j + 1);
            }
          }
          if (j + 1 < n) {
            umma_commit(&k_empty[(gj + 1) % NKV]);  // K_{j+1} read by the S MMAs issued above
            if (j + 2 == n) umma_commit(q_empty);   // those were the last S MMAs of the item
          }
        }
        g += n;
        ++item_cnt;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (it.n[x] > 0) {
            h[x] += it.n[x];
            ++items[x];
          }
        }
      }
    }
  } else {
    // ================================ softmax + epilogue of tile X ==================
    const int x = (warp_idx - 2) >> 2;  // warps 2..5 -> A, 6..9 -> B
    const int quad = warp_idx & 3;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_base = tmem_base + lane_addr + Cfg::S_COL + x * 128;
    const uint32_t o_base = tmem_base + lane_addr + Cfg::O_COL + x * 128;
    const float scale_log2 = p.scale * 1.4426950408889634f;
    int hx = 0;
    for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
      const Item2 it = decode_item2<CAUSAL>(p, w, q_pairs);
      if (!it.valid || it.n[x] == 0) continue;
      const int nx = it.n[x];
      const int tq = it.t[x];
      const int row = tq * BM + r_in_tile;
      float m_ref = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < nx; ++j) {
        const int idx = hx + j;
        mbar_wait(&s_full[x], idx & 1);
        tcgen05_fence_after();
        const int col_base = j * BN;
        const bool need_mask = (col_base + BN > it.len_k) || (CAUSAL && (col_base + BN - 1 > tq * BM + it.causal_off));
        const int limit = CAUSAL ? min(it.len_k - 1, row + it.causal_off) : it.len_k - 1;  // last visible column

        // ---- pass 1: row max (scores stay in TMEM; only the running max is kept) ----
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sr[64];
          tmem_ld_32x32b_x32(s_base + c * 64, sr);
          tmem_ld_32x32b_x32(s_base + c * 64 + 32, sr + 32);
          tmem_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int i = 0; i < 64; ++i) {
              float v = __uint_as_float(sr[i]);
              if (col_base + c * 64 + i > limit) v = -INFINITY;
              mx4[i & 3] = fmaxf(mx4[i & 3], v);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sr[i]));
          }
        }
        const float m_tile = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));

        bool grow = (m_tile - m_ref) * scale_log2 > RESCALE_THRESHOLD;
        if (m_tile == -INFINITY) grow = false;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2((m_ref - m_tile) * scale_log2) : 1.0f;
          l_sum *= alpha;
          mbar_wait(&o_done[2 * x + ((idx - 1) & 1)], ((idx - 1) >> 1) & 1);  // PV_x(j-1) done: O_x stable until PV_x(j)
          tcgen05_fence_after();
#pragma unroll
          for (int c = 0; c < HD / 32; ++c) {
            uint32_t orow[32];
            tmem_ld_32x32b_x32(o_base + c * 32, orow);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * alpha);
            tmem_st_32x32b_x32(o_base + c * 32, orow);
          }
          tmem_st_wait();
        }
        if (grow) m_ref = m_tile;
        const float m_scaled = (m_ref == -INFINITY) ? 0.f : m_ref * scale_log2;

        // ---- pass 2: P = exp2(S * scale - m), 64 columns at a time; P (bf16 pairs) overwrites S columns that
        //      have already been consumed: P[0,32) after S[0,64), P[32,64) after S[64,128) ----
        float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sr[64];
          tmem_ld_32x32b_x32(s_base + c * 64, sr);
          tmem_ld_32x32b_x32(s_base + c * 64 + 32, sr + 32);
          tmem_ld_wait();
          uint32_t pk[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float v0 = __uint_as_float(sr[2 * i]), v1 = __uint_as_float(sr[2 * i + 1]);
            if (need_mask) {
              if (col_base + c * 64 + 2 * i > limit) v0 = -INFINITY;
              if (col_base + c * 64 + 2 * i + 1 > limit) v1 = -INFINITY;
            }
            const float p0 = fast_exp2(fmaf(v0, scale_log2, -m_scaled));
            const float p1 = fast_exp2(fmaf(v1, scale_log2, -m_scaled));
            ps4[i & 3] += p0 + p1;
            pk[i] = pack_bf16x2(p0, p1);
          }
          tmem_st_32x32b_x32(s_base + c * 32, pk);
        }
        l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&p_ready[x]);
      }
      // ---- epilogue: O / l -> bf16 -> HBM ----
      const int last = hx + nx - 1;
      mbar_wait(&o_done[2 * x + (last & 1)], (last >> 1) & 1);
      tcgen05_fence_after();
      const float inv_l = l_sum > 0.f ? 1.0f / l_sum : 0.f;
      bf16* orow_ptr = p.o + (it.o_row0 + row) * p.o_ld + it.head * HD;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t orow[32];
        tmem_ld_32x32b_x32(o_base + c * 32, orow);
        tmem_ld_wait();
        if (row < it.len_q) {
#pragma unroll
          for (int v8 = 0; v8 < 4; ++v8) {
            uint4 pkv;
            pkv.x = pack_bf16x2(__uint_as_float(orow[v8 * 8 + 0]) * inv_l, __uint_as_float(orow[v8 * 8 + 1]) * inv_l);
            pkv.y = pack_bf16x2(__uint_as_float(orow[v8 * 8 + 2]) * inv_l, __uint_as_float(orow[v8 * 8 + 3]) * inv_l);
            pkv.z = pack_bf16x2(__uint_as_float(orow[v8 * 8 + 4]) * inv_l, __uint_as_float(orow[v8 * 8 + 5]) * inv_l);
            pkv.w = pack_bf16x2(__uint_as_float(orow[v8 * 8 + 6]) * inv_l, __uint_as_float(orow[v8 * 8 + 7]) * inv_l);
            *reinterpret_cast<uint4*>(orow_ptr + c * 32 + v8 * 8) = pkv;
          }
        }
      }
      tcgen05_fence_before();
      mbar_arrive(&o_free[x]);
      hx += nx;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int HD, bool CAUSAL>
int launch_tc2(const AttnParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = Cfg2<HD>;
  auto kern = attn_tc2_kernel<HD, CAUSAL>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const long long q_rows = p.total_q_rows > 0 ? p.total_q_rows
                           : (p.q_batch_rows > 0 ? p.q_batch_rows * p.batch : p.seqlen_q);
  const long long k_rows = p.total_k_rows > 0 ? p.total_k_rows
                           : (p.k_batch_rows > 0 ? p.k_batch_rows * p.batch : p.seqlen_k);
  CUtensorMap tq, tk, tv;
  SLIME_PROPAGATE(slime_get_tmap(p.q, static_cast<int>(q_rows), p.num_heads * HD, p.q_ld, BM, &tq));
  SLIME_PROPAGATE(slime_get_tmap(p.k, static_cast<int>(k_rows), p.num_kv_heads * HD, p.k_ld, BN, &tk));
  SLIME_PROPAGATE(slime_get_tmap(p.v, static_cast<int>(k_rows), p.num_kv_heads * HD, p.v_ld, BN, &tv));
  const int q_tiles = (p.seqlen_q + BM - 1) / BM;
  const int q_pairs = (q_tiles + 1) / 2;
  const int total = q_pairs * p.num_heads * p.batch;
  const int grid = total < num_sms ? total : num_sms;
  double flops = 0.0;
  if (p.cu_q == nullptr)
    flops = 4.0 * p.seqlen_q * static_cast<double>(p.seqlen_k) * HD * p.num_heads * p.batch * (CAUSAL ? 0.5 : 1.0);
  slime_prof_begin(1, flops, stream);
  kern<<<grid, NT, Cfg::SMEM_BYTES, stream>>>(tq, tk, tv, p, q_pairs, total);
  slime_prof_end(stream);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

}  // namespace

int slime_launch_attention_tc2(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (p.batch <= 0 || p.seqlen_q <= 0 || p.seqlen_k <= 0) return SLIME_OK;
  if (p.head_dim == 64) {
    return p.causal ? launch_tc2<64, true>(p, num_sms, stream) : launch_tc2<64, false>(p, num_sms, stream);
  }
  return p.causal ? launch_tc2<128, true>(p, num_sms, stream) : launch_tc2<128, false>(p, num_sms, stream);
}
