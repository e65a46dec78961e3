// tcgen05 / TMEM flash attention forward for sm_100a, two query tiles per CTA ("2Q"): bf16/fp16 in, fp32 softmax and
// accumulation, 16-bit out.
//
// Why two query tiles: with ONE 128-row query tile per CTA every 128 x 128 score tile needs its own K and V tile from
// L2 - 64 KB per 1024 tensor-pipe cycles and SM, 1.5 x what the L2 delivers to 148 SMs at once (~42 B/clk/SM) - and the
// softmax (~2 k cycles per tile: one tile at a time, all warps in the same phase) serialises with the MMAs
// (profiles/r01_attention_experiments.txt: 1615 cycles per tile even with the softmax removed, 2150 with it).  Here a
// CTA works on two query tiles (slots A, B) that share ONE K/V stream: either two q heads of the same GQA group at the
// same query tile, or two consecutive query tiles of one head.  Every K / V tile staged in shared memory feeds two
// score tiles; each slot has its own softmax warpgroup (one thread per query row: no cross-thread exchange at all), so
// while the tensor pipe runs PV_A(j) and S_A(j+1) the SFUs exponentiate slot B's tile and vice versa.
//
//   warp 0       : TMA producer - decodes the work items (posted to the other warps through a ring in shared memory),
//                  Q_A / Q_B per item, K and V tiles through 2/2-stage (hd 128) or 4/4-stage (hd 64) rings
//   warp 1       : tcgen05.mma issuer (one thread) + TMEM allocator.  Per slot X and kv tile j:
//                    S_X(j) = Q_X K_j^T   SS-MMA 128 x 128 x HD -> TMEM S_X;   O_X += P_X(j) V_j   TS-MMA (P from TMEM)
//                  issued as  PV_A(j) S_A(j+1) PV_B(j) S_B(j+1) ...  - one continuous stream, also across items
//   warps 2..5   : softmax of slot A (thread r <-> query row r = TMEM lane r): row max, lazy rescale of O (only when
//                  the max grows by > 2^8), exp2 with the scale folded in (part of the exponentials on the FMA pipe),
//                  P (16-bit pairs) overwrites the first 64 columns of S_A
//   warps 6..9   : softmax of slot B
//   warps 10..13 : epilogue of both slots: O_X / l -> 16 bit -> registers (O_X is handed back to the tensor pipe right
//                  after this read), then transposed through shared memory and stored with coalesced 64-byte row pieces
// TMEM (512 columns): S_A [0,128)  S_B [128,256)  O_A [256,256+HD)  O_B [384,384+HD).
// hd 128 fills TMEM, so P_X overwrites the first 64 columns of S_X and S_X(j+1) has to follow PV_X(j) on the tensor pipe.
// hd 64 leaves room: P_X gets its own 64 columns behind O_X ([320,384) / [448,512)), the softmax releases S_X as soon as it
// has READ the tile (half way through its exponentials) and S_X(j+1) is computed while softmax_X(j) is still running -
// the MMA round trip leaves the softmax-bound critical path of the ViT shape (EARLY_S).
//
// Shapes on the SliME path: CLIP (16 heads x 64, S = 577, non-causal), Resampler cross-attention (8 x 128, 144 / 576
// shared queries x 576 keys), Llama decoder (h x 128, causal, GQA, packed variable-length rows).
#include "attention.h"
#include "errors.h"
#include <cstdlib>
#include <type_traits>

#include "gemm.h"

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int SLAB_BYTES = 128 * 128;  // 128 rows x 64 elements
constexpr int NT = 14 * 32;  // TMA warp + MMA warp + 2 x 4 softmax warps + 4 epilogue warps.  The register file is handed
                             // out per warpGROUP, so 14 warps count as 16: 128 registers per thread (the softmax walks
                             // its 128 score columns in 32-column chunks to stay inside that without spilling)
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units: P stays below 2^8

template <int HD>
struct Cfg2 {
  static constexpr int SLABS = HD / 64;
  static constexpr int TILE_BYTES = SLABS * SLAB_BYTES;
  static constexpr int NK = HD == 128 ? 2 : 4;
  static constexpr int NV = HD == 128 ? 2 : 4;
  static constexpr int BAR_BYTES = 512;            // <= 40 mbarriers + the TMEM base address + the ring head counter
  static constexpr int LSUM_BYTES = 2 * 128 * 4;   // per slot: the row sums of the finished item (softmax -> epilogue)
  static constexpr int STAGE_BYTES = 4 * 32 * 64;  // epilogue: per warp 32 rows x 64 B (one 32-column chunk)
  static constexpr int RING_BYTES = 8 * 64;        // decoded work items
  static constexpr int SMEM_RAW = 1024 + TILE_BYTES * (2 + NK + NV) + BAR_BYTES + LSUM_BYTES + STAGE_BYTES + RING_BYTES;
  // >= 120 KB so that two CTAs can never share an SM (each allocates all 512 TMEM columns)
  static constexpr int SMEM_BYTES = SMEM_RAW > 120 * 1024 ? SMEM_RAW : 120 * 1024;
  static constexpr int TMEM_COLS = 512;
  static constexpr int S_COL = 0, O_COL = 256, SLOT_STRIDE = 128;
  // hd 64: P in its own columns behind O, S released early (see the header comment)
  static constexpr bool EARLY_S = HD == 64;
  static constexpr int P_COL = EARLY_S ? O_COL + 64 : S_COL;
};
static_assert(Cfg2<128>::SMEM_BYTES <= 232448, "shared memory budget (hd 128)");

// MUFU.EX2 without the denormal / range fix-up code exp2f() adds (inputs here are <= 8, -inf -> 0)
SLIME_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SLIME_DEVINL float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
// ---- packed fp32 pairs (FFMA2 / FADD2: one issue slot for two lanes of work) ----
SLIME_DEVINL float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{.reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
SLIME_DEVINL float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{.reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd;}\n"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// 2^x for a pair WITHOUT the SFU: Cody-Waite split x = n + f, n = round(x), f in [-0.5, 0.5], minimax polynomial for
// 2^f on the FMA pipe (relative error 1.0e-4 at degree 3 - far below the bf16 rounding of P at 3.9e-3 - and 2.9e-6
// at degree 4 for the fp16 build), n added into the exponent field.  The SFUs retire 16 ex2 per clock and SM: a
// 128 x 128 score tile costs 1024 SFU cycles, exactly its two MMAs - moving part of the exponentials to the FMA pipe
// keeps the softmax off the critical path.  Inputs are <= 8 (lazy rescale) and clamped at -126.
SLIME_DEVINL float2 exp2_poly2(float2 x) {
  const float MAGIC = 12582912.0f;  // 1.5 * 2^23: x + MAGIC holds round(x) in its low mantissa bits
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 xr = fadd2(x, make_float2(MAGIC, MAGIC));
  const float2 n = fadd2(xr, make_float2(-MAGIC, -MAGIC));
  const float2 f = ffma2(n, make_float2(-1.0f, -1.0f), x);
#ifdef SLIME_FP16
  float2 q = ffma2(make_float2(0.009582850150763988f, 0.009582850150763988f), f,
                   make_float2(0.055906426161527634f, 0.055906426161527634f));
  q = ffma2(q, f, make_float2(0.24024099111557007f, 0.24024099111557007f));
  q = ffma2(q, f, make_float2(0.6931241750717163f, 0.6931241750717163f));
#else
  float2 q = ffma2(make_float2(0.05500892549753189f, 0.05500892549753189f), f,
                   make_float2(0.2422109693288803f, 0.2422109693288803f));
  q = ffma2(q, f, make_float2(0.6932829022407532f, 0.6932829022407532f));
#endif
  q = ffma2(q, f, make_float2(1.0f, 1.0f));
  float2 r;
  r.x = __int_as_float(__float_as_int(q.x) + (__float_as_int(xr.x) << 23));
  r.y = __int_as_float(__float_as_int(q.y) + (__float_as_int(xr.y) << 23));
  return r;
}

// Which of every 8 column pairs go through the polynomial: P = 0 none (all MUFU), 2 -> 2 of 8, 3 -> 3 of 8, 4 -> 4 of 8
SLIME_DEVINL constexpr bool pair_is_poly(int c, int P) {
  return P == 2 ? ((c & 3) == 1) : P == 3 ? ((c & 7) == 1 || (c & 7) == 4 || (c & 7) == 6) : P == 4 ? ((c & 1) == 1) : false;
}

SLIME_DEVINL int ld_acquire_cta(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];\n" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
SLIME_DEVINL void st_release_cta(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// max over 32 score columns held in registers; with `need_mask` columns beyond `limit` count as -inf
template <bool MASKED>
SLIME_DEVINL float rowmax32(const uint32_t (&sr)[32], int col_base, int limit) {
  float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains
  if (MASKED) {
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      float v = __uint_as_float(sr[c]);
      if (col_base + c > limit) v = -INFINITY;
      mx4[c & 3] = fmaxf(mx4[c & 3], v);
    }
  } else {
#pragma unroll
    for (int c = 0; c < 32; c += 2)  // FMNMX3: two values per instruction
      mx4[(c >> 1) & 3] = fmax3(mx4[(c >> 1) & 3], __uint_as_float(sr[c]), __uint_as_float(sr[c + 1]));
  }
  return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
}

// P = 2^(s * scale_log2 - m_scaled) for 32 columns -> 16 packed 16-bit pairs; returns the fp32 row sum of the 32 values.
// `plain` = the tile needs no masking (no -inf inputs), so the packed / polynomial arithmetic of variant P may be used.
template <int P, bool MASKED>
SLIME_DEVINL float softmax_exp32(const uint32_t (&sr)[32], int col_base, int limit, float scale_log2,
                                 float m_scaled, uint32_t (&pk)[16]) {
  if (!MASKED) {
    const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-m_scaled, -m_scaled);
    float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float2 x = ffma2(make_float2(__uint_as_float(sr[2 * c]), __uint_as_float(sr[2 * c + 1])), sc2, nm2);
      float2 pv;
      if (pair_is_poly(c, P)) {
        pv = exp2_poly2(x);
      } else {
        pv.x = fast_exp2(x.x);
        pv.y = fast_exp2(x.y);
      }
      acc[c & 1] = fadd2(acc[c & 1], pv);
      pk[c] = pack_bf16x2(pv.x, pv.y);
    }
    return (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y);
  }
  float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    float s0 = __uint_as_float(sr[2 * c]), s1 = __uint_as_float(sr[2 * c + 1]);
    if (col_base + 2 * c > limit) s0 = -INFINITY;
    if (col_base + 2 * c + 1 > limit) s1 = -INFINITY;
    const float p0 = fast_exp2(fmaf(s0, scale_log2, -m_scaled));
    const float p1 = fast_exp2(fmaf(s1, scale_log2, -m_scaled));
    ps4[c & 3] += p0 + p1;
    pk[c] = pack_bf16x2(p0, p1);
  }
  return (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
}

// One unit of work: up to two 128-row query tiles (slots) of one sequence that share a K/V stream of n_tiles tiles.
// Decoded once by the producer warp and handed to the other warps through an 8-entry ring in shared memory (`more`
// says whether another item follows; a CTA without any work posts one entry with n_tiles == 0).
struct alignas(64) Item2 {
  int n_tiles;      // kv tiles both slots run over
  int valid[2];     // slot X has a query tile
  int t[2];         // query tile index of slot X inside the sequence
  int head[2];      // q head of slot X
  int kv_head;
  int len_q, len_k, causal_off;
  int q_row0, k_row0;  // first row of this sequence in the q / kv matrices
  int more;            // another item follows in this CTA's list
  long long o_row0;
};
static_assert(sizeof(Item2) == 64, "ring slot");

template <bool CAUSAL>
SLIME_DEVINL Item2 decode_item2(const AttnParams& p, int w, int units, int hsel_count, int pair_heads) {
  Item2 it;
  int b, u, hsel;
  if (CAUSAL) {
    // heavy (late) query tiles first for load balance; consecutive w = the head pairs of one (sequence, tile): they
    // share K/V in L2
    hsel = w % hsel_count;
    const int rest = w / hsel_count;
    b = rest % p.batch;
    u = units - 1 - rest / p.batch;
  } else {
    u = w % units;
    const int rest = w / units;
    hsel = rest % hsel_count;
    b = rest / hsel_count;
  }
  if (pair_heads) {
    it.head[0] = 2 * hsel;
    it.head[1] = 2 * hsel + 1;
    it.t[0] = it.t[1] = u;
  } else {
    it.head[0] = it.head[1] = hsel;
    it.t[0] = 2 * u;
    it.t[1] = 2 * u + 1;
  }
  it.kv_head = it.head[0] / (p.num_heads / p.num_kv_heads);
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
  int n = 0;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int m0 = it.t[x] * BM;
    int nx = 0;
    if (m0 < it.len_q && it.len_k > 0) {
      int last = it.len_k;
      if (CAUSAL) last = min(it.len_k, m0 + BM + it.causal_off);
      nx = max(0, (last + BN - 1) / BN);
    }
    it.valid[x] = nx > 0;
    n = max(n, nx);
  }
  // both slots run over the same n kv tiles (a slot whose causal extent ends one tile earlier sees that tile fully
  // masked: P = 0) - the two MMA / softmax streams stay in lockstep
  it.n_tiles = n;
  return it;
}

template <int HD, bool CAUSAL, int PV>
__global__ void __launch_bounds__(NT, 1)
attn2q_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
              const __grid_constant__ CUtensorMap tmap_v, const AttnParams p, int units, int hsel_count, int pair_heads,
              int total_items, int per_cta) {
  using Cfg = Cfg2<HD>;
  constexpr int NK = Cfg::NK, NV = Cfg::NV;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                        // [2] slots
  uint8_t* sK = sQ + 2 * Cfg::TILE_BYTES;    // [NK]
  uint8_t* sV = sK + NK * Cfg::TILE_BYTES;   // [NV]
  uint8_t* aux = sV + NV * Cfg::TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(aux);
  uint64_t* q_full = bars + 0;    // [2] per slot; phase = the slot's item count
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;    // [2] per slot; phase = the slot's tile count
  uint64_t* p_ready = bars + 6;   // [2]
  uint64_t* o_done = bars + 8;    // [2] one commit per PV of the slot
  uint64_t* o_free = bars + 10;   // [2] epilogue has read O_X of the slot's previous item
  uint64_t* l_ready = bars + 12;  // [2] softmax has posted the item's row sums
  uint64_t* k_full = bars + 14;   // [NK <= 4]
  uint64_t* k_empty = bars + 18;  // [NK]
  uint64_t* v_full = bars + 22;   // [NV <= 4]
  uint64_t* v_empty = bars + 26;  // [NV]
  uint64_t* s_free = bars + 30;   // [2] EARLY_S: the slot's softmax has read S_X (it may be overwritten by the next score MMA)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 32);
  int* ring_head = reinterpret_cast<int*>(bars + 33);  // items posted so far
  float* lsum = reinterpret_cast<float*>(aux + Cfg::BAR_BYTES);                       // [2][128]
  uint8_t* stage_all = aux + Cfg::BAR_BYTES + Cfg::LSUM_BYTES;                        // [4 warps][32 rows][64 B]
  Item2* ring = reinterpret_cast<Item2*>(aux + Cfg::BAR_BYTES + Cfg::LSUM_BYTES + Cfg::STAGE_BYTES);  // [8] x 64 B

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
      mbar_init(&p_ready[s], 128);
      mbar_init(&o_done[s], 1);
      mbar_init(&o_free[s], 128);
      mbar_init(&l_ready[s], 128);
      mbar_init(&s_free[s], 128);
    }
    for (int s = 0; s < NK; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
    }
    for (int s = 0; s < NV; ++s) {
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
    }
    *ring_head = 0;
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
    // ================================ TMA producer ================================
    if (lane == 0) {
      // work list of this CTA: causal -> w = blockIdx + k * grid (heavy tiles first, interleaved for balance);
      // non-causal -> a CONTIGUOUS block of the list, so the query-tile pairs of one (sequence, head) run back to back on
      // the same SM and their common K/V is fetched from DRAM once (a second CTA asking for the same lines at the same
      // moment does not hit in L2: profiles/r01_ncu_attention_vit.txt, 3.7 x the algorithmic DRAM reads)
      int w = CAUSAL ? static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x) * per_cta;
      const int w_end = CAUSAL ? total_items : min(total_items, w + per_cta);
      const int w_step = CAUSAL ? static_cast<int>(gridDim.x) : 1;
      auto next_valid = [&]() {
        Item2 it;
        while (w < w_end) {
          it = decode_item2<CAUSAL>(p, w, units, hsel_count, pair_heads);
          w += w_step;
          if (it.n_tiles > 0) return it;
        }
        it.n_tiles = 0;
        it.valid[0] = it.valid[1] = 0;
        it.more = 0;
        return it;
      };
      int iq[2] = {0, 0};  // items in which slot X was used so far
      int g = 0;           // kv tiles loaded so far (ring index)
      Item2 cur = next_valid();
      for (int k = 0;; ++k) {
        Item2 nxt;
        nxt.n_tiles = 0;
        if (cur.n_tiles > 0) {
          nxt = next_valid();
          cur.more = nxt.n_tiles > 0;
        }
        ring[k & 7] = cur;
        st_release_cta(ring_head, k + 1);
        if (cur.n_tiles == 0) break;
        const int n = cur.n_tiles;
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
        // K(0) first: its ring stage is free long before the slots' Q buffers are (they are released by the LAST score
        // MMA of the previous item), so only the Q load sits between that MMA and the first score MMA of this item
        load_k(0);
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (!cur.valid[x]) continue;
          mbar_wait(&q_empty[x], (iq[x] & 1) ^ 1);
          mbar_arrive_expect_tx(&q_full[x], Cfg::TILE_BYTES);
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s)
            tma_load_2d(sQ + x * Cfg::TILE_BYTES + s * SLAB_BYTES, &tmap_q, &q_full[x], cur.head[x] * HD + s * 64,
                        cur.q_row0 + cur.t[x] * BM);
          ++iq[x];
        }
        constexpr int KA = NK - 1;  // K tiles in flight ahead of V
        for (int j = 1; j < KA && j < n; ++j) load_k(j);
        for (int j = 0; j < n; ++j) {
          if (j + KA < n) load_k(j + KA);
          load_v(j);
        }
        g += n;
        cur = nxt;
      }
    }
  } else if (warp_idx == 1) {
    // ================================ MMA issuer ==================================
    // The WHOLE warp runs this loop converged and one elected lane issues the tcgen05 instructions.  Every value the
    // descriptors depend on is warp-uniform and known to be so (ring fields go through __shfl_sync), so they live in
    // uniform registers and consecutive MMAs are a couple of uniform adds apart.  (Issued from inside `if (lane == 0)`
    // the operands sit in per-thread registers and every tcgen05.mma pays ~14 instructions of R2UR / ELECT / branch:
    // ~95 cycles per 64-cycle MMA in the first version of this kernel - the issue thread, not the tensor pipe, set the
    // pace.)
    {
      constexpr uint32_t idesc_s = make_idesc_bf16_major(BM, BN, 0, 0);   // Q, K both K-major
      constexpr uint32_t idesc_pv = make_idesc_bf16_major(BM, HD, 0, 1);  // P from TMEM, V MN-major
      auto uni = [](int v) { return __shfl_sync(0xffffffffu, v, 0); };
      const uint32_t sQ_u = static_cast<uint32_t>(uni(static_cast<int>(smem_u32(sQ))));
      const uint32_t sK_u = sQ_u + 2 * Cfg::TILE_BYTES;
      const uint32_t sV_u = sK_u + NK * Cfg::TILE_BYTES;
      const uint32_t tmem_u = static_cast<uint32_t>(uni(static_cast<int>(tmem_base)));
      int gs[2] = {0, 0};  // tiles of slot X so far  (s_full / p_ready / o_done phases)
      int is[2] = {0, 0};  // items of slot X so far  (q_full / o_free phases)
      int g = 0;           // kv ring index of tile 0 of the current item

      // S_X(j) of the item whose tile 0 has ring index g0; `last_user`: no later score MMA reads this K tile
      auto issue_s = [&](int x, int g0, int j, int n, bool last_user, bool k_waited = false) {
        const int gi = g0 + j, st = gi % NK;
        if (!k_waited) mbar_wait(&k_full[st], (gi / NK) & 1);
        tcgen05_fence_after();
        const uint64_t dq = make_umma_desc_sw128(sQ_u + x * Cfg::TILE_BYTES);
        const uint64_t dk = make_umma_desc_sw128(sK_u + st * Cfg::TILE_BYTES);
        const uint32_t tmem_s = tmem_u + Cfg::S_COL + x * Cfg::SLOT_STRIDE;
        if (elect_one_sync()) {
#pragma unroll
          for (int s = 0; s < Cfg::SLABS; ++s) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              umma_bf16_ss(tmem_s, dq + (s * (SLAB_BYTES >> 4) + 2 * kk), dk + (s * (SLAB_BYTES >> 4) + 2 * kk), idesc_s,
                           (s | kk) != 0 ? 1u : 0u);
          }
          if (last_user) umma_commit(&k_empty[st]);
          umma_commit(&s_full[x]);
          if (j == n - 1) umma_commit(&q_empty[x]);  // the slot's Q buffer may be reloaded
        }
        __syncwarp();
      };
      // O_X += P_X(j) V_j; P_X(j) and V(j) have been waited for
      auto issue_pv = [&](int x, int vs, bool first, bool last_user) {
        const uint32_t tmem_p = tmem_u + (Cfg::EARLY_S ? Cfg::P_COL : Cfg::S_COL) + x * Cfg::SLOT_STRIDE;
        const uint32_t tmem_o = tmem_u + Cfg::O_COL + x * Cfg::SLOT_STRIDE;
        const uint64_t dv = make_umma_desc_mn_sw128(sV_u + vs * Cfg::TILE_BYTES, SLAB_BYTES);
        if (elect_one_sync()) {
#pragma unroll
          for (int kk = 0; kk < BN / 16; ++kk) {
            // A: 16 kv positions = 8 TMEM columns of packed 16-bit pairs;  B: 16 kv rows = 2048 bytes further down
            umma_bf16_ts(tmem_o, tmem_p + kk * 8, dv + static_cast<uint64_t>(kk * (2048 >> 4)), idesc_pv,
                         (!first || kk != 0) ? 1u : 0u);
          }
          if (last_user) umma_commit(&v_empty[vs]);
          umma_commit(&o_done[x]);
        }
        __syncwarp();
      };

      while (ld_acquire_cta(ring_head) <= 0) {
      }
      int n = uni(ring[0].n_tiles), more = uni(ring[0].more);
      int valid[2] = {uni(ring[0].valid[0]), uni(ring[0].valid[1])};
      bool started[2] = {false, false};  // S_X(0) of the current item already issued
      for (int k = 0; n > 0; ++k) {
        const int last_slot = valid[1] ? 1 : 0;
        // first score tile of every slot that did not get it at the end of the previous item
#pragma unroll
        for (int x = 0; x < 2; ++x) {
          if (valid[x] && !started[x]) {
            mbar_wait(&q_full[x], is[x] & 1);
            issue_s(x, g, 0, n, x == last_slot);
          }
          started[x] = false;
        }
        int nx_n = 0, nx_more = 0, nx_valid[2] = {0, 0};
        bool have_next = false;
        // the slot's next score tile: same item, or tile 0 of the next item (one continuous stream)
        auto next_s = [&](int x, int j) {
          if (j + 1 < n) {
            issue_s(x, g, j + 1, n, x == last_slot, !Cfg::EARLY_S);
          } else {
            ++is[x];
            if (more) {
              if (!have_next) {
                while (ld_acquire_cta(ring_head) <= k + 1) {
                }
                const Item2* pn = &ring[(k + 1) & 7];
                nx_n = uni(pn->n_tiles);
                nx_more = uni(pn->more);
                nx_valid[0] = uni(pn->valid[0]);
                nx_valid[1] = uni(pn->valid[1]);
                have_next = true;
              }
              if (nx_valid[x]) {
                mbar_wait(&q_full[x], is[x] & 1);
                issue_s(x, g + n, 0, nx_n, x == (nx_valid[1] ? 1 : 0));
                started[x] = true;
              }
            }
          }
        };
        for (int j = 0; j < n; ++j) {
          const int gv = g + j, vs = gv % NV;
          if constexpr (Cfg::EARLY_S) {
            // S_X(j+1) as soon as the slot's softmax has READ S_X(j) - it is still exponentiating
#pragma unroll
            for (int x = 0; x < 2; ++x) {
              if (!valid[x]) continue;
              mbar_wait(&s_free[x], (gs[x] + 0) & 1);
              next_s(x, j);
            }
          }
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            if (!valid[x]) continue;
            // everything this step needs besides P is waited for FIRST (it arrived long ago, but every mbarrier wait
            // costs ~100 cycles of latency): once P_X(j) lands, PV_X(j) (and S_X(j+1)) go out back to back
            const int isx = Cfg::EARLY_S ? (j + 1 < n ? is[x] : is[x] - 1) : is[x];  // items before this one
            if (j == 0) mbar_wait(&o_free[x], (isx & 1) ^ 1);  // epilogue of the slot's previous item has read O_X
            mbar_wait(&v_full[vs], (gv / NV) & 1);
            if (!Cfg::EARLY_S && j + 1 < n) mbar_wait(&k_full[(gv + 1) % NK], ((gv + 1) / NK) & 1);
            const bool tr = p.trace != nullptr && blockIdx.x == 0 && x == 0 && gs[0] < 64 && lane == 0;
            if (tr) p.trace[gs[0] * 16 + 8] = clock64();
            mbar_wait(&p_ready[x], gs[x] & 1);
            if (tr) p.trace[gs[0] * 16 + 9] = clock64();
            tcgen05_fence_after();
            issue_pv(x, vs, j == 0, x == last_slot);
            if (tr) p.trace[gs[0] * 16 + 11] = clock64();
            ++gs[x];
            if constexpr (!Cfg::EARLY_S) next_s(x, j);  // P_X aliases S_X: the next score tile follows PV_X(j)
          }
        }
        if (!more) break;
        // (have_next is always set here: a valid item has at least one valid slot)
        g += n;
        n = nx_n;
        more = nx_more;
        valid[0] = nx_valid[0];
        valid[1] = nx_valid[1];
      }
    }
  } else if (warp_idx < 10) {
    // ================================ softmax of slot X (4 warps, one thread per query row) ================
    const int X = (warp_idx - 2) >> 2;
    const int quad = warp_idx & 3;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float scale_log2 = p.scale * 1.4426950408889634f;
    const uint32_t s_base = tmem_base + lane_addr + Cfg::S_COL + X * Cfg::SLOT_STRIDE;
    const uint32_t p_base = tmem_base + lane_addr + Cfg::P_COL + X * Cfg::SLOT_STRIDE;  // == s_base unless EARLY_S
    const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + X * Cfg::SLOT_STRIDE;
    int gx = 0;  // tiles of this slot so far
    int ix = 0;  // items of this slot so far
    for (int k = 0;; ++k) {
      if (lane == 0) {
        while (ld_acquire_cta(ring_head) <= k) {
        }
      }
      __syncwarp();
      const Item2* pi = &ring[k & 7];
      const int n_tiles = pi->n_tiles;
      if (n_tiles == 0) break;
      const int more = pi->more;
      if (pi->valid[X]) {
        const int it_t = pi->t[X], len_k = pi->len_k, causal_off = pi->causal_off;
        const int row = it_t * BM + r_in_tile;  // query index inside the sequence
        const int limit = CAUSAL ? min(len_k - 1, row + causal_off) : len_k - 1;  // last visible column of this row
        float m_cur = -INFINITY;  // raw-score max the exponentials are taken against
        float l_sum = 0.f;
        for (int j = 0; j < n_tiles; ++j, ++gx) {
          const bool tr = p.trace != nullptr && blockIdx.x == 0 && X == 0 && warp_idx == 2 && lane == 0 && gx < 64;
          if (tr) p.trace[gx * 16 + 0] = clock64();
          mbar_wait(&s_full[X], gx & 1);
          tcgen05_fence_after();
          if (tr) p.trace[gx * 16 + 1] = clock64();
          // Masking by 32-column chunk, decided per WARP: chunks [0, n_full) are visible to every row of the warp (packed
          // arithmetic), chunks [n_full, n_any) are cut by the causal diagonal / the end of the keys for some row
          // (per-element compare), chunks [n_any, 4) are invisible to all 32 rows: skipped, P = 0.
          int n_full = 4, n_any = 4;
          if ((j * BN + BN > len_k) || (CAUSAL && (j * BN + BN - 1 > it_t * BM + causal_off))) {
            const int vis = limit - j * BN + 1;  // visible columns of this tile for this row (may be <= 0 or >= 128)
            n_full = __reduce_min_sync(0xffffffffu, max(0, min(4, vis >> 5)));
            n_any = __reduce_max_sync(0xffffffffu, max(0, min(4, (vis + 31) >> 5)));
          }
          // ---- pass 1: row max, 32 columns at a time; the load of chunk c+1 is in flight while chunk c is reduced.
          //      (`mt` = this tile needs masking at all: a compile-time split, so the common unmasked tile carries no
          //      per-chunk tests - as a run-time flag the compiler predicated BOTH variants into one instruction stream.)
          const bool masked_tile = n_full < 4;  // warp-uniform
          auto pass1 = [&](auto mt) {
            constexpr bool MT = decltype(mt)::value;
            float m = -INFINITY;
            uint32_t sr[2][32];
            tmem_ld_32x32b_x32(s_base, sr[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              tmem_ld_wait();
              if (c + 1 < 4) tmem_ld_32x32b_x32(s_base + (c + 1) * 32, sr[(c + 1) & 1]);
              if (!MT) {
                m = fmaxf(m, rowmax32<false>(sr[c & 1], 0, 0));
              } else if (c < n_any) {
                m = fmaxf(m, c < n_full ? rowmax32<false>(sr[c & 1], 0, 0) : rowmax32<true>(sr[c & 1], j * BN + c * 32, limit));
              }
            }
            return m;
          };
          const float m_tile = masked_tile ? pass1(std::true_type{}) : pass1(std::false_type{});
          if (tr) p.trace[gx * 16 + 2] = clock64();
          // ---- lazy rescale: only move the reference max when it grew by more than 2^8 (always on the first tile)
          bool grow = (m_tile - m_cur) * scale_log2 > RESCALE_THRESHOLD;
          if (m_tile == -INFINITY) grow = false;
          if (j > 0 && __any_sync(0xffffffffu, grow)) {
            const float alpha = grow ? fast_exp2((m_cur - m_tile) * scale_log2) : 1.0f;  // m_cur == -inf -> 0
            l_sum *= alpha;
            mbar_wait(&o_done[X], (gx - 1) & 1);  // PV_X(j-1) finished: O_X is stable until PV_X(j)
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
              uint32_t orow[32];
              tmem_ld_32x32b_x32(o_addr + c * 32, orow);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) orow[i] = __float_as_uint(__uint_as_float(orow[i]) * alpha);
              tmem_st_32x32b_x32(o_addr + c * 32, orow);
            }
            tmem_st_wait();
          }
          if (grow) m_cur = m_tile;
          const float m_scaled = (m_cur == -INFINITY) ? 0.f : m_cur * scale_log2;
          // ---- pass 2: exponentials, 32 columns at a time (next chunk's load in flight); P (16-bit pairs): chunk c ->
          //      columns [16 c, 16 c + 16) of P_X.  Without EARLY_S P_X is the head of S_X (columns already consumed);
          //      with it P_X is a buffer of its own: it must have been read by PV_X(j-1), and S_X is handed back to the
          //      tensor pipe as soon as its last chunk is in registers (half way through the exponentials).
          if (Cfg::EARLY_S && gx > 0) mbar_wait(&o_done[X], (gx - 1) & 1);
          auto pass2 = [&](auto mt) {
            constexpr bool MT = decltype(mt)::value;
            float l = 0.f;
            uint32_t sr[2][32];
            tmem_ld_32x32b_x32(s_base, sr[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint32_t pk[16];
              if (!(Cfg::EARLY_S && c == 3)) tmem_ld_wait();
              if (c + 1 < 4) tmem_ld_32x32b_x32(s_base + (c + 1) * 32, sr[(c + 1) & 1]);
              if (Cfg::EARLY_S && c == 2) {
                tmem_ld_wait();  // chunk 3 is in registers too: S_X(j) is no longer needed
                tcgen05_fence_before();
                mbar_arrive(&s_free[X]);
              }
              if (!MT) {
                l += softmax_exp32<PV, false>(sr[c & 1], 0, 0, scale_log2, m_scaled, pk);
              } else if (c < n_full) {
                l += softmax_exp32<PV, false>(sr[c & 1], 0, 0, scale_log2, m_scaled, pk);
              } else if (c < n_any) {
                l += softmax_exp32<PV, true>(sr[c & 1], j * BN + c * 32, limit, scale_log2, m_scaled, pk);
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = 0u;
              }
              tmem_st_32x32b_x16(p_base + c * 16, pk);
            }
            return l;
          };
          l_sum += masked_tile ? pass2(std::true_type{}) : pass2(std::false_type{});
          if (tr) p.trace[gx * 16 + 3] = clock64();
          tmem_st_wait();
          tcgen05_fence_before();
          mbar_arrive(&p_ready[X]);
          if (tr) p.trace[gx * 16 + 4] = clock64();
        }
        // ---- item end: hand the row sums to the epilogue warps.  The previous item's sums must have been consumed
        //      (the epilogue arrives on o_free after reading them) - a one-tile item can finish before that.
        if (ix > 0) mbar_wait(&o_free[X], (ix - 1) & 1);
        lsum[X * 128 + r_in_tile] = l_sum;
        mbar_arrive(&l_ready[X]);
        ++ix;
      }
      if (!more) break;
    }
  } else {
    // ================================ epilogue of both slots (4 warps) ================================
    const int quad = warp_idx & 3;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    uint8_t* stage = stage_all + (warp_idx - 10) * (32 * 64);
    int ge[2] = {0, 0};  // tiles of slot X through the end of the current item
    int ie[2] = {0, 0};  // items of slot X so far
    for (int k = 0;; ++k) {
      if (lane == 0) {
        while (ld_acquire_cta(ring_head) <= k) {
        }
      }
      __syncwarp();
      const Item2* pi = &ring[k & 7];
      const int n_tiles = pi->n_tiles;
      if (n_tiles == 0) break;
      const int more = pi->more;
      const int len_q = pi->len_q;
      const long long o_row0 = pi->o_row0;
      int vx[2], tx[2], hx[2];
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        vx[x] = pi->valid[x];
        tx[x] = pi->t[x];
        hx[x] = pi->head[x];
      }
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (!vx[x]) continue;
        ge[x] += n_tiles;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && x == 0 && warp_idx == 10 && lane == 0 && k < 64;
        if (tr) p.trace[k * 16 + 12] = clock64();
        mbar_wait(&l_ready[x], ie[x] & 1);
        const float l_tot = lsum[x * 128 + r_in_tile];
        const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
        mbar_wait(&o_done[x], (ge[x] - 1) & 1);  // the slot's last PV of this item
        tcgen05_fence_after();
        if (tr) p.trace[k * 16 + 13] = clock64();
        const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + x * Cfg::SLOT_STRIDE;
        uint32_t opk[HD / 2];  // the whole row, 16-bit pairs
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t orow[32];
          tmem_ld_32x32b_x32(o_addr + c * 32, orow);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i)
            opk[c * 16 + i] = pack_bf16x2(__uint_as_float(orow[2 * i]) * inv_l, __uint_as_float(orow[2 * i + 1]) * inv_l);
        }
        tcgen05_fence_before();
        mbar_arrive(&o_free[x]);  // O_X (and lsum[x]) may be overwritten by the slot's next item
        ++ie[x];
        if (tr) p.trace[k * 16 + 14] = clock64();
        // transposed through this warp's 2 KB of shared memory (16-byte slots XOR-swizzled by row pair: conflict-free
        // both ways): every store instruction writes 8 rows x 64 contiguous bytes
        const int rows_valid = len_q - tx[x] * BM;
        bf16* out = p.o + (o_row0 + tx[x] * BM + quad * 32) * p.o_ld + hx[x] * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = opk[c * 16 + q * 4 + 0];
            v.y = opk[c * 16 + q * 4 + 1];
            v.z = opk[c * 16 + q * 4 + 2];
            v.w = opk[c * 16 + q * 4 + 3];
            *reinterpret_cast<uint4*>(stage + lane * 64 + ((q ^ ((lane >> 1) & 3)) << 4)) = v;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2), q = lane & 3;  // 8 rows x 4 slots per instruction
            const uint4 v = *reinterpret_cast<const uint4*>(stage + r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
            if (quad * 32 + r < rows_valid)
              *reinterpret_cast<uint4*>(out + static_cast<size_t>(r) * p.o_ld + c * 32 + q * 8) = v;
          }
          __syncwarp();
        }
        if (tr) p.trace[k * 16 + 15] = clock64();
      }
      if (!more) break;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

template <int HD, bool CAUSAL, int PV>
int launch2q_var(const AttnParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = Cfg2<HD>;
  auto kern = attn2q_kernel<HD, CAUSAL, PV>;
  static bool attr_set = false;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  // Tensor maps over the packed row matrices; the kernel addresses heads by column offset.
  const long long q_rows = p.total_q_rows > 0 ? p.total_q_rows
                           : (p.q_batch_rows > 0 ? p.q_batch_rows * p.batch : p.seqlen_q);
  const long long k_rows = p.total_k_rows > 0 ? p.total_k_rows
                           : (p.k_batch_rows > 0 ? p.k_batch_rows * p.batch : p.seqlen_k);
  CUtensorMap tq, tk, tv;
  SLIME_PROPAGATE(slime_get_tmap(p.q, static_cast<int>(q_rows), p.num_heads * HD, p.q_ld, BM, &tq));
  SLIME_PROPAGATE(slime_get_tmap(p.k, static_cast<int>(k_rows), p.num_kv_heads * HD, p.k_ld, BN, &tk));
  SLIME_PROPAGATE(slime_get_tmap(p.v, static_cast<int>(k_rows), p.num_kv_heads * HD, p.v_ld, BN, &tv));
  const int q_tiles = (p.seqlen_q + BM - 1) / BM;
  // slots = two q heads of one GQA group at the same query tile (group size even), else two consecutive query tiles
  const int pair_heads = ((p.num_heads / p.num_kv_heads) % 2 == 0) ? 1 : 0;
  const int units = pair_heads ? q_tiles : (q_tiles + 1) / 2;
  const int hsel_count = pair_heads ? p.num_heads / 2 : p.num_heads;
  const int total = units * hsel_count * p.batch;
  const int grid = total < num_sms ? total : num_sms;
  const int per_cta = (total + grid - 1) / grid;
  double flops = 0.0;
  if (p.cu_q == nullptr)
    flops = 4.0 * p.seqlen_q * static_cast<double>(p.seqlen_k) * HD * p.num_heads * p.batch * (CAUSAL ? 0.5 : 1.0);
  slime_prof_begin(1, flops, stream);
  const cudaError_t le = slime_launch_prefill(kern, dim3(grid), dim3(NT), Cfg::SMEM_BYTES, stream, tq, tk, tv, p, units,
                                              hsel_count, pair_heads, total, per_cta);
  slime_prof_end(stream);
  SLIME_CHECK_CUDA(le);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int g_poly = -2;  // pairs of every 8 exponentiated on the FMA pipe (-2: not read yet; -1: by head dim; SLIME_ATTN_POLY)

template <int HD, bool CAUSAL>
int launch2q(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (g_poly == -2) {
    const char* e = getenv("SLIME_ATTN_POLY");
    g_poly = e != nullptr ? atoi(e) : -1;
  }
  // measured best share (profiles/r02_attention_experiments.txt): 3 of 8 at hd 128 (decoder 0.370 / 0.317 / 0.298 / 0.342 ms
  // for 0 / 2 / 3 / 4), 2 of 8 at hd 64 (ViT 0.291 / 0.260 / 0.269 / 0.288 ms)
  const int poly = g_poly >= 0 ? g_poly : (HD == 128 ? SLIME_ATTN_POLY_HD128 : SLIME_ATTN_POLY_HD64);
  switch (poly) {
    case 0: return launch2q_var<HD, CAUSAL, 0>(p, num_sms, stream);
    case 2: return launch2q_var<HD, CAUSAL, 2>(p, num_sms, stream);
    case 3: return launch2q_var<HD, CAUSAL, 3>(p, num_sms, stream);
    case 4: return launch2q_var<HD, CAUSAL, 4>(p, num_sms, stream);
    default:
      slime_set_error("attention: unknown polynomial share %d (0, 2, 3 or 4 of every 8 pairs)", poly);
      return SLIME_EINVAL;
  }
}

}  // namespace

extern "C" int slime_attention_set_poly(int pairs_of_8) {
  const bool known = pairs_of_8 == -1 || pairs_of_8 == 0 || pairs_of_8 == 2 || pairs_of_8 == 3 || pairs_of_8 == 4;
  if (!known) {
    slime_set_error("attention: polynomial share must be 0, 2, 3 or 4 pairs of every 8 (-1 = default), got %d", pairs_of_8);
    return SLIME_EINVAL;
  }
  g_poly = pairs_of_8;
  return SLIME_OK;
}

int slime_launch_attention_tc2(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (p.batch <= 0 || p.seqlen_q <= 0 || p.seqlen_k <= 0) return SLIME_OK;
  if (p.head_dim == 64) {
    return p.causal ? launch2q<64, true>(p, num_sms, stream) : launch2q<64, false>(p, num_sms, stream);
  }
  return p.causal ? launch2q<128, true>(p, num_sms, stream) : launch2q<128, false>(p, num_sms, stream);
}
