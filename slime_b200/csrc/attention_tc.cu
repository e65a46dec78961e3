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
#include <cstdlib>

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
  // exchange area between the two softmax warps of a row.  column split: half-row max (2 slots x 2 halves) and item
  // sums (2 halves) = 6 x 128 floats;  kv split: the reference max handed from tile to tile (128 floats), the
  // per-item (sum, max) of both groups for 2 items in flight (2 x 2 x 128 float2) and the ring head counter
  static constexpr int XCH_BYTES = 128 * 4 + 2 * 2 * 128 * 8 + 64;
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

// 2^x for a pair WITHOUT the SFU: Cody-Waite split x = n + f, n = round(x), f in [-0.5, 0.5], minimax polynomial
// for 2^f on the FMA pipe (relative error 1.0e-4 at degree 3 - far below the bf16 rounding of P at 3.9e-3 - and
// 2.9e-6 at degree 4 for the fp16 build), n added into the exponent field.  The SFUs retire 16 ex2 per clock per
// SM, which makes a 128 x 128 score tile cost as much SFU time as tensor time (profiles/r01_attention_clock_trace_v2.txt);
// moving a fraction of the exponentials onto the FMA pipe shortens that phase.  Inputs are <= 8 (lazy rescale)
// and clamped at -126 so the exponent arithmetic cannot wrap.
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

// Softmax arithmetic variant of attn_tc_kernel (template parameter VAR; SLIME_ATTN_VARIANT / slime_attention_set_variant):
//   0        : scalar FFMA + MUFU.EX2 per element
//   1 + 2*P  : packed pairs (FFMA2 scale, FADD2 row sums) with P of every 8 pairs exponentiated by exp2_poly2 on
//              the FMA pipe instead of the SFU (P = 2 -> variant 5, the default; P = 4 -> 9); tiles that need
//              masking keep the scalar path (masked scores are -inf, which the polynomial path would turn into
//              2^-126 instead of 0)
//   16 + v   : kv-split softmax (template parameter SPLIT) with arithmetic v; built: 21
//   32       : MEASUREMENT ONLY, garbage output: no softmax at all (floor of the TMA + tcgen05.mma side)
// Measured on B200 (profiles/r01_attention_experiments.txt): decoder shape 0.376 ms (0) -> 0.367 (5); kv-split 0.386.
SLIME_DEVINL constexpr bool pair_is_poly(int c, int P) {
  // spread the polynomial pairs evenly over each group of 8 so SFU and FMA work interleave
  return P == 2 ? ((c & 3) == 1) : P == 3 ? ((c & 7) == 1 || (c & 7) == 4 || (c & 7) == 6) : P == 4 ? ((c & 1) == 1) : false;
}

// ---- named barriers (id 0 is __syncthreads) ----
SLIME_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
SLIME_DEVINL void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(nthreads) : "memory");
}
SLIME_DEVINL int ld_acquire_cta(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared::cta.b32 %0, [%1];\n" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
SLIME_DEVINL void st_release_cta(int* p, int v) {
  asm volatile("st.release.cta.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// max over 64 score columns held in registers; with `need_mask` columns beyond `limit` are set to -inf in place
SLIME_DEVINL float rowmax64(uint32_t (&sr)[64], bool need_mask, int col_base, int limit) {
  float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // independent chains
  if (need_mask) {
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
  return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
}

// P = 2^(s * scale_log2 - m_scaled) for 64 columns -> 32 packed bf16 pairs; returns the fp32 row sum of the 64 values.
// `packed_ok` = the tile needs no masking (no -inf inputs), so the packed / polynomial arithmetic of VAR may be used.
template <int VAR>
SLIME_DEVINL float softmax_exp64(const uint32_t (&sr)[64], bool packed_ok, float scale_log2, float m_scaled,
                                 uint32_t (&pk)[32]) {
  if (VAR != 0 && packed_ok) {
    constexpr int P = VAR >> 1;
    const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-m_scaled, -m_scaled);
    float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
    for (int c = 0; c < 32; ++c) {
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
  for (int c = 0; c < 32; ++c) {
    const float p0 = fast_exp2(fmaf(__uint_as_float(sr[2 * c]), scale_log2, -m_scaled));
    const float p1 = fast_exp2(fmaf(__uint_as_float(sr[2 * c + 1]), scale_log2, -m_scaled));
    ps4[c & 3] += p0 + p1;
    pk[c] = pack_bf16x2(p0, p1);
  }
  return (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
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
  if (o_done_bar != nullptr) mbar_wait(o_done_bar, o_done_parity);  // nullptr: the caller has already waited
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

template <int HD, bool CAUSAL, int VAR, bool SPLIT>
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
  int* ring_head = reinterpret_cast<int*>(aux + Cfg::BAR_BYTES + Cfg::XCH_BYTES - 64);  // items posted so far

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
      mbar_init(&p_ready[s], SPLIT ? 128 : 256);  // kv split: only the 4 warps of the tile's group arrive
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
    *ring_head = 0;
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
        st_release_cta(ring_head, k + 1);  // kv-split softmax warps learn about item k from this counter
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
  } else if constexpr (SPLIT) {
    // ================================ softmax + epilogue, kv-split (2 groups x 4 warps) ================
    // Group G = warps {2..5} / {6..9} handles the kv tiles with global index gj & 1 == G - i.e. S / P buffer G -
    // with ONE thread per query row (TMEM lane r belongs to warp quadrant r / 32 of both groups).  While one group
    // exponentiates tile gj the other finds the row max of tile gj+1, and the 1024 tensor-pipe cycles a group waits
    // for its next S (PV of its last tile, then the S MMA into the same buffer) are filled by the other group's
    // softmax - the column-split version leaves SFU and FMA pipes idle in exactly those phases
    // (profiles/r01_attention_clock_trace_v2.txt: ~900 of 2150 cycles per tile).
    //   * the reference max of a row travels from tile to tile through shared memory (mail[r]) with a named-barrier
    //     hand-off per tile (arrive by the publisher, sync by the next tile's warp): tile gj's warp decides whether
    //     the max grows (lazy rescale, as before) after it has seen the decision of tile gj-1;
    //   * each group keeps its own partial row sum, relative to the last reference max it has seen; both post
    //     (sum, max) per item and combine them in the deferred epilogue;
    //   * O is rescaled (rare) by the deciding thread over all HD columns after PV(gj-1) has completed.
    const int quad = warp_idx & 3;
    const int G = (warp_idx - 2) >> 2;
    const int r_in_tile = quad * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const float scale_log2 = p.scale * 1.4426950408889634f;
    const int pair_bar = 1 + quad;            // both warps of a row, blocking (deferred epilogue)
    const int hand_in = 5 + (1 - G) * 4 + quad;   // published by the other group
    const int hand_out = 5 + G * 4 + quad;        // published by this group
    float* mail = xch;                                             // [128] reference max after the latest decision
    float2* post = reinterpret_cast<float2*>(xch + 128);           // [2 items][2 groups][128] (sum, max)
    const uint32_t s_base = tmem_base + lane_addr + (G ? Cfg::S_COL1 : Cfg::S_COL0);
    int g = 0;
    int pend_k = -1, pend_g_last = -1;
    // The last PV of the pending item (global tile pend_g_last) commits o_done[pend_g_last & 1]; the NEXT commit on that
    // barrier is PV(pend_g_last + 2).  A parity wait must not fall two phases behind, so each group waits for the
    // pending item's O BEFORE it releases its own first tile of the following item (PV(pend_g_last + 2) needs that
    // tile's P, or - in-order issue - the P of pend_g_last + 1), never afterwards.
    auto wait_pending_o = [&]() { mbar_wait(&o_done[pend_g_last & 1], (pend_g_last >> 1) & 1); };
    auto run_epilogue = [&](long long* tr) {
      // both groups have posted their (sum, max) of item pend_k once both are here
      named_bar_sync(pair_bar, 64);
      const float2 a = post[((pend_k & 1) * 2 + 0) * 128 + r_in_tile];
      const float2 b = post[((pend_k & 1) * 2 + 1) * 128 + r_in_tile];
      const float m_fin = fmaxf(a.y, b.y);
      float l_tot = 0.f;
      if (a.y != -INFINITY) l_tot += a.x * fast_exp2((a.y - m_fin) * scale_log2);
      if (b.y != -INFINITY) l_tot += b.x * fast_exp2((b.y - m_fin) * scale_log2);
      const float inv_l = l_tot > 0.f ? 1.0f / l_tot : 0.f;
      const Item* pi = &ring[pend_k & 7];
      const int obuf = pend_k & 1;
      epilogue_tile<HD>(nullptr, 0, &o_free[obuf],
                        tmem_base + lane_addr + Cfg::O_COL + obuf * Cfg::O_STRIDE + G * (HD / 2),
                        stage_all + (warp_idx - 2) * (32 * 64), lane, quad, inv_l, pi->len_q - pi->t * BM,
                        p.o + (pi->o_row0 + pi->t * BM + quad * 32) * p.o_ld + pi->head * HD + G * (HD / 2), p.o_ld, tr);
      pend_k = -1;
    };
    for (int k = 0;; ++k) {
      while (ld_acquire_cta(ring_head) <= k) {
      }
      const int n_tiles = ring[k & 7].n_tiles;
      if (n_tiles == 0) break;
      const int more = ring[k & 7].more;
      const int it_t = ring[k & 7].t, len_k = ring[k & 7].len_k, causal_off = ring[k & 7].causal_off;
      const int row = it_t * BM + r_in_tile;  // query index inside the sequence
      const int limit = CAUSAL ? min(len_k - 1, row + causal_off) : len_k - 1;  // last visible column of this row
      float m_cur = -INFINITY;  // reference max this thread's partial sum is relative to
      float l_part = 0.f;       // sum over the tiles of THIS group
      bool first_done = false;
      for (int j = (g & 1) == G ? 0 : 1; j < n_tiles; j += 2) {
        const int gj = g + j;
        const bool tr = p.trace != nullptr && blockIdx.x == 0 && warp_idx == 2 + 4 * G && lane == 0 && gj < 64;
        if (tr) p.trace[gj * 16 + 0] = clock64();
        mbar_wait(&s_full[G], (gj >> 1) & 1);
        tcgen05_fence_after();
        if (tr) p.trace[gj * 16 + 1] = clock64();
        const bool need_mask = (j * BN + BN > len_k) || (CAUSAL && (j * BN + BN - 1 > it_t * BM + causal_off));
        // ---- pass 1: row max over the 128 columns
        float m_tile;
        {
          uint32_t sr[64];
          tmem_ld_32x32b_x32(s_base, sr);
          tmem_ld_32x32b_x32(s_base + 32, sr + 32);
          tmem_ld_wait();
          if (tr) p.trace[gj * 16 + 2] = clock64();
          m_tile = rowmax64(sr, need_mask, j * BN, limit);
          tmem_ld_32x32b_x32(s_base + 64, sr);
          tmem_ld_32x32b_x32(s_base + 96, sr + 32);
          tmem_ld_wait();
          m_tile = fmaxf(m_tile, rowmax64(sr, need_mask, j * BN + 64, limit));
        }
        // ---- the decision of the previous tile (made by the other group; every tile takes part in the chain so the
        //      barrier use stays uniform - the value only matters inside an item)
        if (gj > 0) {
          named_bar_sync(hand_in, 64);
          if (j > 0) {
            const float m_prev = mail[r_in_tile];
            if (m_prev != m_cur) {  // the other group moved the reference: bring this group's partial sum along
              l_part = (m_cur == -INFINITY) ? 0.f : l_part * fast_exp2((m_cur - m_prev) * scale_log2);
              m_cur = m_prev;
            }
          }
        }
        if (tr) p.trace[gj * 16 + 3] = clock64();
        // ---- lazy rescale: only move the reference max when it grew by more than 2^8 (or on the first tile)
        bool grow = (m_tile - m_cur) * scale_log2 > RESCALE_THRESHOLD;
        if (m_tile == -INFINITY) grow = false;
        if (j > 0 && __any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? fast_exp2((m_cur - m_tile) * scale_log2) : 1.0f;  // m_cur == -inf -> 0
          l_part *= alpha;
          mbar_wait(&o_done[(gj - 1) & 1], ((gj - 1) >> 1) & 1);  // PV(gj-1) finished: O is stable until PV(gj)
          tcgen05_fence_after();
          const uint32_t o_addr = tmem_base + lane_addr + Cfg::O_COL + (k & 1) * Cfg::O_STRIDE;
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
        mail[r_in_tile] = m_cur;
        __threadfence_block();
        named_bar_arrive(hand_out, 64);
        const float m_scaled = (m_cur == -INFINITY) ? 0.f : m_cur * scale_log2;
        // ---- pass 2: exponentials, 64 columns at a time; P (bf16 pairs) overwrites the S columns just consumed
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t sr[64];
          tmem_ld_32x32b_x32(s_base + hh * 64, sr);
          tmem_ld_32x32b_x32(s_base + hh * 64 + 32, sr + 32);
          tmem_ld_wait();
          if (need_mask) {
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (j * BN + hh * 64 + c > limit) sr[c] = __float_as_uint(-INFINITY);
          }
          uint32_t pk[32];
          l_part += softmax_exp64<VAR>(sr, !need_mask, scale_log2, m_scaled, pk);
          tmem_st_32x32b_x32(s_base + hh * 32, pk);
        }
        if (tr) p.trace[gj * 16 + 4] = clock64();
        tmem_st_wait();
        tcgen05_fence_before();
        if (tr) p.trace[gj * 16 + 5] = clock64();
        if (!first_done && pend_k >= 0) wait_pending_o();
        mbar_arrive(&p_ready[G]);
        if (!first_done) {
          first_done = true;
          // previous item's O: its last PV finished long ago; the other group and the tensor pipe keep going
          if (pend_k >= 0) run_epilogue(tr ? p.trace + gj * 16 : nullptr);
        }
      }
      if (!first_done && pend_k >= 0) {  // this group has no tile in a one-tile item
        wait_pending_o();
        run_epilogue(nullptr);
      }
      // ---- item end: post this group's (sum, max); combined in the deferred epilogue
      post[((k & 1) * 2 + G) * 128 + r_in_tile] = make_float2(l_part, m_cur);
      pend_k = k;
      pend_g_last = g + n_tiles - 1;
      g += n_tiles;
      if (!more) break;
    }
    if (pend_k >= 0) {
      wait_pending_o();
      run_epilogue(nullptr);
    }
    // the last tile's publisher arrived on a hand-off barrier nobody synchronises on: complete it
    if (g > 0 && (g & 1) == G) named_bar_sync(hand_in, 64);
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
        if (VAR == 32) {
          // DEBUG / measurement only (garbage output): no softmax at all - P is whatever S left in TMEM.  The tile
          // period of this variant is the floor set by the TMA + tcgen05.mma side alone.
          l_sum = 1.f;
          tcgen05_fence_before();
          mbar_arrive(&p_ready[buf]);
          if (j == 0 && pend_g_last >= 0) {
            run_epilogue(k - 1, nullptr);
            pend_g_last = -1;
          }
          continue;
        }
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
        if (VAR != 0 && !need_mask) {
          constexpr int P = VAR >> 1;
          const float2 sc2 = make_float2(scale_log2, scale_log2), nm2 = make_float2(-m_scaled, -m_scaled);
          float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
          for (int c = 0; c < 32; ++c) {
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
          l_sum += (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y);
        } else {
          float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const float p0 = fast_exp2(fmaf(__uint_as_float(sr[2 * c]), scale_log2, -m_scaled));
            const float p1 = fast_exp2(fmaf(__uint_as_float(sr[2 * c + 1]), scale_log2, -m_scaled));
            ps4[c & 3] += p0 + p1;
            pk[c] = pack_bf16x2(p0, p1);
          }
          l_sum += (ps4[0] + ps4[1]) + (ps4[2] + ps4[3]);
        }
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

template <int HD, bool CAUSAL, int VAR, bool SPLIT>
int launch_tc_var(const AttnParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = TcCfg<HD>;
  auto kern = attn_tc_kernel<HD, CAUSAL, VAR, SPLIT>;
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

int g_attn_variant = -1;  // -1: read SLIME_ATTN_VARIANT / the compile-time default on first use

template <int HD, bool CAUSAL>
int launch_tc(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (g_attn_variant < 0) {
    const char* e = getenv("SLIME_ATTN_VARIANT");
    g_attn_variant = e != nullptr ? atoi(e) : SLIME_ATTN_VARIANT_DEFAULT;
  }
  switch (g_attn_variant) {
    case 0: return launch_tc_var<HD, CAUSAL, 0, false>(p, num_sms, stream);
    case 5: return launch_tc_var<HD, CAUSAL, 5, false>(p, num_sms, stream);
    case 9: return launch_tc_var<HD, CAUSAL, 9, false>(p, num_sms, stream);
    case 32: return launch_tc_var<HD, CAUSAL, 32, false>(p, num_sms, stream);  // measurement only: no softmax
    case 21: return launch_tc_var<HD, CAUSAL, 5, true>(p, num_sms, stream);
    default:
      slime_set_error("attention: unknown softmax variant %d (0, 5, 9, 21, 32)", g_attn_variant);
      return SLIME_EINVAL;
  }
}

}  // namespace

extern "C" int slime_attention_set_variant(int variant) {
  const bool known = variant == -1 || variant == 0 || variant == 5 || variant == 9 || variant == 21 || variant == 32;
  if (!known) {
    slime_set_error("attention: unknown softmax variant %d (0, 5, 9, 21 = kv-split, 32 = no softmax; -1 = default)", variant);
    return SLIME_EINVAL;
  }
  g_attn_variant = variant;
  return SLIME_OK;
}

int slime_launch_attention_tc(const AttnParams& p, int num_sms, cudaStream_t stream) {
  if (p.batch <= 0 || p.seqlen_q <= 0 || p.seqlen_k <= 0) return SLIME_OK;
  if (p.head_dim == 64) {
    return p.causal ? launch_tc<64, true>(p, num_sms, stream) : launch_tc<64, false>(p, num_sms, stream);
  }
  return p.causal ? launch_tc<128, true>(p, num_sms, stream) : launch_tc<128, false>(p, num_sms, stream);
}
