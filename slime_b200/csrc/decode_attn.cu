// Single-query (decode step) attention over the KV cache written by the prefill - the step that follows the
// prefill in generate() (reference llava_llama.py:139 -> HF generation loop).
// HBM-bound: every cached K/V row of the sequence is read once per q head group; one CTA per (q head, sequence),
// 4 warps stride over the cached positions with an online softmax each and are merged through shared memory.
#include <cstdlib>

#include "attention.h"
#include "errors.h"

namespace {

constexpr int NT = 128;

template <int HD>
__global__ void __launch_bounds__(NT) decode_attn_kernel(const bf16* __restrict__ q, int q_ld,
                                                         const bf16* __restrict__ kcache,
                                                         const bf16* __restrict__ vcache, int cache_len,
                                                         const int* __restrict__ lens, int heads, int kv_heads,
                                                         float scale_log2, bf16* __restrict__ out, int out_ld) {
  constexpr int EPL = HD / 32;  // elements per lane
  __shared__ float s_m[4], s_l[4];
  __shared__ float s_acc[4][HD];
  const int head = blockIdx.x, b = blockIdx.y;
  const int kvh = head / (heads / kv_heads);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int len = lens[b] + 1;  // cached tokens + the one appended in this step
  const int KD = kv_heads * HD;

  float qv[EPL];
  {
    const bf16* qp = q + static_cast<long long>(b) * q_ld + head * HD + lane * EPL;
#pragma unroll
    for (int i = 0; i < EPL; ++i) qv[i] = elem_to_float(qp[i]);
  }
  float m = -INFINITY, l = 0.f, acc[EPL];
#pragma unroll
  for (int i = 0; i < EPL; ++i) acc[i] = 0.f;

  const long long base = (static_cast<long long>(b) * cache_len) * KD + kvh * HD + lane * EPL;
  for (int pos = warp; pos < len; pos += 4) {
    const bf16* kp = kcache + base + static_cast<long long>(pos) * KD;
    const bf16* vp = vcache + base + static_cast<long long>(pos) * KD;
    float kv[EPL], vv[EPL];
    if (EPL == 4) {
      const uint2 ku = *reinterpret_cast<const uint2*>(kp);
      const uint2 vu = *reinterpret_cast<const uint2*>(vp);
      float2 t;
      t = unpack_bf16x2(ku.x); kv[0] = t.x; kv[1] = t.y;
      t = unpack_bf16x2(ku.y); kv[2] = t.x; kv[3] = t.y;
      t = unpack_bf16x2(vu.x); vv[0] = t.x; vv[1] = t.y;
      t = unpack_bf16x2(vu.y); vv[2] = t.x; vv[3] = t.y;
    } else {
#pragma unroll
      for (int i = 0; i < EPL; ++i) {
        kv[i] = elem_to_float(kp[i]);
        vv[i] = elem_to_float(vp[i]);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < EPL; ++i) s += qv[i] * kv[i];
    s = warp_sum(s) * scale_log2;
    const float m_new = fmaxf(m, s);
    const float alpha = exp2f(m - m_new);  // m == -inf -> 0
    const float pexp = exp2f(s - m_new);
    l = l * alpha + pexp;
#pragma unroll
    for (int i = 0; i < EPL; ++i) acc[i] = acc[i] * alpha + pexp * vv[i];
    m = m_new;
  }
  if (lane == 0) {
    s_m[warp] = m;
    s_l[warp] = l;
  }
#pragma unroll
  for (int i = 0; i < EPL; ++i) s_acc[warp][lane * EPL + i] = acc[i];
  __syncthreads();
  if (warp == 0) {
    const float mt = fmaxf(fmaxf(s_m[0], s_m[1]), fmaxf(s_m[2], s_m[3]));
    float lt = 0.f, o[EPL];
#pragma unroll
    for (int i = 0; i < EPL; ++i) o[i] = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float a = (s_m[w] == -INFINITY) ? 0.f : exp2f(s_m[w] - mt);
      lt += s_l[w] * a;
#pragma unroll
      for (int i = 0; i < EPL; ++i) o[i] += s_acc[w][lane * EPL + i] * a;
    }
    const float inv = lt > 0.f ? 1.0f / lt : 0.f;
    bf16* op = out + static_cast<long long>(b) * out_ld + head * HD + lane * EPL;
#pragma unroll
    for (int i = 0; i < EPL; ++i) op[i] = float_to_elem(o[i] * inv);
  }
}

// ------------------------------------------------------------------------------------------
// Split-KV ("flash-decoding") variant: the kernel above has one CTA per (q head, sequence), i.e. 32 CTAs for one
// sequence, each walking all cached positions with one 256-byte row in flight per warp, and every K/V row is
// fetched once per q head of its GQA group.  Here a CTA owns (kv head, sequence, kv split): the cached positions
// are divided over `splits` CTAs so that a single sequence still fills the GPU, all G = heads / kv_heads query heads
// of the group share every K/V row, and two rows per 8-lane group are in flight.
//   lanes: 4 position groups x 8 lanes per warp; lane sl of a group holds features [8 sl, 8 sl + 8) and
//   [64 + 8 sl, 64 + 8 sl + 8) of the row, so each 16-byte load instruction of a group reads 128 contiguous bytes;
//   a dot product is 16 FMAs + 3 shuffles per (position, q head).
// Every (warp, position group) keeps its own online-softmax state; the 16 states of a CTA are merged through shared
// memory, the per-split results (max, sum, unnormalised accumulator) go to an fp32 scratch and a small kernel merges
// the splits (splits == 1: the CTA writes the final row itself).
// ------------------------------------------------------------------------------------------
constexpr int DS_HD = 128;
constexpr int DS_POS_PER_ITER = 16;  // 4 warps x 4 position groups

SLIME_DEVINL uint4 ld_nc16(const bf16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
SLIME_DEVINL void unpack16(const uint4& a, const uint4& b, float (&f)[16]) {
  const uint32_t u[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 t = unpack_bf16x2(u[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}

template <int G>
__global__ void __launch_bounds__(NT) decode_attn_split_kernel(const bf16* __restrict__ q, int q_ld,
                                                               const bf16* __restrict__ kcache,
                                                               const bf16* __restrict__ vcache, int cache_len,
                                                               const int* __restrict__ lens, int heads, int kv_heads,
                                                               float scale_log2, int splits, float* __restrict__ part,
                                                               bf16* __restrict__ out, int out_ld,
                                                               const void* __restrict__ pf_ptr, size_t pf_bytes) {
  // merged through shared memory: [16 states][G][HD] accumulators + [16][G] max / sum
  extern __shared__ __align__(16) uint8_t ds_smem[];
  float* s_acc = reinterpret_cast<float*>(ds_smem);
  float* s_m = s_acc + 16 * G * DS_HD;
  float* s_l = s_m + 16 * G;

  pdl_trigger();
  if (pf_bytes != 0)  // opt-in experiment (slime_set_decode_prefetch): pull a later projection's weights into L2
    l2_prefetch_slice(pf_ptr, pf_bytes, (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x,
                      gridDim.x * gridDim.y * gridDim.z);
  pdl_wait();  // q, the cache rows appended in this step and lens come from the kernels before

  const int kvh = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, pg = lane >> 3, sl = lane & 7;
  const int len = min(__ldcg(lens + b) + 1, cache_len);  // cached tokens + the one appended in this step
  const int KD = kv_heads * DS_HD;
  // this split's positions [p0, p1): equal chunks, multiples of 32 (two rounds of 16 positions per iteration)
  const int chunk = ((len + splits - 1) / splits + 2 * DS_POS_PER_ITER - 1) / (2 * DS_POS_PER_ITER) * (2 * DS_POS_PER_ITER);
  const int p0 = split * chunk, p1 = min(len, p0 + chunk);

  float qv[G][16];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    const bf16* qp = q + static_cast<long long>(b) * q_ld + (kvh * G + gi) * DS_HD + sl * 8;
    const uint4 a = __ldcg(reinterpret_cast<const uint4*>(qp)), c = __ldcg(reinterpret_cast<const uint4*>(qp + 64));
    unpack16(a, c, qv[gi]);
#pragma unroll
    for (int e = 0; e < 16; ++e) qv[gi][e] *= scale_log2;
  }
  float m[G], l[G], acc[G][16];
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    m[gi] = -INFINITY;
    l[gi] = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[gi][e] = 0.f;
  }

  const long long base = (static_cast<long long>(b) * cache_len) * KD + kvh * DS_HD + sl * 8;
  struct Row {
    uint4 k0, k1, v0, v1;
  };
  auto fetch = [&](int pos) {
    Row r;
    r.k0 = r.k1 = r.v0 = r.v1 = make_uint4(0u, 0u, 0u, 0u);
    if (pos < p1) {
      const bf16* kp = kcache + base + static_cast<long long>(pos) * KD;
      const bf16* vp = vcache + base + static_cast<long long>(pos) * KD;
      r.k0 = ld_nc16(kp); r.k1 = ld_nc16(kp + 64); r.v0 = ld_nc16(vp); r.v1 = ld_nc16(vp + 64);
    }
    return r;
  };
  auto step = [&](const Row& r, int pos) {
    float kf[16], vf[16];
    unpack16(r.k0, r.k1, kf);
    unpack16(r.v0, r.v1, vf);
    const bool valid = pos < p1;
#pragma unroll
    for (int gi = 0; gi < G; ++gi) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 16; ++e) s = fmaf(qv[gi][e], kf[e], s);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      if (valid) {
        const float m_new = fmaxf(m[gi], s);
        const float alpha = exp2f(m[gi] - m_new);  // first position: exp2(-inf) = 0
        const float pe = exp2f(s - m_new);
        l[gi] = l[gi] * alpha + pe;
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[gi][e] = fmaf(acc[gi][e], alpha, pe * vf[e]);
        m[gi] = m_new;
      }
    }
  };

  // Two positions per lane group and iteration, the next two already in flight: 4 rows x 512 B per 8 lanes = 16 KB per
  // warp in flight.  Every lane of a warp runs the same number of iterations (the shuffles need the full warp).
  int pos = p0 + warp * 4 + pg;
  Row ra = fetch(pos), rb = fetch(pos + DS_POS_PER_ITER);
  const int iters = (p1 - p0 + 2 * DS_POS_PER_ITER - 1) / (2 * DS_POS_PER_ITER);
  for (int it = 0; it < iters; ++it, pos += 2 * DS_POS_PER_ITER) {
    const Row na = fetch(pos + 2 * DS_POS_PER_ITER), nb = fetch(pos + 3 * DS_POS_PER_ITER);
    step(ra, pos);
    step(rb, pos + DS_POS_PER_ITER);
    ra = na;
    rb = nb;
  }

  // ---- merge the 16 (warp, position group) states of this CTA ----
  const int st = warp * 4 + pg;
#pragma unroll
  for (int gi = 0; gi < G; ++gi) {
    if (sl == 0) {
      s_m[st * G + gi] = m[gi];
      s_l[st * G + gi] = l[gi];
    }
    float* dst = s_acc + (st * G + gi) * DS_HD + sl * 8;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      dst[e] = acc[gi][e];
      dst[64 + e] = acc[gi][8 + e];
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < G * DS_HD; o += NT) {
    const int gi = o / DS_HD, e = o % DS_HD;
    float mt = -INFINITY;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) mt = fmaxf(mt, s_m[s2 * G + gi]);
    float lt = 0.f, val = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      const float ms = s_m[s2 * G + gi];
      const float a = (ms == -INFINITY) ? 0.f : exp2f(ms - mt);
      lt += s_l[s2 * G + gi] * a;
      val += s_acc[(s2 * G + gi) * DS_HD + e] * a;
    }
    const int head = kvh * G + gi;
    if (part == nullptr) {
      out[static_cast<long long>(b) * out_ld + head * DS_HD + e] = float_to_elem(lt > 0.f ? val / lt : 0.f);
    } else {
      float* pp = part + ((static_cast<long long>(b) * heads + head) * splits + split) * (DS_HD + 2);
      pp[e] = val;
      if (e == 0) {
        pp[DS_HD] = mt;
        pp[DS_HD + 1] = lt;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Tensor-core variant of the split-KV kernel.  The CUDA-core kernel above spends ~256 instructions per position set
// (bf16 unpacking, 128 FMAs, shuffles, the per-position rescale of G x 128 accumulators) and at batch 16 is
// instruction-issue bound at a third of the HBM rate.  Here a warp owns 16-position K / V tiles staged by cp.async in
// a private 2-stage shared-memory ring (XOR-swizzled 256-byte rows: conflict-free ldmatrix) and runs the FlashAttention
// recurrence on mma.sync.m16n8k16 with the G <= 8 query heads of the group as the (padded) 16-row operand:
//   S[g, pos] = Q[g, :] K[pos, :]^T   8 k-steps x 2 position blocks     (ldmatrix of K rows = the col-major B operand)
//   O[g, :]  += P[g, pos] V[pos, :]   16 feature blocks, P re-used from the S accumulator registers as the A operand,
//                                     V through ldmatrix.trans
// ~180 instructions per 16 positions instead of ~1000.  Per-warp states are merged exactly like the kernel above and
// written in the same partial format, so decode_attn_merge_kernel serves both.
// ------------------------------------------------------------------------------------------
// A staged K / V row is 256 bytes = 16 chunks of 16 bytes; chunk ch of row r sits at position ch ^ (r & 7), so the 8 rows
// of an ldmatrix phase and the 16 chunks of a cp.async row both spread over all banks without padding.  2 stages x
// (K + V) x 4 warps = 64 KB per CTA: two CTAs fit the 132 KB carve-out the weight-streaming GEMM runs with, which
// matters under PDL - an SM cannot be re-partitioned between overlapping kernels, and a projection that inherits a
// 208 KB carve-out (the first, padded 3-stage version of this kernel) loses its L1 and with it 40 % of its bandwidth.
constexpr int DM_ROW = 256;
constexpr int DM_TILE = 16 * DM_ROW;  // one 16-position K (or V) tile
constexpr int DM_STAGES = 2;
constexpr int DM_WARP_BYTES = DM_STAGES * 2 * DM_TILE;
constexpr int DM_SMEM = 4 * DM_WARP_BYTES;  // 65536
SLIME_DEVINL uint32_t dm_off(int row, int chunk) { return static_cast<uint32_t>(row * DM_ROW + ((chunk ^ (row & 7)) << 4)); }

SLIME_DEVINL void cp_async16_zfill(uint32_t dst, const bf16* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}
SLIME_DEVINL void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
SLIME_DEVINL void ldsm4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
// rows 8..15 of the A operand are padding: a2a3 = a6a7 = 0
SLIME_DEVINL void mma_rows8(float* d, uint32_t a0, uint32_t a2, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." SLIME_MMA_SYNC_TYPE "." SLIME_MMA_SYNC_TYPE
               ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(0u), "r"(a2), "r"(0u), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(NT) decode_attn_mma_kernel(const bf16* __restrict__ q, int q_ld,
                                                             const bf16* __restrict__ kcache,
                                                             const bf16* __restrict__ vcache, int cache_len,
                                                             const int* __restrict__ lens, int heads, int kv_heads, int G,
                                                             float scale_log2, int splits, float* __restrict__ part,
                                                             bf16* __restrict__ out, int out_ld,
                                                             int* __restrict__ counters) {
  extern __shared__ __align__(16) uint8_t ds_smem[];
  __shared__ int s_last;
  pdl_trigger();
  pdl_wait();  // q, the cache rows appended in this step and lens come from the kernels before

  const int kvh = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, c = lane & 3;
  const int len = min(__ldcg(lens + b) + 1, cache_len);
  const int KD = kv_heads * DS_HD;
  // this split's positions [p0, p1): equal chunks, multiples of 64 (4 warps x 16-position tiles)
  const int chunk = ((len + splits - 1) / splits + 63) / 64 * 64;
  const int p0 = split * chunk, p1 = min(len, p0 + chunk);
  const int n_tiles = p1 > p0 ? (p1 - p0 + 15) / 16 : 0;
  const int my_tiles = n_tiles > warp ? (n_tiles - warp + 3) / 4 : 0;

  // A operand: row g = query head kvh * G + g (rows >= G are zero), k-step ks = features [16 ks, 16 ks + 16)
  uint32_t qa[8][2];
#pragma unroll
  for (int ks = 0; ks < 8; ++ks) {
    qa[ks][0] = qa[ks][1] = 0u;
    if (g < G) {
      const bf16* qp = q + static_cast<long long>(b) * q_ld + (kvh * G + g) * DS_HD + ks * 16 + 2 * c;
      qa[ks][0] = __ldcg(reinterpret_cast<const uint32_t*>(qp));
      qa[ks][1] = __ldcg(reinterpret_cast<const uint32_t*>(qp + 8));
    }
  }
  float o[16][4];
#pragma unroll
  for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  const bf16* kbase = kcache + (static_cast<long long>(b) * cache_len) * KD + kvh * DS_HD;
  const bf16* vbase = vcache + (static_cast<long long>(b) * cache_len) * KD + kvh * DS_HD;
  const uint32_t wbase = smem_u32(ds_smem) + static_cast<uint32_t>(warp) * DM_WARP_BYTES;
  auto issue = [&](int i) {  // the warp's i-th tile -> stage i % DM_STAGES (always commits a group)
    if (i < my_tiles) {
      const int pos0 = p0 + (warp + 4 * i) * 16;
      const uint32_t sk = wbase + static_cast<uint32_t>(i % DM_STAGES) * (2 * DM_TILE);
#pragma unroll
      for (int j = 0; j < 8; ++j) {  // one instruction = 2 rows x 256 contiguous bytes
        const int row = j * 2 + (lane >> 4), ch = lane & 15;
        const int pos = pos0 + row;
        const bool ok = pos < p1;
        const long long off = ok ? static_cast<long long>(pos) * KD + ch * 8 : 0;
        const uint32_t d = sk + dm_off(row, ch);
        cp_async16_zfill(d, kbase + off, ok ? 16 : 0);            // rows past the end are zero-filled: P = 0 there,
        cp_async16_zfill(d + DM_TILE, vbase + off, ok ? 16 : 0);  // and 0 x garbage could be NaN
      }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
  };

#pragma unroll
  for (int i = 0; i < DM_STAGES - 1; ++i) issue(i);
  for (int i = 0; i < my_tiles; ++i) {
    issue(i + DM_STAGES - 1);
    asm volatile("cp.async.wait_group %0;\n" ::"n"(DM_STAGES - 1) : "memory");
    __syncwarp();
    const uint32_t sk = wbase + static_cast<uint32_t>(i % DM_STAGES) * (2 * DM_TILE);
    const uint32_t sv = sk + DM_TILE;

    // ---- S = Q K^T: s[0] = positions 0..7 of the tile, s[1] = positions 8..15 (this thread: columns 2c, 2c + 1) ----
    float s[2][4];
    s[0][0] = s[0][1] = s[0][2] = s[0][3] = s[1][0] = s[1][1] = s[1][2] = s[1][3] = 0.f;
    const int k_row = ((lane >> 4) << 3) + (lane & 7), k_ch = (lane >> 3) & 1;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t r0, r1, r2, r3;
      ldsm4(sk + dm_off(k_row, ks * 2 + k_ch), r0, r1, r2, r3);
      mma_rows8(s[0], qa[ks][0], qa[ks][1], r0, r1);
      mma_rows8(s[1], qa[ks][0], qa[ks][1], r2, r3);
    }
    // ---- online softmax of row g over the tile's 16 positions ----
    const int pos_t = p0 + (warp + 4 * i) * 16 + 2 * c;
    float sc[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int pos = pos_t + (e >> 1) * 8 + (e & 1);
      sc[e] = pos < p1 ? s[e >> 1][e & 1] * scale_log2 : -INFINITY;
    }
    float mx = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    const float m_new = fmaxf(m_run, mx);  // finite: every issued tile has at least one valid position
    const float alpha = exp2f(m_run - m_new);
    m_run = m_new;
    float pr[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) pr[e] = exp2f(sc[e] - m_new);
    l_run = l_run * alpha + (pr[0] + pr[1]) + (pr[2] + pr[3]);
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      o[n][0] *= alpha;
      o[n][1] *= alpha;
    }
    // ---- O += P V: P (row g; positions 2c, 2c+1 | 8 + 2c, 8 + 2c + 1) is the A operand as it sits in registers ----
    const uint32_t pa0 = pack_bf16x2(pr[0], pr[1]), pa2 = pack_bf16x2(pr[2], pr[3]);
    const int v_row = (((lane >> 3) & 1) << 3) + (lane & 7), v_ch = lane >> 4;
#pragma unroll
    for (int dp = 0; dp < 8; ++dp) {
      uint32_t r0, r1, r2, r3;
      ldsm4_t(sv + dm_off(v_row, dp * 2 + v_ch), r0, r1, r2, r3);
      mma_rows8(o[2 * dp], pa0, pa2, r0, r1);
      mma_rows8(o[2 * dp + 1], pa0, pa2, r2, r3);
    }
    __syncwarp();  // every lane is done with this stage before it is refilled
  }
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
  __syncthreads();  // the merge buffers below alias the rings

  // ---- merge the 4 warp states: [4][8][HD] accumulators + [4][8] max / sum ----
  float* s_acc = reinterpret_cast<float*>(ds_smem);
  float* s_m = s_acc + 4 * 8 * DS_HD;
  float* s_l = s_m + 32;
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 1);
  l_run += __shfl_xor_sync(0xffffffffu, l_run, 2);
  if (g < G) {
    float* dst = s_acc + (warp * 8 + g) * DS_HD + 2 * c;
#pragma unroll
    for (int n = 0; n < 16; ++n) *reinterpret_cast<float2*>(dst + n * 8) = make_float2(o[n][0], o[n][1]);
    if (c == 0) {
      s_m[warp * 8 + g] = m_run;
      s_l[warp * 8 + g] = l_run;
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * DS_HD; idx += NT) {
    const int gi = idx / DS_HD, e = idx % DS_HD;
    const float mt = fmaxf(fmaxf(s_m[gi], s_m[8 + gi]), fmaxf(s_m[16 + gi], s_m[24 + gi]));
    float lt = 0.f, val = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float ms = s_m[w * 8 + gi];
      const float a = (ms == -INFINITY) ? 0.f : exp2f(ms - mt);
      lt += s_l[w * 8 + gi] * a;
      val += s_acc[(w * 8 + gi) * DS_HD + e] * a;
    }
    const int head = kvh * G + gi;
    if (part == nullptr) {
      out[static_cast<long long>(b) * out_ld + head * DS_HD + e] = float_to_elem(lt > 0.f ? val / lt : 0.f);
    } else {
      float* pp = part + ((static_cast<long long>(b) * heads + head) * splits + split) * (DS_HD + 2);
      pp[e] = val;
      if (e == 0) {
        pp[DS_HD] = mt;
        pp[DS_HD + 1] = lt;
      }
    }
  }
  if (part == nullptr || counters == nullptr) return;
  // ---- in-kernel merge of the kv splits (no merge launch): the CTA that takes the last ticket of its (sequence, kv head)
  //      combines the partial states of ALL splits in split order - same arithmetic whoever comes last ----
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counters + b * kv_heads + kvh, 1) == splits - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float* s_w = reinterpret_cast<float*>(ds_smem);  // [8][32] per-split weight 2^(m_s - M); first the raw m_s
  float* s_ls = s_w + 8 * 32;                      // [8][32] l_s
  float* s_lt = s_ls + 8 * 32;                     // [8]
  const float* pbase = part + (static_cast<long long>(b) * heads + kvh * G) * splits * (DS_HD + 2);
  for (int idx = threadIdx.x; idx < G * splits; idx += NT) {  // round trip 1: (m_s, l_s) of every head and split
    const int gi = idx / splits, sp = idx - gi * splits;
    const float* pp = pbase + (static_cast<long long>(gi) * splits + sp) * (DS_HD + 2);
    s_w[gi * 32 + sp] = __ldcg(pp + DS_HD);
    s_ls[gi * 32 + sp] = __ldcg(pp + DS_HD + 1);
  }
  __syncthreads();
  if (threadIdx.x < G) {
    const int gi = threadIdx.x;
    float mt = -INFINITY;
    for (int sp = 0; sp < splits; ++sp) mt = fmaxf(mt, s_w[gi * 32 + sp]);
    float lt = 0.f;
    for (int sp = 0; sp < splits; ++sp) {  // fixed order over the splits: deterministic
      const float ms = s_w[gi * 32 + sp];
      const float a = (ms == -INFINITY) ? 0.f : exp2f(ms - mt);
      s_w[gi * 32 + sp] = a;
      lt += s_ls[gi * 32 + sp] * a;
    }
    s_lt[gi] = lt;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * DS_HD; idx += NT) {  // round trip 2: the partial accumulators
    const int gi = idx / DS_HD, e = idx % DS_HD;
    const float* pp = pbase + static_cast<long long>(gi) * splits * (DS_HD + 2) + e;
    float val = 0.f;
#pragma unroll 8
    for (int sp = 0; sp < splits; ++sp) val += __ldcg(pp + sp * (DS_HD + 2)) * s_w[gi * 32 + sp];
    const float lt = s_lt[gi];
    out[static_cast<long long>(b) * out_ld + (kvh * G + gi) * DS_HD + e] = float_to_elem(lt > 0.f ? val / lt : 0.f);
  }
  if (threadIdx.x == 0) counters[b * kv_heads + kvh] = 0;  // zero again for the next launch
}

// out[b, head] = sum_s acc_s 2^(m_s - M) / sum_s l_s 2^(m_s - M)   (fixed order over the splits)
__global__ void __launch_bounds__(DS_HD) decode_attn_merge_kernel(const float* __restrict__ part, int heads, int splits,
                                                                  bf16* __restrict__ out, int out_ld) {
  __shared__ float s_a[DS_HD];  // per-split weight 2^(m_s - M) (splits <= 128)
  __shared__ float s_red[2];
  pdl_trigger();
  pdl_wait();
  const int head = blockIdx.x, b = blockIdx.y, e = threadIdx.x;
  const float* pp = part + (static_cast<long long>(b) * heads + head) * splits * (DS_HD + 2);
  // round trip 1: thread s fetches (m_s, l_s); the CTA agrees on M and on the weights (one warp: splits <= 32)
  float ms = -INFINITY, ls = 0.f;
  if (e < splits) {
    ms = __ldcg(pp + e * (DS_HD + 2) + DS_HD);
    ls = __ldcg(pp + e * (DS_HD + 2) + DS_HD + 1);
  }
  if (e < 32) {
    const float mt = warp_max(ms);
    const float a = (ms == -INFINITY) ? 0.f : exp2f(ms - mt);
    s_a[e] = a;
    float lt = 0.f;  // fixed order over the splits: deterministic
    for (int s = 0; s < splits; ++s) lt += __shfl_sync(0xffffffffu, ls * a, s);
    if (e == 0) s_red[0] = lt;
  }
  __syncthreads();
  // round trip 2: all partial accumulators of this feature at once
  float val = 0.f;
#pragma unroll 8
  for (int s = 0; s < splits; ++s) val += __ldcg(pp + s * (DS_HD + 2) + e) * s_a[s];
  const float lt = s_red[0];
  out[static_cast<long long>(b) * out_ld + head * DS_HD + e] = float_to_elem(lt > 0.f ? val / lt : 0.f);
}

int g_decode_attn_mode = -1;  // -1 unset; 0 = one CTA per (q head, sequence); 1 = split-KV on CUDA cores; 2 = split-KV on mma.sync

int decode_attn_mode() {
  if (g_decode_attn_mode < 0) {
    const char* e = getenv("SLIME_DECODE_ATTN");
    g_decode_attn_mode = (e != nullptr && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
  }
  return g_decode_attn_mode;
}

bool decode_split_supported(int heads, int kv_heads, int head_dim) {
  const int mode = decode_attn_mode();
  if (mode == 0 || head_dim != DS_HD || kv_heads <= 0 || heads % kv_heads != 0) return false;
  const int G = heads / kv_heads;
  if (mode == 2) return G <= 8;
  return G == 1 || G == 2 || G == 4;  // G = 8 would spill (2 x 128 fp32 of state per lane): old kernel
}

template <int G>
int launch_split(const bf16* q, int q_ld, const bf16* kc, const bf16* vc, int cache_len, const int* lens, int batch,
                 int heads, int kv_heads, float sl2, int splits, float* part, bf16* out, int out_ld, const void* pf_ptr,
                 size_t pf_bytes, cudaStream_t stream) {
  const int smem = (16 * G * DS_HD + 32 * G) * static_cast<int>(sizeof(float));
  static bool attr_set = false;
  if (!attr_set && smem > 48 * 1024) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_split_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set = true;
  }
  dim3 grid(kv_heads, batch, splits);
  SLIME_CHECK_CUDA(slime_launch_kernel(decode_attn_split_kernel<G>, grid, dim3(NT), smem, stream, true, q, q_ld, kc, vc,
                                       cache_len, lens, heads, kv_heads, sl2, splits,
                                       splits > 1 ? part : static_cast<float*>(nullptr), out, out_ld, pf_ptr, pf_bytes));
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

// cache_rows[i] = sample(i) * cache_len + pos_ids[i]   (packed prefill row -> slot in the per-sequence cache); -1 (row
// dropped) when the position lies outside [0, cache_len) or the sample outside the cache's batch
__global__ void cache_rows_kernel(const int* __restrict__ cu, const int* __restrict__ pos_ids, int B, int total,
                                  int cache_batch, int cache_len, int* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  if (i >= cu[B]) {  // zero rows past the real total (no-host-sync prefill: buffers sized by an upper bound)
    rows[i] = -1;
    return;
  }
  int lo = 0, hi = B;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= i) lo = mid; else hi = mid;
  }
  const int pos = pos_ids[i];
  rows[i] = (pos >= 0 && pos < cache_len && lo < cache_batch) ? lo * cache_len + pos : -1;
}

// rows[b] = b * cache_len + lens[b]   (slot of the token appended in this decode step)
__global__ void append_rows_kernel(const int* __restrict__ lens, int B, int cache_len, int* __restrict__ rows) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) rows[b] = lens[b] < cache_len ? b * cache_len + lens[b] : -1;
}

// Appends the K / V slices of this step's qkv rows to the cache: row b -> slot lens[b] of sequence b
// (dropped when the cache is full).  One launch for both planes, 16 bytes per thread.
__global__ void kv_append_kernel(const bf16* __restrict__ k, const bf16* __restrict__ v, int ld, bf16* __restrict__ kc,
                                 bf16* __restrict__ vc, int KD, const int* __restrict__ lens, int B, int cache_len) {
  pdl_trigger();
  pdl_wait();
  const int chunks = KD >> 3;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * chunks * 2) return;
  const int which = i / (B * chunks), r = i % (B * chunks);
  const int b = r / chunks, ch = r % chunks;
  const int pos = __ldcg(lens + b);
  if (pos < 0 || pos >= cache_len) return;
  const bf16* src = (which == 0 ? k : v) + static_cast<long long>(b) * ld + ch * 8;
  bf16* dst = (which == 0 ? kc : vc) + (static_cast<long long>(b) * cache_len + pos) * KD + ch * 8;
  *reinterpret_cast<uint4*>(dst) = __ldcg(reinterpret_cast<const uint4*>(src));
}

}  // namespace

int slime_launch_kv_append(const bf16* k, const bf16* v, int ld, bf16* kcache, bf16* vcache, int KD, const int* lens,
                           int B, int cache_len, cudaStream_t stream) {
  SLIME_REQUIRE(KD % 8 == 0 && ld % 8 == 0, "kv append: bad KD=%d ld=%d", KD, ld);
  if (B <= 0) return SLIME_OK;
  const int total = B * (KD / 8) * 2;
  SLIME_CHECK_CUDA(slime_launch_kernel(kv_append_kernel, dim3((total + 255) / 256), dim3(256), 0, stream, true, k, v, ld,
                                       kcache, vcache, KD, lens, B, cache_len));
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

extern "C" int slime_decode_attention_set_mode(int mode) {
  g_decode_attn_mode = (mode < 0 || mode > 2) ? -1 : mode;
  return SLIME_OK;
}

int slime_decode_attention_splits(int batch, int heads, int kv_heads, int head_dim, int cache_len, int num_sms) {
  if (!decode_split_supported(heads, kv_heads, head_dim) || batch <= 0) return 0;
  const int ctas = kv_heads * batch;
  int s = (2 * num_sms) / ctas;                      // ONE wave of (at most) two CTAs per SM: no tail
  const int by_len = (cache_len + 63) / 64;          // about 64 cached positions per split at least
  if (s > by_len) s = by_len;
  if (s > 32) s = 32;
  return s < 1 ? 1 : s;
}

size_t slime_decode_attention_ws_floats(int batch, int heads, int splits) {
  return splits > 1 ? static_cast<size_t>(batch) * heads * splits * (DS_HD + 2) : 0;
}

int slime_launch_decode_attention(const bf16* q, int q_ld, const bf16* kcache, const bf16* vcache, int cache_len,
                                  const int* lens, int batch, int heads, int kv_heads, int head_dim, float scale,
                                  bf16* out, int out_ld, int splits, float* ws, const void* pf_ptr, size_t pf_bytes,
                                  cudaStream_t stream, int* merge_counters) {
  SLIME_REQUIRE(head_dim == 64 || head_dim == 128, "decode attention: head_dim %d unsupported", head_dim);
  if (batch <= 0) return SLIME_OK;
  const float sl2 = scale * 1.4426950408889634f;
  if (splits >= 1 && decode_split_supported(heads, kv_heads, head_dim)) {
    SLIME_REQUIRE(splits == 1 || ws != nullptr, "decode attention: %d kv splits need a scratch buffer", splits);
    SLIME_REQUIRE(splits <= 32, "decode attention: at most 32 kv splits (%d given)", splits);
    SLIME_REQUIRE(q_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(kcache) & 15) == 0 && (reinterpret_cast<uintptr_t>(vcache) & 15) == 0,
                  "decode attention: q / cache must be 16-byte aligned");
    const int G = heads / kv_heads;
    const double bytes = 2.0 * batch * static_cast<double>(cache_len) * kv_heads * head_dim * sizeof(bf16);  // upper bound
    slime_prof_begin(1, bytes, stream);
    int rc = SLIME_OK;
    if (decode_attn_mode() == 2) {
      static bool attr_set = false;
      if (!attr_set) {
        SLIME_CHECK_CUDA(cudaFuncSetAttribute(decode_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DM_SMEM));
        attr_set = true;
      }
      const cudaError_t le = slime_launch_kernel(decode_attn_mma_kernel, dim3(kv_heads, batch, splits), dim3(NT), DM_SMEM, stream,
                                                 true, q, q_ld, kcache, vcache, cache_len, lens, heads, kv_heads, G, sl2,
                                                 splits, splits > 1 ? ws : static_cast<float*>(nullptr), out, out_ld,
                                                 splits > 1 ? merge_counters : static_cast<int*>(nullptr));
      slime_prof_end(stream);
      SLIME_CHECK_CUDA(le);
      SLIME_AFTER_LAUNCH();
      if (merge_counters != nullptr) return SLIME_OK;  // the kv splits were merged inside the kernel
    } else {
    switch (G) {
      case 1: rc = launch_split<1>(q, q_ld, kcache, vcache, cache_len, lens, batch, heads, kv_heads, sl2, splits, ws, out, out_ld, pf_ptr, pf_bytes, stream); break;
      case 2: rc = launch_split<2>(q, q_ld, kcache, vcache, cache_len, lens, batch, heads, kv_heads, sl2, splits, ws, out, out_ld, pf_ptr, pf_bytes, stream); break;
      default: rc = launch_split<4>(q, q_ld, kcache, vcache, cache_len, lens, batch, heads, kv_heads, sl2, splits, ws, out, out_ld, pf_ptr, pf_bytes, stream); break;
    }
    slime_prof_end(stream);
    }
    SLIME_PROPAGATE(rc);
    if (splits > 1) {
      const float* part = ws;
      SLIME_CHECK_CUDA(slime_launch_kernel(decode_attn_merge_kernel, dim3(heads, batch), dim3(DS_HD), 0, stream, true, part,
                                           heads, splits, out, out_ld));
      SLIME_AFTER_LAUNCH();
    }
    return SLIME_OK;
  }
  dim3 grid(heads, batch);
  slime_prof_begin(1, 0.0, stream);
  if (head_dim == 128) {
    decode_attn_kernel<128><<<grid, NT, 0, stream>>>(q, q_ld, kcache, vcache, cache_len, lens, heads, kv_heads, sl2, out, out_ld);
  } else {
    decode_attn_kernel<64><<<grid, NT, 0, stream>>>(q, q_ld, kcache, vcache, cache_len, lens, heads, kv_heads, sl2, out, out_ld);
  }
  slime_prof_end(stream);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_cache_rows(const int* cu, const int* pos_ids, int B, int total, int cache_batch, int cache_len, int* rows,
                            cudaStream_t stream) {
  if (total <= 0) return SLIME_OK;
  cache_rows_kernel<<<(total + 255) / 256, 256, 0, stream>>>(cu, pos_ids, B, total, cache_batch, cache_len, rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

int slime_launch_append_rows(const int* lens, int B, int cache_len, int* rows, cudaStream_t stream) {
  if (B <= 0) return SLIME_OK;
  append_rows_kernel<<<(B + 127) / 128, 128, 0, stream>>>(lens, B, cache_len, rows);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}
