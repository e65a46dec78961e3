// Weight-streaming GEMM for the decode step: C[M,N] = A[M,K] * W[N,K]^T with M <= 32 rows
// (one token per sequence, reference llava_llama.py:139 -> HF generation loop -> LlamaDecoderLayer on [B,1,H]).
//
// The problem is HBM-bound - every weight byte is used for M <= 32 MACs - so the kernel is built around the
// weight stream, not around the tensor pipe (a 128-row tcgen05 tile would run 1..12 % full and, worse, an
// N = 4096 projection yields only 16..32 tiles for 148 SMs):
//   * a work item is 8 weight rows x one k-split; each WARP owns an item and streams its rows with 16-byte
//     non-allocating loads, 8 rows x 64 contiguous bytes per instruction, two groups of SK_U loads in flight
//     (16 KB per warp) - 16 warps per SM keep far more than the ~43 KB per SM in flight that HBM3e needs;
//   * the activations [M, k-split] are staged once per CTA in shared memory (row stride = 64 mod 128 bytes:
//     the 8-row x 64-byte fragment reads are bank-conflict free);
//   * math on mma.sync.m16n8k16 (fp32 accumulate) with the WEIGHT rows as the n = 8 operand and the activation
//     rows as the m = 16 operand.  A dot product does not care about the order of k, so the 16 bytes a lane
//     loads (k = kb + 8c .. 8c+7) are fed to two MMAs as the k-slots that lane owns in the fragment layout and
//     the activation fragment is read with the same permutation - no ldmatrix, no shuffles;
//   * few-row problems (N = 4096 projections) are split along K so that every warp of the GPU has an item;
//     partial sums go to an fp32 scratch [splits, M, N] and a finishing kernel adds them in a FIXED order
//     (deterministic, unlike atomics), applies the epilogue and - for the two residual projections of a decoder
//     layer - also the RMSNorm that follows (HF modeling_llama.py:62-67), saving a launch and a round trip.
// Epilogues are the tcgen05 GEMM's (gemm_epilogue.cuh): bias, quick/erf GELU, SwiGLU on interleaved (gate, up)
// columns, rotary embedding on interleaved (i, i + hd/2) columns, residual, bf16 or fp32 output.
#include <cstdlib>

#include "elementwise.h"
#include "errors.h"
#include "gemm.h"

namespace {

#include "gemm_epilogue.cuh"  // act_silu / act_quick_gelu / act_gelu_erf

constexpr int SK_THREADS = 512;
constexpr int SK_WARPS = SK_THREADS / 32;
constexpr int SK_U = 8;             // 32-element k-steps per load group (two groups in flight)
constexpr int SK_KC_MAX = 4096;     // activations staged per CTA: 16 * MT rows x kc <= 128 KB

int g_skinny_mode = -1;  // -1 unset, 0 off, 1 on

SLIME_DEVINL uint4 ld_stream16(const bf16* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
SLIME_DEVINL uint4 lds16(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
SLIME_DEVINL void mma_16816(float* d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32." SLIME_MMA_SYNC_TYPE "." SLIME_MMA_SYNC_TYPE
               ".f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Epilogue of one output pair (row m, columns n, n + 1; n even), same order of operations as epi_math8.
SLIME_DEVINL void skinny_store(const GemmParams& p, int epi, int m, int n, float v0, float v1) {
  if (epi == GEMM_EPI_SWIGLU) {  // columns (gate_j, up_j) interleaved
    p.out[static_cast<size_t>(m) * p.out_ld + (n >> 1)] = float_to_elem(act_silu(v0) * v1);
    return;
  }
  if (epi == GEMM_EPI_ROPE && n < p.rope_cols) {
    const int pos = min(max(__ldcg(p.rope_pos + m), 0), p.rope_max_pos - 1);
    const float2 cs = __ldg(p.rope_table + static_cast<size_t>(pos) * p.rope_half + ((n & (2 * p.rope_half - 1)) >> 1));
    const float lo = v0, hi = v1;
    v0 = lo * cs.x - hi * cs.y;
    v1 = hi * cs.x + lo * cs.y;
  }
  if (p.bias != nullptr) {
    const float2 b = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(p.bias + n)));
    v0 += b.x;
    v1 += b.y;
  }
  if (epi == GEMM_EPI_QUICK_GELU) {
    v0 = act_quick_gelu(v0);
    v1 = act_quick_gelu(v1);
  } else if (epi == GEMM_EPI_GELU_ERF) {
    v0 = act_gelu_erf(v0);
    v1 = act_gelu_erf(v1);
  }
  if (p.residual != nullptr) {
    const float2 r = unpack_bf16x2(__ldcg(reinterpret_cast<const uint32_t*>(p.residual + static_cast<size_t>(m) * p.res_ld + n)));
    v0 += r.x;
    v1 += r.y;
  }
  if (p.out_f32 != nullptr) {
    *reinterpret_cast<float2*>(p.out_f32 + static_cast<size_t>(m) * p.out_ld + n) = make_float2(v0, v1);
  } else {
    const uint32_t pk = pack_bf16x2(v0, v1);
    *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(m) * p.out_ld + n) = pk;
    if (p.kv_k != nullptr && n >= p.kv_q_cols) {  // this step's K / V row goes straight into the cache as well
      const int pos = __ldcg(p.kv_lens + m);
      if (pos >= 0 && pos < p.kv_cache_len) {
        const int col = n - p.kv_q_cols;
        bf16* plane = col < p.kv_dim ? p.kv_k : p.kv_v;
        const int cc = col < p.kv_dim ? col : col - p.kv_dim;
        *reinterpret_cast<uint32_t*>(plane + (static_cast<size_t>(m) * p.kv_cache_len + pos) * p.kv_dim + cc) = pk;
      }
    }
  }
}

// grid = ctas_per_split * splits; CTA i works on k-split i % splits.  partial == nullptr: epilogue in place
// (splits == 1); else partial[split][m][n] fp32.
// MT = 16-row activation groups (M <= 16 MT); HALF (MT == 1 only): M <= 8 - only the M real rows are staged (8 KB of
// shared memory at M = 1 instead of 132 KB, i.e. more L1 for the in-flight weight loads) and the other fragment rows
// are fed as zero registers.
// (A 256-thread variant with two CTAs per SM - small enough to become resident next to the attention / finishing CTAs
// under PDL - was measured and is not faster: 3.54 vs 3.54 ms at B = 1, 4.06 vs 3.94 at B = 4, 4.12 vs 4.30 at B = 8.)
template <int MT, bool HALF>
__global__ void __launch_bounds__(SK_THREADS, 1)
gemm_skinny_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw, const GemmParams p,
                   int epi, int splits, int kc, int xs_stride, float* __restrict__ partial, int* __restrict__ counters) {
  extern __shared__ __align__(16) uint8_t sk_smem[];
  __shared__ float s_rstd[32];
  __shared__ float s_part[32 * SK_WARPS];
  bf16* xs = reinterpret_cast<bf16*>(sk_smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, c = lane & 3;
  const int split = blockIdx.x % splits, cta = blockIdx.x / splits, ctas = gridDim.x / splits;
  const int ks = split * kc;
  const int klen = min(p.K, ks + kc) - ks;  // > 0 and a multiple of 32 by construction of the plan
  const int steps = klen >> 5;
  const int n_tiles = p.N >> 3;

  // ---- the weights do not depend on the previous kernel: start this warp's first load group, let the next kernel
  //      of the stream get resident, and only then wait for the producer of the activations (PDL) ----
  uint4 wa[SK_U], wb[SK_U];
  const int t_first = cta * SK_WARPS + warp;
  if (t_first < n_tiles) {
    const bf16* wp0 = W + static_cast<size_t>(t_first * 8 + g) * ldw + ks + c * 8;
#pragma unroll
    for (int u = 0; u < SK_U; ++u)
      if (u < steps) wa[u] = ld_stream16(wp0 + u * 32);
#pragma unroll
    for (int u = 0; u < SK_U; ++u)
      if (SK_U + u < steps) wb[u] = ld_stream16(wp0 + (SK_U + u) * 32);
  }
  pdl_trigger();
  pdl_wait();

  const bool a_norm = p.a_norm_w != nullptr;
  if (a_norm) {
    // ---- RMSNorm of the A rows on their way in (a_norm_w): thread t owns the 16-byte chunks t, t + 512, ... of EVERY row, so
    //      the loads of up to AN_B rows are in flight at once (one L2 round trip per batch of rows); the sum of squares runs
    //      over the FULL row (all K columns), the chunks of this CTA's k-split are parked raw in shared memory and normalised
    //      in place by the thread that parked them once 1/rms is known ----
    constexpr int AN_B = 8;
    const int chunks_k = p.K >> 3, c_lo = ks >> 3, c_hi = (ks + klen) >> 3;
    for (int r0 = 0; r0 < p.M; r0 += AN_B) {
      float sq[AN_B];
#pragma unroll
      for (int j = 0; j < AN_B; ++j) sq[j] = 0.f;
      for (int ch = tid; ch < chunks_k; ch += SK_THREADS) {
        uint4 v[AN_B];
#pragma unroll
        for (int j = 0; j < AN_B; ++j)
          v[j] = (r0 + j < p.M) ? __ldcg(reinterpret_cast<const uint4*>(A + static_cast<size_t>(r0 + j) * lda) + ch)
                                : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < AN_B; ++j) {
          const float2 a = unpack_bf16x2(v[j].x), b = unpack_bf16x2(v[j].y), c2 = unpack_bf16x2(v[j].z), d = unpack_bf16x2(v[j].w);
          sq[j] += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c2.x * c2.x + c2.y * c2.y + d.x * d.x + d.y * d.y;
          if (r0 + j < p.M && ch >= c_lo && ch < c_hi)
            *reinterpret_cast<uint4*>(xs + static_cast<size_t>(r0 + j) * xs_stride + (ch - c_lo) * 8) = v[j];
        }
      }
#pragma unroll
      for (int j = 0; j < AN_B; ++j) {
        const float tsum = warp_sum(sq[j]);
        if (lane == 0 && r0 + j < p.M) s_part[(r0 + j) * SK_WARPS + warp] = tsum;
      }
    }
    __syncthreads();
    if (tid < p.M) {
      float tot = 0.f;
#pragma unroll
      for (int w2 = 0; w2 < SK_WARPS; ++w2) tot += s_part[tid * SK_WARPS + w2];
      s_rstd[tid] = rsqrtf(tot / static_cast<float>(p.K) + p.a_norm_eps);
    }
    __syncthreads();
    for (int ch = c_lo + tid; ch < c_hi; ch += SK_THREADS) {  // same (thread, chunk) ownership as above
      const uint4 g4 = __ldg(reinterpret_cast<const uint4*>(p.a_norm_w) + ch);
      for (int r = 0; r < p.M; ++r) {
        uint4* px = reinterpret_cast<uint4*>(xs + static_cast<size_t>(r) * xs_stride + (ch - c_lo) * 8);
        const float rs = s_rstd[r];
        auto nrm = [rs](uint32_t x, uint32_t gw) {  // HF: weight * (x * rstd).to(dtype)
          const float2 xf = unpack_bf16x2(x), gf = unpack_bf16x2(gw);
          const float a = elem_to_float(float_to_elem(xf.x * rs)), b = elem_to_float(float_to_elem(xf.y * rs));
          return pack_bf16x2(gf.x * a, gf.y * b);
        };
        const uint4 v = *px;
        *px = make_uint4(nrm(v.x, g4.x), nrm(v.y, g4.y), nrm(v.z, g4.z), nrm(v.w, g4.w));
      }
    }
    if (!HALF) {  // fragment rows past M are zero
      const int chunks = klen >> 3;
      for (int idx = p.M * chunks + tid; idx < MT * 16 * chunks; idx += SK_THREADS) {
        const int r = idx / chunks, ch = idx - r * chunks;
        *reinterpret_cast<uint4*>(xs + static_cast<size_t>(r) * xs_stride + ch * 8) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  } else {
    // ---- stage the activations of this k-split (rows >= M are zero) ----
    const int chunks = klen >> 3;
    const int total = (HALF ? p.M : MT * 16) * chunks;  // HALF: exactly the M <= 8 real rows (8 KB at M = 1)
    for (int idx = tid; idx < total; idx += SK_THREADS) {
      const int r = idx / chunks, ch = idx - r * chunks;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (r < p.M) v = __ldcg(reinterpret_cast<const uint4*>(A + static_cast<size_t>(r) * lda + ks + ch * 8));
      *reinterpret_cast<uint4*>(xs + static_cast<size_t>(r) * xs_stride + ch * 8) = v;
    }
  }
  __syncthreads();

  const uint32_t xs_lane = smem_u32(xs) + static_cast<uint32_t>(g * xs_stride + c * 8) * 2u;
  const uint32_t row8 = static_cast<uint32_t>(8 * xs_stride) * 2u;

  for (int t = cta * SK_WARPS + warp; t < n_tiles; t += ctas * SK_WARPS) {
    const int n0 = t * 8;
    const bf16* wp = W + static_cast<size_t>(n0 + g) * ldw + ks + c * 8;
    float acc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) acc[mt][0] = acc[mt][1] = acc[mt][2] = acc[mt][3] = 0.f;

    if (t != t_first) {
#pragma unroll
      for (int u = 0; u < SK_U; ++u)
        if (u < steps) wa[u] = ld_stream16(wp + u * 32);
    }

    auto compute = [&](const uint4(&w)[SK_U], int sbase) {
#pragma unroll
      for (int u = 0; u < SK_U; ++u) {
        const int s = sbase + u;
        if (s < steps) {
          const uint32_t xa_addr = xs_lane + static_cast<uint32_t>(s) * 64u;
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint4 xa = (HALF && g >= p.M) ? make_uint4(0u, 0u, 0u, 0u)
                                                : lds16(xa_addr + static_cast<uint32_t>(mt) * 2u * row8);  // row mt*16 + g
            const uint4 xb = HALF ? make_uint4(0u, 0u, 0u, 0u)
                                  : lds16(xa_addr + static_cast<uint32_t>(mt) * 2u * row8 + row8);   // row mt*16 + g + 8
            mma_16816(acc[mt], xa.x, xb.x, xa.y, xb.y, w[u].x, w[u].y);  // k = kb + 8c + {0,1 | 2,3}
            mma_16816(acc[mt], xa.z, xb.z, xa.w, xb.w, w[u].z, w[u].w);  // k = kb + 8c + {4,5 | 6,7}
          }
        }
      }
    };

    for (int s0 = 0; s0 < steps; s0 += 2 * SK_U) {
      if (t != t_first || s0 != 0) {
#pragma unroll
        for (int u = 0; u < SK_U; ++u)
          if (s0 + SK_U + u < steps) wb[u] = ld_stream16(wp + (s0 + SK_U + u) * 32);
      }
      compute(wa, s0);
#pragma unroll
      for (int u = 0; u < SK_U; ++u)
        if (s0 + 2 * SK_U + u < steps) wa[u] = ld_stream16(wp + (s0 + 2 * SK_U + u) * 32);
      compute(wb, s0 + SK_U);
    }

    // accumulator fragment: acc[mt][0,1] = C[mt*16 + g][n0 + 2c, +1], acc[mt][2,3] = C[mt*16 + g + 8][same]
    const int n = n0 + 2 * c;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = mt * 16 + g + h * 8;
        if (m < p.M) {
          if (partial != nullptr) {
            *reinterpret_cast<float2*>(partial + (static_cast<size_t>(split) * p.M + m) * p.N + n) =
                make_float2(acc[mt][2 * h], acc[mt][2 * h + 1]);
          } else {
            skinny_store(p, epi, m, n, acc[mt][2 * h], acc[mt][2 * h + 1]);
          }
        }
      }
    }
    if (partial != nullptr && counters != nullptr) {
      // ---- in-kernel finish: the warp holding the last ticket of this tile sums the partials of ALL splits in split
      //      order (its own included, read back like the others: the sum does not depend on who is last) ----
      __threadfence();
      __syncwarp();
      int ticket = 0;
      if (lane == 0) ticket = atomicAdd(counters + t, 1);
      ticket = __shfl_sync(0xffffffffu, ticket, 0);
      if (ticket == splits - 1) {
        __threadfence();
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int m = mt * 16 + g + h * 8;
            if (m < p.M) {
              const float* src = partial + static_cast<size_t>(m) * p.N + n;
              const size_t sstride = static_cast<size_t>(p.M) * p.N;
              float v0 = 0.f, v1 = 0.f;
#pragma unroll 4
              for (int s = 0; s < splits; ++s) {
                const float2 tt = __ldcg(reinterpret_cast<const float2*>(src + s * sstride));
                v0 += tt.x;
                v1 += tt.y;
              }
              skinny_store(p, epi, m, n, v0, v1);
            }
          }
        }
        if (lane == 0) counters[t] = 0;  // zero again for the next launch that uses the array
      }
    }
  }
}

// Sum of the k-split partials in split order + epilogue; one thread per output pair.
__global__ void __launch_bounds__(256) skinny_finish_kernel(const float* __restrict__ partial, int splits,
                                                            const GemmParams p, int epi, int work_ctas) {
  pdl_trigger();
  if (static_cast<int>(blockIdx.x) >= work_ctas) {  // extra CTAs: only pull a later projection's weights into L2
    l2_prefetch_slice(p.l2_prefetch, p.l2_prefetch_bytes, blockIdx.x - work_ctas, gridDim.x - work_ctas);
    return;
  }
  pdl_wait();
  const int pairs = p.N >> 1;
  const long long idx = blockIdx.x * 256LL + threadIdx.x;
  if (idx >= static_cast<long long>(p.M) * pairs) return;
  const int m = static_cast<int>(idx / pairs), n = static_cast<int>(idx % pairs) * 2;
  float v0 = 0.f, v1 = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float2 t = __ldcg(reinterpret_cast<const float2*>(partial + (static_cast<size_t>(s) * p.M + m) * p.N + n));
    v0 += t.x;
    v1 += t.y;
  }
  skinny_store(p, epi, m, n, v0, v1);
}

// Same, for out = acc (+bias) + residual followed by RMSNorm of the new row (one CTA per row):
//   out[m] = bf16(sum);  norm_out[m] = norm_w * bf16(out[m] * rsqrt(mean(out[m]^2) + eps))   (HF LlamaRMSNorm)
// Latency-bound (a row is 8..40 KB of partials in L2): 1024 threads, every load of a thread issued before the first
// use, the rounded row kept in registers between the two phases.  (The first version - 256 threads walking the row in
// 8 dependent round trips - took 18 us per launch, 20 % of a batch-16 decode step.)
constexpr int FN_THREADS = 1024;
constexpr int FN_SLOTS = 4;  // column pairs per thread: N <= 2 * FN_THREADS * FN_SLOTS = 8192

template <int SPLITS>  // 0 = run-time count
__global__ void __launch_bounds__(FN_THREADS) skinny_finish_norm_kernel(const float* __restrict__ partial, int splits_rt,
                                                                        const GemmParams p) {
  __shared__ float red[FN_THREADS / 32];
  pdl_trigger();
  if (static_cast<int>(blockIdx.x) >= p.M) {  // extra CTAs: only pull a later projection's weights into L2 (opt-in)
    l2_prefetch_slice(p.l2_prefetch, p.l2_prefetch_bytes, blockIdx.x - p.M, gridDim.x - p.M);
    return;
  }
  const int splits = SPLITS > 0 ? SPLITS : splits_rt;
  const int m = blockIdx.x, tid = threadIdx.x;
  uint32_t wn[FN_SLOTS];  // the norm weight does not depend on the previous kernel: before the wait
#pragma unroll
  for (int i = 0; i < FN_SLOTS; ++i) {
    const int n = (tid + i * FN_THREADS) * 2;
    wn[i] = n < p.N ? __ldg(reinterpret_cast<const uint32_t*>(p.norm_w + n)) : 0u;
  }
  pdl_wait();
  float2 v[FN_SLOTS];
  uint32_t res[FN_SLOTS];
#pragma unroll
  for (int i = 0; i < FN_SLOTS; ++i) {
    const int n = (tid + i * FN_THREADS) * 2;
    v[i] = make_float2(0.f, 0.f);
    res[i] = 0u;
    if (n < p.N) {
      if (p.residual != nullptr)
        res[i] = __ldcg(reinterpret_cast<const uint32_t*>(p.residual + static_cast<size_t>(m) * p.res_ld + n));
      if constexpr (SPLITS > 0) {
        float2 t[SPLITS];
#pragma unroll
        for (int s = 0; s < SPLITS; ++s)
          t[s] = __ldcg(reinterpret_cast<const float2*>(partial + (static_cast<size_t>(s) * p.M + m) * p.N + n));
#pragma unroll
        for (int s = 0; s < SPLITS; ++s) {  // fixed order: deterministic
          v[i].x += t[s].x;
          v[i].y += t[s].y;
        }
      } else {
        for (int s = 0; s < splits; ++s) {
          const float2 t = __ldcg(reinterpret_cast<const float2*>(partial + (static_cast<size_t>(s) * p.M + m) * p.N + n));
          v[i].x += t.x;
          v[i].y += t.y;
        }
      }
    }
  }
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < FN_SLOTS; ++i) {
    const int n = (tid + i * FN_THREADS) * 2;
    if (n < p.N) {
      float v0 = v[i].x, v1 = v[i].y;
      if (p.bias != nullptr) {
        const float2 b = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(p.bias + n)));
        v0 += b.x;
        v1 += b.y;
      }
      if (p.residual != nullptr) {
        const float2 r = unpack_bf16x2(res[i]);
        v0 += r.x;
        v1 += r.y;
      }
      const uint32_t pk = pack_bf16x2(v0, v1);
      *reinterpret_cast<uint32_t*>(p.out + static_cast<size_t>(m) * p.out_ld + n) = pk;
      v[i] = unpack_bf16x2(pk);  // the row as the next kernel will read it
      sq += v[i].x * v[i].x + v[i].y * v[i].y;
    }
  }
  sq = warp_sum(sq);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < FN_THREADS / 32; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / p.N + p.norm_eps);
#pragma unroll
  for (int i = 0; i < FN_SLOTS; ++i) {
    const int n = (tid + i * FN_THREADS) * 2;
    if (n < p.N) {
      const float2 w = unpack_bf16x2(wn[i]);
      const float a = elem_to_float(float_to_elem(v[i].x * rstd)), b = elem_to_float(float_to_elem(v[i].y * rstd));
      *reinterpret_cast<uint32_t*>(p.norm_out + static_cast<size_t>(m) * p.norm_ld + n) = pack_bf16x2(w.x * a, w.y * b);
    }
  }
}

struct SkinnyPlan {
  bool ok = false;
  int mt = 1, splits = 1, kc = 0, xs_stride = 0;
  bool use_partial = false;
};

bool skinny_enabled() {
  if (g_skinny_mode < 0) {
    const char* e = getenv("SLIME_GEMM_SKINNY");
    g_skinny_mode = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return g_skinny_mode != 0;
}

SkinnyPlan make_plan(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi, int num_sms) {
  SkinnyPlan pl;
  if (p.M < 1 || p.M > 32 || p.K % 32 != 0 || p.N % 8 != 0 || p.row_map != nullptr || p.res_period != 0) return pl;
  if (lda % 8 != 0 || ldw % 8 != 0 || (reinterpret_cast<uintptr_t>(A) & 15) != 0 || (reinterpret_cast<uintptr_t>(W) & 15) != 0)
    return pl;
  if (p.out_ld % 2 != 0 || (p.residual != nullptr && p.res_ld % 2 != 0)) return pl;
  if (epi == GEMM_EPI_SWIGLU && (p.N % 16 != 0 || p.out == nullptr)) return pl;
  if (p.kv_k != nullptr && (p.out == nullptr || p.kv_v == nullptr || p.kv_lens == nullptr || p.kv_dim % 2 != 0 ||
                            p.kv_q_cols % 2 != 0 || p.kv_q_cols + 2 * p.kv_dim != p.N))
    return pl;
  const bool fused_norm = p.norm_w != nullptr;
  if (p.a_norm_w != nullptr && (p.K % 8 != 0 || p.M > 32)) return pl;
  if (fused_norm && p.tile_counters != nullptr) return pl;  // one or the other
  if (fused_norm && (epi != GEMM_EPI_NONE || p.out == nullptr || p.out_f32 != nullptr || p.norm_out == nullptr ||
                     p.N > 8192 || p.norm_ld % 2 != 0))
    return pl;
  pl.mt = p.M <= 16 ? 1 : 2;
  const int kc_max = SK_KC_MAX / pl.mt;
  const int k32 = p.K / 32;
  const int items = p.N / 8;
  const int warps = num_sms * SK_WARPS;
  int s = 1;
  if (p.force_splits > 0) {
    s = p.force_splits;
  } else {
    while (s < 8 && items * s < (warps * 2) / 3 && p.K / (s * 2) >= 512) s *= 2;
  }
  while (s < 64 && (k32 + s - 1) / s * 32 > kc_max) s *= 2;
  if (s > num_sms) return pl;
  int kc32 = (k32 + s - 1) / s;
  s = (k32 + kc32 - 1) / kc32;  // drop empty trailing splits
  pl.kc = kc32 * 32;
  if (pl.kc > kc_max) return pl;
  pl.splits = s;
  pl.use_partial = s > 1 || fused_norm;
  if (pl.use_partial) {
    const size_t need = static_cast<size_t>(s) * p.M * p.N;
    if (p.splitk_ws == nullptr || p.splitk_ws_floats < need) {
      if (fused_norm || p.force_splits > 0) return pl;
      // no scratch: run unsplit if the activations still fit
      if (p.K > kc_max) return pl;
      pl.splits = 1;
      pl.kc = p.K;
      pl.use_partial = false;
    }
  }
  pl.xs_stride = pl.kc + ((pl.kc % 64 == 0) ? 32 : 0);  // row stride = 64 (mod 128) bytes
  pl.ok = true;
  return pl;
}

template <int MT, bool HALF>
int launch_skinny(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi, const SkinnyPlan& pl,
                  int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  const int rows = HALF ? 8 : MT * 16;
  const int smem = (HALF ? p.M : rows) * pl.xs_stride * 2;
  if (!attr_set) {
    SLIME_CHECK_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<MT, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          rows * (SK_KC_MAX / MT + 32) * 2));
    attr_set = true;
  }
  const int ctas = (num_sms / pl.splits) * pl.splits;
  slime_prof_begin(2, static_cast<double>(p.N) * p.K * sizeof(bf16), stream);
  const cudaError_t le = slime_launch_kernel(gemm_skinny_kernel<MT, HALF>, dim3(ctas), dim3(SK_THREADS), smem, stream, true, A, lda,
                                             W, ldw, p, epi, pl.splits, pl.kc, pl.xs_stride,
                                             pl.use_partial ? p.splitk_ws : static_cast<float*>(nullptr),
                                             pl.use_partial ? p.tile_counters : static_cast<int*>(nullptr));
  slime_prof_end(stream);
  SLIME_CHECK_CUDA(le);
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}

}  // namespace

extern "C" int slime_gemm_set_skinny_mode(int mode) {
  g_skinny_mode = mode < 0 ? -1 : (mode != 0 ? 1 : 0);
  return SLIME_OK;
}

bool slime_gemm_skinny_enabled() { return skinny_enabled(); }

bool slime_gemm_skinny_applies(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi,
                               int num_sms) {
  if (!skinny_enabled() && p.force_splits <= 0) return false;
  return make_plan(A, lda, W, ldw, p, epi, num_sms).ok;
}

int slime_launch_gemm_skinny(const bf16* A, int lda, const bf16* W, int ldw, const GemmParams& p, int epi,
                             int num_sms, cudaStream_t stream) {
  const SkinnyPlan pl = make_plan(A, lda, W, ldw, p, epi, num_sms);
  SLIME_REQUIRE(pl.ok, "skinny gemm: problem M=%d N=%d K=%d not supported", p.M, p.N, p.K);
  static int rows8 = -1;  // SLIME_SKINNY_ROWS8=0: always stage 16 rows (A/B)
  if (rows8 < 0) {
    const char* e = getenv("SLIME_SKINNY_ROWS8");
    rows8 = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  if (pl.mt == 1 && p.M <= 8 && rows8 != 0) {
    SLIME_PROPAGATE((launch_skinny<1, true>(A, lda, W, ldw, p, epi, pl, num_sms, stream)));
  } else if (pl.mt == 1) {
    SLIME_PROPAGATE((launch_skinny<1, false>(A, lda, W, ldw, p, epi, pl, num_sms, stream)));
  } else {
    SLIME_PROPAGATE((launch_skinny<2, false>(A, lda, W, ldw, p, epi, pl, num_sms, stream)));
  }
  if (!pl.use_partial || p.tile_counters != nullptr) return SLIME_OK;  // finished in the kernel
  const float* part = p.splitk_ws;
  const int pf_ctas = (p.l2_prefetch != nullptr && p.l2_prefetch_bytes > 0) ? num_sms / 2 : 0;
  if (p.norm_w != nullptr) {
    const dim3 grid(p.M + pf_ctas), block(FN_THREADS);
    switch (pl.splits) {
      case 1: SLIME_CHECK_CUDA(slime_launch_kernel(skinny_finish_norm_kernel<1>, grid, block, 0, stream, true, part, pl.splits, p)); break;
      case 2: SLIME_CHECK_CUDA(slime_launch_kernel(skinny_finish_norm_kernel<2>, grid, block, 0, stream, true, part, pl.splits, p)); break;
      case 4: SLIME_CHECK_CUDA(slime_launch_kernel(skinny_finish_norm_kernel<4>, grid, block, 0, stream, true, part, pl.splits, p)); break;
      case 8: SLIME_CHECK_CUDA(slime_launch_kernel(skinny_finish_norm_kernel<8>, grid, block, 0, stream, true, part, pl.splits, p)); break;
      default: SLIME_CHECK_CUDA(slime_launch_kernel(skinny_finish_norm_kernel<0>, grid, block, 0, stream, true, part, pl.splits, p)); break;
    }
  } else {
    const long long pairs = static_cast<long long>(p.M) * (p.N / 2);
    const int work_ctas = static_cast<int>((pairs + 255) / 256);
    SLIME_CHECK_CUDA(slime_launch_kernel(skinny_finish_kernel, dim3(work_ctas + pf_ctas), dim3(256), 0, stream, true, part,
                                         pl.splits, p, epi, work_ctas));
  }
  SLIME_AFTER_LAUNCH();
  return SLIME_OK;
}
