// Epilogue of the tcgen05 GEMM (included by gemm_sm100.cu / gemm2_sm100.cu inside their anonymous namespaces).
//
// 8 epilogue warps: warp w may only touch TMEM lane quadrant w % 4, so two warps share each
// quadrant and split the accumulator columns in halves.  Each thread owns one output row and walks
// its columns in chunks of 32; the loop is software-pipelined - the tcgen05.ld of chunk c+1 and the
// global loads of its residual values are issued BEFORE chunk c is processed - because with only
// 8 resident warps per SM nothing else hides the TMEM / global-load latency (the first ncu capture
// of the K = 1024 ViT GEMMs showed the tensor pipe 18-46 % active, waiting on the epilogue).
//
// HBM access pattern (GemmParams::epi_mode):
//   0  direct  : every thread writes (and reads the residual of) its own row in 16-byte pieces - one store
//                instruction touches 32 different rows, each 32-byte sector is written by two instructions.
//   1  staged  : each warp transposes the 32 x 32 chunk through its own 2 KB of shared memory (16-byte slots
//                XOR-swizzled by row pair: conflict-free both ways) and writes 8 rows x 64 contiguous bytes per
//                instruction (SwiGLU: 16 rows x 32 bytes) - the same scheme as the attention epilogue; the
//                residual chunk is likewise fetched 8 rows x 64 bytes per instruction and handed to the owning
//                threads through a second 2 KB staging tile.
// The mode is a template parameter of the kernels (STAGED); fp32 outputs (lm_head on B rows) always go direct.
#pragma once

constexpr int EPI_STAGE_BYTES = 4096;  // per epilogue warp: [0, 2048) output transpose, [2048, 4096) residual

// sigmoid(y) = 0.5 + 0.5 * tanh(y / 2): ONE MUFU op (tanh.approx) instead of ex2 + rcp.  The SFUs retire only
// 16 results per clock per SM, and the ncu capture of the K = 1024 ViT fc1 GEMM showed its quick-GELU epilogue
// (2 MUFU per element) exactly as long as the main loop (tensor pipe 70 % active).
SLIME_DEVINL float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
SLIME_DEVINL float fast_sigmoid(float y) { return fmaf(0.5f, fast_tanh(0.5f * y), 0.5f); }
SLIME_DEVINL float act_quick_gelu(float x) { return x * fast_sigmoid(1.702f * x); }
SLIME_DEVINL float act_gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
SLIME_DEVINL float act_silu(float x) { return x * fast_sigmoid(x); }

struct EpiRow {
  bool store_ok;
  int out_row;
  int res_row;
  int pos;      // GEMM_EPI_ROPE: position id of this row
  float scale;  // GemmParams::row_scale of this row (1 when absent)
  float ssq;    // GemmParams::sumsq_out: running sum of squares of the (bf16-rounded) outputs of the current 64 columns
};

// Row bookkeeping of the coalesced phases: in iteration i a lane handles row r_i = i * (32 / SLOTS) + lane / SLOTS
// of its warp's 32 rows (SLOTS = 16-byte slots per staged row) - the owner's values are fetched by shuffle.
struct EpiCoal {
  int out_row[4];
  int res_row[4];
  unsigned ok;  // bit i: row r_i is stored
};

template <int SLOTS>
SLIME_DEVINL uint32_t epi_stage_off(int r, int q) {
  if constexpr (SLOTS == 4) {
    return static_cast<uint32_t>(r * 64 + ((q ^ ((r >> 1) & 3)) << 4));
  } else {
    return static_cast<uint32_t>(r * 32 + ((q ^ ((r >> 2) & 1)) << 4));
  }
}

template <int SLOTS>
SLIME_DEVINL EpiCoal epi_make_coal(const EpiRow& er, int lane) {
  EpiCoal ec;
  ec.ok = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (i * (32 / SLOTS) + lane / SLOTS) & 31;
    ec.out_row[i] = __shfl_sync(0xffffffffu, er.out_row, r);
    ec.res_row[i] = __shfl_sync(0xffffffffu, er.res_row, r);
    const int ok = __shfl_sync(0xffffffffu, er.store_ok ? 1 : 0, r);
    if (i < SLOTS && ok) ec.ok |= 1u << i;
  }
  return ec;
}

// ---- residual fetch of one chunk, issued one chunk ahead ----
template <int EPI, bool STAGED>
SLIME_DEVINL void epi_issue_residual(const GemmParams& p, const EpiRow& er, const EpiCoal& ec, int lane, int col0,
                                     uint4 (&res)[4]) {
  if constexpr (EPI != GEMM_EPI_SWIGLU) {
    if (p.residual == nullptr) return;
    if constexpr (STAGED) {
      const int q = lane & 3;
      if (col0 + q * 8 < p.N) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ec.ok & (1u << i))
            res[i] = *reinterpret_cast<const uint4*>(p.residual + static_cast<size_t>(ec.res_row[i]) * p.res_ld + col0 +
                                                     q * 8);
        }
      }
    } else {
      if (!er.store_ok) return;
      const bf16* rp = p.residual + static_cast<size_t>(er.res_row) * p.res_ld + col0;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (col0 + g * 8 < p.N) res[g] = *reinterpret_cast<const uint4*>(rp + g * 8);
      }
    }
  }
}

// ---- bias fetch of one chunk (32 columns, the same for every row), issued one chunk ahead like the residual: the
//      ncu source view of the K = 1024 ViT fc1 GEMM had 55 % of all stall samples on the first use of these loads
//      (the 227 KB shared-memory carve-out leaves almost no L1, so each is an L2 round trip) ----
SLIME_DEVINL void epi_issue_bias(const GemmParams& p, int col0, uint4 (&b)[4]) {
  if (p.bias == nullptr) return;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (col0 + g * 8 < p.N) b[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + col0 + g * 8));
  }
}

// bias / activation / residual on 8 accumulator columns -> 8 floats
template <int EPI>
SLIME_DEVINL void epi_math8(const GemmParams& p, const EpiRow& er, int col, const uint32_t* r8, const uint4& resq,
                            const uint4& b, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r8[j]) * er.scale;
  if constexpr (EPI == GEMM_EPI_ROPE) {
    // columns (2i, 2i+1) of a head hold features (i, i + half): out_i = x_i cos_i - x_{i+half} sin_i,
    // out_{i+half} = x_{i+half} cos_i + x_i sin_i, on the fp32 accumulators (one rounding instead of two)
    if (col < p.rope_cols) {
      const float4* cs = reinterpret_cast<const float4*>(p.rope_table + static_cast<size_t>(er.pos) * p.rope_half +
                                                         ((col & (2 * p.rope_half - 1)) >> 1));
      const float4 t0 = __ldg(cs), t1 = __ldg(cs + 1);
      const float c[4] = {t0.x, t0.z, t1.x, t1.z}, sn[4] = {t0.y, t0.w, t1.y, t1.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float lo = v[2 * j], hi = v[2 * j + 1];
        v[2 * j] = lo * c[j] - hi * sn[j];
        v[2 * j + 1] = hi * c[j] + lo * sn[j];
      }
    }
  }
  if (p.bias != nullptr) {
    const float2 b0 = unpack_bf16x2(b.x), b1 = unpack_bf16x2(b.y), b2 = unpack_bf16x2(b.z), b3 = unpack_bf16x2(b.w);
    v[0] += b0.x; v[1] += b0.y; v[2] += b1.x; v[3] += b1.y;
    v[4] += b2.x; v[5] += b2.y; v[6] += b3.x; v[7] += b3.y;
  }
  if constexpr (EPI == GEMM_EPI_QUICK_GELU) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_quick_gelu(v[j]);
  } else if constexpr (EPI == GEMM_EPI_GELU_ERF) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = act_gelu_erf(v[j]);
  }
  if (p.residual != nullptr) {
    const float2 q0 = unpack_bf16x2(resq.x), q1 = unpack_bf16x2(resq.y), q2 = unpack_bf16x2(resq.z),
                 q3 = unpack_bf16x2(resq.w);
    v[0] += q0.x; v[1] += q0.y; v[2] += q1.x; v[3] += q1.y;
    v[4] += q2.x; v[5] += q2.y; v[6] += q3.x; v[7] += q3.y;
  }
}

SLIME_DEVINL uint4 epi_pack8(const float (&v)[8]) {
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]);
  pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]);
  pk.w = pack_bf16x2(v[6], v[7]);
  return pk;
}

// SwiGLU on 16 interleaved (gate, up) accumulator columns -> 8 outputs
SLIME_DEVINL uint4 epi_swiglu8(const uint32_t* r16, float scale) {
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    o[j] = act_silu(__uint_as_float(r16[2 * j]) * scale) * (__uint_as_float(r16[2 * j + 1]) * scale);
  return epi_pack8(o);
}
// sum of squares of 8 packed 16-bit outputs (the values the next layer will actually read)
SLIME_DEVINL float epi_sumsq8(const uint4& pk) {
  const float2 a = unpack_bf16x2(pk.x), b = unpack_bf16x2(pk.y), c = unpack_bf16x2(pk.z), d = unpack_bf16x2(pk.w);
  return (a.x * a.x + a.y * a.y) + (b.x * b.x + b.y * b.y) + (c.x * c.x + c.y * c.y) + (d.x * d.x + d.y * d.y);
}

// 16-byte output store with an L2 eviction policy (GemmParams::store_hint)
SLIME_DEVINL void epi_store16(void* dst, const uint4& v, int hint) {
  if (hint == 0) {
    *reinterpret_cast<uint4*>(dst) = v;
  } else {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;\n" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w), "l"(pol)
                 : "memory");
  }
}

// ---- mode 0: every thread stores its own row ----
template <int EPI>
SLIME_DEVINL void epi_process_chunk_direct(const GemmParams& p, EpiRow& er, int col0, const uint32_t (&r)[32],
                                           const uint4 (&res)[4], const uint4 (&bia)[4]) {
  if (!er.store_ok || col0 >= p.N) return;
  if constexpr (EPI == GEMM_EPI_SWIGLU) {
    // columns are (gate_j, up_j) interleaved -> 16 outputs per 32 accumulator columns
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const int col = col0 + g * 16;
      if (col >= p.N) break;
      epi_store16(p.out + static_cast<size_t>(er.out_row) * p.out_ld + (col >> 1), epi_swiglu8(r + g * 16, er.scale),
                  p.store_hint);
    }
  } else {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = col0 + g * 8;
      if (col >= p.N) break;
      float v[8];
      epi_math8<EPI>(p, er, col, r + g * 8, res[g], bia[g], v);
      if (p.out_f32 != nullptr) {
        float4* dst = reinterpret_cast<float4*>(p.out_f32 + static_cast<size_t>(er.out_row) * p.out_ld + col);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      } else {
        const uint4 pk = epi_pack8(v);
        epi_store16(p.out + static_cast<size_t>(er.out_row) * p.out_ld + col, pk, p.store_hint);
        if (p.sumsq_out != nullptr) er.ssq += epi_sumsq8(pk);
      }
    }
  }
}

// ---- modes 1 / 2: transpose through the warp's staging tile, coalesced stores (and residual loads) ----
template <int EPI>
SLIME_DEVINL void epi_process_chunk_staged(const GemmParams& p, const EpiRow& er, const EpiCoal& ec, int lane, int col0,
                                           const uint32_t (&r)[32], const uint4 (&res)[4], const uint4 (&bia)[4],
                                           uint8_t* stage) {
  if (col0 >= p.N) return;  // warp-uniform
  constexpr int SLOTS = EPI == GEMM_EPI_SWIGLU ? 2 : 4;
  uint8_t* stage_out = stage;
  if constexpr (EPI == GEMM_EPI_SWIGLU) {
#pragma unroll
    for (int g = 0; g < 2; ++g)
      *reinterpret_cast<uint4*>(stage_out + epi_stage_off<2>(lane, g)) = epi_swiglu8(r + g * 16, er.scale);
  } else {
    uint4 own[4];
    if (p.residual != nullptr) {
      {
        uint8_t* stage_res = stage + 2048;
        const int q = lane & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(stage_res + epi_stage_off<4>(i * 8 + (lane >> 2), q)) = res[i];
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 4; ++g) own[g] = *reinterpret_cast<const uint4*>(stage_res + epi_stage_off<4>(lane, g));
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int col = col0 + g * 8;
      float v[8];
      if (col < p.N) {
        epi_math8<EPI>(p, er, col, r + g * 8, own[g], bia[g], v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      *reinterpret_cast<uint4*>(stage_out + epi_stage_off<4>(lane, g)) = epi_pack8(v);
    }
  }
  __syncwarp();
  const int q = lane % SLOTS;
  const int out_col = (EPI == GEMM_EPI_SWIGLU ? (col0 >> 1) : col0) + q * 8;
  const bool col_ok = (EPI == GEMM_EPI_SWIGLU ? col0 + q * 16 : col0 + q * 8) < p.N;
#pragma unroll
  for (int i = 0; i < SLOTS; ++i) {
    const int rr = i * (32 / SLOTS) + lane / SLOTS;
    const uint4 v = *reinterpret_cast<const uint4*>(stage_out + epi_stage_off<SLOTS>(rr, q));
    if (col_ok && (ec.ok & (1u << i)))
      epi_store16(p.out + static_cast<size_t>(ec.out_row[i]) * p.out_ld + out_col, v, p.store_hint);
  }
  __syncwarp();
}

// One accumulator tile (128 x BLOCK_N fp32 in TMEM at column tmem_acc) -> HBM.
//   quad : TMEM lane quadrant of this warp;  half : which half of the columns this warp covers
//   stage: this warp's EPI_STAGE_BYTES of shared memory
template <int BLOCK_N, int EPI, bool STAGED>
SLIME_DEVINL void epilogue_tile(const GemmParams& p, uint32_t tmem_acc, int m0, int n0, int quad, int half, int lane,
                                uint8_t* stage) {
  // 32-column chunks per warp: the two warps of a lane quadrant split the tile's columns in halves; a 192-wide tile is split
  // 128 + 64 so that every warp owns whole 64-column groups (the sum-of-squares partials are per 64 columns)
  constexpr int NCH = BLOCK_N == 192 ? 4 : BLOCK_N / 64;
  const int nch = (BLOCK_N == 192 && half == 1) ? 2 : NCH;
  const int col_off = BLOCK_N == 192 ? half * 128 : half * (BLOCK_N / 2);
  constexpr int SLOTS = EPI == GEMM_EPI_SWIGLU ? 2 : 4;
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
  er.scale = (p.row_scale != nullptr && row_ok) ? p.row_scale[row] : 1.0f;
  er.ssq = 0.f;
  EpiCoal ec = {};
  if constexpr (STAGED) ec = epi_make_coal<SLOTS>(er, lane);

  const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) + col_off;
  const int col_begin = n0 + col_off;

  uint32_t acc[2][32];
  uint4 res[2][4];
  uint4 bia[2][4];
  tmem_ld_32x32b_x32(taddr, acc[0]);
  if constexpr (EPI != GEMM_EPI_SWIGLU) epi_issue_bias(p, col_begin, bia[0]);
  epi_issue_residual<EPI, STAGED>(p, er, ec, lane, col_begin, res[0]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    if (i >= nch) break;
    if (i + 1 < nch) {
      tmem_ld_32x32b_x32(taddr + (i + 1) * 32, acc[(i + 1) & 1]);
      if constexpr (EPI != GEMM_EPI_SWIGLU) epi_issue_bias(p, col_begin + (i + 1) * 32, bia[(i + 1) & 1]);
      epi_issue_residual<EPI, STAGED>(p, er, ec, lane, col_begin + (i + 1) * 32, res[(i + 1) & 1]);
    }
    if constexpr (STAGED) {
      epi_process_chunk_staged<EPI>(p, er, ec, lane, col_begin + i * 32, acc[i & 1], res[i & 1], bia[i & 1], stage);
    } else {
      epi_process_chunk_direct<EPI>(p, er, col_begin + i * 32, acc[i & 1], res[i & 1], bia[i & 1]);
      if constexpr (EPI == GEMM_EPI_NONE) {
        // one partial per 64 output columns (two chunks): [row][column / 64]
        if (p.sumsq_out != nullptr && (i & 1) == 1) {
          const int part = (col_begin + i * 32) >> 6;
          if (er.store_ok && part < p.sumsq_parts) p.sumsq_out[static_cast<size_t>(row) * p.sumsq_parts + part] = er.ssq;
          er.ssq = 0.f;
        }
      }
    }
    if (i + 1 < nch) tmem_ld_wait();
  }
}
