// Common device/host helpers for the slime_b200 sm_100a kernels.
// Raw PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (UMMA + TMEM).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda.h>
#include <stdint.h>

// `bf16` is the library's 16-bit element type.  The default build computes in bfloat16 (libslime_b200.so); the
// same sources compiled with -DSLIME_FP16 compute in IEEE half (libslime_b200_fp16.so) - the reference's
// inference dtype (llava/model/builder.py:43).  Only the conversions, the UMMA operand-format bits, the TMA
// element type and the mma.sync type suffix differ; tiles, pipelines and fp32 accumulation are identical.
#ifdef SLIME_FP16
typedef __half bf16;
typedef __half2 bf162;
#define SLIME_ELEM_DTYPE 2                               /* slime_elem_dtype(): 0 bf16, 2 fp16 */
#define SLIME_UMMA_AB_FORMAT 0u                          /* kind::f16 operand format: 0 = F16, 1 = BF16 */
#define SLIME_TMAP_ELEM CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define SLIME_MMA_SYNC_TYPE "f16"
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;
#define SLIME_ELEM_DTYPE 0
#define SLIME_UMMA_AB_FORMAT 1u
#define SLIME_TMAP_ELEM CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define SLIME_MMA_SYNC_TYPE "bf16"
#endif

#define SLIME_DEVINL __device__ __forceinline__

// ----------------------------------------------------------------------------------------
// shared-memory address helpers
// ----------------------------------------------------------------------------------------
SLIME_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

SLIME_DEVINL uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      ".reg .b32 R1;\n\t"
      "elect.sync R1|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
SLIME_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
SLIME_DEVINL void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
SLIME_DEVINL void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
SLIME_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
SLIME_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
SLIME_DEVINL uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
SLIME_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ----------------------------------------------------------------------------------------
// TMA
// ----------------------------------------------------------------------------------------
SLIME_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled load: coordinates (c0 = innermost element index, c1 = row index)
SLIME_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
SLIME_DEVINL void tma_load_2d_hint(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                   int c1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
SLIME_DEVINL void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
SLIME_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
template <int N>
SLIME_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory");
}
template <int N>
SLIME_DEVINL void tma_store_wait() {
  asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
template <int NCOLS>
SLIME_DEVINL void tmem_alloc(uint32_t* smem_holder) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_holder)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <int NCOLS>
SLIME_DEVINL void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(NCOLS)
               : "memory");
}
SLIME_DEVINL void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
}
SLIME_DEVINL void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]; kind::f16 covers bf16/fp16 operands, fp32 accumulate.
SLIME_DEVINL void umma_bf16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
SLIME_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// TMEM -> registers: each thread of the warp reads its own lane (row), 32 consecutive columns.
SLIME_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
SLIME_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// UMMA shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes
// (64 bf16) with the 128-byte hardware swizzle (what TMA SWIZZLE_128B writes):
//   start address  bits [0,14)   (byte address >> 4)
//   LBO            bits [16,30)  (ignored for swizzled K-major; canonical value 1)
//   SBO            bits [32,46)  (8 rows * 128 B = 1024 B between 8-row core-matrix groups) >> 4
//   version        bits [46,48)  = 1 on sm_100
//   layout type    bits [61,64)  = 2 (SWIZZLE_128B)
SLIME_DEVINL uint64_t make_umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor: bf16 x bf16 -> fp32, both operands K-major.
//   c_format [4,6)=1 (F32), a_format [7,10)=1 (BF16), b_format [10,13)=1 (BF16),
//   a_major bit15 = 0, b_major bit16 = 0, n_dim [17,23) = N>>3, m_dim [24,29) = M>>4.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
  return (1u << 4) | (SLIME_UMMA_AB_FORMAT << 7) | (SLIME_UMMA_AB_FORMAT << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// ----------------------------------------------------------------------------------------
// small math helpers
// ----------------------------------------------------------------------------------------
#ifdef SLIME_FP16
SLIME_DEVINL float elem_to_float(bf16 x) { return __half2float(x); }
SLIME_DEVINL bf16 float_to_elem(float x) { return __float2half_rn(x); }
SLIME_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  bf162 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
SLIME_DEVINL float2 unpack_bf16x2(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __half22float2(v);
}
#else
SLIME_DEVINL float elem_to_float(bf16 x) { return __bfloat162float(x); }
SLIME_DEVINL bf16 float_to_elem(float x) { return __float2bfloat16_rn(x); }
SLIME_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  bf162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
SLIME_DEVINL float2 unpack_bf16x2(uint32_t u) {
  bf162 v = *reinterpret_cast<bf162*>(&u);
  return __bfloat1622float2(v);
}
#endif
// ---- programmatic dependent launch (PDL): a kernel launched with the attribute may become resident while its
// predecessor in the stream is still running; it must execute pdl_wait() before touching anything the predecessor
// (or, transitively, anything earlier) writes or reads-then-overwrites.  What comes BEFORE the wait - in the decode
// GEMM the first 8 KB per warp of the (immutable) weight stream - overlaps the predecessor's tail and the launch
// latency.  pdl_trigger() lets the NEXT kernel in the stream do the same with this one.  Both are no-ops for a
// kernel launched without the attribute.
SLIME_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
SLIME_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// Fire-and-forget L2 prefetch of part `part` of `nparts` of [ptr, ptr + bytes) (16-byte aligned), 8 KB per request,
// spread over the threads of the calling CTA.  Used by the latency-bound kernels of the decode step to keep HBM busy
// with the NEXT projection's weights while they run (the weights are immutable, so no ordering is needed).
SLIME_DEVINL void l2_prefetch_slice(const void* ptr, size_t bytes, int part, int nparts) {
  constexpr size_t CH = 8192;
  const size_t nch = (bytes + CH - 1) / CH;
  for (size_t i = static_cast<size_t>(part) * blockDim.x + threadIdx.x; i < nch; i += static_cast<size_t>(nparts) * blockDim.x) {
    const char* a = static_cast<const char*>(ptr) + i * CH;
    const size_t left = bytes - i * CH;
    const uint32_t sz = static_cast<uint32_t>((left < CH ? left : CH) & ~static_cast<size_t>(15));
    if (sz != 0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(a), "r"(sz) : "memory");
  }
}

SLIME_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
SLIME_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ----------------------------------------------------------------------------------------
// additional tcgen05 helpers (attention kernel): TMEM store, A-from-TMEM MMA, MN-major descriptors
// ----------------------------------------------------------------------------------------
// registers -> TMEM: each thread of the warp writes its own lane (row), 32 consecutive 32-bit columns.
SLIME_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
      "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
      "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
// registers -> TMEM, 16 consecutive 32-bit columns of the thread's own lane
SLIME_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
      "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
SLIME_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]: A is read from tensor memory (row = lane, two bf16 per 32-bit column).
SLIME_DEVINL void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory descriptor of an MN-major operand stored as [K rows][64 MN elements] slabs of 128-byte rows
// with the 128-byte swizzle (a TMA box {64, rows}):  MN blocks of 64 elements are `lbo_bytes` apart (the next
// 64-column slab), groups of 8 K rows are 1024 bytes apart (SBO).
SLIME_DEVINL uint64_t make_umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor with explicit operand majors (0 = K-major, 1 = MN-major).
__host__ __device__ constexpr uint32_t make_idesc_bf16_major(int M, int N, int a_mn, int b_mn) {
  return make_idesc_bf16(M, N) | (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16);
}
