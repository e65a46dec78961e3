// Image pre-processing on the GPU: the step right before the hot path.
//
// Replaces, bit for bit, what the reference does on DataLoader workers with PIL + CLIPImageProcessor
// (llava/mm_utils.py:99-153,177-210,231-259): `Image.resize` (Pillow's antialiased bicubic: two 8-bit passes with
// 22-bit fixed-point coefficients, Resample.c), paste on a black canvas / centre crop, cut into 336 px tiles,
// rescale + normalise (a 3 x 256 float table built by the caller) and channels-first layout in the model dtype.
//
// Work split: the HOST computes the coefficient tables in double precision exactly as Pillow does (they depend
// only on the sizes - a few KB per job) and ships them with the job descriptors in ONE copy; the DEVICE does
// the byte work in two launches for the whole batch of images:
//   resize_rows_kernel : source bytes -> 8-bit intermediate [virt_h, out_w, 3]     (horizontal pass)
//   resize_cols_kernel : intermediate -> canvas pixel -> table lookup -> crops [n, 3, 336, 336] (vertical pass)
// Both are integer kernels bound by load/store issue, not by HBM (algorithmic bytes = source image + output crops):
// the interleaved RGB bytes are therefore fetched as aligned 32-bit words (four pixels = three words per tap,
// re-aligned with a funnel shift), coefficients as int4, and the vertical pass produces four pixels per thread and
// stores 8 bytes per channel.  Taps whose coefficient rounded to zero are trimmed on the host and a pass that does
// not change the size is skipped, as in Pillow.
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "errors.h"

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;  // Resample.c: 8-bit pixels, 2 bits of head-room

struct DevJob {
  long long src_off;  // byte offset of the image in the source buffer (multiple of 4)
  long long tmp_off;  // byte offset of this job's intermediate inside the workspace (multiple of 256)
  int src_w, src_h;
  int virt_w, virt_h, virt_x, virt_y;
  int out_w, out_h;
  int canvas_w, canvas_h;
  int paste_x, paste_y;
  int first_crop;
  int ksize_h, ksize_v;                // padded to multiples of 4 (int4 loads); the padding is zero
  int hb_off, hk_off, vb_off, vk_off;  // int32 offsets into the table area: bounds (2 per output) and coefficients
  int skip_rows_pass;                  // width unchanged and no padding: the vertical pass reads the source itself
  int padded_source;                   // expand2square: the horizontal pass needs the per-pixel inside test
  unsigned char fill[4];
};

// Pillow's bicubic kernel, a = -0.5 (Resample.c bicubic_filter); evaluated in the same operation order.
double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

int coeff_ksize(int in_size, int out_size) {
  double filterscale = static_cast<double>(in_size) / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  return static_cast<int>(std::ceil(support)) * 2 + 1;
}
int coeff_ksize_padded(int in_size, int out_size) { return (coeff_ksize(in_size, out_size) + 3) / 4 * 4; }

// Resample.c precompute_coeffs (box = whole image) + normalize_coeffs_8bpc.
// bounds: 2 ints per output (first tap, tap count); kk: ksize ints per output, zero beyond the tap count.
void precompute_coeffs(int in_size, int out_size, int ksize, int32_t* bounds, int32_t* kk) {
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  const double ss = 1.0 / filterscale;
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      const double v = bicubic_filter((x + xmin - center + 0.5) * ss);
      w[x] = v;
      ww += v;
    }
    int32_t* k = kk + static_cast<size_t>(xx) * ksize;
    for (int x = 0; x < xmax; ++x) {
      double v = w[x];
      if (ww != 0.0) v /= ww;
      k[x] = (v < 0) ? static_cast<int32_t>(-0.5 + v * (1 << PRECISION_BITS))
                     : static_cast<int32_t>(0.5 + v * (1 << PRECISION_BITS));
    }
    // taps whose fixed-point coefficient rounded to zero contribute nothing: drop them from both ends (the
    // identity resize collapses to one tap; Pillow keeps them, the sums are identical)
    int first = 0;
    while (xmax - first > 1 && k[first] == 0) ++first;
    while (xmax - first > 1 && k[xmax - 1] == 0) --xmax;
    if (first > 0) {
      for (int x = first; x < xmax; ++x) k[x - first] = k[x];
      xmin += first;
      xmax -= first;
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
}

__device__ __forceinline__ int clip8(int acc) {
  const int v = acc >> PRECISION_BITS;  // arithmetic shift, like the C reference
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// 12 consecutive bytes (four RGB pixels) starting at byte `off` of the 4-byte-aligned array `base`, as three words.
// `off` may be unaligned, negative or past the end: word indices are clamped to [0, last_word], and whatever is
// fetched from a clamped position is either multiplied by a zero coefficient or masked by the caller.
__device__ __forceinline__ void load12(const unsigned char* base, long long off, long long last_word, uint32_t (&b)[3]) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(base);
  const long long w0 = off >> 2;
  const unsigned sh = (static_cast<unsigned>(off) & 3u) * 8u;
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long wi = w0 + i;
    wi = wi < 0 ? 0 : (wi > last_word ? last_word : wi);
    w[i] = __ldg(wp + wi);
  }
  b[0] = __funnelshift_r(w[0], w[1], sh);
  b[1] = __funnelshift_r(w[1], w[2], sh);
  b[2] = __funnelshift_r(w[2], w[3], sh);
}
__device__ __forceinline__ int byte_of(const uint32_t (&b)[3], int j) { return (b[j >> 2] >> ((j & 3) * 8)) & 255; }

// horizontal pass, plain source: one thread per (source row, output column), four taps per iteration
__global__ void __launch_bounds__(256) resize_rows_kernel(const unsigned char* __restrict__ src,
                                                          const DevJob* __restrict__ jobs,
                                                          const int32_t* __restrict__ tables,
                                                          unsigned char* __restrict__ ws) {
  const DevJob jb = jobs[blockIdx.y];
  if (jb.skip_rows_pass || jb.padded_source) return;  // ImagingResample: need_horizontal == false / generic kernel
  const long long total = static_cast<long long>(jb.src_h) * jb.out_w;
  const unsigned char* img = src + jb.src_off;
  unsigned char* tmp = ws + jb.tmp_off;
  const int32_t* bounds = tables + jb.hb_off;
  const int32_t* kk = tables + jb.hk_off;
  const long long last_word = (static_cast<long long>(jb.src_w) * jb.src_h * 3 + 3) / 4 - 1;
  const long long row_bytes = static_cast<long long>(jb.src_w) * 3;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int y = static_cast<int>(i / jb.out_w);
    const int xo = static_cast<int>(i - static_cast<long long>(y) * jb.out_w);
    const int xmin = bounds[2 * xo], n = bounds[2 * xo + 1];
    const int4* k4 = reinterpret_cast<const int4*>(kk + static_cast<size_t>(xo) * jb.ksize_h);
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    long long off = y * row_bytes + static_cast<long long>(xmin) * 3;
    for (int t = 0; t < n; t += 4) {
      uint32_t b[3];
      load12(img, off, last_word, b);
      const int4 c = __ldg(k4 + (t >> 2));  // zero beyond the tap count: bytes past the window do not matter
      s0 += byte_of(b, 0) * c.x + byte_of(b, 3) * c.y + byte_of(b, 6) * c.z + byte_of(b, 9) * c.w;
      s1 += byte_of(b, 1) * c.x + byte_of(b, 4) * c.y + byte_of(b, 7) * c.z + byte_of(b, 10) * c.w;
      s2 += byte_of(b, 2) * c.x + byte_of(b, 5) * c.y + byte_of(b, 8) * c.z + byte_of(b, 11) * c.w;
      off += 12;
    }
    unsigned char* o = tmp + static_cast<size_t>(i) * 3;
    o[0] = static_cast<unsigned char>(clip8(s0));
    o[1] = static_cast<unsigned char>(clip8(s1));
    o[2] = static_cast<unsigned char>(clip8(s2));
  }
}

// horizontal pass, padded ("virtual") source of expand2square: per-pixel inside test, byte loads
__global__ void __launch_bounds__(256) resize_rows_padded_kernel(const unsigned char* __restrict__ src,
                                                                 const DevJob* __restrict__ jobs,
                                                                 const int32_t* __restrict__ tables,
                                                                 unsigned char* __restrict__ ws) {
  const DevJob jb = jobs[blockIdx.y];
  if (jb.skip_rows_pass || !jb.padded_source) return;
  const long long total = static_cast<long long>(jb.virt_h) * jb.out_w;
  const unsigned char* img = src + jb.src_off;
  unsigned char* tmp = ws + jb.tmp_off;
  const int32_t* bounds = tables + jb.hb_off;
  const int32_t* kk = tables + jb.hk_off;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int y = static_cast<int>(i / jb.out_w);
    const int xo = static_cast<int>(i - static_cast<long long>(y) * jb.out_w);
    const int xmin = bounds[2 * xo], n = bounds[2 * xo + 1];
    const int32_t* k = kk + static_cast<size_t>(xo) * jb.ksize_h;
    int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
    const int sy = y - jb.virt_y;
    const bool row_in = sy >= 0 && sy < jb.src_h;
    const unsigned char* row = img + static_cast<size_t>(row_in ? sy : 0) * jb.src_w * 3;
    for (int t = 0; t < n; ++t) {
      const int sx = xmin + t - jb.virt_x;
      int p0 = jb.fill[0], p1 = jb.fill[1], p2 = jb.fill[2];
      if (row_in && sx >= 0 && sx < jb.src_w) {
        p0 = row[sx * 3];
        p1 = row[sx * 3 + 1];
        p2 = row[sx * 3 + 2];
      }
      const int c = k[t];
      s0 += p0 * c;
      s1 += p1 * c;
      s2 += p2 * c;
    }
    unsigned char* o = tmp + static_cast<size_t>(i) * 3;
    o[0] = static_cast<unsigned char>(clip8(s0));
    o[1] = static_cast<unsigned char>(clip8(s1));
    o[2] = static_cast<unsigned char>(clip8(s2));
  }
}

template <typename T>
__device__ __forceinline__ void store4(T* p, float a, float b, float c, float d);
template <>
__device__ __forceinline__ void store4<float>(float* p, float a, float b, float c, float d) {
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
template <>
__device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&lo);
  v.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}
template <>
__device__ __forceinline__ void store4<__half>(__half* p, float a, float b, float c, float d) {
  __half2 lo = __floats2half2_rn(a, b), hi = __floats2half2_rn(c, d);
  uint2 v;
  v.x = *reinterpret_cast<uint32_t*>(&lo);
  v.y = *reinterpret_cast<uint32_t*>(&hi);
  *reinterpret_cast<uint2*>(p) = v;
}

// vertical pass + canvas placement + table lookup: one thread per FOUR horizontally adjacent canvas pixels
// (crop % 4 == 0, so the four share a tile and an aligned 8-byte / 16-byte store per channel)
template <typename T>
__global__ void __launch_bounds__(256) resize_cols_kernel(const unsigned char* __restrict__ src,
                                                          const DevJob* __restrict__ jobs,
                                                          const int32_t* __restrict__ tables,
                                                          const float* __restrict__ lut,
                                                          const unsigned char* __restrict__ ws, T* __restrict__ out,
                                                          int crop) {
  const DevJob jb = jobs[blockIdx.y];
  const int groups_x = jb.canvas_w >> 2;
  const long long total = static_cast<long long>(groups_x) * jb.canvas_h;
  // intermediate [virt_h, out_w, 3], or the source itself when the horizontal pass was skipped (same shape)
  const unsigned char* tmp = jb.skip_rows_pass ? src + jb.src_off : ws + jb.tmp_off;
  const long long row_bytes = static_cast<long long>(jb.out_w) * 3;
  const long long last_word = (row_bytes * jb.virt_h + 3) / 4 - 1;
  const int32_t* bounds = tables + jb.vb_off;
  const int32_t* kk = tables + jb.vk_off;
  const int tiles_x = jb.canvas_w / crop;
  const size_t plane = static_cast<size_t>(crop) * crop;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int cy = static_cast<int>(i / groups_x);
    const int cx = static_cast<int>(i - static_cast<long long>(cy) * groups_x) * 4;
    const int rx = cx - jb.paste_x, ry = cy - jb.paste_y;
    const bool row_ok = ry >= 0 && ry < jb.out_h;
    int acc[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) acc[j] = 1 << (PRECISION_BITS - 1);
    if (row_ok && rx + 3 >= 0 && rx < jb.out_w) {
      const int ymin = bounds[2 * ry], n = bounds[2 * ry + 1];
      const int32_t* k = kk + static_cast<size_t>(ry) * jb.ksize_v;
      long long off = ymin * row_bytes + static_cast<long long>(rx) * 3;
      for (int t = 0; t < n; ++t) {
        uint32_t b[3];
        load12(tmp, off, last_word, b);
        const int c = __ldg(k + t);
#pragma unroll
        for (int j = 0; j < 12; ++j) acc[j] += byte_of(b, j) * c;
        off += row_bytes;
      }
    }
    float v[12];
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      const bool ok = row_ok && rx + px >= 0 && rx + px < jb.out_w;  // else Image.new('RGB', ..., (0, 0, 0))
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c * 4 + px] = __ldg(lut + c * 256 + (ok ? clip8(acc[px * 3 + c]) : 0));
    }
    const int tile = (cy / crop) * tiles_x + cx / crop;
    const int ty = cy % crop, tx = cx % crop;
    T* o = out + (static_cast<size_t>(jb.first_crop + tile) * 3) * plane + static_cast<size_t>(ty) * crop + tx;
#pragma unroll
    for (int c = 0; c < 3; ++c) store4<T>(o + c * plane, v[c * 4], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]);
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
  size_t jobs_bytes, lut_off, tables_off, tables_ints, tmp_off, total;
  std::vector<long long> tmp_offs;
};

int validate_jobs(const slime_resize_job* jobs, int n_jobs, int crop) {
  SLIME_REQUIRE(jobs != nullptr && n_jobs > 0, "preprocess: no jobs");
  SLIME_REQUIRE(crop > 0 && crop % 4 == 0, "preprocess: crop size %d must be a positive multiple of 4", crop);
  for (int j = 0; j < n_jobs; ++j) {
    const slime_resize_job& q = jobs[j];
    SLIME_REQUIRE(q.src_w > 0 && q.src_h > 0, "preprocess job %d: empty source %dx%d", j, q.src_w, q.src_h);
    SLIME_REQUIRE(q.virt_w >= q.src_w + q.virt_x && q.virt_h >= q.src_h + q.virt_y && q.virt_x >= 0 && q.virt_y >= 0,
                  "preprocess job %d: the %dx%d image at (%d,%d) does not fit its %dx%d padded source", j, q.src_w,
                  q.src_h, q.virt_x, q.virt_y, q.virt_w, q.virt_h);
    SLIME_REQUIRE(q.out_w > 0 && q.out_h > 0, "preprocess job %d: empty resize target", j);
    SLIME_REQUIRE(q.canvas_w > 0 && q.canvas_h > 0 && q.canvas_w % crop == 0 && q.canvas_h % crop == 0,
                  "preprocess job %d: canvas %dx%d is not a multiple of the %d px tile", j, q.canvas_w, q.canvas_h,
                  crop);
    SLIME_REQUIRE(q.first_crop >= 0 && q.src_offset >= 0 && q.src_offset % 4 == 0,
                  "preprocess job %d: offsets must be non-negative and images must start on a 4-byte boundary", j);
  }
  return SLIME_OK;
}

// ints of one table block: bounds, then the coefficient rows on a 16-byte boundary
size_t table_ints(int in_size, int out_size) {
  return align_up(2 * static_cast<size_t>(out_size), 4) + static_cast<size_t>(out_size) * coeff_ksize_padded(in_size, out_size);
}

Layout make_layout(const slime_resize_job* jobs, int n_jobs) {
  Layout L;
  L.jobs_bytes = align_up(sizeof(DevJob) * n_jobs, 256);
  L.lut_off = L.jobs_bytes;
  L.tables_off = L.lut_off + 3 * 256 * sizeof(float);
  size_t ints = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const slime_resize_job& q = jobs[j];
    ints += table_ints(q.virt_w, q.out_w) + table_ints(q.virt_h, q.out_h);
  }
  L.tables_ints = ints;
  L.tmp_off = align_up(L.tables_off + ints * sizeof(int32_t), 256);
  size_t off = L.tmp_off;
  for (int j = 0; j < n_jobs; ++j) {
    L.tmp_offs.push_back(static_cast<long long>(off));
    off += align_up(static_cast<size_t>(jobs[j].virt_h) * jobs[j].out_w * 3, 256);
  }
  L.total = off;
  return L;
}

}  // namespace

extern "C" size_t slime_preprocess_workspace_bytes(const slime_resize_job* jobs, int n_jobs) {
  if (jobs == nullptr || n_jobs <= 0) return 0;
  return make_layout(jobs, n_jobs).total;
}

extern "C" int slime_preprocess_fwd(const uint8_t* src, const slime_resize_job* jobs, int n_jobs, int crop,
                                    const float* lut, void* out, int out_dtype, void* ws, size_t ws_bytes,
                                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  SLIME_PROPAGATE(validate_jobs(jobs, n_jobs, crop));
  SLIME_REQUIRE(src != nullptr && lut != nullptr && out != nullptr && ws != nullptr, "preprocess: null pointer");
  SLIME_REQUIRE(reinterpret_cast<uintptr_t>(src) % 4 == 0 && reinterpret_cast<uintptr_t>(ws) % 256 == 0 &&
                    reinterpret_cast<uintptr_t>(out) % 16 == 0,
                "preprocess: src must be 4-byte, out 16-byte and ws 256-byte aligned");
  SLIME_REQUIRE(out_dtype >= 0 && out_dtype <= 2, "preprocess: out_dtype %d (0 bf16, 1 fp32, 2 fp16)", out_dtype);
  const Layout L = make_layout(jobs, n_jobs);
  if (ws_bytes < L.total) {
    slime_set_error("preprocess: workspace %zu < %zu bytes", ws_bytes, L.total);
    return SLIME_EWORKSPACE;
  }
  // host staging: [DevJob x n][lut][tables]; one copy
  std::vector<unsigned char> stage(L.tmp_off, 0);
  DevJob* dj = reinterpret_cast<DevJob*>(stage.data());
  std::memcpy(stage.data() + L.lut_off, lut, 3 * 256 * sizeof(float));
  int32_t* tables = reinterpret_cast<int32_t*>(stage.data() + L.tables_off);
  size_t cursor = 0;
  long long max_rows_px = 0, max_canvas_groups = 0;
  bool any_padded = false, any_plain = false;
  double bytes = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const slime_resize_job& q = jobs[j];
    DevJob& d = dj[j];
    d.src_off = q.src_offset;
    d.tmp_off = L.tmp_offs[j];
    d.src_w = q.src_w; d.src_h = q.src_h;
    d.virt_w = q.virt_w; d.virt_h = q.virt_h; d.virt_x = q.virt_x; d.virt_y = q.virt_y;
    d.out_w = q.out_w; d.out_h = q.out_h;
    d.canvas_w = q.canvas_w; d.canvas_h = q.canvas_h;
    d.paste_x = q.paste_x; d.paste_y = q.paste_y;
    d.first_crop = q.first_crop;
    std::memcpy(d.fill, q.fill, 4);
    d.padded_source = (q.virt_w != q.src_w || q.virt_h != q.src_h) ? 1 : 0;
    d.skip_rows_pass = (q.out_w == q.virt_w && !d.padded_source) ? 1 : 0;
    d.ksize_h = coeff_ksize_padded(q.virt_w, q.out_w);
    d.ksize_v = coeff_ksize_padded(q.virt_h, q.out_h);
    d.hb_off = static_cast<int>(cursor); cursor += align_up(2 * static_cast<size_t>(q.out_w), 4);
    d.hk_off = static_cast<int>(cursor); cursor += static_cast<size_t>(q.out_w) * d.ksize_h;
    d.vb_off = static_cast<int>(cursor); cursor += align_up(2 * static_cast<size_t>(q.out_h), 4);
    d.vk_off = static_cast<int>(cursor); cursor += static_cast<size_t>(q.out_h) * d.ksize_v;
    precompute_coeffs(q.virt_w, q.out_w, d.ksize_h, tables + d.hb_off, tables + d.hk_off);
    precompute_coeffs(q.virt_h, q.out_h, d.ksize_v, tables + d.vb_off, tables + d.vk_off);
    const long long rows_px = d.skip_rows_pass ? 0 : static_cast<long long>(q.virt_h) * q.out_w;
    const long long groups = static_cast<long long>(q.canvas_w / 4) * q.canvas_h;
    if (rows_px > max_rows_px) max_rows_px = rows_px;
    if (groups > max_canvas_groups) max_canvas_groups = groups;
    if (!d.skip_rows_pass) (d.padded_source ? any_padded : any_plain) = true;
    bytes += 3.0 * q.src_w * q.src_h + 3.0 * q.canvas_w * q.canvas_h * (out_dtype == 1 ? 4 : 2);
  }
  SLIME_CHECK_CUDA(cudaMemcpyAsync(ws, stage.data(), L.tmp_off, cudaMemcpyHostToDevice, stream));
  unsigned char* wsb = static_cast<unsigned char*>(ws);
  const DevJob* djobs = reinterpret_cast<const DevJob*>(wsb);
  const float* dlut = reinterpret_cast<const float*>(wsb + L.lut_off);
  const int32_t* dtables = reinterpret_cast<const int32_t*>(wsb + L.tables_off);

  slime_prof_begin(2, bytes, stream);
  {
    long long bx = (max_rows_px + 255) / 256;
    if (bx > 4096) bx = 4096;
    if (bx < 1) bx = 1;
    dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n_jobs));
    if (any_plain) {
      resize_rows_kernel<<<grid, 256, 0, stream>>>(src, djobs, dtables, wsb);
      SLIME_AFTER_LAUNCH();
    }
    if (any_padded) {
      resize_rows_padded_kernel<<<grid, 256, 0, stream>>>(src, djobs, dtables, wsb);
      SLIME_AFTER_LAUNCH();
    }
  }
  {
    long long bx = (max_canvas_groups + 255) / 256;
    if (bx > 4096) bx = 4096;
    dim3 grid(static_cast<unsigned>(bx), static_cast<unsigned>(n_jobs));
    if (out_dtype == 0) {
      resize_cols_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(src, djobs, dtables, dlut, wsb,
                                                                 static_cast<__nv_bfloat16*>(out), crop);
    } else if (out_dtype == 1) {
      resize_cols_kernel<float><<<grid, 256, 0, stream>>>(src, djobs, dtables, dlut, wsb, static_cast<float*>(out), crop);
    } else {
      resize_cols_kernel<__half><<<grid, 256, 0, stream>>>(src, djobs, dtables, dlut, wsb, static_cast<__half*>(out), crop);
    }
    SLIME_AFTER_LAUNCH();
  }
  slime_prof_end(stream);
  return SLIME_OK;
}
