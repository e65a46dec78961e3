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
// Both are HBM/L2-bound integer kernels (no tensor cores): algorithmic bytes = source image + output crops.
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
  long long src_off;  // byte offset of the image in the source buffer
  long long tmp_off;  // byte offset of this job's intermediate inside the workspace
  int src_w, src_h;
  int virt_w, virt_h, virt_x, virt_y;
  int out_w, out_h;
  int canvas_w, canvas_h;
  int paste_x, paste_y;
  int first_crop;
  int ksize_h, ksize_v;
  int hb_off, hk_off, vb_off, vk_off;  // int32 offsets into the table area: bounds (2 per output) and coefficients
  int skip_rows_pass;                  // width unchanged and no padding: the vertical pass reads the source itself
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

// Resample.c precompute_coeffs (box = whole image) + normalize_coeffs_8bpc.
// bounds: 2 ints per output (first tap, tap count); kk: ksize ints per output.
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

// horizontal pass: one thread per (row of the virtual source, output column), all three channels
__global__ void __launch_bounds__(256) resize_rows_kernel(const unsigned char* __restrict__ src,
                                                          const DevJob* __restrict__ jobs,
                                                          const int32_t* __restrict__ tables,
                                                          unsigned char* __restrict__ ws) {
  const DevJob jb = jobs[blockIdx.y];
  if (jb.skip_rows_pass) return;  // ImagingResample: need_horizontal == false
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
__device__ __forceinline__ T from_float(float v);
template <>
__device__ __forceinline__ float from_float<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_float<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ __half from_float<__half>(float v) { return __float2half_rn(v); }

// vertical pass + canvas placement + table lookup: one thread per canvas pixel, all three channels
template <typename T>
__global__ void __launch_bounds__(256) resize_cols_kernel(const unsigned char* __restrict__ src,
                                                          const DevJob* __restrict__ jobs,
                                                          const int32_t* __restrict__ tables,
                                                          const float* __restrict__ lut,
                                                          const unsigned char* __restrict__ ws, T* __restrict__ out,
                                                          int crop) {
  const DevJob jb = jobs[blockIdx.y];
  const long long total = static_cast<long long>(jb.canvas_w) * jb.canvas_h;
  const unsigned char* tmp = jb.skip_rows_pass ? src + jb.src_off : ws + jb.tmp_off;
  const int32_t* bounds = tables + jb.vb_off;
  const int32_t* kk = tables + jb.vk_off;
  const int tiles_x = jb.canvas_w / crop;
  const size_t plane = static_cast<size_t>(crop) * crop;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += static_cast<long long>(gridDim.x) * 256) {
    const int cy = static_cast<int>(i / jb.canvas_w);
    const int cx = static_cast<int>(i - static_cast<long long>(cy) * jb.canvas_w);
    const int rx = cx - jb.paste_x, ry = cy - jb.paste_y;
    int v0 = 0, v1 = 0, v2 = 0;  // Image.new('RGB', ..., (0, 0, 0)) outside the pasted image
    if (rx >= 0 && rx < jb.out_w && ry >= 0 && ry < jb.out_h) {
      const int ymin = bounds[2 * ry], n = bounds[2 * ry + 1];
      const int32_t* k = kk + static_cast<size_t>(ry) * jb.ksize_v;
      int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
      const unsigned char* p = tmp + (static_cast<size_t>(ymin) * jb.out_w + rx) * 3;
      const size_t stride = static_cast<size_t>(jb.out_w) * 3;
      for (int t = 0; t < n; ++t) {
        const int c = k[t];
        s0 += p[0] * c;
        s1 += p[1] * c;
        s2 += p[2] * c;
        p += stride;
      }
      v0 = clip8(s0);
      v1 = clip8(s1);
      v2 = clip8(s2);
    }
    const int tile = (cy / crop) * tiles_x + cx / crop;
    const int ty = cy % crop, tx = cx % crop;
    T* o = out + (static_cast<size_t>(jb.first_crop + tile) * 3) * plane + static_cast<size_t>(ty) * crop + tx;
    o[0] = from_float<T>(lut[v0]);
    o[plane] = from_float<T>(lut[256 + v1]);
    o[2 * plane] = from_float<T>(lut[512 + v2]);
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Layout {
  size_t jobs_bytes, lut_off, tables_off, tables_ints, tmp_off, total;
  std::vector<long long> tmp_offs;
};

int validate_jobs(const slime_resize_job* jobs, int n_jobs, int crop) {
  SLIME_REQUIRE(jobs != nullptr && n_jobs > 0, "preprocess: no jobs");
  SLIME_REQUIRE(crop > 0, "preprocess: crop size %d", crop);
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
    SLIME_REQUIRE(q.first_crop >= 0 && q.src_offset >= 0, "preprocess job %d: negative offset", j);
  }
  return SLIME_OK;
}

Layout make_layout(const slime_resize_job* jobs, int n_jobs) {
  Layout L;
  L.jobs_bytes = align_up(sizeof(DevJob) * n_jobs, 256);
  L.lut_off = L.jobs_bytes;
  L.tables_off = L.lut_off + 3 * 256 * sizeof(float);
  size_t ints = 0;
  for (int j = 0; j < n_jobs; ++j) {
    const slime_resize_job& q = jobs[j];
    ints += static_cast<size_t>(q.out_w) * (2 + coeff_ksize(q.virt_w, q.out_w));
    ints += static_cast<size_t>(q.out_h) * (2 + coeff_ksize(q.virt_h, q.out_h));
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
  long long max_rows_px = 0, max_canvas_px = 0;
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
    d.skip_rows_pass = (q.out_w == q.virt_w && q.virt_w == q.src_w && q.virt_h == q.src_h) ? 1 : 0;
    d.ksize_h = coeff_ksize(q.virt_w, q.out_w);
    d.ksize_v = coeff_ksize(q.virt_h, q.out_h);
    d.hb_off = static_cast<int>(cursor); cursor += 2 * static_cast<size_t>(q.out_w);
    d.hk_off = static_cast<int>(cursor); cursor += static_cast<size_t>(q.out_w) * d.ksize_h;
    d.vb_off = static_cast<int>(cursor); cursor += 2 * static_cast<size_t>(q.out_h);
    d.vk_off = static_cast<int>(cursor); cursor += static_cast<size_t>(q.out_h) * d.ksize_v;
    precompute_coeffs(q.virt_w, q.out_w, d.ksize_h, tables + d.hb_off, tables + d.hk_off);
    precompute_coeffs(q.virt_h, q.out_h, d.ksize_v, tables + d.vb_off, tables + d.vk_off);
    const long long rows_px = d.skip_rows_pass ? 0 : static_cast<long long>(q.virt_h) * q.out_w;
    const long long canvas_px = static_cast<long long>(q.canvas_w) * q.canvas_h;
    if (rows_px > max_rows_px) max_rows_px = rows_px;
    if (canvas_px > max_canvas_px) max_canvas_px = canvas_px;
    bytes += 3.0 * q.src_w * q.src_h + 3.0 * canvas_px * (out_dtype == 1 ? 4 : 2);
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
    resize_rows_kernel<<<grid, 256, 0, stream>>>(src, djobs, dtables, wsb);
    SLIME_AFTER_LAUNCH();
  }
  {
    long long bx = (max_canvas_px + 255) / 256;
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
