// Image preprocessing of the recognizer on the GPU: what doc2tex/utils/predict_utils.py::resize (14-115)
// does per image on the host with PIL / cv2 / albumentations — optional integer INTER_AREA down-sampling, the crop-to-ink of
// data_utils.py::pad (10-47), minmax_size (62-82: Pillow LANCZOS shrink / white canvas), Normalize — for a whole list of
// differently sized crops in a handful of launches.  All of it is byte / integer work bound by HBM; the only floating point
// is the reference's own float64 min-max stretch (256 possible inputs per image: a lookup table) and the final fp32 affine.
#pragma once
#include "../../include/doc2tex_b200.h"
#include "common.cuh"

namespace d2t {

// grey value of the (down-sampled) source: cv2.resize(INTER_AREA) with an integer scale = rounded box mean
__device__ __forceinline__ int prep_src(const uint8_t* __restrict__ base, const d2t_prep_image& im, int y, int x) {
  if (im.ds <= 1) return base[im.src_off + (long long)y * im.w0 + x];
  int s = 0;
  for (int dy = 0; dy < im.ds; ++dy)
    for (int dx = 0; dx < im.ds; ++dx) s += base[im.src_off + (long long)(y * im.ds + dy) * im.w0 + (x * im.ds + dx)];
  const int n = im.ds * im.ds;
  return (2 * s + n) / (2 * n);
}

// The reference's min-max stretch (data_utils.py:21): float64, value = (v - vmin) / (255 - vmin) * 255 — the LA conversion's
// constant alpha makes data.max() 255.  256 possible inputs: evaluated once per image into a table.
__device__ __forceinline__ double prep_stretch(int v, int vmin) { return (double)(v - vmin) / (double)(255 - vmin) * 255.0; }

// One block per image: smallest grey value, polarity (mean of the stretched image > 128), ink bounding box.
// stats[i] = {x, y, w, h, inverted, vmin, status (0 ok, 1 blank image, 2 nothing crosses the threshold), 0}
__global__ void __launch_bounds__(1024)
prep_stats_kernel(const uint8_t* __restrict__ packed, const d2t_prep_image* __restrict__ imgs, int32_t* __restrict__ stats) {
  __shared__ int s_min, s_x0, s_x1, s_y0, s_y1, s_inv;
  __shared__ unsigned long long s_sum;
  __shared__ unsigned char s_ink[256];
  const d2t_prep_image im = imgs[blockIdx.x];
  const int H = im.h0 / max(im.ds, 1), W = im.w0 / max(im.ds, 1);
  const long long n = (long long)H * W;
  if (threadIdx.x == 0) { s_min = 255; s_sum = 0ull; s_x0 = W; s_x1 = -1; s_y0 = H; s_y1 = -1; }
  __syncthreads();
  int mn = 255;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) mn = min(mn, prep_src(packed, im, (int)(i / W), (int)(i % W)));
  for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  if ((threadIdx.x & 31) == 0) atomicMin(&s_min, mn);
  __syncthreads();
  const int vmin = s_min;
  int32_t* out = stats + (size_t)blockIdx.x * 8;
  if (vmin == 255) {
    if (threadIdx.x == 0) { out[0] = out[1] = out[2] = out[3] = out[4] = 0; out[5] = 255; out[6] = 1; out[7] = 0; }
    return;
  }
  unsigned long long sum = 0ull;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) sum += (unsigned)(prep_src(packed, im, (int)(i / W), (int)(i % W)) - vmin);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0) atomicAdd(&s_sum, sum);
  __syncthreads();
  if (threadIdx.x == 0) {
    // mean((v - vmin) * 255 / (255 - vmin)) > 128  <=>  255 * sum > 128 * n * (255 - vmin), in integers
    s_inv = (255ull * s_sum > 128ull * (unsigned long long)n * (unsigned long long)(255 - vmin)) ? 0 : 1;
  }
  __syncthreads();
  const int inv = s_inv;
  if (threadIdx.x < 256) {
    const double val = prep_stretch((int)threadIdx.x < vmin ? vmin : (int)threadIdx.x, vmin);
    s_ink[threadIdx.x] = inv ? (val > 128.0) : (val < 128.0);     // data_utils.py:23-29
  }
  __syncthreads();
  int x0 = W, x1 = -1, y0 = H, y1 = -1;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int y = (int)(i / W), x = (int)(i % W);
    if (s_ink[prep_src(packed, im, y, x)]) { x0 = min(x0, x); x1 = max(x1, x); y0 = min(y0, y); y1 = max(y1, y); }
  }
  for (int o = 16; o > 0; o >>= 1) {
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&s_x0, x0); atomicMax(&s_x1, x1); atomicMin(&s_y0, y0); atomicMax(&s_y1, y1); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const bool none = s_x1 < 0;
    out[0] = none ? 0 : s_x0; out[1] = none ? 0 : s_y0;
    out[2] = none ? 0 : s_x1 - s_x0 + 1; out[3] = none ? 0 : s_y1 - s_y0 + 1;
    out[4] = inv; out[5] = vmin; out[6] = none ? 2 : 0; out[7] = 0;
  }
}

// Stage B of an image: with the crop-to-ink (plan.use_crop) the stretched (and, for light-on-dark images, inverted) grey
// values of the ink box, truncated to 8 bits, top-left in a canvas of zeros whose sides are multiples of 32
// (data_utils.py:31-46: Image.new("L", dims) is BLACK); otherwise the (down-sampled) source as it is.
__global__ void prep_crop_kernel(const uint8_t* __restrict__ packed, const d2t_prep_image* __restrict__ imgs,
                                 const d2t_prep_plan* __restrict__ plans, uint8_t* __restrict__ scratch) {
  __shared__ unsigned char s_lut[256];
  const d2t_prep_image im = imgs[blockIdx.y];
  const d2t_prep_plan pl = plans[blockIdx.y];
  if (pl.use_crop && threadIdx.x < 256) {
    double val = prep_stretch((int)threadIdx.x < pl.vmin ? pl.vmin : (int)threadIdx.x, pl.vmin);
    if (pl.inverted) val = 255.0 - val;
    s_lut[threadIdx.x] = (unsigned char)(int)val;     // ndarray.astype(uint8) truncates
  }
  __syncthreads();
  const long long n = (long long)pl.hb * pl.wb;
  uint8_t* dst = scratch + pl.off_b;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / pl.wb), x = (int)(i % pl.wb);
    int v = 0;
    if (pl.use_crop) {
      if (y < pl.crop_h && x < pl.crop_w) v = s_lut[prep_src(packed, im, y + pl.crop_y, x + pl.crop_x)];
    } else {
      v = prep_src(packed, im, y, x);
    }
    dst[i] = (uint8_t)v;
  }
}

// One 8-bit pass of Pillow's resampler (Resample.c ImagingResampleHorizontal_8bpc / Vertical_8bpc): 22 fractional bits,
// rounding offset 1 << 21, arithmetic shift, clamp to [0, 255].  coef rows: [xmin, count, k[0..ksize)].
__global__ void prep_resample_kernel(const d2t_prep_plan* __restrict__ plans, const int32_t* __restrict__ coefs,
                                     uint8_t* __restrict__ scratch, int vertical) {
  const d2t_prep_plan pl = plans[blockIdx.y];
  if (!pl.do_resize) return;
  const int in_h = vertical ? pl.hb : pl.hb, in_w = vertical ? pl.rw : pl.wb;
  const int out_h = vertical ? pl.rh : pl.hb, out_w = pl.rw;
  const uint8_t* src = scratch + (vertical ? pl.off_t : pl.off_b);
  uint8_t* dst = scratch + (vertical ? pl.off_r : pl.off_t);
  const int ksize = vertical ? pl.ky_ksize : pl.kx_ksize;
  const int32_t* tab = coefs + (vertical ? pl.ky_off : pl.kx_off);
  const long long n = (long long)out_h * out_w;
  (void)in_h;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / out_w), x = (int)(i % out_w);
    const int32_t* row = tab + (size_t)(vertical ? y : x) * (ksize + 2);
    const int first = row[0], cnt = row[1];
    int acc = 1 << 21;
    if (vertical) for (int k = 0; k < cnt; ++k) acc += (int)src[(size_t)(first + k) * in_w + x] * row[2 + k];
    else for (int k = 0; k < cnt; ++k) acc += (int)src[(size_t)y * in_w + first + k] * row[2 + k];
    acc >>= 22;
    dst[i] = (uint8_t)(acc < 0 ? 0 : (acc > 255 ? 255 : acc));
  }
}

// Final canvas (white where minmax_size enlarges, data_utils.py:72-80), albumentations Normalize in fp32
// ((v - mean*255) * (1 / (std*255))), written into the image's slot of its (H, W) bucket tensor.
__global__ void prep_finish_kernel(const d2t_prep_plan* __restrict__ plans, const uint8_t* __restrict__ scratch,
                                   float sub, float mul) {
  const d2t_prep_plan pl = plans[blockIdx.y];
  const uint8_t* src = scratch + (pl.do_resize ? pl.off_r : pl.off_b);
  const int ih = pl.do_resize ? pl.rh : pl.hb, iw = pl.do_resize ? pl.rw : pl.wb;
  float* dst = reinterpret_cast<float*>(pl.dst);
  const long long n = (long long)pl.out_h * pl.out_w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / pl.out_w), x = (int)(i % pl.out_w);
    const int v = (y < ih && x < iw) ? src[(size_t)y * iw + x] : 255;
    float f = (float)v;
    f -= sub;
    f *= mul;
    dst[i] = f;
  }
}

}  // namespace d2t
