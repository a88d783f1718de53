// Bandwidth-bound helper kernels around the implicit-GEMM convolution: weight packing, the stem's
// normalise + space-to-depth, max-pool, and the heads' bilinear upsample into the NCHW fp32 outputs.
#include "fp_internal.h"
#include "fp_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fp {

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == FP_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  return static_cast<uint16_t>(fp::f32x2_to_f16x2_sat(v, 0.f) & 0xFFFFu);
}

__device__ __forceinline__ float from16(uint16_t u, int fmt) {
  if (fmt == FP_BF16) return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&u));
  return __half2float(*reinterpret_cast<__half*>(&u));
}
// v -> (hi, lo) with hi = round16(v), lo = round16(v - hi)
__device__ __forceinline__ void split16(float v, int fmt, uint16_t& hi, uint16_t& lo) {
  hi = to16(v, fmt);
  lo = to16(v - from16(hi, fmt), fmt);
}

// ---------------------------------------------------------------------------------- weight packing
__global__ void pack_conv_weight_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw, int cpad,
                                        int npad, int kpad, int fmt, int flip, uint16_t* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(npad) * kpad) return;
  const int n = static_cast<int>(idx / kpad);
  const int k = static_cast<int>(idx - static_cast<long long>(n) * kpad);
  const int inner = flip ? cout : cin;   // channel-minor extent of K (padded to cpad in the packed layout)
  const int rows = flip ? cin : cout;
  const int tap = k / cpad;
  const int c = k - tap * cpad;
  float v = 0.f;
  if (n < rows && tap < kh * kw && c < inner) {
    const int ky = tap / kw, kx = tap - ky * kw;
    if (!flip) v = w[((static_cast<long long>(n) * cin + c) * kh + ky) * kw + kx];
    else v = w[((static_cast<long long>(c) * cin + n) * kh + (kh - 1 - ky)) * kw + (kw - 1 - kx)];
  }
  out[idx] = to16(v, fmt);
}

__global__ void pack_stem_weight_kernel(const float* __restrict__ w, int cout, int npad, int fmt,
                                        uint16_t* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= npad * 256) return;
  const int n = idx >> 8, k = idx & 255;
  const int tap = k >> 4, ch = k & 15;
  const int bh = tap >> 2, bw = tap & 3;
  const int sub = ch >> 2, c = ch & 3;
  const int ky = 2 * bh + (sub >> 1) - 1, kx = 2 * bw + (sub & 1) - 1;
  float v = 0.f;
  if (n < cout && c < 3 && ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((n * 3 + c) * 7 + ky) * 7 + kx];
  out[idx] = to16(v, fmt);
}

// One launch for every weight tensor of the network (training re-packs all ~142 operands each step: forward and
// 180-degree-rotated data-gradient layouts).  `jobs` lives in device memory and is sorted by `first`.
__global__ void pack_batched_kernel(const FpPackJob* __restrict__ jobs, int njobs, long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  int lo = 0, hi = njobs - 1;                       // last job whose first <= idx
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (jobs[mid].first <= idx) lo = mid; else hi = mid - 1;
  }
  const FpPackJob j = jobs[lo];
  const long long local = idx - j.first;
  const float* __restrict__ w = j.w;
  uint16_t* __restrict__ out = static_cast<uint16_t*>(j.out);
  float v = 0.f;
  if (j.flip == 2) {                  // split precision: K index = tap*3*cg + part*cg + c, part 2 carries the low halves
    const int n = static_cast<int>(local / j.kpad);
    const int k = static_cast<int>(local - static_cast<long long>(n) * j.kpad);
    const int cg = j.stem ? 16 : j.cin;
    const int tap = k / (3 * cg);
    const int r = k - tap * 3 * cg;
    const int part = r / cg;
    const int c = r - part * cg;
    if (j.stem) {
      const int bh = tap >> 2, bw = tap & 3;
      const int sub = c >> 2, ci = c & 3;
      const int ky = 2 * bh + (sub >> 1) - 1, kx = 2 * bw + (sub & 1) - 1;
      if (n < j.cout && tap < 16 && ci < 3 && ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((n * 3 + ci) * 7 + ky) * 7 + kx];
    } else if (n < j.cout && tap < j.kh * j.kw) {
      const int ky = tap / j.kw, kx = tap - ky * j.kw;
      v = w[((static_cast<long long>(n) * j.cin + c) * j.kh + ky) * j.kw + kx];
    }
    uint16_t hi, lo;
    split16(v, j.dtype, hi, lo);
    out[local] = part == 2 ? lo : hi;
    return;
  }
  if (j.stem) {
    const int n = static_cast<int>(local >> 8), k = static_cast<int>(local & 255);
    const int tap = k >> 4, ch = k & 15;
    const int bh = tap >> 2, bw = tap & 3;
    const int sub = ch >> 2, c = ch & 3;
    const int ky = 2 * bh + (sub >> 1) - 1, kx = 2 * bw + (sub & 1) - 1;
    if (n < j.cout && c < 3 && ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((n * 3 + c) * 7 + ky) * 7 + kx];
  } else {
    const int n = static_cast<int>(local / j.kpad);
    const int k = static_cast<int>(local - static_cast<long long>(n) * j.kpad);
    const int inner = j.flip ? j.cout : j.cin;
    const int rows = j.flip ? j.cin : j.cout;
    const int tap = k / j.cpad;
    const int c = k - tap * j.cpad;
    if (n < rows && tap < j.kh * j.kw && c < inner) {
      const int ky = tap / j.kw, kx = tap - ky * j.kw;
      if (!j.flip) v = w[((static_cast<long long>(n) * j.cin + c) * j.kh + ky) * j.kw + kx];
      else v = w[((static_cast<long long>(c) * j.cin + n) * j.kh + (j.kh - 1 - ky)) * j.kw + (j.kw - 1 - kx)];
    }
  }
  out[local] = to16(v, j.dtype);
}

// ---------------------------------------------------------------------------------- stem prep
// one thread per 2x2 block: reads 2 rows x float2 per channel, writes 32 contiguous bytes
__global__ void stem_prep_kernel(const float* __restrict__ img, int batch, int h, int w, int fmt,
                                 uint4* __restrict__ out) {
  const int hb = h >> 1, wb = w >> 1;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * hb * wb) return;
  const int xb = static_cast<int>(idx % wb);
  const int yb = static_cast<int>((idx / wb) % hb);
  const int b = static_cast<int>(idx / (static_cast<long long>(wb) * hb));
  uint16_t v[16];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const float2 t = *reinterpret_cast<const float2*>(
          img + ((static_cast<long long>(b) * 3 + c) * h + (2 * yb + dy)) * w + 2 * xb);
      v[(dy * 2 + 0) * 4 + c] = to16((t.x - 0.45f) / 0.225f, fmt);
      v[(dy * 2 + 1) * 4 + c] = to16((t.y - 0.45f) / 0.225f, fmt);
    }
  }
#pragma unroll
  for (int s = 0; s < 4; ++s) v[s * 4 + 3] = 0;
  uint4 o0, o1;
  o0.x = v[0] | (v[1] << 16); o0.y = v[2] | (v[3] << 16); o0.z = v[4] | (v[5] << 16); o0.w = v[6] | (v[7] << 16);
  o1.x = v[8] | (v[9] << 16); o1.y = v[10] | (v[11] << 16); o1.z = v[12] | (v[13] << 16); o1.w = v[14] | (v[15] << 16);
  out[idx * 2] = o0;
  out[idx * 2 + 1] = o1;
}

// ---------------------------------------------------------------------------------- max-pool
template <typename V2>
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
  V2 r = __hmax2(*reinterpret_cast<V2*>(&a), *reinterpret_cast<V2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <typename V2>
__global__ void maxpool3x3s2_kernel(const uint4* __restrict__ in, int batch, int h, int w, int c8,
                                    uint4* __restrict__ out) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * ho * wo * c8) return;
  const int cc = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int ox = static_cast<int>(t % wo); t /= wo;
  const int oy = static_cast<int>(t % ho);
  const int b = static_cast<int>(t / ho);
  uint4 m;
  bool first = true;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = 2 * oy + dy;
    if (iy < 0 || iy >= h) continue;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = 2 * ox + dx;
      if (ix < 0 || ix >= w) continue;
      const uint4 v = in[((static_cast<long long>(b) * h + iy) * w + ix) * c8 + cc];
      if (first) { m = v; first = false; }
      else { m.x = max2<V2>(m.x, v.x); m.y = max2<V2>(m.y, v.y); m.z = max2<V2>(m.z, v.z); m.w = max2<V2>(m.w, v.w); }
    }
  }
  out[idx] = m;
}

// ---------------------------------------------------------------------------------- split-precision helpers
// K index = tap*3*cin + part*cin + c  (stem: tap*48 + part*16 + ch of the space-to-depth formulation)
__global__ void pack_conv_weight_split_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw, int npad,
                                              int kpad, int fmt, int stem, uint16_t* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(npad) * kpad) return;
  const int n = static_cast<int>(idx / kpad);
  const int k = static_cast<int>(idx - static_cast<long long>(n) * kpad);
  const int cg = stem ? 16 : cin;                 // channels per tap in the gathered tensor
  const int tap = k / (3 * cg);
  const int r = k - tap * 3 * cg;
  const int part = r / cg;
  const int c = r - part * cg;
  float v = 0.f;
  if (stem) {
    const int bh = tap >> 2, bw = tap & 3;
    const int sub = c >> 2, ci = c & 3;
    const int ky = 2 * bh + (sub >> 1) - 1, kx = 2 * bw + (sub & 1) - 1;
    if (n < cout && tap < 16 && ci < 3 && ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = w[((n * 3 + ci) * 7 + ky) * 7 + kx];
  } else if (n < cout && tap < kh * kw) {
    const int ky = tap / kw, kx = tap - ky * kw;
    v = w[((static_cast<long long>(n) * cin + c) * kh + ky) * kw + kx];
  }
  uint16_t hi, lo;
  split16(v, fmt, hi, lo);
  out[idx] = part == 2 ? lo : hi;
}

__global__ void stem_prep_split_kernel(const float* __restrict__ img, int batch, int h, int w, int fmt,
                                       uint4* __restrict__ out_hi, uint4* __restrict__ out_lo) {
  const int hb = h >> 1, wb = w >> 1;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * hb * wb) return;
  const int xb = static_cast<int>(idx % wb);
  const int yb = static_cast<int>((idx / wb) % hb);
  const int b = static_cast<int>(idx / (static_cast<long long>(wb) * hb));
  uint16_t vh[16], vl[16];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      const float2 t = *reinterpret_cast<const float2*>(
          img + ((static_cast<long long>(b) * 3 + c) * h + (2 * yb + dy)) * w + 2 * xb);
      split16((t.x - 0.45f) / 0.225f, fmt, vh[(dy * 2 + 0) * 4 + c], vl[(dy * 2 + 0) * 4 + c]);
      split16((t.y - 0.45f) / 0.225f, fmt, vh[(dy * 2 + 1) * 4 + c], vl[(dy * 2 + 1) * 4 + c]);
    }
  }
#pragma unroll
  for (int s4 = 0; s4 < 4; ++s4) { vh[s4 * 4 + 3] = 0; vl[s4 * 4 + 3] = 0; }
  uint4 o0, o1;
  o0.x = vh[0] | (vh[1] << 16); o0.y = vh[2] | (vh[3] << 16); o0.z = vh[4] | (vh[5] << 16); o0.w = vh[6] | (vh[7] << 16);
  o1.x = vh[8] | (vh[9] << 16); o1.y = vh[10] | (vh[11] << 16); o1.z = vh[12] | (vh[13] << 16); o1.w = vh[14] | (vh[15] << 16);
  out_hi[idx * 2] = o0; out_hi[idx * 2 + 1] = o1;
  o0.x = vl[0] | (vl[1] << 16); o0.y = vl[2] | (vl[3] << 16); o0.z = vl[4] | (vl[5] << 16); o0.w = vl[6] | (vl[7] << 16);
  o1.x = vl[8] | (vl[9] << 16); o1.y = vl[10] | (vl[11] << 16); o1.z = vl[12] | (vl[13] << 16); o1.w = vl[14] | (vl[15] << 16);
  out_lo[idx * 2] = o0; out_lo[idx * 2 + 1] = o1;
}

// max over hi + lo (fp32), split again (exact: the maximum is one of the inputs)
__global__ void maxpool3x3s2_split_kernel(const uint16_t* __restrict__ in_hi, const uint16_t* __restrict__ in_lo, int batch,
                                          int h, int w, int c, int fmt, uint16_t* __restrict__ out_hi,
                                          uint16_t* __restrict__ out_lo) {
  const int ho = (h + 1) / 2, wo = (w + 1) / 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * ho * wo * c) return;
  const int cc = static_cast<int>(idx % c);
  long long t = idx / c;
  const int ox = static_cast<int>(t % wo); t /= wo;
  const int oy = static_cast<int>(t % ho);
  const int b = static_cast<int>(t / ho);
  float best = -3.0e38f;
  uint16_t bh = 0, bl = 0;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = 2 * oy + dy;
    if (iy < 0 || iy >= h) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = 2 * ox + dx;
      if (ix < 0 || ix >= w) continue;
      const long long o = ((static_cast<long long>(b) * h + iy) * w + ix) * c + cc;
      const uint16_t a = in_hi[o], l = in_lo[o];
      const float v = from16(a, fmt) + from16(l, fmt);
      if (v > best) { best = v; bh = a; bl = l; }
    }
  }
  out_hi[idx] = bh;
  out_lo[idx] = bl;
}

// ---------------------------------------------------------------------------------- heads -> NCHW
// torch upsample_bilinear2d, align_corners=False: src = max((dst+0.5)/scale - 0.5, 0)
__global__ void heads_to_nchw_kernel(const float2* __restrict__ mask_head, const float2* __restrict__ depth_head,
                                     int batch, int h, int w, int scale, float* __restrict__ out) {
  const int H = h * scale, W = w * scale;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * H * W) return;
  const int X = static_cast<int>(idx % W);
  const int Y = static_cast<int>((idx / W) % H);
  const int b = static_cast<int>(idx / (static_cast<long long>(W) * H));
  float2 m, d;
  if (scale == 1) {
    m = mask_head[idx];
    d = depth_head[idx];
  } else {
    const float rs = 1.f / static_cast<float>(scale);
    const float sy = fmaxf(rs * (Y + 0.5f) - 0.5f, 0.f), sx = fmaxf(rs * (X + 0.5f) - 0.5f, 0.f);
    const int y0 = static_cast<int>(sy), x0 = static_cast<int>(sx);
    const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
    const float ly = sy - y0, lx = sx - x0, hy = 1.f - ly, hx = 1.f - lx;
    const long long base = static_cast<long long>(b) * h * w;
    const long long i00 = base + static_cast<long long>(y0) * w + x0, i01 = base + static_cast<long long>(y0) * w + x1;
    const long long i10 = base + static_cast<long long>(y1) * w + x0, i11 = base + static_cast<long long>(y1) * w + x1;
    const float2 a = mask_head[i00], bb = mask_head[i01], c = mask_head[i10], e = mask_head[i11];
    m.x = hy * (hx * a.x + lx * bb.x) + ly * (hx * c.x + lx * e.x);
    m.y = hy * (hx * a.y + lx * bb.y) + ly * (hx * c.y + lx * e.y);
    const float2 a2 = depth_head[i00], b2 = depth_head[i01], c2 = depth_head[i10], e2 = depth_head[i11];
    d.x = hy * (hx * a2.x + lx * b2.x) + ly * (hx * c2.x + lx * e2.x);
    d.y = hy * (hx * a2.y + lx * b2.y) + ly * (hx * c2.y + lx * e2.y);
  }
  const long long plane = static_cast<long long>(H) * W;
  float* o = out + static_cast<long long>(b) * 4 * plane + static_cast<long long>(Y) * W + X;
  __stcs(o, m.x);
  __stcs(o + plane, m.y);
  __stcs(o + 2 * plane, d.x);
  __stcs(o + 3 * plane, d.y);
}

static inline unsigned blocks_for(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

}  // namespace fp

using namespace fp;

extern "C" int fp_pack_conv_weight(const float* w, int cout, int cin, int kh, int kw, int cpad, int npad, int kpad,
                                   int dtype, int flip, void* out, void* stream) {
  if (!w || !out) return set_error(FP_ERR_ARG, "fp_pack_conv_weight: null argument");
  const int rows = flip ? cin : cout, inner = flip ? cout : cin;
  if (cpad < inner || npad < rows || kpad < kh * kw * cpad)
    return set_error(FP_ERR_ARG, "fp_pack_conv_weight: cpad/npad/kpad too small");
  const long long n = static_cast<long long>(npad) * kpad;
  pack_conv_weight_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, kh, kw, cpad, npad, kpad, dtype, flip, static_cast<uint16_t*>(out));
  FP_CHECK_LAUNCH("fp_pack_conv_weight");
  return FP_OK;
}

extern "C" int fp_pack_stem_weight(const float* w, int cout, int npad, int dtype, void* out, void* stream) {
  if (!w || !out || npad < cout) return set_error(FP_ERR_ARG, "fp_pack_stem_weight: bad argument");
  pack_stem_weight_kernel<<<blocks_for(static_cast<long long>(npad) * 256, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, npad, dtype, static_cast<uint16_t*>(out));
  FP_CHECK_LAUNCH("fp_pack_stem_weight");
  return FP_OK;
}

extern "C" int fp_pack_weights_batched(const FpPackJob* jobs_dev, int njobs, long long total, void* stream) {
  if (!jobs_dev || njobs < 1 || total < 1) return set_error(FP_ERR_ARG, "fp_pack_weights_batched: bad argument");
  pack_batched_kernel<<<blocks_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(jobs_dev, njobs, total);
  FP_CHECK_LAUNCH("fp_pack_weights_batched");
  return FP_OK;
}

extern "C" int fp_stem_prep(const float* img, int batch, int h, int w, int dtype, void* out, void* stream) {
  if (!img || !out || (h & 1) || (w & 1)) return set_error(FP_ERR_ARG, "fp_stem_prep: bad argument (h, w must be even)");
  const long long n = static_cast<long long>(batch) * (h / 2) * (w / 2);
  stem_prep_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(img, batch, h, w, dtype,
                                                                                      static_cast<uint4*>(out));
  FP_CHECK_LAUNCH("fp_stem_prep");
  return FP_OK;
}

extern "C" int fp_maxpool3x3s2(const void* in, int batch, int h, int w, int c, int dtype, void* out, void* stream) {
  if (!in || !out || (c & 7)) return set_error(FP_ERR_ARG, "fp_maxpool3x3s2: bad argument (c must be a multiple of 8)");
  const long long n = static_cast<long long>(batch) * ((h + 1) / 2) * ((w + 1) / 2) * (c / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == FP_BF16)
    maxpool3x3s2_kernel<__nv_bfloat162><<<blocks_for(n, 256), 256, 0, st>>>(static_cast<const uint4*>(in), batch, h, w, c / 8, static_cast<uint4*>(out));
  else
    maxpool3x3s2_kernel<__half2><<<blocks_for(n, 256), 256, 0, st>>>(static_cast<const uint4*>(in), batch, h, w, c / 8, static_cast<uint4*>(out));
  FP_CHECK_LAUNCH("fp_maxpool3x3s2");
  return FP_OK;
}

extern "C" int fp_pack_conv_weight_split(const float* w, int cout, int cin, int kh, int kw, int npad, int kpad, int dtype,
                                         void* out, void* stream) {
  if (!w || !out || npad < cout || kpad < 3 * kh * kw * cin) return set_error(FP_ERR_ARG, "fp_pack_conv_weight_split: bad argument");
  const long long n = static_cast<long long>(npad) * kpad;
  pack_conv_weight_split_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, cin, kh, kw, npad, kpad, dtype, 0, static_cast<uint16_t*>(out));
  FP_CHECK_LAUNCH("fp_pack_conv_weight_split");
  return FP_OK;
}

extern "C" int fp_pack_stem_weight_split(const float* w, int cout, int npad, int dtype, void* out, void* stream) {
  if (!w || !out || npad < cout) return set_error(FP_ERR_ARG, "fp_pack_stem_weight_split: bad argument");
  const long long n = static_cast<long long>(npad) * 768;
  pack_conv_weight_split_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      w, cout, 3, 7, 7, npad, 768, dtype, 1, static_cast<uint16_t*>(out));
  FP_CHECK_LAUNCH("fp_pack_stem_weight_split");
  return FP_OK;
}

extern "C" int fp_stem_prep_split(const float* img, int batch, int h, int w, int dtype, void* out_hi, void* out_lo,
                                  void* stream) {
  if (!img || !out_hi || !out_lo || (h & 1) || (w & 1)) return set_error(FP_ERR_ARG, "fp_stem_prep_split: bad argument");
  const long long n = static_cast<long long>(batch) * (h / 2) * (w / 2);
  stem_prep_split_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      img, batch, h, w, dtype, static_cast<uint4*>(out_hi), static_cast<uint4*>(out_lo));
  FP_CHECK_LAUNCH("fp_stem_prep_split");
  return FP_OK;
}

extern "C" int fp_maxpool3x3s2_split(const void* in_hi, const void* in_lo, int batch, int h, int w, int c, int dtype,
                                     void* out_hi, void* out_lo, void* stream) {
  if (!in_hi || !in_lo || !out_hi || !out_lo) return set_error(FP_ERR_ARG, "fp_maxpool3x3s2_split: null argument");
  const long long n = static_cast<long long>(batch) * ((h + 1) / 2) * ((w + 1) / 2) * c;
  maxpool3x3s2_split_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(in_hi), static_cast<const uint16_t*>(in_lo), batch, h, w, c, dtype,
      static_cast<uint16_t*>(out_hi), static_cast<uint16_t*>(out_lo));
  FP_CHECK_LAUNCH("fp_maxpool3x3s2_split");
  return FP_OK;
}

// Output edge of the batched test-set inference (footprints/evaluation/inference.py:104-108 + datasets/inference_dataset.py:35-39):
// full-resolution prediction only, sigmoid on the two mask channels ("not applied in network for stability of BCE
// loss"), stored as float16 -- what the reference computes with three extra passes (sigmoid in place, .cpu(), astype(np.float16))
__global__ void heads_to_nchw_f16_post_kernel(const float2* __restrict__ mask_head, const float2* __restrict__ depth_head,
                                              long long hw, long long n, __half* __restrict__ out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const long long b = idx / hw, pix = idx - b * hw;
  const float2 m = mask_head[idx], d = depth_head[idx];
  __half* o = out + b * 4 * hw + pix;
  o[0] = __float2half_rn(1.f / (1.f + __expf(-m.x)));
  o[hw] = __float2half_rn(1.f / (1.f + __expf(-m.y)));
  o[2 * hw] = __float2half_rn(d.x);
  o[3 * hw] = __float2half_rn(d.y);
}

extern "C" int fp_heads_to_nchw_f16_post(const float* mask_head, const float* depth_head, int batch, int h, int w, void* out_f16,
                                         void* stream) {
  if (!mask_head || !depth_head || !out_f16) return set_error(FP_ERR_ARG, "fp_heads_to_nchw_f16_post: null argument");
  const long long hw = static_cast<long long>(h) * w, n = hw * batch;
  heads_to_nchw_f16_post_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(mask_head), reinterpret_cast<const float2*>(depth_head), hw, n, static_cast<__half*>(out_f16));
  FP_CHECK_LAUNCH("fp_heads_to_nchw_f16_post");
  return FP_OK;
}

extern "C" int fp_heads_to_nchw(const float* mask_head, const float* depth_head, int batch, int h, int w, int scale,
                                float* out, void* stream) {
  if (!mask_head || !depth_head || !out || scale < 1) return set_error(FP_ERR_ARG, "fp_heads_to_nchw: bad argument");
  const long long n = static_cast<long long>(batch) * h * scale * w * scale;
  heads_to_nchw_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(mask_head), reinterpret_cast<const float2*>(depth_head), batch, h, w, scale, out);
  FP_CHECK_LAUNCH("fp_heads_to_nchw");
  return FP_OK;
}

// ---------------------------------------------------------------------------------- pyramid pooling (Segmentor, use_PSP=True)
// preprocessing/segmentation/network.py:174-207: for s in (1, 2, 4, 6): AdaptiveAvgPool2d(s) -> 1x1 conv 512 -> 128 (no
// bias) -> bilinear upsample back to h x w (align_corners=True); output = cat([x, x6, x4, x2, x1], channels).  One CTA per
// (pyramid level, image): pooled cells and their 128-channel reductions live in shared memory, so the block reads x once and
// writes the 1024-channel NHWC tensor the first decoder convolution consumes (the concat is a channel offset).
namespace fp {
namespace {
constexpr int kPspC = 512, kPspR = 128, kPspThreads = 256;

__global__ void __launch_bounds__(kPspThreads) psp_kernel(const uint16_t* __restrict__ x, int h, int w, int fmt,
                                                         const float* __restrict__ w1, const float* __restrict__ w2,
                                                         const float* __restrict__ w4, const float* __restrict__ w6,
                                                         uint16_t* __restrict__ out) {
  extern __shared__ float s_psp[];
  const int level = blockIdx.x, b = blockIdx.y;              // level 0..3 -> pool 6, 4, 2, 1 (the order of the concat)
  const int s = level == 0 ? 6 : (level == 1 ? 4 : (level == 2 ? 2 : 1));
  const float* __restrict__ wr = level == 0 ? w6 : (level == 1 ? w4 : (level == 2 ? w2 : w1));
  float* pooled = s_psp;                                     // [s*s][512]
  float* reduced = s_psp + 36 * kPspC;                       // [s*s][128]
  const uint16_t* xb = x + static_cast<long long>(b) * h * w * kPspC;
  uint16_t* ob = out + static_cast<long long>(b) * h * w * (2 * kPspC);
  const int tid = threadIdx.x;

  // x itself -> channels [0, 512): this CTA copies its quarter (128 channels = 16 vectors of 8 per pixel)
  for (int i = tid; i < h * w * 16; i += kPspThreads) {
    const int pix = i >> 4, v = i & 15;
    reinterpret_cast<uint4*>(ob + static_cast<long long>(pix) * (2 * kPspC) + level * 128)[v] =
        __ldg(reinterpret_cast<const uint4*>(xb + static_cast<long long>(pix) * kPspC + level * 128) + v);
  }
  // adaptive average pooling: cell (i, j) covers rows [floor(i*h/s), ceil((i+1)*h/s)), columns likewise
  for (int cell = 0; cell < s * s; ++cell) {
    const int ci = cell / s, cj = cell - ci * s;
    const int y0 = (ci * h) / s, y1 = ((ci + 1) * h + s - 1) / s;
    const int x0 = (cj * w) / s, x1 = ((cj + 1) * w + s - 1) / s;
    const float inv = 1.f / static_cast<float>((y1 - y0) * (x1 - x0));
    for (int c = tid; c < kPspC; c += kPspThreads) {
      float acc = 0.f;
      for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) acc += from16(xb[(static_cast<long long>(yy) * w + xx) * kPspC + c], fmt);
      pooled[cell * kPspC + c] = acc * inv;
    }
  }
  __syncthreads();
  // 1x1 convolution 512 -> 128 on the pooled cells: one warp per output channel, lanes split the input channels
  const int warp = tid >> 5, lane = tid & 31;
  for (int co = warp; co < kPspR; co += kPspThreads / 32) {
    float wv[kPspC / 32];
#pragma unroll
    for (int k = 0; k < kPspC / 32; ++k) wv[k] = __ldg(wr + co * kPspC + lane + 32 * k);
    for (int cell = 0; cell < s * s; ++cell) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < kPspC / 32; ++k) acc += wv[k] * pooled[cell * kPspC + lane + 32 * k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) reduced[cell * kPspR + co] = acc;
    }
  }
  __syncthreads();
  // bilinear upsample s x s -> h x w, align_corners=True: src = dst * (s - 1) / (out - 1)
  const float sy = h > 1 ? static_cast<float>(s - 1) / static_cast<float>(h - 1) : 0.f;
  const float sx = w > 1 ? static_cast<float>(s - 1) / static_cast<float>(w - 1) : 0.f;
  for (int i = tid; i < h * w * kPspR; i += kPspThreads) {
    const int pix = i >> 7, co = i & (kPspR - 1);
    const int yy = pix / w, xx = pix - yy * w;
    const float fy = sy * yy, fx = sx * xx;
    const int iy0 = static_cast<int>(fy), ix0 = static_cast<int>(fx);
    const int iy1 = iy0 + (iy0 < s - 1 ? 1 : 0), ix1 = ix0 + (ix0 < s - 1 ? 1 : 0);
    const float ly1 = fy - iy0, lx1 = fx - ix0, ly0 = 1.f - ly1, lx0 = 1.f - lx1;
    const float v = ly0 * (lx0 * reduced[(iy0 * s + ix0) * kPspR + co] + lx1 * reduced[(iy0 * s + ix1) * kPspR + co]) +
                    ly1 * (lx0 * reduced[(iy1 * s + ix0) * kPspR + co] + lx1 * reduced[(iy1 * s + ix1) * kPspR + co]);
    ob[static_cast<long long>(pix) * (2 * kPspC) + kPspC + level * kPspR + co] = to16(v, fmt);
  }
}
}  // namespace
}  // namespace fp

extern "C" int fp_psp_fwd(const void* x, int batch, int h, int w, int dtype, const float* w_pool1, const float* w_pool2,
                          const float* w_pool4, const float* w_pool6, void* out, void* stream) {
  if (!x || !w_pool1 || !w_pool2 || !w_pool4 || !w_pool6 || !out) return fp::set_error(FP_ERR_ARG, "fp_psp_fwd: null argument");
  if (batch < 1 || batch > 65535 || h < 1 || w < 1) return fp::set_error(FP_ERR_ARG, "fp_psp_fwd: bad geometry %dx%dx%d", batch, h, w);
  const int smem = (36 * fp::kPspC + 36 * fp::kPspR) * 4;
  int rc = fp::ensure_max_smem(reinterpret_cast<const void*>(fp::psp_kernel), smem, "psp_kernel");
  if (rc) return rc;
  fp::psp_kernel<<<dim3(4, batch), fp::kPspThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint16_t*>(x), h, w, dtype, w_pool1, w_pool2, w_pool4, w_pool6, static_cast<uint16_t*>(out));
  FP_CHECK_LAUNCH("fp_psp_fwd");
  return FP_OK;
}
