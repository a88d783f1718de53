// Bandwidth-bound helper kernels around the implicit-GEMM convolution: weight packing, the stem's
// normalise + space-to-depth, max-pool, and the heads' bilinear upsample into the NCHW fp32 outputs.
#include "fp_internal.h"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fp {

__device__ __forceinline__ uint16_t to16(float v, int fmt) {
  if (fmt == FP_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    return *reinterpret_cast<uint16_t*>(&h);
  }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
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

extern "C" int fp_heads_to_nchw(const float* mask_head, const float* depth_head, int batch, int h, int w, int scale,
                                float* out, void* stream) {
  if (!mask_head || !depth_head || !out || scale < 1) return set_error(FP_ERR_ARG, "fp_heads_to_nchw: bad argument");
  const long long n = static_cast<long long>(batch) * h * scale * w * scale;
  heads_to_nchw_kernel<<<blocks_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(mask_head), reinterpret_cast<const float2*>(depth_head), batch, h, w, scale, out);
  FP_CHECK_LAUNCH("fp_heads_to_nchw");
  return FP_OK;
}
