// Input edge of the hot path: uint8 RGB frames -> the float32 NCHW tensor FootprintNetwork.forward takes.
//
// Replaces, bit for bit, what the reference does on the host for every frame
//   transforms.Resize((H, W), interpolation=Image.ANTIALIAS)   predict_simple.py:41-42,54; datasets/inference_dataset.py:25-27,46
//   transforms.ToTensor()                                       predict_simple.py:43,55;    datasets/inference_dataset.py:47
// i.e. Pillow's two-pass Lanczos (a=3) resampling in 22-bit fixed point with a uint8 intermediate image, then /255.
// The per-geometry tap tables (bounds + fixed-point weights) are built once by the host (footprints_b200/preprocess.py,
// double precision like Pillow's precompute_coeffs) and stay resident; the pixel work is the two kernels below.
//
// Byte/integer streaming work, no tensor cores.  Default kernels (FP_RESIZE_ROWS=0 selects the first-generation pair kept
// below them, which ncu showed LSU-bound on byte-wide shared-memory loads: l1tex 94 %, 22-25 instructions per tap):
//   resize_h_rows_kernel<KS>: a CTA stages 6 consecutive source rows in shared memory as RGBX words (24 byte loads in
//     flight per thread), every thread keeps the KS fixed-point weights of its output column in registers and walks the rows:
//     per tap one LDS.32, three byte extracts (ALU pipe) and three IMADs (FMA pipe);
//   resize_v_rows_kernel: a CTA owns two adjacent output rows x 128 columns; their two tap windows are laid over the
//     union of source rows as zero-padded weight pairs in shared memory, so each intermediate pixel is loaded once and
//     feeds both rows with unconditional IMADs.
// The RGBX (uchar4) intermediate (Ht x Wout x 4 B per frame) is written and consumed through L2.
#include "fp_internal.h"

using fp::set_error;

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Pillow Resample.c PRECISION_BITS

__device__ __forceinline__ unsigned char clip8(int acc) {
  return static_cast<unsigned char>(min(max(acc >> kPrecisionBits, 0), 255));
}

// grid (ht, batch); one CTA per needed source row.
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, long long total_bytes, int hin, int win,
                                                       int y0, int ht, int wout, const int2* __restrict__ bounds,
                                                       const int32_t* __restrict__ kk, uchar4* __restrict__ tmp) {
  extern __shared__ uint4 s_row4[];
  uint8_t* s_bytes = reinterpret_cast<uint8_t*>(s_row4);
  const int y = blockIdx.x, b = blockIdx.y;
  const long long row_off = (static_cast<long long>(b) * hin + y0 + y) * win * 3;
  const int lead = static_cast<int>((reinterpret_cast<uintptr_t>(src) + row_off) & 15);
  const long long a_off = row_off - lead;                      // may be negative / the tail may pass the end: guarded below
  const int nvec = (lead + win * 3 + 15) >> 4;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    const long long o = a_off + static_cast<long long>(i) * 16;
    if (o >= 0 && o + 16 <= total_bytes) {
      s_row4[i] = __ldg(reinterpret_cast<const uint4*>(src + o));
    } else {
      for (int j = 0; j < 16; ++j) s_bytes[i * 16 + j] = (o + j >= 0 && o + j < total_bytes) ? src[o + j] : 0;
    }
  }
  __syncthreads();
  const uint8_t* row = s_bytes + lead;
  uchar4* trow = tmp + (static_cast<long long>(b) * ht + y) * wout;
  for (int xo = threadIdx.x; xo < wout; xo += blockDim.x) {
    const int2 bd = bounds[xo];
    int r = 1 << (kPrecisionBits - 1), g = r, bl = r;
    const uint8_t* p = row + bd.x * 3;
    for (int t = 0; t < bd.y; ++t, p += 3) {
      const int k = __ldg(kk + static_cast<long long>(t) * wout + xo);
      r += p[0] * k;
      g += p[1] * k;
      bl += p[2] * k;
    }
    trow[xo] = make_uchar4(clip8(r), clip8(g), clip8(bl), 0);
  }
}

// grid (ceil(wout/128), hout, batch)
__global__ void __launch_bounds__(128) resize_v_kernel(const uchar4* __restrict__ tmp, int ht, int wout, int hout,
                                                       const int2* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                       float* __restrict__ out) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x, yo = blockIdx.y, b = blockIdx.z;
  if (xo >= wout) return;
  const int2 bd = bounds[yo];
  int r = 1 << (kPrecisionBits - 1), g = r, bl = r;
  const uchar4* p = tmp + (static_cast<long long>(b) * ht + bd.x) * wout + xo;
  for (int t = 0; t < bd.y; ++t, p += wout) {
    const int k = __ldg(kk + t * hout + yo);
    const uchar4 v = *p;
    r += v.x * k;
    g += v.y * k;
    bl += v.z * k;
  }
  const long long plane = static_cast<long long>(hout) * wout;
  float* o = out + static_cast<long long>(b) * 3 * plane + static_cast<long long>(yo) * wout + xo;
  o[0] = static_cast<float>(clip8(r)) / 255.0f;            // ToTensor: IEEE division, not a reciprocal multiply
  o[plane] = static_cast<float>(clip8(g)) / 255.0f;
  o[2 * plane] = static_cast<float>(clip8(bl)) / 255.0f;
}


// ---- second generation (default) ----
constexpr int kRowsPerCta = 6;
constexpr int kVSpanMax = 64;

// grid (ceil(ht / rows_per_cta), batch); 256 threads
template <int KS>
__global__ void __launch_bounds__(256) resize_h_rows_kernel(const uint8_t* __restrict__ src, int hin, int win, int y0, int ht,
                                                            int wout, int rows_per_cta, const int2* __restrict__ bounds,
                                                            const int32_t* __restrict__ kk, uchar4* __restrict__ tmp) {
  extern __shared__ uint32_t s_px[];                         // [rows * win + KS] RGBX words, rows back to back
  const int r0 = blockIdx.x * rows_per_cta, b = blockIdx.y;
  const int nr = min(rows_per_cta, ht - r0);
  // the nr source rows are one contiguous byte span; a tap that runs past the end of its row has weight zero, so it may
  // read the next row's pixels (or the KS zero words after the last row) -- no per-row padding, no division.
  // Eight pixels (24 byte loads) per thread are in flight before the first dependent store.
  const uint8_t* base = src + (static_cast<long long>(b) * hin + y0 + r0) * win * 3;
  const int npx = nr * win;
  constexpr int U = 8;
  for (int p0 = threadIdx.x; p0 < npx; p0 += 256 * U) {
    uint32_t c[U][3];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * 256;
      if (p < npx) {
        c[u][0] = __ldg(base + 3 * p); c[u][1] = __ldg(base + 3 * p + 1); c[u][2] = __ldg(base + 3 * p + 2);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = p0 + u * 256;
      if (p < npx) s_px[p] = c[u][0] | (c[u][1] << 8) | (c[u][2] << 16);
    }
  }
  if (threadIdx.x < KS) s_px[npx + threadIdx.x] = 0u;
  const int rs = win;
  __syncthreads();
  for (int xo = threadIdx.x; xo < wout; xo += blockDim.x) {
    int k[KS];
#pragma unroll
    for (int t = 0; t < KS; ++t) k[t] = __ldg(kk + t * wout + xo);      // weights past the column's tap count are zero
    const uint32_t* p = s_px + bounds[xo].x;
    uchar4* o = tmp + (static_cast<long long>(b) * ht + r0) * wout + xo;
    for (int r = 0; r < nr; ++r, p += rs, o += wout) {
      int cr = 1 << (kPrecisionBits - 1), cg = cr, cb = cr;
#pragma unroll
      for (int t = 0; t < KS; ++t) {
        const uint32_t v = p[t];
        cr += static_cast<int>(__byte_perm(v, 0u, 0x4440)) * k[t];     // one PRMT per channel (ALU pipe), IMADs on the FMA pipe
        cg += static_cast<int>(__byte_perm(v, 0u, 0x4441)) * k[t];
        cb += static_cast<int>(__byte_perm(v, 0u, 0x4442)) * k[t];
      }
      *o = make_uchar4(clip8(cr), clip8(cg), clip8(cb), 0);
    }
  }
}

// grid (ceil(wout / 128), ceil(hout / 2), batch); 128 threads
__global__ void __launch_bounds__(128) resize_v_rows_kernel(const uint32_t* __restrict__ tmp, int ht, int wout, int hout,
                                                            const int2* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                            float* __restrict__ out) {
  __shared__ int2 s_k[kVSpanMax];                            // weight pair (row yo0, row yo0 + 1) per source row of the union
  const int yo0 = blockIdx.y * 2, b = blockIdx.z;
  const bool two = yo0 + 1 < hout;
  const int2 b0 = bounds[yo0], b1 = two ? bounds[yo0 + 1] : make_int2(b0.x, 0);
  const int shift = b1.x - b0.x;                             // >= 0: the windows move down with the output row
  const int span = max(b0.y, shift + b1.y);
  for (int j = threadIdx.x; j < span; j += blockDim.x) {
    const int t1 = j - shift;
    s_k[j] = make_int2(j < b0.y ? __ldg(kk + j * hout + yo0) : 0,
                       (t1 >= 0 && t1 < b1.y) ? __ldg(kk + t1 * hout + yo0 + 1) : 0);
  }
  __syncthreads();
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  if (xo >= wout) return;
  int a0r = 1 << (kPrecisionBits - 1), a0g = a0r, a0b = a0r, a1r = a0r, a1g = a0r, a1b = a0r;
  const uint32_t* p = tmp + (static_cast<long long>(b) * ht + b0.x) * wout + xo;
#pragma unroll 4
  for (int j = 0; j < span; ++j, p += wout) {
    const uint32_t v = *p;
    const int2 k = s_k[j];
    const int cr = static_cast<int>(__byte_perm(v, 0u, 0x4440)), cg = static_cast<int>(__byte_perm(v, 0u, 0x4441)),
              cb = static_cast<int>(__byte_perm(v, 0u, 0x4442));
    a0r += cr * k.x; a0g += cg * k.x; a0b += cb * k.x;
    a1r += cr * k.y; a1g += cg * k.y; a1b += cb * k.y;
  }
  const long long plane = static_cast<long long>(hout) * wout;
  float* o = out + static_cast<long long>(b) * 3 * plane + static_cast<long long>(yo0) * wout + xo;
  o[0] = static_cast<float>(clip8(a0r)) / 255.0f;            // ToTensor: IEEE division, not a reciprocal multiply
  o[plane] = static_cast<float>(clip8(a0g)) / 255.0f;
  o[2 * plane] = static_cast<float>(clip8(a0b)) / 255.0f;
  if (two) {
    o[wout] = static_cast<float>(clip8(a1r)) / 255.0f;
    o[plane + wout] = static_cast<float>(clip8(a1g)) / 255.0f;
    o[2 * plane + wout] = static_cast<float>(clip8(a1b)) / 255.0f;
  }
}

template <int KS>
int launch_h_rows(const uint8_t* src, int batch, int hin, int win, int y0, int ht, int wout, const int2* bounds,
                  const int32_t* taps, uchar4* tmp, cudaStream_t st) {
  const long long row_bytes = static_cast<long long>(win) * 4;
  long long fit = (48 * 1024 - KS * 4) / row_bytes;
  int rows = static_cast<int>(fit < kRowsPerCta ? fit : kRowsPerCta);
  if (rows < 1) rows = 1;
  const long long smem = rows * row_bytes + KS * 4;
  if (smem > 200 * 1024) return FP_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    const int rc = fp::ensure_max_smem(reinterpret_cast<const void*>(resize_h_rows_kernel<KS>), static_cast<int>(smem),
                                       "resize_h_rows_kernel");
    if (rc) return rc;
  }
  resize_h_rows_kernel<KS><<<dim3((ht + rows - 1) / rows, batch), 256, smem, st>>>(src, hin, win, y0, ht, wout, rows, bounds, taps, tmp);
  return FP_OK;
}

}  // namespace

extern "C" int fp_resize_u8_horizontal(const void* src, int batch, int hin, int win, int y0, int ht, int wout,
                                       const int32_t* bounds, const int32_t* taps, int ksize, void* tmp, void* stream) {
  if (!src || !bounds || !taps || !tmp) return set_error(FP_ERR_ARG, "fp_resize_u8_horizontal: null argument");
  if (batch < 1 || hin < 1 || win < 1 || wout < 1 || ht < 1 || y0 < 0 || y0 + ht > hin || ksize < 1)
    return set_error(FP_ERR_ARG, "fp_resize_u8_horizontal: bad geometry (batch %d, %dx%d, rows %d+%d, wout %d)", batch, hin, win,
                     y0, ht, wout);
  if (batch > 65535) return set_error(FP_ERR_ARG, "fp_resize_u8_horizontal: batch %d > 65535", batch);
  if (fp::opts().resize_rows != 0 && ksize <= 32 && static_cast<long long>(batch) * hin * win * 3 < (1ll << 40)) {
    int rc = FP_ERR_UNSUPPORTED;
    const uint8_t* s8 = static_cast<const uint8_t*>(src);
    const int2* bd = reinterpret_cast<const int2*>(bounds);
    uchar4* t4 = static_cast<uchar4*>(tmp);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (ksize) {
#define FP_H_CASE(K) case K: rc = launch_h_rows<K>(s8, batch, hin, win, y0, ht, wout, bd, taps, t4, st); break;
      FP_H_CASE(1) FP_H_CASE(2) FP_H_CASE(3) FP_H_CASE(4) FP_H_CASE(5) FP_H_CASE(6) FP_H_CASE(7) FP_H_CASE(8)
      FP_H_CASE(9) FP_H_CASE(10) FP_H_CASE(11) FP_H_CASE(12) FP_H_CASE(13) FP_H_CASE(14) FP_H_CASE(15) FP_H_CASE(16)
      FP_H_CASE(17) FP_H_CASE(18) FP_H_CASE(19) FP_H_CASE(20) FP_H_CASE(21) FP_H_CASE(22) FP_H_CASE(23) FP_H_CASE(24)
      FP_H_CASE(25) FP_H_CASE(26) FP_H_CASE(27) FP_H_CASE(28) FP_H_CASE(29) FP_H_CASE(30) FP_H_CASE(31) FP_H_CASE(32)
#undef FP_H_CASE
      default: break;
    }
    if (rc != FP_ERR_UNSUPPORTED) {
      if (rc) return rc;
      FP_CHECK_LAUNCH("fp_resize_u8_horizontal");
      return FP_OK;
    }
  }
  const int smem = ((15 + win * 3 + 15) / 16 + 1) * 16;
  if (smem > 48 * 1024) {
    int rc = fp::ensure_max_smem(reinterpret_cast<const void*>(resize_h_kernel), smem, "resize_h_kernel");
    if (rc) return rc;
  }
  const long long total = static_cast<long long>(batch) * hin * win * 3;
  resize_h_kernel<<<dim3(ht, batch), 256, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), total, hin, win, y0, ht, wout, reinterpret_cast<const int2*>(bounds), taps,
      static_cast<uchar4*>(tmp));
  FP_CHECK_LAUNCH("fp_resize_u8_horizontal");
  return FP_OK;
}

extern "C" int fp_resize_u8_vertical_to_f32(const void* tmp, int batch, int ht, int wout, int hout, const int32_t* bounds,
                                            const int32_t* taps, int ksize, float* out, void* stream) {
  if (!tmp || !bounds || !taps || !out) return set_error(FP_ERR_ARG, "fp_resize_u8_vertical_to_f32: null argument");
  if (batch < 1 || ht < 1 || wout < 1 || hout < 1 || ksize < 1 || hout > 65535 || batch > 65535)
    return set_error(FP_ERR_ARG, "fp_resize_u8_vertical_to_f32: bad geometry (batch %d, ht %d, %dx%d)", batch, ht, hout, wout);
  if (fp::opts().resize_rows != 0 && ksize <= kVSpanMax / 2 && (hout + 1) / 2 <= 65535) {
    // two adjacent rows' windows overlap in all but `shift` rows: span = max(n0, shift + n1) <= 2 * ksize <= kVSpanMax
    resize_v_rows_kernel<<<dim3((wout + 127) / 128, (hout + 1) / 2, batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint32_t*>(tmp), ht, wout, hout, reinterpret_cast<const int2*>(bounds), taps, out);
    FP_CHECK_LAUNCH("fp_resize_u8_vertical_to_f32");
    return FP_OK;
  }
  resize_v_kernel<<<dim3((wout + 127) / 128, hout, batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uchar4*>(tmp), ht, wout, hout, reinterpret_cast<const int2*>(bounds), taps, out);
  FP_CHECK_LAUNCH("fp_resize_u8_vertical_to_f32");
  return FP_OK;
}
