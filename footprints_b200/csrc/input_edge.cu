// Input edge of the hot path: uint8 RGB frames -> the float32 NCHW tensor FootprintNetwork.forward takes.
//
// Replaces, bit for bit, what the reference does on the host for every frame
//   transforms.Resize((H, W), interpolation=Image.ANTIALIAS)   predict_simple.py:41-42,54; datasets/inference_dataset.py:25-27,46
//   transforms.ToTensor()                                       predict_simple.py:43,55;    datasets/inference_dataset.py:47
// i.e. Pillow's two-pass Lanczos (a=3) resampling in 22-bit fixed point with a uint8 intermediate image, then /255.
// The per-geometry tap tables (bounds + fixed-point weights) are built once by the host (footprints_b200/preprocess.py,
// double precision like Pillow's precompute_coeffs) and stay resident; the pixel work is the two kernels below.
//
// Both kernels are byte/integer streaming work (HBM/L2 bound, no tensor cores): the horizontal pass stages one source row
// in shared memory with 16-byte loads and writes an RGBX (uchar4) intermediate so that the vertical pass reads and writes
// fully coalesced 4-byte lanes.  The intermediate (Ht x Wout x 4 B per frame) is consumed from L2.
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

}  // namespace

extern "C" int fp_resize_u8_horizontal(const void* src, int batch, int hin, int win, int y0, int ht, int wout,
                                       const int32_t* bounds, const int32_t* taps, int ksize, void* tmp, void* stream) {
  if (!src || !bounds || !taps || !tmp) return set_error(FP_ERR_ARG, "fp_resize_u8_horizontal: null argument");
  if (batch < 1 || hin < 1 || win < 1 || wout < 1 || ht < 1 || y0 < 0 || y0 + ht > hin || ksize < 1)
    return set_error(FP_ERR_ARG, "fp_resize_u8_horizontal: bad geometry (batch %d, %dx%d, rows %d+%d, wout %d)", batch, hin, win,
                     y0, ht, wout);
  if (batch > 65535) return set_error(FP_ERR_ARG, "fp_resize_u8_horizontal: batch %d > 65535", batch);
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
  resize_v_kernel<<<dim3((wout + 127) / 128, hout, batch), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uchar4*>(tmp), ht, wout, hout, reinterpret_cast<const int2*>(bounds), taps, out);
  FP_CHECK_LAUNCH("fp_resize_u8_vertical_to_f32");
  return FP_OK;
}
