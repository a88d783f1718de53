// Shared im2col gather machinery of the forward/dgrad (conv_igemm.cu) and wgrad (conv_wgrad.cu) kernels.
//
// A convolution's A operand row m (an output pixel) x 16-byte K chunk (8 channels of one filter tap) is fetched
// with one cp.async from one of up to two NHWC sources.  Per output pixel a "row info" word pair encodes where the
// taps land in the source: the element offset of a reference source pixel plus 4-bit per-tap row/column deltas
// (15 = the tap reads padding zeros).  That encoding covers zero and reflection padding, stride, the nearest x2
// upsample of source 0 (footprints/network.py:98,154), the channel concat (network.py:155) and — for the data
// gradient of stride-2 convolutions — the zero-insertion of a transposed convolution.
#pragma once
#include "fp_ptx.cuh"
#include "fp_internal.h"

namespace fp {

struct GatherGeom {
  const uint8_t* src0;
  const uint8_t* src1;
  int C0, C1, H0, W0;
  int up0;                   // 0: as stored, 1: nearest x2 upsample, 2: zero-inserted x2 (transposed stride 2)
  int Hin, Win;              // virtual input size (after upsample) == source 1 size
  int cin_log2, taps, KH, KW, stride, pad_t, pad_l, pad_mode;
  int Ho, Wo, HoWo, M;
  // split-precision gather (fp_conv_fwd_split): every source is a pair of 16-bit tensors hi + lo (~22 significant bits)
  // and the K axis per tap is [hi(src0,src1) | lo(src0,src1) | hi(src0,src1)] against weights [Whi | Whi | Wlo]
  const uint8_t* src0_lo;
  const uint8_t* src1_lo;
  int split;
};

__device__ __forceinline__ void make_row_info(int n, int y0, int x0, int Hv, int Wv, int up, int Hs, int Ws, int Cs,
                                              int KH, int KW, int reflect, uint32_t& base, uint32_t& code) {
  const int sh = up ? 1 : 0;
  const int refy = max(y0, 0) >> sh, refx = max(x0, 0) >> sh;
  base = static_cast<uint32_t>(((n * Hs + refy) * Ws + refx) * Cs);   // Cs = 1: pixel index (TMA gather rows)
  uint32_t cd = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    int vy = y0 + t, vx = x0 + t;
    bool oky = t < KH, okx = t < KW;
    if (up == 2) {           // transposed stride 2: only even virtual positions carry data
      oky = oky && vy >= 0 && !(vy & 1) && (vy >> 1) < Hs;
      okx = okx && vx >= 0 && !(vx & 1) && (vx >> 1) < Ws;
    } else if (reflect) {
      vy = vy < 0 ? -vy : (vy >= Hv ? 2 * Hv - 2 - vy : vy);
      vx = vx < 0 ? -vx : (vx >= Wv ? 2 * Wv - 2 - vx : vx);
    } else {
      oky = oky && static_cast<unsigned>(vy) < static_cast<unsigned>(Hv);
      okx = okx && static_cast<unsigned>(vx) < static_cast<unsigned>(Wv);
    }
    const uint32_t dy = oky ? static_cast<uint32_t>((vy >> sh) - refy) : 15u;
    const uint32_t dx = okx ? static_cast<uint32_t>((vx >> sh) - refx) : 15u;
    cd |= (dy & 15u) << (4 * t);
    cd |= (dx & 15u) << (16 + 4 * t);
  }
  code = cd;
}

// {base0, code0, base1, code1} for output pixel m (caller guarantees m < g.M); pixel_units: bases are pixel
// indices (row coordinates of the TMA gather tensor maps) instead of element offsets
__device__ __forceinline__ uint4 make_row_info_for_pixel(const GatherGeom& g, int m, bool pixel_units = false) {
  const int n = m / g.HoWo;
  const int rem = m - n * g.HoWo;
  const int oy = rem / g.Wo;
  const int ox = rem - oy * g.Wo;
  const int y0 = oy * g.stride - g.pad_t, x0 = ox * g.stride - g.pad_l;
  uint4 ri = make_uint4(0u, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
  make_row_info(n, y0, x0, g.Hin, g.Win, g.up0, g.H0, g.W0, pixel_units ? 1 : g.C0, g.KH, g.KW, g.pad_mode, ri.x, ri.y);
  if (g.C1 > 0)
    make_row_info(n, y0, x0, g.Hin, g.Win, 0, g.Hin, g.Win, pixel_units ? 1 : g.C1, g.KH, g.KW, g.pad_mode, ri.z, ri.w);
  return ri;
}

// K-chunk table entry: chunk j covers channels [8j, 8j+8) of the tap-major K axis.
//   [0,5) row-code shift (4*th)   [5,10) col-code shift (16+4*tw)   bit 10 source 1   bit 11 chunk is K padding
//   [12,32) channel offset inside the source pixel
//   bit 31: the chunk reads the LOW halves of the split-precision sources
__device__ __forceinline__ uint32_t make_ktab_entry(const GatherGeom& g, int j, int num_chunks) {
  const int ch0 = j * 8;
  int tap, cc;
  uint32_t lo = 0u;
  if (g.split) {
    const int cin = g.C0 + g.C1;
    tap = ch0 / (3 * cin);
    const int r = ch0 - tap * 3 * cin;
    const int part = r / cin;                   // 0: hi x Whi, 1: lo x Whi, 2: hi x Wlo
    cc = r - part * cin;
    lo = part == 1 ? 1u : 0u;
  } else {
    tap = ch0 >> g.cin_log2;
    cc = ch0 & ((g.C0 + g.C1) - 1);
  }
  if (j >= num_chunks || tap >= g.taps) return 1u << 11;
  const int th = tap / g.KW, tw = tap - th * g.KW;
  const uint32_t second = cc >= g.C0 ? 1u : 0u;
  if (second) cc -= g.C0;
  return static_cast<uint32_t>(4 * th) | (static_cast<uint32_t>(16 + 4 * tw) << 5) | (second << 10) |
         (static_cast<uint32_t>(cc) << 12) | (lo << 31);
}

struct KDecode {
  const uint8_t* sbase;
  uint32_t pstride, rstride, shy, shx, cc, padmask;
  bool second;
};

__device__ __forceinline__ KDecode decode_ktab(const GatherGeom& g, uint32_t ke) {
  KDecode k;
  k.second = (ke & (1u << 10)) != 0u;
  k.padmask = (ke >> 8) & 8u;                                   // bit 11 -> 8: forces "invalid"
  k.shy = ke & 31u; k.shx = (ke >> 5) & 31u; k.cc = (ke >> 12) & 0x7FFFFu;
  k.sbase = (ke >> 31) ? (k.second ? g.src1_lo : g.src0_lo) : (k.second ? g.src1 : g.src0);
  k.pstride = k.second ? g.C1 : g.C0;                           // elements per source pixel
  k.rstride = k.second ? g.Win * g.C1 : g.W0 * g.C0;            // elements per source row
  return k;
}

// element offset of the 8-channel chunk for the row described by `ri`; ok=false -> zero fill
__device__ __forceinline__ uint32_t gather_offset(const KDecode& k, const uint4& ri, bool& ok) {
  const uint32_t code = k.second ? ri.w : ri.y;
  const uint32_t base = k.second ? ri.z : ri.x;
  const uint32_t dy = (code >> k.shy) & 15u, dx = (code >> k.shx) & 15u;
  ok = ((dy | dx | k.padmask) & 8u) == 0u;
  return ok ? base + dy * k.rstride + dx * k.pstride + k.cc : 0u;
}

// TMA-gather flavour: source row (pixel) index of the chunk's tap for the row described by `ri`, or `oob_row`
// (any index >= the tensor's row count: the TMA fills zeros) when the tap reads padding
__device__ __forceinline__ int gather_row_index(const KDecode& k, const uint4& ri, int row_pitch_pixels, int oob_row) {
  const uint32_t code = k.second ? ri.w : ri.y;
  const uint32_t base = k.second ? ri.z : ri.x;
  const uint32_t dy = (code >> k.shy) & 15u, dx = (code >> k.shx) & 15u;
  const bool ok = ((dy | dx | k.padmask) & 8u) == 0u;
  return ok ? static_cast<int>(base + dy * row_pitch_pixels + dx) : oob_row;
}

// host: validate an FpConvDesc's gather part and fill the geometry
// need_pow2 = false: callers that never decode k -> (tap, channel) by shifts (the row-strip kernels) accept any c0+c1
int fill_gather_geom(const FpConvDesc* d, const void* src0, const void* src1, const char* who, GatherGeom& g,
                     bool need_pow2 = true);

}  // namespace fp
