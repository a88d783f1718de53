// Dedicated CUDA-core kernels for the 2-channel output heads (OutConvBlock, footprints/network.py:161-183): a
// 3x3 reflection-padded convolution Cin -> 2.  With N = 2 the tensor-core implicit GEMM wastes >98 % of every
// MMA, so forward and data gradient are plain FMA kernels here.  The default paths are the shared-memory tiled
// head_fwd_tile_kernel / head_dgrad_tile_kernel (16 x 32 pixel tiles, 4 pixels x 16 channels per thread, FMA-issue
// bound); the per-pixel / per-chunk kernels remain as the FP_HEAD_TILE=0 variants and for C % 16 != 0.  The weight
// gradient of the heads runs on the tensor-core row-strip kernel (conv_wgrad.cu); head_wgrad_kernel is the CUDA-core
// reference implementation the tests pin it against.
#include "fp_internal.h"
#include "fp_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

namespace fp {

__device__ __forceinline__ float2 h_up2(uint32_t u, int fmt) {
  if (fmt == FP_BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ uint32_t h_pk2(float a, float b, int fmt) {
  if (fmt == FP_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return fp::f32x2_to_f16x2_sat(a, b);
}
__device__ __forceinline__ void h_unpack8(const uint4& v, int fmt, float (&f)[8]) {
  float2 t;
  t = h_up2(v.x, fmt); f[0] = t.x; f[1] = t.y;
  t = h_up2(v.y, fmt); f[2] = t.x; f[3] = t.y;
  t = h_up2(v.z, fmt); f[4] = t.x; f[5] = t.y;
  t = h_up2(v.w, fmt); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ int reflect1(int v, int n) { return v < 0 ? -v : (v >= n ? 2 * n - 2 - v : v); }

constexpr int HEAD_THREADS = 256;
constexpr int HEAD_MAX_C = 128;

// ---------------------------------------------------------------------------------------------- forward
// One thread per (output pixel, 8-channel chunk): the c8 lanes of a pixel read 16 consecutive bytes each (coalesced),
// accumulate their channels' share of both outputs and combine with a shuffle tree.  Weights sit in shared memory as
// [tap][i][chunk] so that the lanes of a warp hit consecutive addresses.
__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_kernel(const uint4* __restrict__ x, const float* __restrict__ w /*[2][C][3][3]*/, const float* __restrict__ bias,
                int batch, int H, int W, int C, int sigmoid, int fmt, float2* __restrict__ out) {
  __shared__ float2 sw[9 * HEAD_MAX_C];
  const int c8 = C >> 3;
  for (int i = threadIdx.x; i < 9 * C; i += HEAD_THREADS) {
    const int t = i / C, c = i - t * C;
    sw[(t * 8 + (c & 7)) * c8 + (c >> 3)] = make_float2(w[(0 * C + c) * 9 + t], w[(1 * C + c) * 9 + t]);
  }
  __syncthreads();
  const long long idx = blockIdx.x * static_cast<long long>(HEAD_THREADS) + threadIdx.x;
  const long long P = static_cast<long long>(batch) * H * W;
  const long long pix = idx / c8;
  const int cc = static_cast<int>(idx - pix * c8);
  const bool valid = pix < P;
  float a0 = 0.f, a1 = 0.f;
  if (valid) {
    const int px = static_cast<int>(pix % W);
    const int py = static_cast<int>((pix / W) % H);
    const long long img = pix / (static_cast<long long>(W) * H) * H;
#pragma unroll
    for (int ty = 0; ty < 3; ++ty) {
      const int yy = reflect1(py + ty - 1, H);
#pragma unroll
      for (int tx = 0; tx < 3; ++tx) {
        const int xx = reflect1(px + tx - 1, W);
        float f[8];
        h_unpack8(x[((img + yy) * W + xx) * c8 + cc], fmt, f);
        const float2* wt = sw + (ty * 3 + tx) * 8 * c8 + cc;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ww = wt[i * c8];
          a0 = fmaf(f[i], ww.x, a0);
          a1 = fmaf(f[i], ww.y, a1);
        }
      }
    }
  }
  // c8 is 1, 2, 4, 8 or 16: lanes of one pixel are contiguous and aligned inside the warp
  for (int off = c8 >> 1; off > 0; off >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, off);
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
  }
  if (valid && cc == 0) {
    a0 += bias[0];
    a1 += bias[1];
    if (sigmoid) {
      a0 = 1.f / (1.f + expf(-a0));
      a1 = 1.f / (1.f + expf(-a1));
    }
    out[pix] = make_float2(a0, a1);
  }
}

// Variant for C <= 32: one thread per output pixel (64 contiguous bytes per tap and thread, weights broadcast from
// shared memory).  Measured 9x faster than the chunk-per-thread mapping at 192x640x32 (HBM-bound), slower for C >= 64.
__global__ void __launch_bounds__(HEAD_THREADS)
head_fwd_pixel_kernel(const uint4* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, int batch,
                      int H, int W, int C, int sigmoid, int fmt, float2* __restrict__ out) {
  __shared__ float2 sw[9 * 32];
  for (int i = threadIdx.x; i < 9 * C; i += HEAD_THREADS) {
    const int t = i / C, c = i - t * C;
    sw[i] = make_float2(w[(0 * C + c) * 9 + t], w[(1 * C + c) * 9 + t]);
  }
  __syncthreads();
  const long long idx = blockIdx.x * static_cast<long long>(HEAD_THREADS) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * H * W) return;
  const int px = static_cast<int>(idx % W);
  const int py = static_cast<int>((idx / W) % H);
  const long long img = idx / (static_cast<long long>(W) * H) * H;
  const int c8 = C >> 3;
  float a0 = bias[0], a1 = bias[1];
#pragma unroll
  for (int ty = 0; ty < 3; ++ty) {
    const int yy = reflect1(py + ty - 1, H);
#pragma unroll
    for (int tx = 0; tx < 3; ++tx) {
      const int xx = reflect1(px + tx - 1, W);
      const uint4* src = x + ((img + yy) * W + xx) * c8;
      const float2* wt = sw + (ty * 3 + tx) * C;
      for (int cc = 0; cc < c8; ++cc) {
        float f[8];
        h_unpack8(src[cc], fmt, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ww = wt[cc * 8 + i];
          a0 = fmaf(f[i], ww.x, a0);
          a1 = fmaf(f[i], ww.y, a1);
        }
      }
    }
  }
  if (sigmoid) {
    a0 = 1.f / (1.f + expf(-a0));
    a1 = 1.f / (1.f + expf(-a1));
  }
  out[idx] = make_float2(a0, a1);
}

// Shared-memory tiled forward (the default): a block of 128 threads owns a 16 x 32 tile of output pixels and walks the
// channels in chunks of 32.  Per chunk the (16+2) x (32+2) reflected halo tile is staged once (coalesced 16-byte
// loads; 80-byte pixel stride = 64 data + 16 pad bytes, so that the 32 lanes of a warp -- consecutive pixels -- hit all
// banks on 16-byte reads), then each thread produces 4 vertically adjacent pixels: for every (8-channel group, dx) it
// reads 6 input rows once and applies the three dy taps from registers, so a weight pair fetched from shared memory is
// used for 4 pixels and an input value for up to 3 -- FMA-bound instead of L1-wavefront-bound (the per-pixel kernels
// above issue 16 wavefronts per 64-byte-strided global load).
constexpr int HT_ROWS = 16, HT_COLS = 32, HT_THREADS = 128, HT_CB = 32;
constexpr int HT_PIX_BYTES = HT_CB * 2 + 16;
constexpr int HT_HALO_PIX = (HT_ROWS + 2) * (HT_COLS + 2);
constexpr int HT_SMEM_BYTES = HT_HALO_PIX * HT_PIX_BYTES + 9 * HT_CB * static_cast<int>(sizeof(float2));

__global__ void __launch_bounds__(HT_THREADS)
head_fwd_tile_kernel(const uint4* __restrict__ x, const float* __restrict__ w /*[2][C][3][3]*/,
                     const float* __restrict__ bias, int H, int W, int C, int tiles_x, int tiles_y, int sigmoid, int fmt,
                     float2* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t ht_smem[];
  uint8_t* tile = ht_smem;
  float2* sw = reinterpret_cast<float2*>(ht_smem + HT_HALO_PIX * HT_PIX_BYTES);     // [tap][32]
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int tx0 = (bid % tiles_x) * HT_COLS; bid /= tiles_x;
  const int ty0 = (bid % tiles_y) * HT_ROWS;
  const long long img = bid / tiles_y;
  const int c8 = C >> 3;
  const int lx = tid & 31, ly = (tid >> 5) * 4;          // this thread: column lx, rows ly .. ly+3 of the tile
  float acc[4][2];
#pragma unroll
  for (int r = 0; r < 4; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }

  for (int cb = 0; cb < C; cb += HT_CB) {
    const int nch = min(HT_CB, C - cb);                  // 8, 16, 24 or 32 channels in this chunk
    const int nq = nch >> 3;
    __syncthreads();                                     // previous chunk fully consumed
    for (int i = tid; i < HT_HALO_PIX * nq; i += HT_THREADS) {
      const int pix = i / nq, q = i - pix * nq;
      const int hy = pix / (HT_COLS + 2), hx = pix - hy * (HT_COLS + 2);
      int yy = reflect1(ty0 + hy - 1, H), xx = reflect1(tx0 + hx - 1, W);
      yy = min(max(yy, 0), H - 1); xx = min(max(xx, 0), W - 1);   // rows / columns past a ragged tile: any valid pixel
      *reinterpret_cast<uint4*>(tile + pix * HT_PIX_BYTES + q * 16) = __ldg(&x[((img * H + yy) * W + xx) * c8 + (cb >> 3) + q]);
    }
    for (int i = tid; i < 9 * nch; i += HT_THREADS) {
      const int t = i / nch, c = i - t * nch;
      sw[t * HT_CB + c] = make_float2(w[(0 * C + cb + c) * 9 + t], w[(1 * C + cb + c) * 9 + t]);
    }
    __syncthreads();
    for (int q = 0; q < nq; ++q) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float f[6][8];
#pragma unroll
        for (int r = 0; r < 6; ++r)
          h_unpack8(*reinterpret_cast<const uint4*>(tile + ((ly + r) * (HT_COLS + 2) + lx + dx) * HT_PIX_BYTES + q * 16), fmt, f[r]);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float2* wt = sw + (dy * 3 + dx) * HT_CB + q * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 ww = wt[i];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              acc[r][0] = fmaf(f[r + dy][i], ww.x, acc[r][0]);
              acc[r][1] = fmaf(f[r + dy][i], ww.y, acc[r][1]);
            }
          }
        }
      }
    }
  }
  const float b0 = bias[0], b1 = bias[1];
  const int px = tx0 + lx;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int py = ty0 + ly + r;
    if (py < H && px < W) {
      float a0 = acc[r][0] + b0, a1 = acc[r][1] + b1;
      if (sigmoid) {
        a0 = 1.f / (1.f + expf(-a0));
        a1 = 1.f / (1.f + expf(-a1));
      }
      out[(img * H + py) * W + px] = make_float2(a0, a1);
    }
  }
}

// Same idea for the data gradient at C == 32: one thread per pixel owns all 32 channels, so each dZ value is loaded
// once per tap instead of once per chunk.
__global__ void __launch_bounds__(HEAD_THREADS)
head_dgrad_pixel32_kernel(const uint4* __restrict__ dz8, const float* __restrict__ w, int batch, int H, int W,
                          int accumulate, const uint4* __restrict__ yact, int act, int act_fmt, int grad_fmt,
                          uint4* __restrict__ dst) {
  constexpr int C = 32;
  __shared__ float2 sw[9 * C];
  for (int i = threadIdx.x; i < 9 * C; i += HEAD_THREADS) {
    const int t = i / C, c = i - t * C;
    sw[i] = make_float2(w[(0 * C + c) * 9 + t], w[(1 * C + c) * 9 + t]);
  }
  __syncthreads();
  const long long idx = blockIdx.x * static_cast<long long>(HEAD_THREADS) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * H * W) return;
  const int px = static_cast<int>(idx % W);
  const int py = static_cast<int>((idx / W) % H);
  const long long img = idx / (static_cast<long long>(W) * H) * H;
  int vy[2], vx[2], ny = 1, nx = 1;
  vy[0] = py; vx[0] = px;
  if (py == 1) vy[ny++] = -1;
  if (py == H - 2) vy[ny++] = H;
  if (px == 1) vx[nx++] = -1;
  if (px == W - 2) vx[nx++] = W;
  float acc[C];
#pragma unroll
  for (int i = 0; i < C; ++i) acc[i] = 0.f;
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        const int qy = vy[iy] - ty + 1;
        if (qy < 0 || qy >= H) continue;
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          const int qx = vx[ix] - tx + 1;
          if (qx < 0 || qx >= W) continue;
          const float2 g = h_up2(reinterpret_cast<const uint32_t*>(dz8 + (img + qy) * W + qx)[0], grad_fmt);
          const float2* wt = sw + (ty * 3 + tx) * C;
#pragma unroll
          for (int i = 0; i < C; ++i) acc[i] = fmaf(wt[i].x, g.x, fmaf(wt[i].y, g.y, acc[i]));
        }
      }
    }
#pragma unroll
  for (int cc = 0; cc < C / 8; ++cc) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = acc[cc * 8 + i];
    if (accumulate) {
      float f[8];
      h_unpack8(dst[idx * (C / 8) + cc], grad_fmt, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] += f[i];
    }
    if (yact != nullptr) {
      float fy[8];
      h_unpack8(yact[idx * (C / 8) + cc], act_fmt, fy);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (act == FP_ACT_ELU) a[i] *= fy[i] > 0.f ? 1.f : fy[i] + 1.f;
        else if (act == FP_ACT_RELU) a[i] = fy[i] > 0.f ? a[i] : 0.f;
      }
    }
    uint4 o;
    o.x = h_pk2(a[0], a[1], grad_fmt); o.y = h_pk2(a[2], a[3], grad_fmt);
    o.z = h_pk2(a[4], a[5], grad_fmt); o.w = h_pk2(a[6], a[7], grad_fmt);
    dst[idx * (C / 8) + cc] = o;
  }
}

// ---------------------------------------------------------------------------------------------- data gradient
// dX[p][c] = sum over the virtual (reflection-padded) positions v that map onto p, taps t: W[co][c][t] * dZ[v - t + 1][co],
// optionally (+ previous content of dst) and * act'(y) of the layer that produced x.  One thread per (pixel, 8 channels).
__global__ void __launch_bounds__(HEAD_THREADS)
head_dgrad_kernel(const uint4* __restrict__ dz8 /*[B][H][W][8] 16-bit, ch 0..1 valid*/, const float* __restrict__ w,
                  int batch, int H, int W, int C, int accumulate, const uint4* __restrict__ yact, int act, int act_fmt,
                  int grad_fmt, uint4* __restrict__ dst) {
  __shared__ float2 sw[9 * HEAD_MAX_C];
  const int c8 = C >> 3;
  for (int i = threadIdx.x; i < 9 * C; i += HEAD_THREADS) {
    const int t = i / C, c = i - t * C;
    sw[(t * 8 + (c & 7)) * c8 + (c >> 3)] = make_float2(w[(0 * C + c) * 9 + t], w[(1 * C + c) * 9 + t]);
  }
  __syncthreads();
  const long long idx = blockIdx.x * static_cast<long long>(HEAD_THREADS) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * H * W * c8) return;
  const int cc = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int px = static_cast<int>(t % W); t /= W;
  const int py = static_cast<int>(t % H);
  const long long img = t / H * H;
  // virtual rows / columns (in [-1, H]) that reflect onto (py, px)
  int vy[2], vx[2], ny = 1, nx = 1;
  vy[0] = py; vx[0] = px;
  if (py == 1) vy[ny++] = -1;
  if (py == H - 2) vy[ny++] = H;
  if (px == 1) vx[nx++] = -1;
  if (px == W - 2) vx[nx++] = W;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (int iy = 0; iy < ny; ++iy)
    for (int ix = 0; ix < nx; ++ix) {
#pragma unroll
      for (int ty = 0; ty < 3; ++ty) {
        const int qy = vy[iy] - ty + 1;
        if (qy < 0 || qy >= H) continue;
#pragma unroll
        for (int tx = 0; tx < 3; ++tx) {
          const int qx = vx[ix] - tx + 1;
          if (qx < 0 || qx >= W) continue;
          const float2 g = h_up2(reinterpret_cast<const uint32_t*>(dz8 + (img + qy) * W + qx)[0], grad_fmt);
          const float2* wt = sw + (ty * 3 + tx) * 8 * c8 + cc;
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(wt[i * c8].x, g.x, fmaf(wt[i * c8].y, g.y, acc[i]));
        }
      }
    }
  if (accumulate) {
    float f[8];
    h_unpack8(dst[idx], grad_fmt, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += f[i];
  }
  if (yact != nullptr) {
    float fy[8];
    h_unpack8(yact[idx], act_fmt, fy);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (act == FP_ACT_ELU) acc[i] *= fy[i] > 0.f ? 1.f : fy[i] + 1.f;
      else if (act == FP_ACT_RELU) acc[i] = fy[i] > 0.f ? acc[i] : 0.f;
    }
  }
  uint4 o;
  o.x = h_pk2(acc[0], acc[1], grad_fmt); o.y = h_pk2(acc[2], acc[3], grad_fmt);
  o.z = h_pk2(acc[4], acc[5], grad_fmt); o.w = h_pk2(acc[6], acc[7], grad_fmt);
  dst[idx] = o;
}

// Shared-memory tiled data gradient (the default for C % 16 == 0): a block of 128 threads owns a 16 x 32 pixel tile.
// Stage 1 builds G[tap][pixel] = sum of dZ over the virtual positions that reflect onto the pixel (the adjoint of the
// reflection padding, 1 / 2 / 4 terms) once per tile; stage 2 is a plain FMA loop dX[p][c] = sum_tap G[tap][p] . W[.][c][tap]
// with every thread producing 4 pixels x 16 channels per pass (weight pairs broadcast from shared memory are used for 4
// pixels, G values for 16 channels) and writing whole 32-byte sectors.
constexpr int HD_SMEM_G = 9 * HT_ROWS * HT_COLS * static_cast<int>(sizeof(float2));     // 36 KB
constexpr int HD_SMEM_DZ = (HT_ROWS + 2) * (HT_COLS + 2) * static_cast<int>(sizeof(float2));

// HM = true (fp16 gradients): the FMA loop runs on packed half2 arithmetic -- the kernel is FMA-issue bound (18 terms per
// output element; 22 % of the fp32 FMA peak with the scalar loop), HFMA2 halves the instruction count; the 18-term sums in
// fp16 carry ~2e-3 relative error, the accuracy class of the 16-bit gradient tensors they are stored into.  G / dZ tiles
// then hold {half2(gx, gx), half2(gy, gy)} and the weights {half2(w0[c], w0[c+1]), half2(w1[c], w1[c+1])} per channel pair.
// EP = 1 (default; FP_HEAD_EPI=0 selects 0): the epilogue's operands (the activation for ELU' / ReLU', the gradient to
// accumulate into) are requested BEFORE the FMA loop of their channel block, so their DRAM latency hides behind it -- with
// the loads after the loop the kernel spent 40-64 % of its issue slots stalled on them (ncu, profiles/r02F_head_dgrad.md) --
// and, with fp16 activations and gradients, the epilogue stays in packed half2 (clamp, ELU' factor min(y + 1, 1), product,
// accumulate: 2.5 instead of ~8.5 instructions per element).
template <bool HM, int EP, bool ACC>
__global__ void __launch_bounds__(HT_THREADS, (HM && EP) ? (ACC ? 3 : 4) : 1)
head_dgrad_tile_kernel(const uint4* __restrict__ dz8, const float* __restrict__ w, int H, int W, int C, int tiles_x,
                       int tiles_y, const uint4* __restrict__ yact, int act, int act_fmt, int grad_fmt,
                       uint4* __restrict__ dst) {
  extern __shared__ __align__(16) uint8_t hd_smem[];
  float2* G = reinterpret_cast<float2*>(hd_smem);                       // [9][16*32]
  float2* dzs = reinterpret_cast<float2*>(hd_smem + HD_SMEM_G);         // [18][34] dZ halo tile
  float2* sw = dzs + (HT_ROWS + 2) * (HT_COLS + 2);                     // [9][C]
  const int tid = threadIdx.x;
  int bid = blockIdx.x;
  const int tx0 = (bid % tiles_x) * HT_COLS; bid /= tiles_x;
  const int ty0 = (bid % tiles_y) * HT_ROWS;
  const long long img = bid / tiles_y;
  const int c8 = C >> 3;
  auto pack_g = [](float2 v) -> float2 {          // fp32 pair, or the two broadcast half2 registers of the packed loop
    if (!HM) return v;
    float2 o;                                   // saturating: a sum of up to four saturated dZ values must not become inf
    o.x = __uint_as_float(f32x2_to_f16x2_sat(v.x, v.x));
    o.y = __uint_as_float(f32x2_to_f16x2_sat(v.y, v.y));
    return o;
  };
  if (HM) {
    for (int i = tid; i < 9 * (C >> 1); i += HT_THREADS) {
      const int t = i / (C >> 1), c = (i - t * (C >> 1)) * 2;
      const __half2 a = __floats2half2_rn(w[(0 * C + c) * 9 + t], w[(0 * C + c + 1) * 9 + t]);
      const __half2 b = __floats2half2_rn(w[(1 * C + c) * 9 + t], w[(1 * C + c + 1) * 9 + t]);
      sw[i] = make_float2(__uint_as_float(*reinterpret_cast<const uint32_t*>(&a)), __uint_as_float(*reinterpret_cast<const uint32_t*>(&b)));
    }
  } else {
    for (int i = tid; i < 9 * C; i += HT_THREADS) {
      const int t = i / C, c = i - t * C;
      sw[i] = make_float2(w[(0 * C + c) * 9 + t], w[(1 * C + c) * 9 + t]);
    }
  }
  // tiles that hold none of the rows 1, H-2 / columns 1, W-2 have no reflected terms: G[t][p] is just dZ shifted by the
  // tap, read straight from the halo tile (3/4 of the tiles at 192x640) -- the G table is only built for border tiles
  const bool interior = !(ty0 <= 1) && !(ty0 <= H - 2 && H - 2 < ty0 + HT_ROWS) &&
                        !(tx0 <= 1) && !(tx0 <= W - 2 && W - 2 < tx0 + HT_COLS);
  // dZ halo tile (1 pixel each side, zero outside the image) staged once; every G term lies inside it.  Interior tiles of
  // the packed loop keep it in the loop's operand form (the conversion happens once here, not per tap and channel block)
  for (int i = tid; i < (HT_ROWS + 2) * (HT_COLS + 2); i += HT_THREADS) {
    const int hy = i / (HT_COLS + 2), hx = i - hy * (HT_COLS + 2);
    const int qy = ty0 + hy - 1, qx = tx0 + hx - 1;
    float2 d = make_float2(0.f, 0.f);
    if (qy >= 0 && qy < H && qx >= 0 && qx < W)
      d = h_up2(__ldg(reinterpret_cast<const uint32_t*>(dz8 + (img * H + qy) * W + qx)), grad_fmt);
    dzs[i] = (HM && interior) ? pack_g(d) : d;
  }
  __syncthreads();
  for (int i = interior ? 9 * HT_ROWS * HT_COLS : tid; i < 9 * HT_ROWS * HT_COLS; i += HT_THREADS) {
    const int t = i / (HT_ROWS * HT_COLS), pix = i - t * (HT_ROWS * HT_COLS);
    const int ry = pix / HT_COLS, rx = pix - ry * HT_COLS;
    const int py = ty0 + ry, px = tx0 + rx;
    const int tyy = t / 3, txx = t - tyy * 3;
    float2 g = make_float2(0.f, 0.f);
    if (py < H && px < W) {
      // virtual rows / columns (in [-1, H] x [-1, W]) that reflect onto (py, px)
      int vy[3], vx[3], ny = 1, nx = 1;
      vy[0] = py; vx[0] = px;
      if (py == 1) vy[ny++] = -1;
      if (py == H - 2) vy[ny++] = H;
      if (px == 1) vx[nx++] = -1;
      if (px == W - 2) vx[nx++] = W;
      for (int iy = 0; iy < ny; ++iy) {
        const int qy = vy[iy] - tyy + 1;
        if (qy < 0 || qy >= H) continue;
        for (int ix = 0; ix < nx; ++ix) {
          const int qx = vx[ix] - txx + 1;
          if (qx < 0 || qx >= W) continue;
          const float2 d = dzs[(qy - ty0 + 1) * (HT_COLS + 2) + (qx - tx0 + 1)];
          g.x += d.x; g.y += d.y;
        }
      }
    }
    G[i] = pack_g(g);
  }
  __syncthreads();
  const int lx = tid & 31, ly = (tid >> 5) * 4;
  const int px = tx0 + lx;
  for (int cb = 0; cb < C; cb += 16) {
    float acc[4][16];
    uint4 pre_y[4][2], pre_d[4][2];
    if (EP) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int py = ty0 + ly + r;
        const bool ok = py < H && px < W;
        const long long o = ok ? ((img * H + py) * W + px) * c8 + (cb >> 3) : 0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          pre_y[r][h] = (ok && yact != nullptr) ? __ldg(&yact[o + h]) : make_uint4(0u, 0u, 0u, 0u);
          pre_d[r][h] = (ok && ACC) ? dst[o + h] : make_uint4(0u, 0u, 0u, 0u);
        }
      }
    }
    if (HM) {
      __half2 hacc[4][8];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) hacc[r][i] = __float2half2_rn(0.f);
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float2 g[4];
        if (interior) {
#pragma unroll
          for (int r = 0; r < 4; ++r) g[r] = dzs[(ly + r + 2 - t / 3) * (HT_COLS + 2) + lx + 2 - t % 3];
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) g[r] = G[t * (HT_ROWS * HT_COLS) + (ly + r) * HT_COLS + lx];
        }
        const float2* wt = sw + t * (C >> 1) + (cb >> 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 ww = wt[i];
          const uint32_t w0b = __float_as_uint(ww.x), w1b = __float_as_uint(ww.y);
          const __half2 w0 = *reinterpret_cast<const __half2*>(&w0b), w1 = *reinterpret_cast<const __half2*>(&w1b);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const uint32_t gxb = __float_as_uint(g[r].x), gyb = __float_as_uint(g[r].y);
            hacc[r][i] = __hfma2(w0, *reinterpret_cast<const __half2*>(&gxb),
                                 __hfma2(w1, *reinterpret_cast<const __half2*>(&gyb), hacc[r][i]));
          }
        }
      }
      if (EP && act_fmt == FP_F16) {
        const __half2 one2 = __float2half2_rn(1.f), zero2 = __float2half2_rn(0.f);
        const __half2 hi2 = __float2half2_rn(65504.f), lo2 = __float2half2_rn(-65504.f);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int py = ty0 + ly + r;
          if (py >= H || px >= W) continue;
          const long long o = ((img * H + py) * W + px) * c8 + (cb >> 3);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t yv[4] = {pre_y[r][h].x, pre_y[r][h].y, pre_y[r][h].z, pre_y[r][h].w};
            const uint32_t dv[4] = {pre_d[r][h].x, pre_d[r][h].y, pre_d[r][h].z, pre_d[r][h].w};
            uint32_t ov[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              __half2 v = __hmin2(__hmax2(hacc[r][h * 4 + q], lo2), hi2);      // an overflowed sum (inf) clips first: inf * 0 = NaN
              // dst holds the other consumer's gradient w.r.t. the activation: the derivative applies to the sum
              if (ACC) v = __hmin2(__hmax2(__hadd2(v, *reinterpret_cast<const __half2*>(&dv[q])), lo2), hi2);
              if (yact != nullptr) {
                const __half2 y2 = *reinterpret_cast<const __half2*>(&yv[q]);
                if (act == FP_ACT_ELU) v = __hmul2(v, __hmin2(__hadd2(y2, one2), one2));   // ELU'(y) = y > 0 ? 1 : y + 1
                else if (act == FP_ACT_RELU) v = __hmul2(v, __hgt2(y2, zero2));
              }
              ov[q] = *reinterpret_cast<const uint32_t*>(&v);
            }
            dst[o + h] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
          }
        }
        continue;
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 v = __half22float2(hacc[r][i]);          // an overflowed sum (inf) clips like every other conversion
          acc[r][2 * i] = fminf(fmaxf(v.x, -65504.f), 65504.f);
          acc[r][2 * i + 1] = fminf(fmaxf(v.y, -65504.f), 65504.f);
        }
    } else {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[r][i] = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      float2 g[4];
      if (interior) {
#pragma unroll
        for (int r = 0; r < 4; ++r) g[r] = dzs[(ly + r + 2 - t / 3) * (HT_COLS + 2) + lx + 2 - t % 3];
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) g[r] = G[t * (HT_ROWS * HT_COLS) + (ly + r) * HT_COLS + lx];
      }
      const float2* wt = sw + t * C + cb;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float2 ww = wt[i];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][i] = fmaf(ww.x, g[r].x, fmaf(ww.y, g[r].y, acc[r][i]));
      }
    }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int py = ty0 + ly + r;
      if (py >= H || px >= W) continue;
      const long long o = ((img * H + py) * W + px) * c8 + (cb >> 3);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = acc[r][h * 8 + i];
        if (ACC) {
          float f[8];
          h_unpack8(EP ? pre_d[r][h] : dst[o + h], grad_fmt, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] += f[i];
        }
        if (yact != nullptr) {
          float fy[8];
          h_unpack8(EP ? pre_y[r][h] : __ldg(&yact[o + h]), act_fmt, fy);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (act == FP_ACT_ELU) a[i] *= fy[i] > 0.f ? 1.f : fy[i] + 1.f;
            else if (act == FP_ACT_RELU) a[i] = fy[i] > 0.f ? a[i] : 0.f;
          }
        }
        uint4 v;
        v.x = h_pk2(a[0], a[1], grad_fmt); v.y = h_pk2(a[2], a[3], grad_fmt);
        v.z = h_pk2(a[4], a[5], grad_fmt); v.w = h_pk2(a[6], a[7], grad_fmt);
        dst[o + h] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- weight gradient
// Stage 1: block b reduces its pixel range; thread = (pixel sub-range g, tap t, 8-channel chunk cc) accumulates
// 2 x 8 sums, then a fixed-order shared-memory reduction over g.  partial[b][co][c][t] (+ bias sums at the end).
__global__ void __launch_bounds__(HEAD_THREADS)
head_wgrad_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dz8, int batch, int H, int W, int C, int act_fmt,
                  int grad_fmt, long long pix_per_block, float* __restrict__ partial) {
  extern __shared__ float red[];          // [groups][lanes_per_group][16]
  const int c8 = C >> 3;
  const int lanes = 9 * c8;               // (tap, chunk) pairs
  const int groups = HEAD_THREADS / lanes;
  const int g = threadIdx.x / lanes, l = threadIdx.x - g * lanes;
  const long long P = static_cast<long long>(batch) * H * W;
  const long long p0 = blockIdx.x * pix_per_block, p1 = min(P, p0 + pix_per_block);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float b0 = 0.f, b1 = 0.f;
  if (g < groups) {
    const int t = l / c8, cc = l - t * c8;
    const int ty = t / 3, tx = t - ty * 3;
    constexpr int U = 4;                  // independent loads in flight per thread
    for (long long pb = p0 + g; pb < p1; pb += static_cast<long long>(groups) * U) {
      uint32_t dv[U];
      uint4 xv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long p = pb + static_cast<long long>(u) * groups;
        dv[u] = 0u;
        xv[u] = make_uint4(0u, 0u, 0u, 0u);
        if (p < p1) {
          const int px = static_cast<int>(p % W);
          const int py = static_cast<int>((p / W) % H);
          const long long img = p / (static_cast<long long>(W) * H) * H;
          const int yy = reflect1(py + ty - 1, H), xx = reflect1(px + tx - 1, W);
          dv[u] = reinterpret_cast<const uint32_t*>(dz8 + p)[0];
          xv[u] = x[((img + yy) * W + xx) * c8 + cc];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float2 d = h_up2(dv[u], grad_fmt);      // zero for p >= p1
        float f[8];
        h_unpack8(xv[u], act_fmt, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] = fmaf(f[i], d.x, acc[i]);
          acc[8 + i] = fmaf(f[i], d.y, acc[8 + i]);
        }
        if (l == 0) { b0 += d.x; b1 += d.y; }
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) red[(g * lanes + l) * 16 + i] = acc[i];
  }
  __shared__ float redb[HEAD_THREADS / 9][2];
  if (g < groups && l == 0) { redb[g][0] = b0; redb[g][1] = b1; }
  __syncthreads();
  float* out = partial + static_cast<long long>(blockIdx.x) * (2 * C * 9 + 2);
  for (int i = threadIdx.x; i < lanes * 16; i += HEAD_THREADS) {
    const int ll = i >> 4, j = i & 15;
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[(gg * lanes + ll) * 16 + j];
    const int t = ll / c8, cc = ll - t * c8;
    const int co = j >> 3, c = cc * 8 + (j & 7);
    out[(co * C + c) * 9 + t] = s;
  }
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += redb[gg][threadIdx.x];
    out[2 * C * 9 + threadIdx.x] = s;
  }
}

// Tiled stage 1 (default): a persistent block walks 8 x 32-pixel tiles; the input tile with its reflected halo
// ((8+2) x (32+2) pixels x C channels) and the tile's dZ pairs are staged in shared memory ONCE, so the nine taps read
// shared memory instead of re-reading global memory nine times.  thread = (pixel group g, tap t, 8-channel chunk cc)
// keeps its 2 x 8 sums in registers across all its tiles; one fixed-order reduction over g at the end.
// Same partial layout as head_wgrad_kernel.  The 2-channel heads on the tensor-core weight-gradient kernel cost as much
// as a 32-channel layer (the cost is staging X and issuing MMAs, 126 of 128 M rows are padding): this kernel is ~2x faster.
constexpr int HW_TH = 8, HW_TW = 32;
__global__ void __launch_bounds__(HEAD_THREADS)
head_wgrad_tile_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dz8, int batch, int H, int W, int C, int act_fmt,
                       int grad_fmt, int tiles_x, int tiles_y, float* __restrict__ partial) {
  extern __shared__ uint4 hw_smem[];
  const int c8 = C >> 3;
  uint4* sx = hw_smem;                                              // [(TH+2)*(TW+2)][c8]
  float2* sdz = reinterpret_cast<float2*>(sx + (HW_TH + 2) * (HW_TW + 2) * c8);   // [TH*TW]
  float* red = reinterpret_cast<float*>(sdz + HW_TH * HW_TW);         // [groups][lanes][16] (reused at the end)
  const int lanes = 9 * c8;
  const int groups = HEAD_THREADS / lanes;
  const int g = threadIdx.x / lanes, l = threadIdx.x - g * lanes;
  const int t = l / c8, cc = l - t * c8;
  const int ty = t / 3, tx = t - ty * 3;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  float b0 = 0.f, b1 = 0.f;
  const int tiles_per_img = tiles_x * tiles_y;
  const int total = batch * tiles_per_img;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int img = tile / tiles_per_img;
    const int r = tile - img * tiles_per_img;
    const int y0 = (r / tiles_x) * HW_TH, x0 = (r % tiles_x) * HW_TW;
    __syncthreads();                                                 // previous tile fully consumed
    for (int i = threadIdx.x; i < (HW_TH + 2) * (HW_TW + 2) * c8; i += HEAD_THREADS) {
      const int ch = i % c8, pp = i / c8;
      const int py = pp / (HW_TW + 2), px = pp - py * (HW_TW + 2);
      const int yy = reflect1(min(y0 + py - 1, H), H), xx = reflect1(min(x0 + px - 1, W), W);   // (clamped: rows/cols past a
      sx[i] = x[((static_cast<long long>(img) * H + yy) * W + xx) * c8 + ch];                   //  ragged tile pair with dz = 0)
    }
    for (int i = threadIdx.x; i < HW_TH * HW_TW; i += HEAD_THREADS) {
      const int py = i / HW_TW, px = i - py * HW_TW;
      float2 d = make_float2(0.f, 0.f);
      if (y0 + py < H && x0 + px < W)
        d = h_up2(reinterpret_cast<const uint32_t*>(dz8 + (static_cast<long long>(img) * H + y0 + py) * W + x0 + px)[0], grad_fmt);
      sdz[i] = d;
    }
    __syncthreads();
    if (g < groups) {
      for (int p = g; p < HW_TH * HW_TW; p += groups) {
        const int py = p / HW_TW, px = p - py * HW_TW;
        const float2 d = sdz[p];
        float f[8];
        h_unpack8(sx[((py + ty) * (HW_TW + 2) + px + tx) * c8 + cc], act_fmt, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc[i] = fmaf(f[i], d.x, acc[i]);
          acc[8 + i] = fmaf(f[i], d.y, acc[8 + i]);
        }
        if (l == 0) { b0 += d.x; b1 += d.y; }
      }
    }
  }
  __syncthreads();
  if (g < groups) {
#pragma unroll
    for (int i = 0; i < 16; ++i) red[(g * lanes + l) * 16 + i] = acc[i];
  }
  __shared__ float redb2[HEAD_THREADS / 9][2];
  if (g < groups && l == 0) { redb2[g][0] = b0; redb2[g][1] = b1; }
  __syncthreads();
  float* out = partial + static_cast<long long>(blockIdx.x) * (2 * C * 9 + 2);
  for (int i = threadIdx.x; i < lanes * 16; i += HEAD_THREADS) {
    const int ll = i >> 4, j = i & 15;
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += red[(gg * lanes + ll) * 16 + j];
    const int t2 = ll / c8, cc2 = ll - t2 * c8;
    const int co = j >> 3, c = cc2 * 8 + (j & 7);
    out[(co * C + c) * 9 + t2] = s;
  }
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int gg = 0; gg < groups; ++gg) s += redb2[gg][threadIdx.x];
    out[2 * C * 9 + threadIdx.x] = s;
  }
}

// Stage 2: fixed-order sum over blocks
__global__ void head_wgrad_reduce_kernel(const float* __restrict__ partial, int nblk, int n, int accumulate,
                                         float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n + 2) return;
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += partial[static_cast<long long>(b) * (n + 2) + i];
  float* o = i < n ? dw + i : db + (i - n);
  *o = (accumulate ? *o : 0.f) + static_cast<float>(s);
}

}  // namespace fp

using namespace fp;

static int head_check(const char* who, int C, int H, int W) {
  if ((C & 7) || C > HEAD_MAX_C || C < 8) return set_error(FP_ERR_UNSUPPORTED, "%s: C=%d must be a multiple of 8 <= %d", who, C, HEAD_MAX_C);
  if (H < 2 || W < 2) return set_error(FP_ERR_ARG, "%s: reflection padding needs H, W >= 2", who);
  return FP_OK;
}

extern "C" int fp_head_conv_fwd(const void* x, const float* w_oihw, const float* bias, int batch, int h, int w, int c,
                                int sigmoid, int dtype, float* out, void* stream) {
  if (!x || !w_oihw || !bias || !out) return set_error(FP_ERR_ARG, "fp_head_conv_fwd: null argument");
  int rc = head_check("fp_head_conv_fwd", c, h, w);
  if (rc != FP_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (opts().head_tile != 0) {
    const int tiles_x = (w + HT_COLS - 1) / HT_COLS, tiles_y = (h + HT_ROWS - 1) / HT_ROWS;
    { const int arc = ensure_max_smem(reinterpret_cast<const void*>(head_fwd_tile_kernel), HT_SMEM_BYTES, "head_fwd_tile_kernel"); if (arc != FP_OK) return arc; }
    head_fwd_tile_kernel<<<static_cast<unsigned>(static_cast<long long>(batch) * tiles_x * tiles_y), HT_THREADS, HT_SMEM_BYTES, st>>>(
        static_cast<const uint4*>(x), w_oihw, bias, h, w, c, tiles_x, tiles_y, sigmoid, dtype, reinterpret_cast<float2*>(out));
  } else if (c <= 32) {
    const long long n = static_cast<long long>(batch) * h * w;             // one thread per pixel
    head_fwd_pixel_kernel<<<static_cast<unsigned>((n + HEAD_THREADS - 1) / HEAD_THREADS), HEAD_THREADS, 0, st>>>(
        static_cast<const uint4*>(x), w_oihw, bias, batch, h, w, c, sigmoid, dtype, reinterpret_cast<float2*>(out));
  } else {
    const long long n = static_cast<long long>(batch) * h * w * (c / 8);   // one thread per (pixel, 8-channel chunk)
    head_fwd_kernel<<<static_cast<unsigned>((n + HEAD_THREADS - 1) / HEAD_THREADS), HEAD_THREADS, 0, st>>>(
        static_cast<const uint4*>(x), w_oihw, bias, batch, h, w, c, sigmoid, dtype, reinterpret_cast<float2*>(out));
  }
  FP_CHECK_LAUNCH("fp_head_conv_fwd");
  return FP_OK;
}

extern "C" int fp_head_conv_dgrad(const void* dz8, const float* w_oihw, int batch, int h, int w, int c, int accumulate,
                                  const void* y_act, int act, int act_dtype, int grad_dtype, void* dst, void* stream) {
  if (!dz8 || !w_oihw || !dst) return set_error(FP_ERR_ARG, "fp_head_conv_dgrad: null argument");
  int rc = head_check("fp_head_conv_dgrad", c, h, w);
  if (rc != FP_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c % 16 == 0 && opts().head_tile != 0) {
    const int tiles_x = (w + HT_COLS - 1) / HT_COLS, tiles_y = (h + HT_ROWS - 1) / HT_ROWS;
    const int smem = HD_SMEM_G + HD_SMEM_DZ + 9 * c * static_cast<int>(sizeof(float2));
    const bool hm = grad_dtype == FP_F16 && opts().head_hmath != 0;
    const bool ep = opts().head_epi != 0;
    auto pick = [&](auto acc) {
      constexpr bool A = decltype(acc)::value;
      return hm ? (ep ? head_dgrad_tile_kernel<true, 1, A> : head_dgrad_tile_kernel<true, 0, A>)
                : (ep ? head_dgrad_tile_kernel<false, 1, A> : head_dgrad_tile_kernel<false, 0, A>);
    };
    auto kern = accumulate ? pick(std::true_type{}) : pick(std::false_type{});
    { const int arc = ensure_max_smem(reinterpret_cast<const void*>(kern), HD_SMEM_G + HD_SMEM_DZ + 9 * HEAD_MAX_C * static_cast<int>(sizeof(float2)), "head_dgrad_tile_kernel"); if (arc != FP_OK) return arc; }
    kern<<<static_cast<unsigned>(static_cast<long long>(batch) * tiles_x * tiles_y), HT_THREADS, smem, st>>>(
        static_cast<const uint4*>(dz8), w_oihw, h, w, c, tiles_x, tiles_y, static_cast<const uint4*>(y_act), act,
        act_dtype, grad_dtype, static_cast<uint4*>(dst));
  } else if (c == 32) {
    const long long n = static_cast<long long>(batch) * h * w;
    head_dgrad_pixel32_kernel<<<static_cast<unsigned>((n + HEAD_THREADS - 1) / HEAD_THREADS), HEAD_THREADS, 0, st>>>(
        static_cast<const uint4*>(dz8), w_oihw, batch, h, w, accumulate, static_cast<const uint4*>(y_act), act, act_dtype,
        grad_dtype, static_cast<uint4*>(dst));
  } else {
    const long long n = static_cast<long long>(batch) * h * w * (c / 8);
    head_dgrad_kernel<<<static_cast<unsigned>((n + HEAD_THREADS - 1) / HEAD_THREADS), HEAD_THREADS, 0, st>>>(
        static_cast<const uint4*>(dz8), w_oihw, batch, h, w, c, accumulate, static_cast<const uint4*>(y_act), act, act_dtype,
        grad_dtype, static_cast<uint4*>(dst));
  }
  FP_CHECK_LAUNCH("fp_head_conv_dgrad");
  return FP_OK;
}

extern "C" long long fp_head_conv_wgrad_workspace_floats(int c) { return 1184ll * (2 * c * 9 + 2); }

extern "C" int fp_head_conv_wgrad(const void* x, const void* dz8, int batch, int h, int w, int c, int act_dtype,
                                  int grad_dtype, int accumulate, float* dw_oihw, float* dbias, float* workspace,
                                  void* stream) {
  if (!x || !dz8 || !dw_oihw || !dbias || !workspace) return set_error(FP_ERR_ARG, "fp_head_conv_wgrad: null argument");
  int rc = head_check("fp_head_conv_wgrad", c, h, w);
  if (rc != FP_OK) return rc;
  const int lanes = 9 * (c / 8);
  if (lanes > HEAD_THREADS) return set_error(FP_ERR_UNSUPPORTED, "fp_head_conv_wgrad: C too large");
  const int groups = HEAD_THREADS / lanes;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int nblk;
  if (opts().head_tile != 0) {
    const int tiles_x = (w + HW_TW - 1) / HW_TW, tiles_y = (h + HW_TH - 1) / HW_TH;
    const long long tiles = static_cast<long long>(batch) * tiles_x * tiles_y;
    const int smem = (HW_TH + 2) * (HW_TW + 2) * (c / 8) * 16 + HW_TH * HW_TW * 8 + groups * lanes * 16 * 4;
    const int per_sm = smem > 100 * 1024 ? 2 : (smem > 56 * 1024 ? 3 : 4);
    nblk = static_cast<int>(tiles < per_sm * num_sms() ? tiles : per_sm * num_sms());
    if (nblk > 1184) nblk = 1184;
    { const int arc = ensure_max_smem(reinterpret_cast<const void*>(head_wgrad_tile_kernel), 200 * 1024, "head_wgrad_tile_kernel"); if (arc != FP_OK) return arc; }
    head_wgrad_tile_kernel<<<nblk, HEAD_THREADS, smem, st>>>(static_cast<const uint4*>(x), static_cast<const uint4*>(dz8), batch, h, w,
                                                            c, act_dtype, grad_dtype, tiles_x, tiles_y, workspace);
  } else {
    const long long P = static_cast<long long>(batch) * h * w;
    long long want = (P + 4095) / 4096;
    nblk = static_cast<int>(want < 1 ? 1 : (want > 1184 ? 1184 : want));
    const long long ppb = (P + nblk - 1) / nblk;
    head_wgrad_kernel<<<nblk, HEAD_THREADS, groups * lanes * 16 * sizeof(float), st>>>(
        static_cast<const uint4*>(x), static_cast<const uint4*>(dz8), batch, h, w, c, act_dtype, grad_dtype, ppb, workspace);
  }
  FP_CHECK_LAUNCH("fp_head_conv_wgrad");
  const int n = 2 * c * 9;
  head_wgrad_reduce_kernel<<<(n + 2 + 127) / 128, 128, 0, st>>>(workspace, nblk, n, accumulate, dw_oihw, dbias);
  FP_CHECK_LAUNCH("fp_head_conv_wgrad(reduce)");
  return FP_OK;
}
