// Bandwidth-bound kernels of the training step: BatchNorm batch statistics / apply / backward, the fused
// multi-scale loss + gradient, the heads' bilinear-transpose backward, the fold of a convolution's
// padded-input gradient back onto its sources (adjoint of reflection pad / nearest upsample / concat),
// bias gradients and Adam.  All reductions are two-stage and deterministic (no floating-point atomics),
// mirroring the reference's cudnn.deterministic=True (footprints/training/train.py:36).
#include "fp_internal.h"
#include "fp_ptx.cuh"

#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fp {

__device__ __forceinline__ float2 up2(uint32_t u, int fmt) {
  if (fmt == FP_BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}
__device__ __forceinline__ uint32_t pk2(float a, float b, int fmt) {
  if (fmt == FP_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return fp::f32x2_to_f16x2_sat(a, b);
}
__device__ __forceinline__ void unpack8(const uint4& v, int fmt, float (&f)[8]) {
  float2 t;
  t = up2(v.x, fmt); f[0] = t.x; f[1] = t.y;
  t = up2(v.y, fmt); f[2] = t.x; f[3] = t.y;
  t = up2(v.z, fmt); f[4] = t.x; f[5] = t.y;
  t = up2(v.w, fmt); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8], int fmt) {
  uint4 o;
  o.x = pk2(f[0], f[1], fmt); o.y = pk2(f[2], f[3], fmt); o.z = pk2(f[4], f[5], fmt); o.w = pk2(f[6], f[7], fmt);
  return o;
}

constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = 592;   // 4 x 148 SMs

// ------------------------------------------------------------------------------------------------
// Column reductions over a [M][C] 16-bit matrix.  MODE 0: (sum z, sum z^2)  (BatchNorm statistics)
//                                                 MODE 1: (sum g, sum g*xhat) with g = dy * (y > 0)  (BatchNorm backward)
//                                                 MODE 2: (sum dz, -)  (bias gradient)
// partial layout: [block][2][C] fp32.  Each thread owns one 8-channel chunk and strides over rows.
// ------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(RED_THREADS, MODE == 1 ? 3 : 4)   // mode 1 holds two rows of three tensors in flight
col_reduce_kernel(const uint4* __restrict__ a, int a_fmt, const uint4* __restrict__ y, const uint4* __restrict__ z,
                  int act_fmt, const float* __restrict__ mean, const float* __restrict__ invstd, long long M, int C,
                  float* __restrict__ partial) {
  extern __shared__ float red[];   // [rows_per_pass][2][C]
  const int c8 = C >> 3;
  const int tx = threadIdx.x % c8, ty = threadIdx.x / c8;
  const int rows_per_pass = RED_THREADS / c8;
  float s0[8], s1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s0[i] = s1[i] = 0.f;
  float mu[8], is[8];
  if (MODE == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { mu[i] = mean[tx * 8 + i]; is[i] = invstd[tx * 8 + i]; }
  }
  if (ty < rows_per_pass) {
    // two rows per iteration: twice the loads in flight per thread (the reductions are latency-, not bandwidth-limited
    // at 4 blocks x 256 threads per SM otherwise)
    const long long step = static_cast<long long>(gridDim.x) * rows_per_pass;
    auto accumulate_row = [&](const uint4& va, const uint4& vy, const uint4& vz) {
      float f[8];
      unpack8(va, a_fmt, f);
      if (MODE == 0) {
        if (y != nullptr) {            // split precision: z = hi + lo (the low halves ride in the y slot)
          float fl[8];
          unpack8(vy, a_fmt, fl);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] += fl[i];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) { s0[i] += f[i]; s1[i] += f[i] * f[i]; }
      } else if (MODE == 1) {
        float fy[8], fz[8];
        if (y != nullptr) unpack8(vy, act_fmt, fy);   // y == nullptr: no ReLU follows this BatchNorm
        unpack8(vz, act_fmt, fz);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float g = (y == nullptr || fy[i] > 0.f) ? f[i] : 0.f;
          s0[i] += g;
          s1[i] += g * (fz[i] - mu[i]) * is[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) s0[i] += f[i];
      }
    };
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    long long r = static_cast<long long>(blockIdx.x) * rows_per_pass + ty;
    for (; r + step < M; r += 2 * step) {
      const long long i0 = r * c8 + tx, i1 = (r + step) * c8 + tx;
      const uint4 a0 = a[i0], a1 = a[i1];
      uint4 y0 = zero4, y1 = zero4, z0 = zero4, z1 = zero4;
      if (MODE == 1) {
        if (y != nullptr) { y0 = y[i0]; y1 = y[i1]; }
        z0 = z[i0]; z1 = z[i1];
      } else if (MODE == 0 && y != nullptr) {
        y0 = y[i0]; y1 = y[i1];
      }
      accumulate_row(a0, y0, z0);
      accumulate_row(a1, y1, z1);
    }
    if (r < M) {
      const long long i0 = r * c8 + tx;
      uint4 y0 = zero4, z0 = zero4;
      if (MODE == 1) {
        if (y != nullptr) y0 = y[i0];
        z0 = z[i0];
      } else if (MODE == 0 && y != nullptr) {
        y0 = y[i0];
      }
      accumulate_row(a[i0], y0, z0);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      red[(ty * 2 + 0) * C + tx * 8 + i] = s0[i];
      red[(ty * 2 + 1) * C + tx * 8 + i] = s1[i];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * C; j += RED_THREADS) {
    float acc = 0.f;
    for (int r = 0; r < rows_per_pass; ++r) acc += red[r * 2 * C + j];
    partial[static_cast<long long>(blockIdx.x) * 2 * C + j] = acc;
  }
}

// Second stage of the column reductions: one 128-thread block per channel sums the per-block partials (threads stride
// over the blocks, fp64, shuffle tree per warp, then the four warp sums in order) -- deterministic.  These kernels are
// pure latency (<= 1184 partials per channel): 128 threads per channel keep the dependent-load chains <= 10 long.
constexpr int SUM_THREADS = 128;
__device__ __forceinline__ void block_sum_partials(const float* __restrict__ partial, int nblk, int C, int c, double& s,
                                                   double& q) {
  __shared__ double sh[2][SUM_THREADS / 32];
  s = 0.0; q = 0.0;
  for (int b = threadIdx.x; b < nblk; b += SUM_THREADS) {
    s += partial[(static_cast<long long>(b) * 2 + 0) * C + c];
    q += partial[(static_cast<long long>(b) * 2 + 1) * C + c];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, off);
    q += __shfl_xor_sync(0xffffffffu, q, off);
  }
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
  __syncthreads();
  s = (sh[0][0] + sh[0][1]) + (sh[0][2] + sh[0][3]);
  q = (sh[1][0] + sh[1][1]) + (sh[1][2] + sh[1][3]);
}

// BatchNorm2d training statistics (torch.nn.functional.batch_norm, training=True): biased variance for
// normalisation, unbiased for the running estimate, momentum 0.1 (torchvision resnet.py BatchNorm2d defaults).
__global__ void bn_finalize_kernel(const float* __restrict__ partial, int nblk, long long M, int C,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int c = blockIdx.x;                                        // one block per channel
  double s, q;
  block_sum_partials(partial, nblk, C, c, s, q);
  if (threadIdx.x != 0) return;
  const double n = static_cast<double>(M);
  const double mean = s / n;
  double var = q / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - static_cast<float>(mean) * sc;
  save_mean[c] = static_cast<float>(mean);
  save_invstd[c] = invstd;
  if (running_mean != nullptr) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

// sums the [nblk][2][C] partials: out0[c] = sum row 0, out1[c] = sum row 1 (fp64 accumulate)
__global__ void partial_sum_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out0,
                                   float* __restrict__ out1) {
  const int c = blockIdx.x;                                        // one block per channel
  double s, q;
  block_sum_partials(partial, nblk, C, c, s, q);
  if (threadIdx.x != 0) return;
  if (out0 != nullptr) out0[c] = static_cast<float>(s);
  if (out1 != nullptr) out1[c] = static_cast<float>(q);
}

// y = act(z*scale + shift (+ residual))
__global__ void bn_apply_kernel(const uint4* __restrict__ z, const float* __restrict__ scale,
                                const float* __restrict__ shift, const uint4* __restrict__ res, int relu, long long n8,
                                int c8, int fmt, uint4* __restrict__ y) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  const int cb = static_cast<int>(idx % c8) * 8;
  float f[8];
  unpack8(z[idx], fmt, f);
  float r[8];
  if (res != nullptr) unpack8(res[idx], fmt, r);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = f[i] * scale[cb + i] + shift[cb + i];
    if (res != nullptr) v += r[i];
    f[i] = relu ? fmaxf(v, 0.f) : v;
  }
  y[idx] = pack8(f, fmt);
}

// split-precision flavour: z, residual and y are hi + lo pairs of 16-bit tensors
__global__ void bn_apply_split_kernel(const uint4* __restrict__ zh, const uint4* __restrict__ zl,
                                      const float* __restrict__ scale, const float* __restrict__ shift,
                                      const uint4* __restrict__ rh, const uint4* __restrict__ rl, int relu, long long n8,
                                      int c8, int fmt, uint4* __restrict__ yh, uint4* __restrict__ yl) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  const int cb = static_cast<int>(idx % c8) * 8;
  float f[8], l[8], r[8], q[8];
  unpack8(zh[idx], fmt, f);
  unpack8(zl[idx], fmt, l);
  if (rh != nullptr) { unpack8(rh[idx], fmt, r); unpack8(rl[idx], fmt, q); }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float v = (f[i] + l[i]) * scale[cb + i] + shift[cb + i];
    if (rh != nullptr) v += r[i] + q[i];
    f[i] = relu ? fmaxf(v, 0.f) : v;
  }
  const uint4 hi = pack8(f, fmt);
  float h[8];
  unpack8(hi, fmt, h);
#pragma unroll
  for (int i = 0; i < 8; ++i) l[i] = f[i] - h[i];
  yh[idx] = hi;
  yl[idx] = pack8(l, fmt);
}

// BatchNorm backward (training statistics), with the ReLU that follows it folded in:
//   g = dy * (y > 0);  dz = gamma*invstd * (g - mean(g) - xhat * mean(g*xhat));  g is also the gradient that
//   flows into the residual branch.  dgamma = sum(g*xhat), dbeta = sum(g) come from col_reduce<1>.
__global__ void bn_bwd_apply_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y,
                                    const uint4* __restrict__ z, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ dbeta, const float* __restrict__ dgamma, float inv_n,
                                    int relu, long long n8, int c8, int act_fmt, int grad_fmt, uint4* __restrict__ dz,
                                    uint4* __restrict__ g_out) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  const int cb = static_cast<int>(idx % c8) * 8;
  float g[8], fy[8], fz[8], o[8];
  unpack8(dy[idx], grad_fmt, g);
  unpack8(z[idx], act_fmt, fz);
  if (relu) {
    unpack8(y[idx], act_fmt, fy);
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = fy[i] > 0.f ? g[i] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = cb + i;
    const float xhat = (fz[i] - mean[c]) * invstd[c];
    o[i] = gamma[c] * invstd[c] * (g[i] - dbeta[c] * inv_n - xhat * dgamma[c] * inv_n);
  }
  dz[idx] = pack8(o, grad_fmt);
  if (g_out != nullptr) g_out[idx] = pack8(g, grad_fmt);
}

// ------------------------------------------------------------------------------------------------
// Fused multi-scale loss + gradient (LossManager.__call__, footprints/training/losses.py:31-92).
// One thread per pixel reads the 6 target planes once and all 4 scales x 4 channels of predictions,
// accumulates the 16 loss sums and writes d(loss)/d(prediction) for every scale.
// ------------------------------------------------------------------------------------------------
struct LossParams {
  const float* pred[4];
  float* grad[4];
  const float* depth; const float* visible; const float* all_ground; const float* ground_depth;
  const float* moving; const float* depth_mask;
  int B; long long HW;
  float min_disp, disp_range, prior_w, gscale;   // gscale = 1 / (4 * P)
};

__device__ __forceinline__ float bce_logits(float z, float t) {
  return fmaxf(z, 0.f) - z * t + log1pf(expf(-fabsf(z)));
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }

constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS) loss_kernel(const LossParams p, float* __restrict__ partial) {
  __shared__ float red[LOSS_THREADS / 32][16];
  const long long P = static_cast<long long>(p.B) * p.HW;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (long long idx = blockIdx.x * static_cast<long long>(LOSS_THREADS) + threadIdx.x; idx < P;
       idx += static_cast<long long>(gridDim.x) * LOSS_THREADS) {
    const long long b = idx / p.HW, pix = idx - b * p.HW;
    const float t_d = p.depth[idx], t_vis = p.visible[idx], t_all = p.all_ground[idx], t_gd = p.ground_depth[idx];
    const float mv = 1.f - p.moving[idx];                              // losses.py:44
    const float m = (t_all + p.depth_mask[idx]) > 0.f ? 1.f : 0.f;    // losses.py:140
    const float vd = t_d > 0.f ? 1.f : 0.f, vgd = t_gd > 0.f ? 1.f : 0.f;
    const long long o = b * 4 * p.HW + pix;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float z0 = p.pred[s][o], z1 = p.pred[s][o + p.HW], s2 = p.pred[s][o + 2 * p.HW], s3 = p.pred[s][o + 3 * p.HW];
      // visible ground: BCE with logits (losses.py:53-55, 121)
      acc[s * 4 + 0] += bce_logits(z0, t_vis);
      // hidden ground: 3-class loss with prior (losses.py:129-152)
      acc[s * 4 + 1] += bce_logits(z1, t_all) * m * mv + p.prior_w * bce_logits(z1, 0.f) * (1.f - m);
      // depth / ground depth: log(|d - t| + 1) on valid pixels, d = 1/(min_disp + range*sigmoid) (utils.py:36-42)
      const float d2 = 1.f / (p.min_disp + p.disp_range * s2), d3 = 1.f / (p.min_disp + p.disp_range * s3);
      const float e2 = d2 - t_d, e3 = d3 - t_gd;
      acc[s * 4 + 2] += logf(fabsf(e2) + 1.f) * vd;
      acc[s * 4 + 3] += logf(fabsf(e3) + 1.f) * vgd;
      if (p.grad[s] != nullptr) {
        const float sg1 = sigmoidf_(z1);
        float* g = p.grad[s] + o;
        g[0] = (sigmoidf_(z0) - t_vis) * p.gscale;
        g[p.HW] = ((sg1 - t_all) * m * mv + p.prior_w * sg1 * (1.f - m)) * p.gscale;
        const float sgn2 = e2 > 0.f ? 1.f : (e2 < 0.f ? -1.f : 0.f), sgn3 = e3 > 0.f ? 1.f : (e3 < 0.f ? -1.f : 0.f);
        g[2 * p.HW] = sgn2 / (fabsf(e2) + 1.f) * vd * (-p.disp_range * d2 * d2) * p.gscale;
        g[3 * p.HW] = sgn3 / (fabsf(e3) + 1.f) * vgd * (-p.disp_range * d3 * d3) * p.gscale;
      }
    }
  }
  // deterministic block reduction: warp shuffle tree, then fixed-order sum over warps
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float v = acc[i];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    float v = 0.f;
    for (int w = 0; w < LOSS_THREADS / 32; ++w) v += red[w][threadIdx.x];
    partial[blockIdx.x * 16 + threadIdx.x] = v;
  }
}

// out[21]: per scale s (order '1/8','1/4','1/2','1/1'): [s*5+0..3] = visible_ground, all_ground, depth,
// ground_depth means; [s*5+4] = their sum; [20] = mean over scales (losses.py:80-88)
__global__ void loss_finalize_kernel(const float* __restrict__ partial, int nblk, double inv_p, float* __restrict__ out) {
  __shared__ double sums[16];
  if (threadIdx.x < 16) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += partial[b * 16 + threadIdx.x];
    sums[threadIdx.x] = s * inv_p;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int s = 0; s < 4; ++s) {
      const float vis = static_cast<float>(sums[s * 4 + 0]), all = static_cast<float>(sums[s * 4 + 1]);
      const float dep = static_cast<float>(sums[s * 4 + 2]), gd = static_cast<float>(sums[s * 4 + 3]);
      out[s * 5 + 0] = vis; out[s * 5 + 1] = all; out[s * 5 + 2] = dep; out[s * 5 + 3] = gd;
      const float l = dep + vis + all + gd;                             // losses.py:80-83 order
      out[s * 5 + 4] = l;
      total += l;
    }
    out[20] = total / 4.f;                                             // losses.py:87
  }
}

// ------------------------------------------------------------------------------------------------
// Heads backward: adjoint of the bilinear upsample (align_corners=False, integer scale) of one decoder's
// two channels, times sigmoid' for the depth decoder.  Output: dZ of the head conv, [B][h][w][8] 16-bit
// (channels 2..7 zero so that the tensor-core dgrad/wgrad kernels see 16-byte pixel rows).
// ------------------------------------------------------------------------------------------------
__global__ void heads_bwd_kernel(const float* __restrict__ dout /*[B][4][H][W]*/, int ch0,
                                 const float2* __restrict__ head_post /*[B][h][w][2] or null*/, int batch, int h, int w,
                                 int scale, int grad_fmt, float gscale, uint4* __restrict__ dz) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * h * w) return;
  const int j = static_cast<int>(idx % w);
  const int i = static_cast<int>((idx / w) % h);
  const int b = static_cast<int>(idx / (static_cast<long long>(w) * h));
  const int H = h * scale, W = w * scale;
  const float* p0 = dout + (static_cast<long long>(b) * 4 + ch0) * H * W;
  const float* p1 = p0 + static_cast<long long>(H) * W;
  float a0 = 0.f, a1 = 0.f;
  if (scale == 1) {
    a0 = p0[static_cast<long long>(i) * W + j];
    a1 = p1[static_cast<long long>(i) * W + j];
  } else {
    const float rs = 1.f / static_cast<float>(scale);
    const int Y0 = max(0, scale * i - scale / 2 - 1), Y1 = min(H - 1, scale * i + (3 * scale) / 2);
    const int X0 = max(0, scale * j - scale / 2 - 1), X1 = min(W - 1, scale * j + (3 * scale) / 2);
    for (int Y = Y0; Y <= Y1; ++Y) {
      const float sy = fmaxf(rs * (Y + 0.5f) - 0.5f, 0.f);
      const int y0 = static_cast<int>(sy), y1 = y0 + (y0 < h - 1 ? 1 : 0);
      const float ly = sy - y0;
      const float wy = (y0 == i ? 1.f - ly : 0.f) + (y1 == i ? ly : 0.f);
      if (wy == 0.f) continue;
      for (int X = X0; X <= X1; ++X) {
        const float sx = fmaxf(rs * (X + 0.5f) - 0.5f, 0.f);
        const int x0 = static_cast<int>(sx), x1 = x0 + (x0 < w - 1 ? 1 : 0);
        const float lx = sx - x0;
        const float wx = (x0 == j ? 1.f - lx : 0.f) + (x1 == j ? lx : 0.f);
        const float wgt = wy * wx;
        a0 += wgt * p0[static_cast<long long>(Y) * W + X];
        a1 += wgt * p1[static_cast<long long>(Y) * W + X];
      }
    }
  }
  if (head_post != nullptr) {   // sigmoid'(z) = s (1 - s) from the saved post-sigmoid head map
    const float2 s = head_post[idx];
    a0 *= s.x * (1.f - s.x);
    a1 *= s.y * (1.f - s.y);
  }
  uint4 o;
  o.x = pk2(a0 * gscale, a1 * gscale, grad_fmt);      // loss scale: keeps fp16 inter-layer gradients in range
  o.y = o.z = o.w = 0u;
  dz[idx] = o;
}

// ------------------------------------------------------------------------------------------------
// Fold: adjoint of [nearest x2 upsample] -> [concat] -> [ReflectionPad2d(1)] applied to the gradient of the
// padded virtual input dVp [B][Hv+2][Wv+2][Ccat].  Destination is either the pre-upsample source (UP) or a
// source stored at the virtual resolution.  Optionally accumulates into the destination and applies the
// producing layer's activation derivative from its saved output (ELU: y>0 ? 1 : y+1; ReLU: y>0).
// ------------------------------------------------------------------------------------------------
template <bool UP>
__global__ void fold_kernel(const uint4* __restrict__ dvp, int Hv, int Wv, int ccat8, int coff8, int batch, int Hd,
                            int Wd, int c8, int accumulate, const uint4* __restrict__ yact, int act, int act_fmt,
                            int grad_fmt, uint4* __restrict__ dst) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * Hd * Wd * c8) return;
  const int cc = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int x = static_cast<int>(t % Wd); t /= Wd;
  const int y = static_cast<int>(t % Hd);
  const int b = static_cast<int>(t / Hd);
  const int Wp = Wv + 2, Hp = Hv + 2;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  constexpr int R = UP ? 2 : 1;
  // interior fast path (all but the two rows / columns next to each border): every virtual pixel has exactly one padded
  // source, no reflected terms -- straight-line loads, all of them in flight at once
  const int vy0 = UP ? 2 * y : y, vx0 = UP ? 2 * x : x;
  if (vy0 > 1 && vy0 + R - 1 < Hv - 2 && vx0 > 1 && vx0 + R - 1 < Wv - 2) {
    uint4 v[R * R];
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
      for (int bb = 0; bb < R; ++bb)
        v[a * R + bb] = dvp[((static_cast<long long>(b) * Hp + vy0 + a + 1) * Wp + vx0 + bb + 1) * ccat8 + coff8 + cc];
    uint4 prev = make_uint4(0u, 0u, 0u, 0u), ya = make_uint4(0u, 0u, 0u, 0u);
    if (accumulate) prev = dst[idx];
    if (yact != nullptr) ya = yact[idx];
#pragma unroll
    for (int k = 0; k < R * R; ++k) {
      float f[8];
      unpack8(v[k], grad_fmt, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
    if (accumulate) {
      float f[8];
      unpack8(prev, grad_fmt, f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
    if (yact != nullptr) {
      float fy[8];
      unpack8(ya, act_fmt, fy);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (act == FP_ACT_ELU) acc[i] *= fy[i] > 0.f ? 1.f : fy[i] + 1.f;
        else if (act == FP_ACT_RELU) acc[i] = fy[i] > 0.f ? acc[i] : 0.f;
      }
    }
    dst[idx] = pack8(acc, grad_fmt);
    return;
  }
#pragma unroll
  for (int a = 0; a < R; ++a) {
    const int vy = UP ? 2 * y + a : y;
    // padded rows that reflect onto virtual row vy: vy+1, plus 0 if vy==1, plus Hv+1 if vy==Hv-2
    int ys[3]; int ny = 0;
    ys[ny++] = vy + 1;
    if (vy == 1) ys[ny++] = 0;
    if (vy == Hv - 2) ys[ny++] = Hv + 1;
#pragma unroll
    for (int bb = 0; bb < R; ++bb) {
      const int vx = UP ? 2 * x + bb : x;
      int xs[3]; int nx = 0;
      xs[nx++] = vx + 1;
      if (vx == 1) xs[nx++] = 0;
      if (vx == Wv - 2) xs[nx++] = Wv + 1;
      for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
          float f[8];
          unpack8(dvp[((static_cast<long long>(b) * Hp + ys[iy]) * Wp + xs[ix]) * ccat8 + coff8 + cc], grad_fmt, f);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] += f[i];
        }
    }
  }
  if (accumulate) {
    float f[8];
    unpack8(dst[idx], grad_fmt, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += f[i];
  }
  if (yact != nullptr) {
    float fy[8];
    unpack8(yact[idx], act_fmt, fy);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (act == FP_ACT_ELU) acc[i] *= fy[i] > 0.f ? 1.f : fy[i] + 1.f;
      else if (act == FP_ACT_RELU) acc[i] = fy[i] > 0.f ? acc[i] : 0.f;
    }
  }
  dst[idx] = pack8(acc, grad_fmt);
}

// dz = dy * act'(y) (elementwise; used where no fold precedes the activation backward)
__global__ void act_bwd_kernel(const uint4* __restrict__ dy, const uint4* __restrict__ y, int act, long long n8,
                               int act_fmt, int grad_fmt, uint4* __restrict__ dz) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  float g[8], fy[8];
  unpack8(dy[idx], grad_fmt, g);
  unpack8(y[idx], act_fmt, fy);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (act == FP_ACT_ELU) g[i] *= fy[i] > 0.f ? 1.f : fy[i] + 1.f;
    else if (act == FP_ACT_RELU) g[i] = fy[i] > 0.f ? g[i] : 0.f;
  }
  dz[idx] = pack8(g, grad_fmt);
}

// a += b  (16-bit gradient buffers)
__global__ void grad_add_kernel(uint4* __restrict__ a, const uint4* __restrict__ b, long long n8, int grad_fmt) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n8) return;
  float fa[8], fb[8];
  unpack8(a[idx], grad_fmt, fa);
  unpack8(b[idx], grad_fmt, fb);
#pragma unroll
  for (int i = 0; i < 8; ++i) fa[i] += fb[i];
  a[idx] = pack8(fa, grad_fmt);
}

// max-pool 3x3/2 pad 1 backward (h, w even): one thread per 2x2 block of input pixels and 8-channel chunk.  The block
// is covered by the four windows (oy, ox) in {j, j+1} x {i, i+1}; each window's argmax (first maximum in scan order
// wins ties, like ATen) is found once and its gradient routed to whichever of the four pixels holds it -- a gather,
// so the result is deterministic and every dx element is written exactly once.
__global__ void __launch_bounds__(256)
maxpool_bwd_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy, int batch, int h, int w, int c8,
                   int act_fmt, int grad_fmt, int accumulate, uint4* __restrict__ dx) {
  const int ho = h / 2, wo = w / 2;
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(batch) * ho * wo * c8) return;
  const int cc = static_cast<int>(idx % c8);
  long long t = idx / c8;
  const int i = static_cast<int>(t % wo); t /= wo;
  const int j = static_cast<int>(t % ho);
  const long long b = t / ho;
  float acc[4][8];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[q][e] = 0.f;
#pragma unroll
  for (int wy = 0; wy < 2; ++wy) {
    const int oy = j + wy;
    if (oy >= ho) continue;
#pragma unroll
    for (int wx = 0; wx < 2; ++wx) {
      const int ox = i + wx;
      if (ox >= wo) continue;
      float best[8]; int pos[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) { best[e] = -INFINITY; pos[e] = -1; }
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        const int yy = 2 * oy + ky - 1;
        if (yy < 0 || yy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int xx = 2 * ox + kx - 1;
          if (xx < 0 || xx >= w) continue;
          float f[8];
          unpack8(__ldg(&x[((b * h + yy) * w + xx) * c8 + cc]), act_fmt, f);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (f[e] > best[e]) { best[e] = f[e]; pos[e] = ky * 3 + kx; }
        }
      }
      float g[8];
      unpack8(__ldg(&dy[((b * ho + oy) * wo + ox) * c8 + cc]), grad_fmt, g);
      // window (j + wy, i + wx) sees block pixel (a, bb) at tap (ky, kx) = (a + 1 - 2 wy, bb + 1 - 2 wx)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
          const int ky = a + 1 - 2 * wy, kx = bb + 1 - 2 * wx;
          if (ky < 0 || ky > 2 || kx < 0 || kx > 2) continue;
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (pos[e] == ky * 3 + kx) acc[a * 2 + bb][e] += g[e];
        }
    }
  }
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int bb = 0; bb < 2; ++bb) {
      const long long o = ((b * h + 2 * j + a) * w + 2 * i + bb) * c8 + cc;
      if (accumulate) {
        float f[8];
        unpack8(dx[o], grad_fmt, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[a * 2 + bb][e] += f[e];
      }
      dx[o] = pack8(acc[a * 2 + bb], grad_fmt);
    }
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam defaults: betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad;
// footprints/model_manager.py:27) over one flat fp32 parameter buffer.
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float beta1, float beta2, float eps,
                            float bc1, float bc2_sqrt, float grad_scale) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float gi = g[idx] * grad_scale;
  const float mi = beta1 * m[idx] + (1.f - beta1) * gi;
  const float vi = beta2 * v[idx] + (1.f - beta2) * gi * gi;
  m[idx] = mi;
  v[idx] = vi;
  // torch: step_size = lr / bias_correction1; denom = sqrt(v) / sqrt(bias_correction2) + eps
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[idx] -= (lr / bc1) * (mi / denom);
}

// same update with the per-step scalars read from device memory {lr, beta1, beta2, eps, bc1, sqrt(bc2), grad_scale}: a
// CUDA graph of the training step bakes kernel arguments in, these seven floats are refreshed by the host before each
// replay (StepLR changes lr, the bias corrections change every step)
__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, long long n, const float* __restrict__ hyper) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float lr = hyper[0], beta1 = hyper[1], beta2 = hyper[2], eps = hyper[3], bc1 = hyper[4], bc2_sqrt = hyper[5];
  const float gi = g[idx] * hyper[6];
  const float mi = beta1 * m[idx] + (1.f - beta1) * gi;
  const float vi = beta2 * v[idx] + (1.f - beta2) * gi * gi;
  m[idx] = mi;
  v[idx] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[idx] -= (lr / bc1) * (mi / denom);
}

static inline unsigned nblocks(long long n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }
static inline int red_blocks(long long M, int C, int blocks_per_sm = 4) {
  const int rows_per_pass = RED_THREADS / (C / 8);
  long long want = (M + rows_per_pass * 8 - 1) / (rows_per_pass * 8);
  if (want < 1) want = 1;
  const int cap = blocks_per_sm * 148 < RED_MAX_BLOCKS ? blocks_per_sm * 148 : RED_MAX_BLOCKS;   // one resident wave
  return static_cast<int>(want < cap ? want : cap);
}

}  // namespace fp

using namespace fp;

extern "C" int fp_reduce_workspace_floats(int C) { return RED_MAX_BLOCKS * 2 * C; }

extern "C" int fp_bn_stats(const void* z, long long M, int C, int dtype, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps, float* scale,
                           float* shift, float* save_mean, float* save_invstd, float* workspace, void* stream) {
  if (!z || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd || !workspace)
    return set_error(FP_ERR_ARG, "fp_bn_stats: null argument");
  if ((C & 7) || C > 2048 || RED_THREADS % (C / 8)) return set_error(FP_ERR_ARG, "fp_bn_stats: unsupported C=%d", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = red_blocks(M, C);
  const int rows_per_pass = RED_THREADS / (C / 8);
  col_reduce_kernel<0><<<nblk, RED_THREADS, rows_per_pass * 2 * C * sizeof(float), st>>>(
      static_cast<const uint4*>(z), dtype, nullptr, nullptr, dtype, nullptr, nullptr, M, C, workspace);
  FP_CHECK_LAUNCH("fp_bn_stats(reduce)");
  bn_finalize_kernel<<<C, SUM_THREADS, 0, st>>>(workspace, nblk, M, C, gamma, beta, running_mean, running_var,
                                                      momentum, eps, scale, shift, save_mean, save_invstd);
  FP_CHECK_LAUNCH("fp_bn_stats(finalize)");
  return FP_OK;
}

extern "C" int fp_bn_stats_split(const void* z_hi, const void* z_lo, long long M, int C, int dtype, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                 float* scale, float* shift, float* save_mean, float* save_invstd, float* workspace,
                                 void* stream) {
  if (!z_hi || !z_lo || !gamma || !beta || !scale || !shift || !save_mean || !save_invstd || !workspace)
    return set_error(FP_ERR_ARG, "fp_bn_stats_split: null argument");
  if ((C & 7) || C > 2048 || RED_THREADS % (C / 8)) return set_error(FP_ERR_ARG, "fp_bn_stats_split: unsupported C=%d", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = red_blocks(M, C);
  const int rows_per_pass = RED_THREADS / (C / 8);
  col_reduce_kernel<0><<<nblk, RED_THREADS, rows_per_pass * 2 * C * sizeof(float), st>>>(
      static_cast<const uint4*>(z_hi), dtype, static_cast<const uint4*>(z_lo), nullptr, dtype, nullptr, nullptr, M, C, workspace);
  FP_CHECK_LAUNCH("fp_bn_stats_split(reduce)");
  bn_finalize_kernel<<<C, SUM_THREADS, 0, st>>>(workspace, nblk, M, C, gamma, beta, running_mean, running_var,
                                                      momentum, eps, scale, shift, save_mean, save_invstd);
  FP_CHECK_LAUNCH("fp_bn_stats_split(finalize)");
  return FP_OK;
}

extern "C" int fp_bn_apply_split(const void* z_hi, const void* z_lo, const float* scale, const float* shift,
                                 const void* res_hi, const void* res_lo, int relu, long long M, int C, int dtype, void* y_hi,
                                 void* y_lo, void* stream) {
  if (!z_hi || !z_lo || !scale || !shift || !y_hi || !y_lo || (C & 7) || ((res_hi == nullptr) != (res_lo == nullptr)))
    return set_error(FP_ERR_ARG, "fp_bn_apply_split: bad argument");
  const long long n8 = M * (C / 8);
  bn_apply_split_kernel<<<nblocks(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(z_hi), static_cast<const uint4*>(z_lo), scale, shift, static_cast<const uint4*>(res_hi),
      static_cast<const uint4*>(res_lo), relu, n8, C / 8, dtype, static_cast<uint4*>(y_hi), static_cast<uint4*>(y_lo));
  FP_CHECK_LAUNCH("fp_bn_apply_split");
  return FP_OK;
}

extern "C" int fp_bn_apply(const void* z, const float* scale, const float* shift, const void* residual, int relu,
                           long long M, int C, int dtype, void* y, void* stream) {
  if (!z || !scale || !shift || !y || (C & 7)) return set_error(FP_ERR_ARG, "fp_bn_apply: bad argument");
  const long long n8 = M * (C / 8);
  bn_apply_kernel<<<nblocks(n8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(z), scale, shift, static_cast<const uint4*>(residual), relu, n8, C / 8, dtype,
      static_cast<uint4*>(y));
  FP_CHECK_LAUNCH("fp_bn_apply");
  return FP_OK;
}

extern "C" int fp_bn_bwd(const void* dy, const void* y, const void* z, const float* gamma, const float* save_mean,
                         const float* save_invstd, int relu, long long M, int C, int act_dtype, int grad_dtype,
                         void* dz, void* g_out, float* dgamma, float* dbeta, float* workspace, void* stream) {
  if (!dy || !z || !gamma || !save_mean || !save_invstd || !dz || !dgamma || !dbeta || !workspace || (relu && !y))
    return set_error(FP_ERR_ARG, "fp_bn_bwd: null argument");
  if ((C & 7) || RED_THREADS % (C / 8)) return set_error(FP_ERR_ARG, "fp_bn_bwd: unsupported C=%d", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = red_blocks(M, C, 3);
  const int rows_per_pass = RED_THREADS / (C / 8);
  col_reduce_kernel<1><<<nblk, RED_THREADS, rows_per_pass * 2 * C * sizeof(float), st>>>(
      static_cast<const uint4*>(dy), grad_dtype, relu ? static_cast<const uint4*>(y) : nullptr,
      static_cast<const uint4*>(z), act_dtype, save_mean, save_invstd, M, C, workspace);
  FP_CHECK_LAUNCH("fp_bn_bwd(reduce)");
  partial_sum_kernel<<<C, SUM_THREADS, 0, st>>>(workspace, nblk, C, dbeta, dgamma);
  FP_CHECK_LAUNCH("fp_bn_bwd(sum)");
  const long long n8 = M * (C / 8);
  bn_bwd_apply_kernel<<<nblocks(n8, 256), 256, 0, st>>>(
      static_cast<const uint4*>(dy), static_cast<const uint4*>(y), static_cast<const uint4*>(z), gamma, save_mean,
      save_invstd, dbeta, dgamma, 1.f / static_cast<float>(M), relu, n8, C / 8, act_dtype, grad_dtype,
      static_cast<uint4*>(dz), static_cast<uint4*>(g_out));
  FP_CHECK_LAUNCH("fp_bn_bwd(apply)");
  return FP_OK;
}

extern "C" int fp_bias_grad(const void* dz, long long M, int C, int grad_dtype, float* dbias, float* workspace,
                            void* stream) {
  if (!dz || !dbias || !workspace) return set_error(FP_ERR_ARG, "fp_bias_grad: null argument");
  if ((C & 7) || RED_THREADS % (C / 8)) return set_error(FP_ERR_ARG, "fp_bias_grad: unsupported C=%d", C);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nblk = red_blocks(M, C);
  const int rows_per_pass = RED_THREADS / (C / 8);
  col_reduce_kernel<2><<<nblk, RED_THREADS, rows_per_pass * 2 * C * sizeof(float), st>>>(
      static_cast<const uint4*>(dz), grad_dtype, nullptr, nullptr, grad_dtype, nullptr, nullptr, M, C, workspace);
  FP_CHECK_LAUNCH("fp_bias_grad(reduce)");
  partial_sum_kernel<<<C, SUM_THREADS, 0, st>>>(workspace, nblk, C, dbias, nullptr);
  FP_CHECK_LAUNCH("fp_bias_grad(sum)");
  return FP_OK;
}

extern "C" int fp_loss_fwd_bwd(const float* const* pred, float* const* grad, const float* depth, const float* visible,
                               const float* all_ground, const float* ground_depth, const float* moving,
                               const float* depth_mask, int batch, int h, int w, float min_depth, float max_depth,
                               float prior_weight, float* losses21, float* workspace, void* stream) {
  if (!pred || !depth || !visible || !all_ground || !ground_depth || !moving || !depth_mask || !losses21 || !workspace)
    return set_error(FP_ERR_ARG, "fp_loss_fwd_bwd: null argument");
  LossParams p{};
  for (int s = 0; s < 4; ++s) {
    if (!pred[s]) return set_error(FP_ERR_ARG, "fp_loss_fwd_bwd: null prediction %d", s);
    p.pred[s] = pred[s];
    p.grad[s] = grad ? grad[s] : nullptr;
  }
  p.depth = depth; p.visible = visible; p.all_ground = all_ground; p.ground_depth = ground_depth;
  p.moving = moving; p.depth_mask = depth_mask;
  p.B = batch; p.HW = static_cast<long long>(h) * w;
  p.min_disp = 1.f / max_depth;                       // utils.py:38-40
  p.disp_range = 1.f / min_depth - 1.f / max_depth;
  p.prior_w = prior_weight;
  const long long P = static_cast<long long>(batch) * h * w;
  p.gscale = static_cast<float>(1.0 / (4.0 * static_cast<double>(P)));
  long long want = (P + LOSS_THREADS * 4 - 1) / (LOSS_THREADS * 4);
  const int nblk = static_cast<int>(want < 1184 ? (want < 1 ? 1 : want) : 1184);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  loss_kernel<<<nblk, LOSS_THREADS, 0, st>>>(p, workspace);
  FP_CHECK_LAUNCH("fp_loss_fwd_bwd");
  loss_finalize_kernel<<<1, 32, 0, st>>>(workspace, nblk, 1.0 / static_cast<double>(P), losses21);
  FP_CHECK_LAUNCH("fp_loss_fwd_bwd(finalize)");
  return FP_OK;
}

extern "C" int fp_heads_bwd(const float* dout_nchw, int ch0, const float* head_post, int batch, int h, int w, int scale,
                            int grad_dtype, float grad_scale, void* dz8, void* stream) {
  if (!dout_nchw || !dz8 || scale < 1) return set_error(FP_ERR_ARG, "fp_heads_bwd: bad argument");
  const long long n = static_cast<long long>(batch) * h * w;
  heads_bwd_kernel<<<nblocks(n, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      dout_nchw, ch0, reinterpret_cast<const float2*>(head_post), batch, h, w, scale, grad_dtype, grad_scale,
      static_cast<uint4*>(dz8));
  FP_CHECK_LAUNCH("fp_heads_bwd");
  return FP_OK;
}

extern "C" int fp_fold_bwd(const void* dvp, int batch, int hv, int wv, int ccat, int coff, int c, int up,
                           int accumulate, const void* y_act, int act, int act_dtype, int grad_dtype, void* dst,
                           void* stream) {
  if (!dvp || !dst || (ccat & 7) || (coff & 7) || (c & 7) || hv < 2 || wv < 2)
    return set_error(FP_ERR_ARG, "fp_fold_bwd: bad argument");
  if (up && ((hv & 1) || (wv & 1))) return set_error(FP_ERR_ARG, "fp_fold_bwd: upsampled size must be even");
  const int hd = up ? hv / 2 : hv, wd = up ? wv / 2 : wv;
  const long long n = static_cast<long long>(batch) * hd * wd * (c / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (up)
    fold_kernel<true><<<nblocks(n, 256), 256, 0, st>>>(static_cast<const uint4*>(dvp), hv, wv, ccat / 8, coff / 8, batch,
                                                      hd, wd, c / 8, accumulate, static_cast<const uint4*>(y_act), act,
                                                      act_dtype, grad_dtype, static_cast<uint4*>(dst));
  else
    fold_kernel<false><<<nblocks(n, 256), 256, 0, st>>>(static_cast<const uint4*>(dvp), hv, wv, ccat / 8, coff / 8, batch,
                                                       hd, wd, c / 8, accumulate, static_cast<const uint4*>(y_act), act,
                                                       act_dtype, grad_dtype, static_cast<uint4*>(dst));
  FP_CHECK_LAUNCH("fp_fold_bwd");
  return FP_OK;
}

extern "C" int fp_act_bwd(const void* dy, const void* y, int act, long long n, int act_dtype, int grad_dtype, void* dz,
                          void* stream) {
  if (!dy || !y || !dz || (n & 7)) return set_error(FP_ERR_ARG, "fp_act_bwd: bad argument");
  act_bwd_kernel<<<nblocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(dy), static_cast<const uint4*>(y), act, n / 8, act_dtype, grad_dtype,
      static_cast<uint4*>(dz));
  FP_CHECK_LAUNCH("fp_act_bwd");
  return FP_OK;
}

extern "C" int fp_grad_add(void* a, const void* b, long long n, int grad_dtype, void* stream) {
  if (!a || !b || (n & 7)) return set_error(FP_ERR_ARG, "fp_grad_add: bad argument");
  grad_add_kernel<<<nblocks(n / 8, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(a), static_cast<const uint4*>(b), n / 8, grad_dtype);
  FP_CHECK_LAUNCH("fp_grad_add");
  return FP_OK;
}

extern "C" int fp_maxpool3x3s2_bwd(const void* x, const void* dy, int batch, int h, int w, int c, int act_dtype,
                                   int grad_dtype, int accumulate, void* dx, void* stream) {
  if (!x || !dy || !dx || (c & 7) || (h & 1) || (w & 1))
    return set_error(FP_ERR_ARG, "fp_maxpool3x3s2_bwd: bad argument (c %% 8, even h and w required)");
  const long long n = static_cast<long long>(batch) * (h / 2) * (w / 2) * (c / 8);
  maxpool_bwd_kernel<<<nblocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x), static_cast<const uint4*>(dy), batch, h, w, c / 8, act_dtype, grad_dtype, accumulate,
      static_cast<uint4*>(dx));
  FP_CHECK_LAUNCH("fp_maxpool3x3s2_bwd");
  return FP_OK;
}

extern "C" int fp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, float lr,
                            float beta1, float beta2, float eps, int step, float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || step < 1) return set_error(FP_ERR_ARG, "fp_adam_step: bad argument");
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_kernel<<<nblocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, static_cast<float>(bc1),
      static_cast<float>(sqrt(bc2)), grad_scale);
  FP_CHECK_LAUNCH("fp_adam_step");
  return FP_OK;
}

extern "C" int fp_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                const float* hyper7_dev, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper7_dev) return set_error(FP_ERR_ARG, "fp_adam_step_dev: bad argument");
  adam_dev_kernel<<<nblocks(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, hyper7_dev);
  FP_CHECK_LAUNCH("fp_adam_step_dev");
  return FP_OK;
}
