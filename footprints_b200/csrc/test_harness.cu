// Standalone GPU harness for libfootprints_b200.so (no torch): checks fp_conv_fwd against a naive direct
// convolution kernel on the same 16-bit-rounded operands, and times the large layer shapes.
//   usage: test_harness <case-index | all | perf>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "../../include/footprints_b200.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);  \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)
#define FPCK(x)                                                        \
  do {                                                                 \
    int r_ = (x);                                                      \
    if (r_ != 0) {                                                     \
      printf("fp error %d: %s (%s:%d)\n", r_, fp_last_error(), __FILE__, __LINE__); \
      exit(3);                                                         \
    }                                                                  \
  } while (0)

static uint32_t rng_state = 12345;
static float frand() {  // uniform [-1, 1)
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFFFF) / 8388608.0f - 1.0f;
}

__host__ __device__ inline float from16(uint16_t h, int fmt) {
#ifdef __CUDA_ARCH__
  if (fmt == FP_BF16) return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&h));
  return __half2float(*reinterpret_cast<const __half*>(&h));
#else
  if (fmt == FP_BF16) { uint32_t u = static_cast<uint32_t>(h) << 16; float f; memcpy(&f, &u, 4); return f; }
  return __half2float(*reinterpret_cast<const __half*>(&h));
#endif
}
static uint16_t to16h(float v, int fmt) {
  if (fmt == FP_BF16) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
  __half h = __float2half_rn(v);
  return *reinterpret_cast<uint16_t*>(&h);
}

// naive reference: one thread per (m, n)
__global__ void ref_conv(FpConvDesc d, const uint16_t* s0, const uint16_t* s1, const float* w /*OIHW, pre-rounded*/,
                         const float* scale, const float* shift, const uint16_t* res, float* out) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long M = (long long)d.batch * d.ho * d.wo;
  if (idx >= M * d.cout) return;
  const int n = (int)(idx % d.cout);
  const long long m = idx / d.cout;
  const int ox = (int)(m % d.wo), oy = (int)((m / d.wo) % d.ho), b = (int)(m / ((long long)d.wo * d.ho));
  const int Hin = d.up0 ? 2 * d.h0 : d.h0, Win = d.up0 ? 2 * d.w0 : d.w0;
  const int cin = d.c0 + d.c1;
  float acc = 0.f;
  for (int ky = 0; ky < d.kh; ++ky)
    for (int kx = 0; kx < d.kw; ++kx) {
      int iy = oy * d.stride - d.pad_t + ky, ix = ox * d.stride - d.pad_l + kx;
      if (d.pad_mode == FP_PAD_REFLECT) {
        if (iy < 0) iy = -iy; if (iy >= Hin) iy = 2 * Hin - 2 - iy;
        if (ix < 0) ix = -ix; if (ix >= Win) ix = 2 * Win - 2 - ix;
      } else if (iy < 0 || iy >= Hin || ix < 0 || ix >= Win) continue;
      for (int c = 0; c < cin; ++c) {
        float a;
        if (c < d.c0) {
          const int sy = d.up0 ? iy / 2 : iy, sx = d.up0 ? ix / 2 : ix;
          a = from16(s0[(((long long)b * d.h0 + sy) * d.w0 + sx) * d.c0 + c], d.dtype);
        } else {
          a = from16(s1[(((long long)b * Hin + iy) * Win + ix) * d.c1 + (c - d.c0)], d.dtype);
        }
        acc += a * w[(((long long)n * cin + c) * d.kh + ky) * d.kw + kx];
      }
    }
  float y = acc * (scale ? scale[n] : 1.f) + (shift ? shift[n] : 0.f);
  if (res) y += from16(res[m * d.res_ld + n], d.dtype);
  if (d.act == FP_ACT_RELU) y = fmaxf(y, 0.f);
  else if (d.act == FP_ACT_ELU) y = y > 0.f ? y : expm1f(y);
  else if (d.act == FP_ACT_SIGMOID) y = 1.f / (1.f + expf(-y));
  out[m * d.cout + n] = y;
}

struct Case {
  const char* name;
  int batch, c0, h0, w0, up0, c1, k, stride, pad, pad_mode, cout, bn, act, dtype, out_f32, use_res, use_scale, tma;
};

static int run_case(const Case& c, bool time_it) {
  FpConvDesc d{};
  d.batch = c.batch; d.c0 = c.c0; d.h0 = c.h0; d.w0 = c.w0; d.up0 = c.up0; d.c1 = c.c1;
  d.kh = d.kw = c.k; d.stride = c.stride; d.pad_t = d.pad_l = c.pad; d.pad_mode = c.pad_mode;
  const int Hin = c.up0 ? 2 * c.h0 : c.h0, Win = c.up0 ? 2 * c.w0 : c.w0;
  d.ho = (Hin + 2 * c.pad - c.k) / c.stride + 1; d.wo = (Win + 2 * c.pad - c.k) / c.stride + 1;
  d.cout = c.cout; d.bn = c.bn; d.act = c.act; d.dtype = c.dtype; d.out_f32 = c.out_f32;
  const int cin = c.c0 + c.c1;
  d.kpad = (c.k * c.k * cin + 63) / 64 * 64;
  d.npad = (c.cout + c.bn - 1) / c.bn * c.bn;
  d.out_ld = c.cout; d.res_ld = c.cout; d.b_via_tma = c.tma;
  const long long M = (long long)d.batch * d.ho * d.wo;
  printf("[%s] M=%lld N=%d K=%d (kpad %d) bn=%d dtype=%s tma=%d\n", c.name, M, d.cout, c.k * c.k * cin, d.kpad, d.bn,
         c.dtype ? "bf16" : "f16", c.tma);

  const size_t n0 = (size_t)c.batch * c.h0 * c.w0 * c.c0, n1 = (size_t)c.batch * Hin * Win * c.c1;
  std::vector<uint16_t> h0(n0), h1(n1 ? n1 : 1), hres((size_t)M * c.cout);
  for (auto& v : h0) v = to16h(frand(), c.dtype);
  for (auto& v : h1) v = to16h(frand(), c.dtype);
  for (auto& v : hres) v = to16h(frand(), c.dtype);
  const size_t nw = (size_t)c.cout * cin * c.k * c.k;
  std::vector<float> hw(nw), hscale(c.cout), hshift(c.cout);
  const float wstd = 1.0f / sqrtf((float)(cin * c.k * c.k));
  for (auto& v : hw) v = from16(to16h(frand() * wstd * 1.7f, c.dtype), c.dtype);
  for (auto& v : hscale) v = 0.75f + 0.5f * fabsf(frand());
  for (auto& v : hshift) v = 0.2f * frand();

  uint16_t *ds0, *ds1, *dres, *dwp; float *dw, *dscale, *dshift, *dref; void* dout;
  CK(cudaMalloc(&ds0, n0 * 2)); CK(cudaMalloc(&ds1, (n1 ? n1 : 1) * 2)); CK(cudaMalloc(&dres, hres.size() * 2));
  CK(cudaMalloc(&dw, nw * 4)); CK(cudaMalloc(&dwp, (size_t)d.npad * d.kpad * 2));
  CK(cudaMalloc(&dscale, c.cout * 4)); CK(cudaMalloc(&dshift, c.cout * 4));
  CK(cudaMalloc(&dref, (size_t)M * c.cout * 4)); CK(cudaMalloc(&dout, (size_t)M * c.cout * 4));
  CK(cudaMemcpy(ds0, h0.data(), n0 * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(ds1, h1.data(), h1.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dres, hres.data(), hres.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw, hw.data(), nw * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dscale, hscale.data(), c.cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dshift, hshift.data(), c.cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0xFF, (size_t)M * c.cout * 4));

  FPCK(fp_pack_conv_weight(dw, c.cout, cin, c.k, c.k, cin, d.npad, d.kpad, c.dtype, 0, dwp, nullptr));
  const float* sc = c.use_scale ? dscale : nullptr;
  const void* rs = c.use_res ? dres : nullptr;
  FPCK(fp_conv_fwd(&d, ds0, n1 ? ds1 : nullptr, dwp, sc, dshift, rs, dout, nullptr));
  CK(cudaDeviceSynchronize());

  int rc = 0;
  if (!time_it || M * c.cout <= (1ll << 26)) {
    const long long tot = M * c.cout;
    ref_conv<<<(unsigned)((tot + 255) / 256), 256>>>(d, ds0, ds1, dw, sc, dshift, (const uint16_t*)rs, dref);
    CK(cudaDeviceSynchronize());
    std::vector<float> ref(tot);
    CK(cudaMemcpy(ref.data(), dref, tot * 4, cudaMemcpyDeviceToHost));
    std::vector<float> got(tot);
    if (c.out_f32) CK(cudaMemcpy(got.data(), dout, tot * 4, cudaMemcpyDeviceToHost));
    else {
      std::vector<uint16_t> g16(tot);
      CK(cudaMemcpy(g16.data(), dout, tot * 2, cudaMemcpyDeviceToHost));
      for (long long i = 0; i < tot; ++i) got[i] = from16(g16[i], c.dtype);
    }
    double max_err = 0, max_ref = 0; long long bad = 0, first_bad = -1;
    const double tol_rel = c.out_f32 ? 2e-4 : (c.dtype == FP_BF16 ? 1.0 / 128 : 1.0 / 1024);
    for (long long i = 0; i < tot; ++i) {
      const double e = fabs((double)got[i] - ref[i]);
      if (!(e <= 2e-3 + tol_rel * fabs(ref[i]))) { if (first_bad < 0) first_bad = i; ++bad; }
      if (e > max_err || e != e) max_err = e;
      if (fabs(ref[i]) > max_ref) max_ref = fabs(ref[i]);
    }
    printf("  max_abs_err=%.3e max_ref=%.3e bad=%lld/%lld %s\n", max_err, max_ref, bad, tot, bad ? "FAIL" : "PASS");
    if (bad) {
      printf("  first bad idx %lld (m=%lld n=%lld): got %f ref %f\n", first_bad, first_bad / c.cout, first_bad % c.cout,
             got[first_bad], ref[first_bad]);
      rc = 1;
    }
  }
  if (getenv("FP_TRACE")) {
    static long long tr[3][2048];
    FPCK(fp_conv_fwd(&d, ds0, n1 ? ds1 : nullptr, dwp, sc, dshift, rs, dout, nullptr));
    CK(cudaDeviceSynchronize());
    FPCK(fp_debug_read_trace(&tr[0][0]));
    const long long t0 = tr[0][0];
    printf("  trace (cycles rel. to first producer stamp): kb: producer_free  mma_full | tile: epi_start epi_end\n");
    for (int i = 0; i < 64; ++i) printf("   it %3d  P %8lld  M %8lld\n", i, tr[0][i] - t0, tr[1][i] - t0);
    for (int i = 0; i < 8; ++i) printf("   tile %d  E %8lld .. %8lld\n", i, tr[2][2 * i] - t0, tr[2][2 * i + 1] - t0);
  }
  if (time_it) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) FPCK(fp_conv_fwd(&d, ds0, n1 ? ds1 : nullptr, dwp, sc, dshift, rs, dout, nullptr));
    const int iters = 20;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) FPCK(fp_conv_fwd(&d, ds0, n1 ? ds1 : nullptr, dwp, sc, dshift, rs, dout, nullptr));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); ms /= iters;
    const double flops = 2.0 * M * c.cout * c.k * c.k * cin;
    const double bytes = (double)(n0 + n1) * 2 + (double)M * c.cout * (c.out_f32 ? 4 : 2) + (c.use_res ? (double)M * c.cout * 2 : 0);
    printf("  time %.3f ms  %.1f TFLOP/s  %.1f GB/s (algorithmic)\n", ms, flops / ms * 1e-9, bytes / ms * 1e-6);
  }
  cudaFree(ds0); cudaFree(ds1); cudaFree(dres); cudaFree(dw); cudaFree(dwp); cudaFree(dscale); cudaFree(dshift);
  cudaFree(dref); cudaFree(dout);
  return rc;
}

//            name                   B  c0  h0  w0 up c1 k s p pm  cout bn act dt f32 res scale tma
static const Case kCases[] = {
    {"c64_3x3_f16_cpasyncB",         2, 64, 24, 40, 0, 0, 3, 1, 1, 0, 64, 64, 1, 0, 0, 0, 1, 0},
    {"c64_3x3_f16_tmaB",             2, 64, 24, 40, 0, 0, 3, 1, 1, 0, 64, 64, 1, 0, 0, 0, 1, 1},
    {"c64_3x3_bf16_tmaB",            2, 64, 24, 40, 0, 0, 3, 1, 1, 0, 64, 64, 1, 1, 0, 0, 1, 1},
    {"c128_3x3_s2_res_bn128",        2, 64, 24, 40, 0, 0, 3, 2, 1, 0, 128, 128, 1, 0, 0, 1, 1, 1},
    {"c64_1x1_s2_bn128",             2, 64, 24, 40, 0, 0, 1, 2, 0, 0, 128, 128, 0, 0, 0, 0, 1, 1},
    {"up_concat_reflect_elu",        2, 64, 12, 20, 1, 64, 3, 1, 1, 1, 64, 64, 2, 0, 0, 0, 0, 1},
    {"c256_bn256_mtail",             3, 256, 6, 10, 0, 0, 3, 1, 1, 0, 256, 256, 1, 0, 0, 1, 1, 1},
    {"head_c64_n2_sigmoid_f32",      2, 64, 24, 40, 0, 0, 3, 1, 1, 1, 2, 16, 3, 0, 1, 0, 0, 1},
    {"c32_kpad_reflect_elu_bn32",    2, 32, 24, 40, 0, 0, 3, 1, 1, 1, 32, 32, 2, 0, 0, 0, 0, 1},
    {"c512_bn256_two_ntiles",        2, 512, 6, 20, 0, 0, 3, 1, 1, 0, 512, 256, 1, 0, 0, 1, 1, 1},
    {"s2d_like_c16_4x4",             2, 16, 24, 40, 0, 0, 4, 1, 2, 0, 64, 64, 1, 0, 0, 0, 1, 1},
    {"multi_tile_persistent",        8, 64, 48, 160, 0, 0, 3, 1, 1, 0, 64, 64, 1, 0, 0, 1, 1, 1},
    // wide rows -> row-strip kernel (shifted UMMA descriptors)
    {"strip_reflect_c64_w320",       1, 64, 6, 320, 0, 0, 3, 1, 1, 1, 64, 64, 2, 0, 0, 0, 0, 1},
    {"strip_upcat_reflect_w320",     1, 64, 3, 160, 1, 64, 3, 1, 1, 1, 64, 64, 2, 0, 0, 0, 0, 1},
    {"strip_fullcorr_pad2_w322",     2, 64, 5, 320, 0, 0, 3, 1, 2, 0, 128, 128, 0, 1, 0, 1, 1, 1},
    {"strip_c128_bn32_zero_w640",    1, 128, 4, 640, 0, 0, 3, 1, 1, 0, 32, 32, 1, 0, 0, 0, 1, 1},
    {"strip_up_c64_bn32_w640",       2, 64, 3, 320, 1, 0, 3, 1, 1, 1, 32, 32, 2, 0, 0, 0, 0, 1},
    // 32-channel rows -> SWIZZLE_64B strips
    {"strip64B_c32_reflect_w640",    1, 32, 5, 640, 0, 0, 3, 1, 1, 1, 32, 32, 2, 0, 0, 0, 0, 1},
    {"strip64B_c32_fullcorr_w322",   2, 32, 4, 320, 0, 0, 3, 1, 2, 0, 64, 64, 0, 1, 0, 1, 1, 1},
    {"strip64B_c32_cat32_zero_w320", 1, 32, 6, 320, 0, 32, 3, 1, 1, 0, 32, 32, 1, 0, 0, 0, 1, 1},
};
static const Case kPerf[] = {
    {"perf_layer1_64x48x160_c64",    64, 64, 48, 160, 0, 0, 3, 1, 1, 0, 64, 64, 1, 0, 0, 1, 1, 1},
    {"perf_layer2_c128",             64, 128, 24, 80, 0, 0, 3, 1, 1, 0, 128, 128, 1, 0, 0, 1, 1, 1},
    {"perf_layer3_c256",             64, 256, 12, 40, 0, 0, 3, 1, 1, 0, 256, 256, 1, 0, 0, 1, 1, 1},
    {"perf_layer3_c256_bn128",       64, 256, 12, 40, 0, 0, 3, 1, 1, 0, 256, 128, 1, 0, 0, 1, 1, 1},
    {"perf_layer4_c512",             64, 512, 6, 20, 0, 0, 3, 1, 1, 0, 512, 256, 1, 0, 0, 1, 1, 1},
    {"perf_layer4_c512_bn128",       64, 512, 6, 20, 0, 0, 3, 1, 1, 0, 512, 128, 1, 0, 0, 1, 1, 1},
    {"perf_block4_post2_96x320_c64", 64, 64, 96, 320, 0, 0, 3, 1, 1, 1, 64, 64, 2, 0, 0, 0, 0, 1},
    {"perf_block4_post1_upcat",      64, 64, 48, 160, 1, 64, 3, 1, 1, 1, 64, 64, 2, 0, 0, 0, 0, 1},
    {"perf_outconv4_c1_up_64to32",   64, 64, 96, 320, 1, 0, 3, 1, 1, 1, 32, 32, 2, 0, 0, 0, 0, 1},
    {"perf_outconv4_c2_32to32",      64, 32, 192, 640, 0, 0, 3, 1, 1, 1, 32, 32, 2, 0, 0, 0, 0, 1},
    {"perf_head_32to2",              64, 32, 192, 640, 0, 0, 3, 1, 1, 1, 2, 16, 3, 0, 1, 0, 0, 1},
};

int main(int argc, char** argv) {
  const std::string arg = argc > 1 ? argv[1] : "all";
  if (fp_check_device() != 0) { printf("device check failed: %s\n", fp_last_error()); return 4; }
  const int ncases = (int)(sizeof(kCases) / sizeof(kCases[0]));
  const int nperf = (int)(sizeof(kPerf) / sizeof(kPerf[0]));
  int rc = 0;
  if (arg == "all") { for (int i = 0; i < ncases; ++i) rc |= run_case(kCases[i], false); }
  else if (arg == "perf") { for (int i = 0; i < nperf; ++i) rc |= run_case(kPerf[i], true); }
  else if (arg == "count") { printf("%d\n", ncases); return 0; }
  else if (arg.rfind("perf", 0) == 0) { int i = atoi(arg.c_str() + 4); if (i >= 0 && i < nperf) rc = run_case(kPerf[i], true); }
  else { int i = atoi(arg.c_str()); if (i < 0 || i >= ncases) return 5; rc = run_case(kCases[i], false); }
  printf(rc ? "HARNESS FAIL\n" : "HARNESS OK\n");
  return rc;
}
