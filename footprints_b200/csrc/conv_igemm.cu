// Implicit-GEMM convolution for sm_100a: tcgen05.mma (UMMA) with TMEM accumulators.
//
//   D[M = batch*ho*wo][N = cout] = A[M][K = taps*cin] * W[N][K]^T,  fp32 accumulate, 16-bit operands.
//
// Persistent, warp-specialised CTA (one per SM, 544 threads):
//   warps 0-7  producers : gather the im2col A tile (128 pixels x 64 K-elements per stage) with 16-byte
//                          cp.async straight into the 128B-swizzled K-major layout UMMA reads; per-row
//                          tap tables (make_row_info) implement zero / reflection padding, stride, the
//                          nearest x2 upsample of source 0 and the channel concat with source 1, so none
//                          of those tensors is ever materialised (reference: footprints/network.py:124-158).
//                          Weights (B tile, BN x 64) arrive by TMA (cp.async.bulk.tensor, SWIZZLE_128B).
//   warp  8    MMA issuer: one elected lane issues 4 x tcgen05.mma (M128 x BN x K16) per stage and
//                          tcgen05.commit's the stage back to the producers / the accumulator to the epilogue.
//   warps 9-16 epilogue  : tcgen05.ld the fp32 accumulator (double-buffered in TMEM so the next tile's
//                          mainloop overlaps; the buffer is released right after the TMEM reads), apply
//                          scale/shift (+residual) + activation, store NHWC.
//
// Two kernels share this structure:
//   conv_igemm_kernel<BN>           any convolution (1x1 .. 4x4, stride 1/2, transposed stride 2): per-tap gather.
//   conv_igemm_strip_kernel<BN,CB>  3x3 stride-1 on wide rows (fp_conv_fwd picks it when >= 80 % of the 128-pixel row
//                                   segments are real pixels): one staged 130-pixel input row segment per (dy, channel
//                                   block) serves the three dx taps through shifted UMMA descriptors; strips by TMA
//                                   for zero-padded single-source inputs; the filter resident in shared memory.
//   conv_rowstream_kernel<BN,CB>    3x3 stride-1, Cout <= 64, rows >= 80 pixels: the default for the thin high-resolution
//                                   layers (documented at the kernel).
// What bounds them (measured): the deep layers L2->SM operand bandwidth (~50 B/clk/SM against the 94 B/clk a
// 128x256 tile consumes at the tensor pipe's rate); the thin layers the epilogue's instruction stream -- neither HBM (DRAM
// traffic = algorithmic bytes, 18-39 % busy) nor the tensor pipe (14-49 %).
#include "fp_ptx.cuh"
#include "fp_internal.h"
#include "igemm_common.cuh"

#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace fp {

constexpr int BM = 128;          // output pixels per tile (UMMA M)
constexpr int BK = 64;           // K elements per stage = one 128-byte swizzle row
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int NUM_PRODUCER_WARPS = 8;
constexpr int NUM_PRODUCER_THREADS = NUM_PRODUCER_WARPS * 32;
constexpr int MMA_WARP = NUM_PRODUCER_WARPS;
constexpr int EPI_FIRST_WARP = MMA_WARP + 1;
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_EPI_THREADS = NUM_EPI_WARPS * 32;
constexpr int NUM_THREADS = (EPI_FIRST_WARP + NUM_EPI_WARPS) * 32;   // 544
constexpr int ROWS_PER_THREAD = BM / (NUM_PRODUCER_THREADS / 8);     // 4
constexpr int KTAB_ENTRIES = 2048;   // one entry per 16-byte K chunk: kpad <= 16384 (split precision triples K)

// perf triage hooks (FP_DEBUG_FLAGS: 1 skip A gather, 2 skip MMA, 4 skip epilogue math, 8 skip B load, 16 record
// clock64() stamps of CTA 0 into g_trace[role][event]) are compiled in only with -DFP_TRIAGE (make TRIAGE=1):
// even as dead branches they cost instructions in the producers' latency-bound K loop.
#ifdef FP_TRIAGE
constexpr bool kTriage = true;
#else
constexpr bool kTriage = false;
#endif
__device__ long long g_trace[3][2048];

template <int BN>
struct IgemmCfg {
  static constexpr int B_STAGE_BYTES = BN * BK * 2;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator stages, power of two
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) +
                                    (2 * STAGES + 4) * 8 + 32 + 2 * BN * 4 + 2 * BM * 16 + KTAB_ENTRIES * 4;
};

struct IgemmParams {
  GatherGeom g;
  const uint8_t* w_packed;   // only used when !b_via_tma
  const float* scale;
  const float* shift;
  const uint8_t* residual;
  const uint8_t* residual_lo;   // split precision: low half of the residual
  void* out;
  void* out_lo;                 // split precision: low half of the output (y = hi + lo, both 16-bit)
  int N;
  int num_kb, kpad, m_tiles, n_tiles, total_tiles;
  int act, fmt, w_fmt, out_f32, out_ld, res_ld, b_via_tma;
  int a_mode;                // 0: cp.async gather (any channel count), 1: TMA tile::gather4 (C0, C1 multiples of 64)
  int rows0, rows1;          // pixel counts of the two sources (first out-of-range TMA row index)
  int debug;                 // FP_DEBUG_FLAGS (perf triage only): 1 skip A gather, 2 skip MMA, 4 skip epilogue math, 8 skip B load
};


// exp(x) by ex2.approx (one FMUL + one MUFU); |rel err| ~ 2^-22, far below the 16-bit output rounding
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
// Branch-free activations with the kind as a compile-time constant: the epilogue is on the critical path of the
// small-N layers (one accumulator drain per 128x64 tile), so no per-element jump table / divergent branches.
template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == FP_ACT_RELU) return fmaxf(x, 0.f);
  if (ACT == FP_ACT_ELU) return x > 0.f ? x : fast_exp(x) - 1.f;          // select, both sides evaluated
  if (ACT == FP_ACT_SIGMOID) return __frcp_rn(1.f + fast_exp(-x));
  return x;
}

template <int FMT>
__device__ __forceinline__ uint32_t pack2t(float a, float b) {
  if (FMT == FP_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return fp::f32x2_to_f16x2_sat(a, b);
}
template <int FMT>
__device__ __forceinline__ float2 unpack2t(uint32_t u) {
  if (FMT == FP_BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

__device__ __forceinline__ uint32_t pack2(float a, float b, int fmt) {
  if (fmt == FP_BF16) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  return fp::f32x2_to_f16x2_sat(a, b);
}
__device__ __forceinline__ float2 unpack2(uint32_t u, int fmt) {
  if (fmt == FP_BF16) return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
  return __half22float2(*reinterpret_cast<__half2*>(&u));
}

// y = act(acc*scale + shift (+ residual)) -> 16-bit or fp32 NHWC row, for 16 accumulator columns held in registers.
template <int ACT, int FMT>
__device__ __forceinline__ void epilogue_store16(const IgemmParams& p, const uint32_t* v, int j, int n_base, int m,
                                                 const float* s_scale, const float* s_shift) {
  constexpr int CW = 16;
  const int n0 = n_base + j;
  if (n0 >= p.N) return;
  if (p.out_f32 && p.N - n0 == 2 && p.out_ld == 2 && p.residual == nullptr) {
    // the 2-channel heads (N padded to 16): touch only the two real columns, one 8-byte store per pixel
    const float a = apply_act<ACT>(__uint_as_float(v[0]) * lds_f1(s_scale + j) + lds_f1(s_shift + j));
    const float b = apply_act<ACT>(__uint_as_float(v[1]) * lds_f1(s_scale + j + 1) + lds_f1(s_shift + j + 1));
    *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * 2 + n0) = make_float2(a, b);
    return;
  }
  float f[CW];
#pragma unroll
  for (int x = 0; x < CW; x += 4) {        // per-channel scale / shift as 128-bit shared-memory loads (j is a multiple of 16)
    const float4 sh = lds_f4(s_shift + j + x);
    if (p.scale != nullptr) {
      const float4 sc = lds_f4(s_scale + j + x);
      f[x] = __uint_as_float(v[x]) * sc.x + sh.x; f[x + 1] = __uint_as_float(v[x + 1]) * sc.y + sh.y;
      f[x + 2] = __uint_as_float(v[x + 2]) * sc.z + sh.z; f[x + 3] = __uint_as_float(v[x + 3]) * sc.w + sh.w;
    } else {                                // decoder layers (bias only) and raw outputs: no multiply
      f[x] = __uint_as_float(v[x]) + sh.x; f[x + 1] = __uint_as_float(v[x + 1]) + sh.y;
      f[x + 2] = __uint_as_float(v[x + 2]) + sh.z; f[x + 3] = __uint_as_float(v[x + 3]) + sh.w;
    }
  }
  const int ncols = min(CW, p.N - n0);
  if (p.residual != nullptr) {
    const uint8_t* rp = p.residual + (static_cast<size_t>(m) * p.res_ld + n0) * 2;
    if (ncols == CW) {
#pragma unroll
      for (int x = 0; x < CW / 8; ++x) {
        const uint4 r = *reinterpret_cast<const uint4*>(rp + x * 16);
        float2 t;
        t = unpack2t<FMT>(r.x); f[x * 8 + 0] += t.x; f[x * 8 + 1] += t.y;
        t = unpack2t<FMT>(r.y); f[x * 8 + 2] += t.x; f[x * 8 + 3] += t.y;
        t = unpack2t<FMT>(r.z); f[x * 8 + 4] += t.x; f[x * 8 + 5] += t.y;
        t = unpack2t<FMT>(r.w); f[x * 8 + 6] += t.x; f[x * 8 + 7] += t.y;
      }
    } else {
#pragma unroll
      for (int x = 0; x < CW; ++x)
        if (x < ncols) f[x] += unpack2t<FMT>(*reinterpret_cast<const uint16_t*>(rp + x * 2)).x;
    }
    if (p.residual_lo != nullptr) {
      const uint8_t* rl = p.residual_lo + (static_cast<size_t>(m) * p.res_ld + n0) * 2;
#pragma unroll
      for (int x = 0; x < CW; ++x)
        if (x < ncols) f[x] += unpack2t<FMT>(*reinterpret_cast<const uint16_t*>(rl + x * 2)).x;
    }
  }
#pragma unroll
  for (int x = 0; x < CW; ++x) f[x] = apply_act<ACT>(f[x]);
  if (p.out_f32) {
    float* op = reinterpret_cast<float*>(p.out) + static_cast<size_t>(m) * p.out_ld + n0;
    if (ncols == CW && (p.out_ld & 3) == 0) {
#pragma unroll
      for (int x = 0; x < CW / 4; ++x)
        *reinterpret_cast<float4*>(op + 4 * x) = make_float4(f[4 * x], f[4 * x + 1], f[4 * x + 2], f[4 * x + 3]);
    } else {
#pragma unroll
      for (int x = 0; x < CW; ++x) if (x < ncols) op[x] = f[x];
    }
  } else {
    uint8_t* op = reinterpret_cast<uint8_t*>(p.out) + (static_cast<size_t>(m) * p.out_ld + n0) * 2;
    if (p.out_lo != nullptr) {       // split precision: lo = y - hi, rounded once more (the difference is exact in fp32)
      uint8_t* ol = reinterpret_cast<uint8_t*>(p.out_lo) + (static_cast<size_t>(m) * p.out_ld + n0) * 2;
      float l[CW];
#pragma unroll
      for (int x = 0; x < CW; x += 2) {
        const float2 hf = unpack2t<FMT>(pack2t<FMT>(f[x], f[x + 1]));
        l[x] = f[x] - hf.x;
        l[x + 1] = f[x + 1] - hf.y;
      }
      if (ncols == CW && (p.out_ld & 7) == 0) {
#pragma unroll
        for (int x = 0; x < CW / 8; ++x) {
          uint4 o, q;
          o.x = pack2t<FMT>(f[8 * x], f[8 * x + 1]); o.y = pack2t<FMT>(f[8 * x + 2], f[8 * x + 3]);
          o.z = pack2t<FMT>(f[8 * x + 4], f[8 * x + 5]); o.w = pack2t<FMT>(f[8 * x + 6], f[8 * x + 7]);
          q.x = pack2t<FMT>(l[8 * x], l[8 * x + 1]); q.y = pack2t<FMT>(l[8 * x + 2], l[8 * x + 3]);
          q.z = pack2t<FMT>(l[8 * x + 4], l[8 * x + 5]); q.w = pack2t<FMT>(l[8 * x + 6], l[8 * x + 7]);
          *reinterpret_cast<uint4*>(op + 16 * x) = o;
          *reinterpret_cast<uint4*>(ol + 16 * x) = q;
        }
      } else {
#pragma unroll
        for (int x = 0; x < CW; ++x)
          if (x < ncols) {
            *reinterpret_cast<uint16_t*>(op + 2 * x) = static_cast<uint16_t>(pack2t<FMT>(f[x], 0.f) & 0xFFFF);
            *reinterpret_cast<uint16_t*>(ol + 2 * x) = static_cast<uint16_t>(pack2t<FMT>(l[x], 0.f) & 0xFFFF);
          }
      }
    } else if (ncols == CW && (p.out_ld & 7) == 0) {
#pragma unroll
      for (int x = 0; x < CW / 8; ++x) {
        uint4 o;
        o.x = pack2t<FMT>(f[8 * x], f[8 * x + 1]);
        o.y = pack2t<FMT>(f[8 * x + 2], f[8 * x + 3]);
        o.z = pack2t<FMT>(f[8 * x + 4], f[8 * x + 5]);
        o.w = pack2t<FMT>(f[8 * x + 6], f[8 * x + 7]);
        *reinterpret_cast<uint4*>(op + 16 * x) = o;
      }
    } else {
#pragma unroll
      for (int x = 0; x < CW; ++x)
        if (x < ncols) *reinterpret_cast<uint16_t*>(op + 2 * x) = static_cast<uint16_t>(pack2t<FMT>(f[x], 0.f) & 0xFFFF);
    }
  }
}

// Drain NC accumulator columns of this thread's row (TMEM lane) starting at column `col0`.  The columns are pulled into
// registers in passes of <= 32; as soon as the LAST pass's TMEM reads have completed the accumulator buffer is handed
// back to the MMA warp (arrive on `release_bar`), so scale/shift, residual, activation, packing and the global stores
// run off the tensor pipe's critical path (they used to delay the next-but-one tile's first MMA).
template <int ACT, int FMT, int NC>
__device__ __forceinline__ void epilogue_cols(const IgemmParams& p, uint32_t tbase, int col0, int n_base, int m, bool row_ok,
                                              const float* s_scale, const float* s_shift, uint64_t* release_bar,
                                              int lane) {
  constexpr int PASS = NC > 32 ? 32 : NC;
#pragma unroll 1
  for (int c0 = 0; c0 < NC; c0 += PASS) {
    uint32_t v[PASS];
#pragma unroll
    for (int j = 0; j < PASS; j += 16) tmem_ld_32x32b_x16(tbase + col0 + c0 + j, *reinterpret_cast<uint32_t(*)[16]>(&v[j]));
    tmem_ld_wait();
    if (c0 + PASS >= NC) {
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(release_bar);
    }
    if (row_ok) {
#pragma unroll
      for (int j = 0; j < PASS; j += 16) epilogue_store16<ACT, FMT>(p, &v[j], col0 + c0 + j, n_base, m, s_scale, s_shift);
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_kernel(const __grid_constant__ IgemmParams p, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_a0, const __grid_constant__ CUtensorMap tmap_a1) {
  using Cfg = IgemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int B_STAGE_BYTES = Cfg::B_STAGE_BYTES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);   // SWIZZLE_128B needs 1024-B alignment
  uint8_t* a_smem = smem;
  uint8_t* b_smem = smem + STAGES * A_STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(b_smem + STAGES * B_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 4) + 15) & ~static_cast<uintptr_t>(15));   // float4 loads
  float* s_shift = s_scale + BN;
  uint4* s_rowinfo = reinterpret_cast<uint4*>(s_shift + BN);   // [2][BM] {base0, code0, base1, code1}
  uint32_t* s_ktab = reinterpret_cast<uint32_t*>(s_rowinfo + 2 * BM);   // per 16-byte K chunk: tap/channel decode

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      // cp.async gather: one async arrive per producer thread (+1 for the weight TMA's expect_tx);
      // TMA gather: a single expect_tx arrive, everything else is transaction bytes
      mbar_init(&full_bar[s], p.a_mode == 1 ? 1 : NUM_PRODUCER_THREADS + (p.b_via_tma ? 1 : 0));
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    fence_barrier_init();
    if (p.b_via_tma) tma_prefetch_desc(&tmap_w);
    if (p.a_mode == 1) { tma_prefetch_desc(&tmap_a0); if (p.g.C1 > 0) tma_prefetch_desc(&tmap_a1); }
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  for (int j = tid; j < p.num_kb * 8; j += NUM_THREADS) s_ktab[j] = make_ktab_entry(p.g, j, p.num_kb * 8);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (p.a_mode == 1 && warp < NUM_PRODUCER_WARPS) {
    // =============================== PRODUCERS (TMA tile::gather4) ===============================
    // Warps 0-3: warp w owns K blocks kb = w (mod 4); lane L fetches rows 4L..4L+3 of the tile with one gather4
    // (four source pixels x 64 channels = 4 x 128 B, swizzled by the TMA unit).  Completion is pure transaction
    // bytes on the stage's mbarrier, so nobody waits for loads and the MMA warp sees one arrival per stage.
    if (warp < 4) {
      uint32_t ti = 0;
      uint32_t it_base = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti, it_base += p.num_kb) {
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        uint4* info_buf = s_rowinfo + (ti & 1) * BM;
        {
          const int m = m_tile * BM + tid;      // tid < 128 here
          uint4 ri = make_uint4(0u, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
          if (m < p.g.M) ri = make_row_info_for_pixel(p.g, m, true);
          info_buf[tid] = ri;
        }
        named_bar_sync(2, 128);
        uint4 info[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) info[i] = info_buf[4 * lane + i];
        for (int kb = warp; kb < p.num_kb; kb += 4) {
          const uint32_t it = it_base + kb;
          const int s = it % STAGES;
          if (lane == 0) {
            mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1, 1);
            if ((kTriage && p.debug & 16) && blockIdx.x == 0 && it < 2048) g_trace[0][it] = clock64();
            mbar_arrive_expect_tx(&full_bar[s], A_STAGE_BYTES + (p.b_via_tma ? B_STAGE_BYTES : 0));
            if (p.b_via_tma)
              tma_load_2d(smem_u32(b_smem) + s * B_STAGE_BYTES, &tmap_w, &full_bar[s], kb * BK, n_tile * BN);
          }
          __syncwarp();
          const KDecode kd = decode_ktab(p.g, s_ktab[kb * 8]);
          const int pitch = kd.second ? p.g.Win : p.g.W0;
          const int oob = kd.second ? p.rows1 : p.rows0;
          const int r0 = gather_row_index(kd, info[0], pitch, oob), r1 = gather_row_index(kd, info[1], pitch, oob);
          const int r2 = gather_row_index(kd, info[2], pitch, oob), r3 = gather_row_index(kd, info[3], pitch, oob);
          tma_gather4_2d(smem_u32(a_smem) + s * A_STAGE_BYTES + lane * 512, kd.second ? &tmap_a1 : &tmap_a0, &full_bar[s],
                         static_cast<int>(kd.cc), r0, r1, r2, r3);
        }
      }
    }
  } else if (warp < NUM_PRODUCER_WARPS) {
    // =============================== PRODUCERS (cp.async gather) ===============================
    const int c = tid & 7;    // 16-byte chunk (8 channels) within the 128-byte K row
    const int r0 = tid >> 3;  // rows r0 + 32*i, i = 0..3
    const uint32_t swz = static_cast<uint32_t>(r0 * 128 + ((c ^ (r0 & 7)) << 4));
    const uint32_t a_dst0 = smem_u32(a_smem) + swz;
    const uint32_t b_dst0 = smem_u32(b_smem) + swz;
    uint32_t it = 0, ti = 0;
    int s = 0;                 // running stage index and phase (no div/mod in the K loop)
    uint32_t ph = 0;
    // per-source constants of the gather (selected per K chunk by the table's source bit)
    const uint8_t* const sbase0 = p.g.src0;
    const uint8_t* const sbase1 = p.g.src1;
    const uint32_t pstride0 = p.g.C0, pstride1 = p.g.C1;
    const uint32_t rstride0 = p.g.W0 * p.g.C0, rstride1 = p.g.Win * p.g.C1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      uint4* info_buf = s_rowinfo + (ti & 1) * BM;
      if (tid < BM) {
        const int m = m_tile * BM + tid;
        uint4 ri = make_uint4(0u, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
        if (m < p.g.M) ri = make_row_info_for_pixel(p.g, m);
        info_buf[tid] = ri;
      }
      named_bar_sync(2, NUM_PRODUCER_THREADS);
      uint4 info[ROWS_PER_THREAD];
#pragma unroll
      for (int i = 0; i < ROWS_PER_THREAD; ++i) info[i] = info_buf[r0 + 32 * i];

      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        // every lane polls (uniform, no divergence); the slot is almost always free already
        mbar_wait(&empty_bar[s], ph ^ 1, 1);
        if ((kTriage && p.debug & 16) && blockIdx.x == 0 && tid == 0 && it < 2048) g_trace[0][it] = clock64();
        if (p.b_via_tma) {
          if (tid == 0) {
            if (kTriage && p.debug & 8) mbar_arrive(&full_bar[s]);
            else {
              mbar_arrive_expect_tx(&full_bar[s], B_STAGE_BYTES);
              tma_load_2d(smem_u32(b_smem) + s * B_STAGE_BYTES, &tmap_w, &full_bar[s], kb * BK, n_tile * BN);
            }
          }
        } else {
          const uint8_t* wsrc = p.w_packed + (static_cast<size_t>(n_tile) * BN * p.kpad + kb * BK + c * 8) * 2;
#pragma unroll
          for (int i = 0; i < (BN + 31) / 32; ++i) {
            const int r = r0 + 32 * i;
            if (r < BN) cp_async_16(b_dst0 + s * B_STAGE_BYTES + i * 4096, wsrc + static_cast<size_t>(r) * p.kpad * 2, 16);
          }
        }
        // decode this thread's K chunk (tap, channel slice, source) from the table built at kernel start
        const uint32_t ke = s_ktab[kb * 8 + c];
        KDecode kd;
        kd.second = (ke & (1u << 10)) != 0u;
        kd.padmask = (ke >> 8) & 8u;
        kd.shy = ke & 31u; kd.shx = (ke >> 5) & 31u; kd.cc = (ke >> 12) & 0x7FFFFu;
        kd.sbase = (ke >> 31) ? (kd.second ? p.g.src1_lo : p.g.src0_lo) : (kd.second ? sbase1 : sbase0);
        kd.pstride = kd.second ? pstride1 : pstride0;
        kd.rstride = kd.second ? rstride1 : rstride0;
        const uint32_t dst = a_dst0 + s * A_STAGE_BYTES;
        if (!(kTriage && p.debug & 1)) {
#pragma unroll
          for (int i = 0; i < ROWS_PER_THREAD; ++i) {
            bool ok;
            const uint32_t off = gather_offset(kd, info[i], ok);
            cp_async_16_zfill(dst + i * 4096, kd.sbase + static_cast<size_t>(off) * 2, ok);
          }
        }
        // the stage is published to the MMA warp when every producer thread's copies have landed
        cp_async_mbar_arrive_noinc(&full_bar[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BM, BN, p.fmt, p.w_fmt, 0, 0);
      uint32_t it = 0, ti = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1, 2);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1, 3);
          if ((kTriage && p.debug & 16) && blockIdx.x == 0 && it < 2048) g_trace[1][it] = clock64();
          tcgen05_fence_after();
          const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(a_smem) + s * A_STAGE_BYTES, 16, 1024);
          const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(b_smem) + s * B_STAGE_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (kTriage && p.debug & 2) break;
            // advancing 16 K-elements = 32 bytes inside the swizzle row = +2 in the (addr >> 4) field
            umma_f16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE ===============================
    // 8 warps: warp%4 selects the TMEM lane quadrant it may read, (warp-EPI_FIRST_WARP)/4 the column half.
    const int q = warp & 3;
    const int half = (warp - EPI_FIRST_WARP) >> 2;
    const int row = q * 32 + lane;
    const int e = tid - EPI_FIRST_WARP * 32;
    constexpr int COLS_PER_HALF = (BN >= 32) ? BN / 2 : BN;
    const int col_begin = (BN >= 32) ? half * COLS_PER_HALF : 0;
    const bool active = (BN >= 32) || (half == 0);
    uint32_t ti = 0;
    int cur_n = -1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const uint32_t acc = ti & 1;
      if (n_tile != cur_n) {     // (re)load the per-channel scale / shift only when the channel tile changes: the two
        cur_n = n_tile;          // barriers and the global-load latency sat on every tile's epilogue critical path
        named_bar_sync(1, NUM_EPI_THREADS);        // everyone finished reading the previous tile's scale/shift
        for (int j = e; j < BN; j += NUM_EPI_THREADS) {
          const int n = n_tile * BN + j;
          s_scale[j] = (p.scale != nullptr && n < p.N) ? p.scale[n] : 1.f;
          s_shift[j] = (p.shift != nullptr && n < p.N) ? p.shift[n] : 0.f;
        }
        named_bar_sync(1, NUM_EPI_THREADS);
      }
      mbar_wait<32>(&tmem_full_bar[acc], (ti >> 1) & 1, 4);
      if ((kTriage && p.debug & 16) && blockIdx.x == 0 && e == 0 && ti < 1024) g_trace[2][2 * ti] = clock64();
      tcgen05_fence_after();
      const int m = m_tile * BM + row;
      const bool m_ok = m < p.g.M;
      if (active) {
        const uint32_t tbase = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        const bool skip = (kTriage && p.debug & 4) != 0;
#define FP_EPI(A, F) epilogue_cols<A, F, COLS_PER_HALF>(p, tbase, col_begin, n_tile * BN, m, m_ok && !skip, s_scale, s_shift, &tmem_empty_bar[acc], lane)
        if (p.fmt == FP_BF16) {
          switch (p.act) {
            case FP_ACT_RELU: FP_EPI(FP_ACT_RELU, FP_BF16); break;
            case FP_ACT_ELU: FP_EPI(FP_ACT_ELU, FP_BF16); break;
            case FP_ACT_SIGMOID: FP_EPI(FP_ACT_SIGMOID, FP_BF16); break;
            default: FP_EPI(FP_ACT_NONE, FP_BF16); break;
          }
        } else {
          switch (p.act) {
            case FP_ACT_RELU: FP_EPI(FP_ACT_RELU, FP_F16); break;
            case FP_ACT_ELU: FP_EPI(FP_ACT_ELU, FP_F16); break;
            case FP_ACT_SIGMOID: FP_EPI(FP_ACT_SIGMOID, FP_F16); break;
            default: FP_EPI(FP_ACT_NONE, FP_F16); break;
          }
        }
#undef FP_EPI
      }
      if (!active) {          // BN < 32: the second column half has nothing to read, it only releases the buffer
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
      if ((kTriage && p.debug & 16) && blockIdx.x == 0 && e == 0 && ti < 1024) g_trace[2][2 * ti + 1] = clock64();
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// --------------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------------
template <int BN>
static int launch_igemm(const IgemmParams& p, const CUtensorMap& tmap, const CUtensorMap& tmap_a0,
                        const CUtensorMap& tmap_a1, cudaStream_t stream) {
  using Cfg = IgemmCfg<BN>;
  auto kern = conv_igemm_kernel<BN>;
  { const int arc = ensure_max_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_BYTES, "kern"); if (arc != FP_OK) return arc; }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  kern<<<grid, NUM_THREADS, Cfg::SMEM_BYTES, stream>>>(p, tmap, tmap_a0, tmap_a1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "conv_igemm launch (BN=%d): %s", BN, cudaGetErrorString(e));
  return FP_OK;
}

// ==================================================================================================
// Row-strip variant for 3x3 stride-1 convolutions on wide images (the 96x320 / 192x640 layers that carry most of
// the network's FLOPs).  A tile is 128 consecutive output pixels of ONE image row.  For each filter row dy and each
// 64-channel block the producers stage the input row segment once — 130 pixels x 64 channels, 128-byte rows, 128B
// swizzle — and the three horizontal taps dx = 0,1,2 are three UMMA operands that start 0, 1 and 2 rows (128 bytes)
// into that strip: the tensor core applies the swizzle to absolute shared-memory addresses (probe_umma_shift.cu),
// so a K-major SW128 descriptor may begin at any 128-byte row.  Shared-memory fills and L2->SM traffic drop from 9 to
// 3 (+1.6 %) per tile, which is what bounds these layers.  Padding (zero / reflection), the nearest x2 upsample of
// source 0 and the channel concat are resolved per strip pixel exactly as in the general kernel.
// ==================================================================================================
constexpr int STRIP_PIX = BM + 2;            // 130 input pixels feed 128 outputs of a 3-tap row
constexpr int STRIP_ROWS = 136;              // rounded up to whole 8-row swizzle groups

// CB = channels per strip row: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B — the 32-channel
// layers of the 192x640 tail; both layouts verified by probe_umma_shift.cu)
template <int BN, int CB>
struct StripCfg {
  static constexpr int ROW_BYTES = CB * 2;
  static constexpr int A_BYTES = (STRIP_ROWS * ROW_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_TILE = BN * ROW_BYTES;                        // one dx tap's weights [BN][CB]
  static constexpr int B_BYTES = 3 * B_TILE;
  static constexpr int STAGE_BYTES = A_BYTES + (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGES = (CB == 32) ? 8 : ((BN <= 32) ? 6 : (BN <= 64 ? 5 : 3));
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  static constexpr int MAX_STAGES = 8;
  // everything but the stage ring and the resident filter: alignment slack, mbarriers, TMEM pointer, scale/shift, maps
  static constexpr int FIXED_BYTES = 1024 + (2 * MAX_STAGES + 5) * 8 + 32 + 2 * BN * 4 + 2 * (2 * STRIP_ROWS + 8) * 4;
  static constexpr int SMEM_MAX = 232448;      // 227 KB opt-in limit per CTA
};

struct StripParams {
  IgemmParams b;             // geometry, epilogue, tile counts (m_tiles = batch*Ho*tiles_per_row)
  int tiles_per_row;         // ceil(Wo / 128)
  int cblocks;               // (C0 + C1) / CB
  int x_via_tma;             // strips by one 4-D TMA box each (zero padding = out-of-bounds fill; single source, no upsample)
  // weight residency: when the layer's whole packed filter (9 * cblocks tiles of [BN][CB]) fits next to the strip
  // ring it is loaded ONCE per CTA instead of once per output tile.  The thin 192x640 / 96x320 layers are bound by
  // L2->SM bandwidth (~20 B/clk/SM with all SMs streaming), and re-fetching 18-147 KB of weights for every 128-pixel
  // tile was 40-60 % of that traffic.
  int w_resident;
  int stages;                // pipeline depth actually used (<= StripCfg::MAX_STAGES)
  int stage_bytes;           // A_BYTES when resident (no per-stage weight area), else STAGE_BYTES
  int w_bytes;               // resident filter bytes (0 when streaming)
};

template <int BN, int CB>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_igemm_strip_kernel(const __grid_constant__ StripParams sp, const __grid_constant__ CUtensorMap tmap_w,
                        const __grid_constant__ CUtensorMap tmap_x) {
  using Cfg = StripCfg<BN, CB>;
  constexpr int ROW_BYTES = Cfg::ROW_BYTES;
  constexpr int CPR = ROW_BYTES / 16;                    // 16-byte chunks per strip row
  constexpr int ROWS_PER_PASS = NUM_PRODUCER_THREADS / CPR;
  constexpr uint64_t LAYOUT = (CB == 64) ? 2ull : 4ull;  // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  constexpr uint32_t SBO = 8 * ROW_BYTES;                // 8-row swizzle group
  constexpr int MAX_STAGES = Cfg::MAX_STAGES;
  const int STAGES = sp.stages;
  const int STAGE_BYTES = sp.stage_bytes;
  const IgemmParams& p = sp.b;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* stage0 = smem;                                               // [STAGES][strip (| B0 | B1 | B2 when streaming)]
  uint8_t* w_res = smem + STAGES * STAGE_BYTES;                         // resident filter: [9 * cblocks][BN][CB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(w_res + sp.w_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* tmem_full_bar = empty_bar + MAX_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* w_bar = tmem_empty_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 4) + 15) & ~static_cast<uintptr_t>(15));   // float4 loads
  float* s_shift = s_scale + BN;
  // per tile: element offset of every strip pixel inside its source row (0xFFFFFFFF = padding zero), for both
  // sources, and the element offset of the three source rows (double-buffered)
  uint32_t* s_xoff = reinterpret_cast<uint32_t*>(s_shift + BN);        // [2][2*STRIP_ROWS + 8]

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], sp.x_via_tma ? 1 : NUM_PRODUCER_THREADS + 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], NUM_EPI_WARPS);
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_w);
    if (sp.x_via_tma) tma_prefetch_desc(&tmap_x);
    if (sp.w_resident) {
      mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(sp.w_bytes));
      for (int t = 0; t < 9 * sp.cblocks; ++t)
        tma_load_2d(smem_u32(w_res) + t * Cfg::B_TILE, &tmap_w, w_bar, t * CB, 0);
    }
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  const int stages_per_tile = 3 * sp.cblocks;

  if (warp < NUM_PRODUCER_WARPS && sp.x_via_tma) {
    // =============================== PRODUCER (TMA) ===============================
    // one thread: per stage a {CB, 130, 1, 1} box of the NHWC input lands in the strip layout (the tensor map's swizzle
    // is the strip's), rows / columns outside the image are zero-filled by the copy engine
    if (tid == 0) {
      int s = 0;
      uint32_t ph = 0;
      const GatherGeom& g = p.g;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        const int xseg = m_tile % sp.tiles_per_row;
        const int yrow = m_tile / sp.tiles_per_row;          // b * Ho + y
        const int img = yrow / g.Ho;
        const int y = yrow - img * g.Ho;
        for (int dy = 0; dy < 3; ++dy)
          for (int cb = 0; cb < sp.cblocks; ++cb) {
            mbar_wait(&empty_bar[s], ph ^ 1, 21);
            const uint32_t stage = smem_u32(stage0) + s * STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[s], (sp.w_resident ? 0 : Cfg::B_BYTES) + STRIP_PIX * ROW_BYTES);
            tma_load_4d(stage, &tmap_x, &full_bar[s], cb * CB, xseg * BM - g.pad_l, y - g.pad_t + dy, img);
            if (!sp.w_resident) {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
                tma_load_2d(stage + Cfg::A_BYTES + dx * Cfg::B_TILE, &tmap_w, &full_bar[s],
                            ((dy * 3 + dx) * sp.cblocks + cb) * CB, n_tile * BN);
            }
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp < NUM_PRODUCER_WARPS) {
    // =============================== PRODUCERS ===============================
    const int c = tid % CPR;   // 16-byte chunk of the strip row
    const int r0 = tid / CPR;  // strip rows r0 + ROWS_PER_PASS*i
    uint32_t ti = 0;
    int s = 0;
    uint32_t ph = 0;
    const GatherGeom& g = p.g;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const int xseg = m_tile % sp.tiles_per_row;
      const int yrow = m_tile / sp.tiles_per_row;          // b * Ho + y
      const int img = yrow / g.Ho;
      const int y = yrow - img * g.Ho;
      const int x0 = xseg * BM;
      uint32_t* xo = s_xoff + (ti & 1) * (2 * STRIP_ROWS + 8);
      if (tid < STRIP_ROWS) {
        // column map of strip pixel j: virtual x = x0 - pad_l + j
        int vx = x0 - g.pad_l + tid;
        bool ok = tid < STRIP_PIX;
        if (g.pad_mode == FP_PAD_REFLECT) vx = vx < 0 ? -vx : (vx >= g.Win ? 2 * g.Win - 2 - vx : vx);
        ok = ok && static_cast<unsigned>(vx) < static_cast<unsigned>(g.Win);
        xo[tid] = ok ? static_cast<uint32_t>((g.up0 ? vx >> 1 : vx) * g.C0) : 0xFFFFFFFFu;
        xo[STRIP_ROWS + tid] = ok ? static_cast<uint32_t>(vx * g.C1) : 0xFFFFFFFFu;
      } else if (tid < STRIP_ROWS + 3) {
        // row map of filter row dy: virtual y = y - pad_t + dy -> element offset of that source row
        const int dy = tid - STRIP_ROWS;
        int vy = y - g.pad_t + dy;
        if (g.pad_mode == FP_PAD_REFLECT) vy = vy < 0 ? -vy : (vy >= g.Hin ? 2 * g.Hin - 2 - vy : vy);
        const bool ok = static_cast<unsigned>(vy) < static_cast<unsigned>(g.Hin);
        xo[2 * STRIP_ROWS + dy] = ok ? static_cast<uint32_t>(((img * g.H0 + (g.up0 ? vy >> 1 : vy)) * g.W0) * g.C0) : 0xFFFFFFFFu;
        xo[2 * STRIP_ROWS + 4 + dy] = ok ? static_cast<uint32_t>(((img * g.Hin + vy) * g.Win) * g.C1) : 0xFFFFFFFFu;
      }
      named_bar_sync(2, NUM_PRODUCER_THREADS);
      for (int dy = 0; dy < 3; ++dy) {
        for (int cb = 0; cb < sp.cblocks; ++cb) {
          mbar_wait(&empty_bar[s], ph ^ 1, 21);
          const uint32_t stage = smem_u32(stage0) + s * STAGE_BYTES;
          if (tid == 0) {
            if (sp.w_resident) {
              mbar_arrive(&full_bar[s]);
            } else {
              mbar_arrive_expect_tx(&full_bar[s], Cfg::B_BYTES);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx)
                tma_load_2d(stage + Cfg::A_BYTES + dx * Cfg::B_TILE, &tmap_w, &full_bar[s],
                            ((dy * 3 + dx) * sp.cblocks + cb) * CB, n_tile * BN);
            }
          }
          const int ch = cb * CB;
          const bool second = ch >= g.C0;
          const uint8_t* sbase = second ? g.src1 : g.src0;
          const uint32_t rowoff = xo[2 * STRIP_ROWS + (second ? 4 : 0) + dy];
          const uint32_t* xcol = xo + (second ? STRIP_ROWS : 0);
          const uint32_t cc = static_cast<uint32_t>(second ? ch - g.C0 : ch) + c * 8;
#pragma unroll
          for (int i = 0; i < (STRIP_PIX + ROWS_PER_PASS - 1) / ROWS_PER_PASS; ++i) {
            const int r = r0 + ROWS_PER_PASS * i;
            if (r < STRIP_PIX) {
              const uint32_t xoff = xcol[r];
              const bool ok = (rowoff != 0xFFFFFFFFu) && (xoff != 0xFFFFFFFFu);
              const uint32_t off = ok ? rowoff + xoff + cc : 0u;
              // absolute-address swizzle: SW128: chunk ^ (row & 7); SW64: chunk ^ ((row >> 1) & 3)
              const int sw = (CB == 64) ? (c ^ (r & 7)) : (c ^ ((r >> 1) & 3));
              cp_async_16_zfill(stage + r * ROW_BYTES + (sw << 4), sbase + static_cast<size_t>(off) * 2, ok);
            }
          }
          cp_async_mbar_arrive_noinc(&full_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(BM, BN, p.fmt, p.w_fmt, 0, 0);
      uint32_t ti = 0;
      int s = 0;
      uint32_t ph = 0;
      if (sp.w_resident) mbar_wait(w_bar, 0, 25);
      constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (static_cast<uint32_t>(LAYOUT) << 29);   // SBO, version 1, layout
      auto mk_desc = [](uint32_t lo) -> uint64_t { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; };
      const uint32_t stage0_lo = ((smem_u32(stage0) >> 4) & 0x3FFF) | (1u << 16);                 // LBO = 16 bytes
      const uint32_t wres_lo = ((smem_u32(w_res) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t stage_step = static_cast<uint32_t>(STAGE_BYTES) >> 4;
      const uint32_t b_step = sp.w_resident ? sp.cblocks * (Cfg::B_TILE >> 4) : (Cfg::B_TILE >> 4);
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1;
        mbar_wait(&tmem_empty_bar[acc], ((ti >> 1) & 1) ^ 1, 22);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t b_tile = 0;                       // resident filter tile (dy*3*cblocks + cb) in units of B_TILE/16
        int cbi = 0;
        for (int st = 0; st < stages_per_tile; ++st) {
          mbar_wait(&full_bar[s], ph, 23);
          tcgen05_fence_after();
          // descriptor low words: (address >> 4) | LBO field; all strip / filter addresses are < 256 KB, so the tap,
          // K-step and stage offsets are plain additions (no re-masking) -- the single issuing thread is the critical
          // resource on the thin layers, every instruction between two UTCHMMAs counts
          const uint32_t a_lo = stage0_lo + static_cast<uint32_t>(s) * stage_step;
          const uint32_t b_lo = sp.w_resident ? wres_lo + b_tile : a_lo + (Cfg::A_BYTES >> 4);
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            // the tap's A operand = the strip advanced by dx pixels (whole rows)
#pragma unroll
            for (int k = 0; k < CB / 16; ++k)
              umma_f16(d_tmem, mk_desc(a_lo + dx * (ROW_BYTES >> 4) + 2 * k), mk_desc(b_lo + dx * b_step + 2 * k), idesc,
                       (dx | k) != 0 ? 1u : (st != 0 ? 1u : 0u));
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
          // next stage = next channel block, then next filter row (tile index (dy*3 + dx)*cblocks + cb)
          if (++cbi == sp.cblocks) { cbi = 0; b_tile += (2 * sp.cblocks + 1) * (Cfg::B_TILE >> 4); }
          else b_tile += Cfg::B_TILE >> 4;
        }
        umma_commit(&tmem_full_bar[acc]);
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE ===============================
    const int q = warp & 3;
    const int half = (warp - EPI_FIRST_WARP) >> 2;
    const int row = q * 32 + lane;
    const int e = tid - EPI_FIRST_WARP * 32;
    constexpr int COLS_PER_HALF = (BN >= 32) ? BN / 2 : BN;
    const int col_begin = (BN >= 32) ? half * COLS_PER_HALF : 0;
    const bool active = (BN >= 32) || (half == 0);
    uint32_t ti = 0;
    int cur_n = -1;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++ti) {
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const int xseg = m_tile % sp.tiles_per_row;
      const int yrow = m_tile / sp.tiles_per_row;
      const int x0 = xseg * BM;
      const uint32_t acc = ti & 1;
      if (n_tile != cur_n) {
        cur_n = n_tile;
        named_bar_sync(1, NUM_EPI_THREADS);
        for (int j = e; j < BN; j += NUM_EPI_THREADS) {
          const int n = n_tile * BN + j;
          s_scale[j] = (p.scale != nullptr && n < p.N) ? p.scale[n] : 1.f;
          s_shift[j] = (p.shift != nullptr && n < p.N) ? p.shift[n] : 0.f;
        }
        named_bar_sync(1, NUM_EPI_THREADS);
      }
      mbar_wait<32>(&tmem_full_bar[acc], (ti >> 1) & 1, 24);
      tcgen05_fence_after();
      const int m = yrow * p.g.Wo + x0 + row;
      const bool m_ok = (x0 + row) < p.g.Wo;
      if (active) {
        const uint32_t tbase = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
#define FP_EPI(A, F) epilogue_cols<A, F, COLS_PER_HALF>(p, tbase, col_begin, n_tile * BN, m, m_ok, s_scale, s_shift, &tmem_empty_bar[acc], lane)
        if (p.fmt == FP_BF16) {
          switch (p.act) {
            case FP_ACT_RELU: FP_EPI(FP_ACT_RELU, FP_BF16); break;
            case FP_ACT_ELU: FP_EPI(FP_ACT_ELU, FP_BF16); break;
            case FP_ACT_SIGMOID: FP_EPI(FP_ACT_SIGMOID, FP_BF16); break;
            default: FP_EPI(FP_ACT_NONE, FP_BF16); break;
          }
        } else {
          switch (p.act) {
            case FP_ACT_RELU: FP_EPI(FP_ACT_RELU, FP_F16); break;
            case FP_ACT_ELU: FP_EPI(FP_ACT_ELU, FP_F16); break;
            case FP_ACT_SIGMOID: FP_EPI(FP_ACT_SIGMOID, FP_F16); break;
            default: FP_EPI(FP_ACT_NONE, FP_F16); break;
          }
        }
#undef FP_EPI
      }
      if (!active) {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ==================================================================================================
// Row-streaming 3x3 convolution: the DEFAULT for Cout <= 64 on rows of >= 80 pixels (every 48x160 / 96x320 / 192x640
// layer, forward and data gradient, and the 2-channel heads with N padded to 16); FP_ROWSTREAM=0 falls back to the strip
// kernel.  Validated on B200 against the strip kernel and torch (tests/test_gpu_kernel_variants.py).
//
// Why: on the thin layers (Cout <= 64) a tcgen05.mma costs max(M, N)/2 = 64 cycles whatever N is, and L2->SM bandwidth
// caps the strip fills; the strip kernel stages every input row three times (once per output row that uses it) and
// issues 9*C/16 N = Cout instructions per tile.  Here a CTA walks DOWN a column of 128-pixel segments: every input row
// segment is staged once and, per (dx, K step), ONE instruction with N = 3*BN accumulates it into the three output
// rows it feeds -- the B operand is [W(dy=0,dx) | W(dy=1,dx) | W(dy=2,dx)] (resident, stored in that order) and the
// three accumulators are adjacent TMEM column blocks ordered by DEcreasing output row.  TMEM (512 columns) is split
// in two halves of H = 256/BN row slots; a work item is H consecutive output rows of one column (H + 2 input rows,
// i.e. 1.25x (BN=32) / 1.5x (BN=64) fills per output row instead of 3x, and 3x fewer instructions).  A CTA works on a
// PAIR of items at a time, one per TMEM half, each with its own MMA-issuing warp and two epilogue groups (see below).
//   input row j of an item (virtual row v = yc - pad_t + j, j = 0 .. H+1) feeds output rows y = yc + j - dy:
//   dy in [max(0, j-H+1), min(2, j)];  slot(y) = H-1-(y-yc)  (so dy = 0,1,2 are increasing columns).
//   The first instruction touching a row slot (dy = 0 block, cb = dx = k = 0) is issued alone with accumulate = 0.
// ==================================================================================================
struct RowStreamParams {
  IgemmParams b;
  int tiles_per_row;         // ceil(Wo / 128)
  int cblocks;               // (C0 + C1) / CB
  int x_via_tma;
  int stages;                // strip ring depth (<= StripCfg::MAX_STAGES)
  int w_bytes;               // resident filter bytes: 9 * cblocks * B_TILE
  int H;                     // output rows per item = row slots per TMEM half
  int chunks_per_col;        // ceil(Ho / H)
  int total_items;           // batch * tiles_per_row * chunks_per_col
};

// warp roles of the row-streaming kernel: 4 producer warps (input strips: by TMA one thread suffices, the cp.async gather
// is not what paces the kernel either), 2 MMA-issuing warps (one per TMEM half / item of the pair), 16 epilogue warps =
// 4 groups of 4 (one warp per TMEM lane quadrant).  Triage on B200 (scripts/triage_rowstream.py, outconv4.0.conv2
// forward, 0.37 ms): without the epilogue's math + stores 0.22 ms, without the MMAs 0.30, without the loads 0.33, the
// bare barrier skeleton 0.11 -- a row's drain is ~330 dependent instructions per thread (bias, ELU = 5 ops with a MUFU,
// pack, stores), i.e. ~1200 cycles per warp, so the number of rows drained CONCURRENTLY sets the pace: four groups.
constexpr int RS_PRODUCER_WARPS = 4;
constexpr int RS_PRODUCER_THREADS = RS_PRODUCER_WARPS * 32;
constexpr int RS_MMA_WARP = RS_PRODUCER_WARPS;                            // warps 4, 5
constexpr int RS_EPI_FIRST_WARP = RS_MMA_WARP + 2;
constexpr int RS_EPI_WARPS = 16;
constexpr int RS_EPI_THREADS = RS_EPI_WARPS * 32;
constexpr int RS_THREADS = (RS_EPI_FIRST_WARP + RS_EPI_WARPS) * 32;       // 704

template <int BN, int CB>
__global__ void __launch_bounds__(RS_THREADS, 1)
conv_rowstream_kernel(const __grid_constant__ RowStreamParams rp, const __grid_constant__ CUtensorMap tmap_w,
                      const __grid_constant__ CUtensorMap tmap_x) {
  using Cfg = StripCfg<BN, CB>;
  constexpr int ROW_BYTES = Cfg::ROW_BYTES;
  constexpr int CPR = ROW_BYTES / 16;
  constexpr int ROWS_PER_PASS = RS_PRODUCER_THREADS / CPR;
  constexpr uint64_t LAYOUT = (CB == 64) ? 2ull : 4ull;
  constexpr uint32_t SBO = 8 * ROW_BYTES;
  constexpr int MAX_STAGES = Cfg::MAX_STAGES;
  constexpr int SLOTS = 512 / BN;                    // row slots in TMEM (two halves of H = SLOTS / 2)
  const int STAGES = rp.stages;
  const int H = rp.H;
  const IgemmParams& p = rp.b;
  const GatherGeom& g = p.g;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* stage0 = smem;                                               // [STAGES][strip]
  uint8_t* w_res = smem + STAGES * Cfg::A_BYTES;                        // [(dx*cblocks + cb)*3 + dy][BN][CB]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(w_res + rp.w_bytes);
  uint64_t* empty_bar = full_bar + MAX_STAGES;
  uint64_t* row_full = empty_bar + MAX_STAGES;                          // [SLOTS] MMA -> epilogue, one per output-row slot
  uint64_t* row_empty = row_full + SLOTS;                               // [SLOTS] epilogue -> MMA
  uint64_t* w_bar = row_empty + SLOTS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(w_bar + 1);
  float* s_scale = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_ptr_smem + 4) + 15) & ~static_cast<uintptr_t>(15));   // float4 loads
  float* s_shift = s_scale + BN;
  uint32_t* s_xoff = reinterpret_cast<uint32_t*>(s_shift + BN);        // [2 pair parities][2 items][2*STRIP_ROWS] column maps

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], rp.x_via_tma ? 1 : RS_PRODUCER_THREADS);
      mbar_init(&empty_bar[s], 1);
    }
    for (int q = 0; q < SLOTS; ++q) {
      mbar_init(&row_full[q], 1);
      mbar_init(&row_empty[q], NUM_EPI_WARPS / 2);         // the four warps of the group that drains this row
    }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_w);
    if (rp.x_via_tma) tma_prefetch_desc(&tmap_x);
    // resident filter, regrouped so that the three dy tiles of a (dx, cb) are adjacent: one N = 3*BN B operand
    mbar_arrive_expect_tx(w_bar, static_cast<uint32_t>(rp.w_bytes));
    for (int dx = 0; dx < 3; ++dx)
      for (int cb = 0; cb < rp.cblocks; ++cb)
        for (int dy = 0; dy < 3; ++dy)
          tma_load_2d(smem_u32(w_res) + ((dx * rp.cblocks + cb) * 3 + dy) * Cfg::B_TILE, &tmap_w, w_bar,
                      ((dy * 3 + dx) * rp.cblocks + cb) * CB, 0);
  }
  if (warp == RS_MMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // item -> (image, row segment, first output row)
  auto decode = [&](int item, int& img, int& x0, int& yc) {
    const int chunk = item % rp.chunks_per_col;
    const int t = item / rp.chunks_per_col;
    x0 = (t % rp.tiles_per_row) * BM;
    img = t / rp.tiles_per_row;
    yc = chunk * H;
  };

  // A CTA works on PAIRS of items at once: item 2*pair + sub lives in TMEM half `sub`, is issued by MMA warp `sub` and
  // drained by epilogue group `sub`; the producers alternate the two items' rows through one stage ring (even stages =
  // sub 0, odd = sub 1).  Why two issuers: on the thin layers a row costs 6-12 tcgen05.mma (~340-700 tensor cycles) but
  // the ~150 dependent instructions one thread spends per row (waits, descriptor arithmetic, the ELECT wrappers ptxas
  // puts around each MMA) took ~1300 cycles (ncu r02j: tensor pipe 18-29 %, DRAM 18-31 %, 8.4 cycles per issued
  // instruction): the single issuing thread, not the tensor pipe or memory, paced the kernel.
  const int total_pairs = (rp.total_items + 1) >> 1;

  if (warp < RS_PRODUCER_WARPS && rp.x_via_tma) {
    // =============================== PRODUCER (TMA) ===============================
    if (tid == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x) {
        int img[2], x0[2], yc[2];
        bool ex[2];
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          ex[sub] = 2 * pair + sub < rp.total_items;
          decode(ex[sub] ? 2 * pair + sub : 2 * pair, img[sub], x0[sub], yc[sub]);
        }
        for (int j = 0; j < H + 2; ++j)
          for (int cb = 0; cb < rp.cblocks; ++cb)
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              mbar_wait(&empty_bar[s], ph ^ 1, 41);
              if (ex[sub] && !(kTriage && (p.debug & 1))) {
                mbar_arrive_expect_tx(&full_bar[s], STRIP_PIX * ROW_BYTES);
                tma_load_4d(smem_u32(stage0) + s * Cfg::A_BYTES, &tmap_x, &full_bar[s], cb * CB, x0[sub] - g.pad_l,
                            yc[sub] - g.pad_t + j, img[sub]);
              } else {
                mbar_arrive(&full_bar[s]);                 // odd item count: the missing partner's stages stay empty
              }
              if (++s == STAGES) { s = 0; ph ^= 1; }
            }
      }
    }
  } else if (warp < RS_PRODUCER_WARPS) {
    // =============================== PRODUCERS (cp.async) ===============================
    const int c = tid % CPR;
    const int r0 = tid / CPR;
    int s = 0;
    uint32_t ph = 0, pit = 0;
    for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x, ++pit) {
      int img[2], x0[2], yc[2];
      bool ex[2];
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        ex[sub] = 2 * pair + sub < rp.total_items;
        decode(ex[sub] ? 2 * pair + sub : 2 * pair, img[sub], x0[sub], yc[sub]);
      }
      uint32_t* xo2 = s_xoff + (pit & 1) * (4 * STRIP_ROWS);       // [sub][2 * STRIP_ROWS]
      for (int t = tid; t < STRIP_ROWS; t += RS_PRODUCER_THREADS) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          int vx = x0[sub] - g.pad_l + t;
          bool ok = t < STRIP_PIX;
          if (g.pad_mode == FP_PAD_REFLECT) vx = vx < 0 ? -vx : (vx >= g.Win ? 2 * g.Win - 2 - vx : vx);
          ok = ok && static_cast<unsigned>(vx) < static_cast<unsigned>(g.Win);
          xo2[sub * 2 * STRIP_ROWS + t] = ok ? static_cast<uint32_t>((g.up0 ? vx >> 1 : vx) * g.C0) : 0xFFFFFFFFu;
          xo2[sub * 2 * STRIP_ROWS + STRIP_ROWS + t] = ok ? static_cast<uint32_t>(vx * g.C1) : 0xFFFFFFFFu;
        }
      }
      named_bar_sync(2, RS_PRODUCER_THREADS);
      for (int j = 0; j < H + 2; ++j) {
        for (int cb = 0; cb < rp.cblocks; ++cb) {
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            mbar_wait(&empty_bar[s], ph ^ 1, 42);
            if (ex[sub] && !(kTriage && (p.debug & 1))) {
              int vy = yc[sub] - g.pad_t + j;
              if (g.pad_mode == FP_PAD_REFLECT) vy = vy < 0 ? -vy : (vy >= g.Hin ? 2 * g.Hin - 2 - vy : vy);
              const bool row_ok = static_cast<unsigned>(vy) < static_cast<unsigned>(g.Hin);   // rows past a ragged item: zeros
              const uint32_t* xo = xo2 + sub * 2 * STRIP_ROWS;
              const uint32_t stage = smem_u32(stage0) + s * Cfg::A_BYTES;
              const int ch = cb * CB;
              const bool second = ch >= g.C0;
              const uint8_t* sbase = second ? g.src1 : g.src0;
              const uint32_t rowoff = second ? static_cast<uint32_t>(((img[sub] * g.Hin + vy) * g.Win) * g.C1)
                                             : static_cast<uint32_t>(((img[sub] * g.H0 + (g.up0 ? vy >> 1 : vy)) * g.W0) * g.C0);
              const uint32_t* xcol = xo + (second ? STRIP_ROWS : 0);
              const uint32_t cc = static_cast<uint32_t>(second ? ch - g.C0 : ch) + c * 8;
#pragma unroll
              for (int i = 0; i < (STRIP_PIX + ROWS_PER_PASS - 1) / ROWS_PER_PASS; ++i) {
                const int r = r0 + ROWS_PER_PASS * i;
                if (r < STRIP_PIX) {
                  const uint32_t xoff = xcol[r];
                  const bool ok = row_ok && (xoff != 0xFFFFFFFFu);
                  const uint32_t off = ok ? rowoff + xoff + cc : 0u;
                  const int sw = (CB == 64) ? (c ^ (r & 7)) : (c ^ ((r >> 1) & 3));
                  cp_async_16_zfill(stage + r * ROW_BYTES + (sw << 4), sbase + static_cast<size_t>(off) * 2, ok);
                }
              }
            }
            cp_async_mbar_arrive_noinc(&full_bar[s]);
            if (++s == STAGES) { s = 0; ph ^= 1; }
          }
        }
      }
    }
  } else if (warp < RS_EPI_FIRST_WARP) {
    // =============================== MMA ISSUERS (one warp per TMEM half) ===============================
    const int half = warp - RS_MMA_WARP;
    if (lane == 0) {
      const uint32_t idesc1 = umma_idesc_f16(BM, BN, p.fmt, p.w_fmt, 0, 0);
      const uint32_t idesc2 = umma_idesc_f16(BM, 2 * BN, p.fmt, p.w_fmt, 0, 0);
      const uint32_t idesc3 = umma_idesc_f16(BM, 3 * BN, p.fmt, p.w_fmt, 0, 0);
      constexpr uint32_t DESC_HI = (SBO >> 4) | (1u << 14) | (static_cast<uint32_t>(LAYOUT) << 29);
      auto mk_desc = [](uint32_t lo) -> uint64_t { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; };
      const uint32_t stage0_lo = ((smem_u32(stage0) >> 4) & 0x3FFF) | (1u << 16);
      const uint32_t wres_lo = ((smem_u32(w_res) >> 4) & 0x3FFF) | (1u << 16);
      constexpr uint32_t BT = Cfg::B_TILE >> 4;
      int s = half;                                          // this warp's stages: half, half + 2, ...
      uint32_t ph = 0, pit = 0;
      mbar_wait(w_bar, 0, 43);
      for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x, ++pit) {
        const bool ex = 2 * pair + half < rp.total_items;
        const uint32_t use_par = pit & 1;                    // parity of this half's use count
        for (int j = 0; j < H + 2; ++j) {
          const int dy_lo = j > H - 1 ? j - (H - 1) : 0;
          const int dy_hi = j < 2 ? j : 2;
          const int ndy = dy_hi - dy_lo + 1;
          const int q_hi = (H - 1) - (j - dy_lo);          // slot of the dy_lo block's output row
          const uint32_t d_col = tmem_base + (half * H + q_hi) * BN;
          if (ex && dy_lo == 0) {                          // first touch of this row slot in this item
            mbar_wait(&row_empty[half * H + q_hi], use_par ^ 1, 44);
            tcgen05_fence_after();
          }
          const uint32_t idesc_n = ndy == 3 ? idesc3 : (ndy == 2 ? idesc2 : idesc1);
          const uint32_t idesc_rest = ndy == 3 ? idesc2 : idesc1;   // the dy >= 1 blocks when dy = 0 is split off
          for (int cb = 0; cb < rp.cblocks; ++cb) {
            mbar_wait(&full_bar[s], ph, 45);
            tcgen05_fence_after();
            if (ex && !(kTriage && (p.debug & 2))) {
              const uint32_t a_lo = stage0_lo + static_cast<uint32_t>(s) * (Cfg::A_BYTES >> 4);
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
                const uint32_t b_lo = wres_lo + ((dx * rp.cblocks + cb) * 3 + dy_lo) * BT;
#pragma unroll
                for (int k = 0; k < CB / 16; ++k) {
                  const uint64_t a_desc = mk_desc(a_lo + dx * (ROW_BYTES >> 4) + 2 * k);
                  if (dy_lo == 0 && cb == 0 && dx == 0 && k == 0) {
                    umma_f16(d_col, a_desc, mk_desc(b_lo), idesc1, 0u);                       // opens the row's accumulator
                    if (ndy > 1) umma_f16(d_col + BN, a_desc, mk_desc(b_lo + BT), idesc_rest, 1u);
                  } else {
                    umma_f16(d_col, a_desc, mk_desc(b_lo + 2 * k), idesc_n, 1u);
                  }
                }
              }
            }
            umma_commit(&empty_bar[s]);
            s += 2;
            if (s >= STAGES) { s -= STAGES; ph ^= 1; }
          }
          if (ex && j >= 2) umma_commit(&row_full[half * H + (H - 1) - (j - 2)]);   // output row yc + j - 2 is complete
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE ===============================
    // four groups of four warps (one warp per TMEM lane quadrant): group grp drains the rows r = (grp >> 1) (mod 2) of the
    // pair's item hf = grp & 1
    const int q = warp & 3;
    const int grp = (warp - RS_EPI_FIRST_WARP) >> 2;
    const int hf = grp & 1;
    const int row = q * 32 + lane;
    const int e = tid - RS_EPI_FIRST_WARP * 32;
    for (int j = e; j < BN; j += RS_EPI_THREADS) {
      s_scale[j] = (p.scale != nullptr && j < p.N) ? p.scale[j] : 1.f;
      s_shift[j] = (p.shift != nullptr && j < p.N) ? p.shift[j] : 0.f;
    }
    named_bar_sync(1, RS_EPI_THREADS);
    uint32_t pit = 0;
    for (int pair = blockIdx.x; pair < total_pairs; pair += gridDim.x, ++pit) {
      const int item = 2 * pair + hf;
      if (item >= rp.total_items) continue;
      int img, x0, yc;
      decode(item, img, x0, yc);
      const uint32_t use_par = pit & 1;
      for (int r = grp >> 1; r < H; r += 2) {
        const int slot = hf * H + (H - 1 - r);
        const int y = yc + r;
        mbar_wait<32>(&row_full[slot], use_par, 46);
        tcgen05_fence_after();
        const int m = (img * g.Ho + y) * g.Wo + x0 + row;
        const bool m_ok = (x0 + row) < g.Wo && y < g.Ho && !(kTriage && (p.debug & 4));
        const uint32_t tbase = tmem_base + slot * BN + (static_cast<uint32_t>(q * 32) << 16);
#define FP_EPI(A, F) epilogue_cols<A, F, BN>(p, tbase, 0, 0, m, m_ok, s_scale, s_shift, &row_empty[slot], lane)
        if (p.fmt == FP_BF16) {
          switch (p.act) {
            case FP_ACT_RELU: FP_EPI(FP_ACT_RELU, FP_BF16); break;
            case FP_ACT_ELU: FP_EPI(FP_ACT_ELU, FP_BF16); break;
            case FP_ACT_SIGMOID: FP_EPI(FP_ACT_SIGMOID, FP_BF16); break;
            default: FP_EPI(FP_ACT_NONE, FP_BF16); break;
          }
        } else {
          switch (p.act) {
            case FP_ACT_RELU: FP_EPI(FP_ACT_RELU, FP_F16); break;
            case FP_ACT_ELU: FP_EPI(FP_ACT_ELU, FP_F16); break;
            case FP_ACT_SIGMOID: FP_EPI(FP_ACT_SIGMOID, FP_F16); break;
            default: FP_EPI(FP_ACT_NONE, FP_F16); break;
          }
        }
#undef FP_EPI
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == RS_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


template <int BN, int CB>
static int launch_strip(const StripParams& sp_in, const CUtensorMap& tmap, const CUtensorMap& tmap_x, cudaStream_t stream) {
  using Cfg = StripCfg<BN, CB>;
  auto kern = conv_igemm_strip_kernel<BN, CB>;
  { const int arc = ensure_max_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_MAX, "kern"); if (arc != FP_OK) return arc; }
  StripParams sp = sp_in;
  // resident filter if it fits beside >= 3 strip stages (single channel tile only), else stream it per stage
  const int w_bytes = 9 * sp.cblocks * Cfg::B_TILE;
  const int room = Cfg::SMEM_MAX - Cfg::FIXED_BYTES - w_bytes;
  if (sp.b.n_tiles == 1 && room >= 3 * Cfg::A_BYTES && opts().strip_wres != 0) {
    sp.w_resident = 1;
    sp.w_bytes = w_bytes;
    sp.stage_bytes = Cfg::A_BYTES;
    sp.stages = room / Cfg::A_BYTES < Cfg::MAX_STAGES ? room / Cfg::A_BYTES : Cfg::MAX_STAGES;
  } else {
    sp.w_resident = 0;
    sp.w_bytes = 0;
    sp.stage_bytes = Cfg::STAGE_BYTES;
    sp.stages = Cfg::STAGES;
  }
  const int smem_bytes = Cfg::FIXED_BYTES + sp.stages * sp.stage_bytes + sp.w_bytes;
  const int grid = sp.b.total_tiles < num_sms() ? sp.b.total_tiles : num_sms();
  kern<<<grid, NUM_THREADS, smem_bytes, stream>>>(sp, tmap, tmap_x);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "conv_igemm_strip launch (BN=%d): %s", BN, cudaGetErrorString(e));
  return FP_OK;
}

// Row-streaming launch (experimental, FP_ROWSTREAM=1): returns FP_ERR_UNSUPPORTED when the filter + 3 strip stages do
// not fit, so that the caller falls back to the strip kernel.
template <int BN, int CB>
static int launch_rowstream(const StripParams& sp, const CUtensorMap& tmap, const CUtensorMap& tmap_x, cudaStream_t stream) {
  using Cfg = StripCfg<BN, CB>;
  constexpr int SLOTS = 512 / BN;
  constexpr int RS_FIXED = 1024 + (2 * Cfg::MAX_STAGES + 2 * SLOTS + 1) * 8 + 32 + 2 * BN * 4 + 2 * 4 * STRIP_ROWS * 4;
  RowStreamParams rp{};
  rp.b = sp.b;
  rp.tiles_per_row = sp.tiles_per_row;
  rp.cblocks = sp.cblocks;
  rp.x_via_tma = sp.x_via_tma;
  rp.w_bytes = 9 * sp.cblocks * Cfg::B_TILE;
  const int room = Cfg::SMEM_MAX - RS_FIXED - rp.w_bytes;
  if (sp.b.n_tiles != 1 || room < 4 * Cfg::A_BYTES) return FP_ERR_UNSUPPORTED;
  rp.stages = (room / Cfg::A_BYTES < Cfg::MAX_STAGES ? room / Cfg::A_BYTES : Cfg::MAX_STAGES) & ~1;   // even: the two MMA warps own alternate stages
  rp.H = SLOTS / 2;
  rp.chunks_per_col = (sp.b.g.Ho + rp.H - 1) / rp.H;
  rp.total_items = sp.b.g.M / sp.b.g.HoWo * sp.tiles_per_row * rp.chunks_per_col;
  auto kern = conv_rowstream_kernel<BN, CB>;
  { const int arc = ensure_max_smem(reinterpret_cast<const void*>(kern), Cfg::SMEM_MAX, "kern"); if (arc != FP_OK) return arc; }
  const int smem_bytes = RS_FIXED + rp.stages * Cfg::A_BYTES + rp.w_bytes;
  const int pairs = (rp.total_items + 1) / 2;
  const int grid = pairs < num_sms() ? pairs : num_sms();
  kern<<<grid, RS_THREADS, smem_bytes, stream>>>(rp, tmap, tmap_x);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "conv_rowstream launch (BN=%d): %s", BN, cudaGetErrorString(e));
  return FP_OK;
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return (1 << l) == v ? l : -1;
}

}  // namespace fp

using namespace fp;

int fp::fill_gather_geom(const FpConvDesc* d, const void* src0, const void* src1, const char* who, GatherGeom& g,
                         bool need_pow2) {
  const int cin = d->c0 + d->c1;
  const int cl2 = ilog2_exact(cin);
  if (cl2 < 3 && (need_pow2 || cin < 8)) return set_error(FP_ERR_ARG, "%s: c0+c1=%d must be a power of two >= 8", who, cin);
  if ((d->c0 & 7) || (d->c1 & 7)) return set_error(FP_ERR_ARG, "%s: channel counts must be multiples of 8", who);
  if (d->c1 > 0 && src1 == nullptr) return set_error(FP_ERR_ARG, "%s: src1 is null but c1=%d", who, d->c1);
  if (d->kpad % 64 || d->kpad < d->kh * d->kw * cin) return set_error(FP_ERR_ARG, "%s: bad kpad %d", who, d->kpad);
  if (d->dtype != FP_F16 && d->dtype != FP_BF16) return set_error(FP_ERR_ARG, "%s: bad dtype", who);
  if (d->up0 < 0 || d->up0 > 2 || (d->up0 == 2 && (d->c1 > 0 || d->pad_mode != FP_PAD_ZERO || d->stride != 1)))
    return set_error(FP_ERR_ARG, "%s: bad up0 mode", who);
  const int Hin = d->up0 ? 2 * d->h0 : d->h0, Win = d->up0 ? 2 * d->w0 : d->w0;
  if (d->pad_mode == FP_PAD_REFLECT && (d->pad_t > 1 || d->pad_l > 1 || Hin < 2 || Win < 2 || d->kh > 3 || d->kw > 3))
    return set_error(FP_ERR_UNSUPPORTED, "%s: reflection padding is implemented for pad 1, 3x3 only", who);
  if (d->kh > 4 || d->kw > 4 || d->kh < 1 || d->kw < 1) return set_error(FP_ERR_UNSUPPORTED, "%s: kernel size must be 1..4", who);
  if (d->batch <= 0 || static_cast<long long>(d->batch) * d->h0 * d->w0 * d->c0 >= (1ll << 31) ||
      static_cast<long long>(d->batch) * Hin * Win * d->c1 >= (1ll << 31) ||
      static_cast<long long>(d->batch) * d->ho * d->wo >= (1ll << 31))
    return set_error(FP_ERR_UNSUPPORTED, "%s: a tensor has >= 2^31 elements; split the batch", who);
  g.src0 = static_cast<const uint8_t*>(src0);
  g.src1 = static_cast<const uint8_t*>(src1);
  g.C0 = d->c0; g.C1 = d->c1; g.H0 = d->h0; g.W0 = d->w0; g.up0 = d->up0;
  g.Hin = Hin; g.Win = Win;
  g.cin_log2 = cl2; g.taps = d->kh * d->kw; g.KH = d->kh; g.KW = d->kw; g.stride = d->stride;
  g.pad_t = d->pad_t; g.pad_l = d->pad_l; g.pad_mode = d->pad_mode;
  g.Ho = d->ho; g.Wo = d->wo; g.HoWo = d->ho * d->wo; g.M = d->batch * d->ho * d->wo;
  return FP_OK;
}

// split == nullptr: plain 16-bit operands.  Otherwise the low halves of the sources / residual / output (fp_conv_fwd_split).
static int conv_fwd_impl(const FpConvDesc* d, const void* src0, const void* src1, const void* w_packed,
                         const float* scale, const float* shift, const void* residual, void* out, void* stream,
                         const FpSplitLow* split) {
  if (d == nullptr || src0 == nullptr || w_packed == nullptr || out == nullptr)
    return set_error(FP_ERR_ARG, "fp_conv_fwd: null argument");
  IgemmParams p{};
  int rc = fill_gather_geom(d, src0, src1, "fp_conv_fwd", p.g, split == nullptr);
  if (rc != FP_OK) return rc;
  if (split != nullptr) {
    if (split->src0_lo == nullptr || (d->c1 > 0 && split->src1_lo == nullptr) || (residual != nullptr && split->residual_lo == nullptr))
      return set_error(FP_ERR_ARG, "fp_conv_fwd_split: a low half is missing");
    if (d->kpad < 3 * d->kh * d->kw * (d->c0 + d->c1)) return set_error(FP_ERR_ARG, "fp_conv_fwd_split: kpad %d < 3*taps*cin", d->kpad);
    if (d->up0 == 2 || d->out_f32 && split->out_lo != nullptr) return set_error(FP_ERR_ARG, "fp_conv_fwd_split: unsupported mode");
    p.g.split = 1;
    p.g.src0_lo = static_cast<const uint8_t*>(split->src0_lo);
    p.g.src1_lo = static_cast<const uint8_t*>(split->src1_lo);
    p.residual_lo = static_cast<const uint8_t*>(split->residual_lo);
    p.out_lo = split->out_lo;
  }
  if (d->kpad > 8 * KTAB_ENTRIES) return set_error(FP_ERR_UNSUPPORTED, "fp_conv_fwd: kpad %d too large", d->kpad);
  if (d->npad % d->bn || d->npad < d->cout) return set_error(FP_ERR_ARG, "fp_conv_fwd: bad npad %d (bn %d, cout %d)", d->npad, d->bn, d->cout);
  rc = fp_check_device();
  if (rc != FP_OK) return rc;

  p.w_packed = static_cast<const uint8_t*>(w_packed);
  p.scale = scale; p.shift = shift;
  p.residual = static_cast<const uint8_t*>(residual);
  p.out = out;
  p.N = d->cout;
  p.num_kb = d->kpad / 64; p.kpad = d->kpad;
  p.m_tiles = (p.g.M + BM - 1) / BM; p.n_tiles = d->npad / d->bn; p.total_tiles = p.m_tiles * p.n_tiles;
  p.act = d->act; p.fmt = d->dtype; p.w_fmt = d->w_dtype_plus1 ? d->w_dtype_plus1 - 1 : d->dtype;
  p.out_f32 = d->out_f32; p.out_ld = d->out_ld; p.res_ld = d->res_ld;
  p.b_via_tma = d->b_via_tma;
  p.debug = opts().debug_flags;

  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (d->b_via_tma) {
    rc = make_tmap_2d_sw128(&tmap, w_packed, static_cast<uint64_t>(d->kpad), static_cast<uint64_t>(d->npad), 64u,
                            static_cast<uint32_t>(d->bn));
    if (rc != FP_OK) return rc;
  }
  // A operand by TMA tile::gather4 when every 128-byte K row is 64 channels of one tap of one source
  CUtensorMap tmap_a0, tmap_a1;
  memset(&tmap_a0, 0, sizeof(tmap_a0));
  memset(&tmap_a1, 0, sizeof(tmap_a1));
  const bool gather_ok = (d->c0 % 64 == 0) && (d->c1 % 64 == 0) && d->b_via_tma && split == nullptr;
  p.a_mode = (gather_ok && opts().a_mode == 1) ? 1 : 0;   // opt-in (FP_A_MODE=1): measured no faster than cp.async
  if (p.a_mode == 1) {
    p.rows0 = d->batch * d->h0 * d->w0;
    p.rows1 = d->batch * p.g.Hin * p.g.Win;
    rc = make_tmap_2d_sw128(&tmap_a0, src0, static_cast<uint64_t>(d->c0), static_cast<uint64_t>(p.rows0), 64u, 1u);
    if (rc != FP_OK) return rc;
    if (d->c1 > 0) {
      rc = make_tmap_2d_sw128(&tmap_a1, src1, static_cast<uint64_t>(d->c1), static_cast<uint64_t>(p.rows1), 64u, 1u);
      if (rc != FP_OK) return rc;
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // row-strip kernel: 3x3 stride-1, 64-channel blocks, wide rows (>= 80 % of the 128-pixel row segments are used)
  if (split == nullptr) {
    const int segs = (d->wo + BM - 1) / BM;
    const int cb = (d->c0 % 64 == 0 && d->c1 % 64 == 0) ? 64 : ((d->c0 % 32 == 0 && d->c1 % 32 == 0) ? 32 : 0);
    const bool strip_ok = d->kh == 3 && d->kw == 3 && d->stride == 1 && d->up0 != 2 && cb != 0 && d->b_via_tma &&
                          d->bn <= 128 && d->bn >= 16 && (cb == 64 || d->bn <= 64) && d->kpad >= 9 * (d->c0 + d->c1) &&
                          opts().strip != 0;
    // used fraction of the 128-pixel row segments: the strip kernel (3 fills per tile instead of 9) pays off from ~80 %;
    // the row-streaming kernel (<= 1.5 fills per output row, N = 3*Cout instructions) already from ~60 % -- that brings
    // the 48x160 / 24x80 layers (Wo = 160 / 80: 62.5 %) onto it
    const int eff = d->wo * 100 / (segs * BM);
    const bool rs_ok = strip_ok && opts().rowstream != 0 && d->bn <= 64 && d->npad == d->bn && eff >= opts().rowstream_mineff;
    if (strip_ok && (rs_ok || (eff >= opts().strip_mineff && d->bn >= 32))) {      // bn = 16: row-streaming kernel only
      StripParams sp{};
      sp.b = p;
      sp.tiles_per_row = segs;
      sp.cblocks = (d->c0 + d->c1) / cb;
      sp.b.m_tiles = d->batch * d->ho * segs;
      sp.b.total_tiles = sp.b.m_tiles * sp.b.n_tiles;
      CUtensorMap tmap_s = tmap;
      if (cb == 32) {     // 64-byte rows: weights tiles [bn][32] with SWIZZLE_64B
        rc = make_tmap_2d(&tmap_s, w_packed, static_cast<uint64_t>(d->kpad), static_cast<uint64_t>(d->npad), 32u,
                          static_cast<uint32_t>(d->bn), 64);
        if (rc != FP_OK) return rc;
      }
      // zero-padded single-source inputs (every data-gradient full correlation): strips straight from the copy engine
      CUtensorMap tmap_x;
      memset(&tmap_x, 0, sizeof(tmap_x));
      if (d->pad_mode == FP_PAD_ZERO && d->up0 == 0 && d->c1 == 0 && opts().strip_tma != 0) {
        rc = make_tmap_nhwc(&tmap_x, src0, static_cast<uint64_t>(d->c0), static_cast<uint64_t>(d->w0),
                            static_cast<uint64_t>(d->h0), static_cast<uint64_t>(d->batch), static_cast<uint32_t>(cb),
                            static_cast<uint32_t>(STRIP_PIX), cb * 2);
        if (rc != FP_OK) return rc;
        sp.x_via_tma = 1;
      }
      {   // row-streaming kernel (see conv_rowstream_kernel): Cout tile <= 64, single channel tile
        if (rs_ok) {
          int rrc = FP_ERR_UNSUPPORTED;
          if (cb == 64 && d->bn == 16) rrc = launch_rowstream<16, 64>(sp, tmap_s, tmap_x, st);
          else if (cb == 32 && d->bn == 16) rrc = launch_rowstream<16, 32>(sp, tmap_s, tmap_x, st);
          else if (cb == 64 && d->bn == 32) rrc = launch_rowstream<32, 64>(sp, tmap_s, tmap_x, st);
          else if (cb == 64 && d->bn == 64) rrc = launch_rowstream<64, 64>(sp, tmap_s, tmap_x, st);
          else if (cb == 32 && d->bn == 32) rrc = launch_rowstream<32, 32>(sp, tmap_s, tmap_x, st);
          else if (cb == 32 && d->bn == 64) rrc = launch_rowstream<64, 32>(sp, tmap_s, tmap_x, st);
          if (rrc != FP_ERR_UNSUPPORTED) return rrc;
        }
        if (eff < opts().strip_mineff || d->bn < 32) goto general_kernel;     // too ragged / too narrow for the plain strip kernel
      }
      if (cb == 64) {
        switch (d->bn) {
          case 32: return launch_strip<32, 64>(sp, tmap_s, tmap_x, st);
          case 64: return launch_strip<64, 64>(sp, tmap_s, tmap_x, st);
          case 128: return launch_strip<128, 64>(sp, tmap_s, tmap_x, st);
          default: break;
        }
      } else {
        switch (d->bn) {
          case 32: return launch_strip<32, 32>(sp, tmap_s, tmap_x, st);
          case 64: return launch_strip<64, 32>(sp, tmap_s, tmap_x, st);
          default: break;
        }
      }
    }
  }
general_kernel:
  switch (d->bn) {
    case 16: return launch_igemm<16>(p, tmap, tmap_a0, tmap_a1, st);
    case 32: return launch_igemm<32>(p, tmap, tmap_a0, tmap_a1, st);
    case 64: return launch_igemm<64>(p, tmap, tmap_a0, tmap_a1, st);
    case 128: return launch_igemm<128>(p, tmap, tmap_a0, tmap_a1, st);
    case 256: return launch_igemm<256>(p, tmap, tmap_a0, tmap_a1, st);
    default: return set_error(FP_ERR_ARG, "fp_conv_fwd: bn=%d not in {16,32,64,128,256}", d->bn);
  }
}

extern "C" int fp_conv_fwd(const FpConvDesc* d, const void* src0, const void* src1, const void* w_packed,
                           const float* scale, const float* shift, const void* residual, void* out, void* stream) {
  return conv_fwd_impl(d, src0, src1, w_packed, scale, shift, residual, out, stream, nullptr);
}

extern "C" int fp_conv_fwd_split(const FpConvDesc* d, const void* src0_hi, const void* src1_hi, const FpSplitLow* low,
                                 const void* w_packed3, const float* scale, const float* shift, const void* residual_hi,
                                 void* out_hi, void* stream) {
  if (low == nullptr) return set_error(FP_ERR_ARG, "fp_conv_fwd_split: null argument");
  return conv_fwd_impl(d, src0_hi, src1_hi, w_packed3, scale, shift, residual_hi, out_hi, stream, low);
}

// perf triage helper: copies CTA 0's clock stamps (FP_DEBUG_FLAGS & 16) to the host; out[3][2048]
extern "C" int fp_debug_read_trace(long long* out) {
  cudaError_t e = cudaMemcpyFromSymbol(out, fp::g_trace, sizeof(long long) * 3 * 2048);
  return e == cudaSuccess ? FP_OK : fp::set_error(FP_ERR_CUDA, "fp_debug_read_trace: %s", cudaGetErrorString(e));
}
