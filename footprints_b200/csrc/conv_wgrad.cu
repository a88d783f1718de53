// Weight-gradient convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[co][k] = sum over output pixels m of dZ[m][co] * A[m][k],   k = tap*cin + c  (the packed-weight order)
//
// where A is the same on-the-fly im2col gather as the forward convolution (zero / reflection padding, stride,
// nearest x2 upsample of source 0, channel concat with source 1).  As a GEMM the reduction runs over pixels, so
// both operands are "MN-major": a shared-memory tile [128 pixels][64 channels] (128-byte rows, 128B swizzle) is
// consumed by UMMA with the pixel index as K (8 x K16 steps per tile).  One CTA owns a 128(co) x (up to 512)(k) block
// of dW — all of TMEM — for a contiguous range of pixel blocks (split-K), so each dZ tile and each pixel block's
// gather table is fetched/computed once per four 128-column chunks; fp32 partials go to a workspace and a second kernel sums
// the splits in a fixed order (deterministic) and scatters into the OIHW fp32 gradient.
//
//   warps 0-7  producers : dZ tile (2 x [128 px][64 co]) by TMA, im2col tile (2 x [128 px][64 k]) by cp.async gather;
//                          asynchronous mbarrier arrive (never blocks)
//   warp  8    MMA issuer: 8 x tcgen05.mma (M128 x N128 x K16, both operands MN-major) per im2col stage
//   warps 9-12 epilogue  : once, at the end: TMEM -> fp32 partial block
#include "fp_ptx.cuh"
#include "fp_internal.h"
#include "igemm_common.cuh"

namespace fp {

constexpr int WG_SUBTILE = 128 * 64 * 2;               // 16 KB: [128 pixels][64 channels] 16-bit
constexpr int WG_X_STAGES = 4;                         // im2col ring: 2 sub-tiles (128 K columns) per stage
constexpr int WG_X_STAGE_BYTES = 2 * WG_SUBTILE;
constexpr int WG_DZ_STAGES = 2;                        // dZ ring: 2 sub-tiles (128 output channels) per stage
constexpr int WG_DZ_STAGE_BYTES = 2 * WG_SUBTILE;
constexpr int WG_MAX_CHUNKS = 4;                       // 128-column chunks accumulated per CTA (4 x 128 = 512 TMEM columns)
constexpr int WG_PRODUCER_WARPS = 8;
constexpr int WG_PRODUCER_THREADS = WG_PRODUCER_WARPS * 32;
constexpr int WG_MMA_WARP = WG_PRODUCER_WARPS;
constexpr int WG_EPI_FIRST_WARP = WG_MMA_WARP + 1;
constexpr int WG_THREADS = (WG_EPI_FIRST_WARP + 4) * 32;   // 416
constexpr int WG_KTAB = 16 * WG_MAX_CHUNKS;            // 16-byte K chunks per CTA
constexpr int WG_BAR_BYTES = 128;                      // 13 mbarriers, padded so that what follows stays 16-byte aligned
constexpr int WG_SMEM_BYTES = 1024 + WG_X_STAGES * WG_X_STAGE_BYTES + WG_DZ_STAGES * WG_DZ_STAGE_BYTES + WG_BAR_BYTES + 16 +
                              2 * 128 * 16 + WG_KTAB * 4;

struct WgradParams {
  GatherGeom g;              // forward-input gather geometry
  int dz_ld, Cout, kpad;
  int k_chunks;              // 128-column chunks of dW
  int chunks_per_group, k_groups;   // a CTA accumulates `chunks_per_group` consecutive chunks
  int co_tiles, splits, blocks_per_split, num_pix_blocks;
  float* partial;            // [splits][co_tiles*128][kpad]
  int a_fmt, b_fmt;          // dZ format, activation format
};

__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_kernel(const __grid_constant__ WgradParams p, const __grid_constant__ CUtensorMap tmap_dz) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* x_tiles = smem;
  uint8_t* dz_tiles = smem + WG_X_STAGES * WG_X_STAGE_BYTES;
  uint64_t* x_full = reinterpret_cast<uint64_t*>(dz_tiles + WG_DZ_STAGES * WG_DZ_STAGE_BYTES);
  uint64_t* x_empty = x_full + WG_X_STAGES;
  uint64_t* dz_full = x_empty + WG_X_STAGES;
  uint64_t* dz_empty = dz_full + WG_DZ_STAGES;
  uint64_t* done_bar = dz_empty + WG_DZ_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(x_full) + WG_BAR_BYTES);
  uint4* s_rowinfo = reinterpret_cast<uint4*>(tmem_ptr_smem + 4);     // [2][128]
  uint32_t* s_ktab = reinterpret_cast<uint32_t*>(s_rowinfo + 2 * 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int unit = blockIdx.x;
  const int k_group = unit % p.k_groups; unit /= p.k_groups;
  const int co_tile = unit % p.co_tiles;
  const int split = unit / p.co_tiles;
  const int chunk0 = k_group * p.chunks_per_group;
  const int nchunks = min(p.chunks_per_group, p.k_chunks - chunk0);
  const int pb_begin = split * p.blocks_per_split;
  const int pb_end = min(p.num_pix_blocks, pb_begin + p.blocks_per_split);
  const int nblk = pb_end - pb_begin;

  if (tid == 0) {
    for (int s = 0; s < WG_X_STAGES; ++s) {
      mbar_init(&x_full[s], WG_PRODUCER_THREADS);    // one async arrive per im2col gather thread
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < WG_DZ_STAGES; ++s) {
      mbar_init(&dz_full[s], 1);                     // expect_tx arrive; the data is TMA transaction bytes
      mbar_init(&dz_empty[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_dz);
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  if (tid < WG_KTAB) s_ktab[tid] = make_ktab_entry(p.g, chunk0 * 16 + tid, p.kpad / 8);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < WG_PRODUCER_WARPS) {
    // =============================== PRODUCERS ===============================
    const int c = tid & 7, r0 = tid >> 3;
    const uint32_t swz = static_cast<uint32_t>(r0 * 128 + ((c ^ (r0 & 7)) << 4));
    const uint32_t x_u32 = smem_u32(x_tiles), dz_u32 = smem_u32(dz_tiles);
    uint32_t xs = 0;                                  // running im2col stage counter
    for (int i = 0; i < nblk; ++i) {
      const int pb = pb_begin + i;
      // row info of this pixel block: computed once, used for every K chunk of the group
      uint4* info_buf = s_rowinfo + (i & 1) * 128;
      if (tid < 128) {
        const int m = pb * 128 + tid;
        uint4 ri = make_uint4(0u, 0xFFFFFFFFu, 0u, 0xFFFFFFFFu);
        if (m < p.g.M) ri = make_row_info_for_pixel(p.g, m);
        info_buf[tid] = ri;
      }
      named_bar_sync(2, WG_PRODUCER_THREADS);
      // dZ tile: plain rows of the [M][dz_ld] matrix -> two TMA boxes of [128 pixels][64 channels]; columns or rows
      // past the matrix are zero-filled by the TMA unit (Cout < 128, last pixel block)
      if (tid == 0) {
        const int ds = i % WG_DZ_STAGES;
        mbar_wait(&dz_empty[ds], ((i / WG_DZ_STAGES) & 1) ^ 1, 10);
        mbar_arrive_expect_tx(&dz_full[ds], WG_DZ_STAGE_BYTES);
        tma_load_2d(dz_u32 + ds * WG_DZ_STAGE_BYTES, &tmap_dz, &dz_full[ds], co_tile * 128, pb * 128);
        tma_load_2d(dz_u32 + ds * WG_DZ_STAGE_BYTES + WG_SUBTILE, &tmap_dz, &dz_full[ds], co_tile * 128 + 64, pb * 128);
      }
      uint4 info[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) info[r] = info_buf[r0 + 32 * r];
      for (int j = 0; j < nchunks; ++j, ++xs) {
        const int s = xs % WG_X_STAGES;
        if (lane == 0) mbar_wait(&x_empty[s], ((xs / WG_X_STAGES) & 1) ^ 1, 11);
        __syncwarp();
        const uint32_t stage_base = x_u32 + s * WG_X_STAGE_BYTES + swz;
        // im2col sub-tiles: K chunks (chunk0 + j)*16 + h*8 + c
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const KDecode kd = decode_ktab(p.g, s_ktab[j * 16 + h * 8 + c]);
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            bool ok;
            const uint32_t off = gather_offset(kd, info[r], ok);
            cp_async_16_zfill(stage_base + h * WG_SUBTILE + r * 4096, kd.sbase + static_cast<size_t>(off) * 2, ok);
          }
        }
        cp_async_mbar_arrive_noinc(&x_full[s]);
      }
    }
  } else if (warp == WG_MMA_WARP) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, 128, p.a_fmt, p.b_fmt, 1, 1);
      uint32_t xs = 0;
      for (int i = 0; i < nblk; ++i) {
        const int ds = i % WG_DZ_STAGES;
        mbar_wait(&dz_full[ds], (i / WG_DZ_STAGES) & 1, 12);
        const uint32_t a_base = smem_u32(dz_tiles) + ds * WG_DZ_STAGE_BYTES;
        for (int j = 0; j < nchunks; ++j, ++xs) {
          const int s = xs % WG_X_STAGES;
          mbar_wait(&x_full[s], (xs / WG_X_STAGES) & 1, 13);
          tcgen05_fence_after();
          const uint32_t b_base = smem_u32(x_tiles) + s * WG_X_STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) {
            // MN-major SW128: 16 pixels (K) = 2 groups of 8 rows, SBO = 1024 B; 64-channel blocks LBO = 16 KB apart
            const uint64_t a_desc = umma_smem_desc_sw128(a_base + ks * 2048, WG_SUBTILE, 1024);
            const uint64_t b_desc = umma_smem_desc_sw128(b_base + ks * 2048, WG_SUBTILE, 1024);
            umma_f16(tmem_base + j * 128, a_desc, b_desc, idesc, (i | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&x_empty[s]);
        }
        umma_commit(&dz_empty[ds]);
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE (once) ===============================
    // thread <-> accumulator row = output channel co_tile*128 + row
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait<256>(done_bar, 0, 14);
    tcgen05_fence_after();
    float* dst_row = p.partial + (static_cast<size_t>(split) * p.co_tiles * 128 + co_tile * 128 + row) * p.kpad;
    for (int j = 0; j < nchunks; ++j) {
      float* dst = dst_row + (chunk0 + j) * 128;
      const int ncols = min(128, p.kpad - (chunk0 + j) * 128);
#pragma unroll 1
      for (int cidx = 0; cidx < 128; cidx += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(tmem_base + j * 128 + cidx + (static_cast<uint32_t>(q * 32) << 16), v);
        tmem_ld_wait();
        if (cidx < ncols) {
#pragma unroll
          for (int x = 0; x < 4; ++x)
            *reinterpret_cast<float4*>(dst + cidx + 4 * x) =
                nblk > 0 ? make_float4(__uint_as_float(v[4 * x]), __uint_as_float(v[4 * x + 1]),
                                       __uint_as_float(v[4 * x + 2]), __uint_as_float(v[4 * x + 3]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW_oihw[co][ci][ky][kx] (+)= sum over splits of partial[split][co][(ky*kw+kx)*cin + ci]; stem = 1 inverts the
// space-to-depth packing of fp_pack_stem_weight.
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int splits, int co_rows, int kpad, int cout,
                                    int cin, int cpad, int kh, int kw, int stem, int accumulate, float* __restrict__ dw) {
  // one thread per (co, k) of the packed layout: consecutive threads read consecutive k (coalesced) in every split
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<long long>(cout) * kpad) return;
  const int k = static_cast<int>(idx % kpad);
  const int co = static_cast<int>(idx / kpad);
  int ci, ky, kx;
  if (stem) {
    const int tap = k >> 4, ch = k & 15;
    const int bh = tap >> 2, bw = tap & 3, sub = ch >> 2;
    ci = ch & 3;
    ky = 2 * bh + (sub >> 1) - 1;
    kx = 2 * bw + (sub & 1) - 1;
    if (ci >= 3 || ky < 0 || ky >= 7 || kx < 0 || kx >= 7 || tap >= 16) return;
  } else {
    const int tap = k / cpad;
    ci = k - tap * cpad;
    if (tap >= kh * kw || ci >= cin) return;
    ky = tap / kw;
    kx = tap - ky * kw;
  }
  const size_t stride = static_cast<size_t>(co_rows) * kpad;
  const float* src = partial + static_cast<size_t>(co) * kpad + k;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;        // four independent chains, fixed order
  int sp = 0;
  for (; sp + 4 <= splits; sp += 4) {
    s0 += src[(sp + 0) * stride]; s1 += src[(sp + 1) * stride]; s2 += src[(sp + 2) * stride]; s3 += src[(sp + 3) * stride];
  }
  for (; sp < splits; ++sp) s0 += src[sp * stride];
  const double s = (s0 + s1) + (s2 + s3);
  const size_t o = ((static_cast<size_t>(co) * cin + ci) * kh + ky) * kw + kx;
  dw[o] = (accumulate ? dw[o] : 0.f) + static_cast<float>(s);
}

struct WgradPlan { int k_chunks, chunks_per_group, k_groups, co_tiles, splits, blocks_per_split, num_pix_blocks; };

static WgradPlan wgrad_plan(const FpConvDesc* d) {
  WgradPlan w;
  const long long M = static_cast<long long>(d->batch) * d->ho * d->wo;
  w.num_pix_blocks = static_cast<int>((M + 127) / 128);
  w.k_chunks = (d->kpad + 127) / 128;
  w.k_groups = (w.k_chunks + WG_MAX_CHUNKS - 1) / WG_MAX_CHUNKS;
  w.chunks_per_group = (w.k_chunks + w.k_groups - 1) / w.k_groups;      // balanced groups of <= 4 chunks
  w.k_groups = (w.k_chunks + w.chunks_per_group - 1) / w.chunks_per_group;
  w.co_tiles = (d->cout + 127) / 128;
  const int base = w.k_groups * w.co_tiles;
  int want = (2 * num_sms() + base - 1) / base;
  if (want < 1) want = 1;
  if (want > w.num_pix_blocks) want = w.num_pix_blocks;
  w.blocks_per_split = (w.num_pix_blocks + want - 1) / want;
  w.splits = (w.num_pix_blocks + w.blocks_per_split - 1) / w.blocks_per_split;
  return w;
}

}  // namespace fp

using namespace fp;

extern "C" long long fp_conv_wgrad_workspace_bytes(const FpConvDesc* d) {
  if (d == nullptr) return -1;
  const WgradPlan w = wgrad_plan(d);
  return static_cast<long long>(w.splits) * w.co_tiles * 128 * d->kpad * 4;
}

extern "C" int fp_conv_wgrad(const FpConvDesc* d, const void* src0, const void* src1, const void* dz, int dz_ld,
                             int grad_dtype, int stem, int accumulate, float* dw_oihw, float* workspace, void* stream) {
  if (!d || !src0 || !dz || !dw_oihw || !workspace) return set_error(FP_ERR_ARG, "fp_conv_wgrad: null argument");
  WgradParams p{};
  int rc = fill_gather_geom(d, src0, src1, "fp_conv_wgrad", p.g);
  if (rc != FP_OK) return rc;
  if (dz_ld < d->cout || (dz_ld & 7)) return set_error(FP_ERR_ARG, "fp_conv_wgrad: dz_ld=%d must be a multiple of 8 >= cout", dz_ld);
  rc = fp_check_device();
  if (rc != FP_OK) return rc;
  p.dz_ld = dz_ld; p.Cout = d->cout; p.kpad = d->kpad;
  const WgradPlan w = wgrad_plan(d);
  p.k_chunks = w.k_chunks; p.chunks_per_group = w.chunks_per_group; p.k_groups = w.k_groups;
  p.co_tiles = w.co_tiles; p.splits = w.splits; p.blocks_per_split = w.blocks_per_split; p.num_pix_blocks = w.num_pix_blocks;
  p.partial = workspace;
  p.a_fmt = grad_dtype; p.b_fmt = d->dtype;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
    if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "cudaFuncSetAttribute(wgrad): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = p.k_groups * p.co_tiles * p.splits;
  CUtensorMap tmap_dz;
  rc = make_tmap_2d_sw128(&tmap_dz, dz, static_cast<uint64_t>(dz_ld), static_cast<uint64_t>(p.g.M), 64u, 128u);
  if (rc != FP_OK) return rc;
  conv_wgrad_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, st>>>(p, tmap_dz);
  FP_CHECK_LAUNCH("fp_conv_wgrad");
  const int cin = stem ? 3 : d->c0 + d->c1;
  const int kh = stem ? 7 : d->kh, kw = stem ? 7 : d->kw;
  const long long total = static_cast<long long>(d->cout) * d->kpad;
  wgrad_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
      workspace, p.splits, p.co_tiles * 128, d->kpad, d->cout, cin, stem ? 16 : d->c0 + d->c1, kh, kw, stem, accumulate,
      dw_oihw);
  FP_CHECK_LAUNCH("fp_conv_wgrad(reduce)");
  return FP_OK;
}
