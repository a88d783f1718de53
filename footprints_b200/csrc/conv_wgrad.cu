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

// ==================================================================================================
// Row-strip variant (3x3 stride-1 convolutions on wide rows; channel blocks of CB = 64 / 32 / 16): a pixel block is a
// 128-pixel segment of one image row; for a filter row dy and a channel block the input row segment (130 pixels) is
// staged once and the three horizontal taps are three MN-major B operands starting 0, 1, 2 rows into the strip (the
// swizzle is a function of absolute shared-memory addresses, probe_umma_shift.cu).  Strip rows are CB*2 bytes
// (SWIZZLE_128B / 64B / 32B).  A CTA owns (128 output channels) x (ndy filter rows) x (ncb channel blocks) x 3 taps x CB
// TMEM columns for a range of row segments; dZ arrives by a 3-D TMA box whose out-of-row pixels are zero-filled.
// ==================================================================================================
// A row segment is `seg` output pixels (any multiple of 16 up to 160: K of the MMA is the pixel axis, so unlike the
// forward strips the segment need not be 128): Wo = 160 / 320 / 640 -> 160, Wo = 80 -> 80, Wo = 40 -> 48.
constexpr int WS_MAX_SEG = 160;
constexpr int WS_STRIP_ROWS = 168;                             // >= WS_MAX_SEG + 2, whole 8-row swizzle groups
constexpr int WS_DZ_SUB_MAX = WS_MAX_SEG * 128;                // one 64-channel dZ block of a segment
constexpr int WS_DZ_STAGE_MAX = 2 * WS_DZ_SUB_MAX;
constexpr int WS_X_STAGES = 6;
constexpr int WS_DZ_STAGES = 2;
constexpr int WS_MAX_COLS = 384;                               // accumulator columns per CTA
constexpr int WS_BAR_BYTES = 192;
constexpr int WS_XOFF_WORDS = 2 * WS_STRIP_ROWS + 8;

template <int CB> constexpr int ws_smem_bytes() {
  return 1024 + WS_X_STAGES * WS_STRIP_ROWS * CB * 2 + WS_DZ_STAGES * WS_DZ_STAGE_MAX + WS_BAR_BYTES + 16 +
         2 * WS_XOFF_WORDS * 4;
}

struct WgradStripParams {
  GatherGeom g;
  int Cout, kpad, cin;
  int seg;                    // pixels per row segment (multiple of 16, <= WS_MAX_SEG)
  int tiles_per_row;          // ceil(Wo / seg)
  int num_segments;           // batch * Ho * tiles_per_row
  int cblocks, cb_per_unit, cb_groups;
  int ndy, dy_groups;         // filter rows per CTA (3 or 1)
  int co_tiles, splits, segs_per_split;
  float* partial;
  int a_fmt, b_fmt;
  int fuse_taps;              // 1: one N = 3*CB instruction per K step (overlapping MN blocks)
};

template <int CB>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_strip_kernel(const __grid_constant__ WgradStripParams p, const __grid_constant__ CUtensorMap tmap_dz) {
  constexpr int ROW_BYTES = CB * 2;
  constexpr int STRIP_BYTES = WS_STRIP_ROWS * ROW_BYTES;        // whole swizzle periods (1024 / 512 / 256 B) for every CB
  constexpr int CHUNKS = CB / 8;                                // 16-byte chunks per strip row
  constexpr int ROWS_PER_PASS = WG_PRODUCER_THREADS / CHUNKS;
  constexpr int PASSES = (WS_MAX_SEG + 2 + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
  constexpr uint32_t LAYOUT = CB == 64 ? 2u : (CB == 32 ? 4u : 6u);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* dz_tiles = smem;
  uint8_t* x_tiles = smem + WS_DZ_STAGES * WS_DZ_STAGE_MAX;
  uint64_t* x_full = reinterpret_cast<uint64_t*>(x_tiles + WS_X_STAGES * STRIP_BYTES);
  uint64_t* x_empty = x_full + WS_X_STAGES;
  uint64_t* dz_full = x_empty + WS_X_STAGES;
  uint64_t* dz_empty = dz_full + WS_DZ_STAGES;
  uint64_t* done_bar = dz_empty + WS_DZ_STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(x_full) + WS_BAR_BYTES);
  uint32_t* s_xoff = tmem_ptr_smem + 4;                         // [2][WS_XOFF_WORDS]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // unit -> (cb group, dy group, co tile, split)
  int unit = blockIdx.x;
  const int cbg = unit % p.cb_groups; unit /= p.cb_groups;
  const int dy0 = (unit % p.dy_groups) * p.ndy; unit /= p.dy_groups;
  const int co_tile = unit % p.co_tiles;
  const int split = unit / p.co_tiles;
  const int cb0 = cbg * p.cb_per_unit;
  const int ncb = min(p.cb_per_unit, p.cblocks - cb0);
  const int ndy = p.ndy;
  const int seg_begin = split * p.segs_per_split;
  const int seg_end = min(p.num_segments, seg_begin + p.segs_per_split);
  const int nseg = max(seg_end - seg_begin, 0);
  const bool two_halves = p.Cout - co_tile * 128 > 64;          // second 64-channel dZ block carries data
  const int strip_pix = p.seg + 2;                              // input pixels feeding one segment of a 3-tap row
  const uint32_t dz_sub = static_cast<uint32_t>(p.seg) * 128u;  // bytes of one 64-channel dZ block (multiple of 1024)
  const int ksteps = p.seg >> 4;
  const GatherGeom& g = p.g;

  if (tid == 0) {
    for (int s = 0; s < WS_X_STAGES; ++s) {
      mbar_init(&x_full[s], WG_PRODUCER_THREADS);
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < WS_DZ_STAGES; ++s) {
      mbar_init(&dz_full[s], 1);
      mbar_init(&dz_empty[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_dz);
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < WG_PRODUCER_WARPS) {
    // =============================== PRODUCERS ===============================
    const int c = tid % CHUNKS, r0 = tid / CHUNKS;
    const uint32_t x_u32 = smem_u32(x_tiles), dz_u32 = smem_u32(dz_tiles);
    uint32_t xs = 0;
    for (int i = 0; i < nseg; ++i) {
      const int seg = seg_begin + i;
      const int xseg = seg % p.tiles_per_row;
      const int yrow = seg / p.tiles_per_row;            // b * Ho + y
      const int img = yrow / g.Ho;
      const int y = yrow - img * g.Ho;
      const int x0 = xseg * p.seg;
      uint32_t* xo = s_xoff + (i & 1) * WS_XOFF_WORDS;
      if (tid < WS_STRIP_ROWS) {
        int vx = x0 - g.pad_l + tid;
        bool ok = tid < strip_pix;
        if (g.pad_mode == FP_PAD_REFLECT) vx = vx < 0 ? -vx : (vx >= g.Win ? 2 * g.Win - 2 - vx : vx);
        ok = ok && static_cast<unsigned>(vx) < static_cast<unsigned>(g.Win);
        xo[tid] = ok ? static_cast<uint32_t>((g.up0 ? vx >> 1 : vx) * g.C0) : 0xFFFFFFFFu;
        xo[WS_STRIP_ROWS + tid] = ok ? static_cast<uint32_t>(vx * g.C1) : 0xFFFFFFFFu;
      } else if (tid < WS_STRIP_ROWS + 3) {
        const int dy = tid - WS_STRIP_ROWS;
        int vy = y - g.pad_t + dy;
        if (g.pad_mode == FP_PAD_REFLECT) vy = vy < 0 ? -vy : (vy >= g.Hin ? 2 * g.Hin - 2 - vy : vy);
        const bool ok = static_cast<unsigned>(vy) < static_cast<unsigned>(g.Hin);
        xo[2 * WS_STRIP_ROWS + dy] = ok ? static_cast<uint32_t>(((img * g.H0 + (g.up0 ? vy >> 1 : vy)) * g.W0) * g.C0) : 0xFFFFFFFFu;
        xo[2 * WS_STRIP_ROWS + 4 + dy] = ok ? static_cast<uint32_t>(((img * g.Hin + vy) * g.Win) * g.C1) : 0xFFFFFFFFu;
      }
      named_bar_sync(2, WG_PRODUCER_THREADS);
      if (tid == 0) {
        const int ds = i % WS_DZ_STAGES;
        mbar_wait(&dz_empty[ds], ((i / WS_DZ_STAGES) & 1) ^ 1, 30);
        mbar_arrive_expect_tx(&dz_full[ds], two_halves ? 2 * dz_sub : dz_sub);
        tma_load_3d(dz_u32 + ds * WS_DZ_STAGE_MAX, &tmap_dz, &dz_full[ds], co_tile * 128, x0, yrow);
        if (two_halves)
          tma_load_3d(dz_u32 + ds * WS_DZ_STAGE_MAX + dz_sub, &tmap_dz, &dz_full[ds], co_tile * 128 + 64, x0, yrow);
      }
      for (int dyi = 0; dyi < ndy; ++dyi)
        for (int j = 0; j < ncb; ++j, ++xs) {
          const int s = xs % WS_X_STAGES;
          mbar_wait(&x_empty[s], ((xs / WS_X_STAGES) & 1) ^ 1, 31);
          const uint32_t stage = x_u32 + s * STRIP_BYTES;
          const int ch = (cb0 + j) * CB;
          const bool second = ch >= g.C0;
          const uint8_t* sbase = second ? g.src1 : g.src0;
          const uint32_t rowoff = xo[2 * WS_STRIP_ROWS + (second ? 4 : 0) + dy0 + dyi];
          const uint32_t* xcol = xo + (second ? WS_STRIP_ROWS : 0);
          const uint32_t cc = static_cast<uint32_t>(second ? ch - g.C0 : ch) + c * 8;
#pragma unroll
          for (int k = 0; k < PASSES; ++k) {
            const int r = r0 + ROWS_PER_PASS * k;
            if (r < strip_pix) {
              const uint32_t xoff = xcol[r];
              const bool ok = (rowoff != 0xFFFFFFFFu) && (xoff != 0xFFFFFFFFu);
              const uint32_t off = ok ? rowoff + xoff + cc : 0u;
              const int sw = CB == 64 ? (r & 7) : (CB == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
              cp_async_16_zfill(stage + r * ROW_BYTES + ((c ^ sw) << 4), sbase + static_cast<size_t>(off) * 2, ok);
            }
          }
          cp_async_mbar_arrive_noinc(&x_full[s]);
        }
    }
  } else if (warp == WG_MMA_WARP) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_f16(128, CB, p.a_fmt, p.b_fmt, 1, 1);
      const uint32_t idesc3 = umma_idesc_f16(128, 3 * CB, p.a_fmt, p.b_fmt, 1, 1);
      uint32_t xs = 0;
      for (int i = 0; i < nseg; ++i) {
        const int ds = i % WS_DZ_STAGES;
        mbar_wait(&dz_full[ds], (i / WS_DZ_STAGES) & 1, 32);
        const uint32_t a_base = smem_u32(dz_tiles) + ds * WS_DZ_STAGE_MAX;
        for (int dyi = 0; dyi < ndy; ++dyi)
          for (int j = 0; j < ncb; ++j, ++xs) {
            const int s = xs % WS_X_STAGES;
            mbar_wait(&x_full[s], (xs / WS_X_STAGES) & 1, 33);
            tcgen05_fence_after();
            const uint32_t strip = smem_u32(x_tiles) + s * STRIP_BYTES;
            const uint32_t d_tmem = tmem_base + ((dyi * ncb + j) * 3) * CB;
            if (p.fuse_taps) {
              // one N = 3*CB instruction per K step: B's three MN blocks are the strip at pixel shifts 0, 1, 2 (block
              // stride LBO = one strip row), so dZ (the A operand) is read from shared memory once instead of 3 times
#pragma unroll 2
              for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t a_desc = umma_smem_desc_sw128(a_base + ks * 2048, dz_sub, 1024);
                const uint64_t b_desc = umma_smem_desc(strip + ks * 16 * ROW_BYTES, ROW_BYTES, 8 * ROW_BYTES, LAYOUT);
                umma_f16(d_tmem, a_desc, b_desc, idesc3, (i | ks) != 0 ? 1u : 0u);
              }
            } else {
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll 2
                for (int ks = 0; ks < ksteps; ++ks) {
                  const uint64_t a_desc = umma_smem_desc_sw128(a_base + ks * 2048, dz_sub, 1024);
                  // B = the strip advanced by dx pixels (rows); N = CB channels = one MN block (LBO unused)
                  const uint64_t b_desc = umma_smem_desc(strip + dx * ROW_BYTES + ks * 16 * ROW_BYTES, 8 * ROW_BYTES,
                                                         8 * ROW_BYTES, LAYOUT);
                  umma_f16(d_tmem + dx * CB, a_desc, b_desc, idesc, (i | ks) != 0 ? 1u : 0u);
                }
              }
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
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait<256>(done_bar, 0, 34);
    tcgen05_fence_after();
    const bool row_ok = co_tile * 128 + row < p.Cout;
    float* dst_row = p.partial + (static_cast<size_t>(split) * p.co_tiles * 128 + co_tile * 128 + row) * p.kpad;
    if (co_tile * 128 + q * 32 < p.Cout)                       // warp-uniform: this quarter holds real output channels
      for (int dyi = 0; dyi < ndy; ++dyi)
        for (int j = 0; j < ncb; ++j)
          for (int dx = 0; dx < 3; ++dx) {
            float* dst = dst_row + ((dy0 + dyi) * 3 + dx) * p.cin + (cb0 + j) * CB;
#pragma unroll 1
            for (int cidx = 0; cidx < CB; cidx += 16) {
              uint32_t v[16];
              tmem_ld_32x32b_x16(tmem_base + ((dyi * ncb + j) * 3 + dx) * CB + cidx + (static_cast<uint32_t>(q * 32) << 16), v);
              tmem_ld_wait();
              if (row_ok) {
#pragma unroll
                for (int x = 0; x < 4; ++x)
                  *reinterpret_cast<float4*>(dst + cidx + 4 * x) =
                      nseg > 0 ? make_float4(__uint_as_float(v[4 * x]), __uint_as_float(v[4 * x + 1]),
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

// ==================================================================================================
// Rolling, row-paired variant: the DEFAULT weight gradient for Cout <= 64 on strip-eligible layers (FP_WGRAD_ROLL=0 falls
// back to the strip kernel above).  Validated on B200 against the strip kernel and autograd
// (tests/test_gpu_kernel_variants.py); 58-61 % tensor-pipe active for 64-channel blocks (profiles/r02A_ncu_full_table.md).
//
// For Cout <= 64 the strip kernel above wastes half of every instruction's M = 128 and stages each input row and each dZ
// row three times (once per filter row).  Here a CTA walks DOWN a column of row segments of one image: every input row
// segment X[v] and every dZ row segment is staged once, and per K step (16 pixels) two instructions are issued:
//   pair  : A = [dZ[y-1] (64 ch) ; dZ[y] (64 ch)]  (two adjacent dZ ring slots, LBO = slot stride), B = X[v], 3 taps fused
//           -> TMEM lanes 0-63 += filter row dy = 1, lanes 64-127 += filter row dy = 0         (y = v + pad)
//   single: A = [dZ[y-2] ; (ignored)], same B  -> lanes 0-63 of a second accumulator block += filter row dy = 2
// The dZ ring has 4 slots plus a 5th that duplicates every row with y % 4 == 0, so "slot(y-1) + 1" always holds row y.
// Rows outside the image arrive as zeros (4-D TMA box, out-of-bounds fill), so no instruction is conditional.
// ==================================================================================================
constexpr int WR_DZ_SLOTS = 5;
constexpr int WR_BAR_BYTES = 256;

template <int CB> constexpr int wr_x_stages() { return CB == 64 ? 5 : 6; }
template <int CB> constexpr int wr_smem_bytes() {
  return 1024 + WR_DZ_SLOTS * WS_DZ_SUB_MAX + wr_x_stages<CB>() * WS_STRIP_ROWS * CB * 2 + WR_BAR_BYTES + 16 +
         2 * WS_XOFF_WORDS * 4;
}

struct WgradRollParams {
  GatherGeom g;
  int Cout, kpad, cin;
  int seg, tiles_per_row;
  int num_cols;               // batch * tiles_per_row columns of row segments
  int cblocks, cb_per_unit, cb_groups;
  int splits, cols_per_split;
  float* partial;             // [splits][64][kpad]
  int a_fmt, b_fmt;
};

template <int CB>
__global__ void __launch_bounds__(WG_THREADS, 1)
conv_wgrad_roll_kernel(const __grid_constant__ WgradRollParams p, const __grid_constant__ CUtensorMap tmap_dz) {
  constexpr int ROW_BYTES = CB * 2;
  constexpr int STRIP_BYTES = WS_STRIP_ROWS * ROW_BYTES;
  constexpr int CHUNKS = CB / 8;
  constexpr int ROWS_PER_PASS = WG_PRODUCER_THREADS / CHUNKS;
  constexpr int PASSES = (WS_MAX_SEG + 2 + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
  constexpr uint32_t LAYOUT = CB == 64 ? 2u : (CB == 32 ? 4u : 6u);
  constexpr int X_STAGES = wr_x_stages<CB>();
  constexpr uint32_t SINGLE_COL = 256;                          // accumulator block of filter row dy = 2
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* dz_ring = smem;                                      // [5][seg px][64 ch], slot stride WS_DZ_SUB_MAX
  uint8_t* x_tiles = smem + WR_DZ_SLOTS * WS_DZ_SUB_MAX;
  uint64_t* x_full = reinterpret_cast<uint64_t*>(x_tiles + X_STAGES * STRIP_BYTES);
  uint64_t* x_empty = x_full + X_STAGES;
  uint64_t* dz_full = x_empty + X_STAGES;
  uint64_t* dz_empty = dz_full + WR_DZ_SLOTS;
  uint64_t* done_bar = dz_empty + WR_DZ_SLOTS;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(x_full) + WR_BAR_BYTES);
  uint32_t* s_xoff = tmem_ptr_smem + 4;                         // [2][2 * WS_STRIP_ROWS] column maps of both sources

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cbg = blockIdx.x % p.cb_groups;
  const int split = blockIdx.x / p.cb_groups;
  const int cb0 = cbg * p.cb_per_unit;
  const int ncb = min(p.cb_per_unit, p.cblocks - cb0);
  const int col_begin = split * p.cols_per_split;
  const int col_end = min(p.num_cols, col_begin + p.cols_per_split);
  const int strip_pix = p.seg + 2;
  const uint32_t dz_bytes = static_cast<uint32_t>(p.seg) * 128u;
  const int ksteps = p.seg >> 4;
  const GatherGeom& g = p.g;

  if (tid == 0) {
    for (int s = 0; s < X_STAGES; ++s) {
      mbar_init(&x_full[s], WG_PRODUCER_THREADS);
      mbar_init(&x_empty[s], 1);
    }
    for (int s = 0; s < WR_DZ_SLOTS; ++s) {
      mbar_init(&dz_full[s], 1);
      mbar_init(&dz_empty[s], 1);
    }
    mbar_init(done_bar, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmap_dz);
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_ptr_smem, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < WG_PRODUCER_WARPS) {
    // =============================== PRODUCERS ===============================
    const int c = tid % CHUNKS, r0 = tid / CHUNKS;
    const uint32_t x_u32 = smem_u32(x_tiles), dz_u32 = smem_u32(dz_ring);
    uint32_t xs = 0, ci = 0;
    uint32_t empty_ph = 0;                                       // bit s: phase of dz_empty[s] this thread waits for next
    for (int col = col_begin; col < col_end; ++col, ++ci) {
      const int img = col / p.tiles_per_row;
      const int x0 = (col - img * p.tiles_per_row) * p.seg;
      uint32_t* xo = s_xoff + (ci & 1) * (2 * WS_STRIP_ROWS);
      if (tid < WS_STRIP_ROWS) {
        int vx = x0 - g.pad_l + tid;
        bool ok = tid < strip_pix;
        if (g.pad_mode == FP_PAD_REFLECT) vx = vx < 0 ? -vx : (vx >= g.Win ? 2 * g.Win - 2 - vx : vx);
        ok = ok && static_cast<unsigned>(vx) < static_cast<unsigned>(g.Win);
        xo[tid] = ok ? static_cast<uint32_t>((g.up0 ? vx >> 1 : vx) * g.C0) : 0xFFFFFFFFu;
        xo[WS_STRIP_ROWS + tid] = ok ? static_cast<uint32_t>(vx * g.C1) : 0xFFFFFFFFu;
      }
      named_bar_sync(2, WG_PRODUCER_THREADS);
      for (int y = -2; y < g.Ho + 2; ++y) {
        if (tid == 0) {
          // dZ row y (zeros outside the image) -> ring slot y mod 4, and a copy in slot 4 when y % 4 == 0
          const int slot = (y + 4) & 3;
          mbar_wait(&dz_empty[slot], ((empty_ph >> slot) & 1) ^ 1, 50);
          empty_ph ^= 1u << slot;
          mbar_arrive_expect_tx(&dz_full[slot], dz_bytes);
          tma_load_4d(dz_u32 + slot * WS_DZ_SUB_MAX, &tmap_dz, &dz_full[slot], 0, x0, y, img);
          if (slot == 0 && y >= 0) {
            mbar_wait(&dz_empty[4], ((empty_ph >> 4) & 1) ^ 1, 51);
            empty_ph ^= 1u << 4;
            mbar_arrive_expect_tx(&dz_full[4], dz_bytes);
            tma_load_4d(dz_u32 + 4 * WS_DZ_SUB_MAX, &tmap_dz, &dz_full[4], 0, x0, y, img);
          }
        }
        if (y < 0) continue;                                     // rows -2, -1 only prime the ring with zeros
        int vy = y - g.pad_t;
        if (g.pad_mode == FP_PAD_REFLECT) vy = vy < 0 ? -vy : (vy >= g.Hin ? 2 * g.Hin - 2 - vy : vy);
        const bool row_ok = static_cast<unsigned>(vy) < static_cast<unsigned>(g.Hin);
        const uint32_t rowoff0 = static_cast<uint32_t>(((img * g.H0 + (g.up0 ? vy >> 1 : vy)) * g.W0) * g.C0);
        const uint32_t rowoff1 = static_cast<uint32_t>(((img * g.Hin + vy) * g.Win) * g.C1);
        for (int j = 0; j < ncb; ++j, ++xs) {
          const int s = xs % X_STAGES;
          mbar_wait(&x_empty[s], ((xs / X_STAGES) & 1) ^ 1, 52);
          const uint32_t stage = x_u32 + s * STRIP_BYTES;
          const int ch = (cb0 + j) * CB;
          const bool second = ch >= g.C0;
          const uint8_t* sbase = second ? g.src1 : g.src0;
          const uint32_t rowoff = second ? rowoff1 : rowoff0;
          const uint32_t* xcol = xo + (second ? WS_STRIP_ROWS : 0);
          const uint32_t cc = static_cast<uint32_t>(second ? ch - g.C0 : ch) + c * 8;
#pragma unroll
          for (int k = 0; k < PASSES; ++k) {
            const int r = r0 + ROWS_PER_PASS * k;
            if (r < strip_pix) {
              const uint32_t xoff = xcol[r];
              const bool ok = row_ok && (xoff != 0xFFFFFFFFu);
              const uint32_t off = ok ? rowoff + xoff + cc : 0u;
              const int sw = CB == 64 ? (r & 7) : (CB == 32 ? ((r >> 1) & 3) : ((r >> 2) & 1));
              cp_async_16_zfill(stage + r * ROW_BYTES + ((c ^ sw) << 4), sbase + static_cast<size_t>(off) * 2, ok);
            }
          }
          cp_async_mbar_arrive_noinc(&x_full[s]);
        }
      }
    }
  } else if (warp == WG_MMA_WARP) {
    // =============================== MMA ISSUER ===============================
    if (lane == 0) {
      const uint32_t idesc3 = umma_idesc_f16(128, 3 * CB, p.a_fmt, p.b_fmt, 1, 1);
      const uint32_t ring_u32 = smem_u32(dz_ring);
      uint32_t xs = 0;
      uint32_t full_ph = 0;                                      // bit s: phase of dz_full[s] to wait for next
      bool first = true;
      auto wait_full = [&](int slot) {
        mbar_wait(&dz_full[slot], (full_ph >> slot) & 1, 53);
        full_ph ^= 1u << slot;
      };
      for (int col = col_begin; col < col_end; ++col) {
        for (int y = 0; y < g.Ho + 2; ++y) {
          if (y == 0) { wait_full(2); wait_full(3); }            // rows -2, -1 (slots 2, 3)
          const int s_new = y & 3;
          wait_full(s_new);
          if (s_new == 0) wait_full(4);
          const uint32_t a_pair = ring_u32 + ((y + 3) & 3) * WS_DZ_SUB_MAX;      // slot(y-1); its successor holds row y
          const uint32_t a_single = ring_u32 + ((y + 2) & 3) * WS_DZ_SUB_MAX;    // slot(y-2)
          for (int j = 0; j < ncb; ++j, ++xs) {
            const int s = xs % X_STAGES;
            mbar_wait(&x_full[s], (xs / X_STAGES) & 1, 54);
            tcgen05_fence_after();
            const uint32_t strip = smem_u32(x_tiles) + s * STRIP_BYTES;
            const uint32_t d_pair = tmem_base + j * 3 * CB;
#pragma unroll 2
            for (int ks = 0; ks < ksteps; ++ks) {
              const uint64_t b_desc = umma_smem_desc(strip + ks * 16 * ROW_BYTES, ROW_BYTES, 8 * ROW_BYTES, LAYOUT);
              const uint32_t acc = (first && ks == 0) ? 0u : 1u;
              umma_f16(d_pair, umma_smem_desc_sw128(a_pair + ks * 2048, WS_DZ_SUB_MAX, 1024), b_desc, idesc3, acc);
              umma_f16(d_pair + SINGLE_COL, umma_smem_desc_sw128(a_single + ks * 2048, WS_DZ_SUB_MAX, 1024), b_desc, idesc3,
                       acc);
            }
            umma_commit(&x_empty[s]);
          }
          first = false;
          // row y-2 is no longer referenced (it was the "single" row); the duplicate slot is free after its pair
          umma_commit(&dz_empty[(y + 2) & 3]);
          if (s_new == 0) umma_commit(&dz_empty[4]);
        }
        // rows Ho, Ho+1 (zero rows) still occupy their slots: hand them back for the next column
        umma_commit(&dz_empty[g.Ho & 3]);
        umma_commit(&dz_empty[(g.Ho + 1) & 3]);
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  } else {
    // =============================== EPILOGUE (once) ===============================
    const int q = warp & 3;
    const int row = q * 32 + lane;                               // TMEM lane: 0-63 -> (dy 1 | dy 2), 64-127 -> dy 0
    mbar_wait<256>(done_bar, 0, 55);
    tcgen05_fence_after();
    const int co = row & 63;
    const bool row_ok = co < p.Cout;                             // (the host never creates an empty split)
    float* dst_row = p.partial + (static_cast<size_t>(split) * 64 + co) * p.kpad;
    for (int blk = 0; blk < 2; ++blk) {                          // 0: pair accumulator, 1: single accumulator
      if (blk == 1 && row >= 64) break;                          // warp-uniform (q >= 2): lanes 64-127 of "single" are unused
      const int dy = blk == 1 ? 2 : (row < 64 ? 1 : 0);
      for (int j = 0; j < ncb; ++j)
        for (int dx = 0; dx < 3; ++dx) {
          float* dst = dst_row + (dy * 3 + dx) * p.cin + (cb0 + j) * CB;
#pragma unroll 1
          for (int cidx = 0; cidx < CB; cidx += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(tmem_base + blk * SINGLE_COL + (j * 3 + dx) * CB + cidx + (static_cast<uint32_t>(q * 32) << 16), v);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int x = 0; x < 4; ++x)
                *reinterpret_cast<float4*>(dst + cidx + 4 * x) =
                    make_float4(__uint_as_float(v[4 * x]), __uint_as_float(v[4 * x + 1]), __uint_as_float(v[4 * x + 2]),
                                __uint_as_float(v[4 * x + 3]));
            }
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

struct WgradStripPlan {
  int cb, seg, tiles_per_row, num_segments, cblocks, cb_per_unit, cb_groups, ndy, dy_groups, co_tiles, splits, segs_per_split;
};

static int wgrad_strip_cb(const FpConvDesc* d) {
  const int c0 = d->c0, c1 = d->c1;
  if (c0 % 64 == 0 && c1 % 64 == 0) return 64;
  if (c0 % 32 == 0 && c1 % 32 == 0) return 32;
  if (c0 % 16 == 0 && c1 % 16 == 0) return 16;
  return 0;
}

// segment length: split the row into the fewest segments of <= WS_MAX_SEG pixels, rounded up to the MMA's K = 16
// (measured: 160-pixel segments beat 128 for 64-channel blocks -- fewer barrier round trips per pixel -- but not for the
// 32/16-channel blocks that keep all three filter rows in one CTA, which stay at 128 when the row divides evenly)
static int wgrad_strip_seg(int wo, int cb) {
  if (cb < 64 && wo % 128 == 0) return 128;
  const int n = (wo + WS_MAX_SEG - 1) / WS_MAX_SEG;
  const int seg = ((wo + n - 1) / n + 15) / 16 * 16;
  return seg < 32 ? 32 : seg;
}

static bool wgrad_strip_ok(const FpConvDesc* d) {
  const int cb = wgrad_strip_cb(d);
  const int seg = wgrad_strip_seg(d->wo, cb);
  const int segs = (d->wo + seg - 1) / seg;
  const int cin = d->c0 + d->c1;
  return d->kh == 3 && d->kw == 3 && d->stride == 1 && d->up0 != 2 && cb != 0 && d->kpad >= 9 * cin &&
         (cb == 64 || cin / cb <= 2) && d->wo >= 64 && d->wo * 100 >= segs * seg * 80 && opts().wgrad_strip != 0;
}

static WgradStripPlan wgrad_strip_plan(const FpConvDesc* d) {
  WgradStripPlan w;
  w.cb = wgrad_strip_cb(d);
  w.seg = wgrad_strip_seg(d->wo, w.cb);
  w.tiles_per_row = (d->wo + w.seg - 1) / w.seg;
  w.num_segments = d->batch * d->ho * w.tiles_per_row;
  w.cblocks = (d->c0 + d->c1) / w.cb;
  const int max_cb = WS_MAX_COLS / (3 * w.cb);               // channel blocks whose 3 taps fit the accumulator budget
  w.cb_groups = (w.cblocks + max_cb - 1) / max_cb;
  w.cb_per_unit = (w.cblocks + w.cb_groups - 1) / w.cb_groups;
  w.cb_groups = (w.cblocks + w.cb_per_unit - 1) / w.cb_per_unit;
  w.ndy = 9 * w.cb * w.cb_per_unit <= WS_MAX_COLS ? 3 : 1;
  w.dy_groups = 3 / w.ndy;
  w.co_tiles = (d->cout + 127) / 128;
  const int base = w.cb_groups * w.dy_groups * w.co_tiles;
  int want = num_sms() / base;                                 // one resident CTA per SM: a single full wave
  if (want < 1) want = 1;
  if (want > w.num_segments) want = w.num_segments;
  w.segs_per_split = (w.num_segments + want - 1) / want;
  w.splits = (w.num_segments + w.segs_per_split - 1) / w.segs_per_split;
  return w;
}

struct WgradRollPlan { int cb, seg, tiles_per_row, num_cols, cblocks, cb_per_unit, cb_groups, splits, cols_per_split; };

// rolling variant (default; FP_WGRAD_ROLL=0 falls back to the strip kernel): strip-eligible layers with Cout <= 64
static bool wgrad_roll_ok(const FpConvDesc* d) {
  return opts().wgrad_roll != 0 && d->cout <= 64 && wgrad_strip_ok(d);
}

static WgradRollPlan wgrad_roll_plan(const FpConvDesc* d) {
  WgradRollPlan w;
  w.cb = wgrad_strip_cb(d);
  w.seg = wgrad_strip_seg(d->wo, 64);                          // always the longest segments: one TMA box per dZ row
  w.tiles_per_row = (d->wo + w.seg - 1) / w.seg;
  w.num_cols = d->batch * w.tiles_per_row;
  w.cblocks = (d->c0 + d->c1) / w.cb;
  const int max_cb = 256 / (3 * w.cb);                         // pair block and single block: 256 TMEM columns each
  w.cb_groups = (w.cblocks + max_cb - 1) / max_cb;
  w.cb_per_unit = (w.cblocks + w.cb_groups - 1) / w.cb_groups;
  w.cb_groups = (w.cblocks + w.cb_per_unit - 1) / w.cb_per_unit;
  int want = num_sms() / w.cb_groups;
  if (want < 1) want = 1;
  if (want > w.num_cols) want = w.num_cols;
  w.cols_per_split = (w.num_cols + want - 1) / want;
  w.splits = (w.num_cols + w.cols_per_split - 1) / w.cols_per_split;
  return w;
}

template <int CB>
static int launch_wgrad_roll(const WgradRollParams& rp, const CUtensorMap& tm, int grid, cudaStream_t st) {
  { const int arc = ensure_max_smem(reinterpret_cast<const void*>(conv_wgrad_roll_kernel<CB>), wr_smem_bytes<CB>(), "conv_wgrad_roll_kernel"); if (arc != FP_OK) return arc; }
  conv_wgrad_roll_kernel<CB><<<grid, WG_THREADS, wr_smem_bytes<CB>(), st>>>(rp, tm);
  FP_CHECK_LAUNCH("fp_conv_wgrad(roll)");
  return FP_OK;
}

template <int CB>
static int launch_wgrad_strip(const WgradStripParams& sp, const CUtensorMap& tm, int grid, cudaStream_t st) {
  { const int arc = ensure_max_smem(reinterpret_cast<const void*>(conv_wgrad_strip_kernel<CB>), ws_smem_bytes<CB>(), "conv_wgrad_strip_kernel"); if (arc != FP_OK) return arc; }
  conv_wgrad_strip_kernel<CB><<<grid, WG_THREADS, ws_smem_bytes<CB>(), st>>>(sp, tm);
  FP_CHECK_LAUNCH("fp_conv_wgrad(strip)");
  return FP_OK;
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
#pragma unroll 4                                          // 16 loads in flight: the kernel is pure load latency
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
  // one CTA is resident per SM: never exceed two full waves (rounding UP used to leave a third wave of 1-4 CTAs,
  // i.e. +50 % on the 68 general weight-gradient launches of a step)
  int want = 2 * num_sms() / base;
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
  if (wgrad_roll_ok(d)) {
    const WgradRollPlan wr = wgrad_roll_plan(d);
    return static_cast<long long>(wr.splits) * 64 * d->kpad * 4;
  }
  if (wgrad_strip_ok(d)) {
    const WgradStripPlan ws = wgrad_strip_plan(d);
    return static_cast<long long>(ws.splits) * ws.co_tiles * 128 * d->kpad * 4;
  }
  const WgradPlan w = wgrad_plan(d);
  return static_cast<long long>(w.splits) * w.co_tiles * 128 * d->kpad * 4;
}

extern "C" int fp_conv_wgrad(const FpConvDesc* d, const void* src0, const void* src1, const void* dz, int dz_ld,
                             int grad_dtype, int stem, int accumulate, float* dw_oihw, float* workspace, void* stream) {
  if (!d || !src0 || !dz || !dw_oihw || !workspace) return set_error(FP_ERR_ARG, "fp_conv_wgrad: null argument");
  WgradParams p{};
  int rc = fill_gather_geom(d, src0, src1, "fp_conv_wgrad", p.g, stem || !wgrad_strip_ok(d));
  if (rc != FP_OK) return rc;
  if (dz_ld < d->cout || (dz_ld & 7)) return set_error(FP_ERR_ARG, "fp_conv_wgrad: dz_ld=%d must be a multiple of 8 >= cout", dz_ld);
  rc = fp_check_device();
  if (rc != FP_OK) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cin_r = stem ? 3 : d->c0 + d->c1;
  const int kh_r = stem ? 7 : d->kh, kw_r = stem ? 7 : d->kw;
  if (!stem && wgrad_roll_ok(d)) {
    WgradRollParams rp{};
    rp.g = p.g;
    const WgradRollPlan wr = wgrad_roll_plan(d);
    rp.Cout = d->cout; rp.kpad = d->kpad; rp.cin = d->c0 + d->c1;
    rp.seg = wr.seg; rp.tiles_per_row = wr.tiles_per_row; rp.num_cols = wr.num_cols;
    rp.cblocks = wr.cblocks; rp.cb_per_unit = wr.cb_per_unit; rp.cb_groups = wr.cb_groups;
    rp.splits = wr.splits; rp.cols_per_split = wr.cols_per_split;
    rp.partial = workspace; rp.a_fmt = grad_dtype; rp.b_fmt = d->dtype;
    CUtensorMap tmap_dz4;
    rc = make_tmap_nhwc(&tmap_dz4, dz, static_cast<uint64_t>(dz_ld), static_cast<uint64_t>(d->wo),
                        static_cast<uint64_t>(d->ho), static_cast<uint64_t>(d->batch), 64u,
                        static_cast<uint32_t>(wr.seg), 128);
    if (rc != FP_OK) return rc;
    const int grid_r = wr.cb_groups * wr.splits;
    rc = wr.cb == 64 ? launch_wgrad_roll<64>(rp, tmap_dz4, grid_r, st)
                     : (wr.cb == 32 ? launch_wgrad_roll<32>(rp, tmap_dz4, grid_r, st) : launch_wgrad_roll<16>(rp, tmap_dz4, grid_r, st));
    if (rc != FP_OK) return rc;
    const long long total_r = static_cast<long long>(d->cout) * d->kpad;
    wgrad_reduce_kernel<<<static_cast<unsigned>((total_r + 255) / 256), 256, 0, st>>>(
        workspace, wr.splits, 64, d->kpad, d->cout, cin_r, d->c0 + d->c1, kh_r, kw_r, 0, accumulate, dw_oihw);
    FP_CHECK_LAUNCH("fp_conv_wgrad(roll reduce)");
    return FP_OK;
  }
  if (!stem && wgrad_strip_ok(d)) {
    WgradStripParams sp{};
    sp.g = p.g;
    const WgradStripPlan ws = wgrad_strip_plan(d);
    sp.Cout = d->cout; sp.kpad = d->kpad; sp.cin = d->c0 + d->c1;
    sp.seg = ws.seg; sp.tiles_per_row = ws.tiles_per_row; sp.num_segments = ws.num_segments; sp.cblocks = ws.cblocks;
    sp.cb_per_unit = ws.cb_per_unit; sp.cb_groups = ws.cb_groups; sp.ndy = ws.ndy; sp.dy_groups = ws.dy_groups;
    sp.co_tiles = ws.co_tiles; sp.splits = ws.splits; sp.segs_per_split = ws.segs_per_split;
    sp.partial = workspace; sp.a_fmt = grad_dtype; sp.b_fmt = d->dtype;
    sp.fuse_taps = opts().wgrad_fuse_taps;
    CUtensorMap tmap_dz3;
    rc = make_tmap_3d_sw128(&tmap_dz3, dz, static_cast<uint64_t>(dz_ld), static_cast<uint64_t>(d->wo),
                            static_cast<uint64_t>(d->batch) * d->ho, 64u, static_cast<uint32_t>(ws.seg));
    if (rc != FP_OK) return rc;
    const int grid_s = ws.cb_groups * ws.dy_groups * ws.co_tiles * ws.splits;
    rc = ws.cb == 64 ? launch_wgrad_strip<64>(sp, tmap_dz3, grid_s, st)
                     : (ws.cb == 32 ? launch_wgrad_strip<32>(sp, tmap_dz3, grid_s, st) : launch_wgrad_strip<16>(sp, tmap_dz3, grid_s, st));
    if (rc != FP_OK) return rc;
    const long long total_s = static_cast<long long>(d->cout) * d->kpad;
    wgrad_reduce_kernel<<<static_cast<unsigned>((total_s + 255) / 256), 256, 0, st>>>(
        workspace, ws.splits, ws.co_tiles * 128, d->kpad, d->cout, cin_r, d->c0 + d->c1, kh_r, kw_r, 0, accumulate, dw_oihw);
    FP_CHECK_LAUNCH("fp_conv_wgrad(strip reduce)");
    return FP_OK;
  }
  p.dz_ld = dz_ld; p.Cout = d->cout; p.kpad = d->kpad;
  const WgradPlan w = wgrad_plan(d);
  p.k_chunks = w.k_chunks; p.chunks_per_group = w.chunks_per_group; p.k_groups = w.k_groups;
  p.co_tiles = w.co_tiles; p.splits = w.splits; p.blocks_per_split = w.blocks_per_split; p.num_pix_blocks = w.num_pix_blocks;
  p.partial = workspace;
  p.a_fmt = grad_dtype; p.b_fmt = d->dtype;
  { const int arc = ensure_max_smem(reinterpret_cast<const void*>(conv_wgrad_kernel), WG_SMEM_BYTES, "conv_wgrad_kernel"); if (arc != FP_OK) return arc; }
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
