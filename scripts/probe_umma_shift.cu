// Hardware probe (not part of the library): can a K-major SWIZZLE_128B UMMA operand start at a 128-byte row that is
// not a multiple of 8 rows from the 1024-byte-aligned swizzle pattern?  If so, the three horizontal taps of a 3x3
// convolution can share one staged row strip (descriptor start advanced by one row per tap) instead of three fills.
//   usage: probe_umma_shift        prints, per (shift, base_offset mode), whether D[m][n] == A[m + shift][n]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "fp_ptx.cuh"

using namespace fp;

constexpr int STRIP_ROWS = 144;   // 128 + 16 spare rows

__global__ void __launch_bounds__(128) probe_kernel(const __half* a_gl /*[144][64]*/, const __half* b_gl /*[16][64]*/,
                                                    int shift, int base_offset_mode, float* d_out /*[128][16]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_s = smem;                       // 144 rows x 128 B
  uint8_t* b_s = smem + 18432 + 1024 - (18432 % 1024);   // next 1024-aligned
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 2048);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // fill A and B with the canonical SW128 pattern by ABSOLUTE position: 16-byte chunk c of row r -> c ^ (r & 7)
  for (int i = tid; i < STRIP_ROWS * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(a_gl + r * 64 + c * 8);
  }
  for (int i = tid; i < 16 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b_gl + r * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 32);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 16, 0, 0, 0, 0);
    const uint32_t a_addr = smem_u32(a_s) + shift * 128;
    uint64_t a_desc = umma_smem_desc_sw128(a_addr, 16, 1024);
    if (base_offset_mode == 1) a_desc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
    const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(b_s), 16, 1024);
    for (int k = 0; k < 4; ++k) umma_f16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 99);
  tcgen05_fence_after();
  uint32_t v[16];
  tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int x = 0; x < 16; ++x) d_out[(warp * 32 + (tid & 31)) * 16 + x] = __uint_as_float(v[x]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// ---- probe 2: K-major SWIZZLE_64B (64-byte rows = 32 x 16-bit): does a start advanced by whole rows work?
__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t lbo, uint32_t sbo, uint64_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= layout << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_sw64_kernel(const __half* a_gl /*[144][32]*/, const __half* b_gl /*[16][32]*/,
                                                         int shift, float* d_out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_s = smem;                 // 144 rows x 64 B = 9216
  uint8_t* b_s = smem + 10240;         // 16 rows x 64 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  // Swizzle<2,4,3>: 16-byte chunk c of a 64-byte row at byte address a -> c ^ ((a >> 7) & 3)
  for (int i = tid; i < STRIP_ROWS * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(a_s + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(a_gl + r * 32 + c * 8);
  }
  for (int i = tid; i < 16 * 4; i += 128) {
    const int r = i >> 2, c = i & 3;
    *reinterpret_cast<uint4*>(b_s + r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = *reinterpret_cast<const uint4*>(b_gl + r * 32 + c * 8);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 32);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 16, 0, 0, 0, 0);
    const uint64_t a_desc = desc_generic(smem_u32(a_s) + shift * 64, 16, 512, 4);
    const uint64_t b_desc = desc_generic(smem_u32(b_s), 16, 512, 4);
    for (int k = 0; k < 2; ++k) umma_f16(tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k != 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 98);
  tcgen05_fence_after();
  uint32_t v[16];
  tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int x = 0; x < 16; ++x) d_out[(warp * 32 + (tid & 31)) * 16 + x] = __uint_as_float(v[x]);
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 32);
}

// ---- probe 3: MN-major SWIZZLE_128B operands (the weight-gradient GEMM: rows = pixels = K): B start advanced by rows
__global__ void __launch_bounds__(128) probe_mn_kernel(const __half* a_gl /*[16 px][128 co]*/, const __half* b_gl /*[32 px][64 ch]*/,
                                                       int shift, float* d_out /*[128][64]*/) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* a_s = smem;                 // two sub-tiles [16 px][64 co] = 2 x 2048
  uint8_t* b_s = smem + 4096;          // [32 px][64 ch] = 4096
  uint64_t* bar = reinterpret_cast<uint64_t*>(b_s + 4096);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 2 * 16 * 8; i += 128) {
    const int sub = i >> 7, r = (i >> 3) & 15, c = i & 7;
    *reinterpret_cast<uint4*>(a_s + sub * 2048 + r * 128 + ((c ^ (r & 7)) << 4)) =
        *reinterpret_cast<const uint4*>(a_gl + r * 128 + sub * 64 + c * 8);
  }
  for (int i = tid; i < 32 * 8; i += 128) {
    const int r = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(b_s + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b_gl + r * 64 + c * 8);
  }
  if (tid == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(tptr, 64);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tptr;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_f16(128, 64, 0, 0, 1, 1);
    const uint64_t a_desc = umma_smem_desc_sw128(smem_u32(a_s), 2048, 1024);
    const uint64_t b_desc = umma_smem_desc_sw128(smem_u32(b_s) + shift * 128, 4096, 1024);
    umma_f16(tmem, a_desc, b_desc, idesc, 0);
    umma_commit(bar);
  }
  mbar_wait(bar, 0, 97);
  tcgen05_fence_after();
  for (int j = 0; j < 64; j += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + j + (static_cast<uint32_t>(warp * 32) << 16), v);
    tmem_ld_wait();
    for (int x = 0; x < 16; ++x) d_out[(warp * 32 + (tid & 31)) * 64 + j + x] = __uint_as_float(v[x]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static int probe_sw64_and_mn() {
  cudaFuncSetAttribute(probe_sw64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaFuncSetAttribute(probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  {
    std::vector<__half> a(STRIP_ROWS * 32), b(16 * 32);
    for (int r = 0; r < STRIP_ROWS; ++r) for (int k = 0; k < 32; ++k) a[r * 32 + k] = __float2half(static_cast<float>((r * 32 + k) % 2048));
    for (int n = 0; n < 16; ++n) for (int k = 0; k < 32; ++k) b[n * 32 + k] = __float2half(k == 2 * n + 1 ? 1.f : 0.f);
    __half *da, *db; float* dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 128 * 16 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    std::vector<float> d(128 * 16);
    for (int shift = 0; shift <= 5; ++shift) {
      probe_sw64_kernel<<<1, 128, 32768>>>(da, db, shift, dd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("SW64 shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 16; ++n)
        if (d[m * 16 + n] != static_cast<float>(((m + shift) * 32 + 2 * n + 1) % 2048)) ++bad;
      printf("K-major SWIZZLE_64B shift %d rows: %s (%d/2048 wrong; d[0]=%.0f d[1]=%.0f d[16]=%.0f)\n", shift, bad ? "MISMATCH" : "OK", bad,
             d[0], d[1], d[16]);
    }
  }
  {
    std::vector<__half> a(16 * 128), b(32 * 64);
    for (int p = 0; p < 16; ++p) for (int co = 0; co < 128; ++co) a[p * 128 + co] = __float2half(co == p ? 1.f : 0.f);
    for (int r = 0; r < 32; ++r) for (int c = 0; c < 64; ++c) b[r * 64 + c] = __float2half(static_cast<float>((r * 64 + c) % 2048));
    __half *da, *db; float* dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    std::vector<float> d(128 * 64);
    for (int shift = 0; shift <= 9; ++shift) {
      probe_mn_kernel<<<1, 128, 32768>>>(da, db, shift, dd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MN shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        const float expect = m < 16 ? static_cast<float>(((m + shift) * 64 + n) % 2048) : 0.f;
        if (d[m * 64 + n] != expect) ++bad;
      }
      printf("MN-major SWIZZLE_128B B-start shift %d rows: %s (%d/8192 wrong; d[0]=%.0f d[1]=%.0f d[64]=%.0f)\n", shift, bad ? "MISMATCH" : "OK",
             bad, d[0], d[1], d[64]);
    }
  }
  return 0;
}

int main() {
  if (probe_sw64_and_mn()) return 1;
  std::vector<__half> a(STRIP_ROWS * 64), b(16 * 64);
  // A[r][k] = 4*r + (k / 16) for k < 64 restricted to the probed columns; B = 16 selected columns: n -> k = 4*n + 1
  for (int r = 0; r < STRIP_ROWS; ++r)
    for (int k = 0; k < 64; ++k) a[r * 64 + k] = __float2half(static_cast<float>((r * 64 + k) % 2048));
  for (int n = 0; n < 16; ++n)
    for (int k = 0; k < 64; ++k) b[n * 64 + k] = __float2half(k == 4 * n + 1 ? 1.f : 0.f);
  __half *da, *db; float* dd;
  cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 128 * 16 * 4);
  cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  std::vector<float> d(128 * 16);
  for (int mode = 0; mode < 2; ++mode)
    for (int shift = 0; shift <= 9; ++shift) {
      cudaMemset(dd, 0, 128 * 16 * 4);
      probe_kernel<<<1, 128, 32768>>>(da, db, shift, mode, dd);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, shift, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, first = -1;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
          const float expect = static_cast<float>(((m + shift) * 64 + 4 * n + 1) % 2048);
          if (d[m * 16 + n] != expect) { if (first < 0) first = m * 16 + n; ++bad; }
        }
      printf("base_offset_mode %d shift %d: %s (%d/2048 wrong", mode, shift, bad ? "MISMATCH" : "OK", bad);
      if (bad) printf("; first at m=%d n=%d got %.0f want %.0f; row0: %.0f %.0f row1: %.0f row7: %.0f row8: %.0f", first / 16, first % 16,
                      d[first], static_cast<float>(((first / 16 + shift) * 64 + 4 * (first % 16) + 1) % 2048), d[0], d[1], d[16], d[7 * 16], d[8 * 16]);
      printf(")\n");
    }
  return 0;
}
