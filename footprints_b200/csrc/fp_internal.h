// Internal host-side helpers shared by the translation units of libfootprints_b200.so.
#pragma once
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/footprints_b200.h"

namespace fp {

// printf-style thread-local error message; returns `code` so callers can `return set_error(...)`.
int set_error(int code, const char* fmt, ...);
int num_sms();       // SM count of the CURRENT device (cached per device ordinal)

// Tuning / triage switches.  Read from the environment ONCE (first use) instead of ~10 getenv()
// calls per convolution launch; fp_reload_options() re-reads them (the tests flip switches inside one process).
struct Options {
  int strip, strip_mineff, strip_tma, strip_wres, rowstream, rowstream_mineff, a_mode, debug_flags;
  int wgrad_strip, wgrad_roll, wgrad_fuse_taps, head_tile, head_hmath, head_epi, resize_rows;
};
const Options& opts();

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (device, kernel): function attributes are per device, so
// a process-wide `static bool` would skip the opt-in on every device but the first.
int ensure_max_smem(const void* kernel, int bytes, const char* what);
// 2-D row-major 16-bit matrix [rows][cols] -> tensor map with box {box_cols, box_rows}, SWIZZLE_128B.
int make_tmap_2d_sw128(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                       uint32_t box_rows);
// same with an explicit swizzle span in bytes (128 or 64; box_cols * 2 must equal it)
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                 int swizzle_bytes);

int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t n, uint32_t box_c,
                   uint32_t box_w, int swizzle_bytes);
int make_tmap_3d_sw128(CUtensorMap* out, const void* base, uint64_t cols, uint64_t width, uint64_t rows, uint32_t box_cols,
                       uint32_t box_width);

#define FP_CHECK_LAUNCH(name)                                                                  \
  do {                                                                                         \
    cudaError_t e__ = cudaGetLastError();                                                      \
    if (e__ != cudaSuccess) return fp::set_error(FP_ERR_CUDA, "%s: %s", name, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace fp
