// C-ABI plumbing: error strings, device check, TMA descriptor creation (driver entry point fetched at
// run time so the library links against cudart only).
#include "fp_internal.h"

#include <cstdarg>
#include <cstdio>
#include <cudaTypedefs.h>

namespace fp {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static int g_num_sms = 0;
static int g_dev_status = 1;  // 1 = unknown

int num_sms() { return g_num_sms > 0 ? g_num_sms : 148; }

static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

int make_tmap_2d_sw128(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                       uint32_t box_rows) {
  return make_tmap_2d(out, base, cols, rows, box_cols, box_rows, 128);
}

int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows,
                 int swizzle_bytes) {
  auto enc = get_encode();
  if (enc == nullptr) return set_error(FP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FP_ERR_CUDA, "cuTensorMapEncodeTiled failed: %d", static_cast<int>(r));
  return FP_OK;
}

// NHWC 16-bit activation tensor [n][h][w][c] -> box {box_c, box_w, 1, 1}; coordinates outside the tensor (negative or
// past the end in w / h) are zero-filled, which is exactly a zero-padded convolution's halo
int make_tmap_nhwc(CUtensorMap* out, const void* base, uint64_t c, uint64_t w, uint64_t h, uint64_t n, uint32_t box_c,
                   uint32_t box_w, int swizzle_bytes) {
  auto enc = get_encode();
  if (enc == nullptr) return set_error(FP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[4] = {c, w, h, n};
  cuuint64_t strides[3] = {c * 2, c * w * 2, c * w * h * 2};
  cuuint32_t box[4] = {box_c, box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                               : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FP_ERR_CUDA, "cuTensorMapEncodeTiled(nhwc) failed: %d", static_cast<int>(r));
  return FP_OK;
}

// [rows][width][cols] 16-bit tensor (cols contiguous) -> box {box_cols, box_width, 1}, SWIZZLE_128B; out-of-range
// width positions are zero-filled, which masks the overhang of 128-pixel row segments
int make_tmap_3d_sw128(CUtensorMap* out, const void* base, uint64_t cols, uint64_t width, uint64_t rows, uint32_t box_cols,
                       uint32_t box_width) {
  auto enc = get_encode();
  if (enc == nullptr) return set_error(FP_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {cols, width, rows};
  cuuint64_t strides[2] = {cols * 2, cols * width * 2};
  cuuint32_t box[3] = {box_cols, box_width, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(FP_ERR_CUDA, "cuTensorMapEncodeTiled(3d) failed: %d", static_cast<int>(r));
  return FP_OK;
}

}  // namespace fp

extern "C" int fp_version(void) { return FP_VERSION; }
extern "C" const char* fp_last_error(void) { return fp::g_err; }

extern "C" int fp_check_device(void) {
  using namespace fp;
  if (g_dev_status <= 0) return g_dev_status == 0 ? FP_OK : set_error(g_dev_status, "%s", g_err);
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10) {
    g_dev_status = FP_ERR_ARCH;
    return set_error(FP_ERR_ARCH, "device is sm_%d%d; footprints_b200 kernels are sm_100a only", prop.major, prop.minor);
  }
  g_num_sms = prop.multiProcessorCount;
  g_dev_status = 0;
  return FP_OK;
}
