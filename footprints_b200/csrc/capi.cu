// C-ABI plumbing: error strings, device check, TMA descriptor creation (driver entry point fetched at
// run time so the library links against cudart only).
#include "fp_internal.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cudaTypedefs.h>
#include <mutex>
#include <unordered_set>

namespace fp {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// per-device state (CUDA function attributes and properties are per device; one process may drive several GPUs)
constexpr int kMaxDevices = 64;
static int g_num_sms[kMaxDevices];            // 0 = unknown
static int g_dev_status[kMaxDevices];         // 0 = unknown, 1 = ok, < 0 = error code
static std::mutex g_mu;
static std::unordered_set<unsigned long long> g_attr_done;

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

int num_sms() {
  const int dev = current_device();
  return (dev >= 0 && g_num_sms[dev] > 0) ? g_num_sms[dev] : 148;
}

int ensure_max_smem(const void* kernel, int bytes, const char* what) {
  const int dev = current_device();
  if (dev < 0) return set_error(FP_ERR_CUDA, "%s: no current CUDA device", what);
  const unsigned long long key = (reinterpret_cast<unsigned long long>(kernel) << 6) ^ static_cast<unsigned long long>(dev);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_attr_done.count(key)) return FP_OK;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "cudaFuncSetAttribute(%s, smem=%d): %s", what, bytes, cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(g_mu);
  g_attr_done.insert(key);
  return FP_OK;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
static Options g_opts;
static std::atomic<int> g_opts_loaded{0};
static void load_options() {
  Options o;
  o.strip = env_int("FP_STRIP", 1);
  o.strip_mineff = env_int("FP_STRIP_MINEFF", 80);
  o.strip_tma = env_int("FP_STRIP_TMA", 1);
  o.strip_wres = env_int("FP_STRIP_WRES", 1);
  o.rowstream = env_int("FP_ROWSTREAM", 1);
  o.rowstream_mineff = env_int("FP_ROWSTREAM_MINEFF", 60);
  o.a_mode = env_int("FP_A_MODE", 0);
  o.debug_flags = env_int("FP_DEBUG_FLAGS", 0);
  o.wgrad_strip = env_int("FP_WGRAD_STRIP", 1);
  o.wgrad_roll = env_int("FP_WGRAD_ROLL", 1);
  o.wgrad_fuse_taps = env_int("FP_WGRAD_FUSE_TAPS", 1);
  o.head_tile = env_int("FP_HEAD_TILE", 1);
  o.head_hmath = env_int("FP_HEAD_HMATH", 1);
  o.head_epi = env_int("FP_HEAD_EPI", 1);
  o.resize_rows = env_int("FP_RESIZE_ROWS", 1);
  std::lock_guard<std::mutex> lk(g_mu);
  g_opts = o;
  g_opts_loaded.store(1);
}
const Options& opts() {
  if (!g_opts_loaded.load()) load_options();
  return g_opts;
}

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

extern "C" int fp_reload_options(void) {
  fp::load_options();
  return FP_OK;
}

extern "C" int fp_check_device(void) {
  using namespace fp;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "no CUDA device: %s (there is no CPU fallback)", cudaGetErrorString(e));
  if (dev < 0 || dev >= kMaxDevices) return set_error(FP_ERR_UNSUPPORTED, "device ordinal %d out of range", dev);
  if (g_dev_status[dev] == 1) return FP_OK;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) return set_error(FP_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
  if (prop.major != 10 || prop.minor != 0) {       // the library carries an sm_100a image only (sm_103 cannot run it)
    g_dev_status[dev] = FP_ERR_ARCH;
    return set_error(FP_ERR_ARCH, "device %d is sm_%d%d; footprints_b200 kernels are sm_100a only", dev, prop.major, prop.minor);
  }
  g_num_sms[dev] = prop.multiProcessorCount;
  g_dev_status[dev] = 1;
  return FP_OK;
}
