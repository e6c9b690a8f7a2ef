// Library-wide plumbing: version, thread-local error string, device check, TMA descriptor factory.
#include <stdarg.h>
#include <string.h>

#include "ce_common.cuh"

namespace ce {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

static thread_local int g_dev_checked = -1;  // device ordinal that passed the check
static thread_local int g_sms = 0;

int check_device() {
  // PyTorch runs backward() on its own thread: the driver entry points used for tensor maps need the
  // primary context bound to THAT thread, which only a runtime call that touches the device does
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) { cudaFree(nullptr); ctx_bound = true; }
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    return CE_ERR_ARCH;
  }
  if (dev == g_dev_checked) return CE_OK;
  int major = 0, minor = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10 || minor != 0)
    return fail(CE_ERR_ARCH, "device %d is sm_%d%d; clip_event_b200 is built for sm_100a (B200) only", dev, major, minor);
  g_dev_checked = dev;
  g_sms = sms;
  return CE_OK;
}

int num_sms() { return g_sms > 0 ? g_sms : 148; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, const void* ptr, bool fp32, uint64_t inner, uint64_t outer,
              uint64_t row_stride_elems, uint32_t box_inner, uint32_t box_outer, int swizzle) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(CE_ERR_ARCH, "cuTensorMapEncodeTiled is not available from this driver");
  const uint64_t esz = fp32 ? 4 : 2;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_elems * esz) % 16)
    return fail(CE_ERR_ALIGN, "TMA operand must be 16-byte aligned with a 16-byte multiple row stride");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_elems * esz};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                  : swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CE_ERR_ARG, "cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_elems,
                box_inner, box_outer);
  return CE_OK;
}

}  // namespace ce

extern "C" int ce_version(void) { return 100; }
extern "C" const char* ce_last_error(void) { return ce::g_err; }
extern "C" int ce_device_check(void) { return ce::check_device(); }
extern "C" unsigned long long ce_debug_launch_count(void) { return ce::g_launches; }
