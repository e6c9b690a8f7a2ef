// Library-wide plumbing: version, thread-local error string, device check, TMA descriptor factory.
#include <algorithm>
#include <stdarg.h>
#include <stdlib.h>
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

namespace {
thread_local bool g_pdl_prev_kernel = false;
// read per launch (not cached): a captured graph keeps the edges it was captured with, so a tool can capture the
// same step both ways in one process (tools/pdl_check.py)
bool pdl_enabled() {
  const char* e = getenv("CE_PDL");
  return e != nullptr ? e[0] != '0' : true;
}
}  // namespace
bool pdl_use() { return g_pdl_prev_kernel && pdl_enabled(); }
void pdl_mark() { g_pdl_prev_kernel = true; }
void pdl_break() { g_pdl_prev_kernel = false; }

int check_device() {
  pdl_break();   // every API call starts a chain of its own
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

// ------------------------------------------------------------------------------------------
// Device-side exchange over peer memory (NVLink / NVSwitch): the sharded loss head's all-gathers and its
// gradient reduce-scatter as plain loads from the peers' symmetric buffers -- utils.py:192-206 without a
// collective library in the data path.  The caller provides the peers' buffer addresses (same layout on
// every rank) and orders the steps with its own barrier (torch symmetric memory: hdl.barrier()).
// ------------------------------------------------------------------------------------------
namespace ce {
namespace {
constexpr int kMaxPeers = 16;
struct PeerPtrs { const void* p[kMaxPeers]; };

__global__ void p2p_gather_kernel(PeerPtrs pp, int world, int64_t vec_each, uint4* dst) {
  const int64_t total = vec_each * world;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec_each);
    dst[i] = reinterpret_cast<const uint4*>(pp.p[r])[i - r * vec_each];
  }
}
// dst[i] = sum_r peer_r[offset + i] (fixed order r = 0..world-1: every rank forms the same bits), 16 bytes per thread;
// tail: tail_dst[j] = sum_r peer_r[tail_offset + j] for a few scalars riding in the same buffer
__global__ void p2p_reduce_kernel(PeerPtrs pp, int world, int64_t offset, int64_t n4, float4* dst, int64_t tail_offset,
                                  int tail_n, float* tail_dst) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 v = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(pp.p[r]) + offset)[i];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    dst[i] = s;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < tail_n) {
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += reinterpret_cast<const float*>(pp.p[r])[tail_offset + threadIdx.x];
    tail_dst[threadIdx.x] = s;
  }
}
int peer_ptrs(const int64_t* host_ptrs, int world, PeerPtrs* out) {
  if (world < 1 || world > kMaxPeers) return fail(CE_ERR_ARG, "p2p: world size %d outside 1..%d", world, kMaxPeers);
  for (int r = 0; r < world; ++r) {
    if (host_ptrs[r] == 0 || (host_ptrs[r] & 15)) return fail(CE_ERR_ALIGN, "p2p: peer buffer %d is null or not 16-byte aligned", r);
    out->p[r] = reinterpret_cast<const void*>(host_ptrs[r]);
  }
  return CE_OK;
}
}  // namespace
// 3-D tensor map over a batch of row-major [rows, inner] bf16 matrices (batch stride in elements), 128-byte
// swizzle, out-of-bounds rows / samples read as zeros and are clipped on stores (csrc/ot_wide.cu).
int make_tmap3d_bf16(CUtensorMap* out, const void* ptr, uint64_t inner, uint64_t rows, uint64_t batch,
                     uint64_t row_stride_elems, uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return fail(CE_ERR_ARCH, "cuTensorMapEncodeTiled is not available from this driver");
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (row_stride_elems * 2) % 16 || (batch_stride_elems * 2) % 16)
    return fail(CE_ERR_ALIGN, "TMA operand must be 16-byte aligned with 16-byte multiple strides");
  cuuint64_t dims[3] = {inner, rows, batch};
  cuuint64_t strides[2] = {row_stride_elems * 2, batch_stride_elems * 2};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(CE_ERR_ARG, "cuTensorMapEncodeTiled (3-D) failed (%d) inner=%llu rows=%llu batch=%llu box=%ux%u", (int)r,
                (unsigned long long)inner, (unsigned long long)rows, (unsigned long long)batch, box_inner, box_rows);
  return CE_OK;
}

}  // namespace ce

extern "C" int ce_p2p_gather(const int64_t* peer_ptrs_host, int world, int64_t bytes_each, void* dst, ce_stream_t stream) {
  CE_TRY(ce::check_device());
  ce::PeerPtrs pp{};
  CE_TRY(ce::peer_ptrs(peer_ptrs_host, world, &pp));
  if (bytes_each <= 0 || bytes_each % 16 || ((uintptr_t)dst & 15)) return ce::fail(CE_ERR_ALIGN, "p2p gather: sizes and pointers must be multiples of 16 bytes");
  const int64_t vec = bytes_each / 16;
  const int blocks = (int)std::min<int64_t>((vec * world + 255) / 256, 148 * 8);
  ce::p2p_gather_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pp, world, vec, reinterpret_cast<uint4*>(dst));
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_p2p_reduce_f32(const int64_t* peer_ptrs_host, int world, int64_t offset_elems, int64_t n, float* dst,
                                 int64_t tail_offset_elems, int tail_n, float* tail_dst, ce_stream_t stream) {
  CE_TRY(ce::check_device());
  ce::PeerPtrs pp{};
  CE_TRY(ce::peer_ptrs(peer_ptrs_host, world, &pp));
  if (n <= 0 || n % 4 || offset_elems % 4 || ((uintptr_t)dst & 15)) return ce::fail(CE_ERR_ALIGN, "p2p reduce: n and offset must be multiples of 4 floats");
  if (tail_n < 0 || tail_n > 32 || (tail_n > 0 && tail_dst == nullptr)) return ce::fail(CE_ERR_ARG, "p2p reduce: bad tail");
  const int blocks = (int)std::min<int64_t>((n / 4 + 255) / 256, 148 * 8);
  ce::p2p_reduce_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(pp, world, offset_elems, n / 4, reinterpret_cast<float4*>(dst),
                                                                               tail_offset_elems, tail_n, tail_dst);
  CE_LAUNCH_CHECK();
  return CE_OK;
}

extern "C" int ce_version(void) { return 100; }
extern "C" const char* ce_last_error(void) { return ce::g_err; }
extern "C" int ce_device_check(void) { return ce::check_device(); }
extern "C" unsigned long long ce_debug_launch_count(void) { return ce::g_launches; }
