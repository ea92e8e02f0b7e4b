#include "host_util.cuh"

#include <mutex>
#include <vector>
#include <string.h>

namespace lx {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// libcuda is resolved at run time through the runtime API so the library links (and loads) on a box without a driver.
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 8) != 0) {
    set_error("tensor map: base must be 16-byte aligned and row stride a multiple of 8 elements (ld=%llu)",
              (unsigned long long)ld);
    return LX_ERR_ARG;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d) failed: %d (rows=%llu cols=%llu ld=%llu box=%ux%u)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
    return LX_ERR_CUDA;
  }
  return LX_OK;
}

// fp32 variant (box_cols * 4 bytes must be 128): used as the destination of cp.reduce.async.bulk.tensor (+= tiles)
int make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LX_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) != 0) {
    set_error("tensor map (f32): base must be 16-byte aligned and row stride a multiple of 4 elements");
    return LX_ERR_ARG;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 4};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  // swizzle span = the inner box extent: 32 floats -> 128-byte swizzle, 16 floats -> 64-byte swizzle
  const CUtensorMapSwizzle sw = box_cols * 4 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : box_cols * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(2d f32) failed: %d", (int)r);
    return LX_ERR_CUDA;
  }
  return LX_OK;
}

int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t outer, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint64_t outer_stride, uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
    return LX_ERR_CUDA;
  }
  cuuint64_t dims[3] = {cols, rows, outer};
  cuuint64_t strides[2] = {ld * 2, outer_stride * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed: %d", (int)r);
    return LX_ERR_CUDA;
  }
  return LX_OK;
}

static int g_pdl = 1;
bool pdl_enabled() { return g_pdl != 0; }
}  // namespace lx
extern "C" void lx_debug_set_pdl(int on) { lx::g_pdl = on; }
namespace lx {

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    n = prop.multiProcessorCount;
  }
  return n;
}

// ------------------------------------------------------------------------------------------- exchange workspace
// One caller-owned buffer bound to one stream (lx_set_workspace).  The split-work kernels (attention, GEMM) exchange
// partial tiles through it; a launch on any other stream, or with no / too small a workspace, uses the unsplit schedule.
static void* g_ws_ptr = nullptr;
static int64_t g_ws_bytes = 0;
static void* g_ws_stream = nullptr;
static bool g_ws_set = false;

void* workspace_region(void* stream, int which, size_t need_bytes) {
  if (!g_ws_set || stream != g_ws_stream || which < 0 || which >= WS_REGIONS) return nullptr;
  const size_t region = (size_t)(g_ws_bytes / WS_REGIONS) & ~(size_t)1023;
  if (need_bytes > region) return nullptr;
  return static_cast<char*>(g_ws_ptr) + (size_t)which * region;
}

static int g_skip_mask = 0;
int debug_skip_mask() { return g_skip_mask; }
static long long* g_tl_buf = nullptr;
static int g_tl_cap = 0, g_tl_next = 0;
static std::vector<int> g_tl_cls;
long long* timeline_next(int cls) {
  if (g_tl_buf == nullptr || g_tl_next >= g_tl_cap) return nullptr;
  g_tl_cls.push_back(cls);  // host side: nothing may be enqueued between two kernels (it would change what is measured)
  return g_tl_buf + 4LL * g_tl_next++;
}

// ------------------------------------------------------------------------------------------- launch accounting
struct ProfRec {
  cudaEvent_t e0, e1;
  int cls;
  double work;
};
static long long g_launches[KC_COUNT] = {0, 0, 0, 0};
static bool g_profiling = false;
static std::vector<ProfRec> g_recs;

LaunchScope::LaunchScope(int cls, void* stream, double work) : slot_(-1), stream_(stream) {
  g_launches[cls]++;
  if (!g_profiling) return;
  ProfRec r;
  r.cls = cls;
  r.work = work;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return;
  cudaEventRecord(r.e0, static_cast<cudaStream_t>(stream));
  g_recs.push_back(r);
  slot_ = (int)g_recs.size() - 1;
}
LaunchScope::~LaunchScope() {
  if (slot_ >= 0) cudaEventRecord(g_recs[slot_].e1, static_cast<cudaStream_t>(stream_));
}

}  // namespace lx

extern "C" {

int64_t lx_launch_count(int32_t cls) {
  if (cls < 0) {
    long long t = 0;
    for (int i = 0; i < lx::KC_COUNT; ++i) t += lx::g_launches[i];
    return t;
  }
  return cls < lx::KC_COUNT ? lx::g_launches[cls] : 0;
}
void lx_launch_count_reset(void) {
  for (int i = 0; i < lx::KC_COUNT; ++i) lx::g_launches[i] = 0;
}
int lx_profile_begin(void) {
  lx::g_recs.clear();
  lx::g_profiling = true;
  return LX_OK;
}
// Synchronises the device, then fills per-class totals: ms[c], launches[c], work[c] (arrays of 4).
int lx_profile_end(double* ms, int64_t* launches, double* work) {
  lx::g_profiling = false;
  LX_CHECK_ARG(ms && launches && work, "lx_profile_end: null output");
  LX_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < lx::KC_COUNT; ++i) { ms[i] = 0; launches[i] = 0; work[i] = 0; }
  for (auto& r : lx::g_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) {
      ms[r.cls] += t;
      launches[r.cls] += 1;
      work[r.cls] += r.work;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  lx::g_recs.clear();
  return LX_OK;
}

void lx_debug_skip(int mask) { lx::g_skip_mask = mask; }
void lx_debug_timeline(long long* device_buffer, int capacity) {
  lx::g_tl_buf = device_buffer;
  lx::g_tl_cap = capacity;
  lx::g_tl_next = 0;
  lx::g_tl_cls.clear();
}
int lx_debug_timeline_count(void) { return lx::g_tl_next; }
int lx_debug_timeline_class(int i) { return i >= 0 && i < (int)lx::g_tl_cls.size() ? lx::g_tl_cls[i] : -1; }

int lx_set_workspace(void* ptr, int64_t bytes, void* stream) {
  if (ptr == nullptr || bytes <= 0) {
    lx::g_ws_set = false;
    lx::g_ws_ptr = nullptr;
    lx::g_ws_bytes = 0;
    return LX_OK;
  }
  LX_CHECK_ARG((reinterpret_cast<uintptr_t>(ptr) & 1023) == 0, "lx_set_workspace: pointer must be 1024-byte aligned");
  LX_CHECK_ARG(bytes >= (int64_t)lx::WS_REGIONS * (lx::WS_FLAG_BYTES + 1024), "lx_set_workspace: %lld bytes is too small",
               (long long)bytes);
  const size_t region = (size_t)(bytes / lx::WS_REGIONS) & ~(size_t)1023;
  // the flag words at the head of every region start (and, by protocol, return to) zero
  for (int i = 0; i < lx::WS_REGIONS; ++i)
    LX_CUDA(cudaMemsetAsync(static_cast<char*>(ptr) + i * region, 0, lx::WS_FLAG_BYTES, static_cast<cudaStream_t>(stream)));
  lx::g_ws_ptr = ptr;
  lx::g_ws_bytes = bytes;
  lx::g_ws_stream = stream;
  lx::g_ws_set = true;
  return LX_OK;
}

const char* lx_last_error(void) { return lx::g_err; }
int lx_version(void) { return 100; }

int lx_device_info(int32_t* out3) {
  LX_CHECK_ARG(out3 != nullptr, "lx_device_info: null output");
  int dev = 0;
  LX_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  LX_CUDA(cudaGetDeviceProperties(&prop, dev));
  out3[0] = prop.multiProcessorCount;
  out3[1] = prop.major;
  out3[2] = prop.minor;
  return LX_OK;
}
}
