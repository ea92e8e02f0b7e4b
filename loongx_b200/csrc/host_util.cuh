// Host-side helpers shared by all translation units: error reporting and TMA tensor-map encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/loongx_b200.h"

namespace lx {

void set_error(const char* fmt, ...);

#define LX_CHECK_ARG(cond, ...)   \
  do {                            \
    if (!(cond)) {                \
      lx::set_error(__VA_ARGS__); \
      return LX_ERR_ARG;          \
    }                             \
  } while (0)

#define LX_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      lx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));        \
      return LX_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

// 2-D bf16 tensor map: `rows` x `cols` (cols contiguous), row stride `ld` elements, box = box_rows x box_cols,
// 128-byte swizzle (box_cols must be 64 bf16 = 128 B).  Out-of-bounds elements are zero-filled on load and
// clipped on store.
int make_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint32_t box_rows, uint32_t box_cols);
int make_tmap_2d_f32(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                     uint32_t box_cols);
// 3-D variant: dims (cols, rows, outer) with strides (ld, outer_stride) in elements.
int make_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t outer, uint64_t rows, uint64_t cols, uint64_t ld,
                      uint64_t outer_stride, uint32_t box_rows, uint32_t box_cols);

int num_sms();
// Exchange workspace registered with lx_set_workspace: region `which` (0 = attention, 1 = GEMM) if the workspace is bound
// to `stream` and the region holds `need_bytes`, else nullptr.  Every region starts with WS_FLAG_BYTES of flag words that
// are zero between launches.
constexpr int WS_REGIONS = 2;
constexpr int WS_FLAG_BYTES = 4096;
void* workspace_region(void* stream, int which, size_t need_bytes);
// Programmatic dependent launch switch (default on; lx_debug_set_pdl(0) turns it off for A/B timing).
bool pdl_enabled();

// Launch accounting + optional per-kernel-class CUDA-event timing (bench.py's roofline numbers): every extern "C"
// launcher opens a LaunchScope; with profiling off it only bumps a counter.
enum KernelClass { KC_GEMM = 0, KC_ATTENTION = 1, KC_ROW = 2, KC_CS3DGF = 3, KC_COUNT = 4 };
// Kernel launch with the programmatic-dependent-launch attribute (when enabled).  The kernel MUST call pdl_wait() before
// its first global-memory access (ptx.cuh).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Plain stream-ordered launch (no programmatic overlap with the predecessor).  Used, defensively, for the first kernel
// after a cudaMemsetAsync whose target the kernel accumulates into.  (A first version of the VAE's GroupNorm statistics
// - memset + atomics under programmatic launch - showed percent-level run-to-run noise; scripts/pdl_order_probe.cu did
// NOT reproduce a memset / kernel reordering in isolation, so the cause is not established.  The VAE now has no memset
// and no atomics and is bit-reproducible; this launcher keeps the remaining memset + accumulate site on plain ordering.)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_ordered(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Development aid: device-side timeline of the three kernels of the DiT loop.  lx_debug_timeline(buf, capacity) hands the
// library a device buffer of [capacity][4] int64 = (first CTA entry, first CTA past its programmatic-dependency wait, last
// CTA exit, unused) in %globaltimer ns, pre-filled by the caller with (INT64_MAX, INT64_MAX, 0, 0); every launch
// of lx_gemm_bf16 / lx_attention / lx_ln_modulate takes the next row.  NULL switches it off (the default).
long long* timeline_next(int cls);
// lx_debug_skip(mask): bit (1 << kernel class) set = that class's launcher returns without launching (timing experiments:
// the marginal cost of a kernel inside the real loop, with the buffers keeping their last realistic contents)
int debug_skip_mask();

// radix-2 FFT of d channels of L (power of two) values (cs3_dgf.cu): S4 kernel generation and its backward
int s4_fft_launch(bool inverse, const double2* in_c, const float* in_r, double2* out_c, float* out_r, int d, int L,
                  void* stream);

// 128 x 128 x 16 fp32 SIMT GEMM tile for the larger products of the CS3 / DGF path (cs3_dgf.cu; shared by the forward
// lx_sgemm_f32 and the backward lx_sgemm_ex): C[b] = epilogue(alpha op(A[b]) op(B[b])), op = stored [M,K] / [K,N] or
// transposed; reduce mode adds (batch element, K slice) partials into one C with fp32 atomics.
struct Sgemm128 {
  const float* A;
  const float* B;
  float* C;
  const float* bias;  // [M] added per output row, or NULL
  const float* R;     // residual with C's layout, or NULL
  int64_t lda, ldb, ldc, a_bs, b_bs, c_bs;
  int M, N, K;
  int trans_a, trans_b, act;  // act: 0 none, 1 relu, 2 sigmoid
  int reduce_batch, ksplit;
  float alpha, beta;
};
int sgemm128_launch(const Sgemm128& p, int batch, void* stream);

struct LaunchScope {
  LaunchScope(int cls, void* stream, double work);  // work: FLOPs (tensor kernels) or bytes (HBM-bound kernels)
  ~LaunchScope();
  int slot_;
  void* stream_;
};

}  // namespace lx
