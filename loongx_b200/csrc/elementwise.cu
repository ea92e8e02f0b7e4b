// HBM-bound row kernels of the DiT path (coalesced 16-byte loads/stores, one warp per activation row, fp32 math):
//   ln_modulate      LayerNorm(no affine, eps) * (1 + scale) + shift               (block.py:192-207,238-253,301-305)
//   timestep_embed   sinusoidal Timesteps(256, flip_sin_to_cos=True)                (transformer.py:102-114)
//   add_silu         silu(a + b [+ c])  -> operand of every AdaLN modulation GEMM
//   euler_step       FlowMatchEulerDiscreteScheduler.step                            (generate.py:349)
//   rope_table       FluxPosEmbed (float64 internally, like the reference)          (transformer.py:130-134)
//   pack / unpack    FluxPipeline._pack_latents / _unpack_latents (bit-exact index shuffles, generate.py:262,375)
#include <stdlib.h>

#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

constexpr int ROW_WARPS = 8;  // warps (rows) per CTA

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float* x) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void load8_ldg(const __nv_bfloat16* p, float* x) {
  uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  return u;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int LN_WARPS = 4;      // rows per CTA: one warp owns one activation row, no shared memory, no __syncthreads
constexpr int LN_THREADS = LN_WARPS * 32;
constexpr int LN_MAX_NV = 12;    // D <= 3072: each lane owns up to 12 chunks of 8 columns (lane l: chunks l, l+32, ...)

// One warp per row.  All 16-byte loads of the row are issued before the first use (up to 12 per lane = 6 KB per warp in
// flight), the row stays in registers as packed bf16, mean and variance are two register passes with one shuffle
// reduction each, and the modulation vectors (shared by every row of a (stream, batch) pair: L1 hits) are fetched in
// the store pass.  The round-1 version (one 128-thread CTA per row, two block reductions, four __syncthreads) ran at
// 19 % of the HBM peak on L2-resident data; this form is bound by one L2 round trip per row.
__global__ void __launch_bounds__(LN_THREADS) ln_modulate_kernel(const lx_lnmod_desc_t d, long long* tl) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (threadIdx.x == 0) timeline_mark(tl, 0);
  if (row >= d.rows) {  // (whole warp) still take part in the PDL protocol
    pdl_wait();
    pdl_launch_dependents();
    return;
  }
  const lx_tile_meta_t meta = d.tile_meta[row >> 7];  // static table: read ahead of the PDL wait
  const int npl = d.D >> 8;  // 16-byte chunks per lane (D is a multiple of 256)
  const uint4* x = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.x) + (size_t)row * d.ldx);
  // (selects instead of dynamic indexing: kernel-parameter arrays indexed at run time are copied to local memory)
  const int st = meta.stream;
  const void* shift_p = st == 0 ? d.shift[0] : (st == 1 ? d.shift[1] : d.shift[2]);
  const void* scale_p = st == 0 ? d.scale[0] : (st == 1 ? d.scale[1] : d.scale[2]);
  const int64_t mstride = st == 0 ? d.stride[0] : (st == 1 ? d.stride[1] : d.stride[2]);
  const uint4* shift = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(shift_p) + (size_t)meta.batch * mstride);
  const uint4* scale = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(scale_p) + (size_t)meta.batch * mstride);
  pdl_wait();  // PDL: x is the previous kernel's output
  pdl_launch_dependents();  // (releasing the next GEMM's CTAs only at the end of the row was measured: no faster, and
                            //  the GEMM loses the overlap of its prologue / first weight tiles)
  if (threadIdx.x == 0) timeline_mark(tl, 1);
  uint4 raw[LN_MAX_NV];
#pragma unroll
  for (int i = 0; i < LN_MAX_NV; ++i)
    if (i < npl) raw[i] = x[i * 32 + lane];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_NV; ++i) {
    if (i < npl) {
      const float2 a = unpack_bf16(raw[i].x), b = unpack_bf16(raw[i].y), c = unpack_bf16(raw[i].z), e = unpack_bf16(raw[i].w);
      sum += ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (e.x + e.y));
    }
  }
  const float mean = warp_sum(sum) / d.D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_NV; ++i) {
    if (i < npl) {
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(u[e]);
        const float c0 = f.x - mean, c1 = f.y - mean;
        sq += c0 * c0 + c1 * c1;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / d.D + d.eps);
  uint4* out = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(d.out) + (size_t)row * d.ldo);
#pragma unroll
  for (int i = 0; i < LN_MAX_NV; ++i) {
    if (i < npl) {
      const uint4 s4 = __ldg(scale + i * 32 + lane), h4 = __ldg(shift + i * 32 + lane);
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
      const uint32_t su[4] = {s4.x, s4.y, s4.z, s4.w}, hu[4] = {h4.x, h4.y, h4.z, h4.w};
      uint32_t o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = unpack_bf16(u[e]), sc = unpack_bf16(su[e]), sh = unpack_bf16(hu[e]);
        o[e] = pack_bf16((f.x - mean) * rstd * (1.0f + sc.x) + sh.x, (f.y - mean) * rstd * (1.0f + sc.y) + sh.y);
      }
      out[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
  if (threadIdx.x == 0) timeline_mark(tl, 2);  // one mark per CTA: 2560 same-address atomics would dominate the kernel
}

// x[m, 0:D] = bf16(float(x[m]) + float(r[m])): the controlnet residual of transformer.py:172-181, 230-239 on the image rows
__global__ void add_rows_kernel(__nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ r, int64_t ldr,
                                int rows, int D) {
  const int per_row = D >> 3;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 8 elements per thread
  if (idx >= (int64_t)rows * per_row) return;
  const int m = (int)(idx / per_row), c0 = (int)(idx % per_row) * 8;
  float a[8], b[8];
  load8(x + (size_t)m * ldx + c0, a);
  load8_ldg(r + (size_t)m * ldr + c0, b);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] += b[e];
  *reinterpret_cast<uint4*>(x + (size_t)m * ldx + c0) = pack8(a);
}

// out[m, 0:128] = cos(t*f_j), out[m, 128:256] = sin(t*f_j), f_j = exp(-ln(1e4) j / 128)
__global__ void timestep_embed_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int64_t ldo, int M,
                                      float mult) {
  const int m = blockIdx.x;
  const int j = threadIdx.x;  // 128 threads
  if (m >= M) return;
  const float f = expf(-9.210340371976184f * (float)j / 128.0f);
  const float e = t[m] * mult * f;
  out[(size_t)m * ldo + j] = __float2bfloat16_rn(cosf(e));
  out[(size_t)m * ldo + 128 + j] = __float2bfloat16_rn(sinf(e));
}

// out[m] = silu(a[m] + b[m] + c[m % c_rows])  (b, c optional), bf16 -> bf16, row-strided output
__global__ void add_silu_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                const __nv_bfloat16* __restrict__ c, int c_rows, __nv_bfloat16* __restrict__ out,
                                int64_t ldo, int M, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // 8 elements per thread
  const int per_row = D >> 3;
  if (idx >= M * per_row) return;
  const int m = idx / per_row, c0 = (idx % per_row) * 8;
  float x[8], y[8];
  load8(a + (size_t)m * D + c0, x);
  if (b != nullptr) {
    load8(b + (size_t)m * D + c0, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
  }
  if (c != nullptr) {
    load8(c + (size_t)(m % c_rows) * D + c0, y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] = silu(x[e]);
  *reinterpret_cast<uint4*>(out + (size_t)m * ldo + c0) = pack8(x);
}

// x_next = bf16( float(x) + dt * float(v) )  (scheduler.step: fp32 update, cast back to the model dtype)
__global__ void euler_step_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ v,
                                  __nv_bfloat16* __restrict__ out, float dt, int64_t n8) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  float a[8], b[8];
  load8(x + idx * 8, a);
  load8(v + idx * 8, b);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = __fadd_rn(a[e], __fmul_rn(dt, b[e]));  // two roundings, like torch (no FMA)
  *reinterpret_cast<uint4*>(out + idx * 8) = pack8(a);
}

// ids [S,3] fp32 -> table [S, 64, 2] fp32 (cos, sin); axes (d0, d1, d2) with d0+d1+d2 = 128; float64 internally.
__global__ void rope_table_kernel(const float* __restrict__ ids, float* __restrict__ table, int S, int d0, int d1,
                                  int d2, double theta) {
  const int s = blockIdx.x;
  const int p = threadIdx.x;  // rotary pair 0..63
  if (s >= S || p >= 64) return;
  int axis, j, dim;
  if (p < d0 / 2) { axis = 0; j = p; dim = d0; }
  else if (p < (d0 + d1) / 2) { axis = 1; j = p - d0 / 2; dim = d1; }
  else { axis = 2; j = p - (d0 + d1) / 2; dim = d2; }
  const double freq = 1.0 / pow(theta, (double)(2 * j) / (double)dim);
  const double ang = (double)ids[s * 3 + axis] * freq;
  table[((size_t)s * 64 + p) * 2 + 0] = (float)cos(ang);
  table[((size_t)s * 64 + p) * 2 + 1] = (float)sin(ang);
}

// _pack_latents: [B, C, h, w] -> [B, (h/2)(w/2), C*4]   (view(B,C,h/2,2,w/2,2).permute(0,2,4,1,3,5))
template <typename T>
__global__ void pack_latents_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Cc, int h, int w) {
  const int64_t n = (int64_t)B * Cc * h * w;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  // idx enumerates the OUTPUT: (b, i, j, c, di, dj)
  int64_t r = idx;
  const int dj = r % 2; r /= 2;
  const int di = r % 2; r /= 2;
  const int c = r % Cc; r /= Cc;
  const int j = r % (w / 2); r /= (w / 2);
  const int i = r % (h / 2); r /= (h / 2);
  const int b = (int)r;
  out[idx] = in[(((int64_t)b * Cc + c) * h + (2 * i + di)) * w + (2 * j + dj)];
}
// _unpack_latents: [B, (h/2)(w/2), C*4] -> [B, C, h, w]
template <typename T>
__global__ void unpack_latents_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Cc, int h, int w) {
  const int64_t n = (int64_t)B * Cc * h * w;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  // idx enumerates the OUTPUT: (b, c, y, x)
  int64_t r = idx;
  const int x = r % w; r /= w;
  const int y = r % h; r /= h;
  const int c = r % Cc; r /= Cc;
  const int b = (int)r;
  const int i = y / 2, di = y % 2, j = x / 2, dj = x % 2;
  out[idx] = in[(((((int64_t)b * (h / 2) + i) * (w / 2) + j) * Cc + c) * 2 + di) * 2 + dj];
}

}  // namespace lx

using namespace lx;

extern "C" int lx_ln_modulate(const lx_lnmod_desc_t* desc, void* stream) {
  LX_CHECK_ARG(desc != nullptr, "lx_ln_modulate: null descriptor");
  const lx_lnmod_desc_t& d = *desc;
  LX_CHECK_ARG(d.rows > 0 && d.D > 0 && d.D % 256 == 0 && d.D <= LN_MAX_NV * 256,
               "lx_ln_modulate: D=%d must be a multiple of 256 and <= %d", d.D, LN_MAX_NV * 256);
  LX_CHECK_ARG(d.x && d.out && d.tile_meta, "lx_ln_modulate: null pointer");
  LX_CHECK_ARG(d.ldx % 8 == 0 && d.ldo % 8 == 0 && d.ldo >= d.D && d.ldx >= d.D, "lx_ln_modulate: bad strides");
  if (lx::debug_skip_mask() & 4) return LX_OK;  // timing experiments only (lx_debug_skip): marginal cost of this kernel
  LaunchScope scope(KC_ROW, stream, 4.0 * d.rows * d.D);  // bytes: read + write bf16 rows
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((d.rows + LN_WARPS - 1) / LN_WARPS);
  cfg.blockDim = dim3(LN_THREADS);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  long long* tl = timeline_next(KC_ROW);
  cudaLaunchKernelEx(&cfg, ln_modulate_kernel, d, tl);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_timestep_embed(const float* t, void* out, int64_t ldo, int32_t M, float mult, void* stream) {
  LX_CHECK_ARG(t && out && M > 0 && ldo >= 256, "lx_timestep_embed: bad arguments");
  LaunchScope scope(KC_ROW, stream, 512.0 * M);
  timestep_embed_kernel<<<M, 128, 0, static_cast<cudaStream_t>(stream)>>>(t, reinterpret_cast<__nv_bfloat16*>(out), ldo,
                                                                         M, mult);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_add_silu_bcast(const void* a, const void* b, const void* c, int32_t c_rows, void* out, int64_t ldo,
                                 int32_t M, int32_t D, void* stream) {
  LX_CHECK_ARG(a && out && M > 0 && D > 0 && D % 8 == 0 && ldo % 8 == 0 && ldo >= D, "lx_add_silu_bcast: bad arguments");
  LX_CHECK_ARG(c == nullptr || c_rows > 0, "lx_add_silu_bcast: c_rows must be positive");
  const int n = M * (D / 8);
  LaunchScope scope(KC_ROW, stream, 8.0 * M * D);
  add_silu_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(a), reinterpret_cast<const __nv_bfloat16*>(b),
      reinterpret_cast<const __nv_bfloat16*>(c), c_rows, reinterpret_cast<__nv_bfloat16*>(out), ldo, M, D);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_add_rows(void* x, int64_t ldx, const void* r, int64_t ldr, int32_t rows, int32_t D, void* stream) {
  LX_CHECK_ARG(x && r && rows > 0 && D > 0 && D % 8 == 0 && ldx % 8 == 0 && ldr % 8 == 0 && ldx >= D && ldr >= D,
               "lx_add_rows: bad arguments (rows=%d D=%d)", rows, D);
  const int64_t n = (int64_t)rows * (D / 8);
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * D);
  add_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(x), ldx, reinterpret_cast<const __nv_bfloat16*>(r), ldr, rows, D);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_euler_step(const void* x, const void* v, void* out, float dt, int64_t n, void* stream) {
  LX_CHECK_ARG(x && v && out && n > 0 && n % 8 == 0, "lx_euler_step: n=%lld must be a positive multiple of 8",
               (long long)n);
  const int64_t n8 = n / 8;
  LaunchScope scope(KC_ROW, stream, 6.0 * n);
  euler_step_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(v),
      reinterpret_cast<__nv_bfloat16*>(out), dt, n8);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_rope_table(const float* ids, float* table, int32_t S, int32_t d0, int32_t d1, int32_t d2, double theta,
                             void* stream) {
  LX_CHECK_ARG(ids && table && S > 0, "lx_rope_table: bad arguments");
  LX_CHECK_ARG(d0 + d1 + d2 == 128 && d0 % 2 == 0 && d1 % 2 == 0 && d2 % 2 == 0,
               "lx_rope_table: axes dims must be even and sum to 128");
  LaunchScope scope(KC_ROW, stream, 512.0 * S);
  rope_table_kernel<<<S, 64, 0, static_cast<cudaStream_t>(stream)>>>(ids, table, S, d0, d1, d2, theta);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_pack_latents(const void* in, void* out, int32_t B, int32_t C, int32_t h, int32_t w, int32_t elem_bytes,
                               int32_t unpack, void* stream) {
  LX_CHECK_ARG(in && out && B > 0 && C > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0, "lx_pack_latents: bad shape");
  LX_CHECK_ARG(elem_bytes == 2 || elem_bytes == 4, "lx_pack_latents: elem_bytes must be 2 or 4");
  const int64_t n = (int64_t)B * C * h * w;
  const unsigned grid = (unsigned)((n + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LaunchScope scope(KC_ROW, stream, 2.0 * n * elem_bytes);
  if (elem_bytes == 2) {
    if (unpack) unpack_latents_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, B, C, h, w);
    else pack_latents_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)in, (uint16_t*)out, B, C, h, w);
  } else {
    if (unpack) unpack_latents_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, B, C, h, w);
    else pack_latents_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, B, C, h, w);
  }
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
