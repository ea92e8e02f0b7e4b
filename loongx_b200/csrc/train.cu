// Training-step kernels (OminiModel.step, model.py:569-729): the forward pieces that keep the intermediates the
// backward needs, and the backward row kernels of the DiT blocks (block.py:179-339 differentiated by hand).
//
// All activations / gradients are bf16 rows in the stream-major layout [txt(B*Nt) | img(B*Ni) | cond(B*Nc)]
// (lx_tile_meta_t maps a 128-row tile to its (stream, batch)); arithmetic is fp32 with one rounding on store;
// reductions over rows (gradients of the AdaLN modulation vectors, LoRA factor gradients) accumulate in fp32 with
// one atomicAdd per (128-row tile, column).  HBM-bound: every kernel reads / writes each operand once with
// 16-byte (row kernels) or 4-byte-per-lane coalesced (column-reduction kernels) accesses.
#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

struct Vec3 {
  const __nv_bfloat16* p[3];
  int64_t stride[3];
};
struct Acc3 {
  float* p[3];
  int64_t stride[3];
};

__device__ __forceinline__ void ld8(const __nv_bfloat16* p, float* x) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void st8(__nv_bfloat16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// d/dx of 0.5 x (1 + tanh(k0 (x + k1 x^3)))
__device__ __forceinline__ float gelu_tanh_grad(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u = k0 * (x + k1 * x * x * x);
  const float t = tanhf(u);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * x * x);
}

// ---------------------------------------------------------------------------------------------------------------
// elementwise: GELU forward / backward on [rows, cols] views (8 columns per thread)
// ---------------------------------------------------------------------------------------------------------------
__global__ void gelu_fwd_kernel(const __nv_bfloat16* __restrict__ pre, int64_t ld_pre, __nv_bfloat16* __restrict__ out,
                                int64_t ldo, int rows, int cols8) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)rows * cols8) return;
  const int r = (int)(idx / cols8), c = (int)(idx % cols8) * 8;
  float v[8];
  ld8(pre + (size_t)r * ld_pre + c, v);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = gelu_tanh(v[e]);
  st8(out + (size_t)r * ldo + c, v);
}

__global__ void gelu_bwd_kernel(const __nv_bfloat16* __restrict__ pre, int64_t ld_pre, const __nv_bfloat16* dy,
                                int64_t ld_dy, __nv_bfloat16* dx /* may alias dy */, int64_t ld_dx, int rows, int cols8) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)rows * cols8) return;
  const int r = (int)(idx / cols8), c = (int)(idx % cols8) * 8;
  float v[8], g[8];
  ld8(pre + (size_t)r * ld_pre + c, v);
  ld8(dy + (size_t)r * ld_dy + c, g);
#pragma unroll
  for (int e = 0; e < 8; ++e) g[e] *= gelu_tanh_grad(v[e]);
  st8(dx + (size_t)r * ld_dx + c, g);
}

// ---------------------------------------------------------------------------------------------------------------
// out = res + gate[stream, batch] * y     (block.py:224-234, 268-274, 328-334 un-fused from the GEMM epilogue so that
// y survives for the gate gradient)
// ---------------------------------------------------------------------------------------------------------------
__global__ void gate_residual_fwd_kernel(const __nv_bfloat16* res, const __nv_bfloat16* __restrict__ y,
                                         __nv_bfloat16* out /* may alias res (element-local) */, int64_t ld, int rows, int D8,
                                         const lx_tile_meta_t* __restrict__ tm, Vec3 gate) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)rows * D8) return;
  const int r = (int)(idx / D8), c = (int)(idx % D8) * 8;
  const lx_tile_meta_t m = tm[r >> 7];
  float a[8], b[8], g[8];
  ld8(res + (size_t)r * ld + c, a);
  ld8(y + (size_t)r * ld + c, b);
  ld8(gate.p[m.stream] + (size_t)m.batch * gate.stride[m.stream] + c, g);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] += g[e] * b[e];
  st8(out + (size_t)r * ld + c, a);
}

// dy = gate * dout ; dgate[stream, batch, col] += sum_rows dout * y.   One CTA = one 128-row tile x 256 columns; each
// thread owns 2 adjacent columns and walks the 128 rows (a warp reads 128 contiguous bytes per row).
__global__ void __launch_bounds__(128) gate_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                       const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dy,
                                                       int64_t ld, int rows, int D, const lx_tile_meta_t* __restrict__ tm,
                                                       Vec3 gate, Acc3 dgate) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int tile = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x * 2;
  if (c >= D) return;
  const lx_tile_meta_t m = tm[tile];
  const float2 g = unpack_bf16(*reinterpret_cast<const uint32_t*>(gate.p[m.stream] + (size_t)m.batch * gate.stride[m.stream] + c));
  float ax = 0.f, ay = 0.f;
  const int r0 = tile * 128, r1 = min(rows, r0 + 128);
  float* acc = dgate.p[m.stream];
  if (acc != nullptr) {
    for (int r = r0; r < r1; ++r) {
      const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(dout + (size_t)r * ld + c));
      const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(y + (size_t)r * ld + c));
      ax += d.x * v.x;
      ay += d.y * v.y;
      *reinterpret_cast<uint32_t*>(dy + (size_t)r * ld + c) = pack_bf16(g.x * d.x, g.y * d.y);
    }
  } else {  // no gate gradient wanted for this stream: y is not read at all (it may not even have been recomputed)
    for (int r = r0; r < r1; ++r) {
      const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(dout + (size_t)r * ld + c));
      *reinterpret_cast<uint32_t*>(dy + (size_t)r * ld + c) = pack_bf16(g.x * d.x, g.y * d.y);
    }
  }
  if (acc != nullptr) {
    acc += (size_t)m.batch * dgate.stride[m.stream] + c;
    atomicAdd(acc, ax);
    atomicAdd(acc + 1, ay);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// backward of  xn = LayerNorm(x) * (1 + scale) + shift   (AdaLayerNormZero / -Single / norm2 + FiLM / norm_out)
//   row kernel:     dx = dres + rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dxn * (1 + scale);  stats[row] = (mean, rstd)
//   column kernel:  dscale[stream,batch,col] += sum_rows dxn * xhat ;  dshift += sum_rows dxn
// ---------------------------------------------------------------------------------------------------------------
constexpr int LNB_THREADS = 128;
constexpr int LNB_MAX_NV = 3;  // D <= 3072

__device__ __forceinline__ float block_sum128(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const float r = red[0] + red[1] + red[2] + red[3];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(LNB_THREADS) ln_mod_bwd_row_kernel(
    const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dxn, const __nv_bfloat16* dres,
    __nv_bfloat16* dx /* may alias dres (row-local) */, int64_t ld, int D, const lx_tile_meta_t* __restrict__ tm, Vec3 scale, float eps,
    float2* __restrict__ stats) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float red[4];
  const int row = blockIdx.x;
  const lx_tile_meta_t m = tm[row >> 7];
  const int nchunk = D >> 3;
  const __nv_bfloat16* sc = scale.p[m.stream] + (size_t)m.batch * scale.stride[m.stream];
  float v[LNB_MAX_NV * 8], g[LNB_MAX_NV * 8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LNB_MAX_NV; ++i) {
    const int ch = i * LNB_THREADS + threadIdx.x;
    if (ch < nchunk) {
      float s[8];
      ld8(x + (size_t)row * ld + ch * 8, &v[i * 8]);
      ld8(dxn + (size_t)row * ld + ch * 8, &g[i * 8]);
      ld8(sc + ch * 8, s);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        sum += v[i * 8 + e];
        g[i * 8 + e] *= 1.0f + s[e];
      }
    }
  }
  const float mean = block_sum128(sum, red) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LNB_MAX_NV; ++i)
    if (i * LNB_THREADS + threadIdx.x < nchunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[i * 8 + e] -= mean;
        sq += v[i * 8 + e] * v[i * 8 + e];
      }
    }
  const float rstd = rsqrtf(block_sum128(sq, red) / D + eps);
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < LNB_MAX_NV; ++i)
    if (i * LNB_THREADS + threadIdx.x < nchunk) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        v[i * 8 + e] *= rstd;  // xhat
        sg += g[i * 8 + e];
        sgx += g[i * 8 + e] * v[i * 8 + e];
      }
    }
  const float mg = block_sum128(sg, red) / D;
  const float mgx = block_sum128(sgx, red) / D;
#pragma unroll
  for (int i = 0; i < LNB_MAX_NV; ++i) {
    const int ch = i * LNB_THREADS + threadIdx.x;
    if (ch < nchunk) {
      float o[8];
      if (dres != nullptr) ld8(dres + (size_t)row * ld + ch * 8, o);
      else {
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = 0.f;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += rstd * (g[i * 8 + e] - mg - v[i * 8 + e] * mgx);
      st8(dx + (size_t)row * ld + ch * 8, o);
    }
  }
  if (threadIdx.x == 0 && stats != nullptr) stats[row] = make_float2(mean, rstd);
}

__global__ void __launch_bounds__(128) ln_mod_bwd_col_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ dxn, int64_t ld, int rows,
                                                             int D, const lx_tile_meta_t* __restrict__ tm,
                                                             const float2* __restrict__ stats, Acc3 dscale, Acc3 dshift) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int tile = blockIdx.x, c = blockIdx.y * 256 + threadIdx.x * 2;
  if (c >= D) return;
  const lx_tile_meta_t m = tm[tile];
  if (dscale.p[m.stream] == nullptr && dshift.p[m.stream] == nullptr) return;
  float sx = 0.f, sy = 0.f, hx = 0.f, hy = 0.f;
  const int r0 = tile * 128, r1 = min(rows, r0 + 128);
  for (int r = r0; r < r1; ++r) {
    const float2 st = stats[r];
    const float2 v = unpack_bf16(*reinterpret_cast<const uint32_t*>(x + (size_t)r * ld + c));
    const float2 d = unpack_bf16(*reinterpret_cast<const uint32_t*>(dxn + (size_t)r * ld + c));
    sx += d.x * (v.x - st.x) * st.y;
    sy += d.y * (v.y - st.x) * st.y;
    hx += d.x;
    hy += d.y;
  }
  if (dscale.p[m.stream] != nullptr) {
    float* a = dscale.p[m.stream] + (size_t)m.batch * dscale.stride[m.stream] + c;
    atomicAdd(a, sx);
    atomicAdd(a + 1, sy);
  }
  if (dshift.p[m.stream] != nullptr) {
    float* a = dshift.p[m.stream] + (size_t)m.batch * dshift.stride[m.stream] + c;
    atomicAdd(a, hx);
    atomicAdd(a + 1, hy);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// q/k/v post-processing un-fused from the GEMM epilogue: per-head RMSNorm(q, k) * w, RoPE, scatter to [B,H,S,128]
// (block.py:34-41, 60-67, 74-99) and its backward.  One warp per (row, head): lane owns 4 adjacent elements = 2 rotary
// pairs.
// ---------------------------------------------------------------------------------------------------------------
struct RmsW {
  const float* q[3];
  const float* k[3];
};

__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float* x) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float* v) {
  uint2 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

__global__ void __launch_bounds__(128) qkv_post_fwd_kernel(const __nv_bfloat16* __restrict__ pre, int64_t ld, int rows,
                                                           int heads, const lx_tile_meta_t* __restrict__ tm,
                                                           __nv_bfloat16* __restrict__ q, __nv_bfloat16* __restrict__ k,
                                                           __nv_bfloat16* __restrict__ v, int seq_total, RmsW w,
                                                           const float* __restrict__ rope, float eps) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const lx_tile_meta_t m = tm[row >> 7];
  const int seq = m.seq_row + (row & 127);  // batch * S + position in the joint sequence
  const int b = seq / seq_total, s = seq % seq_total;
  const int D = heads * 128;
  float cs[4] = {1.f, 0.f, 1.f, 0.f};
  if (rope != nullptr) {
    const float4 t = *reinterpret_cast<const float4*>(rope + ((size_t)s * 64 + lane * 2) * 2);
    cs[0] = t.x; cs[1] = t.y; cs[2] = t.z; cs[3] = t.w;
  }
  for (int h = warp; h < heads; h += 4) {
    const size_t dst = (((size_t)b * heads + h) * seq_total + s) * 128 + lane * 4;
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      float x[4];
      ld4(pre + (size_t)row * ld + which * D + h * 128 + lane * 4, x);
      if (which == 2) {
        st4(v + dst, x);
        continue;
      }
      const float* wt = which == 0 ? w.q[m.stream] : w.k[m.stream];
      if (wt != nullptr) {
        const float ss = warp_sum(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
        const float r = rsqrtf(ss * (1.0f / 128.0f) + eps);
        const float4 ww = *reinterpret_cast<const float4*>(wt + lane * 4);
        x[0] *= r * ww.x; x[1] *= r * ww.y; x[2] *= r * ww.z; x[3] *= r * ww.w;
      }
      float o[4];
      o[0] = x[0] * cs[0] - x[1] * cs[1];
      o[1] = x[1] * cs[0] + x[0] * cs[1];
      o[2] = x[2] * cs[2] - x[3] * cs[3];
      o[3] = x[3] * cs[2] + x[2] * cs[3];
      st4((which == 0 ? q : k) + dst, o);
    }
  }
}

__global__ void __launch_bounds__(128) qkv_post_bwd_kernel(const __nv_bfloat16* __restrict__ pre, int64_t ld,
                                                           const __nv_bfloat16* __restrict__ dq,
                                                           const __nv_bfloat16* __restrict__ dk,
                                                           const __nv_bfloat16* __restrict__ dv,
                                                           __nv_bfloat16* __restrict__ dpre, int64_t ldo, int rows, int heads,
                                                           const lx_tile_meta_t* __restrict__ tm, int seq_total, RmsW w,
                                                           const float* __restrict__ rope, float eps) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const lx_tile_meta_t m = tm[row >> 7];
  const int seq = m.seq_row + (row & 127);
  const int b = seq / seq_total, s = seq % seq_total;
  const int D = heads * 128;
  float cs[4] = {1.f, 0.f, 1.f, 0.f};
  if (rope != nullptr) {
    const float4 t = *reinterpret_cast<const float4*>(rope + ((size_t)s * 64 + lane * 2) * 2);
    cs[0] = t.x; cs[1] = t.y; cs[2] = t.z; cs[3] = t.w;
  }
  for (int h = warp; h < heads; h += 4) {
    const size_t src = (((size_t)b * heads + h) * seq_total + s) * 128 + lane * 4;
#pragma unroll
    for (int which = 0; which < 3; ++which) {
      __nv_bfloat16* out = dpre + (size_t)row * ldo + which * D + h * 128 + lane * 4;
      float g[4];
      ld4((which == 0 ? dq : which == 1 ? dk : dv) + src, g);
      if (which == 2) {
        st4(out, g);
        continue;
      }
      // RoPE^T
      float t[4];
      t[0] = g[0] * cs[0] + g[1] * cs[1];
      t[1] = g[1] * cs[0] - g[0] * cs[1];
      t[2] = g[2] * cs[2] + g[3] * cs[3];
      t[3] = g[3] * cs[2] - g[2] * cs[3];
      const float* wt = which == 0 ? w.q[m.stream] : w.k[m.stream];
      if (wt != nullptr) {
        float x[4];
        ld4(pre + (size_t)row * ld + which * D + h * 128 + lane * 4, x);
        const float4 ww = *reinterpret_cast<const float4*>(wt + lane * 4);
        t[0] *= ww.x; t[1] *= ww.y; t[2] *= ww.z; t[3] *= ww.w;
        const float ss = warp_sum(x[0] * x[0] + x[1] * x[1] + x[2] * x[2] + x[3] * x[3]);
        const float r = rsqrtf(ss * (1.0f / 128.0f) + eps);
        const float dot = warp_sum(t[0] * x[0] + t[1] * x[1] + t[2] * x[2] + t[3] * x[3]);
        const float cfac = r * r * r * dot * (1.0f / 128.0f);
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = r * t[e] - x[e] * cfac;
      }
      st4(out, t);
    }
  }
}

// rows [R, ld] (head h in columns [128h, 128h+128)) -> [B, H, S, 128]
__global__ void __launch_bounds__(128) rows_to_heads_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld, int rows,
                                                            int heads, const lx_tile_meta_t* __restrict__ tm,
                                                            __nv_bfloat16* __restrict__ out, int seq_total) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int row = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const lx_tile_meta_t m = tm[row >> 7];
  const int seq = m.seq_row + (row & 127);
  const int b = seq / seq_total, s = seq % seq_total;
  for (int h = warp; h < heads; h += 4) {
    const uint2 u = *reinterpret_cast<const uint2*>(in + (size_t)row * ld + h * 128 + lane * 4);
    *reinterpret_cast<uint2*>(out + (((size_t)b * heads + h) * seq_total + s) * 128 + lane * 4) = u;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// LoRA factor gradients (peft LoRA Linear, y = x W^T + s (x A^T) B^T):
//   dA[r, K] += s (dy B)^T x        dB[N, r] += s dy^T (x A^T)
// kernel 1: P[m, 0:r] = sum_c Z[m, c] * F[c, j]           (Z = dy, F = B  or  Z = x, F = A^T)   one warp per row
// kernel 2: G[c, j]  += s * sum_m Z[m, c] * P[m, j]       (128-row chunk x 256 columns per CTA, atomics)
// ---------------------------------------------------------------------------------------------------------------
constexpr int LORA_MAX_R = 16;
constexpr int LORA_CCHUNK = 2048;  // columns per CTA of the projection kernel (fills the GPU when M is small)

// RPW rows per warp: the factor values F[c, 0:r] are loaded once per lane and reused across the rows, so the kernel
// streams Z (bf16, 4 bytes per lane per row, RPW independent rows in flight) instead of re-reading F for every row.
template <int RMAX, int RPW>
__global__ void __launch_bounds__(128) lora_project_kernel(const __nv_bfloat16* __restrict__ Z, int64_t ldz, int M, int C,
                                                           const float* __restrict__ F, int64_t f_stride_c,
                                                           int64_t f_stride_j, int r, float* __restrict__ P) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int row0 = (blockIdx.x * 4 + (threadIdx.x >> 5)) * RPW, lane = threadIdx.x & 31;
  if (row0 >= M) return;
  const int c_begin = blockIdx.y * LORA_CCHUNK, c_end = min(C, c_begin + LORA_CCHUNK);  // column split: partial sums
  float acc[RPW][RMAX];
#pragma unroll
  for (int i = 0; i < RPW; ++i)
#pragma unroll
    for (int j = 0; j < RMAX; ++j) acc[i][j] = 0.f;
  for (int c = c_begin + lane * 2; c < c_end; c += 64) {
    float f0[RMAX], f1[RMAX];
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      f0[j] = j < r ? F[(size_t)c * f_stride_c + j * f_stride_j] : 0.f;
      f1[j] = j < r ? F[(size_t)(c + 1) * f_stride_c + j * f_stride_j] : 0.f;
    }
    float2 z[RPW];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int row = min(row0 + i, M - 1);
      z[i] = unpack_bf16(*reinterpret_cast<const uint32_t*>(Z + (size_t)row * ldz + c));
    }
#pragma unroll
    for (int i = 0; i < RPW; ++i)
#pragma unroll
      for (int j = 0; j < RMAX; ++j) acc[i][j] += z[i].x * f0[j] + z[i].y * f1[j];
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i)
#pragma unroll
    for (int j = 0; j < RMAX; ++j) {
      const float sm = warp_sum(acc[i][j]);
      if (lane == 0 && j < r && row0 + i < M) atomicAdd(P + (size_t)(row0 + i) * r + j, sm);  // P zeroed by the launcher
    }
}

// G[c, j] += s * sum_m Z[m, c] P[m, j]: one CTA = 128 rows x 512 columns, 4 adjacent columns (8 bytes) per thread,
// four rows in flight per thread; P rows are broadcast from shared memory.
template <int RMAX>
__global__ void __launch_bounds__(128) lora_reduce_kernel(const __nv_bfloat16* __restrict__ Z, int64_t ldz, int M, int C,
                                                          const float* __restrict__ P, int r, float scaling,
                                                          float* __restrict__ G, int64_t g_stride_c, int64_t g_stride_j) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float ps[128 * RMAX];
  const int r0 = blockIdx.x * 128, r1 = min(M, r0 + 128);
  for (int i = threadIdx.x; i < 128 * RMAX; i += 128) {
    const int m = i / RMAX, j = i % RMAX;
    ps[i] = (r0 + m < r1 && j < r) ? P[(size_t)(r0 + m) * r + j] : 0.f;
  }
  __syncthreads();
  const int c = blockIdx.y * 512 + threadIdx.x * 4;
  if (c >= C) return;
  float acc[4][RMAX];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int j = 0; j < RMAX; ++j) acc[e][j] = 0.f;
  const int nrow = r1 - r0;
  for (int m = 0; m < nrow; m += 4) {
    uint2 u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mm = min(m + i, nrow - 1);  // rows past the end re-read the last row and multiply by ps = 0
      u[i] = *reinterpret_cast<const uint2*>(Z + (size_t)(r0 + mm) * ldz + c);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 a = unpack_bf16(u[i].x), b = unpack_bf16(u[i].y);
      const float* pr = ps + (m + i < nrow ? m + i : 127) * RMAX;
      const float valid = (m + i < nrow) ? 1.f : 0.f;
#pragma unroll
      for (int j = 0; j < RMAX; ++j) {
        const float pv = pr[j] * valid;
        acc[0][j] += a.x * pv;
        acc[1][j] += a.y * pv;
        acc[2][j] += b.x * pv;
        acc[3][j] += b.y * pv;
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int j = 0; j < RMAX; ++j)
      if (j < r) atomicAdd(G + (size_t)(c + e) * g_stride_c + j * g_stride_j, scaling * acc[e][j]);
}

// out[n, k] = bf16(W[n, k] + s * sum_j B[n, j] A[j, k])   (the merged panel the LoRA-active row group multiplies by)
__global__ void lora_merge_kernel(const __nv_bfloat16* __restrict__ W, int64_t ldw, const float* __restrict__ A,
                                  const float* __restrict__ Bw, __nv_bfloat16* __restrict__ out, int64_t ldo, int N, int K,
                                  int r, float scaling) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int K2 = K >> 1;
  if (idx >= (int64_t)N * K2) return;
  const int n = (int)(idx / K2), k = (int)(idx % K2) * 2;
  const float2 w = unpack_bf16(*reinterpret_cast<const uint32_t*>(W + (size_t)n * ldw + k));
  float ax = 0.f, ay = 0.f;
  for (int j = 0; j < r; ++j) {
    const float b = Bw[(size_t)n * r + j];
    ax += b * A[(size_t)j * K + k];
    ay += b * A[(size_t)j * K + k + 1];
  }
  *reinterpret_cast<uint32_t*>(out + (size_t)n * ldo + k) = pack_bf16(w.x + scaling * ax, w.y + scaling * ay);
}

// out[c, r] = in[r, c]   (32x32 shared-memory tiles; builds the K-major W^T panels the dX GEMMs multiply by)
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t ld_in, __nv_bfloat16* __restrict__ out,
                                      int64_t ld_out, int rows, int cols) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ __nv_bfloat16 t[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) t[i][threadIdx.x] = in[(size_t)r * ld_in + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * ld_out + r] = t[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------------------------
// rectified-flow objective (model.py:590-594, 726-727)
// ---------------------------------------------------------------------------------------------------------------
__global__ void flow_noise_mix_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ x1,
                                      const float* __restrict__ t, __nv_bfloat16* __restrict__ xt, int64_t per_sample8,
                                      int64_t n8) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n8) return;
  const float tt = t[idx / per_sample8];
  float a[8], b[8];
  ld8(x0 + idx * 8, a);
  ld8(x1 + idx * 8, b);
#pragma unroll
  for (int e = 0; e < 8; ++e) a[e] = (1.0f - tt) * a[e] + tt * b[e];
  st8(xt + idx * 8, a);
}

__global__ void __launch_bounds__(256) flow_mse_kernel(const __nv_bfloat16* __restrict__ pred,
                                                       const __nv_bfloat16* __restrict__ x0,
                                                       const __nv_bfloat16* __restrict__ x1, float* __restrict__ loss,
                                                       __nv_bfloat16* __restrict__ dpred, int64_t n8, float inv_n,
                                                       float grad_scale) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float red[8];
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float s = 0.f;
  if (idx < n8) {
    float p[8], a[8], b[8];
    ld8(pred + idx * 8, p);
    ld8(x0 + idx * 8, a);
    ld8(x1 + idx * 8, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      // the reference subtracts in the model dtype: (x_1 - x_0) is a bf16 tensor (model.py:726)
      const float target = __bfloat162float(__float2bfloat16_rn(b[e] - a[e]));
      const float d = p[e] - target;
      s += d * d;
      p[e] = grad_scale * 2.0f * d * inv_n;
    }
    if (dpred != nullptr) st8(dpred + idx * 8, p);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += red[i];
    atomicAdd(loss, tot * inv_n);
  }
}

}  // namespace lx

using namespace lx;

namespace {
inline cudaStream_t cs(void* s) { return static_cast<cudaStream_t>(s); }
inline const __nv_bfloat16* bf(const void* p) { return reinterpret_cast<const __nv_bfloat16*>(p); }
inline __nv_bfloat16* bf(void* p) { return reinterpret_cast<__nv_bfloat16*>(p); }
Vec3 vec3(const void* const* p, const int64_t* stride) {
  Vec3 v;
  for (int i = 0; i < 3; ++i) { v.p[i] = bf(p[i]); v.stride[i] = stride[i]; }
  return v;
}
Acc3 acc3(float* const* p, const int64_t* stride) {
  Acc3 v;
  for (int i = 0; i < 3; ++i) { v.p[i] = p ? p[i] : nullptr; v.stride[i] = stride ? stride[i] : 0; }
  return v;
}
}  // namespace

extern "C" int lx_gelu_fwd(const void* pre, int64_t ld_pre, void* out, int64_t ldo, int32_t rows, int32_t cols,
                           void* stream) {
  LX_CHECK_ARG(pre && out && rows > 0 && cols > 0 && cols % 8 == 0 && ld_pre % 8 == 0 && ldo % 8 == 0, "lx_gelu_fwd: bad arguments");
  const int64_t n = (int64_t)rows * (cols / 8);
  LaunchScope scope(KC_ROW, stream, 4.0 * rows * cols);
  launch_pdl(gelu_fwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, cs(stream), bf(pre), ld_pre, bf(out), ldo, rows, cols / 8);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_gelu_bwd(const void* pre, int64_t ld_pre, const void* dy, int64_t ld_dy, void* dx, int64_t ld_dx,
                           int32_t rows, int32_t cols, void* stream) {
  LX_CHECK_ARG(pre && dy && dx && rows > 0 && cols > 0 && cols % 8 == 0 && ld_pre % 8 == 0 && ld_dy % 8 == 0 && ld_dx % 8 == 0,
               "lx_gelu_bwd: bad arguments");
  const int64_t n = (int64_t)rows * (cols / 8);
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * cols);
  launch_pdl(gelu_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, cs(stream), bf(pre), ld_pre, bf(dy), ld_dy, bf(dx), ld_dx, rows,
                                                                      cols / 8);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_gate_residual_fwd(const void* res, const void* y, void* out, int64_t ld, int32_t rows, int32_t D,
                                    const lx_tile_meta_t* tile_meta, const void* const gate[3], const int64_t gate_stride[3],
                                    void* stream) {
  LX_CHECK_ARG(res && y && out && tile_meta && gate && gate_stride && rows > 0 && D > 0 && D % 8 == 0 && ld % 8 == 0,
               "lx_gate_residual_fwd: bad arguments");
  const int64_t n = (int64_t)rows * (D / 8);
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * D);
  launch_pdl(gate_residual_fwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, cs(stream), bf(res), bf(y), bf(out), ld, rows, D / 8,
                                                                               tile_meta, vec3(gate, gate_stride));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_gate_bwd(const void* dout, const void* y, void* dy, int64_t ld, int32_t rows, int32_t D,
                           const lx_tile_meta_t* tile_meta, const void* const gate[3], const int64_t gate_stride[3],
                           float* const dgate[3], const int64_t dgate_stride[3], void* stream) {
  LX_CHECK_ARG(dout && y && dy && tile_meta && gate && gate_stride && rows > 0 && rows % 128 == 0 && D > 0 && D % 2 == 0 &&
                   ld % 2 == 0,
               "lx_gate_bwd: bad arguments");
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * D);
  launch_pdl(gate_bwd_kernel, dim3(rows / 128, (D + 255) / 256), dim3(128), 0, cs(stream), bf(dout), bf(y), bf(dy), ld, rows, D, tile_meta,
                                                                            vec3(gate, gate_stride), acc3(dgate, dgate_stride));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_ln_modulate_bwd(const void* x, const void* dxn, const void* dres, void* dx, int64_t ld, int32_t rows,
                                  int32_t D, const lx_tile_meta_t* tile_meta, const void* const scale[3],
                                  const int64_t scale_stride[3], float* const dscale[3], float* const dshift[3],
                                  const int64_t dstride[3], float eps, float* stats_workspace, void* stream) {
  LX_CHECK_ARG(x && dxn && dx && tile_meta && scale && scale_stride && rows > 0 && rows % 128 == 0, "lx_ln_modulate_bwd: bad arguments");
  LX_CHECK_ARG(D > 0 && D % 8 == 0 && D <= LNB_MAX_NV * LNB_THREADS * 8 && ld % 8 == 0 && ld >= D,
               "lx_ln_modulate_bwd: D=%d must be a multiple of 8 and <= %d", D, LNB_MAX_NV * LNB_THREADS * 8);
  bool cols = false;
  for (int i = 0; i < 3; ++i) cols |= (dscale && dscale[i]) || (dshift && dshift[i]);
  LX_CHECK_ARG(!cols || stats_workspace, "lx_ln_modulate_bwd: stats workspace [rows,2] fp32 needed for dscale / dshift");
  {
    LaunchScope scope(KC_ROW, stream, (dres ? 8.0 : 6.0) * rows * D);
    launch_pdl(ln_mod_bwd_row_kernel, dim3(rows), dim3(LNB_THREADS), 0, cs(stream), bf(x), bf(dxn), bf(dres), bf(dx), ld, D, tile_meta,
                                                               vec3(scale, scale_stride), eps,
                                                               reinterpret_cast<float2*>(stats_workspace));
    LX_CUDA(cudaGetLastError());
  }
  if (cols) {
    LaunchScope scope(KC_ROW, stream, 4.0 * rows * D);
    launch_pdl(ln_mod_bwd_col_kernel, dim3(rows / 128, (D + 255) / 256), dim3(128), 0, cs(stream), 
        bf(x), bf(dxn), ld, rows, D, tile_meta, reinterpret_cast<const float2*>(stats_workspace), acc3(dscale, dstride),
        acc3(dshift, dstride));
    LX_CUDA(cudaGetLastError());
  }
  return LX_OK;
}

namespace {
RmsW rmsw(const float* const* q, const float* const* k) {
  RmsW w;
  for (int i = 0; i < 3; ++i) { w.q[i] = q ? q[i] : nullptr; w.k[i] = k ? k[i] : nullptr; }
  return w;
}
}  // namespace

extern "C" int lx_qkv_post_fwd(const void* qkv_pre, int64_t ld, int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta,
                               void* q, void* k, void* v, int32_t seq_total, const float* const rms_q[3],
                               const float* const rms_k[3], const float* rope, float eps, void* stream) {
  LX_CHECK_ARG(qkv_pre && q && k && v && tile_meta && rows > 0 && heads > 0 && seq_total > 0 && ld % 4 == 0 && ld >= 3 * heads * 128,
               "lx_qkv_post_fwd: bad arguments");
  LaunchScope scope(KC_ROW, stream, 12.0 * rows * heads * 128);
  launch_pdl(qkv_post_fwd_kernel, dim3(rows), dim3(128), 0, cs(stream), bf(qkv_pre), ld, rows, heads, tile_meta, bf(q), bf(k), bf(v), seq_total,
                                                   rmsw(rms_q, rms_k), rope, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_qkv_post_bwd(const void* qkv_pre, int64_t ld, const void* dq, const void* dk, const void* dv,
                               void* dqkv_pre, int64_t ldo, int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta,
                               int32_t seq_total, const float* const rms_q[3], const float* const rms_k[3], const float* rope,
                               float eps, void* stream) {
  LX_CHECK_ARG(qkv_pre && dq && dk && dv && dqkv_pre && tile_meta && rows > 0 && heads > 0 && seq_total > 0 && ld % 4 == 0 &&
                   ldo % 4 == 0 && ld >= 3 * heads * 128 && ldo >= 3 * heads * 128,
               "lx_qkv_post_bwd: bad arguments");
  LaunchScope scope(KC_ROW, stream, 18.0 * rows * heads * 128);
  launch_pdl(qkv_post_bwd_kernel, dim3(rows), dim3(128), 0, cs(stream), bf(qkv_pre), ld, bf(dq), bf(dk), bf(dv), bf(dqkv_pre), ldo, rows, heads,
                                                   tile_meta, seq_total, rmsw(rms_q, rms_k), rope, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_rows_to_heads(const void* rows_in, int64_t ld, void* heads_out, int32_t rows, int32_t heads,
                                const lx_tile_meta_t* tile_meta, int32_t seq_total, void* stream) {
  LX_CHECK_ARG(rows_in && heads_out && tile_meta && rows > 0 && heads > 0 && seq_total > 0 && ld % 4 == 0 && ld >= heads * 128,
               "lx_rows_to_heads: bad arguments");
  LaunchScope scope(KC_ROW, stream, 4.0 * rows * heads * 128);
  launch_pdl(rows_to_heads_kernel, dim3(rows), dim3(128), 0, cs(stream), bf(rows_in), ld, rows, heads, tile_meta, bf(heads_out), seq_total);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_lora_grad(const void* x, int64_t ldx, const void* dy, int64_t ldy, const float* A, const float* Bw,
                            float* dA, float* dB, int32_t M, int32_t K, int32_t N, int32_t r, float scaling,
                            float* workspace, void* stream) {
  LX_CHECK_ARG(x && dy && A && Bw && dA && dB && workspace && M > 0 && K > 0 && N > 0, "lx_lora_grad: bad arguments");
  LX_CHECK_ARG(r > 0 && r <= LORA_MAX_R && K % 4 == 0 && N % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0,
               "lx_lora_grad: rank %d must be in [1, %d], K / N / strides multiples of 4", r, LORA_MAX_R);
  float* P1 = workspace;                  // [M, r] = dy B
  float* P2 = workspace + (size_t)M * r;  // [M, r] = x A^T
  cudaStream_t st = cs(stream);
  const unsigned gr = (M + 127) / 128;
  LaunchScope scope(KC_ROW, stream, 4.0 * M * ((double)K + N));
  LX_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * 2 * (size_t)M * r, st));
  const unsigned gn = (N + LORA_CCHUNK - 1) / LORA_CCHUNK, gk = (K + LORA_CCHUNK - 1) / LORA_CCHUNK;
  if (r <= 4) {
    const unsigned gm = (M + 31) / 32;  // 4 warps x 8 rows
    // (the first kernel after the memset is launched stream-ordered, see launch_ordered)
    launch_ordered(lora_project_kernel<4, 8>, dim3(gm, gn), dim3(128), 0, st, bf(dy), ldy, M, N, Bw, r, 1, r, P1);  // F[c=n, j] = B[n*r + j]
    launch_pdl(lora_project_kernel<4, 8>, dim3(gm, gk), dim3(128), 0, st, bf(x), ldx, M, K, A, 1, K, r, P2);    // F[c=k, j] = A[j*K + k]
    launch_pdl(lora_reduce_kernel<4>, dim3(gr, (K + 511) / 512), dim3(128), 0, st, bf(x), ldx, M, K, P1, r, scaling, dA, 1, K);
    launch_pdl(lora_reduce_kernel<4>, dim3(gr, (N + 511) / 512), dim3(128), 0, st, bf(dy), ldy, M, N, P2, r, scaling, dB, r, 1);
  } else {
    const unsigned gm = (M + 7) / 8;  // 4 warps x 2 rows
    launch_ordered(lora_project_kernel<LORA_MAX_R, 2>, dim3(gm, gn), dim3(128), 0, st, bf(dy), ldy, M, N, Bw, r, 1, r, P1);
    launch_pdl(lora_project_kernel<LORA_MAX_R, 2>, dim3(gm, gk), dim3(128), 0, st, bf(x), ldx, M, K, A, 1, K, r, P2);
    launch_pdl(lora_reduce_kernel<LORA_MAX_R>, dim3(gr, (K + 511) / 512), dim3(128), 0, st, bf(x), ldx, M, K, P1, r, scaling, dA, 1, K);
    launch_pdl(lora_reduce_kernel<LORA_MAX_R>, dim3(gr, (N + 511) / 512), dim3(128), 0, st, bf(dy), ldy, M, N, P2, r, scaling, dB, r, 1);
  }
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_lora_merge(const void* W, int64_t ldw, const float* A, const float* Bw, void* out, int64_t ldo, int32_t N,
                             int32_t K, int32_t r, float scaling, void* stream) {
  LX_CHECK_ARG(W && A && Bw && out && N > 0 && K > 0 && K % 2 == 0 && r > 0 && ldw % 2 == 0 && ldo % 2 == 0,
               "lx_lora_merge: bad arguments");
  const int64_t n = (int64_t)N * (K / 2);
  LaunchScope scope(KC_ROW, stream, 4.0 * N * K);
  launch_pdl(lora_merge_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, cs(stream), bf(W), ldw, A, Bw, bf(out), ldo, N, K, r, scaling);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_transpose_bf16(const void* in, int64_t ld_in, void* out, int64_t ld_out, int32_t rows, int32_t cols,
                                 void* stream) {
  LX_CHECK_ARG(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows, "lx_transpose_bf16: bad arguments");
  LaunchScope scope(KC_ROW, stream, 4.0 * rows * cols);
  launch_pdl(transpose_bf16_kernel, dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, cs(stream), bf(in), ld_in, bf(out),
                                                                                                 ld_out, rows, cols);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_flow_noise_mix(const void* x0, const void* x1, const float* t, void* xt, int32_t B, int64_t per_sample,
                                 void* stream) {
  LX_CHECK_ARG(x0 && x1 && t && xt && B > 0 && per_sample > 0 && per_sample % 8 == 0, "lx_flow_noise_mix: bad arguments");
  const int64_t n8 = (int64_t)B * per_sample / 8;
  LaunchScope scope(KC_ROW, stream, 6.0 * B * per_sample);
  launch_pdl(flow_noise_mix_kernel, dim3((unsigned)((n8 + 255) / 256)), dim3(256), 0, cs(stream), bf(x0), bf(x1), t, bf(xt), per_sample / 8, n8);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_flow_mse_loss(const void* pred, const void* x0, const void* x1, float* loss, void* dpred, int64_t n,
                                float grad_scale, void* stream) {
  LX_CHECK_ARG(pred && x0 && x1 && loss && n > 0 && n % 8 == 0, "lx_flow_mse_loss: bad arguments");
  const int64_t n8 = n / 8;
  LaunchScope scope(KC_ROW, stream, 8.0 * n);
  launch_pdl(flow_mse_kernel, dim3((unsigned)((n8 + 255) / 256)), dim3(256), 0, cs(stream), bf(pred), bf(x0), bf(x1), loss, bf(dpred), n8,
                                                                       1.0f / (float)n, grad_scale);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
