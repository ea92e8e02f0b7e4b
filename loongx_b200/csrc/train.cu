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

__device__ __forceinline__ void unpack8(const uint4& u, float* x) {
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}

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

// dy = gate * dout ; dgate[stream, batch, col] += sum_rows dout * y.   One CTA = one 128-row tile x 256 columns: a warp
// owns 32 of the rows, a lane 8 adjacent columns (16-byte loads, a warp reads 512 contiguous bytes per row, four rows in
// flight); the four warps' column sums meet in shared memory, then one atomicAdd per column.
__global__ void __launch_bounds__(128) gate_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                       const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ dy,
                                                       int64_t ld, int rows, int D, const lx_tile_meta_t* __restrict__ tm,
                                                       Vec3 gate, Acc3 dgate) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float part[4][256];
  const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * 256 + lane * 8;
  const lx_tile_meta_t m = tm[tile];
  float* acc = dgate.p[m.stream];
  float sum[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sum[e] = 0.f;
  if (c < D) {
    float g[8];
    ld8(gate.p[m.stream] + (size_t)m.batch * gate.stride[m.stream] + c, g);
    const int r0 = tile * 128 + warp * 32, r1 = min(rows, r0 + 32);
    for (int r = r0; r < r1; r += 4) {
      uint4 d[4], v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const size_t off = (size_t)min(r + i, r1 - 1) * ld + c;
        d[i] = *reinterpret_cast<const uint4*>(dout + off);
        // no gate gradient wanted for this stream: y is not read at all (it may not even have been recomputed)
        v[i] = acc != nullptr ? *reinterpret_cast<const uint4*>(y + off) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r + i < r1) {
          float df[8], vf[8], o[8];
          unpack8(d[i], df);
          unpack8(v[i], vf);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            sum[e] += df[e] * vf[e];
            o[e] = g[e] * df[e];
          }
          st8(dy + (size_t)(r + i) * ld + c, o);
        }
      }
    }
  }
  if (acc == nullptr) return;  // uniform over the CTA
#pragma unroll
  for (int e = 0; e < 8; ++e) part[warp][lane * 8 + e] = sum[e];
  __syncthreads();
  for (int k = threadIdx.x; k < 256; k += 128) {
    const int col = blockIdx.y * 256 + k;
    if (col < D)
      atomicAdd(acc + (size_t)m.batch * dgate.stride[m.stream] + col, part[0][k] + part[1][k] + part[2][k] + part[3][k]);
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

// CTA = one 128-row tile x 256 columns: a warp owns 32 of the rows, a lane 8 adjacent columns (16-byte loads, four rows in
// flight); the four warps' column sums meet in shared memory, then one atomicAdd per column and output.
__global__ void __launch_bounds__(128) ln_mod_bwd_col_kernel(const __nv_bfloat16* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ dxn, int64_t ld, int rows,
                                                             int D, const lx_tile_meta_t* __restrict__ tm,
                                                             const float2* __restrict__ stats, Acc3 dscale, Acc3 dshift) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float part[2][4][256];
  const int tile = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.y * 256 + lane * 8;
  const lx_tile_meta_t m = tm[tile];
  if (dscale.p[m.stream] == nullptr && dshift.p[m.stream] == nullptr) return;  // uniform over the CTA
  float ss[8], hh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) ss[e] = hh[e] = 0.f;
  if (c < D) {
    const int r0 = tile * 128 + warp * 32, r1 = min(rows, r0 + 32);
    for (int r = r0; r < r1; r += 4) {
      uint4 xv[4], dv[4];
      float2 st[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = min(r + i, r1 - 1);
        xv[i] = *reinterpret_cast<const uint4*>(x + (size_t)rr * ld + c);
        dv[i] = *reinterpret_cast<const uint4*>(dxn + (size_t)rr * ld + c);
        st[i] = stats[rr];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (r + i < r1) {
          float xf[8], df[8];
          unpack8(xv[i], xf);
          unpack8(dv[i], df);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            ss[e] += df[e] * (xf[e] - st[i].x) * st[i].y;
            hh[e] += df[e];
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    part[0][warp][lane * 8 + e] = ss[e];
    part[1][warp][lane * 8 + e] = hh[e];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 256; k += 128) {
    const int col = blockIdx.y * 256 + k;
    if (col >= D) continue;
    if (dscale.p[m.stream] != nullptr)
      atomicAdd(dscale.p[m.stream] + (size_t)m.batch * dscale.stride[m.stream] + col,
                part[0][0][k] + part[0][1][k] + part[0][2][k] + part[0][3][k]);
    if (dshift.p[m.stream] != nullptr)
      atomicAdd(dshift.p[m.stream] + (size_t)m.batch * dshift.stride[m.stream] + col,
                part[1][0][k] + part[1][1][k] + part[1][2][k] + part[1][3][k]);
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

// kDqF32: dq is the fp32 accumulation buffer of lx_attention_bwd (read directly: no separate cast pass)
template <bool kDqF32>
__global__ void __launch_bounds__(128) qkv_post_bwd_kernel(const __nv_bfloat16* __restrict__ pre, int64_t ld,
                                                           const void* __restrict__ dq,
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
      if (kDqF32 && which == 0) {
        const float4 t4 = *reinterpret_cast<const float4*>(static_cast<const float*>(dq) + src);
        g[0] = t4.x; g[1] = t4.y; g[2] = t4.z; g[3] = t4.w;
      } else {
        ld4((which == 0 ? static_cast<const __nv_bfloat16*>(dq) : which == 1 ? dk : dv) + src, g);
      }
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

// ---------------------------------------------------------------------------------------------------------------
// Stacked LoRA factor gradients: G <= 4 sub-Linears that read the same x and whose outputs are adjacent column blocks
// of one dy (to_q | to_k | to_v (| proj_mlp)), rank r in {4, 8, 16}, RT = G * r <= 16.  Four launches for the whole
// group, each reading its operand once with 16-byte (8-byte in the widest reduce) loads:
//   project_x:  PX[m, g r + j] = sum_k x[m, k] A_g[j, k]                 project_dy: PY[m, g r + j] = sum_n dy[m, c_g + n] B_g[n, j]
//   reduce_x:   dA_g[j, k] += s_g sum_m PY[m, g r + j] x[m, k]           reduce_dy:  dB_g[n, j] += s_g sum_m dy[m, c_g + n] PX[m, g r + j]
// ---------------------------------------------------------------------------------------------------------------
constexpr int LORA_STACK_G = 4;
constexpr int LORA_STACK_RT = 16;
struct LoraStack {
  int G, r;
  int col0[LORA_STACK_G + 1];          // first dy column of group g; col0[G] = total width
  const float* a_row[LORA_STACK_RT];   // a_row[g r + j] = A_g + j K
  float* da_row[LORA_STACK_RT];        // dA_g + j K
  float s_row[LORA_STACK_RT];          // s_g
  float s_grp[LORA_STACK_G];
  const float* B[LORA_STACK_G];
  float* dB[LORA_STACK_G];
};

// The two projections run as ONE launch (the first CTAs take x, the rest dy), and so do the two reductions: each of the
// four pieces alone leaves SMs idle on the DiT's shapes (M = the condition rows, 25-100 MB per operand).
//
// Projection: a warp owns a strip of 32 * COLS columns (lane = COLS adjacent columns) and LP_RB rows.  Its factor values
// F[COLS][RT] are loaded ONCE into registers; the rows then stream through with eight 16-byte (8-byte) loads in flight per
// lane and nothing else on the load path.  Per row a lane holds RT partial sums; 32 of them (32 / RT rows) are reduced
// across the warp with a 31-shuffle transposing butterfly (lane l ends up with the total of value l).  The 8 warps of a CTA
// take 8 neighbouring strips of the same rows and meet in shared memory, so the strip sums reach P with one atomicAdd per
// (row, output, CTA).  (The first version re-loaded the factors for every 4 rows x 256 columns: 1.4 TB/s, latency-bound.)
constexpr int LP_RB = 64;     // rows per CTA
constexpr int LP_WARPS = 8;   // strips per CTA

__device__ __forceinline__ float butterfly32(float (&v)[32], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int o = 16 >> s;  // lane distance = number of values that survive this stage
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int k = 0; k < o; ++k) {
      const float send = up ? v[k] : v[k + o];
      const float keep = up ? v[k + o] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// rows [row0, row0 + LP_RB) of Z times F -> slot[(row - row0) * RT + j] (this warp's strip only)
template <int RT, int COLS>
__device__ __forceinline__ void lora_strip_rows(const __nv_bfloat16* __restrict__ Z, int64_t ldz, int M, int row0, int c,
                                                bool live, const float (&F)[COLS][RT], float* __restrict__ slot) {
  constexpr int U = 32 / RT;  // rows per butterfly
  const int lane = threadIdx.x & 31;
  for (int rb = 0; rb < LP_RB; rb += 8) {
    float z[8][COLS];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const __nv_bfloat16* src = Z + (size_t)min(row0 + rb + i, M - 1) * ldz + c;
      if constexpr (COLS == 8) {
        unpack8(live ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0), z[i]);
      } else {
        const uint2 u = live ? __ldg(reinterpret_cast<const uint2*>(src)) : make_uint2(0, 0);
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        z[i][0] = a.x; z[i][1] = a.y; z[i][2] = b.x; z[i][3] = b.y;
      }
    }
#pragma unroll
    for (int u = 0; u < 8 / U; ++u) {
      float v[32];
#pragma unroll
      for (int q = 0; q < U; ++q)
#pragma unroll
        for (int j = 0; j < RT; ++j) {
          float a = 0.f;
#pragma unroll
          for (int e = 0; e < COLS; ++e) a += z[u * U + q][e] * F[e][j];
          v[q * RT + j] = a;
        }
      slot[(rb + u * U) * RT + lane] = butterfly32(v, lane);  // value l = (row l / RT, output l % RT)
    }
  }
}

struct LoraStackGrid {
  int x_bx, x_n;   // CTAs [0, x_n) work on x: (bx, by) = (b % x_bx, b / x_bx); the rest on dy: b - x_n -> (% dy_bx, / dy_bx)
  int dy_bx;
};

__device__ __forceinline__ int lora_group_of(const LoraStack& p, int c) {
  int g = 0;
#pragma unroll
  for (int i = 1; i < LORA_STACK_G; ++i) g += (i < p.G && c >= p.col0[i]) ? 1 : 0;
  return g;
}

// RT: outputs per row of the x part (G r, 12 rounded up to 16); R: rank = outputs per row of the dy part
template <int RT, int R>
__global__ void __launch_bounds__(256, 2) lora_stack_project_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                    const __nv_bfloat16* __restrict__ dy, int64_t ldy, int M,
                                                                    int K, const LoraStackGrid gr, const LoraStack p,
                                                                    float* __restrict__ PX, float* __restrict__ PY) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int RS = RT > R ? RT : R;
  __shared__ __align__(16) float slots[LP_WARPS][LP_RB * RS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  const int rt_tot = p.G * p.r;  // row stride of PX / PY
  if (b < gr.x_n) {
    constexpr int COLS = RT == 16 ? 4 : 8;
    const int row0 = (b % gr.x_bx) * LP_RB;
    const int strip = (b / gr.x_bx) * LP_WARPS + warp;
    const int c = strip * (32 * COLS) + lane * COLS;
    const bool live = c < K;  // (K is a multiple of 8: a lane is all in or all out)
    float F[COLS][RT];
#pragma unroll
    for (int j = 0; j < RT; ++j) {
      const bool on = live && j < rt_tot;
      if constexpr (COLS == 8) {
        const float4 f0 = on ? __ldg(reinterpret_cast<const float4*>(p.a_row[j] + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 f1 = on ? __ldg(reinterpret_cast<const float4*>(p.a_row[j] + c + 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        F[0][j] = f0.x; F[1][j] = f0.y; F[2][j] = f0.z; F[3][j] = f0.w;
        F[4][j] = f1.x; F[5][j] = f1.y; F[6][j] = f1.z; F[7][j] = f1.w;
      } else {
        const float4 f0 = on ? __ldg(reinterpret_cast<const float4*>(p.a_row[j] + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        F[0][j] = f0.x; F[1][j] = f0.y; F[2][j] = f0.z; F[3][j] = f0.w;
      }
    }
    const int n_strip = (K + 32 * COLS - 1) / (32 * COLS);
    if (strip < n_strip) lora_strip_rows<RT, COLS>(x, ldx, M, row0, live ? c : 0, live, F, slots[warp]);
    __syncthreads();
    const int n_w = min(LP_WARPS, n_strip - (b / gr.x_bx) * LP_WARPS);
    for (int i = threadIdx.x; i < LP_RB * RT; i += 256) {
      const int row = row0 + i / RT, j = i % RT;
      if (row >= M || j >= rt_tot) continue;
      float a = 0.f;
      for (int w = 0; w < n_w; ++w) a += slots[w][i];
      atomicAdd(PX + (size_t)row * rt_tot + j, a);
    }
  } else {
    constexpr int COLS = R == 16 ? 4 : 8;
    const int bb = b - gr.x_n;
    const int row0 = (bb % gr.dy_bx) * LP_RB;
    const int strip0 = (bb / gr.dy_bx) * LP_WARPS, strip = strip0 + warp;
    const int N = p.col0[LORA_STACK_G], n_strip = N / (32 * COLS);  // (every group width is a multiple of 256)
    const int c = strip * (32 * COLS) + lane * COLS;
    if (strip < n_strip) {
      const int g = lora_group_of(p, c);
      // (selects: run-time indexing of kernel-parameter arrays goes through local memory)
      const float* __restrict__ Bg = g == 0 ? p.B[0] : (g == 1 ? p.B[1] : (g == 2 ? p.B[2] : p.B[3]));
      const int cg = g == 0 ? 0 : (g == 1 ? p.col0[1] : (g == 2 ? p.col0[2] : p.col0[3]));  // B_g row of dy column c: c - cg
      float F[COLS][R];
#pragma unroll
      for (int e = 0; e < COLS; ++e)
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(Bg + (size_t)(c + e - cg) * R + q * 4));
          F[e][q * 4] = t.x; F[e][q * 4 + 1] = t.y; F[e][q * 4 + 2] = t.z; F[e][q * 4 + 3] = t.w;
        }
      lora_strip_rows<R, COLS>(dy, ldy, M, row0, c, true, F, slots[warp]);
    }
    __syncthreads();
    const int n_w = min(LP_WARPS, n_strip - strip0);
    for (int i = threadIdx.x; i < LP_RB * R; i += 256) {
      const int row = row0 + i / R, j = i % R;
      if (row >= M) continue;
      float a = 0.f;
      for (int w = 0; w < n_w; ++w) {  // neighbouring strips of one group are summed before they go out
        a += slots[w][i];
        const int g = lora_group_of(p, (strip0 + w) * (32 * COLS));
        if (w + 1 == n_w || lora_group_of(p, (strip0 + w + 1) * (32 * COLS)) != g) {
          atomicAdd(PY + (size_t)row * rt_tot + g * R + j, a);
          a = 0.f;
        }
      }
    }
  }
}

// reduce pieces: CTA = 32 rows x 512 columns of x (4 adjacent columns per thread, 8 rows in flight) or 32 rows x
// (128 * COLS) columns of dy (a thread's COLS adjacent columns lie inside one group); P rows from shared memory.
constexpr int LORA_RROWS = 32;
template <int RT>
__device__ __forceinline__ void lora_stack_reduce_x(const __nv_bfloat16* __restrict__ x, int64_t ldx, int M, int K,
                                                    const LoraStack& p, const float* __restrict__ PY, float* ps, int bx, int by) {
  const int r0 = bx * LORA_RROWS, nrow = min(M - r0, LORA_RROWS);
  for (int i = threadIdx.x; i < LORA_RROWS * RT; i += 128) ps[i] = (i / RT < nrow) ? PY[(size_t)r0 * RT + i] : 0.f;
  __syncthreads();
  const int c = by * 512 + threadIdx.x * 4;
  if (c >= K) return;
  float acc[4][RT];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int j = 0; j < RT; ++j) acc[e][j] = 0.f;
  for (int m = 0; m < nrow; m += 8) {
    uint2 u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)  // rows past the end re-read the last row; their P rows are zero
      u[i] = __ldg(reinterpret_cast<const uint2*>(x + (size_t)(r0 + min(m + i, nrow - 1)) * ldx + c));
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 a = unpack_bf16(u[i].x), b = unpack_bf16(u[i].y);
      const float* pr = ps + (m + i) * RT;  // m + i < LORA_RROWS always (m a multiple of 8)
#pragma unroll
      for (int q = 0; q < RT / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(pr + q * 4);
        const float pv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          acc[0][q * 4 + jj] += a.x * pv[jj];
          acc[1][q * 4 + jj] += a.y * pv[jj];
          acc[2][q * 4 + jj] += b.x * pv[jj];
          acc[3][q * 4 + jj] += b.y * pv[jj];
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < RT; ++j) {
    const float s = p.s_row[j];
    atomicAdd(reinterpret_cast<float4*>(p.da_row[j] + c), make_float4(s * acc[0][j], s * acc[1][j], s * acc[2][j], s * acc[3][j]));
  }
}

template <int R, int COLS>
__device__ __forceinline__ void lora_stack_reduce_dy(const __nv_bfloat16* __restrict__ dy, int64_t ldy, int M, const LoraStack& p,
                                                     const float* __restrict__ PX, float* ps, int bx, int by) {
  const int RT = p.G * R;
  const int r0 = bx * LORA_RROWS, nrow = min(M - r0, LORA_RROWS);
  for (int i = threadIdx.x; i < LORA_RROWS * RT; i += 128) ps[i] = (i / RT < nrow) ? PX[(size_t)r0 * RT + i] : 0.f;
  __syncthreads();
  const int c = (by * 128 + threadIdx.x) * COLS;
  if (c >= p.col0[LORA_STACK_G]) return;
  int g = 0;
#pragma unroll
  for (int i = 1; i < LORA_STACK_G; ++i) g += (i < p.G && c >= p.col0[i]) ? 1 : 0;
  float acc[COLS][R];
#pragma unroll
  for (int e = 0; e < COLS; ++e)
#pragma unroll
    for (int j = 0; j < R; ++j) acc[e][j] = 0.f;
  constexpr int INF = 64 / COLS;  // rows in flight: 128 bytes per thread
  for (int m = 0; m < nrow; m += INF) {
    float z[INF][COLS];
#pragma unroll
    for (int i = 0; i < INF; ++i) {
      const __nv_bfloat16* src = dy + (size_t)(r0 + min(m + i, nrow - 1)) * ldy + c;
      if constexpr (COLS == 8) {
        unpack8(__ldg(reinterpret_cast<const uint4*>(src)), z[i]);
      } else {
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(src));
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
        z[i][0] = a.x; z[i][1] = a.y; z[i][2] = b.x; z[i][3] = b.y;
      }
    }
#pragma unroll
    for (int i = 0; i < INF; ++i) {
      const float* pr = ps + (m + i) * RT + g * R;  // m + i < LORA_RROWS always (INF divides it)
      float pv[R];
#pragma unroll
      for (int q = 0; q < R / 4; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(pr + q * 4);
        pv[q * 4] = t.x; pv[q * 4 + 1] = t.y; pv[q * 4 + 2] = t.z; pv[q * 4 + 3] = t.w;
      }
#pragma unroll
      for (int e = 0; e < COLS; ++e)
#pragma unroll
        for (int j = 0; j < R; ++j) acc[e][j] += z[i][e] * pv[j];
    }
  }
  const float s = g == 0 ? p.s_grp[0] : (g == 1 ? p.s_grp[1] : (g == 2 ? p.s_grp[2] : p.s_grp[3]));
  float* dBg = g == 0 ? p.dB[0] : (g == 1 ? p.dB[1] : (g == 2 ? p.dB[2] : p.dB[3]));
  const int cg = g == 0 ? 0 : (g == 1 ? p.col0[1] : (g == 2 ? p.col0[2] : p.col0[3]));
  float* out = dBg + (size_t)(c - cg) * R;
#pragma unroll
  for (int e = 0; e < COLS; ++e)
#pragma unroll
    for (int q = 0; q < R / 4; ++q)
      atomicAdd(reinterpret_cast<float4*>(out + (size_t)e * R + q * 4),
                make_float4(s * acc[e][q * 4], s * acc[e][q * 4 + 1], s * acc[e][q * 4 + 2], s * acc[e][q * 4 + 3]));
}

template <int RT, int R, int COLS>
__global__ void __launch_bounds__(128) lora_stack_reduce_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx,
                                                                const __nv_bfloat16* __restrict__ dy, int64_t ldy, int M, int K,
                                                                const LoraStackGrid gr, const LoraStack p,
                                                                const float* __restrict__ PX, const float* __restrict__ PY) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) float ps[LORA_RROWS * LORA_STACK_RT];
  const int b = blockIdx.x;
  if (b < gr.x_n) lora_stack_reduce_x<RT>(x, ldx, M, K, p, PY, ps, b % gr.x_bx, b / gr.x_bx);
  else lora_stack_reduce_dy<R, COLS>(dy, ldy, M, p, PX, ps, (b - gr.x_n) % gr.dy_bx, (b - gr.x_n) / gr.dy_bx);
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

// The same merge writing BOTH panels in one pass over W: out[n, k] and outT[k, n] (the K-major copy the dX GEMMs multiply
// by).  CTA = 64 x 64 tile, 256 threads; the merged tile goes out row-major directly and transposed through shared memory
// (16-byte stores on both sides).  After every optimizer step all 343 LoRA targets are rebuilt: W is read once (24 GB for
// FLUX.1-dev) instead of merge (read W, write out) + transpose (read out, write outT).
constexpr int LMT = 64;
__global__ void __launch_bounds__(256) lora_merge_t_kernel(const __nv_bfloat16* __restrict__ W, int64_t ldw,
                                                           const float* __restrict__ A, const float* __restrict__ Bw,
                                                           __nv_bfloat16* __restrict__ out, int64_t ldo,
                                                           __nv_bfloat16* __restrict__ outT, int64_t ldt, int N, int K, int r,
                                                           float scaling) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float sA[LORA_MAX_R][LMT];        // A[j, k0 + c]
  __shared__ float sB[LMT][LORA_MAX_R + 1];    // B[n0 + row, j]
  __shared__ __align__(16) __nv_bfloat16 sT[LMT][LMT + 8];  // merged tile [row][swizzled col] (+8: rows stay 16-byte aligned)
  const int n0 = blockIdx.y * LMT, k0 = blockIdx.x * LMT;
  for (int i = threadIdx.x; i < r * LMT; i += 256) {
    const int j = i / LMT, c = i % LMT;
    sA[j][c] = (k0 + c < K) ? A[(size_t)j * K + k0 + c] : 0.f;
  }
  for (int i = threadIdx.x; i < LMT * r; i += 256) {
    const int row = i / r, j = i % r;
    sB[row][j] = (n0 + row < N) ? Bw[(size_t)(n0 + row) * r + j] : 0.f;
  }
  __syncthreads();
  // thread -> (row, 8-column chunk): 64 rows x 8 chunks = 512 slots, two per thread
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int slot = threadIdx.x + h * 256, row = slot >> 3, c8 = (slot & 7) * 8;
    const int n = n0 + row, k = k0 + c8;
    float v[8];
    if (n < N && k < K) {
      ld8(W + (size_t)n * ldw + k, v);
      float ba[8];  // (B A)[n, k..k+8) summed in the order of lora_merge_kernel: the two kernels agree bit for bit
#pragma unroll
      for (int e = 0; e < 8; ++e) ba[e] = 0.f;
      for (int j = 0; j < r; ++j) {
        const float b = sB[row][j];
#pragma unroll
        for (int e = 0; e < 8; ++e) ba[e] = fmaf(b, sA[j][c8 + e], ba[e]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(scaling, ba[e], v[e]);
      st8(out + (size_t)n * ldo + k, v);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
    st8(&sT[row][c8 ^ (((row >> 3) & 7) << 3)], v);  // 8-column chunks XOR-swizzled by the row octet: conflict-free transposed reads
  }
  __syncthreads();
  // transposed write: thread -> (column of the tile = row of outT, 8-row chunk)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int slot = threadIdx.x + h * 256, col = slot >> 3, r8 = (slot & 7) * 8;
    const int k = k0 + col, n = n0 + r8;
    if (k < K && n < N) {
      __nv_bfloat16 t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) t[e] = sT[r8 + e][(((col >> 3) ^ ((r8 >> 3) & 7)) << 3) | (col & 7)];
      *reinterpret_cast<uint4*>(outT + (size_t)k * ldt + n) = *reinterpret_cast<const uint4*>(t);
    }
  }
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

// ONE CTA walks the whole prediction (a few hundred KB once per step) and reduces in a fixed order: the loss is
// bit-reproducible from run to run (a grid of CTAs meeting in an atomicAdd was not).
__global__ void __launch_bounds__(1024) flow_mse_kernel(const __nv_bfloat16* __restrict__ pred,
                                                        const __nv_bfloat16* __restrict__ x0,
                                                        const __nv_bfloat16* __restrict__ x1, float* __restrict__ loss,
                                                        __nv_bfloat16* __restrict__ dpred, int64_t n8, float inv_n,
                                                        float grad_scale) {
  pdl_wait();  // PDL: inputs are the previous kernel's outputs
  pdl_launch_dependents();
  __shared__ float red[32];
  float s = 0.f;
  for (int64_t idx = threadIdx.x; idx < n8; idx += 1024) {
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
    for (int i = 0; i < 32; ++i) tot += red[i];
    *loss += tot * inv_n;  // accumulates (micro-batches): the caller zeroes it
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
  LX_CHECK_ARG(dout && y && dy && tile_meta && gate && gate_stride && rows > 0 && rows % 128 == 0 && D > 0 && D % 8 == 0 &&
                   ld % 8 == 0,
               "lx_gate_bwd: bad arguments (rows a multiple of 128, D / ld multiples of 8)");
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
  launch_pdl(qkv_post_bwd_kernel<false>, dim3(rows), dim3(128), 0, cs(stream), bf(qkv_pre), ld, dq, bf(dk), bf(dv), bf(dqkv_pre), ldo, rows, heads,
                                                   tile_meta, seq_total, rmsw(rms_q, rms_k), rope, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_qkv_post_bwd_f32dq(const void* qkv_pre, int64_t ld, const float* dq, const void* dk, const void* dv,
                                     void* dqkv_pre, int64_t ldo, int32_t rows, int32_t heads, const lx_tile_meta_t* tile_meta,
                                     int32_t seq_total, const float* const rms_q[3], const float* const rms_k[3],
                                     const float* rope, float eps, void* stream) {
  LX_CHECK_ARG(qkv_pre && dq && dk && dv && dqkv_pre && tile_meta && rows > 0 && heads > 0 && seq_total > 0 && ld % 4 == 0 &&
                   ldo % 4 == 0 && ld >= 3 * heads * 128 && ldo >= 3 * heads * 128,
               "lx_qkv_post_bwd_f32dq: bad arguments");
  LaunchScope scope(KC_ROW, stream, 20.0 * rows * heads * 128);
  launch_pdl(qkv_post_bwd_kernel<true>, dim3(rows), dim3(128), 0, cs(stream), bf(qkv_pre), ld, static_cast<const void*>(dq), bf(dk), bf(dv),
             bf(dqkv_pre), ldo, rows, heads, tile_meta, seq_total, rmsw(rms_q, rms_k), rope, eps);
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

extern "C" int lx_lora_grad_stacked(const void* x, int64_t ldx, const void* dy, int64_t ldy, int32_t M, int32_t K,
                                    const lx_lora_stack_t* st, float* workspace, void* stream) {
  LX_CHECK_ARG(x && dy && st && workspace && M > 0 && K > 0, "lx_lora_grad_stacked: bad arguments");
  const int G = st->groups, r = st->r, RT = G * r;
  LX_CHECK_ARG(G >= 1 && G <= LORA_STACK_G && (r == 4 || r == 8 || r == 16) && RT <= LORA_STACK_RT,
               "lx_lora_grad_stacked: %d groups of rank %d (rank in {4, 8, 16}, groups * rank <= %d)", G, r, LORA_STACK_RT);
  LX_CHECK_ARG(K % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "lx_lora_grad_stacked: K / strides must be multiples of 8");
  LoraStack p{};
  p.G = G;
  p.r = r;
  int chunk = 2048, N = 0;
  for (int g = 0; g < G; ++g) {
    LX_CHECK_ARG(st->A[g] && st->B[g] && st->dA[g] && st->dB[g] && st->width[g] > 0 && st->width[g] % 256 == 0,
                 "lx_lora_grad_stacked: group %d: null factor or width %d not a multiple of 256", g, st->width[g]);
    LX_CHECK_ARG(((uintptr_t)st->A[g] | (uintptr_t)st->B[g] | (uintptr_t)st->dA[g] | (uintptr_t)st->dB[g]) % 16 == 0,
                 "lx_lora_grad_stacked: group %d: factors must be 16-byte aligned", g);
    p.col0[g] = N;
    N += st->width[g];
    while (st->width[g] % chunk) chunk >>= 1;
    p.B[g] = st->B[g];
    p.dB[g] = st->dB[g];
    p.s_grp[g] = st->scaling[g];
    for (int j = 0; j < r; ++j) {
      p.a_row[g * r + j] = st->A[g] + (size_t)j * K;
      p.da_row[g * r + j] = st->dA[g] + (size_t)j * K;
      p.s_row[g * r + j] = st->scaling[g];
    }
  }
  for (int g = G; g <= LORA_STACK_G; ++g) p.col0[g] = N;
  float* PX = workspace;                   // [M, RT] = x A^T (all groups)
  float* PY = workspace + (size_t)M * RT;  // [M, RT] = dy_g B_g
  cudaStream_t s = cs(stream);
  LaunchScope scope(KC_ROW, stream, 4.0 * M * ((double)K + N));
  LX_CUDA(cudaMemsetAsync(workspace, 0, sizeof(float) * 2 * (size_t)M * RT, s));
  // projection launch: CTA = LP_RB rows x LP_WARPS strips of 32 * COLS columns; the x part first, then the dy part
  const int nrb = (M + LP_RB - 1) / LP_RB;
  const int rtx = RT == 12 ? 16 : RT;  // (three groups of rank 4 run the 16-output variant with four idle outputs)
  const int strip_x = 32 * (rtx == 16 ? 4 : 8), strip_y = 32 * (r == 16 ? 4 : 8);
  const int nsx = (K + strip_x - 1) / strip_x, nsy = N / strip_y;
  const LoraStackGrid gp{nrb, nrb * ((nsx + LP_WARPS - 1) / LP_WARPS), nrb};
  const unsigned n_project = (unsigned)(gp.x_n + nrb * ((nsy + LP_WARPS - 1) / LP_WARPS));
  (void)chunk;
  const int grr = (M + LORA_RROWS - 1) / LORA_RROWS;
  const int cols = r == 16 ? 4 : 8;
  const LoraStackGrid gq{grr, grr * ((K + 511) / 512), grr};
  const unsigned n_reduce = (unsigned)(gq.x_n + grr * ((N + 128 * cols - 1) / (128 * cols)));
  // (the first kernel after the memset is launched stream-ordered, see launch_ordered)
#define LX_STACK(RTV, RTXV, RV, COLSV)                                                                                        \
  LX_CUDA(launch_ordered(lora_stack_project_kernel<RTXV, RV>, dim3(n_project), dim3(256), 0, s, bf(x), ldx, bf(dy), ldy, M, K, \
                         gp, p, PX, PY));                                                                                    \
  LX_CUDA(launch_pdl(lora_stack_reduce_kernel<RTV, RV, COLSV>, dim3(n_reduce), dim3(128), 0, s, bf(x), ldx, bf(dy), ldy, M, K, \
                     gq, p, static_cast<const float*>(PX), static_cast<const float*>(PY)))
  if (r == 4) {
    if (RT == 4) { LX_STACK(4, 4, 4, 8); }
    else if (RT == 8) { LX_STACK(8, 8, 4, 8); }
    else if (RT == 12) { LX_STACK(12, 16, 4, 8); }
    else { LX_STACK(16, 16, 4, 8); }
  } else if (r == 8) {
    if (RT == 8) { LX_STACK(8, 8, 8, 8); }
    else { LX_STACK(16, 16, 8, 8); }
  } else {
    LX_STACK(16, 16, 16, 4);
  }
#undef LX_STACK
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

extern "C" int lx_lora_merge_t(const void* W, int64_t ldw, const float* A, const float* Bw, void* out, int64_t ldo, void* outT,
                               int64_t ldt, int32_t N, int32_t K, int32_t r, float scaling, void* stream) {
  LX_CHECK_ARG(W && A && Bw && out && outT && N > 0 && K > 0 && N % 8 == 0 && K % 8 == 0 && r > 0 && r <= LORA_MAX_R &&
                   ldw % 8 == 0 && ldo % 8 == 0 && ldt % 8 == 0,
               "lx_lora_merge_t: bad arguments (N, K and the strides multiples of 8, r <= %d)", LORA_MAX_R);
  LX_CHECK_ARG(((uintptr_t)W | (uintptr_t)out | (uintptr_t)outT) % 16 == 0, "lx_lora_merge_t: panels must be 16-byte aligned");
  LaunchScope scope(KC_ROW, stream, 6.0 * N * K);
  launch_pdl(lora_merge_t_kernel, dim3((K + LMT - 1) / LMT, (N + LMT - 1) / LMT), dim3(256), 0, cs(stream), bf(W), ldw, A, Bw,
             bf(out), ldo, bf(outT), ldt, N, K, r, scaling);
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
  launch_pdl(flow_mse_kernel, dim3(1), dim3(1024), 0, cs(stream), bf(pred), bf(x0), bf(x1), loss, bf(dpred), n8,
                                                                       1.0f / (float)n, grad_scale);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
