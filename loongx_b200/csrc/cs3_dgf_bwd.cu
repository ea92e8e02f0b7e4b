// Backward of the CS3 signal encoders and the DGF (DUAN) fusion, fp32 — the parameter gradients the reference's
// OminiModel.step() + loss.backward() leaves on every encoder / fusion parameter (src/train/model.py:656-701 is inside
// the autograd graph; Lightning's DDP all-reduces them with the LoRA factors, train.py:181-183).
//
//   sgemm_ex            C = alpha op(A) op(B) + beta C, per batch element or summed over the batch (weight gradients of
//                       every 1x1 conv / token-axis / channel Linear: dW = sum_b dY X^T;  input gradients: dX = W^T dY)
//   sum_rows / sum_last bias gradients
//   ln_relu_rows_bwd    Linear -> LayerNorm -> ReLU of the projection MLPs (model.py:60-72)
//   dropout             nn.Dropout(0.3) of the projection MLPs in train mode (counter-based mask, same mask both ways)
//   token_linear_bwd    Unflatten(512, 8) -> Linear(8, 4096)
//   adaptive_pool_bwd   nn.AdaptiveAvgPool1d
//   channel_ln_bwd      LayerNorm over the channels of [B, d, L] (S4Block post-norm)
//   gelu_erf_bwd, s4_conv (reversible, optional GELU), s4_conv_wgrad, s4_kernel_gen_bwd   the S4 layer
//   duan_bwd_*          DUAN: FiLM / mixed-statistics / gate paths (the top-k mask is a constant, like in autograd)
#include <string.h>

#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

// ------------------------------------------------------------------------------------------------ general fp32 GEMM
struct SgemmEx {
  const float* A;
  const float* Bm;
  float* C;
  int64_t lda, ldb, ldc, a_bs, b_bs, c_bs;
  int M, N, K, batch;
  int trans_a, trans_b, reduce_batch;
  int ksplit;  // reduce mode: grid.z = batch * ksplit CTAs add their (batch element, K range) partial into C with fp32 atomics
  float alpha, beta;
};

// 64x64x16 tiles, 256 threads, 4x4 outputs per thread.  op(A) is [M, K]: stored [M, K] (trans_a = 0) or [K, M]
// (trans_a = 1); op(B) is [K, N]: stored [K, N] (trans_b = 0) or [N, K] (trans_b = 1).
__global__ void __launch_bounds__(256) sgemm_ex_kernel(const SgemmEx p) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  // reduce mode: CTA z = (batch element, K slice); the slices are multiples of 16
  const int b = p.reduce_batch ? blockIdx.z / p.ksplit : blockIdx.z;
  const int kper = p.reduce_batch ? ((p.K + p.ksplit * 16 - 1) / (p.ksplit * 16)) * 16 : p.K;
  const int k_lo = p.reduce_batch ? (blockIdx.z % p.ksplit) * kper : 0, k_hi = min(p.K, k_lo + kper);
  {
    const float* Ab = p.A + (size_t)b * p.a_bs;
    const float* Bb = p.Bm + (size_t)b * p.b_bs;
    // The next k-tile's global loads are issued before this tile's FMAs (registers as the second buffer); per output
    // element the accumulation order is unchanged (k ascending inside the slice).
    float ra[4], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = threadIdx.x + t * 256;
        int m, k;
        if (p.trans_a) { k = i >> 6; m = i & 63; } else { m = i >> 4; k = i & 15; }
        const bool oka = m0 + m < p.M && k0 + k < k_hi;
        ra[t] = oka ? (p.trans_a ? Ab[(size_t)(k0 + k) * p.lda + m0 + m] : Ab[(size_t)(m0 + m) * p.lda + k0 + k]) : 0.f;
        int n;
        if (p.trans_b) { n = i >> 4; k = i & 15; } else { k = i >> 6; n = i & 63; }
        const bool okb = k0 + k < k_hi && n0 + n < p.N;
        rb[t] = okb ? (p.trans_b ? Bb[(size_t)(n0 + n) * p.ldb + k0 + k] : Bb[(size_t)(k0 + k) * p.ldb + n0 + n]) : 0.f;
      }
    };
    if (k_lo < k_hi) fetch(k_lo);
    for (int k0 = k_lo; k0 < k_hi; k0 += 16) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int i = threadIdx.x + t * 256;
        if (p.trans_a) sA[i >> 6][i & 63] = ra[t]; else sA[i & 15][i >> 4] = ra[t];
        if (p.trans_b) sB[i & 15][i >> 4] = rb[t]; else sB[i >> 6][i & 63] = rb[t];
      }
      __syncthreads();
      if (k0 + 16 < k_hi) fetch(k0 + 16);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        float a[4], bb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) bb[j] = sB[k][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
      }
      __syncthreads();
    }
  }
  float* Cb = p.C + (p.reduce_batch ? 0 : (size_t)blockIdx.z * p.c_bs);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (m < p.M && n < p.N) {
        float* c = Cb + (size_t)m * p.ldc + n;
        if (p.reduce_batch) atomicAdd(c, p.alpha * acc[i][j]);  // beta == 1 (checked by the launcher): C accumulates
        else *c = p.alpha * acc[i][j] + (p.beta != 0.f ? p.beta * *c : 0.f);
      }
    }
  }
}

// out[j] (+)= sum_r in[r * ld + j]
__global__ void sum_rows_kernel(const float* __restrict__ in, int64_t ld, int rows, int n, float* __restrict__ out,
                                int accumulate) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int r = 0; r < rows; ++r) s += in[(size_t)r * ld + j];
  out[j] = (accumulate ? out[j] : 0.f) + s;
}

// in [B, C, L]: out[c] (+)= sum_b sum_l a[b,c,l] (* b2[b,c,l] when given); one CTA per channel
__global__ void __launch_bounds__(256) sum_last_kernel(const float* __restrict__ a, const float* __restrict__ b2, int B,
                                                       int Cc, int L, float* __restrict__ out, int accumulate) {
  __shared__ float red[8];
  const int c = blockIdx.x;
  float s = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* r = a + ((size_t)b * Cc + c) * L;
    const float* r2 = b2 ? b2 + ((size_t)b * Cc + c) * L : nullptr;
    for (int i = threadIdx.x; i < L; i += 256) s += r2 ? r[i] * r2[i] : r[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[c] = (accumulate ? out[c] : 0.f) + t;
  }
}

// ------------------------------------------------------------------------------------------------ projection MLP
__device__ __forceinline__ float block_sum_256(float v, float* red, float* bc) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    *bc = t;
  }
  __syncthreads();
  return *bc;
}

// y = relu(LN(x) w + b): dx, and dw / db accumulated with atomics (rows is the batch, <= a few dozen); one CTA per row
__global__ void __launch_bounds__(256) ln_relu_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bvec, const float* __restrict__ dy,
                                                          float* __restrict__ dx, float* __restrict__ dw,
                                                          float* __restrict__ db, int n, float eps) {
  __shared__ float red[8];
  __shared__ float bc;
  const float* xr = x + (size_t)blockIdx.x * n;
  const float* dyr = dy + (size_t)blockIdx.x * n;
  float* dxr = dx + (size_t)blockIdx.x * n;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += xr[i];
  const float mean = block_sum_256(s, red, &bc) / n;
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float c = xr[i] - mean;
    v += c * c;
  }
  const float rstd = rsqrtf(block_sum_256(v, red, &bc) / n + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float xh = (xr[i] - mean) * rstd;
    const float pre = xh * w[i] + bvec[i];
    const float g = pre > 0.f ? dyr[i] : 0.f;
    if (g != 0.f) {
      atomicAdd(&dw[i], g * xh);
      atomicAdd(&db[i], g);
    }
    const float dxh = g * w[i];
    s1 += dxh;
    s2 += dxh * xh;
  }
  const float m1 = block_sum_256(s1, red, &bc) / n;
  const float m2 = block_sum_256(s2, red, &bc) / n;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float xh = (xr[i] - mean) * rstd;
    const float pre = xh * w[i] + bvec[i];
    const float dxh = pre > 0.f ? dyr[i] * w[i] : 0.f;
    dxr[i] = rstd * (dxh - m1 - xh * m2);
  }
}

// counter-based Bernoulli mask: element i of stream `seed` is kept with probability 1 - p (splitmix64 finaliser)
__device__ __forceinline__ bool dropout_keep(uint64_t seed, uint64_t i, float p) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ull + i + 0x632BE59BD9B4E019ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(z >> 40) * (1.0f / 16777216.0f) >= p;
}
__global__ void dropout_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, uint64_t seed) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  y[i] = dropout_keep(seed, (uint64_t)i, p) ? x[i] * (1.0f / (1.0f - p)) : 0.f;
}

// Unflatten(tokens, 8) -> Linear(8, n_out): dh[b, 8t + k] = sum_o dout[b, t, o] W[o, k]; one CTA per (t, b)
__global__ void __launch_bounds__(256) token_linear_dh_kernel(const float* __restrict__ dout, const float* __restrict__ W,
                                                              float* __restrict__ dh, int tokens, int n_out,
                                                              int64_t out_bstride) {
  __shared__ float red[8][8];
  const int t = blockIdx.x, b = blockIdx.y;
  const float* d = dout + (size_t)b * out_bstride + (size_t)t * n_out;
  float acc[8] = {};
  for (int o = threadIdx.x; o < n_out; o += 256) {
    const float g = d[o];
    const float4 w0 = *reinterpret_cast<const float4*>(W + (size_t)o * 8);
    const float4 w1 = *reinterpret_cast<const float4*>(W + (size_t)o * 8 + 4);
    acc[0] += g * w0.x; acc[1] += g * w0.y; acc[2] += g * w0.z; acc[3] += g * w0.w;
    acc[4] += g * w1.x; acc[5] += g * w1.y; acc[6] += g * w1.z; acc[7] += g * w1.w;
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float s = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float s = 0.f;
    for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
    dh[((size_t)b * tokens + t) * 8 + threadIdx.x] = s;
  }
}
// dW[o, k] += sum_{b,t} dout[b,t,o] h[b, 8t + k]; dbias[o] += sum_{b,t} dout[b,t,o]; one thread per o, `rows_per` (b,t)
// rows per CTA along grid.y, atomics across those slices
__global__ void __launch_bounds__(256) token_linear_dw_kernel(const float* __restrict__ dout, const float* __restrict__ h,
                                                              float* __restrict__ dW, float* __restrict__ dbias, int B,
                                                              int tokens, int n_out, int64_t out_bstride, int rows_per) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const int r0 = blockIdx.y * rows_per, r1 = min(B * tokens, r0 + rows_per);
  float acc[8] = {}, accb = 0.f;
  for (int r = r0; r < r1; ++r) {
    const int b = r / tokens, t = r - b * tokens;
    const float g = dout[(size_t)b * out_bstride + (size_t)t * n_out + o];
    const float* hv = h + (size_t)r * 8;
    accb += g;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += g * hv[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) atomicAdd(&dW[(size_t)o * 8 + k], acc[k]);
  atomicAdd(&dbias[o], accb);
}

// AdaptiveAvgPool1d backward: din[b,c,l] += dfeat[b*bs + c*cs + i*is + off] / (bin length) for every l of bin i
__global__ void adaptive_pool_bwd_kernel(const float* __restrict__ dfeat, float* __restrict__ din, int Cc, int L, int O,
                                         int64_t f_bstride, int cs, int is, int off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (i >= O) return;
  const int s = (int)(((int64_t)i * L) / O);
  const int e = (int)((((int64_t)(i + 1)) * L + O - 1) / O);
  const float g = dfeat[(size_t)b * f_bstride + (size_t)c * cs + (size_t)i * is + off] / (float)(e - s);
  float* x = din + ((size_t)b * Cc + c) * L;
  for (int l = s; l < e; ++l) atomicAdd(&x[l], g);
}

// ------------------------------------------------------------------------------------------------ S4 model
constexpr int CLB_MAX = 64;
// y = LN_j(v) w + b over the d channels at each (b, l): dv; dw / db accumulated (shared-memory, then global atomics)
__global__ void __launch_bounds__(128) channel_ln_bwd_kernel(const float* __restrict__ v, const float* __restrict__ w,
                                                             const float* __restrict__ dy, float* __restrict__ dv,
                                                             float* __restrict__ dw, float* __restrict__ db, int d, int L,
                                                             float eps) {
  __shared__ float sdw[CLB_MAX], sdb[CLB_MAX];
  if (threadIdx.x < CLB_MAX) sdw[threadIdx.x] = sdb[threadIdx.x] = 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < L) {
    float x[CLB_MAX], g[CLB_MAX];
    float mean = 0.f;
#pragma unroll
    for (int j = 0; j < CLB_MAX; ++j)
      if (j < d) {
        x[j] = v[((size_t)b * d + j) * L + l];
        g[j] = dy[((size_t)b * d + j) * L + l];
        mean += x[j];
      }
    mean /= d;
    float var = 0.f;
#pragma unroll
    for (int j = 0; j < CLB_MAX; ++j)
      if (j < d) var += (x[j] - mean) * (x[j] - mean);
    const float rstd = rsqrtf(var / d + eps);
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int j = 0; j < CLB_MAX; ++j)
      if (j < d) {
        x[j] = (x[j] - mean) * rstd;  // x-hat
        atomicAdd(&sdw[j], g[j] * x[j]);
        atomicAdd(&sdb[j], g[j]);
        g[j] *= w[j];
        m1 += g[j];
        m2 += g[j] * x[j];
      }
    m1 /= d;
    m2 /= d;
#pragma unroll
    for (int j = 0; j < CLB_MAX; ++j)
      if (j < d) dv[((size_t)b * d + j) * L + l] = rstd * (g[j] - m1 - x[j] * m2);
  }
  __syncthreads();
  if (threadIdx.x < d) {
    atomicAdd(&dw[threadIdx.x], sdw[threadIdx.x]);
    atomicAdd(&db[threadIdx.x], sdb[threadIdx.x]);
  }
}

__device__ __forceinline__ float gelu_erf_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// ds = dg * d/ds gelu_erf(s)
__global__ void gelu_erf_bwd_kernel(const float* __restrict__ s, const float* __restrict__ dg, float* __restrict__ ds,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = s[i];
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * expf(-0.5f * x * x);
  ds[i] = dg[i] * (cdf + x * pdf);
}

constexpr int CONVB_T = 128;
// dK[c, j] (+)= sum_b sum_{l >= j} ds[b,c,l] h[b,c,l-j]; one CTA = 128 lags of one channel
__global__ void __launch_bounds__(CONVB_T) s4_conv_wgrad_kernel(const float* __restrict__ ds, const float* __restrict__ h,
                                                                float* __restrict__ dK, int B, int d, int L, int accumulate) {
  __shared__ float sD[CONVB_T];
  __shared__ float sH[2 * CONVB_T];
  const int c = blockIdx.y;
  const int j0 = blockIdx.x * CONVB_T;
  const int t = threadIdx.x;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* dsb = ds + ((size_t)b * d + c) * L;
    const float* hb = h + ((size_t)b * d + c) * L;
    // l-chunks [l0, l0+128) with l0 + 127 >= j0
    for (int l0 = (j0 / CONVB_T) * CONVB_T; l0 < L; l0 += CONVB_T) {
      __syncthreads();
      sD[t] = (l0 + t < L) ? dsb[l0 + t] : 0.f;
      const int base = l0 - j0 - (CONVB_T - 1);
      const int i0 = base + t, i1 = base + CONVB_T + t;
      sH[t] = (i0 >= 0 && i0 < L) ? hb[i0] : 0.f;
      sH[CONVB_T + t] = (i1 >= 0 && i1 < L) ? hb[i1] : 0.f;
      __syncthreads();
      // h[l - j] with l = l0 + ll, j = j0 + t: index (l0 + ll - j0 - t) - base = ll - t + 127
#pragma unroll 16
      for (int ll = 0; ll < CONVB_T; ++ll) acc = fmaf(sD[ll], sH[ll - t + (CONVB_T - 1)], acc);
    }
  }
  if (j0 + t < L) {
    float* o = dK + (size_t)c * L + j0 + t;
    *o = (accumulate ? *o : 0.f) + acc;
  }
}

// ---- S4 kernel generation backward (float64) ---------------------------------------------------------------------
struct cdd {
  double x, y;
};
__device__ __forceinline__ cdd zmul(cdd a, cdd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cdd zadd(cdd a, cdd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cdd zsub(cdd a, cdd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cdd zconj(cdd a) { return {a.x, -a.y}; }
__device__ __forceinline__ cdd zdiv(cdd a, cdd b) {
  const double den = b.x * b.x + b.y * b.y;
  return {(a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den};
}
__device__ __forceinline__ cdd zf(float2 v) { return {(double)v.x, (double)v.y}; }

// G_at[c, l] = (1/L) sum_j dK[c, j] exp(-2 pi i j l / L)   (gradient convention dL/dRe + i dL/dIm, as torch)
__global__ void s4_dft_grad_kernel(const float* __restrict__ dK, double2* __restrict__ G, int L) {
  extern __shared__ double2 tw[];  // exp(-2 pi i m / L)
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    double sn, cs;
    sincospi(-2.0 * (double)m / (double)L, &sn, &cs);
    tw[m] = make_double2(cs, sn);
  }
  __syncthreads();
  const int c = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= L) return;
  const float* g = dK + (size_t)c * L;
  double ar = 0.0, ai = 0.0;
  int idx = 0;
  for (int j = 0; j < L; ++j) {
    const double v = (double)g[j];
    ar += v * tw[idx].x;
    ai += v * tw[idx].y;
    idx += l;
    if (idx >= L) idx -= L;
  }
  G[(size_t)c * L + l] = make_double2(ar / L, ai / L);
}

// per (c, l): recompute the four Cauchy sums and their g-derivatives; emit the gradients of k00, k01, k10 (which carry
// the parameters B / Ct) and this root's contribution to d log_step.  Gk [d, L, 3] complex128, dls [d, L] float64.
__global__ void s4_cauchy_bwd_kernel(const float2* __restrict__ lam, const float2* __restrict__ p, const float2* __restrict__ q,
                                     const float2* __restrict__ Bm, const float2* __restrict__ Ct,
                                     const float* __restrict__ log_step, const double2* __restrict__ G,
                                     double2* __restrict__ Gk, double* __restrict__ dls, int d, int n, int L) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (l >= L) return;
  const double step = exp((double)log_step[c]);
  const cdd Gat = {G[(size_t)c * L + l].x, G[(size_t)c * L + l].y};
  double2* gk = Gk + ((size_t)c * L + l) * 3;
  if (2 * l == L) {
    // at = step/2 * sum_n conj(Ct_n) B_n: handled in the (c, n) kernel; d at / d log_step = at
    cdd s = {0, 0};
    for (int j = 0; j < n; ++j) s = zadd(s, zmul(zconj(zf(Ct[c * n + j])), zf(Bm[c * n + j])));
    const cdd at = {0.5 * step * s.x, 0.5 * step * s.y};
    dls[(size_t)c * L + l] = Gat.x * at.x + Gat.y * at.y;  // Re(conj(G) at)
    gk[0] = gk[1] = gk[2] = make_double2(0.0, 0.0);
    return;
  }
  double sn, cs;
  sincospi(-2.0 * (double)l / (double)L, &sn, &cs);
  const cdd w = {cs, sn}, one = {1.0, 0.0};
  const cdd opw = zadd(one, w);
  cdd g = zdiv(zsub(one, w), opw);
  g.x *= 2.0 / step;
  g.y *= 2.0 / step;
  const cdd cc = zdiv({2.0, 0.0}, opw);
  cdd k00 = {0, 0}, k01 = {0, 0}, k10 = {0, 0}, k11 = {0, 0};
  cdd e00 = {0, 0}, e01 = {0, 0}, e10 = {0, 0}, e11 = {0, 0};  // sums with inv^2 (= -d k / d g)
  for (int j = 0; j < n; ++j) {
    const cdd lj = zf(lam[j]), pj = zf(p[j]), qj = zf(q[j]);
    const cdd ct = zf(Ct[c * n + j]), b = zf(Bm[c * n + j]);
    const cdd inv = zdiv(one, zsub(g, lj));
    const cdd inv2 = zmul(inv, inv);
    const cdd a0 = zconj(ct), a1 = zconj(qj);
    const cdd t00 = zmul(a0, b), t01 = zmul(a0, pj), t10 = zmul(a1, b), t11 = zmul(a1, pj);
    k00 = zadd(k00, zmul(t00, inv)); e00 = zadd(e00, zmul(t00, inv2));
    k01 = zadd(k01, zmul(t01, inv)); e01 = zadd(e01, zmul(t01, inv2));
    k10 = zadd(k10, zmul(t10, inv)); e10 = zadd(e10, zmul(t10, inv2));
    k11 = zadd(k11, zmul(t11, inv)); e11 = zadd(e11, zmul(t11, inv2));
  }
  const cdd r = zdiv(one, zadd(one, k11));  // 1 / (1 + k11)
  // at = cc (k00 - k01 r k10): holomorphic partials
  const cdd d00 = cc;
  const cdd d01 = zmul(cc, zmul(r, k10)); // with a minus sign below
  const cdd d10 = zmul(cc, zmul(r, k01));
  const cdd d11 = zmul(cc, zmul(zmul(k01, k10), zmul(r, r)));  // + k01 k10 r^2
  // gradient through a holomorphic map: G_in = conj(d out / d in) G_out
  const cdd G00 = zmul(zconj(d00), Gat);
  cdd G01 = zmul(zconj(d01), Gat);
  G01.x = -G01.x; G01.y = -G01.y;
  cdd G10 = zmul(zconj(d10), Gat);
  G10.x = -G10.x; G10.y = -G10.y;
  gk[0] = make_double2(G00.x, G00.y);
  gk[1] = make_double2(G01.x, G01.y);
  gk[2] = make_double2(G10.x, G10.y);
  // d at / d g = cc( -e00 + (e01 k10 + k01 e10) r - k01 k10 r^2 e11 ),   d g / d log_step = -g
  cdd datdg = zsub(zmul(zadd(zmul(e01, k10), zmul(k01, e10)), r), e00);
  datdg = zsub(datdg, zmul(zmul(zmul(k01, k10), zmul(r, r)), e11));
  datdg = zmul(cc, datdg);
  (void)d11;
  const cdd dat = zmul(datdg, {-g.x, -g.y});
  dls[(size_t)c * L + l] = Gat.x * dat.x + Gat.y * dat.y;
}

// per (c, n): sum over the roots.  dB[c,n] += conj(a0) S00 + conj(a1) S10;  dCt[c,n] += conj( conj(b) S00 + conj(p) S01 )
// with S.. = sum_l conj(inv_n(l)) G_k..(l), plus the w = -1 root.  One CTA per (n, c): 128 threads stride over l, fixed-order
// block reduction of the three complex sums.  Gradients are complex64 (float2), accumulated.
__global__ void __launch_bounds__(128) s4_param_grad_kernel(const float2* __restrict__ lam, const float2* __restrict__ p,
                                                            const float2* __restrict__ q, const float2* __restrict__ Bm,
                                                            const float2* __restrict__ Ct, const float* __restrict__ log_step,
                                                            const double2* __restrict__ G, const double2* __restrict__ Gk,
                                                            float2* __restrict__ dB, float2* __restrict__ dCt, int d, int n,
                                                            int L) {
  __shared__ double red[4][6];
  const int j = blockIdx.x, c = blockIdx.y;
  const double step = exp((double)log_step[c]);
  const cdd lj = zf(lam[j]), pj = zf(p[j]), qj = zf(q[j]);
  const cdd ct = zf(Ct[c * n + j]), b = zf(Bm[c * n + j]);
  const cdd one = {1.0, 0.0};
  cdd S00 = {0, 0}, S01 = {0, 0}, S10 = {0, 0};
  for (int l = threadIdx.x; l < L; l += 128) {
    if (2 * l == L) continue;
    double sn, cs;
    sincospi(-2.0 * (double)l / (double)L, &sn, &cs);
    const cdd w = {cs, sn};
    cdd g = zdiv(zsub(one, w), zadd(one, w));
    g.x *= 2.0 / step;
    g.y *= 2.0 / step;
    const cdd cinv = zconj(zdiv(one, zsub(g, lj)));
    const double2* gk = Gk + ((size_t)c * L + l) * 3;
    S00 = zadd(S00, zmul(cinv, {gk[0].x, gk[0].y}));
    S01 = zadd(S01, zmul(cinv, {gk[1].x, gk[1].y}));
    S10 = zadd(S10, zmul(cinv, {gk[2].x, gk[2].y}));
  }
  double v[6] = {S00.x, S00.y, S01.x, S01.y, S10.x, S10.y};
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v[i];
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  for (int i = 0; i < 6; ++i) v[i] = ((red[0][i] + red[1][i]) + red[2][i]) + red[3][i];
  S00 = {v[0], v[1]};
  S01 = {v[2], v[3]};
  S10 = {v[4], v[5]};
  const cdd a0 = zconj(ct), a1 = zconj(qj);
  cdd gB = zadd(zmul(zconj(a0), S00), zmul(zconj(a1), S10));
  cdd gA0 = zadd(zmul(zconj(b), S00), zmul(zconj(pj), S01));
  if (L % 2 == 0) {  // the w = -1 root: at = step/2 sum conj(Ct) B
    const cdd Gat = {G[(size_t)c * L + L / 2].x, G[(size_t)c * L + L / 2].y};
    const cdd hs = {0.5 * step, 0.0};
    gB = zadd(gB, zmul(zconj(zmul(hs, a0)), Gat));
    gA0 = zadd(gA0, zmul(zconj(zmul(hs, b)), Gat));
  }
  const cdd gCt = zconj(gA0);  // a0 = conj(Ct)
  dB[c * n + j].x += (float)gB.x;
  dB[c * n + j].y += (float)gB.y;
  dCt[c * n + j].x += (float)gCt.x;
  dCt[c * n + j].y += (float)gCt.y;
}

__global__ void __launch_bounds__(256) sum_f64_rows_kernel(const double* __restrict__ in, int L, float* __restrict__ out) {
  __shared__ double red[8];
  const int c = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < L; i += 256) s += in[(size_t)c * L + i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += red[i];
    out[c] += (float)t;
  }
}

// ------------------------------------------------------------------------------------------------ DUAN backward
// per (b, ch) row: r1 = sum_l mask dy, r2 = sum_l mask dy x-hat
__global__ void __launch_bounds__(256) duan_bwd_rows_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            int64_t dy_bstride, const float* __restrict__ mu,
                                                            const float* __restrict__ rsig, const float* __restrict__ mask,
                                                            float* __restrict__ r1, float* __restrict__ r2, int Cc, int L) {
  __shared__ float red[8];
  __shared__ float bc;
  const int row = blockIdx.x;
  const int b = row / Cc, ch = row - b * Cc;
  const float* xr = x + (size_t)row * L;
  const float* dr = dy + (size_t)b * dy_bstride + (size_t)ch * L;
  const float m = mu[row], rs = rsig[row], mk = mask[row];
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) {
    const float g = mk * dr[i];
    s1 += g;
    s2 += g * ((xr[i] - m) * rs);
  }
  const float t1 = block_sum_256(s1, red, &bc);
  const float t2 = block_sum_256(s2, red, &bc);
  if (threadIdx.x == 0) {
    r1[row] = t1;
    r2[row] = t2;
  }
}

// one CTA per batch element: statistics-path gradients.  Outputs per (b, ch): coefA, coefB (dx = mask g1 rsig dy + A + B x),
// dg (gradient of the gate mean g_mix) and dgb = [d gamma | d beta] (the FiLM MLP's output gradient).
__global__ void duan_bwd_mix_kernel(const float* __restrict__ mean_x, const float* __restrict__ m2_x,
                                    const float* __restrict__ g_mix, const float* __restrict__ rsig,
                                    const float* __restrict__ g1, const float* __restrict__ r1, const float* __restrict__ r2,
                                    float* __restrict__ coefA, float* __restrict__ coefB, float* __restrict__ dg,
                                    float* __restrict__ dgb, int Cc, int L, float eps) {
  __shared__ double sh[4];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int ch = 0; ch < Cc; ++ch) tot += (double)mean_x[b * Cc + ch];
    const double mu_l = tot / Cc;
    double m2 = 0.0;
    for (int ch = 0; ch < Cc; ++ch) {
      const double dm = (double)mean_x[b * Cc + ch] - mu_l;
      m2 += (double)m2_x[b * Cc + ch] + dm * dm * L;
    }
    const double sigma_l = sqrt(m2 / ((double)Cc * L) + (double)eps);
    double dmu_l = 0.0, dsig_l = 0.0;
    for (int ch = 0; ch < Cc; ++ch) {
      const int i = b * Cc + ch;
      const double g = g_mix[i];
      const double dmu = -(double)rsig[i] * g1[i] * r1[i];
      const double dsg = -(double)rsig[i] * g1[i] * r2[i];
      dmu_l += (1.0 - g) * dmu;
      dsig_l += (1.0 - g) * dsg;
    }
    sh[0] = mu_l; sh[1] = sigma_l; sh[2] = dmu_l; sh[3] = dsig_l;
  }
  __syncthreads();
  const float mu_l = (float)sh[0], sigma_l = (float)sh[1], dmu_l = (float)sh[2], dsig_l = (float)sh[3];
  const float invL = 1.0f / L, invCL = 1.0f / ((float)Cc * L);
  for (int ch = threadIdx.x; ch < Cc; ch += blockDim.x) {
    const int i = b * Cc + ch;
    const float g = g_mix[i];
    const float mu_c = mean_x[i];
    const float sigma_c = sqrtf(m2_x[i] * invL + eps);
    const float dmu = -rsig[i] * g1[i] * r1[i];
    const float dsg = -rsig[i] * g1[i] * r2[i];
    dg[i] = dmu * (mu_c - mu_l) + dsg * (sigma_c - sigma_l);
    const float dmu_c = g * dmu, dsig_c = g * dsg;
    coefB[i] = dsig_c / sigma_c * invL + dsig_l / sigma_l * invCL;
    coefA[i] = dmu_c * invL - dsig_c * mu_c / sigma_c * invL + dmu_l * invCL - dsig_l * mu_l / sigma_l * invCL;
    dgb[(size_t)b * 2 * Cc + ch] = r2[i];        // d gamma
    dgb[(size_t)b * 2 * Cc + Cc + ch] = r1[i];   // d beta
  }
}

// dx[b,ch,l] (+)= mask g1 rsig dy + A + B x;   dc[b,ch,l] (+)= dmean_c[b,ch] / L
__global__ void duan_bwd_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t dy_bstride,
                                      const float* __restrict__ rsig, const float* __restrict__ g1,
                                      const float* __restrict__ mask, const float* __restrict__ coefA,
                                      const float* __restrict__ coefB, const float* __restrict__ dmean_c,
                                      float* __restrict__ dx, float* __restrict__ dc, int Cc, int L, int acc_dx, int acc_dc) {
  const int row = blockIdx.y;
  const int b = row / Cc, ch = row - b * Cc;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const size_t o = (size_t)row * L + i;
  if (dx != nullptr) {
    const float v = mask[row] * g1[row] * rsig[row] * dy[(size_t)b * dy_bstride + (size_t)ch * L + i] + coefA[row] +
                    coefB[row] * x[o];
    dx[o] = (acc_dx ? dx[o] : 0.f) + v;
  }
  if (dc != nullptr) dc[o] = (acc_dc ? dc[o] : 0.f) + dmean_c[row] / (float)L;
}

// in place: s[b,ch,l] (= sigmoid output) -> d a = (dg[b,ch] / L) s (1 - s)
__global__ void duan_bwd_sigmoid_kernel(float* __restrict__ s, const float* __restrict__ dg, int L, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = s[i];
  s[i] = dg[i / L] / (float)L * v * (1.0f - v);
}
// in place: d[i] = act[i] > 0 ? d[i] : 0
__global__ void relu_mask_kernel(float* __restrict__ d, const float* __restrict__ act, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(act[i] > 0.f)) d[i] = 0.f;
}

// d_in[b, k] += sum_n d[b, n] W[n, k] for a bf16 panel W [N, K] streamed once (the AdaLN / time-text-embedding Linears of
// the DiT: gradient of the conditioning vectors; the stacked AdaLN panels are 2.1 GB each), B <= 8 rows.  CTA = rows_per
// rows x 1024 columns: a thread owns 8 adjacent columns (16-byte loads, four rows in flight), the d values of the row chunk
// are staged in shared memory (broadcast reads), one 16-byte atomicAdd pair per (b, thread) at the end.
// (First form: 4-byte loads, one row at a time, d re-read from global per row: 1.5 TB/s.)
constexpr int SKINNY_MAXB = 8;
constexpr int SKINNY_ROWS = 256;  // rows of d staged at a time
__global__ void __launch_bounds__(128) skinny_xw_bf16_kernel(const float* __restrict__ d, int64_t ldd,
                                                             const __nv_bfloat16* __restrict__ W, int64_t ldw,
                                                             float* __restrict__ out, int64_t ldo, int B, int N, int K,
                                                             int rows_per) {
  __shared__ float ds[SKINNY_MAXB][SKINNY_ROWS];
  const int n0 = blockIdx.x * rows_per, n1 = min(N, n0 + rows_per);
  const int k0 = (blockIdx.y * 128 + threadIdx.x) * 8;
  const bool live = k0 < K;  // (K is a multiple of 8)
  float acc[SKINNY_MAXB][8];
#pragma unroll
  for (int b = 0; b < SKINNY_MAXB; ++b)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[b][e] = 0.f;
  for (int nb = n0; nb < n1; nb += SKINNY_ROWS) {
    const int cnt = min(SKINNY_ROWS, n1 - nb);
    __syncthreads();
    for (int i = threadIdx.x; i < SKINNY_MAXB * SKINNY_ROWS; i += 128) {
      const int b = i / SKINNY_ROWS, r = i % SKINNY_ROWS;
      ds[b][r] = (b < B && r < cnt) ? d[(size_t)b * ldd + nb + r] : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int r = 0; r < cnt; r += 4) {
      uint4 w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)  // rows past the end re-read the last row; their d values are zero
        w[i] = __ldg(reinterpret_cast<const uint4*>(W + (size_t)(nb + min(r + i, cnt - 1)) * ldw + k0));
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 a = unpack_bf16(w[i].x), bq = unpack_bf16(w[i].y), c = unpack_bf16(w[i].z), e4 = unpack_bf16(w[i].w);
        const float wf[8] = {a.x, a.y, bq.x, bq.y, c.x, c.y, e4.x, e4.y};
        const int rr = (r + i < cnt) ? r + i : SKINNY_ROWS - 1;  // a staged zero (cnt < SKINNY_ROWS there) or, for a full
        const float keep = (r + i < cnt) ? 1.f : 0.f;            // chunk, masked explicitly
#pragma unroll
        for (int b = 0; b < SKINNY_MAXB; ++b) {
          const float g = ds[b][rr] * keep;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[b][e] = fmaf(g, wf[e], acc[b][e]);
        }
      }
    }
  }
  if (!live) return;
#pragma unroll
  for (int b = 0; b < SKINNY_MAXB; ++b)
    if (b < B) {
      float* o = out + (size_t)b * ldo + k0;
      atomicAdd(reinterpret_cast<float4*>(o), make_float4(acc[b][0], acc[b][1], acc[b][2], acc[b][3]));
      atomicAdd(reinterpret_cast<float4*>(o + 4), make_float4(acc[b][4], acc[b][5], acc[b][6], acc[b][7]));
    }
}

// dpre[i] = dy[i] * silu'(pre[i]) with bf16 pre-activations summed from up to three bf16 sources (row-broadcast for c)
__global__ void silu_bwd_sum_kernel(const float* __restrict__ dy, const __nv_bfloat16* __restrict__ a,
                                    const __nv_bfloat16* __restrict__ b2, const __nv_bfloat16* __restrict__ c, int c_rows,
                                    float* __restrict__ dpre, int M, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * D) return;
  const int m = i / D, k = i - m * D;
  // the forward rounds the sum to bf16 after every add (add_silu_kernel adds in fp32 once): same fp32 sum here
  float x = __bfloat162float(a[i]);
  if (b2 != nullptr) x += __bfloat162float(b2[i]);
  if (c != nullptr) x += __bfloat162float(c[(size_t)(m % c_rows) * D + k]);
  const float sg = 1.0f / (1.0f + __expf(-x));
  dpre[i] = dy[i] * sg * (1.0f + x * (1.0f - sg));
}
__global__ void silu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ pre, float* __restrict__ dpre,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = pre[i];
  const float sg = 1.0f / (1.0f + __expf(-x));
  dpre[i] = dy[i] * sg * (1.0f + x * (1.0f - sg));
}

}  // namespace lx

using namespace lx;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int lx_sgemm_ex(const lx_sgemm_ex_desc_t* desc, void* stream) {
  LX_CHECK_ARG(desc != nullptr, "lx_sgemm_ex: null descriptor");
  const lx_sgemm_ex_desc_t& d = *desc;
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * d.batch * ((double)d.M * d.K + (double)d.K * d.N) + 4.0 * (double)d.M * d.N);
  LX_CHECK_ARG(d.A && d.Bm && d.C && d.M > 0 && d.N > 0 && d.K > 0 && d.batch > 0, "lx_sgemm_ex: bad arguments");
  SgemmEx p;
  p.A = d.A; p.Bm = d.Bm; p.C = d.C;
  p.lda = d.lda; p.ldb = d.ldb; p.ldc = d.ldc; p.a_bs = d.a_bstride; p.b_bs = d.b_bstride; p.c_bs = d.c_bstride;
  p.M = d.M; p.N = d.N; p.K = d.K; p.batch = d.batch;
  p.trans_a = d.trans_a; p.trans_b = d.trans_b; p.reduce_batch = d.reduce_batch;
  p.alpha = d.alpha; p.beta = d.beta;
  p.ksplit = 1;
  if (d.reduce_batch) {
    LX_CHECK_ARG(d.beta == 1.0f, "lx_sgemm_ex: reduce_batch accumulates into C (beta must be 1)");
    // enough CTAs to fill the GPU a few times over: tiles x batch x K slices >= ~4 x SMs, slices of at least 64
    const long tiles = (long)((d.N + 63) / 64) * ((d.M + 63) / 64) * d.batch;
    const long want = 4L * num_sms();
    while (tiles * p.ksplit < want && d.K / (p.ksplit * 2) >= 64) p.ksplit *= 2;
  }
  {  // the larger products go to the 128 x 128 tile (host_util.cuh: Sgemm128) when that still fills the GPU
    long tiles128 = (long)((d.N + 127) / 128) * ((d.M + 127) / 128) * d.batch;
    int ks = 1;
    if (d.reduce_batch)
      while (tiles128 * ks < 2L * num_sms() && d.K / (ks * 2) >= 64) ks *= 2;
    if (d.M >= 128 && d.N >= 128 && tiles128 * ks >= num_sms()) {
      Sgemm128 q{};
      q.A = d.A; q.B = d.Bm; q.C = d.C;
      q.lda = d.lda; q.ldb = d.ldb; q.ldc = d.ldc; q.a_bs = d.a_bstride; q.b_bs = d.b_bstride; q.c_bs = d.c_bstride;
      q.M = d.M; q.N = d.N; q.K = d.K; q.trans_a = d.trans_a; q.trans_b = d.trans_b;
      q.reduce_batch = d.reduce_batch; q.ksplit = ks; q.alpha = d.alpha; q.beta = d.beta;
      return sgemm128_launch(q, d.batch, stream);
    }
  }
  dim3 grid((d.N + 63) / 64, (d.M + 63) / 64, d.reduce_batch ? d.batch * p.ksplit : d.batch);
  sgemm_ex_kernel<<<grid, 256, 0, ST(stream)>>>(p);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_sum_rows_f32(const float* in, int64_t ld, int32_t rows, int32_t n, float* out, int32_t accumulate,
                               void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * rows * (double)n);
  LX_CHECK_ARG(in && out && rows > 0 && n > 0, "lx_sum_rows_f32: bad arguments");
  sum_rows_kernel<<<(n + 255) / 256, 256, 0, ST(stream)>>>(in, ld, rows, n, out, accumulate);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_sum_last_f32(const float* a, const float* b, int32_t B, int32_t C, int32_t L, float* out,
                               int32_t accumulate, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * B * C * (double)L * (b ? 2 : 1));
  LX_CHECK_ARG(a && out && B > 0 && C > 0 && L > 0, "lx_sum_last_f32: bad arguments");
  sum_last_kernel<<<C, 256, 0, ST(stream)>>>(a, b, B, C, L, out, accumulate);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_ln_relu_rows_bwd(const float* x, const float* w, const float* b, const float* dy, float* dx, float* dw,
                                   float* db, int32_t rows, int32_t n, float eps, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 16.0 * rows * (double)n);
  LX_CHECK_ARG(x && w && b && dy && dx && dw && db && rows > 0 && n > 0, "lx_ln_relu_rows_bwd: bad arguments");
  ln_relu_bwd_kernel<<<rows, 256, 0, ST(stream)>>>(x, w, b, dy, dx, dw, db, n, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 8.0 * n);
  LX_CHECK_ARG(x && y && n > 0 && p >= 0.f && p < 1.f, "lx_dropout_f32: bad arguments");
  dropout_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(x, y, n, p, seed);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_token_linear_bwd(const float* h, const float* W, const float* dout, float* dh, float* dW, float* dbias,
                                   int32_t B, int32_t tokens, int32_t n_out, int64_t out_bstride, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 8.0 * B * tokens * (double)n_out);
  LX_CHECK_ARG(h && W && dout && dh && dW && dbias && B > 0 && tokens > 0 && n_out > 0, "lx_token_linear_bwd: bad arguments");
  token_linear_dh_kernel<<<dim3(tokens, B), 256, 0, ST(stream)>>>(dout, W, dh, tokens, n_out, out_bstride);
  LX_CUDA(cudaGetLastError());
  const int rows_per = 64;
  dim3 grid((n_out + 255) / 256, (B * tokens + rows_per - 1) / rows_per);
  token_linear_dw_kernel<<<grid, 256, 0, ST(stream)>>>(dout, h, dW, dbias, B, tokens, n_out, out_bstride, rows_per);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_adaptive_pool_bwd(const float* dfeat, float* din, int32_t B, int32_t C, int32_t L, int32_t O,
                                    int64_t f_bstride, int32_t cs, int32_t is, int32_t off, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * B * C * ((double)L + O));
  LX_CHECK_ARG(dfeat && din && B > 0 && C > 0 && L > 0 && O > 0, "lx_adaptive_pool_bwd: bad arguments");
  dim3 grid((O + 127) / 128, C, B);
  adaptive_pool_bwd_kernel<<<grid, 128, 0, ST(stream)>>>(dfeat, din, C, L, O, f_bstride, cs, is, off);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_channel_ln_bwd(const float* v, const float* w, const float* dy, float* dv, float* dw, float* db, int32_t B,
                                 int32_t d, int32_t L, float eps, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 12.0 * B * d * (double)L);
  LX_CHECK_ARG(v && w && dy && dv && dw && db && B > 0 && L > 0 && d > 0 && d <= CLB_MAX, "lx_channel_ln_bwd: bad arguments");
  dim3 grid((L + 127) / 128, B);
  channel_ln_bwd_kernel<<<grid, 128, 0, ST(stream)>>>(v, w, dy, dv, dw, db, d, L, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_gelu_erf_bwd(const float* s, const float* dg, float* ds, int64_t n, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 12.0 * n);
  LX_CHECK_ARG(s && dg && ds && n > 0, "lx_gelu_erf_bwd: bad arguments");
  gelu_erf_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(s, dg, ds, n);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_s4_conv_wgrad(const float* ds, const float* h, float* dK, int32_t B, int32_t d, int32_t L,
                                int32_t accumulate, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 8.0 * B * d * (double)L + 4.0 * d * (double)L);
  LX_CHECK_ARG(ds && h && dK && B > 0 && d > 0 && L > 0, "lx_s4_conv_wgrad: bad arguments");
  dim3 grid((L + CONVB_T - 1) / CONVB_T, d);
  s4_conv_wgrad_kernel<<<grid, CONVB_T, 0, ST(stream)>>>(ds, h, dK, B, d, L, accumulate);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_s4_kernel_gen_bwd(const void* lam, const void* p, const void* q, const void* Bm, const void* Ct,
                                    const float* log_step, const float* dK, void* dB, void* dCt, float* dlog_step,
                                    void* workspace, int32_t d, int32_t n, int32_t L, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 72.0 * d * (double)L);
  LX_CHECK_ARG(lam && p && q && Bm && Ct && log_step && dK && dB && dCt && dlog_step && workspace,
               "lx_s4_kernel_gen_bwd: null pointer");
  LX_CHECK_ARG(d > 0 && n > 0 && L > 0 && L * 16 <= 200 * 1024, "lx_s4_kernel_gen_bwd: L=%d too large for the twiddle table", L);
  // workspace (bytes): G [d, L] complex128 | Gk [d, L, 3] complex128 | dls [d, L] float64  = 72 d L
  double2* G = static_cast<double2*>(workspace);
  double2* Gk = G + (size_t)d * L;
  double* dls = reinterpret_cast<double*>(Gk + (size_t)d * L * 3);
  cudaStream_t st = ST(stream);
  const size_t smem = (size_t)L * sizeof(double2);
  static int configured = 0;
  if (smem > 48 * 1024 && (int)smem > configured) {
    LX_CUDA(cudaFuncSetAttribute(s4_dft_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = (int)smem;
  }
  dim3 g1((L + 127) / 128, d);
  if ((L & (L - 1)) == 0 && L >= 2) {  // power of two: G = FFT(dK) / L
    int rc = s4_fft_launch(false, nullptr, dK, G, nullptr, d, L, stream);
    if (rc) return rc;
  } else {
    s4_dft_grad_kernel<<<g1, 128, smem, st>>>(dK, G, L);
    LX_CUDA(cudaGetLastError());
  }
  s4_cauchy_bwd_kernel<<<g1, 128, 0, st>>>((const float2*)lam, (const float2*)p, (const float2*)q, (const float2*)Bm,
                                           (const float2*)Ct, log_step, G, Gk, dls, d, n, L);
  LX_CUDA(cudaGetLastError());
  s4_param_grad_kernel<<<dim3(n, d), 128, 0, st>>>((const float2*)lam, (const float2*)p, (const float2*)q,
                                                              (const float2*)Bm, (const float2*)Ct, log_step, G, Gk,
                                                              (float2*)dB, (float2*)dCt, d, n, L);
  LX_CUDA(cudaGetLastError());
  sum_f64_rows_kernel<<<d, 256, 0, st>>>(dls, L, dlog_step);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

// DUAN backward.  `fwd_ws` is the workspace lx_duan_forward filled for the same (x, c) (statistics, gate mean, FiLM, mask,
// ReLU'd hidden activations); `ws` is scratch: fp32 [B*C*L (sigmoid / d a) + B*hidden*L (d hidden) + B*(8 C + 2 hidden)].
// dx / dc may be NULL (that input needs no gradient); acc_dx / acc_dc != 0 accumulates into them.  The eight weight
// gradients are accumulated (+=).
extern "C" int lx_duan_backward(const lx_duan_weights_t* w, const lx_duan_weights_t* dw, const float* x, const float* c,
                                const float* dy, int64_t dy_bstride, float* dx, float* dc, int32_t acc_dx, int32_t acc_dc,
                                int32_t B, int32_t Cc, int32_t L, const float* fwd_ws, float* ws, void* stream) {
  LX_CHECK_ARG(w && dw && x && c && dy && fwd_ws && ws, "lx_duan_backward: null pointer");
  LX_CHECK_ARG(B > 0 && Cc > 0 && L > 0 && w->hidden > 0 && w->hidden <= 1024, "lx_duan_backward: bad shape");
  const int Hd = w->hidden;
  const size_t BC = (size_t)B * Cc;
  // forward workspace layout (lx_duan_forward)
  const float* mean_x = fwd_ws;
  const float* m2_x = mean_x + BC;
  const float* mean_c = m2_x + BC;
  const float* g_mix = mean_c + BC;
  const float* mu = g_mix + BC;
  const float* rsig = mu + BC;
  const float* g1 = rsig + BC;
  const float* mask = g1 + 3 * BC;          // beta, imp in between
  const float* hid_pool = mask + BC + 2 * BC;  // after gb [B, 2C]
  const float* hid = hid_pool + (size_t)B * Hd;
  // scratch
  float* sg = ws;                                // [B, C, L]
  float* dhid = sg + BC * L;                     // [B, Hd, L]
  float* r1 = dhid + (size_t)B * Hd * L;
  float* r2 = r1 + BC;
  float* coefA = r2 + BC;
  float* coefB = coefA + BC;
  float* dg = coefB + BC;
  float* dgb = dg + BC;                          // [B, 2C]
  float* dmean_c = dgb + 2 * BC;                 // [B, C]
  float* dhp = dmean_c + BC;                     // [B, Hd]
  cudaStream_t st = ST(stream);
  int rc;
  lx_sgemm_ex_desc_t g;
  auto gemm = [&](const float* A, int64_t lda, int64_t abs_, int ta, const float* Bm, int64_t ldb, int64_t bbs, int tb,
                  float* Cm, int64_t ldc, int64_t cbs, int M, int N, int K, int batch, int reduce, float beta) -> int {
    memset(&g, 0, sizeof(g));
    g.A = A; g.lda = lda; g.a_bstride = abs_; g.trans_a = ta; g.Bm = Bm; g.ldb = ldb; g.b_bstride = bbs; g.trans_b = tb;
    g.C = Cm; g.ldc = ldc; g.c_bstride = cbs; g.M = M; g.N = N; g.K = K; g.batch = batch; g.reduce_batch = reduce;
    g.alpha = 1.0f; g.beta = beta;
    return lx_sgemm_ex(&g, stream);
  };
  {
    LaunchScope scope(KC_CS3DGF, stream, 8.0 * BC * L);
    duan_bwd_rows_kernel<<<(unsigned)BC, 256, 0, st>>>(x, dy, dy_bstride, mu, rsig, mask, r1, r2, Cc, L);
    LX_CUDA(cudaGetLastError());
  }
  duan_bwd_mix_kernel<<<B, 256, 0, st>>>(mean_x, m2_x, g_mix, rsig, g1, r1, r2, coefA, coefB, dg, dgb, Cc, L, w->eps);
  LX_CUDA(cudaGetLastError());
  // FiLM MLP: gb = W4 relu(W3 mean_c + b3) + b4
  if ((rc = gemm(dgb, 2 * Cc, 0, 1, hid_pool, Hd, 0, 0, const_cast<float*>(dw->mlp_w2), Hd, 0, 2 * Cc, Hd, B, 1, 0, 1.0f))) return rc;
  if ((rc = lx_sum_rows_f32(dgb, 2 * Cc, B, 2 * Cc, const_cast<float*>(dw->mlp_b2), 1, stream))) return rc;
  if ((rc = gemm(dgb, 2 * Cc, 0, 0, w->mlp_w2, Hd, 0, 0, dhp, Hd, 0, B, Hd, 2 * Cc, 1, 0, 0.0f))) return rc;  // [B, Hd]
  relu_mask_kernel<<<(unsigned)(((size_t)B * Hd + 255) / 256), 256, 0, st>>>(dhp, hid_pool, (int64_t)B * Hd);
  LX_CUDA(cudaGetLastError());
  if ((rc = gemm(dhp, Hd, 0, 1, mean_c, Cc, 0, 0, const_cast<float*>(dw->mlp_w1), Cc, 0, Hd, Cc, B, 1, 0, 1.0f))) return rc;
  if ((rc = lx_sum_rows_f32(dhp, Hd, B, Hd, const_cast<float*>(dw->mlp_b1), 1, stream))) return rc;
  if ((rc = gemm(dhp, Hd, 0, 0, w->mlp_w1, Cc, 0, 0, dmean_c, Cc, 0, B, Cc, Hd, 1, 0, 0.0f))) return rc;      // [B, C]
  // statistics + FiLM-mean paths into dx / dc
  {
    LaunchScope scope(KC_CS3DGF, stream, 16.0 * BC * L);
    dim3 ga((L + 255) / 256, (unsigned)BC);
    duan_bwd_apply_kernel<<<ga, 256, 0, st>>>(x, dy, dy_bstride, rsig, g1, mask, coefA, coefB, dmean_c, dx, dc, Cc, L, acc_dx,
                                              acc_dc);
    LX_CUDA(cudaGetLastError());
  }
  // gate: g_mix = mean_L sigmoid(W2 hid + b2), hid = relu(W1 c + b1)
  lx_sgemm_desc_t f;
  memset(&f, 0, sizeof(f));
  f.A = w->gate_w2; f.lda = Hd; f.Bm = hid; f.ldb = L; f.b_bstride = (int64_t)Hd * L; f.bias = w->gate_b2;
  f.C = sg; f.ldc = L; f.c_bstride = (int64_t)Cc * L; f.M = Cc; f.N = L; f.K = Hd; f.batch = B; f.act = 2;
  if ((rc = lx_sgemm_f32(&f, stream))) return rc;
  duan_bwd_sigmoid_kernel<<<(unsigned)((BC * L + 255) / 256), 256, 0, st>>>(sg, dg, L, (int64_t)(BC * L));  // sg := d a
  LX_CUDA(cudaGetLastError());
  if ((rc = gemm(sg, L, (int64_t)Cc * L, 0, hid, L, (int64_t)Hd * L, 1, const_cast<float*>(dw->gate_w2), Hd, 0, Cc, Hd, L, B, 1, 1.0f))) return rc;
  if ((rc = lx_sum_last_f32(sg, nullptr, B, Cc, L, const_cast<float*>(dw->gate_b2), 1, stream))) return rc;
  if ((rc = gemm(w->gate_w2, Hd, 0, 1, sg, L, (int64_t)Cc * L, 0, dhid, L, (int64_t)Hd * L, Hd, L, Cc, B, 0, 0.0f))) return rc;
  relu_mask_kernel<<<(unsigned)(((size_t)B * Hd * L + 255) / 256), 256, 0, st>>>(dhid, hid, (int64_t)B * Hd * L);
  LX_CUDA(cudaGetLastError());
  if ((rc = gemm(dhid, L, (int64_t)Hd * L, 0, c, L, (int64_t)Cc * L, 1, const_cast<float*>(dw->gate_w1), Cc, 0, Hd, Cc, L, B, 1, 1.0f))) return rc;
  if ((rc = lx_sum_last_f32(dhid, nullptr, B, Hd, L, const_cast<float*>(dw->gate_b1), 1, stream))) return rc;
  if (dc != nullptr)
    if ((rc = gemm(w->gate_w1, Cc, 0, 1, dhid, L, (int64_t)Hd * L, 0, dc, L, (int64_t)Cc * L, Cc, L, Hd, B, 0, 1.0f))) return rc;
  return LX_OK;
}

extern "C" int lx_skinny_xw_bf16(const float* d, int64_t ldd, const void* W, int64_t ldw, float* out, int64_t ldo, int32_t B,
                                 int32_t N, int32_t K, void* stream) {
  LaunchScope scope(KC_ROW, stream, 2.0 * N * (double)K);
  LX_CHECK_ARG(d && W && out && B > 0 && B <= SKINNY_MAXB && N > 0 && K > 0 && K % 8 == 0 && ldw % 8 == 0 && ldo % 4 == 0 &&
                   (uintptr_t)out % 16 == 0 && (uintptr_t)W % 16 == 0,
               "lx_skinny_xw_bf16: bad arguments (B <= %d, K / ldw multiples of 8, out 16-byte aligned)", SKINNY_MAXB);
  const int kchunks = (K + 1023) / 1024;
  const int want = max(1, 6 * num_sms() / kchunks);  // CTAs along the rows: ~6 per SM in total
  int rows_per = (N + want - 1) / want;
  rows_per = max(32, (rows_per + 3) / 4 * 4);
  skinny_xw_bf16_kernel<<<dim3((N + rows_per - 1) / rows_per, kchunks), 128, 0, ST(stream)>>>(
      d, ldd, static_cast<const __nv_bfloat16*>(W), ldw, out, ldo, B, N, K, rows_per);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_silu_bwd_sum(const float* dy, const void* a, const void* b, const void* c, int32_t c_rows, float* dpre,
                               int32_t M, int32_t D, void* stream) {
  LaunchScope scope(KC_ROW, stream, 14.0 * M * (double)D);
  LX_CHECK_ARG(dy && a && dpre && M > 0 && D > 0 && (c == nullptr || c_rows > 0), "lx_silu_bwd_sum: bad arguments");
  silu_bwd_sum_kernel<<<(M * D + 255) / 256, 256, 0, ST(stream)>>>(dy, static_cast<const __nv_bfloat16*>(a),
                                                                   static_cast<const __nv_bfloat16*>(b),
                                                                   static_cast<const __nv_bfloat16*>(c), c_rows, dpre, M, D);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_silu_bwd_f32(const float* dy, const float* pre, float* dpre, int64_t n, void* stream) {
  LaunchScope scope(KC_ROW, stream, 12.0 * n);
  LX_CHECK_ARG(dy && pre && dpre && n > 0, "lx_silu_bwd_f32: bad arguments");
  silu_bwd_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(dy, pre, dpre, n);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
