// CS3 (cross-scale state-space signal encoders) and DGF (DUAN dynamic gated fusion) kernels, all fp32 (the reference
// only runs this path in float32: train/config/seed_512.yaml:2).  Once-per-edit, HBM-bound work:
// coalesced loads along the signal / feature axis, warp-shuffle reductions, no tensor cores (fp32 accuracy contract
// of 1e-4 rules out TF32/bf16 MMA here).
//
//   pad_truncate        OminiModel.spatial_pyramid_pooling                                  model.py:479-511
//   s4_kernel_gen       S4 (DPLR, HiPPO-LegS) convolution kernel: Cauchy sums at the roots of unity + inverse DFT,
//                       float64 internally; cached at load because the weights are frozen at inference
//                       [s4torch S4Layer, SURVEY.md App. B]
//   s4_conv_gelu        y = GELU(causal_conv(u, K) + D u) per channel (the LTI SSM in its convolution form; S4 has no
//                       input-dependent "selective" scan, and its DPLR state matrix is dense, so the recurrent/scan form
//                       would carry an n x n complex state per channel)
//   channel_linear      per-position Linear over channels (+ residual + LayerNorm)         [s4torch S4Block / S4Model]
//   adaptive_pool       nn.AdaptiveAvgPool1d / FeaturePyramidPooling                       model.py:83-103, 345-373
//   gemv / ln_relu      projection MLP Linear -> LayerNorm -> ReLU                         model.py:60-72
//   token_linear        Unflatten(512, 8) -> Linear(8, 4096)                               model.py:70-71
//   sgemm               batched C = act(A . B + bias) (DUAN 1x1 convs, fusion1..4 over the token axis)
//   duan_*              statistics, importance, top-k channel mask, apply                  model.py:989-1035
#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

// ------------------------------------------------------------------------------------------------ pad / truncate
__global__ void pad_truncate_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int Lin, int Lout) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)rows * Lout) return;
  const int r = (int)(idx / Lout), l = (int)(idx % Lout);
  out[idx] = l < Lin ? in[(int64_t)r * Lin + l] : 0.f;
}

// ------------------------------------------------------------------------------------------------ S4 kernel generation
struct cd {
  double x, y;
};
__device__ __forceinline__ cd cmul(cd a, cd b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ cd cadd(cd a, cd b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cd csub(cd a, cd b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cd cconj(cd a) { return {a.x, -a.y}; }
__device__ __forceinline__ cd cdiv(cd a, cd b) {
  const double den = b.x * b.x + b.y * b.y;
  return {(a.x * b.x + a.y * b.y) / den, (a.y * b.x - a.x * b.y) / den};
}

// at_roots[c, l] = c(w) * (k00 - k01 * k10 / (1 + k11)),  w = exp(-2 pi i l / L)
__global__ void s4_cauchy_kernel(const float2* __restrict__ lam, const float2* __restrict__ p, const float2* __restrict__ q,
                                 const float2* __restrict__ Bm, const float2* __restrict__ Ct,
                                 const float* __restrict__ log_step, double2* __restrict__ at_roots, int d, int n, int L) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (l >= L) return;
  const double step = exp((double)log_step[c]);
  if (2 * l == L) {
    // w = -1: g -> infinity; the limit of the generating function is step/2 * sum_n conj(Ct_n) B_n
    cd s = {0, 0};
    for (int j = 0; j < n; ++j) {
      cd ct = {Ct[c * n + j].x, Ct[c * n + j].y}, b = {Bm[c * n + j].x, Bm[c * n + j].y};
      s = cadd(s, cmul(cconj(ct), b));
    }
    at_roots[(size_t)c * L + l] = make_double2(0.5 * step * s.x, 0.5 * step * s.y);
    return;
  }
  double sn, cs;
  sincospi(-2.0 * (double)l / (double)L, &sn, &cs);
  const cd w = {cs, sn};
  const cd one = {1.0, 0.0};
  const cd opw = cadd(one, w);
  cd g = cdiv(csub(one, w), opw);
  g.x *= 2.0 / step;
  g.y *= 2.0 / step;
  const cd cc = cdiv({2.0, 0.0}, opw);
  cd k00 = {0, 0}, k01 = {0, 0}, k10 = {0, 0}, k11 = {0, 0};
  for (int j = 0; j < n; ++j) {
    const cd lj = {lam[j].x, lam[j].y}, pj = {p[j].x, p[j].y}, qj = {q[j].x, q[j].y};
    const cd ct = {Ct[c * n + j].x, Ct[c * n + j].y}, b = {Bm[c * n + j].x, Bm[c * n + j].y};
    const cd inv = cdiv(one, csub(g, lj));
    const cd a0 = cconj(ct), a1 = cconj(qj);
    k00 = cadd(k00, cmul(cmul(a0, b), inv));
    k01 = cadd(k01, cmul(cmul(a0, pj), inv));
    k10 = cadd(k10, cmul(cmul(a1, b), inv));
    k11 = cadd(k11, cmul(cmul(a1, pj), inv));
  }
  const cd corr = cmul(cmul(k01, cdiv(one, cadd(one, k11))), k10);
  const cd r = cmul(cc, csub(k00, corr));
  at_roots[(size_t)c * L + l] = make_double2(r.x, r.y);
}

// K[c, k] = Re( (1/L) sum_l at_roots[c, l] exp(+2 pi i k l / L) )
__global__ void s4_idft_kernel(const double2* __restrict__ at_roots, float* __restrict__ K, int L) {
  extern __shared__ double2 tw[];  // exp(2 pi i m / L)
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    double sn, cs;
    sincospi(2.0 * (double)m / (double)L, &sn, &cs);
    tw[m] = make_double2(cs, sn);
  }
  __syncthreads();
  const int c = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= L) return;
  const double2* a = at_roots + (size_t)c * L;
  double acc = 0.0;
  int idx = 0;
  for (int l = 0; l < L; ++l) {
    const double2 v = a[l], t = tw[idx];
    acc += v.x * t.x - v.y * t.y;
    idx += k;
    if (idx >= L) idx -= L;
  }
  K[(size_t)c * L + k] = (float)(acc / L);
}

// The same transform as a radix-2 FFT when L is a power of two (every signal length of the model is: 4096 / 512 / 256 /
// 128): one CTA per channel, the L complex128 values stay in shared memory through the log2(L) butterfly stages.
// kInverse: exp(+2 pi i jk / L) (kernel generation: K = Re(IDFT(at_roots)));  else exp(-2 pi i jk / L) (its adjoint:
// G_at = DFT(dK) / L).  Input real (in_r) or complex (in_c); output real part (out_r) or complex (out_c), scaled by 1 / L.
template <bool kInverse>
__global__ void __launch_bounds__(512) s4_fft_kernel(const double2* __restrict__ in_c, const float* __restrict__ in_r,
                                                     double2* __restrict__ out_c, float* __restrict__ out_r, int L, int logL) {
  extern __shared__ double2 sf[];
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const int j = (int)(__brev((unsigned)i) >> (32 - logL));
    sf[j] = in_c != nullptr ? in_c[(size_t)c * L + i] : make_double2((double)in_r[(size_t)c * L + i], 0.0);
  }
  __syncthreads();
  for (int len = 2; len <= L; len <<= 1) {
    const int half = len >> 1;
    for (int t = threadIdx.x; t < (L >> 1); t += blockDim.x) {
      const int grp = t / half, pos = t - grp * half;
      const int i = grp * len + pos, j = i + half;
      double sn, cs;
      sincospi((kInverse ? 2.0 : -2.0) * (double)pos / (double)len, &sn, &cs);
      const double2 u = sf[i], v = sf[j];
      const double2 vw = make_double2(v.x * cs - v.y * sn, v.x * sn + v.y * cs);
      sf[i] = make_double2(u.x + vw.x, u.y + vw.y);
      sf[j] = make_double2(u.x - vw.x, u.y - vw.y);
    }
    __syncthreads();
  }
  const double inv = 1.0 / (double)L;
  for (int k = threadIdx.x; k < L; k += blockDim.x) {
    if (out_r != nullptr) out_r[(size_t)c * L + k] = (float)(sf[k].x * inv);
    else out_c[(size_t)c * L + k] = make_double2(sf[k].x * inv, sf[k].y * inv);
  }
}

// ------------------------------------------------------------------------------------------------ S4 convolution
// y[b,c,l] = act( sum_{j<=l} K[c,j] u[b,c,l-j] + D[c] u[b,c,l] ), act = GELU(erf) or identity, optional pre-activation
// output, optional time reversal of u and y (the adjoint of a causal convolution is the causal convolution of the reversed
// sequence: used by the backward).  One CTA = 256 consecutive outputs of one (b, c), four warps: every lane owns EIGHT
// consecutive outputs and slides an 8-value register window over the input, so one shared-memory read of u and one
// (broadcast) read of K feed eight FMAs; the four warps split the 512 taps of a staging round and their partial sums are
// added in a fixed order at the end (bit-reproducible).  The u window is stored with a one-word skew per 32 words, which
// makes the stride-8 window reads conflict-free.  (History, L = 4096: one output per thread, one FMA per two shared reads:
// 84 us for 64 channels; 1024 outputs per CTA with the register window: 94 us, bound by the single longest CTA - also
// 66 us for FOUR channels; this form spreads the longest tap range over 4x more warps.)
constexpr int CONV_T = 128, CONV_R = 8, CONV_OUT = 32 * CONV_R, CONV_J = 512, CONV_JW = CONV_J / 4;
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ int conv_skew(int i) { return i + (i >> 5); }

template <bool kReverse, bool kGelu>
__global__ void __launch_bounds__(CONV_T) s4_conv_kernel(const float* __restrict__ u, const float* __restrict__ K,
                                                         const float* __restrict__ Dp, float* __restrict__ y,
                                                         float* __restrict__ pre, int d, int L) {
  constexpr int WIN = CONV_OUT + CONV_J;  // u[base .. base + WIN), base = l0 - j0 - (CONV_J - 1) (one slot more than needed)
  __shared__ float sK[CONV_J];
  __shared__ float sU[WIN + WIN / 32 + 1];
  __shared__ float sAcc[3][CONV_OUT];
  const int c = blockIdx.y, b = blockIdx.z;
  const int l0 = (gridDim.x - 1 - blockIdx.x) * CONV_OUT;  // the CTAs with the longest tap range are scheduled first
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const float* ub = u + ((size_t)b * d + c) * L;
  const float* Kc = K + (size_t)c * L;
  auto ld_u = [&](int i) -> float { return (i >= 0 && i < L) ? ub[kReverse ? L - 1 - i : i] : 0.f; };
  float acc[CONV_R];
#pragma unroll
  for (int r = 0; r < CONV_R; ++r) acc[r] = 0.f;
  const int l_hi = min(L, l0 + CONV_OUT) - 1;  // last output of this CTA
  for (int j0 = 0; j0 <= l_hi; j0 += CONV_J) {
    __syncthreads();
    for (int i = threadIdx.x; i < CONV_J; i += CONV_T) sK[i] = (j0 + i < L) ? Kc[j0 + i] : 0.f;
    const int base = l0 - j0 - (CONV_J - 1);
    for (int i = threadIdx.x; i < WIN; i += CONV_T) sU[conv_skew(i)] = ld_u(base + i);
    __syncthreads();
    // output l_r = l0 + 8 lane + r, tap j0 + jj: u[l_r - j0 - jj] = window[8 lane + r + (CONV_J - 1) - jj];
    // warp q takes the taps jj in [128 q, 128 q + 128)
    const int jb = q * CONV_JW;
    if (j0 + jb > l_hi) continue;  // (warp-uniform) only zero padding beyond the last output's index
    float w[CONV_R];
#pragma unroll
    for (int r = 0; r < CONV_R; ++r) w[r] = sU[conv_skew(CONV_R * lane + r + CONV_J - 1 - jb)];
#pragma unroll 8
    for (int jj = 0; jj < CONV_JW; ++jj) {
      const float kv = sK[jb + jj];
#pragma unroll
      for (int r = 0; r < CONV_R; ++r) acc[r] = fmaf(kv, w[r], acc[r]);
#pragma unroll
      for (int r = CONV_R - 1; r > 0; --r) w[r] = w[r - 1];
      w[0] = sU[conv_skew(max(CONV_R * lane + CONV_J - 2 - jb - jj, 0))];  // (the value loaded on the very last tap is unused)
    }
  }
  __syncthreads();
  if (q > 0) {
#pragma unroll
    for (int r = 0; r < CONV_R; ++r) sAcc[q - 1][CONV_R * lane + r] = acc[r];
  }
  __syncthreads();
  if (q == 0) {
#pragma unroll
    for (int r = 0; r < CONV_R; ++r) {
      const int l = l0 + CONV_R * lane + r;
      if (l < L) {
        const float tot = ((acc[r] + sAcc[0][CONV_R * lane + r]) + sAcc[1][CONV_R * lane + r]) + sAcc[2][CONV_R * lane + r];
        const float sv = tot + (Dp ? Dp[c] * ld_u(l) : 0.f);
        const size_t o = ((size_t)b * d + c) * L + (kReverse ? L - 1 - l : l);
        if (pre != nullptr) pre[o] = sv;
        y[o] = kGelu ? gelu_erf(sv) : sv;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ channel linear
// out[b, j, l] = sum_c W[j, c] in[b, c, l] + bias[j]; optional residual add + LayerNorm over the d_out channels.
// One CTA = 32 positions x 8 output groups (one warp per group: the weight reads are shared-memory broadcasts, the
// activation reads 128-byte rows); 128 CTAs per sample at L = 4096 (the first version ran 32 CTAs of 128 threads with all
// d_out accumulators in one thread: 60 us per launch for 17 MFLOP).
constexpr int CL_MAX = 64, CL_GROUPS = 8, CL_JPT = CL_MAX / CL_GROUPS;
__global__ void __launch_bounds__(256) channel_linear_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                             const float* __restrict__ bias, const float* __restrict__ residual,
                                                             const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                             float* __restrict__ out, int d_in, int d_out, int L, float eps) {
  extern __shared__ float sW[];  // [d_out * d_in] + [d_out] bias
  __shared__ float red[CL_GROUPS][32];
  for (int i = threadIdx.x; i < d_out * d_in; i += blockDim.x) sW[i] = W[i];
  for (int i = threadIdx.x; i < d_out; i += blockDim.x) sW[d_out * d_in + i] = bias[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int lp = threadIdx.x & 31, jg = threadIdx.x >> 5;
  const int l = blockIdx.x * 32 + lp;
  const int jpt = (d_out + CL_GROUPS - 1) / CL_GROUPS;  // outputs per thread (<= CL_JPT)
  const int jb = jg * jpt;
  const bool lok = l < L;
  float acc[CL_JPT];
#pragma unroll
  for (int i = 0; i < CL_JPT; ++i) acc[i] = (i < jpt && jb + i < d_out) ? sW[d_out * d_in + jb + i] : 0.f;
  if (lok) {
    for (int c = 0; c < d_in; ++c) {
      const float x = in[((size_t)b * d_in + c) * L + l];
#pragma unroll
      for (int i = 0; i < CL_JPT; ++i)
        if (i < jpt && jb + i < d_out) acc[i] = fmaf(sW[(jb + i) * d_in + c], x, acc[i]);
    }
    if (residual != nullptr) {
#pragma unroll
      for (int i = 0; i < CL_JPT; ++i)
        if (i < jpt && jb + i < d_out) acc[i] += residual[((size_t)b * d_out + jb + i) * L + l];
    }
  }
  if (ln_w != nullptr) {  // LayerNorm over the channels of one position = over the 8 groups of this lane
    float ps = 0.f;
#pragma unroll
    for (int i = 0; i < CL_JPT; ++i)
      if (i < jpt && jb + i < d_out) ps += acc[i];
    red[jg][lp] = ps;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < CL_GROUPS; ++g2) mean += red[g2][lp];
    mean /= d_out;
    __syncthreads();
    float pv = 0.f;
#pragma unroll
    for (int i = 0; i < CL_JPT; ++i)
      if (i < jpt && jb + i < d_out) pv += (acc[i] - mean) * (acc[i] - mean);
    red[jg][lp] = pv;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int g2 = 0; g2 < CL_GROUPS; ++g2) var += red[g2][lp];
    const float rstd = rsqrtf(var / d_out + eps);
#pragma unroll
    for (int i = 0; i < CL_JPT; ++i)
      if (i < jpt && jb + i < d_out) acc[i] = (acc[i] - mean) * rstd * ln_w[jb + i] + ln_b[jb + i];
  }
  if (lok) {
#pragma unroll
    for (int i = 0; i < CL_JPT; ++i)
      if (i < jpt && jb + i < d_out) out[((size_t)b * d_out + jb + i) * L + l] = acc[i];
  }
}

// ------------------------------------------------------------------------------------------------ adaptive avg pool
// in [B, C, L] -> out[b * out_bstride + c * cs + i * is + off] = mean(in[b, c, floor(i L / O) : ceil((i+1) L / O)])
__global__ void adaptive_pool_kernel(const float* __restrict__ in, float* __restrict__ out, int Cc, int L, int O,
                                     int64_t out_bstride, int cs, int is, int off) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (i >= O) return;
  const int s = (int)(((int64_t)i * L) / O);
  const int e = (int)((((int64_t)(i + 1)) * L + O - 1) / O);
  const float* x = in + ((size_t)b * Cc + c) * L;
  float acc = 0.f;
  for (int l = s; l < e; ++l) acc += x[l];
  out[(size_t)b * out_bstride + (size_t)c * cs + (size_t)i * is + off] = acc / (float)(e - s);
}
// long bins (O << L, e.g. the 4 bins of 1024 samples over the S4 output, model.py:83-87): one warp per bin, coalesced
__global__ void __launch_bounds__(128) adaptive_pool_warp_kernel(const float* __restrict__ in, float* __restrict__ out, int Cc,
                                                                 int L, int O, int64_t out_bstride, int cs, int is, int off) {
  const int i = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int c = blockIdx.y, b = blockIdx.z;
  if (i >= O) return;
  const int s = (int)(((int64_t)i * L) / O);
  const int e = (int)((((int64_t)(i + 1)) * L + O - 1) / O);
  const float* x = in + ((size_t)b * Cc + c) * L;
  float acc = 0.f;
  for (int l = s + (threadIdx.x & 31); l < e; l += 32) acc += x[l];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) out[(size_t)b * out_bstride + (size_t)c * cs + (size_t)i * is + off] = acc / (float)(e - s);
}

// Several pooling sizes of the same input in one launch (FeaturePyramidPooling: concat over `n` output sizes, model.py:
// 345-373): bin i of pool k goes to out[b * out_bstride + c * cs + off_k + i].
struct PoolList {
  int n;
  int O[8], off[8], start[9];  // start = prefix sums of O (thread index -> pool)
};
__global__ void adaptive_pool_multi_kernel(const float* __restrict__ in, float* __restrict__ out, int Cc, int L,
                                           int64_t out_bstride, int cs, const PoolList pl) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (idx >= pl.start[pl.n]) return;
  int k = 0;
  while (k + 1 < pl.n && idx >= pl.start[k + 1]) ++k;
  const int O = pl.O[k], i = idx - pl.start[k];
  const int s = (int)(((int64_t)i * L) / O);
  const int e = (int)((((int64_t)(i + 1)) * L + O - 1) / O);
  const float* x = in + ((size_t)b * Cc + c) * L;
  float acc = 0.f;
  for (int l = s; l < e; ++l) acc += x[l];
  out[(size_t)b * out_bstride + (size_t)c * cs + pl.off[k] + i] = acc / (float)(e - s);
}

// ------------------------------------------------------------------------------------------------ GEMV (B <= 8)
// y[b, j] = sum_i W[j, i] x[b, i] + bias[j]; one warp per output row j, W streamed once with float4 loads.
constexpr int GEMV_MAXB = 8;
__global__ void __launch_bounds__(256) gemv_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                   const float* __restrict__ x, float* __restrict__ y, int B, int n_out,
                                                   int n_in, int64_t ldx, int64_t ldy) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  if (j >= n_out) return;
  const float* w = W + (size_t)j * n_in;
  float acc[GEMV_MAXB];
#pragma unroll
  for (int b = 0; b < GEMV_MAXB; ++b) acc[b] = 0.f;
  const int n4 = n_in >> 2;
  int i = lane;
  for (; i + 96 < n4; i += 128) {  // four independent 16-byte weight loads in flight per lane (the row is streamed once)
    float4 wv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) wv[q] = __ldg(reinterpret_cast<const float4*>(w) + i + 32 * q);
#pragma unroll
    for (int b = 0; b < GEMV_MAXB; ++b) {
      if (b < B) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 xv = *(reinterpret_cast<const float4*>(x + (size_t)b * ldx) + i + 32 * q);
          acc[b] += wv[q].x * xv.x + wv[q].y * xv.y + wv[q].z * xv.z + wv[q].w * xv.w;
        }
      }
    }
  }
  for (; i < n4; i += 32) {
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w) + i);
#pragma unroll
    for (int b = 0; b < GEMV_MAXB; ++b) {
      if (b < B) {
        const float4 xv = *(reinterpret_cast<const float4*>(x + (size_t)b * ldx) + i);
        acc[b] += wv.x * xv.x + wv.y * xv.y + wv.z * xv.z + wv.w * xv.w;
      }
    }
  }
  for (int i = (n4 << 2) + lane; i < n_in; i += 32) {
    const float wv = w[i];
#pragma unroll
    for (int b = 0; b < GEMV_MAXB; ++b)
      if (b < B) acc[b] += wv * x[(size_t)b * ldx + i];
  }
#pragma unroll
  for (int b = 0; b < GEMV_MAXB; ++b) {
    if (b < B) {
      const float s = warp_sum(acc[b]);
      if (lane == 0) y[(size_t)b * ldy + j] = s + (bias ? bias[j] : 0.f);
    }
  }
}

// rows [R, n]: y = relu(LayerNorm(x) * w + b), one CTA per row
__global__ void __launch_bounds__(256) ln_relu_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bvec, float* __restrict__ y, int n,
                                                      float eps) {
  __shared__ float red[8];
  __shared__ float stat;
  const float* xr = x + (size_t)blockIdx.x * n;
  float* yr = y + (size_t)blockIdx.x * n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += xr[i];
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat = t / n;
  }
  __syncthreads();
  const float mean = stat;
  float v = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float c = xr[i] - mean;
    v += c * c;
  }
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    stat = rsqrtf(t / n + eps);
  }
  __syncthreads();
  const float rstd = stat;
  for (int i = threadIdx.x; i < n; i += 256) yr[i] = fmaxf((xr[i] - mean) * rstd * w[i] + bvec[i], 0.f);
}

// out[b, t, o] = sum_{i<8} W[o, i] h[b, t*8 + i] + bias[o]   (Unflatten(512,8) -> Linear(8, n_out))
__global__ void token_linear_kernel(const float* __restrict__ h, const float* __restrict__ W,
                                    const float* __restrict__ bias, float* __restrict__ out, int tokens, int n_out,
                                    int64_t out_bstride) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = blockIdx.y, b = blockIdx.z;
  if (o >= n_out) return;
  const float* hv = h + ((size_t)b * tokens + t) * 8;
  const float4 w0 = *reinterpret_cast<const float4*>(W + (size_t)o * 8);
  const float4 w1 = *reinterpret_cast<const float4*>(W + (size_t)o * 8 + 4);
  const float acc = w0.x * hv[0] + w0.y * hv[1] + w0.z * hv[2] + w0.w * hv[3] + w1.x * hv[4] + w1.y * hv[5] +
                    w1.z * hv[6] + w1.w * hv[7] + bias[o];
  out[(size_t)b * out_bstride + (size_t)t * n_out + o] = acc;
}

// ------------------------------------------------------------------------------------------------ batched SGEMM
// C[b] = act(A[M,K] . Bm[b][K,N] + bias[m] [+ R[b]]);  A row-major (lda), Bm row-major (ldb), 64x64x16 tiles,
// 256 threads, 4x4 outputs per thread.  act: 0 none, 1 relu, 2 sigmoid.  If rowmean != NULL, instead of storing C the
// kernel accumulates mean_n(act(.)) into rowmean[b, m] (atomicAdd of per-tile partial sums / N).
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, int64_t lda, const float* __restrict__ Bm,
                                                    int64_t ldb, int64_t b_bstride, const float* __restrict__ bias,
                                                    const float* __restrict__ R, float* __restrict__ Cm, int64_t ldc,
                                                    int64_t c_bstride, float* __restrict__ rowmean,
                                                    float* __restrict__ rowpart, int M, int N, int K, int act) {
  __shared__ float sA[16][64 + 4];
  __shared__ float sB[16][64 + 4];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const float* Bb = Bm + (size_t)b * b_bstride;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 x 16 threads
  float acc[4][4] = {};
  // The next k-tile's global loads are issued before this tile's FMAs (registers as the second buffer): the loads of a
  // tile used to be exposed between two __syncthreads.  Accumulation order per output element is unchanged (k ascending).
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256;
      const int m = i >> 4, ka = i & 15;  // A tile: 64 rows x 16 k  -> sA[k][m]
      ra[t] = (m0 + m < M && k0 + ka < K) ? A[(size_t)(m0 + m) * lda + k0 + ka] : 0.f;
      const int kb = i >> 6, n = i & 63;  // B tile: 16 k x 64 n
      rb[t] = (k0 + kb < K && n0 + n < N) ? Bb[(size_t)(k0 + kb) * ldb + n0 + n] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int i = threadIdx.x + t * 256;
      sA[i & 15][i >> 4] = ra[t];
      sB[i >> 6][i & 63] = rb[t];
    }
    __syncthreads();
    if (k0 + 16 < K) fetch(k0 + 16);
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    float rsum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (m < M && n < N) {
        float v = acc[i][j] + (bias ? bias[m] : 0.f);
        if (R != nullptr) v += R[(size_t)b * c_bstride + (size_t)m * ldc + n];
        if (act == 1) v = fmaxf(v, 0.f);
        else if (act == 2) v = 1.0f / (1.0f + expf(-v));
        if (rowmean != nullptr) rsum += v;
        else Cm[(size_t)b * c_bstride + (size_t)m * ldc + n] = v;
      }
    }
    if (rowmean != nullptr) {
      // reduce over the 16 tx lanes that share this row (lanes of a half-warp)
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) rsum += __shfl_xor_sync(0xffffffffu, rsum, o);
      if (tx == 0 && m < M) {
        // deterministic: per-tile partial sums, added up in tile order by rowmean_finish_kernel (no float atomics)
        if (rowpart != nullptr) rowpart[((size_t)b * M + m) * gridDim.x + blockIdx.x] = rsum;
        else atomicAdd(&rowmean[(size_t)b * M + m], rsum / (float)N);
      }
    }
  }
}

__global__ void rowmean_finish_kernel(const float* __restrict__ rowpart, float* __restrict__ rowmean, int rows, int tiles,
                                      int N) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s2 = 0.f;
  for (int t = 0; t < tiles; ++t) s2 += rowpart[(size_t)r * tiles + t];
  rowmean[r] = s2 / (float)N;
}

// ------------------------------------------------------------------------------------------------ DUAN
// per (b, ch) row of x [B, C, L]: mean and M2 = sum (x - mean)^2 (two-pass, like torch.var), plus mean of c.
__global__ void __launch_bounds__(256) duan_row_stats_kernel(const float* __restrict__ x, const float* __restrict__ c,
                                                             float* __restrict__ mean_x, float* __restrict__ m2_x,
                                                             float* __restrict__ mean_c, int L) {
  __shared__ float red[8];
  __shared__ float bc;
  const size_t row = blockIdx.x;
  const float* xr = x + row * L;
  const float* cr = c + row * L;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int i = 0; i < 8; ++i) t += red[i];
      bc = t;
    }
    __syncthreads();
    return bc;
  };
  float s = 0.f, sc = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) {
    s += xr[i];
    sc += cr[i];
  }
  const float mean = block_sum(s) / L;
  const float cm = block_sum(sc) / L;
  float v = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) {
    const float dlt = xr[i] - mean;
    v += dlt * dlt;
  }
  const float m2 = block_sum(v);
  if (threadIdx.x == 0) {
    mean_x[row] = mean;
    m2_x[row] = m2;
    mean_c[row] = cm;
  }
}

// Combines row statistics into the mixed (mu, 1/sigma) and the FiLM (1+gamma, beta) per (b, ch); one CTA per batch.
// layer statistics from row (mean, M2) pairs with Chan's parallel-variance formula.
__global__ void duan_mix_kernel(const float* __restrict__ mean_x, const float* __restrict__ m2_x,
                                const float* __restrict__ g_mix, const float* __restrict__ gamma_beta,
                                float* __restrict__ mu_out, float* __restrict__ rsig_out, float* __restrict__ g1_out,
                                float* __restrict__ beta_out, int Cc, int L, float eps) {
  __shared__ double sh[2];
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    double tot = 0.0;
    for (int ch = 0; ch < Cc; ++ch) tot += (double)mean_x[b * Cc + ch];
    const double mu_l = tot / Cc;  // equal row lengths
    double m2 = 0.0;
    for (int ch = 0; ch < Cc; ++ch) {
      const double dm = (double)mean_x[b * Cc + ch] - mu_l;
      m2 += (double)m2_x[b * Cc + ch] + dm * dm * L;
    }
    sh[0] = mu_l;
    sh[1] = sqrt(m2 / ((double)Cc * L) + (double)eps);
  }
  __syncthreads();
  const float mu_l = (float)sh[0], sigma_l = (float)sh[1];
  for (int ch = threadIdx.x; ch < Cc; ch += blockDim.x) {
    const int i = b * Cc + ch;
    const float g = g_mix[i];
    const float sigma_c = sqrtf(m2_x[i] / L + eps);
    const float mu = g * mean_x[i] + (1.f - g) * mu_l;
    const float sigma = g * sigma_c + (1.f - g) * sigma_l;
    mu_out[i] = mu;
    rsig_out[i] = 1.0f / sigma;
    g1_out[i] = 1.0f + gamma_beta[(size_t)b * 2 * Cc + ch];
    beta_out[i] = gamma_beta[(size_t)b * 2 * Cc + Cc + ch];
  }
}

// imp[b, ch] = mean_l | (1+gamma) (x - mu) / sigma + beta |
__global__ void __launch_bounds__(256) duan_importance_kernel(const float* __restrict__ x, const float* __restrict__ mu,
                                                              const float* __restrict__ rsig, const float* __restrict__ g1,
                                                              const float* __restrict__ beta, float* __restrict__ imp,
                                                              int L) {
  __shared__ float red[8];
  const size_t row = blockIdx.x;
  const float* xr = x + row * L;
  const float m = mu[row], rs = rsig[row], g = g1[row], bt = beta[row];
  float s = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) s += fabsf(g * ((xr[i] - m) * rs) + bt);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    imp[row] = t / L;
  }
}

// mask[b, ch] = 1 if ch is among the k channels with the largest importance (ties -> lower index first)
__global__ void duan_topk_mask_kernel(const float* __restrict__ imp, float* __restrict__ mask, int Cc, int k) {
  extern __shared__ float simp[];
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < Cc; i += blockDim.x) simp[i] = imp[b * Cc + i];
  __syncthreads();
  for (int i = threadIdx.x; i < Cc; i += blockDim.x) {
    const float v = simp[i];
    int rank = 0;
    for (int j = 0; j < Cc; ++j) {
      const float o = simp[j];
      rank += (o > v) || (o == v && j < i);
    }
    mask[b * Cc + i] = rank < k ? 1.f : 0.f;
  }
}

// y[b, ch, :] = mask * ((1+gamma) (x - mu) / sigma + beta), written with an arbitrary batch stride (e.g. into the
// second half of the fusion concat buffer)
__global__ void duan_apply_kernel(const float* __restrict__ x, const float* __restrict__ mu, const float* __restrict__ rsig,
                                  const float* __restrict__ g1, const float* __restrict__ beta,
                                  const float* __restrict__ mask, float* __restrict__ y, int Cc, int L,
                                  int64_t y_bstride) {
  const int row = blockIdx.y;  // b * Cc + ch
  const int b = row / Cc, ch = row % Cc;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  const float v = mask[row] * (g1[row] * ((x[(size_t)row * L + i] - mu[row]) * rsig[row]) + beta[row]);
  y[(size_t)b * y_bstride + (size_t)ch * L + i] = v;
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 v = *reinterpret_cast<const float2*>(in + i);
    *reinterpret_cast<uint32_t*>(out + i) = pack_bf16(v.x, v.y);
  } else if (i < n) {
    out[i] = __float2bfloat16_rn(in[i]);
  }
}
__global__ void cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

// shared by the kernel generation (inverse, real output) and its backward (forward transform of the real dK, complex output)
int s4_fft_launch(bool inverse, const double2* in_c, const float* in_r, double2* out_c, float* out_r, int d, int L,
                  void* stream) {
  int logL = 0;
  while ((1 << logL) < L) ++logL;
  const size_t smem = (size_t)L * sizeof(double2);
  static int configured[2] = {0, 0};
  if (smem > 48 * 1024 && (int)smem > configured[inverse ? 1 : 0]) {
    if (inverse) LX_CUDA(cudaFuncSetAttribute(s4_fft_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else LX_CUDA(cudaFuncSetAttribute(s4_fft_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[inverse ? 1 : 0] = (int)smem;
  }
  const int threads = L / 2 >= 512 ? 512 : (L / 2 >= 32 ? L / 2 : 32);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (inverse) s4_fft_kernel<true><<<d, threads, smem, st>>>(in_c, in_r, out_c, out_r, L, logL);
  else s4_fft_kernel<false><<<d, threads, smem, st>>>(in_c, in_r, out_c, out_r, L, logL);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

}  // namespace lx

using namespace lx;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" int lx_pad_truncate(const float* in, float* out, int32_t rows, int32_t Lin, int32_t Lout, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * rows * ((Lin < Lout ? Lin : Lout) + (double)Lout));  // algorithmic bytes
  LX_CHECK_ARG(in && out && rows > 0 && Lin > 0 && Lout > 0, "lx_pad_truncate: bad arguments");
  const int64_t n = (int64_t)rows * Lout;
  pad_truncate_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>(in, out, rows, Lin, Lout);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_s4_kernel_gen(const void* lam, const void* p, const void* q, const void* Bm, const void* Ct,
                                const float* log_step, float* K, void* workspace, int32_t d, int32_t n, int32_t L,
                                void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 16.0 * d * (double)L);  // algorithmic bytes
  LX_CHECK_ARG(lam && p && q && Bm && Ct && log_step && K && workspace, "lx_s4_kernel_gen: null pointer");
  LX_CHECK_ARG(d > 0 && n > 0 && L > 0 && L * 16 <= 200 * 1024, "lx_s4_kernel_gen: L=%d too large for the twiddle table", L);
  dim3 g1((L + 127) / 128, d);
  s4_cauchy_kernel<<<g1, 128, 0, ST(stream)>>>((const float2*)lam, (const float2*)p, (const float2*)q, (const float2*)Bm,
                                               (const float2*)Ct, log_step, (double2*)workspace, d, n, L);
  LX_CUDA(cudaGetLastError());
  const size_t smem = (size_t)L * sizeof(double2);
  if ((L & (L - 1)) == 0 && L >= 2) {  // power of two: O(L log L) FFT, one CTA per channel
    int rc = lx::s4_fft_launch(true, (const double2*)workspace, nullptr, nullptr, K, d, L, stream);
    return rc;
  }
  static int configured = 0;
  if (smem > 48 * 1024 && (int)smem > configured) {
    LX_CUDA(cudaFuncSetAttribute(s4_idft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = (int)smem;
  }
  s4_idft_kernel<<<g1, 128, smem, ST(stream)>>>((const double2*)workspace, K, L);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_s4_conv_gelu(const float* u, const float* K, const float* D, float* y, int32_t B, int32_t d, int32_t L,
                               void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 8.0 * B * d * (double)L + 4.0 * d * (double)L);  // algorithmic bytes
  LX_CHECK_ARG(u && K && D && y && B > 0 && d > 0 && L > 0, "lx_s4_conv_gelu: bad arguments");
  return lx_s4_conv(u, K, D, y, nullptr, B, d, L, 0, 1, stream);
}

extern "C" int lx_s4_conv(const float* u, const float* K, const float* D, float* y, float* pre, int32_t B, int32_t d,
                          int32_t L, int32_t reverse, int32_t gelu, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 8.0 * B * d * (double)L + 4.0 * d * (double)L);
  LX_CHECK_ARG(u && K && y && B > 0 && d > 0 && L > 0, "lx_s4_conv: bad arguments");
  dim3 grid((L + CONV_OUT - 1) / CONV_OUT, d, B);
  if (reverse && gelu) s4_conv_kernel<true, true><<<grid, CONV_T, 0, ST(stream)>>>(u, K, D, y, pre, d, L);
  else if (reverse) s4_conv_kernel<true, false><<<grid, CONV_T, 0, ST(stream)>>>(u, K, D, y, pre, d, L);
  else if (gelu) s4_conv_kernel<false, true><<<grid, CONV_T, 0, ST(stream)>>>(u, K, D, y, pre, d, L);
  else s4_conv_kernel<false, false><<<grid, CONV_T, 0, ST(stream)>>>(u, K, D, y, pre, d, L);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_adaptive_pool_multi(const float* in, float* out, int32_t B, int32_t C, int32_t L, int32_t n,
                                      const int32_t* O, const int32_t* off, int64_t out_bstride, int32_t cs, void* stream) {
  LX_CHECK_ARG(in && out && O && off && B > 0 && C > 0 && L > 0 && n > 0 && n <= 8, "lx_adaptive_pool_multi: bad arguments");
  PoolList pl;
  pl.n = n;
  pl.start[0] = 0;
  for (int k = 0; k < n; ++k) {
    LX_CHECK_ARG(O[k] > 0, "lx_adaptive_pool_multi: bad output size");
    pl.O[k] = O[k];
    pl.off[k] = off[k];
    pl.start[k + 1] = pl.start[k] + O[k];
  }
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * B * C * ((double)L * n + pl.start[n]));
  dim3 grid((pl.start[n] + 127) / 128, C, B);
  adaptive_pool_multi_kernel<<<grid, 128, 0, ST(stream)>>>(in, out, C, L, out_bstride, cs, pl);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_channel_linear(const float* in, const float* W, const float* bias, const float* residual,
                                 const float* ln_w, const float* ln_b, float* out, int32_t B, int32_t d_in, int32_t d_out,
                                 int32_t L, float eps, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * B * (double)L * (d_in + d_out) + 4.0 * d_in * (double)d_out);  // algorithmic bytes
  LX_CHECK_ARG(in && W && bias && out && B > 0 && L > 0, "lx_channel_linear: bad arguments");
  LX_CHECK_ARG(d_in > 0 && d_in <= CL_MAX && d_out > 0 && d_out <= CL_MAX, "lx_channel_linear: d_in/d_out must be <= %d",
               CL_MAX);
  LX_CHECK_ARG((ln_w == nullptr) == (ln_b == nullptr), "lx_channel_linear: LayerNorm needs both weight and bias");
  dim3 grid((L + 31) / 32, B);
  const size_t smem = (size_t)(d_out * d_in + d_out) * sizeof(float);
  channel_linear_kernel<<<grid, 256, smem, ST(stream)>>>(in, W, bias, residual, ln_w, ln_b, out, d_in, d_out, L, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_adaptive_pool(const float* in, float* out, int32_t B, int32_t C, int32_t L, int32_t O,
                                int64_t out_bstride, int32_t cs, int32_t is, int32_t off, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * B * C * ((double)L + O));  // algorithmic bytes
  LX_CHECK_ARG(in && out && B > 0 && C > 0 && L > 0 && O > 0, "lx_adaptive_pool: bad arguments");
  if (L >= 64 * O) {  // (sums in a different order than the thread-per-bin form: both are plain fp32 sums of the bin)
    adaptive_pool_warp_kernel<<<dim3((O + 3) / 4, C, B), 128, 0, ST(stream)>>>(in, out, C, L, O, out_bstride, cs, is, off);
  } else {
    dim3 grid((O + 127) / 128, C, B);
    adaptive_pool_kernel<<<grid, 128, 0, ST(stream)>>>(in, out, C, L, O, out_bstride, cs, is, off);
  }
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_gemv_f32(const float* W, const float* bias, const float* x, float* y, int32_t B, int32_t n_out,
                           int32_t n_in, int64_t ldx, int64_t ldy, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * n_out * (double)n_in + 4.0 * B * ((double)n_in + n_out));  // algorithmic bytes
  LX_CHECK_ARG(W && x && y && n_out > 0 && n_in > 0, "lx_gemv_f32: bad arguments");
  LX_CHECK_ARG(B > 0 && B <= GEMV_MAXB, "lx_gemv_f32: batch %d outside [1, %d]", B, GEMV_MAXB);
  LX_CHECK_ARG(n_in % 4 == 0 ? ldx % 4 == 0 : true, "lx_gemv_f32: ldx must be a multiple of 4");
  gemv_kernel<<<(n_out + 7) / 8, 256, 0, ST(stream)>>>(W, bias, x, y, B, n_out, n_in, ldx, ldy);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_ln_relu_rows(const float* x, const float* w, const float* b, float* y, int32_t rows, int32_t n, float eps,
                               void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 8.0 * rows * (double)n);  // algorithmic bytes
  LX_CHECK_ARG(x && w && b && y && rows > 0 && n > 0, "lx_ln_relu_rows: bad arguments");
  ln_relu_kernel<<<rows, 256, 0, ST(stream)>>>(x, w, b, y, n, eps);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_token_linear(const float* h, const float* W, const float* bias, float* out, int32_t B, int32_t tokens,
                               int32_t n_out, int64_t out_bstride, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 4.0 * B * tokens * (8.0 + n_out));  // algorithmic bytes
  LX_CHECK_ARG(h && W && bias && out && B > 0 && tokens > 0 && n_out > 0, "lx_token_linear: bad arguments");
  dim3 grid((n_out + 255) / 256, tokens, B);
  token_linear_kernel<<<grid, 256, 0, ST(stream)>>>(h, W, bias, out, tokens, n_out, out_bstride);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

namespace lx {

// 128 x 128 x 16 tiles, 256 threads, 8 x 8 outputs per thread as 2 x 2 blocks of 4 x 4 (rows ty*4.. and 64+ty*4.., columns
// tx*4.. and 64+tx*4..: every operand fragment is one conflict-free 16-byte shared-memory load, 16 FMAs per load), the next
// k-tile in registers while this one is multiplied.  Per output element the products are added in ascending k like the
// 64 x 64 kernels (same results bit for bit where no K split is used); the 64 x 64 kernels reach ~12-16 TFLOP/s on the
// M = 512, N = K = 1024..4096, batch 8 products of the DUANs / fusion layers, this one is used for those.
constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16;
__global__ void __launch_bounds__(256, 2) sgemm128_kernel(const Sgemm128 p) {
  __shared__ __align__(16) float sA[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float sB[SG_BK][SG_BN + 4];
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int b = p.reduce_batch ? blockIdx.z / p.ksplit : blockIdx.z;
  const int kper = p.reduce_batch ? ((p.K + p.ksplit * SG_BK - 1) / (p.ksplit * SG_BK)) * SG_BK : p.K;
  const int k_lo = p.reduce_batch ? (blockIdx.z % p.ksplit) * kper : 0, k_hi = min(p.K, k_lo + kper);
  const float* Ab = p.A + (size_t)b * p.a_bs;
  const float* Bb = p.B + (size_t)b * p.b_bs;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int i = threadIdx.x + t * 256;
      int m, k;
      if (p.trans_a) { k = i >> 7; m = i & 127; } else { m = i >> 4; k = i & 15; }
      const bool oka = m0 + m < p.M && k0 + k < k_hi;
      ra[t] = oka ? (p.trans_a ? Ab[(size_t)(k0 + k) * p.lda + m0 + m] : Ab[(size_t)(m0 + m) * p.lda + k0 + k]) : 0.f;
      int n;
      if (p.trans_b) { n = i >> 4; k = i & 15; } else { k = i >> 7; n = i & 127; }
      const bool okb = k0 + k < k_hi && n0 + n < p.N;
      rb[t] = okb ? (p.trans_b ? Bb[(size_t)(n0 + n) * p.ldb + k0 + k] : Bb[(size_t)(k0 + k) * p.ldb + n0 + n]) : 0.f;
    }
  };
  if (k_lo < k_hi) fetch(k_lo);
  for (int k0 = k_lo; k0 < k_hi; k0 += SG_BK) {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int i = threadIdx.x + t * 256;
      if (p.trans_a) sA[i >> 7][i & 127] = ra[t]; else sA[i & 15][i >> 4] = ra[t];
      if (p.trans_b) sB[i & 15][i >> 4] = rb[t]; else sB[i >> 7][i & 127] = rb[t];
    }
    __syncthreads();
    if (k0 + SG_BK < k_hi) fetch(k0 + SG_BK);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[k][ty * 4]), a1 = *reinterpret_cast<const float4*>(&sA[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[k][tx * 4]), b1 = *reinterpret_cast<const float4*>(&sB[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cb = p.C + (p.reduce_batch ? 0 : (size_t)blockIdx.z * p.c_bs);
  const float* Rb = p.R != nullptr ? p.R + (size_t)blockIdx.z * p.c_bs : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
    if (m >= p.M) continue;
    const float bias = p.bias != nullptr ? p.bias[m] : 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j >> 2) * 64 + tx * 4 + (j & 3);
      if (n >= p.N) continue;
      float* c = Cb + (size_t)m * p.ldc + n;
      if (p.reduce_batch) {
        atomicAdd(c, p.alpha * acc[i][j]);  // beta == 1 (checked by the caller): C accumulates
      } else {
        float v = p.alpha * acc[i][j] + bias;
        if (Rb != nullptr) v += Rb[(size_t)m * p.ldc + n];
        if (p.act == 1) v = fmaxf(v, 0.f);
        else if (p.act == 2) v = 1.0f / (1.0f + expf(-v));
        *c = v + (p.beta != 0.f ? p.beta * *c : 0.f);
      }
    }
  }
}

int sgemm128_launch(const Sgemm128& p, int batch, void* stream) {
  dim3 grid((p.N + SG_BN - 1) / SG_BN, (p.M + SG_BM - 1) / SG_BM, p.reduce_batch ? batch * p.ksplit : batch);
  sgemm128_kernel<<<grid, 256, 0, ST(stream)>>>(p);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

}  // namespace lx

extern "C" int lx_sgemm_f32(const lx_sgemm_desc_t* desc, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, desc ? 4.0 * ((double)desc->M * desc->K + desc->batch * ((double)desc->K * desc->N + (double)desc->M * desc->N)) : 0.0);
  LX_CHECK_ARG(desc != nullptr, "lx_sgemm_f32: null descriptor");
  const lx_sgemm_desc_t& d = *desc;
  LX_CHECK_ARG(d.A && d.Bm && (d.C || d.rowmean) && d.M > 0 && d.N > 0 && d.K > 0 && d.batch > 0,
               "lx_sgemm_f32: bad arguments");
  LX_CHECK_ARG(d.act >= 0 && d.act <= 2, "lx_sgemm_f32: bad activation");
  if (d.rowmean == nullptr && d.M >= 128 && d.N >= 128 && (long)((d.M + 127) / 128) * ((d.N + 127) / 128) * d.batch >= num_sms()) {
    lx::Sgemm128 p{};  // the larger products: 128 x 128 tiles (same results bit for bit)
    p.A = d.A; p.B = d.Bm; p.C = d.C; p.bias = d.bias; p.R = d.R;
    p.lda = d.lda; p.ldb = d.ldb; p.ldc = d.ldc; p.a_bs = 0; p.b_bs = d.b_bstride; p.c_bs = d.c_bstride;
    p.M = d.M; p.N = d.N; p.K = d.K; p.act = d.act; p.ksplit = 1; p.alpha = 1.0f; p.beta = 0.0f;
    return lx::sgemm128_launch(p, d.batch, stream);
  }
  dim3 grid((d.N + 63) / 64, (d.M + 63) / 64, d.batch);
  sgemm_kernel<<<grid, 256, 0, ST(stream)>>>(d.A, d.lda, d.Bm, d.ldb, d.b_bstride, d.bias, d.R, d.C, d.ldc, d.c_bstride,
                                             d.rowmean, d.rowmean ? d.rowpart : nullptr, d.M, d.N, d.K, d.act);
  LX_CUDA(cudaGetLastError());
  if (d.rowmean != nullptr && d.rowpart != nullptr) {
    const int rows = d.batch * d.M;
    rowmean_finish_kernel<<<(rows + 255) / 256, 256, 0, ST(stream)>>>(d.rowpart, d.rowmean, rows, (int)grid.x, d.N);
    LX_CUDA(cudaGetLastError());
  }
  return LX_OK;
}

extern "C" int lx_cast(const void* in, void* out, int64_t n, int32_t to_bf16, void* stream) {
  LaunchScope scope(KC_CS3DGF, stream, 6.0 * n);  // algorithmic bytes
  LX_CHECK_ARG(in && out && n > 0, "lx_cast: bad arguments");
  if (to_bf16)
    cast_f32_bf16_kernel<<<(unsigned)((n / 2 + 256) / 256), 256, 0, ST(stream)>>>((const float*)in, (__nv_bfloat16*)out, n);
  else
    cast_bf16_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ST(stream)>>>((const __nv_bfloat16*)in, (float*)out, n);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

// DUAN.forward (model.py:989-1035).  x, c: fp32 [B, C, L].  Workspace: fp32 [B * (C * 9 + 128 + 2 * C) ] (see below).
extern "C" int lx_duan_forward(const lx_duan_weights_t* w, const float* x, const float* c, float* y, int64_t y_bstride,
                               int32_t B, int32_t Cc, int32_t L, float keep_ratio, float* workspace, void* stream) {
  LX_CHECK_ARG(w && x && c && y && workspace, "lx_duan_forward: null pointer");
  LX_CHECK_ARG(B > 0 && Cc > 0 && L > 0 && w->hidden > 0 && w->hidden <= 1024, "lx_duan_forward: bad shape");
  const int Hd = w->hidden;
  const size_t BC = (size_t)B * Cc;
  float* mean_x = workspace;
  float* m2_x = mean_x + BC;
  float* mean_c = m2_x + BC;
  float* g_mix = mean_c + BC;
  float* mu = g_mix + BC;
  float* rsig = mu + BC;
  float* g1 = rsig + BC;
  float* beta = g1 + BC;
  float* imp = beta + BC;
  float* mask = imp + BC;
  float* gb = mask + BC;               // [B, 2C]
  float* hid_pool = gb + 2 * BC;       // [B, Hd]
  float* hid = hid_pool + (size_t)B * Hd;  // [B, Hd, L] gate hidden
  float* rowpart = hid + (size_t)B * Hd * L;  // [B, C, ceil(L / 64)] partial gate sums (deterministic mean over L)
  cudaStream_t st = ST(stream);

  duan_row_stats_kernel<<<(unsigned)BC, 256, 0, st>>>(x, c, mean_x, m2_x, mean_c, L);
  LX_CUDA(cudaGetLastError());
  // gate: g_mix = mean_L sigmoid(W2 relu(W1 c + b1) + b2)
  lx_sgemm_desc_t g;
  memset(&g, 0, sizeof(g));
  g.A = w->gate_w1; g.lda = Cc; g.Bm = c; g.ldb = L; g.b_bstride = (int64_t)Cc * L; g.bias = w->gate_b1;
  g.C = hid; g.ldc = L; g.c_bstride = (int64_t)Hd * L; g.M = Hd; g.N = L; g.K = Cc; g.batch = B; g.act = 1;
  int rc = lx_sgemm_f32(&g, stream);
  if (rc) return rc;
  memset(&g, 0, sizeof(g));
  g.A = w->gate_w2; g.lda = Hd; g.Bm = hid; g.ldb = L; g.b_bstride = (int64_t)Hd * L; g.bias = w->gate_b2;
  g.rowmean = g_mix; g.rowpart = rowpart; g.M = Cc; g.N = L; g.K = Hd; g.batch = B; g.act = 2;
  rc = lx_sgemm_f32(&g, stream);
  if (rc) return rc;
  // FiLM: [gamma, beta] = W4 relu(W3 mean_L(c) + b3) + b4     (N = 1 "GEMMs")
  memset(&g, 0, sizeof(g));
  g.A = w->mlp_w1; g.lda = Cc; g.Bm = mean_c; g.ldb = 1; g.b_bstride = Cc; g.bias = w->mlp_b1;
  g.C = hid_pool; g.ldc = 1; g.c_bstride = Hd; g.M = Hd; g.N = 1; g.K = Cc; g.batch = B; g.act = 1;
  rc = lx_sgemm_f32(&g, stream);
  if (rc) return rc;
  memset(&g, 0, sizeof(g));
  g.A = w->mlp_w2; g.lda = Hd; g.Bm = hid_pool; g.ldb = 1; g.b_bstride = Hd; g.bias = w->mlp_b2;
  g.C = gb; g.ldc = 1; g.c_bstride = 2 * Cc; g.M = 2 * Cc; g.N = 1; g.K = Hd; g.batch = B; g.act = 0;
  rc = lx_sgemm_f32(&g, stream);
  if (rc) return rc;
  duan_mix_kernel<<<B, 256, 0, st>>>(mean_x, m2_x, g_mix, gb, mu, rsig, g1, beta, Cc, L, w->eps);
  LX_CUDA(cudaGetLastError());
  duan_importance_kernel<<<(unsigned)BC, 256, 0, st>>>(x, mu, rsig, g1, beta, imp, L);
  LX_CUDA(cudaGetLastError());
  int k = (int)((float)Cc * keep_ratio);
  k = (int)((double)Cc * (double)keep_ratio);  // int(C * keep_ratio) as Python computes it
  if (k < 1) k = 1;
  duan_topk_mask_kernel<<<B, 256, Cc * sizeof(float), st>>>(imp, mask, Cc, k);
  LX_CUDA(cudaGetLastError());
  dim3 ga((L + 255) / 256, (unsigned)BC);
  duan_apply_kernel<<<ga, 256, 0, st>>>(x, mu, rsig, g1, beta, mask, y, Cc, L, y_bstride);
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
