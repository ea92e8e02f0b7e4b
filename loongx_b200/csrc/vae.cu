// VAE kernels: the 16-channel FLUX AutoencoderKL either side of the denoising loop (reference call
// sites pipeline_tools.py:7-12 `vae.encode(..).latent_dist.sample()` and generate.py:375-380 `vae.decode` +
// `image_processor.postprocess`; the arithmetic is diffusers 0.31.0's, restated in oracle/vae.py).
//
// Layout: activations are bf16 rows [B*H*W, C] (NHWC), so every convolution is the tcgen05 GEMM of gemm.cu over an
// im2col panel [B*Ho*Wo, taps*C] whose columns are ordered (ky, kx, c).  GroupNorm + SiLU — which precede every
// convolution of a ResnetBlock2D — are folded into the panel write: the normalised activation is never stored.
// GroupNorm statistics are fp32 partial sums per thread combined in fp64 in a fixed order (sum and sum of squares; the
// subtraction E[x^2] - mean^2 happens in fp64): the whole VAE is bit-reproducible run to run.  All kernels here are HBM / L2-bound: 16-byte accesses, each operand touched once
// (the nine taps of a panel re-read the activation through L2).
#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {

namespace {

__device__ __forceinline__ void vld8(const __nv_bfloat16* p, float* x) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}

// SiLU as 0.5 x (1 + tanh(x / 2)): one MUFU op per element (the panel write applies it once per tap, nine times per
// activation element, so exp + IEEE division would make the kernel MUFU-bound); tanh.approx is good to ~2^-11, the
// result is stored in bf16.
__device__ __forceinline__ float silu_tanh(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
  return 0.5f * x * (1.0f + t);
}

inline cudaStream_t vcs(void* s) { return static_cast<cudaStream_t>(s); }

// ---------------------------------------------------------------------------------------------------------------
// GroupNorm statistics, two deterministic stages (no atomics, nothing to zero):
//   gn_stats_kernel     grid (row chunks, B), 256 threads; thread = (row lane, 8-channel vector).  fp32 partial sums per
//                       thread over <= 128/lanes rows, combined through shared memory in a fixed order into one fp64
//                       (sum, sum of squares) per (sample, chunk, group).
//   gn_finalize_kernel  one CTA per (sample, group): fixed-order fp64 reduction over the chunks, then the per-channel
//                       affine coefficients.
// ---------------------------------------------------------------------------------------------------------------
constexpr int GN_THREADS = 256;
constexpr int GN_ROWS_PER_CTA = 128;  // many small CTAs: the kernel is latency-bound with few loads in flight per thread

__global__ void __launch_bounds__(GN_THREADS) gn_stats_kernel(const __nv_bfloat16* __restrict__ x, double* __restrict__ part,
                                                              long long hw, int C, int groups) {
  __shared__ float sh_s[GN_THREADS * 8], sh_q[GN_THREADS * 8];  // [row lane][channel] (lanes * C = 2048 floats each)
  pdl_wait();
  pdl_launch_dependents();
  const int vec = C / 8, lanes = GN_THREADS / vec;
  const int v = threadIdx.x % vec, r0 = threadIdx.x / vec;
  const long long row_begin = (long long)blockIdx.x * GN_ROWS_PER_CTA;
  const long long row_end = min(hw, row_begin + GN_ROWS_PER_CTA);
  const __nv_bfloat16* xb = x + (size_t)blockIdx.y * hw * C + v * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
#pragma unroll 4
  for (long long r = row_begin + r0; r < row_end; r += lanes) {
    float t[8];
    vld8(xb + (size_t)r * C, t);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += t[e];
      q[e] = fmaf(t[e], t[e], q[e]);
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sh_s[r0 * C + v * 8 + e] = s[e];
    sh_q[r0 * C + v * 8 + e] = q[e];
  }
  __syncthreads();
  const int cpg = C / groups;
  for (int g = threadIdx.x; g < groups; g += GN_THREADS) {
    double a = 0.0, b = 0.0;
    for (int l = 0; l < lanes; ++l)
      for (int c = 0; c < cpg; ++c) {
        a += (double)sh_s[l * C + g * cpg + c];
        b += (double)sh_q[l * C + g * cpg + c];
      }
    double* dst = part + (((size_t)blockIdx.y * gridDim.x + blockIdx.x) * groups + g) * 2;
    dst[0] = a;
    dst[1] = b;
  }
}

// per (sample, group): sum the chunk partials, then y = x * a + b with a = gamma * rstd, b = beta - mean * a
__global__ void __launch_bounds__(GN_THREADS) gn_finalize_kernel(const double* __restrict__ part, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float2* __restrict__ coeff,
                                                                 int chunks, long long hw, int C, int groups, float eps) {
  __shared__ double sh[2][GN_THREADS];
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x / groups, g = blockIdx.x % groups, cpg = C / groups;
  double a = 0.0, q = 0.0;
  for (int ch = threadIdx.x; ch < chunks; ch += GN_THREADS) {
    const double* src = part + (((size_t)b * chunks + ch) * groups + g) * 2;
    a += src[0];
    q += src[1];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = q;
  __syncthreads();
  for (int o = GN_THREADS / 2; o > 0; o >>= 1) {  // fixed tree: deterministic
    if ((int)threadIdx.x < o) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
    }
    __syncthreads();
  }
  const double n = (double)hw * cpg;
  const double mean = sh[0][0] / n;
  double var = sh[1][0] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  for (int c = threadIdx.x; c < cpg; c += GN_THREADS) {
    const int cc = g * cpg + c;
    const float ga = gamma[cc] * rstd;
    coeff[(size_t)b * C + cc] = make_float2(ga, beta[cc] - (float)mean * ga);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// im2col panel write with the optional GroupNorm affine + SiLU, nearest x2 up-sampling and stride.
// Thread block = (16-byte vectors of one panel row, rows): threadIdx.y picks the output pixel, threadIdx.x the vector
// j -> (tap, channel vector), so the only integer divisions are the two that decode the pixel (the channel-vector count
// is a power of two for every FLUX layer and becomes a shift).
// ---------------------------------------------------------------------------------------------------------------
struct Im2col {
  const __nv_bfloat16* x;
  __nv_bfloat16* out;
  const float2* coeff;  // [B, C] or null
  int H, W, C;          // stored input
  int up_shift;         // 0: as stored, 1: nearest x2 virtual image
  int Ho, Wo, stride, pad_lo, taps, silu;
  int M;         // output pixels = panel rows
  int kv;        // 16-byte vectors per panel row (ldk / 8)
  int cv_shift;  // log2(C / 8), or -1 when C / 8 is not a power of two
  long long ldk;
};

__global__ void __launch_bounds__(1024) im2col_kernel(const Im2col p) {
  pdl_wait();
  pdl_launch_dependents();
  const int m = blockIdx.x * blockDim.y + threadIdx.y;
  if (m >= p.M) return;
  const int hwo = p.Ho * p.Wo;
  const int b = m / hwo, rem = m - b * hwo;
  const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
  const int cv = p.C / 8;
  __nv_bfloat16* orow = p.out + (size_t)m * p.ldk;
  for (int j = threadIdx.x; j < p.kv; j += blockDim.x) {
    int tap, v;
    if (p.cv_shift >= 0) {
      tap = j >> p.cv_shift;
      v = j & (cv - 1);
    } else {
      tap = j / cv;
      v = j - tap * cv;
    }
    uint4 o = make_uint4(0u, 0u, 0u, 0u);
    if (tap < p.taps) {
      const int ky = p.taps == 9 ? tap / 3 : 0, kx = p.taps == 9 ? tap - 3 * ky : 0;
      const int iy = oy * p.stride + ky - p.pad_lo, ix = ox * p.stride + kx - p.pad_lo;
      if (iy >= 0 && ix >= 0 && iy < (p.H << p.up_shift) && ix < (p.W << p.up_shift)) {
        const int sy = iy >> p.up_shift, sx = ix >> p.up_shift;
        const __nv_bfloat16* src = p.x + (((size_t)b * p.H + sy) * p.W + sx) * p.C + v * 8;
        if (p.coeff == nullptr) {
          o = *reinterpret_cast<const uint4*>(src);
        } else {
          float t[8];
          vld8(src, t);
          const float4* cf = reinterpret_cast<const float4*>(p.coeff + (size_t)b * p.C + v * 8);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 ab = __ldg(cf + e);  // (a, b) of two channels
            t[2 * e] = fmaf(t[2 * e], ab.x, ab.y);
            t[2 * e + 1] = fmaf(t[2 * e + 1], ab.z, ab.w);
          }
          if (p.silu) {
#pragma unroll
            for (int e = 0; e < 8; ++e) t[e] = silu_tanh(t[e]);
          }
          o.x = pack_bf16(t[0], t[1]);
          o.y = pack_bf16(t[2], t[3]);
          o.z = pack_bf16(t[4], t[5]);
          o.w = pack_bf16(t[6], t[7]);
        }
      }
    }
    *reinterpret_cast<uint4*>(orow + (size_t)j * 8) = o;
  }
}

// Scatter form of the same panel for the common case (3x3 kernel, ldk == 9 C): thread = (padded input pixel, channel
// vector).  The activation element is loaded, normalised and passed through SiLU ONCE and stored to the (up to) nine
// panel slots it occupies; the zero border is written by the threads of the padding pixels, so every (row, tap) slot is
// written exactly once and nothing is pre-zeroed.  Per 16-byte store this is ~1/9 of the gather form's arithmetic and L1
// traffic, which leaves the kernel bound by the panel write itself.
__global__ void __launch_bounds__(1024) im2col_scatter_kernel(const Im2col p) {
  pdl_wait();
  pdl_launch_dependents();
  const int Hv = p.H << p.up_shift, Wv = p.W << p.up_shift;  // the image the convolution sees
  const int Hp = Hv + 2, Wp = Wv + 2;                         // with one padding pixel on every side
  const long long pp = (long long)blockIdx.x * blockDim.y + threadIdx.y;
  const int b = (int)(pp / ((long long)Hp * Wp));
  if (b >= p.M) return;  // (p.M holds the batch size in this kernel)
  const int rem = (int)(pp - (long long)b * Hp * Wp);
  const int py = rem / Wp - 1, px = rem - (py + 1) * Wp - 1;
  const int v = threadIdx.x;
  uint4 o = make_uint4(0u, 0u, 0u, 0u);
  if (py >= 0 && py < Hv && px >= 0 && px < Wv) {
    const __nv_bfloat16* src = p.x + (((size_t)b * p.H + (py >> p.up_shift)) * p.W + (px >> p.up_shift)) * p.C + v * 8;
    if (p.coeff == nullptr) {
      o = *reinterpret_cast<const uint4*>(src);
    } else {
      float t[8];
      vld8(src, t);
      const float4* cf = reinterpret_cast<const float4*>(p.coeff + (size_t)b * p.C + v * 8);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 ab = __ldg(cf + e);
        t[2 * e] = fmaf(t[2 * e], ab.x, ab.y);
        t[2 * e + 1] = fmaf(t[2 * e + 1], ab.z, ab.w);
      }
      if (p.silu) {
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = silu_tanh(t[e]);
      }
      o.x = pack_bf16(t[0], t[1]);
      o.y = pack_bf16(t[2], t[3]);
      o.z = pack_bf16(t[4], t[5]);
      o.w = pack_bf16(t[6], t[7]);
    }
  }
  const int smask = p.stride - 1, sshift = p.stride >> 1;  // stride 1 or 2
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int ty = py + p.pad_lo - ky;
    if (ty < 0 || (ty & smask) || (ty >> sshift) >= p.Ho) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int tx = px + p.pad_lo - kx;
      if (tx < 0 || (tx & smask) || (tx >> sshift) >= p.Wo) continue;
      const size_t row = ((size_t)b * p.Ho + (ty >> sshift)) * p.Wo + (tx >> sshift);
      *reinterpret_cast<uint4*>(p.out + row * p.ldk + (size_t)(ky * 3 + kx) * p.C + v * 8) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Row softmax of the mid-block attention (one head of width C over H*W positions): fp32 logits -> bf16 probabilities.
// One CTA per row; three passes over a row that stays in L1/L2 (<= 64 KB).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, float* sh, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // sh may still be read from the previous reduction
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < (int)(blockDim.x >> 5); ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];
  return r;
}

__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, long long lds,
                                                           __nv_bfloat16* __restrict__ p, long long ldp, int n, int n_pad,
                                                           float scale_log2e) {
  __shared__ float sh[8];
  pdl_wait();
  pdl_launch_dependents();
  const float* row = s + (size_t)blockIdx.x * lds;
  __nv_bfloat16* prow = p + (size_t)blockIdx.x * ldp;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < n; c += 256) mx = fmaxf(mx, row[c]);
  mx = block_reduce(mx, sh, true);
  float sum = 0.f;
  for (int c = threadIdx.x; c < n; c += 256) sum += exp2f((row[c] - mx) * scale_log2e);
  sum = block_reduce(sum, sh, false);
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < n_pad; c += 256)
    prow[c] = __float2bfloat16(c < n ? exp2f((row[c] - mx) * scale_log2e) * inv : 0.f);
}

// ---------------------------------------------------------------------------------------------------------------
// Layout changes at the two ends: fp32 NCHW <-> rows
// ---------------------------------------------------------------------------------------------------------------
// fp32 [B, C, hw] -> bf16 rows [B*hw, c_pad] = in * mul + add (channels >= C are zero)
__global__ void nchw_to_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int C, long long hw,
                                    int c_pad, float mul, float add, long long total) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // (row, 8-channel vector)
  if (idx >= total) return;
  const int cv = c_pad / 8;
  const long long m = idx / cv;
  const int c0 = (int)(idx % cv) * 8;
  const long long b = m / hw, pix = m % hw;
  float t[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) t[e] = (c0 + e < C) ? fmaf(in[((size_t)b * C + c0 + e) * hw + pix], mul, add) : 0.f;
  uint4 o;
  o.x = pack_bf16(t[0], t[1]);
  o.y = pack_bf16(t[2], t[3]);
  o.z = pack_bf16(t[4], t[5]);
  o.w = pack_bf16(t[6], t[7]);
  *reinterpret_cast<uint4*>(out + (size_t)m * c_pad + c0) = o;
}

// fp32 rows [B*hw, ld] -> fp32 [B, C, hw]; denormalize = (x / 2 + 0.5).clamp(0, 1) (VaeImageProcessor.denormalize)
__global__ void rows_to_nchw_kernel(const float* __restrict__ in, long long ld, float* __restrict__ out, int C, long long hw,
                                    int denormalize, long long total) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over [B, C, hw]
  if (idx >= total) return;
  const long long pix = idx % hw, bc = idx / hw;
  const int c = (int)(bc % C);
  const long long b = bc / C;
  float v = in[((size_t)b * hw + pix) * ld + c];
  if (denormalize) v = fminf(fmaxf(v * 0.5f + 0.5f, 0.f), 1.f);
  out[idx] = v;
}

// DiagonalGaussianDistribution.sample + pipeline_tools.py:11-12:
// out[b, l, pix] = (mean + exp(0.5 * clamp(logvar, -30, 20)) * eps - shift) * scale, moments rows = [mean(L) | logvar(L)]
__global__ void sample_latents_kernel(const float* __restrict__ mom, long long ld, const float* __restrict__ eps,
                                      float* __restrict__ out, int L, long long hw, float shift, float scale, long long total) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // over [B, L, hw]
  if (idx >= total) return;
  const long long pix = idx % hw, bl = idx / hw;
  const int l = (int)(bl % L);
  const long long b = bl / L;
  const float* r = mom + ((size_t)b * hw + pix) * ld;
  float z = r[l];
  if (eps != nullptr) {
    const float lv = fminf(fmaxf(r[L + l], -30.f), 20.f);
    z = fmaf(expf(0.5f * lv), eps[idx], z);
  }
  out[idx] = (z - shift) * scale;
}

}  // namespace
}  // namespace lx

using namespace lx;

static int g_im2col_force_gather = 0;
// development / test knob: 1 = always use the gather form of the panel kernel
extern "C" void lx_debug_vae_im2col_gather(int on) { g_im2col_force_gather = on; }

extern "C" int64_t lx_vae_group_norm_workspace(int32_t B, int64_t hw, int32_t groups) {
  return 2 * (int64_t)B * ((hw + GN_ROWS_PER_CTA - 1) / GN_ROWS_PER_CTA) * groups;
}

extern "C" int lx_vae_group_norm_coeffs(const void* x, int32_t B, int64_t hw, int32_t C, int32_t groups, const float* gamma,
                                        const float* beta, float eps, double* workspace, float* coeff, void* stream) {
  LX_CHECK_ARG(x && gamma && beta && workspace && coeff && B > 0 && hw > 0, "lx_vae_group_norm_coeffs: null argument or empty input");
  LX_CHECK_ARG(C >= 8 && C % 8 == 0 && GN_THREADS % (C / 8) == 0 && groups > 0 && C % groups == 0 && groups <= GN_THREADS,
               "lx_vae_group_norm_coeffs: C=%d must be 8 * a power of two <= 2048 and divisible by groups=%d", C, groups);
  const long long chunks = (hw + GN_ROWS_PER_CTA - 1) / GN_ROWS_PER_CTA;
  LX_CHECK_ARG(chunks < 65536 * 32, "lx_vae_group_norm_coeffs: hw=%lld too large", (long long)hw);
  LaunchScope scope(KC_ROW, stream, 2.0 * B * (double)hw * C);
  LX_CUDA(launch_pdl(gn_stats_kernel, dim3((unsigned)chunks, B), dim3(GN_THREADS), 0, vcs(stream),
                     reinterpret_cast<const __nv_bfloat16*>(x), workspace, (long long)hw, C, groups));
  LX_CUDA(launch_pdl(gn_finalize_kernel, dim3(B * groups), dim3(GN_THREADS), 0, vcs(stream), (const double*)workspace, gamma, beta,
                     reinterpret_cast<float2*>(coeff), (int)chunks, (long long)hw, C, groups, eps));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_vae_im2col(const lx_vae_im2col_desc_t* d, void* stream) {
  LX_CHECK_ARG(d && d->x && d->out, "lx_vae_im2col: null argument");
  LX_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->C >= 8 && d->C % 8 == 0, "lx_vae_im2col: bad input shape (C=%d must be a multiple of 8)",
               d->C);
  LX_CHECK_ARG(d->taps == 1 || d->taps == 9, "lx_vae_im2col: taps=%d must be 1 or 9", d->taps);
  LX_CHECK_ARG(d->upsample == 1 || d->upsample == 2, "lx_vae_im2col: upsample=%d must be 1 or 2", d->upsample);
  LX_CHECK_ARG(d->stride == 1 || d->stride == 2, "lx_vae_im2col: stride=%d must be 1 or 2", d->stride);
  LX_CHECK_ARG(d->pad_lo == 0 || d->pad_lo == 1, "lx_vae_im2col: pad_lo=%d must be 0 or 1", d->pad_lo);
  LX_CHECK_ARG(d->taps == 9 || (d->stride == 1 && d->pad_lo == 0), "lx_vae_im2col: a 1-tap panel has stride 1 and no padding");
  LX_CHECK_ARG(d->Ho > 0 && d->Wo > 0 && (d->Ho - 1) * d->stride - d->pad_lo < d->H * d->upsample &&
                   (d->Wo - 1) * d->stride - d->pad_lo < d->W * d->upsample,
               "lx_vae_im2col: output %dx%d does not fit the input", d->Ho, d->Wo);
  LX_CHECK_ARG(d->ldk % 8 == 0 && d->ldk >= (int64_t)d->taps * d->C, "lx_vae_im2col: ldk=%lld must be a multiple of 8 and >= taps*C",
               (long long)d->ldk);
  Im2col p;
  p.x = reinterpret_cast<const __nv_bfloat16*>(d->x);
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  p.coeff = reinterpret_cast<const float2*>(d->coeff);
  p.H = d->H; p.W = d->W; p.C = d->C;
  p.up_shift = d->upsample == 2 ? 1 : 0;
  p.Ho = d->Ho; p.Wo = d->Wo; p.stride = d->stride; p.pad_lo = d->pad_lo; p.taps = d->taps; p.silu = d->silu;
  p.ldk = d->ldk;
  const long long M = (long long)d->B * d->Ho * d->Wo;
  LX_CHECK_ARG(M < (1LL << 31) && d->ldk / 8 < (1LL << 31), "lx_vae_im2col: panel too large for one launch");
  p.M = (int)M;
  p.kv = (int)(d->ldk / 8);
  const int cv = d->C / 8;
  p.cv_shift = -1;
  for (int sft = 0; sft < 20; ++sft)
    if ((1 << sft) == cv) p.cv_shift = sft;
  LaunchScope scope(KC_ROW, stream, 2.0 * d->B * (double)d->H * d->W * d->C + 2.0 * M * (double)d->ldk);
  if (d->taps == 9 && d->ldk == 9LL * d->C && cv <= 1024 && !g_im2col_force_gather) {
    // scatter form: one thread per (padded input pixel, channel vector)
    const int ty = max(1, 1024 / cv);
    const long long pixels = (long long)d->B * ((d->H << p.up_shift) + 2) * ((d->W << p.up_shift) + 2);
    LX_CHECK_ARG((pixels + ty - 1) / ty < (1LL << 31), "lx_vae_im2col: input too large for one launch");
    p.M = d->B;
    LX_CUDA(launch_pdl(im2col_scatter_kernel, dim3((unsigned)((pixels + ty - 1) / ty)), dim3(cv, ty), 0, vcs(stream), p));
  } else {
    const int tx = min((p.kv + 31) / 32 * 32, 1024), ty = max(1, 1024 / tx);
    LX_CUDA(launch_pdl(im2col_kernel, dim3((unsigned)((M + ty - 1) / ty)), dim3(tx, ty), 0, vcs(stream), p));
  }
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_vae_softmax_rows(const float* s, int64_t lds, void* p, int64_t ldp, int32_t rows, int32_t n, float scale,
                                   void* stream) {
  LX_CHECK_ARG(s && p && rows > 0 && n > 0 && lds >= n && ldp >= n, "lx_vae_softmax_rows: bad arguments");
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * (double)n);
  LX_CUDA(launch_pdl(softmax_rows_kernel, dim3(rows), dim3(256), 0, vcs(stream), s, (long long)lds,
                     reinterpret_cast<__nv_bfloat16*>(p), (long long)ldp, n, (int)ldp, scale * 1.4426950408889634f));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_vae_nchw_to_rows(const float* in, void* out, int32_t B, int32_t C, int64_t hw, int32_t c_pad, float mul,
                                   float add, void* stream) {
  LX_CHECK_ARG(in && out && B > 0 && C > 0 && hw > 0 && c_pad >= C && c_pad % 8 == 0, "lx_vae_nchw_to_rows: bad arguments");
  const long long total = (long long)B * hw * (c_pad / 8);
  LaunchScope scope(KC_ROW, stream, 4.0 * B * (double)hw * C + 2.0 * B * (double)hw * c_pad);
  LX_CUDA(launch_pdl(nchw_to_rows_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, vcs(stream), in,
                     reinterpret_cast<__nv_bfloat16*>(out), C, (long long)hw, c_pad, mul, add, total));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_vae_rows_to_nchw(const float* in, int64_t ld, float* out, int32_t B, int32_t C, int64_t hw, int32_t denormalize,
                                   void* stream) {
  LX_CHECK_ARG(in && out && B > 0 && C > 0 && hw > 0 && ld >= C, "lx_vae_rows_to_nchw: bad arguments");
  const long long total = (long long)B * C * hw;
  LaunchScope scope(KC_ROW, stream, 8.0 * total);
  LX_CUDA(launch_pdl(rows_to_nchw_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, vcs(stream), in, (long long)ld, out,
                     C, (long long)hw, denormalize, total));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_vae_sample_latents(const float* moments, int64_t ld, const float* eps, float* out, int32_t B, int32_t L,
                                     int64_t hw, float shift, float scale, void* stream) {
  LX_CHECK_ARG(moments && out && B > 0 && L > 0 && hw > 0 && ld >= 2 * L, "lx_vae_sample_latents: bad arguments");
  const long long total = (long long)B * L * hw;
  LaunchScope scope(KC_ROW, stream, 16.0 * total);
  LX_CUDA(launch_pdl(sample_latents_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, vcs(stream), moments,
                     (long long)ld, eps, out, L, (long long)hw, shift, scale, total));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
