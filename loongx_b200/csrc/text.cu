// Text-encoder kernels: the pieces of transformers' T5EncoderModel / CLIPTextModel — which
// FluxPipeline.encode_prompt runs once per edit (generate.py:156-165, pipeline_tools.py:33-52) — that are not a Linear:
// token (+ position) embedding gather, T5 RMS norm / CLIP LayerNorm with fp32 weights, 64-wide-head attention with an
// additive bias table (T5 relative positions) or a causal mask (CLIP), and the gated-GELU product.  Every Linear is the
// tcgen05 GEMM of gemm.cu (fused q|k|v and wi_0|wi_1 panels, residual adds in its GATE_RESIDUAL epilogue).
// Sequences are short (512 / 77 tokens, once per edit): these kernels are latency / L2-bound, written for simplicity.
#include "host_util.cuh"
#include "ptx.cuh"

namespace lx {
namespace {

inline cudaStream_t tcs(void* s) { return static_cast<cudaStream_t>(s); }

__device__ __forceinline__ void tld8(const __nv_bfloat16* p, float* x) {
  uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
__device__ __forceinline__ void tst8(__nv_bfloat16* p, const float* v) {
  uint4 u;
  u.x = pack_bf16(v[0], v[1]); u.y = pack_bf16(v[2], v[3]); u.z = pack_bf16(v[4], v[5]); u.w = pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}

// out[i, :] = table[ids[i], :] (+ pos[i % period, :])
__global__ void embed_rows_kernel(const __nv_bfloat16* __restrict__ table, const int* __restrict__ ids,
                                  const __nv_bfloat16* __restrict__ pos, int period, __nv_bfloat16* __restrict__ out, int n,
                                  int D8, int vocab) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * D8) return;
  const int i = (int)(idx / D8), c = (int)(idx % D8) * 8;
  int id = ids[i];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  float a[8];
  tld8(table + (size_t)id * D8 * 8 + c, a);
  if (pos != nullptr) {
    float b[8];
    tld8(pos + (size_t)(i % period) * D8 * 8 + c, b);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += b[e];
  }
  tst8(out + (size_t)i * D8 * 8 + c, a);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
  return r;
}

// rms != 0: y = x * rsqrt(mean(x^2) + eps) * gamma (T5LayerNorm);  else y = (x - mean) * rstd * gamma + beta (nn.LayerNorm)
__global__ void __launch_bounds__(128) norm_rows_kernel(const __nv_bfloat16* __restrict__ x, long long ldx,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        __nv_bfloat16* __restrict__ out, long long ldo, int D, float eps, int rms) {
  __shared__ float red[4];
  pdl_wait();
  pdl_launch_dependents();
  const __nv_bfloat16* xr = x + (size_t)blockIdx.x * ldx;
  __nv_bfloat16* orow = out + (size_t)blockIdx.x * ldo;
  const int nch = D >> 3;
  float mean = 0.f;
  if (!rms) {
    float s = 0.f;
    for (int ch = threadIdx.x; ch < nch; ch += 128) {
      float v[8];
      tld8(xr + ch * 8, v);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[e];
    }
    mean = block_sum(s, red) / D;
  }
  float q = 0.f;
  for (int ch = threadIdx.x; ch < nch; ch += 128) {
    float v[8];
    tld8(xr + ch * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) q = fmaf(v[e] - mean, v[e] - mean, q);
  }
  const float rstd = rsqrtf(block_sum(q, red) / D + eps);
  for (int ch = threadIdx.x; ch < nch; ch += 128) {
    float v[8];
    tld8(xr + ch * 8, v);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float g = gamma[ch * 8 + e];
      v[e] = (v[e] - mean) * rstd * g + (beta != nullptr ? beta[ch * 8 + e] : 0.f);
    }
    tst8(orow + ch * 8, v);
  }
}

// out = a * b over [rows, cols] views
__global__ void mul_rows_kernel(const __nv_bfloat16* a, long long lda, const __nv_bfloat16* b, long long ldb,
                                __nv_bfloat16* out /* may alias a or b */, long long ldo, int rows, int cols8) {
  pdl_wait();
  pdl_launch_dependents();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)rows * cols8) return;
  const int r = (int)(idx / cols8), c = (int)(idx % cols8) * 8;
  float x[8], y[8];
  tld8(a + (size_t)r * lda + c, x);
  tld8(b + (size_t)r * ldb + c, y);
#pragma unroll
  for (int e = 0; e < 8; ++e) x[e] *= y[e];
  tst8(out + (size_t)r * ldo + c, x);
}

// ---------------------------------------------------------------------------------------------------------------
// Attention for 64-wide heads over short sequences (S <= 512).  One CTA per (batch, head): K and V of the head live in
// shared memory (row stride 66 bf16 = 33 words, so a warp reading one word of 32 different keys hits 32 banks); each warp
// owns query rows i = warp, warp + 16, ... (of the CTA's slice when the rows of a head are split over several CTAs): lane l scores keys l, l+32, ... (q in registers), the softmax runs over the
// warp, the probabilities go through a per-warp shared buffer, and lane l accumulates output dimensions 2l, 2l+1.
// ---------------------------------------------------------------------------------------------------------------
constexpr int SA_DH = 64, SA_WARPS = 16, SA_MAX_S = 512, SA_KSTRIDE = SA_DH + 2;

struct SmallAttn {
  const __nv_bfloat16 *q, *k, *v;  // rows [B*S, ld], head h at column h * 64
  long long ldq, ldk, ldv;
  __nv_bfloat16* out;  // rows [B*S, ldo], head h at column h * 64
  long long ldo;
  const float* bias;  // [H, S, S] additive (may be null)
  int B, H, S, causal;
  int bias_rel;  // bias is [H, 2S-1] indexed by key - query + S - 1
  int qsplit;    // CTAs per (batch, head): each takes a contiguous slice of the query rows (all keys)
  float scale;
};

__global__ void __launch_bounds__(SA_WARPS * 32) small_attention_kernel(const SmallAttn p) {
  extern __shared__ __align__(16) unsigned char sa_smem[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(sa_smem);
  __nv_bfloat16* Vs = Ks + (size_t)p.S * SA_KSTRIDE;
  float* Ps = reinterpret_cast<float*>(Vs + (size_t)p.S * SA_KSTRIDE);  // [SA_WARPS][S]
  pdl_wait();
  pdl_launch_dependents();
  const int bh = blockIdx.x / p.qsplit, part = blockIdx.x % p.qsplit;
  const int b = bh / p.H, h = bh % p.H;
  const int rows_per_part = (p.S + p.qsplit - 1) / p.qsplit;
  const int i_begin = part * rows_per_part, i_end = min(p.S, i_begin + rows_per_part);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t row0 = (size_t)b * p.S;
  // stage K and V: one 4-byte word (two dims) per thread step
  for (int idx = threadIdx.x; idx < p.S * (SA_DH / 2); idx += SA_WARPS * 32) {
    const int j = idx / (SA_DH / 2), w = idx % (SA_DH / 2);
    reinterpret_cast<uint32_t*>(Ks + (size_t)j * SA_KSTRIDE)[w] =
        reinterpret_cast<const uint32_t*>(p.k + (row0 + j) * p.ldk + h * SA_DH)[w];
    reinterpret_cast<uint32_t*>(Vs + (size_t)j * SA_KSTRIDE)[w] =
        reinterpret_cast<const uint32_t*>(p.v + (row0 + j) * p.ldv + h * SA_DH)[w];
  }
  __syncthreads();
  float* Pw = Ps + (size_t)warp * p.S;
  const int nj = (p.S + 31) / 32;
  for (int i = i_begin + warp; i < i_end; i += SA_WARPS) {
    float2 q[SA_DH / 2];
    const uint32_t* qg = reinterpret_cast<const uint32_t*>(p.q + (row0 + i) * p.ldq + h * SA_DH);
#pragma unroll
    for (int w = 0; w < SA_DH / 2; ++w) q[w] = unpack_bf16(qg[w]);
    float sc[SA_MAX_S / 32];
    float mx = -INFINITY;
#pragma unroll
    for (int jj = 0; jj < SA_MAX_S / 32; ++jj) {
      sc[jj] = -INFINITY;
      const int j = jj * 32 + lane;
      if (jj < nj && j < p.S && !(p.causal && j > i)) {
        const uint32_t* kr = reinterpret_cast<const uint32_t*>(Ks + (size_t)j * SA_KSTRIDE);
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < SA_DH / 2; ++w) {
          const float2 kk = unpack_bf16(kr[w]);
          acc = fmaf(q[w].x, kk.x, acc);
          acc = fmaf(q[w].y, kk.y, acc);
        }
        acc *= p.scale;
        if (p.bias != nullptr)
          acc += p.bias_rel ? p.bias[(size_t)h * (2 * p.S - 1) + (j - i + p.S - 1)] : p.bias[((size_t)h * p.S + i) * p.S + j];
        sc[jj] = acc;
        mx = fmaxf(mx, acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int jj = 0; jj < SA_MAX_S / 32; ++jj) {
      const int j = jj * 32 + lane;
      if (jj < nj && j < p.S) {
        const float e = sc[jj] == -INFINITY ? 0.f : __expf(sc[jj] - mx);
        Pw[j] = e;
        sum += e;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    float ox = 0.f, oy = 0.f;
    const int jend = p.causal ? i + 1 : p.S;
    for (int j = 0; j < jend; ++j) {
      const float pj = Pw[j];
      const float2 vv = unpack_bf16(reinterpret_cast<const uint32_t*>(Vs + (size_t)j * SA_KSTRIDE)[lane]);
      ox = fmaf(pj, vv.x, ox);
      oy = fmaf(pj, vv.y, oy);
    }
    const float inv = 1.0f / sum;
    reinterpret_cast<uint32_t*>(p.out + (row0 + i) * p.ldo + h * SA_DH)[lane] = pack_bf16(ox * inv, oy * inv);
    __syncwarp();  // Pw is rewritten by the next row
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core form of the same attention (the default): mma.sync m16n8k16 bf16 tiles with an online softmax.  One CTA per
// (batch, head, 128 query rows), 8 warps x 16 query rows; K [key][dim] and V^T [dim][key] of the head in shared memory
// with row strides of 36 / (S64 + 8) / 2 words = 4 (mod 8), so the eight row groups x four lanes of a fragment load hit 32
// different banks.  Per 64-key block: S = Q K^T (32 MMAs), scale / bias / mask in the accumulator layout, running max and
// sum per row (the four lanes of a row group share a row), P re-used directly as the A fragments of O += P V (32 MMAs).
// tcgen05 would be the wrong tool here: 64-wide heads, 77..512 keys, once per edit.
// ---------------------------------------------------------------------------------------------------------------
constexpr int MA_WARPS = 8, MA_ROWS = 16 * MA_WARPS, MA_KB = 64, MA_KWORDS = 36;

__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(MA_WARPS * 32) small_attention_mma_kernel(const SmallAttn p) {
  extern __shared__ __align__(16) unsigned char sa_smem[];
  const int S = p.S, S64 = (S + MA_KB - 1) / MA_KB * MA_KB, VT = S64 + 8;  // VT: bf16 per V^T row
  uint32_t* Kw = reinterpret_cast<uint32_t*>(sa_smem);                       // [S64][36 words]
  __nv_bfloat16* Vt = reinterpret_cast<__nv_bfloat16*>(Kw + (size_t)S64 * MA_KWORDS);  // [64][VT]
  float* relb = reinterpret_cast<float*>(Vt + (size_t)SA_DH * VT);                      // [2S-1] (bias_rel only)
  pdl_wait();
  pdl_launch_dependents();
  const int chunks = (S + MA_ROWS - 1) / MA_ROWS;
  const int bh = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int b = bh / p.H, h = bh % p.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const size_t row_base = (size_t)b * S;
  const int kneed = p.causal ? min(S, (chunk + 1) * MA_ROWS) : S;  // keys any row of this CTA can see
  const int kstage = (kneed + MA_KB - 1) / MA_KB * MA_KB;
  for (int idx = threadIdx.x; idx < kstage * (SA_DH / 2); idx += MA_WARPS * 32) {
    const int j = idx / (SA_DH / 2), w = idx % (SA_DH / 2);
    uint32_t kw = 0u, vw = 0u;
    if (j < S) {
      kw = reinterpret_cast<const uint32_t*>(p.k + (row_base + j) * p.ldk + h * SA_DH)[w];
      vw = reinterpret_cast<const uint32_t*>(p.v + (row_base + j) * p.ldv + h * SA_DH)[w];
    }
    Kw[(size_t)j * MA_KWORDS + w] = kw;
    reinterpret_cast<uint16_t*>(Vt)[(size_t)(2 * w) * VT + j] = (uint16_t)(vw & 0xffffu);
    reinterpret_cast<uint16_t*>(Vt)[(size_t)(2 * w + 1) * VT + j] = (uint16_t)(vw >> 16);
  }
  if (p.bias != nullptr && p.bias_rel)
    for (int i = threadIdx.x; i < 2 * S - 1; i += MA_WARPS * 32) relb[i] = p.bias[(size_t)h * (2 * S - 1) + i];
  __syncthreads();
  const int row0 = chunk * MA_ROWS + warp * 16;
  if (row0 >= S) return;
  const int r0 = row0 + g, r1 = row0 + g + 8;
  uint32_t qa[4][4];
  {
    const uint32_t* q0 = reinterpret_cast<const uint32_t*>(p.q + (row_base + min(r0, S - 1)) * p.ldq + h * SA_DH);
    const uint32_t* q1 = reinterpret_cast<const uint32_t*>(p.q + (row_base + min(r1, S - 1)) * p.ldq + h * SA_DH);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      qa[ks][0] = q0[8 * ks + t];
      qa[ks][1] = q1[8 * ks + t];
      qa[ks][2] = q0[8 * ks + t + 4];
      qa[ks][3] = q1[8 * ks + t + 4];
    }
  }
  float o[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int kend = p.causal ? min(S, row0 + 16) : S;
  const uint32_t* Vw = reinterpret_cast<const uint32_t*>(Vt);
  const int vt_words = VT / 2;
  for (int kb = 0; kb < kend; kb += MA_KB) {
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const uint32_t* kr = Kw + (size_t)(kb + nt * 8 + g) * MA_KWORDS;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) mma_16816(s[nt], qa[ks], kr[8 * ks + t], kr[8 * ks + t + 4]);
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int row = e < 2 ? r0 : r1, key = kb + nt * 8 + 2 * t + (e & 1);
        float v = s[nt][e] * p.scale;
        if (p.bias != nullptr && row < S && key < S)
          v += p.bias_rel ? relb[key - row + S - 1] : p.bias[((size_t)h * S + row) * S + key];
        if (key >= S || (p.causal && key > row)) v = -INFINITY;
        s[nt][e] = v;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
    const float c0 = mn0 == -INFINITY ? 0.f : mn0, c1 = mn1 == -INFINITY ? 0.f : mn1;  // (only padding rows stay at -inf)
    const float al0 = __expf(m0 - c0), al1 = __expf(m1 - c1);
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = __expf(s[nt][0] - c0);
      s[nt][1] = __expf(s[nt][1] - c0);
      s[nt][2] = __expf(s[nt][2] - c1);
      s[nt][3] = __expf(s[nt][3] - c1);
      rs0 += s[nt][0] + s[nt][1];
      rs1 += s[nt][2] + s[nt][3];
      o[nt][0] *= al0;
      o[nt][1] *= al0;
      o[nt][2] *= al1;
      o[nt][3] *= al1;
    }
    l0 = l0 * al0 + rs0;  // lane-partial row sums: the rescale factor is the same on the four lanes of a row
    l1 = l1 * al1 + rs1;
    m0 = mn0;
    m1 = mn1;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t* vr = Vw + (size_t)(nt * 8 + g) * vt_words + (kb + 16 * kk) / 2;
        mma_16816(o[nt], a, vr[t], vr[t + 4]);
      }
    }
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.0f / l0, i1 = 1.0f / l1;
  if (r0 < S) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.out + (row_base + r0) * p.ldo + h * SA_DH);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dst[nt * 4 + t] = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
  }
  if (r1 < S) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.out + (row_base + r1) * p.ldo + h * SA_DH);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) dst[nt * 4 + t] = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
  }
}

}  // namespace
}  // namespace lx

using namespace lx;

static int g_small_attention_scalar = 0;
// development / test knob: 1 = use the scalar-FMA form of lx_attention_small instead of the mma.sync one
extern "C" void lx_debug_small_attention_scalar(int on) { g_small_attention_scalar = on; }

extern "C" int lx_embed_rows(const void* table, const int32_t* ids, const void* pos, int32_t period, void* out, int32_t n,
                             int32_t D, int32_t vocab, void* stream) {
  LX_CHECK_ARG(table && ids && out && n > 0 && D > 0 && D % 8 == 0 && vocab > 0, "lx_embed_rows: bad arguments");
  LX_CHECK_ARG(pos == nullptr || period > 0, "lx_embed_rows: position table needs period > 0");
  const long long total = (long long)n * (D / 8);
  LaunchScope scope(KC_ROW, stream, 4.0 * n * D);
  LX_CUDA(launch_pdl(embed_rows_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, tcs(stream),
                     reinterpret_cast<const __nv_bfloat16*>(table), (const int*)ids, reinterpret_cast<const __nv_bfloat16*>(pos),
                     period, reinterpret_cast<__nv_bfloat16*>(out), n, D / 8, vocab));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_norm_rows(const void* x, int64_t ldx, const float* gamma, const float* beta, void* out, int64_t ldo,
                            int32_t rows, int32_t D, float eps, int32_t rms, void* stream) {
  LX_CHECK_ARG(x && gamma && out && rows > 0 && D > 0 && D % 8 == 0 && ldx >= D && ldo >= D && ldx % 8 == 0 && ldo % 8 == 0,
               "lx_norm_rows: bad arguments");
  LX_CHECK_ARG(!(rms && beta), "lx_norm_rows: the RMS form has no bias");
  LaunchScope scope(KC_ROW, stream, 4.0 * rows * D);
  LX_CUDA(launch_pdl(norm_rows_kernel, dim3(rows), dim3(128), 0, tcs(stream), reinterpret_cast<const __nv_bfloat16*>(x),
                     (long long)ldx, gamma, beta, reinterpret_cast<__nv_bfloat16*>(out), (long long)ldo, D, eps, rms));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_mul_rows(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int32_t rows,
                           int32_t cols, void* stream) {
  LX_CHECK_ARG(a && b && out && rows > 0 && cols > 0 && cols % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0,
               "lx_mul_rows: bad arguments");
  const long long total = (long long)rows * (cols / 8);
  LaunchScope scope(KC_ROW, stream, 6.0 * rows * cols);
  LX_CUDA(launch_pdl(mul_rows_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, tcs(stream),
                     reinterpret_cast<const __nv_bfloat16*>(a), (long long)lda, reinterpret_cast<const __nv_bfloat16*>(b),
                     (long long)ldb, reinterpret_cast<__nv_bfloat16*>(out), (long long)ldo, rows, cols / 8));
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}

extern "C" int lx_attention_small(const lx_small_attn_desc_t* d, void* stream) {
  LX_CHECK_ARG(d && d->q && d->k && d->v && d->out, "lx_attention_small: null argument");
  LX_CHECK_ARG(d->head_dim == SA_DH, "lx_attention_small: head_dim=%d, only 64 is built", d->head_dim);
  LX_CHECK_ARG(d->B > 0 && d->H > 0 && d->S > 0 && d->S <= SA_MAX_S, "lx_attention_small: S=%d outside [1, %d]", d->S, SA_MAX_S);
  LX_CHECK_ARG(d->ldq % 2 == 0 && d->ldk % 2 == 0 && d->ldv % 2 == 0 && d->ldo % 2 == 0, "lx_attention_small: odd row stride");
  SmallAttn p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(d->q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(d->k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(d->v);
  p.out = reinterpret_cast<__nv_bfloat16*>(d->out);
  p.ldq = d->ldq; p.ldk = d->ldk; p.ldv = d->ldv; p.ldo = d->ldo;
  p.bias = d->bias;
  p.B = d->B; p.H = d->H; p.S = d->S; p.causal = d->causal;
  p.scale = d->scale;
  p.bias_rel = d->bias_relative != 0;
  p.qsplit = 1;
  LaunchScope scope(KC_ATTENTION, stream, 4.0 * d->B * d->H * (double)d->S * d->S * SA_DH);
  static bool attr_set = false;
  if (!attr_set) {
    LX_CUDA(cudaFuncSetAttribute(small_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    LX_CUDA(cudaFuncSetAttribute(small_attention_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (!g_small_attention_scalar) {
    const int S64 = (d->S + MA_KB - 1) / MA_KB * MA_KB;
    const size_t smem = (size_t)S64 * MA_KWORDS * 4 + (size_t)SA_DH * (S64 + 8) * sizeof(__nv_bfloat16) +
                        (size_t)(2 * d->S) * sizeof(float);
    const int chunks = (d->S + MA_ROWS - 1) / MA_ROWS;
    LX_CUDA(launch_pdl(small_attention_mma_kernel, dim3(d->B * d->H * chunks), dim3(MA_WARPS * 32), smem, tcs(stream), p));
  } else {
    // scalar form (kept as a cross-check): few (batch, head) pairs -> split the query rows so that the grid fills the SMs
    const size_t smem = 2 * (size_t)d->S * SA_KSTRIDE * sizeof(__nv_bfloat16) + (size_t)SA_WARPS * d->S * sizeof(float);
    p.qsplit = max(1, min(4, num_sms() / (d->B * d->H)));
    LX_CUDA(launch_pdl(small_attention_kernel, dim3(d->B * d->H * p.qsplit), dim3(SA_WARPS * 32), smem, tcs(stream), p));
  }
  LX_CUDA(cudaGetLastError());
  return LX_OK;
}
