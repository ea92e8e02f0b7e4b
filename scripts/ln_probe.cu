// Micro-benchmark (development aid): why does the AdaLN row kernel take ~15 us for 31.5 MB?  Variants of the thread
// mapping on a [2560, 3072] bf16 activation, timed back to back with CUDA events (L2-resident and DRAM-streaming inputs).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ float2 unpack(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ uint32_t pack(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float wsum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// E: pure copy, warp per row
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_copy(const uint4* x, uint4* out, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * WARPS + warp;
  if (row >= rows) return;
  uint4 raw[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) raw[i] = x[(size_t)row * 384 + i * 32 + lane];
#pragma unroll
  for (int i = 0; i < 12; ++i) out[(size_t)row * 384 + i * 32 + lane] = raw[i];
}

// A: warp per row LN + modulate (ROWS_PER_WARP rows processed one after the other by the same warp)
template <int WARPS, int RPW, bool MOD>
__global__ void __launch_bounds__(WARPS * 32) k_ln(const uint4* x, uint4* out, const uint4* scale, const uint4* shift, int rows) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = 0; rr < RPW; ++rr) {
    const int row = (blockIdx.x * WARPS + warp) * RPW + rr;
    if (row >= rows) return;
    uint4 raw[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) raw[i] = x[(size_t)row * 384 + i * 32 + lane];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      float2 a = unpack(raw[i].x), b = unpack(raw[i].y), c = unpack(raw[i].z), e = unpack(raw[i].w);
      sum += ((a.x + a.y) + (b.x + b.y)) + ((c.x + c.y) + (e.x + e.y));
    }
    const float mean = wsum(sum) / 3072.f;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f = unpack(u[e]);
        float c0 = f.x - mean, c1 = f.y - mean;
        sq += c0 * c0 + c1 * c1;
      }
    }
    const float rstd = rsqrtf(wsum(sq) / 3072.f + 1e-6f);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const uint32_t u[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
      uint32_t o[4];
      if (MOD) {
        const uint4 s4 = __ldg(scale + i * 32 + lane), h4 = __ldg(shift + i * 32 + lane);
        const uint32_t su[4] = {s4.x, s4.y, s4.z, s4.w}, hu[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = unpack(u[e]), sc = unpack(su[e]), sh = unpack(hu[e]);
          o[e] = pack((f.x - mean) * rstd * (1.f + sc.x) + sh.x, (f.y - mean) * rstd * (1.f + sc.y) + sh.y);
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float2 f = unpack(u[e]);
          o[e] = pack((f.x - mean) * rstd, (f.y - mean) * rstd);
        }
      }
      out[(size_t)row * 384 + i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}

// D: CTA (384 threads) per row, one 16-byte chunk per thread, block reduction through shared memory
__global__ void __launch_bounds__(384) k_ln_cta(const uint4* x, uint4* out, const uint4* scale, const uint4* shift, int rows) {
  __shared__ float red[2][12];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x;
  const uint4 raw = x[(size_t)row * 384 + threadIdx.x];
  const uint4 s4 = __ldg(scale + threadIdx.x), h4 = __ldg(shift + threadIdx.x);
  const uint32_t u[4] = {raw.x, raw.y, raw.z, raw.w};
  float sum = 0.f;
  for (int e = 0; e < 4; ++e) { float2 f = unpack(u[e]); sum += f.x + f.y; }
  sum = wsum(sum);
  if (lane == 0) red[0][warp] = sum;
  __syncthreads();
  float tot = 0.f;
  for (int w = 0; w < 12; ++w) tot += red[0][w];
  const float mean = tot / 3072.f;
  float sq = 0.f;
  for (int e = 0; e < 4; ++e) { float2 f = unpack(u[e]); float a = f.x - mean, b = f.y - mean; sq += a * a + b * b; }
  sq = wsum(sq);
  if (lane == 0) red[1][warp] = sq;
  __syncthreads();
  tot = 0.f;
  for (int w = 0; w < 12; ++w) tot += red[1][w];
  const float rstd = rsqrtf(tot / 3072.f + 1e-6f);
  const uint32_t su[4] = {s4.x, s4.y, s4.z, s4.w}, hu[4] = {h4.x, h4.y, h4.z, h4.w};
  uint32_t o[4];
  for (int e = 0; e < 4; ++e) {
    float2 f = unpack(u[e]), sc = unpack(su[e]), sh = unpack(hu[e]);
    o[e] = pack((f.x - mean) * rstd * (1.f + sc.x) + sh.x, (f.y - mean) * rstd * (1.f + sc.y) + sh.y);
  }
  out[(size_t)row * 384 + threadIdx.x] = make_uint4(o[0], o[1], o[2], o[3]);
}

template <typename F>
float timeit(F f, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) f(i);
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) f(i);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms * 1e3f / iters;
}

int main() {
  const int rows = 2560, NB = 16;  // NB buffers of 15.7 MB each: rotating through them defeats the 126 MB L2
  const size_t row_u4 = 384, bytes = (size_t)rows * row_u4 * 16;
  uint4 *x, *out, *sc, *sh;
  cudaMalloc(&x, bytes * NB);
  cudaMalloc(&out, bytes * NB);
  cudaMalloc(&sc, 6144 * 4);
  cudaMalloc(&sh, 6144 * 4);
  cudaMemset(x, 0x11, bytes * NB);
  cudaMemset(sc, 0, 6144 * 4);
  cudaMemset(sh, 0, 6144 * 4);
  for (int rot = 0; rot < 2; ++rot) {
    auto X = [&](int i) { return x + (rot ? (size_t)(i % NB) * rows * row_u4 : 0); };
    auto O = [&](int i) { return out + (rot ? (size_t)(i % NB) * rows * row_u4 : 0); };
    printf("---- %s\n", rot ? "rotating over 16 buffers (DRAM)" : "one buffer (L2-resident)");
    printf("copy   4 warps/CTA          : %6.2f us\n", timeit([&](int i) { k_copy<4><<<rows / 4, 128>>>(X(i), O(i), rows); }, 200));
    printf("copy   8 warps/CTA          : %6.2f us\n", timeit([&](int i) { k_copy<8><<<rows / 8, 256>>>(X(i), O(i), rows); }, 200));
    printf("ln+mod 4 warps/CTA, 1 row/w : %6.2f us\n", timeit([&](int i) { k_ln<4, 1, true><<<rows / 4, 128>>>(X(i), O(i), sc, sh, rows); }, 200));
    printf("ln+mod 8 warps/CTA, 1 row/w : %6.2f us\n", timeit([&](int i) { k_ln<8, 1, true><<<rows / 8, 256>>>(X(i), O(i), sc, sh, rows); }, 200));
    printf("ln+mod 2 warps/CTA, 1 row/w : %6.2f us\n", timeit([&](int i) { k_ln<2, 1, true><<<rows / 2, 64>>>(X(i), O(i), sc, sh, rows); }, 200));
    printf("ln     4 warps/CTA (no mod) : %6.2f us\n", timeit([&](int i) { k_ln<4, 1, false><<<rows / 4, 128>>>(X(i), O(i), sc, sh, rows); }, 200));
    printf("ln+mod 4 warps/CTA, 2 rows/w: %6.2f us\n", timeit([&](int i) { k_ln<4, 2, true><<<rows / 8, 128>>>(X(i), O(i), sc, sh, rows); }, 200));
    printf("ln+mod CTA(384)/row         : %6.2f us\n", timeit([&](int i) { k_ln_cta<<<rows, 384>>>(X(i), O(i), sc, sh, rows); }, 200));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
