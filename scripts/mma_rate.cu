// Micro-benchmark (development aid, not product code): cycles per tcgen05.mma for the operand forms the attention and
// GEMM kernels use.  One CTA per SM, one issuing thread, `reps` back-to-back MMAs + commit; prints clk / MMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/bin/mma_rate scripts/mma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../loongx_b200/csrc/ptx.cuh"

using namespace lx;

// mode 0: SS  M128 N128  A K-major, B K-major      (QK^T)
// mode 1: TS  M128 N128  A TMEM,    B MN-major     (PV)
// mode 2: SS  M128 N256  A K-major, B K-major      (GEMM)
// mode 3: SS  M128 N128  A K-major, B MN-major
// mode 4: 8 x mode 1 then 8 x mode 0, alternating   (attention issue pattern, different accumulators)
// mode 5: TS  M128 N128  A TMEM,    B K-major
// mode 6: SS  M128 N64   K-major both
template <int mode>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    constexpr uint32_t id_kk128 = make_idesc_bf16(128, 128, false, false);
    constexpr uint32_t id_kmn128 = make_idesc_bf16(128, 128, false, true);
    constexpr uint32_t id_kk256 = make_idesc_bf16(128, 256, false, false);
    constexpr uint32_t id_kk64 = make_idesc_bf16(128, 64, false, false);
    long long t0 = clock64();
    for (int i0 = 0; i0 < reps; i0 += 16) {
#pragma unroll
     for (int u = 0; u < 16; ++u) {
      const int i = u;
      const int kk = i & 7;
      const uint32_t off = (kk >> 2) * 16384 + (kk & 3) * 32;
      switch (mode) {
        case 0:
          umma_ss(tm, make_sdesc_sw128(a + off, 16, 1024), make_sdesc_sw128(b + off, 16, 1024), id_kk128, 1);
          break;
        case 1:
          umma_ts(tm + 256, tm + kk * 8, make_sdesc_sw128(b + kk * 2048, 16384, 1024), id_kmn128, 1);
          break;
        case 2:
          umma_ss(tm, make_sdesc_sw128(a + off, 16, 1024), make_sdesc_sw128(b + (kk >> 2) * 32768 + (kk & 3) * 32, 16, 1024),
                  id_kk256, 1);
          break;
        case 3:
          umma_ss(tm, make_sdesc_sw128(a + off, 16, 1024), make_sdesc_sw128(b + kk * 2048, 16384, 1024), id_kmn128, 1);
          break;
        case 4:
          if ((i >> 3) & 1)
            umma_ss(tm, make_sdesc_sw128(a + off, 16, 1024), make_sdesc_sw128(b + off, 16, 1024), id_kk128, 1);
          else
            umma_ts(tm + 256, tm + 128 + kk * 8, make_sdesc_sw128(b + 32768 + kk * 2048, 16384, 1024), id_kmn128, 1);
          break;
        case 5:
          umma_ts(tm + 256, tm + kk * 8, make_sdesc_sw128(b + off, 16, 1024), id_kk128, 1);
          break;
        case 6:
          umma_ss(tm, make_sdesc_sw128(a + off, 16, 1024), make_sdesc_sw128(b + off, 16, 1024), id_kk64, 1);
          break;
      }
     }
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tm, 512);
  }
}

template <int mode>
static void launch1(int ctas, int reps, long long* d) {
  cudaFuncSetAttribute(mma_rate_kernel<mode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 162 * 1024);
  mma_rate_kernel<mode><<<ctas, 128, 162 * 1024>>>(reps, d);
}
static void launch(int mode, int ctas, int reps, long long* d) {
  switch (mode) {
    case 0: launch1<0>(ctas, reps, d); break;
    case 1: launch1<1>(ctas, reps, d); break;
    case 2: launch1<2>(ctas, reps, d); break;
    case 3: launch1<3>(ctas, reps, d); break;
    case 4: launch1<4>(ctas, reps, d); break;
    case 5: launch1<5>(ctas, reps, d); break;
    case 6: launch1<6>(ctas, reps, d); break;
  }
}

int main(int argc, char** argv) {
  int reps = argc > 1 ? atoi(argv[1]) : 4096;
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  const char* names[] = {"SS N128 K/K", "TS N128 B=MN", "SS N256 K/K", "SS N128 B=MN", "attn 8xTS+8xSS", "TS N128 B=K", "SS N64 K/K"};
  for (int ctas : {1, 148}) {
    for (int mode = 0; mode < 7; ++mode) {
      launch(mode, ctas, 64, d);  // warm-up
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      cudaEventRecord(e0);
      launch(mode, ctas, reps, d);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        return 1;
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(ctas * 2);
      cudaMemcpy(h.data(), d, ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost);
      double issue = 0, total = 0;
      for (int i = 0; i < ctas; ++i) {
        issue += h[2 * i];
        total += h[2 * i + 1];
      }
      const int n = mode == 2 ? 256 : (mode == 6 ? 64 : 128);
      const double flop = 2.0 * 128 * n * 16 * reps * ctas;
      printf("ctas %3d  %-16s issue %.1f clk/MMA   complete %.1f clk/MMA   %.3f ms  %.1f TFLOP/s  (%.0f MHz eff)\n", ctas,
             names[mode], issue / ctas / reps, total / ctas / reps, ms, flop / ms / 1e9, total / ctas / ms / 1e3);
    }
  }
  return 0;
}
