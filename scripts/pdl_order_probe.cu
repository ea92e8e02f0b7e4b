// Does a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization stay ordered after a cudaMemsetAsync /
// cudaMemcpyAsync that sits between it and the previous kernel (which triggers launch_dependents early)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o loongx_b200/lib/pdl_order_probe scripts/pdl_order_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void fill(int* p, int v, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
// triggers its dependents at once, then keeps the SMs busy for a while
__global__ void primary(float* sink, int iters) {
  asm volatile("griddepcontrol.launch_dependents;");
  float x = threadIdx.x;
  for (int i = 0; i < iters; ++i) x = x * 1.0001f + 0.5f;
  if (x == 123.f) sink[0] = x;
}
__global__ void check_copy(const int* dst, int expect, int* stale, int n) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && dst[i] != expect) atomicAdd(stale, 1);
}
__global__ void accumulate(int* acc, int n) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicAdd(&acc[i & 63], 1);
}
__global__ void check_sum(const int* acc, int expect, int* bad) {
  int s = 0;
  for (int i = 0; i < 64; ++i) s += acc[i];
  if (s != expect) atomicAdd(bad, 1);
}

// case 3: the primary accumulates with fire-and-forget reductions (RED); the dependent reads the sums after its wait,
// through ordinary loads or through the non-coherent path (const __restrict__ / __ldg)
__global__ void fill_d(double* p, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0.0;
}
__global__ void red_primary(double* acc, float* sink, int spin) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  float x = threadIdx.x;
  for (int i = 0; i < spin * (1 + (blockIdx.x & 3)); ++i) x = x * 1.0001f + 0.5f;
  if (x == 123.f) sink[0] = x;
  atomicAdd(&acc[threadIdx.x & 63], 1.0);
}
__global__ void red_check(const double* acc, double expect, int* bad) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x < 64 && acc[threadIdx.x] != expect) atomicAdd(bad, 1);
}
__global__ void red_check_nc(const double* __restrict__ acc, double expect, int* bad) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x < 64 && __ldg(acc + threadIdx.x) != expect) atomicAdd(bad, 1);
}

template <typename... KArgs, typename... Args>
void launch(bool pdl, void (*k)(KArgs...), int grid, int block, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  CK(cudaLaunchKernelEx(&cfg, k, static_cast<KArgs>(args)...));
}

int main() {
  cudaStream_t st; CK(cudaStreamCreate(&st));
  const int iters = 300, spin = 20000;
  float* sink; CK(cudaMalloc(&sink, 4));
  int *stale, *bad, *acc, *src, *dst;
  CK(cudaMalloc(&stale, 4)); CK(cudaMalloc(&bad, 4)); CK(cudaMalloc(&acc, 256));
  for (int n : {1 << 10, 1 << 20}) {
    CK(cudaMalloc(&src, 4 * n)); CK(cudaMalloc(&dst, 4 * n));
    for (int pdl = 0; pdl < 2; ++pdl) {
      CK(cudaMemset(stale, 0, 4)); CK(cudaMemset(bad, 0, 4)); CK(cudaMemset(dst, 0xff, 4 * n));
      for (int it = 0; it < iters; ++it) {
        fill<<<(n + 255) / 256, 256, 0, st>>>(src, it, n);
        launch(true, primary, 148, 256, st, sink, spin);
        CK(cudaMemcpyAsync(dst, src, 4 * n, cudaMemcpyDeviceToDevice, st));
        launch(pdl != 0, check_copy, (n + 255) / 256, 256, st, (const int*)dst, it, stale, n);
      }
      CK(cudaStreamSynchronize(st));
      int h = 0; CK(cudaMemcpy(&h, stale, 4, cudaMemcpyDeviceToHost));
      printf("[pdl probe] memcpy %8d B, consumer %s: stale elements seen over %d iterations: %d\n", 4 * n, pdl ? "PDL" : "ordered", iters, h);
      for (int it = 0; it < iters; ++it) {
        launch(true, primary, 148, 256, st, sink, spin);
        CK(cudaMemsetAsync(acc, 0, 256, st));
        launch(pdl != 0, accumulate, (n + 255) / 256, 256, st, acc, n);
        check_sum<<<1, 1, 0, st>>>((const int*)acc, n, bad);
      }
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
      printf("[pdl probe] memset 256 B then %d atomic adds, consumer %s: iterations with a wrong sum: %d of %d\n", n, pdl ? "PDL" : "ordered", h, iters);
    }
    CK(cudaFree(src)); CK(cudaFree(dst));
  }
  double* dacc; CK(cudaMalloc(&dacc, 64 * 8));
  for (int nc = 0; nc < 2; ++nc)
    for (int pdl = 0; pdl < 2; ++pdl) {
      CK(cudaMemset(bad, 0, 4));
      const int grid = 512;
      for (int it = 0; it < 1000; ++it) {
        fill_d<<<1, 64, 0, st>>>(dacc, 64);
        launch(true, red_primary, grid, 256, st, dacc, sink, 200);
        if (nc) launch(pdl != 0, red_check_nc, 8, 64, st, (const double*)dacc, (double)(grid * 4), bad);
        else launch(pdl != 0, red_check, 8, 64, st, (const double*)dacc, (double)(grid * 4), bad);
      }
      CK(cudaStreamSynchronize(st));
      int h = 0; CK(cudaMemcpy(&h, bad, 4, cudaMemcpyDeviceToHost));
      printf("[pdl probe] RED.F64 sums by the primary, read by a %s dependent through %s loads: wrong values seen: %d (1000 iterations x 512 slots-reads)\n",
             pdl ? "PDL" : "stream-ordered", nc ? "non-coherent (__ldg)" : "ordinary", h);
    }
  return 0;
}
