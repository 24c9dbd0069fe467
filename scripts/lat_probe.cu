// Dependent-load latency on B200 as a function of the footprint (TLB reach / L2 residency): one lane per warp chases
// a[i] = (i + stride) mod n over n 8-byte elements; `warps` warps run at once.  Diagnostic only (gpurun), not product.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void fill(unsigned long long *a, unsigned long long n, unsigned long long s) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) a[i] = (i + s) & (n - 1);
}
__global__ void chase(const unsigned long long *a, unsigned long long n, int iters, unsigned long long *out, long long *cyc) {
  if (threadIdx.x & 31) return;
  unsigned long long w = (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x) >> 5;
  unsigned long long i = (w * 0x9E3779B97F4A7C15ull) & (n - 1);
  for (int k = 0; k < 64; k++) i = a[i];
  long long t0 = clock64();
  for (int k = 0; k < iters; k++) i = a[i];
  long long t1 = clock64();
  out[w] = i; cyc[w] = t1 - t0;
}
int main(int argc, char **argv) {
  const int iters = 4000;
  for (int lg = 23; lg <= 32; lg += 3) {              // 64 MB, 512 MB, 4 GB, 32 GB
    unsigned long long n = 1ull << lg, *a, *out; long long *cyc;
    if (cudaMalloc(&a, n * 8) != cudaSuccess) { printf("alloc failed at 2^%d\n", lg); break; }
    unsigned long long s = ((unsigned long long)(n * 0.6180339887) | 1ull);
    fill<<<1184, 256>>>(a, n, s);
    for (int warps : {1, 148 * 4, 148 * 16}) {
      cudaMalloc(&out, warps * 8); cudaMalloc(&cyc, warps * 8);
      chase<<<(warps + 3) / 4, 128>>>(a, n, iters, out, cyc);
      cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, cyc, sizeof(long long) * (warps < 4 ? warps : 4), cudaMemcpyDeviceToHost);
      printf("footprint %6.0f MB  warps %5d  cycles/load %7.1f\n", n * 8 / 1048576.0, warps, (double)h[0] / iters);
      cudaFree(out); cudaFree(cyc);
    }
    cudaFree(a);
  }
  return 0;
}
