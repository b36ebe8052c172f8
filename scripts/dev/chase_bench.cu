// dev microbenchmark: cycles per step of a lone thread chasing pointers through shared memory
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void chase(int n, int steps, int mode, long long* out) {
  extern __shared__ unsigned long long w[];
  // a single random-ish cycle: u -> (u * 7 + 3) % n  (n prime-ish)
  for (int u = threadIdx.x; u < n; u += blockDim.x) w[u] = (unsigned long long)((u * 7 + 3) % n) | (1ull << 40);
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    int u = 0;
    unsigned acc = 0;
    if (mode == 0) {            // load only
      for (int i = 0; i < steps; ++i) { u = (int)(w[u] & 0xffff); }
    } else if (mode == 1) {     // load + stamp store (low half) + compare
      uint32_t* lo = reinterpret_cast<uint32_t*>(w);
      unsigned long long x = w[u];
      for (int i = 0; i < steps; ++i) {
        uint32_t l = (uint32_t)x;
        if ((l >> 16) == 0xffffu) break;
        int un = l & 0xffff;
        x = w[un];
        lo[2 * u] = (l & 0xffff) | (5u << 16);
        u = un;
      }
    } else {                    // 32-bit loads only
      uint32_t* lo = reinterpret_cast<uint32_t*>(w);
      for (int i = 0; i < steps; ++i) { u = (int)(lo[2 * u] & 0xffff); }
    }
    long long t1 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = u + acc;
  }
  __syncthreads();
}
int main() {
  long long* d; cudaMalloc(&d, 4096 * 16);
  long long h[2];
  const int n = 1009, steps = 100000;
  for (int mode = 0; mode < 3; ++mode)
    for (int grid : {1, 148, 148 * 8}) {
      chase<<<grid, 128, n * 8>>>(n, steps, mode, d);
      cudaDeviceSynchronize();
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      printf("mode %d grid %4d: %.1f cycles/step (%s)\n", mode, grid, (double)h[0] / steps, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
