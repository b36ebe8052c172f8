// dev microbenchmark (round 2): what does one step of a lone random walk through shared memory cost on B200,
// and which way of writing it gets there?  nvcc -arch=sm_100a -O3 -o walk_bench walk_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t lds16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "h"((uint16_t)v) : "memory"); }

// layout: nbr[n][8] u32 (neighbour | degree << 16), S[n] u16, ring[64] u32
__global__ void walk(int n, int steps, int mode, int lanes, long long* out) {
  extern __shared__ uint32_t sm[];
  uint32_t* nbr = sm;
  uint16_t* S = (uint16_t*)(sm + n * 8);
  uint32_t* ring = sm + n * 8 + (n + 1) / 2 + 1;
  // a 6-regular-ish circulant graph
  for (int u = threadIdx.x; u < n; u += blockDim.x) {
    const int off[6] = {1, n - 1, 17, n - 17, 101, n - 101};
    for (int k = 0; k < 6; ++k) nbr[u * 8 + k] = (uint32_t)((u + off[k]) % n) | (6u << 16);
    S[u] = 0;
  }
  for (int i = threadIdx.x; i < 64; i += blockDim.x) ring[i] = 0x9E3779B9u * (i + 1);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  if (threadIdx.x < 32) {
    const bool active = lane < lanes;
    long long t0 = clock64();
    uint32_t acc = 0;
    if (mode == 0) {            // plain C++ through generic pointers, 1 load chain
      if (active) { int u = lane; for (int i = 0; i < steps; ++i) u = nbr[u * 8] & 0xffff; acc = u; }
    } else if (mode == 1) {     // full step, plain C++ (what the compiler makes of it)
      if (active) {
        int u = lane; uint32_t t = 0; uint32_t e = nbr[u * 8];
        for (int i = 0; i < steps; ++i) {
          const int v = e & 0xffff; ++t;
          S[u] = (uint16_t)v;
          const uint32_t wv = S[v];
          const uint32_t en = nbr[v * 8 + __umulhi(ring[t & 63], e >> 16)];
          if (wv & 0x8000u) break;
          u = v; e = en;
        }
        acc = u;
      }
    } else if (mode == 2) {     // full step, explicit shared addresses
      if (active) {
        const uint32_t a_nbr = (uint32_t)__cvta_generic_to_shared(nbr), a_S = (uint32_t)__cvta_generic_to_shared(S),
                       a_ring = (uint32_t)__cvta_generic_to_shared(ring);
        uint32_t u = lane, t = 0; uint32_t e = lds32(a_nbr + u * 32);
        uint32_t r = lds32(a_ring + 4);
        for (int i = 0; i < steps; ++i) {
          const uint32_t v = e & 0xffffu; ++t;
          sts16(a_S + 2 * u, v);
          const uint32_t en = lds32(a_nbr + v * 32 + 4 * __umulhi(r, e >> 16));
          const uint32_t wv = lds16(a_S + 2 * v);
          r = lds32(a_ring + 4 * ((t + 1) & 63));
          if (wv & 0x8000u) break;
          u = v; e = en;
        }
        acc = u;
      }
    } else if (mode == 3) {     // mode 2 inside an all-lanes loop with a vote per 8 steps (the kernel's shape)
      const uint32_t a_nbr = (uint32_t)__cvta_generic_to_shared(nbr), a_S = (uint32_t)__cvta_generic_to_shared(S),
                     a_ring = (uint32_t)__cvta_generic_to_shared(ring);
      uint32_t u = lane, t = 0; uint32_t e = lds32(a_nbr + u * 32);
      uint32_t r = lds32(a_ring + 4);
      int left = active ? steps : 0;
      while (__any_sync(0xffffffffu, left > 0)) {
        if (left > 0) {
#pragma unroll 1
          for (int k = 0; k < 8; ++k) {
            const uint32_t v = e & 0xffffu; ++t;
            sts16(a_S + 2 * u, v);
            const uint32_t en = lds32(a_nbr + v * 32 + 4 * __umulhi(r, e >> 16));
            const uint32_t wv = lds16(a_S + 2 * v);
            r = lds32(a_ring + 4 * ((t + 1) & 63));
            if (wv & 0x8000u) { left = 0; break; }
            u = v; e = en;
          }
          left -= 8;
        }
      }
      acc = u;
    } else if (mode == 4) {     // dependent LDS only, explicit
      if (active) { const uint32_t a_nbr = (uint32_t)__cvta_generic_to_shared(nbr); uint32_t u = lane;
        for (int i = 0; i < steps; ++i) u = lds32(a_nbr + u * 32) & 0xffffu; acc = u; }
    } else if (mode == 5) {     // LDS chain + a special-register read per step
      if (active) { const uint32_t a_nbr = (uint32_t)__cvta_generic_to_shared(nbr); uint32_t u = lane;
        for (int i = 0; i < steps; ++i) { uint32_t s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); u = lds32(a_nbr + (u + (s >> 20)) * 32) & 0xffffu; } acc = u; }
    }
    long long t1 = clock64();
    if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; }
    out[4096 + threadIdx.x] = acc;
  }
  __syncthreads();
}
int main() {
  long long* d; cudaMalloc(&d, 8192 * 16);
  long long h[2];
  const int n = 1009, steps = 20000;
  const size_t smem = (size_t)n * 32 + n * 2 + 64 * 4 + 64;
  cudaFuncSetAttribute(walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int mode = 0; mode < 6; ++mode)
    for (int lanes : {1, 4, 32})
      for (int threads : {32, 256}) {
        for (int grid : {1, 148 * 4}) {
          walk<<<grid, threads, smem>>>(n, steps, mode, lanes, d);
          cudaDeviceSynchronize();
          cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("mode %d lanes %2d threads %3d grid %4d: %.1f cycles/step (%s)\n", mode, lanes, threads, grid, (double)h[0] / steps,
                 cudaGetErrorString(cudaGetLastError()));
        }
      }
  return 0;
}
