// dev microbenchmark (round 2): one 1-D bulk copy (cp.async.bulk + mbarrier) of a 10 KB row into shared memory
// vs two 128-bit loads per thread; persistent loop like the chain-step kernel's gather.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(const unsigned char* rows, int stride, int n_rows, int iters, int mode, unsigned long long* out) {
  extern __shared__ __align__(16) unsigned char sm[];
  unsigned long long* mb = (unsigned long long*)sm;      // mbarrier at offset 0
  uint4* stage = (uint4*)(sm + 16);
  const int tid = threadIdx.x;
  const uint32_t bar = (uint32_t)__cvta_generic_to_shared(mb), dst = (uint32_t)__cvta_generic_to_shared(stage);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  uint32_t phase = 0, acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const unsigned char* row = rows + (size_t)((blockIdx.x * 131 + it * 17) % n_rows) * stride;
    if (mode == 1) {
      __syncthreads();
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)stride) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(row), "r"((uint32_t)stride), "r"(bar) : "memory");
      }
      uint32_t ok;
      do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(phase) : "memory");
      } while (!ok);
      phase ^= 1;
      for (int w = tid; w * 32 < stride; w += blockDim.x) { uint4 a = stage[2 * w], b = stage[2 * w + 1]; acc += a.x ^ a.w ^ b.y ^ b.z; }
    } else {
      const uint4* r4 = (const uint4*)row;
      for (int w = tid; w * 32 < stride; w += blockDim.x) { uint4 a = __ldcg(r4 + 2 * w), b = __ldcg(r4 + 2 * w + 1); acc += a.x ^ a.w ^ b.y ^ b.z; }
    }
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 0x12345678u) out[4096] = acc;
}
int main() {
  const int stride = 10016, n_rows = 1024, iters = 2000;
  unsigned char* rows; cudaMalloc(&rows, (size_t)stride * n_rows); cudaMemset(rows, 3, (size_t)stride * n_rows);
  unsigned long long* out; cudaMalloc(&out, 8192 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 + stride);
  for (int mode = 0; mode < 2; ++mode)
    for (int grid : {148, 444}) {
      k<<<grid, 320, 16 + stride>>>(rows, stride, n_rows, iters, mode, out);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
      printf("mode %d (%s) grid %d: %.0f cycles per row (%s)\n", mode, mode ? "bulk copy" : "2 x LDG.128 per thread", grid, (double)h / iters, cudaGetErrorString(e));
    }
  return 0;
}
