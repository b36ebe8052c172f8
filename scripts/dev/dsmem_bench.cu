// dev microbenchmark (round 2): where should the adjacency of a LARGE merged pair live while one
// lane per tree walks it (config 4: ~8k-node pairs, ~256 KB of adjacency rows, one dependent load
// per walk step)?  Compares one dependent load per step through
//   mode 0  the CTA's own shared memory                       (config 3: placement level 0)
//   mode 1  distributed shared memory of an 8-CTA cluster     (the design VERDICT r1 item 6 asks about:
//           mapa + ld.shared::cluster, the table striped over the 8 CTAs of the cluster)
//   mode 2  a per-CTA block in global memory                  (config 4 today: placement level 2)
// and reports cycles per walk step of one walker and walk steps per second over the whole GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dsmem_bench dsmem_bench.cu && ./dsmem_bench
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r;
}
__device__ __forceinline__ uint32_t ld_cluster(uint32_t a) {
  uint32_t v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}

// Table: row u = 8 successors (u32).  n_total rows; in mode 1 row u lives in CTA (u >> lsh) of the
// cluster, in modes 0 / 2 the whole table belongs to the CTA.  `walkers` warps walk, lane 0 each.
template <int MODE>
__global__ void chase(uint32_t* gtab, int lsh, int n_total, int steps, int walkers,
                      unsigned long long* cyc, uint32_t* sink) {
  extern __shared__ uint32_t sm[];
  const int n_local = 1 << lsh;
  uint32_t rank = 0;
  if (MODE == 1) rank = cg::this_cluster().block_rank();
  uint32_t* gt = gtab + (size_t)blockIdx.x * n_total * 8;
  if (MODE == 2) {
    for (int i = threadIdx.x; i < n_total * 8; i += blockDim.x) gt[i] = hash32(i * 2654435761u + blockIdx.x) % (uint32_t)n_total;
  } else {
    for (int i = threadIdx.x; i < n_local * 8; i += blockDim.x)
      sm[i] = hash32((rank * n_local * 8 + i) * 2654435761u + blockIdx.x / 8) % (uint32_t)n_total;
  }
  __syncthreads();
  if (MODE == 1) cg::this_cluster().sync();
  const int warp = threadIdx.x >> 5;
  if (warp < walkers && (threadIdx.x & 31) == 0) {
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t u = hash32(blockIdx.x * 16 + warp) % (uint32_t)n_total, r = u * 2654435761u + 12345u;
    const long long t0 = clock64();
    for (int i = 0; i < steps; ++i) {
      r = r * 1664525u + 1013904223u;
      const uint32_t slot = r >> 29;
      if (MODE == 0) u = lds32(base + (u * 8 + slot) * 4);
      else if (MODE == 1) u = ld_cluster(mapa(base + ((u & (n_local - 1)) * 8 + slot) * 4, u >> lsh));
      else u = __ldcg(gt + u * 8 + slot);
    }
    const long long t1 = clock64();
    cyc[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
    sink[blockIdx.x * 16 + warp] = u;
  }
  if (MODE == 1) cg::this_cluster().sync();  // no CTA of the cluster may exit while its rows are read
}

template <int MODE>
static void run(const char* what, int grid, int cluster, int threads, int lsh, int n_total, int steps, int walkers,
                size_t smem, uint32_t* gtab, unsigned long long* d_cyc, uint32_t* d_sink, double ghz) {
  CK(cudaFuncSetAttribute(chase<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (cluster > 8) CK(cudaFuncSetAttribute(chase<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  CK(cudaMemset(d_cyc, 0, sizeof(unsigned long long) * 16 * 4096));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; ++rep) {  // the second run is the one reported
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, chase<MODE>, gtab, lsh, n_total, steps, walkers, d_cyc, d_sink));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned long long> h(16 * grid);
  CK(cudaMemcpy(h.data(), d_cyc, sizeof(unsigned long long) * 16 * grid, cudaMemcpyDeviceToHost));
  double sum = 0; int cnt = 0; unsigned long long mx = 0;
  for (int b = 0; b < grid; ++b) for (int w = 0; w < walkers; ++w) { sum += (double)h[b * 16 + w]; ++cnt; if (h[b * 16 + w] > mx) mx = h[b * 16 + w]; }
  const double cps = sum / cnt / steps;
  // throughput from the walkers' own clocks (the table fill is outside): all walkers run side by side
  const double steps_per_s = (double)cnt * steps / ((double)mx / (ghz * 1e9));
  printf("%-58s grid %4d x %d walkers  %7.1f cycles/step (%6.1f ns)  %8.3f G steps/s  [kernel %.2f ms]\n", what, grid, walkers, cps,
         cps / ghz, steps_per_s / 1e9, ms);
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int khz = 0; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double ghz = khz / 1e6;
  const int sms = prop.multiProcessorCount;
  printf("%s, %d SMs, %.3f GHz\n", prop.name, sms, ghz);
  unsigned long long* d_cyc; uint32_t* d_sink; uint32_t* gtab;
  CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * 16 * 4096));
  CK(cudaMalloc(&d_sink, sizeof(uint32_t) * 16 * 4096));
  const size_t blk = (size_t)16384 * 8 * 4;  // 512 KB per CTA: config 4's adjacency + walker state
  CK(cudaMalloc(&gtab, blk * (size_t)(sms * 5)));
  // -- latency of one walker ------------------------------------------------------------------
  run<0>("own shared memory, 128 KB table", sms, 1, 128, 12, 4096, 200000, 1, 128 << 10, gtab, d_cyc, d_sink, ghz);
  run<1>("DSMEM, 8-CTA cluster, 1 MB table striped over 8 x 128 KB", sms / 8 * 8, 8, 128, 12, 32768, 100000, 1, 128 << 10, gtab, d_cyc, d_sink, ghz);
  run<2>("global block 512 KB per CTA, 148 CTAs (76 MB: in L2)", sms, 1, 128, 0, 16384, 50000, 1, 0, gtab, d_cyc, d_sink, ghz);
  run<2>("global block 512 KB per CTA, 592 CTAs (303 MB: beyond L2)", sms * 4, 1, 128, 0, 16384, 20000, 1, 0, gtab, d_cyc, d_sink, ghz);
  // -- throughput with the walks a design keeps in flight ---------------------------------------
  // cluster design: one CTA per SM owns all its shared memory; a 1.6 MB cluster holds ~4 pairs, each
  // with W = 4 useful speculative trees -> 16 walks per 8 SMs (2 per SM); 4 per SM shown as the upper end
  run<1>("DSMEM, 8-CTA cluster, 2 walkers per SM", sms / 8 * 8, 8, 128, 12, 32768, 100000, 2, 128 << 10, gtab, d_cyc, d_sink, ghz);
  run<1>("DSMEM, 8-CTA cluster, 4 walkers per SM", sms / 8 * 8, 8, 128, 12, 32768, 100000, 4, 128 << 10, gtab, d_cyc, d_sink, ghz);
  run<1>("DSMEM, 8-CTA cluster, 16 walkers per SM", sms / 8 * 8, 8, 512, 12, 32768, 50000, 16, 128 << 10, gtab, d_cyc, d_sink, ghz);
  // global design (what the engine runs at config 4): 4 CTAs per SM x 4 walkers, 512 KB each
  run<2>("global, 4 CTAs/SM x 4 walkers, 303 MB (config 4 today)", sms * 4, 1, 128, 0, 16384, 20000, 4, 0, gtab, d_cyc, d_sink, ghz);
  run<2>("global, 1 CTA/SM x 4 walkers, 76 MB (blocks kept in L2)", sms, 1, 128, 0, 16384, 50000, 4, 0, gtab, d_cyc, d_sink, ghz);
  return 0;
}
