// Per-plan statistics read off the resident chain state:
// the reductions users attach as updaters / constraints behind every yielded state.
//
//   region_splits_kernel  - updaters/county_splits.py:120-161 (tally_region_splits /
//                           total_reg_splits) and :58-117 (compute_county_splits: which
//                           counties hold more than one district)
//   contiguity_kernel     - constraints/contiguity.py:169-181 (contiguous): connected
//                           components of the same-district subgraph
//
// One CTA per chain, grid-stride over chains.  Both read the chain's assignment row and cut
// mask exactly as the step kernel leaves them; nothing here is on the proposal's hot path.
#pragma once
#include "recom_device.cuh"

namespace rc {

// out[chain] = { regions with a cut edge inside them   (total_reg_splits),
//                regions whose nodes lie in > 1 district (CountySplit != NOT_SPLIT),
//                sum over regions of the districts they touch ("pieces") }
// Region ids are dense 0..R-1; a negative id means "no region" (a node the reference's
// graph.nodes[..][reg_attr] lookup would hold None for is given its own id by the host).
// Shared memory: R words (edge flag | first district seen) + R * ceil(k/32) words (district sets).
__global__ void region_splits_kernel(const Params p, const int32_t* __restrict__ region, int R, int32_t* __restrict__ out) {
  extern __shared__ uint32_t sm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int N = p.g.N, E = p.g.E, k = p.cfg.n_districts;
  const int kw = (k + 31) >> 5;
  uint32_t* inner = sm;           // [R] 1 = some cut edge has both ends in the region
  uint32_t* seen = sm + R;        // [R][kw] districts present in the region
  __shared__ int acc[3];
  for (int chain = blockIdx.x; chain < p.cfg.n_chains; chain += gridDim.x) {
    const rc_label_t* assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
    const uint32_t* mask = p.st.d_cut_mask + (size_t)chain * p.st.mask_stride;
    for (int i = tid; i < R * (1 + kw); i += T) sm[i] = 0;
    if (tid < 3) acc[tid] = 0;
    __syncthreads();
    for (int w = tid; w < (E + 31) >> 5; w += T) {
      uint32_t bits = __ldcg(mask + w);
      while (bits) {
        const int e = (w << 5) + __ffs(bits) - 1;
        bits &= bits - 1;
        if (e >= E) break;
        const int2 uv = __ldg(reinterpret_cast<const int2*>(p.g.edge_uv) + e);
        const int ru = __ldg(region + uv.x), rv = __ldg(region + uv.y);
        if (ru == rv && ru >= 0 && ru < R) inner[ru] = 1u;
      }
    }
    for (int v = tid; v < N; v += T) {
      const int r = __ldg(region + v);
      if (r < 0 || r >= R) continue;
      const uint32_t d = assign[v];
      if (d >= (uint32_t)k) continue;  // a plan that never completed (RC_SEED_REMAINING): its status says so
      atomicOr(&seen[r * kw + (d >> 5)], 1u << (d & 31u));
    }
    __syncthreads();
    int a0 = 0, a1 = 0, a2 = 0;
    for (int r = tid; r < R; r += T) {
      int nd = 0;
      for (int q = 0; q < kw; ++q) nd += __popc(seen[r * kw + q]);
      a0 += (int)inner[r];
      a1 += nd > 1;
      a2 += nd;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if ((tid & 31) == 0) { atomicAdd(&acc[0], a0); atomicAdd(&acc[1], a1); atomicAdd(&acc[2], a2); }
    __syncthreads();
    if (tid < 3) out[(size_t)chain * 3 + tid] = acc[tid];
    __syncthreads();
  }
}

// out[chain][d] = number of connected pieces of district d (0: the district owns no node).
// The plan is contiguous iff no entry exceeds 1.  Lock-free union-find over the edges that are
// not cut (hook the larger root under the smaller with atomicCAS, path halving on the way),
// one pass over the edges, then every root is counted for its district.  `parent` is N words
// of shared memory when that fits, else this CTA's slice of global scratch.
__global__ void contiguity_kernel(const Params p, int* __restrict__ gscratch, int use_smem, int32_t* __restrict__ out) {
  extern __shared__ uint32_t sm[];
  const int tid = threadIdx.x, T = blockDim.x;
  const int N = p.g.N, E = p.g.E, k = p.cfg.n_districts;
  int* parent = use_smem ? reinterpret_cast<int*>(sm) : gscratch + (size_t)blockIdx.x * N;
#if RC_LABEL_BITS == 8
  __shared__ int pieces[256];
#endif
  auto find = [&](int x) {
    volatile int* par = parent;
    int px = par[x];
    while (px != x) {
      const int gp = par[px];
      if (gp != px) par[x] = gp;  // halving: only ever replaces an ancestor by a higher ancestor
      x = px;
      px = gp;
    }
    return x;
  };
  for (int chain = blockIdx.x; chain < p.cfg.n_chains; chain += gridDim.x) {
    const rc_label_t* assign = p.st.d_assignment + (size_t)chain * p.st.assign_stride;
    for (int v = tid; v < N; v += T) parent[v] = v;
#if RC_LABEL_BITS == 8
    for (int d = tid; d < 256; d += T) pieces[d] = 0;
#else
    int* pieces = out + (size_t)chain * k;  // up to 32 767 districts: count in the output row
    for (int d = tid; d < k; d += T) pieces[d] = 0;
#endif
    __syncthreads();
    for (int e = tid; e < E; e += T) {
      const int2 uv = __ldg(reinterpret_cast<const int2*>(p.g.edge_uv) + e);
      if (assign[uv.x] != assign[uv.y]) continue;
      int a = uv.x, b = uv.y;
      for (;;) {
        a = find(a);
        b = find(b);
        if (a == b) break;
        if (a < b) { const int t = a; a = b; b = t; }  // a > b: hang a under b
        if (atomicCAS(&parent[a], a, b) == a) break;
      }
    }
    __syncthreads();
    for (int v = tid; v < N; v += T)
      if (((volatile int*)parent)[v] == v && (RC_LABEL_BITS == 8 || (int)assign[v] < k)) atomicAdd(&pieces[assign[v]], 1);
    __syncthreads();
#if RC_LABEL_BITS == 8
    for (int d = tid; d < k; d += T) out[(size_t)chain * k + d] = pieces[d];
#endif
    __syncthreads();
  }
}

}  // namespace rc
