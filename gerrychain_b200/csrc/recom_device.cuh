// recom_device.cuh - device-side building blocks of the batched ReCom engine (sm_100a).
//
// One CTA owns one chain-step.  Everything between "gather the merged pair" and "commit"
// lives in shared memory; HBM holds only the chain state and the read-only CSR.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/recom_b200.h"
#include "../../include/recom_rng.h"

namespace rc {

typedef uint16_t lid_t;  // local node / arc / edge-slot id inside one merged pair
constexpr int kMaxLocal = 32760;  // 2*(n-1) arcs must fit 16 bits

struct DevGraph {
  int32_t N, E, R;
  const int32_t* row_ptr;
  const int32_t* col_idx;
  const int32_t* slot_edge;
  const int32_t* edge_uv;
  const int32_t* pop;
  const int32_t* region_ids;  // [R][N]
  double surcharge[8];
};

struct Params {
  DevGraph g;
  RcConfig cfg;
  RcState st;
  RcReplay rp;        // rp.n_steps > 0: replay mode (device pointers inside)
  int64_t n_steps;    // free-running: accepted steps per chain for this launch
  int32_t cap_nodes, cap_edges;
  // per-CTA global scratch (L2 resident)
  uint16_t* g2l;      // [grid][N16]
  int32_t* nodes;     // [grid][cap_nodes]
  int32_t* eids;      // [grid][cap_edges]
  int64_t g2l_stride;
  uint32_t* badpairs; // [grid][2048] bitset of (a*256+b), pair reselection only
  int32_t* work_counter;  // persistent scheduling ticket
  int32_t keep_counters;  // init_state_kernel: leave step/proposal numbers and counters alone
};

// ---- shared-memory carve-up -------------------------------------------------------
struct Smem {
  // whole step
  int32_t* pop;      // [CN]
  lid_t* eu;         // [CE]
  lid_t* ev;         // [CE]
  lid_t* te;         // [CN] local edge index of tree edge t (ascending)
  // spanning-tree phase
  uint32_t* khi;     // [CE] high word of the 64-bit weight key
  lid_t* comp;       // [CN]
  lid_t* par;        // [CN]
  uint32_t* best_hi; // [CN]
  unsigned long long* best_lo;  // [CN] (low word << 32 | local edge)
  uint32_t* tflag;   // [CE/32]
  // Euler-tour phase (aliases the spanning-tree phase)
  int32_t* toff;     // [CN+1]
  int32_t* cur;      // [CN]
  lid_t* tadj;       // [2CN] arcs leaving each node
  lid_t* apos;       // [2CN] slot of arc a in tadj; later: tour position of arc a
  uint32_t* lrA;     // [2CN] (next << 16 | dist)
  uint32_t* lrB;     // [2CN] ; later the prefix-sum array
  lid_t* ent;        // [CN] tour position of the arc entering the node
  // small
  int32_t* misc;     // [128]
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t smem_layout(Smem* s, unsigned char* base, int CN, int CE) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align16(o + bytes); return r; };
  size_t o_misc = take(128 * 4);
  size_t o_pop = take(size_t(CN) * 4);
  size_t o_eu = take(size_t(CE) * 2);
  size_t o_ev = take(size_t(CE) * 2);
  size_t o_te = take(size_t(CN) * 2);
  size_t phase = o;
  // spanning-tree phase
  size_t o_khi = take(size_t(CE) * 4);
  size_t o_comp = take(size_t(CN) * 2);
  size_t o_par = take(size_t(CN) * 2);
  size_t o_bhi = take(size_t(CN) * 4);
  size_t o_blo = take(size_t(CN) * 8);
  size_t o_tfl = take(size_t((CE + 31) / 32) * 4);
  size_t end_a = o;
  // Euler phase; tflag must survive until te[] is built, so the Euler arrays that are
  // written before that point (none) may overlap it freely: te is built first.
  o = phase;
  size_t o_toff = take(size_t(CN + 1) * 4);
  size_t o_cur = take(size_t(CN) * 4);
  size_t o_tadj = take(size_t(2 * CN) * 2);
  size_t o_apos = take(size_t(2 * CN) * 2);
  size_t o_lrA = take(size_t(2 * CN) * 4);
  size_t o_lrB = take(size_t(2 * CN) * 4);
  size_t o_ent = take(size_t(CN) * 2);
  size_t end_b = o;
  if (s) {
    s->misc = (int32_t*)(base + o_misc);
    s->pop = (int32_t*)(base + o_pop);
    s->eu = (lid_t*)(base + o_eu);
    s->ev = (lid_t*)(base + o_ev);
    s->te = (lid_t*)(base + o_te);
    s->khi = (uint32_t*)(base + o_khi);
    s->comp = (lid_t*)(base + o_comp);
    s->par = (lid_t*)(base + o_par);
    s->best_hi = (uint32_t*)(base + o_bhi);
    s->best_lo = (unsigned long long*)(base + o_blo);
    s->tflag = (uint32_t*)(base + o_tfl);
    s->toff = (int32_t*)(base + o_toff);
    s->cur = (int32_t*)(base + o_cur);
    s->tadj = (lid_t*)(base + o_tadj);
    s->apos = (lid_t*)(base + o_apos);
    s->lrA = (uint32_t*)(base + o_lrA);
    s->lrB = (uint32_t*)(base + o_lrB);
    s->ent = (lid_t*)(base + o_ent);
  }
  return end_a > end_b ? end_a : end_b;
}

// misc[] slots
enum { M_WSUM = 0 /* 0..32: warp sums + total */, M_A = 40, M_B, M_N, M_M, M_ERR, M_PICK,
       M_FLAG, M_CNT, M_TOT, M_ROOT, M_NCOMP, M_SUBSET, M_DELTA, M_TSTAR,
       M_BEST64 = 64 /* 64..65, 8-byte aligned */, M_CTR = 72 /* 72..79 counters */ };

#ifdef __CUDACC__

// Exclusive block scan of one int per thread; *total = block sum.  3 barriers.
__device__ __forceinline__ int block_excl_scan(int v, int* total, int32_t* misc) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) misc[M_WSUM + w] = inc;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    int s = lane < nw ? misc[M_WSUM + lane] : 0, si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, si, o);
      if (lane >= o) si += n;
    }
    misc[M_WSUM + lane] = si - s;
    if (lane == 31) misc[M_WSUM + 32] = si;
  }
  __syncthreads();
  int res = misc[M_WSUM + w] + inc - v;
  *total = misc[M_WSUM + 32];
  __syncthreads();
  return res;
}

__device__ __forceinline__ unsigned long long dbl_key(double w) {
  return (unsigned long long)__double_as_longlong(w);  // w >= 0: order preserving
}

__device__ __forceinline__ void atomic_min_u64_smem(unsigned long long* a, unsigned long long v) {
  unsigned long long old = *a;
  while (v < old) {
    unsigned long long seen = atomicCAS(a, old, v);
    if (seen == old) break;
    old = seen;
  }
}

__device__ __forceinline__ void atomic_max_u64_smem(unsigned long long* a, unsigned long long v) {
  unsigned long long old = *a;
  while (v > old) {
    unsigned long long seen = atomicCAS(a, old, v);
    if (seen == old) break;
    old = seen;
  }
}

#endif  // __CUDACC__

}  // namespace rc
