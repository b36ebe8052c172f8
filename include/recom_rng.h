/*
 * recom_rng.h - the random-draw SCHEDULE shared by the device engine and the CPU oracle.
 *
 * The reference draws from Python's global Mersenne Twister in program order
 * (random.random() tree.py:79, random.choice tree.py:377/:545, tree_proposals.py:106).
 * A stream position cannot be reproduced by a parallel machine, so free-running mode
 * replaces it by a counter-based generator: every draw is Philox4x32-10 of
 *     key     = (seed_lo, seed_hi ^ global_chain_id)
 *     counter = (index, tree_no, proposal_no, purpose << 28 | sub)
 * i.e. a pure function of WHAT is drawn, never of the order it is drawn in.  Anything
 * that follows this header draws identical numbers; that is what lets the CUDA engine be
 * checked bit-for-bit against the sequential oracle at full problem sizes.
 *
 * Usable from C (oracle) and CUDA (engine).
 */
#ifndef RECOM_RNG_H
#define RECOM_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define RC_HD __host__ __device__ __forceinline__
#else
#define RC_HD static inline
#endif

enum {
  RC_DRAW_PAIR = 1,   /* which cut edge (tree_proposals.py:106); index = reselect retry  */
  RC_DRAW_WEIGHT = 2, /* random_weight of the j-th edge of the merged subgraph, edges taken
                         in ascending edge id (tree.py:79); index = j >> 2, word j & 3    */
  RC_DRAW_ROOT = 3,   /* BFS root among tree nodes of degree > 1 (tree.py:377)          */
  RC_DRAW_CUT = 4,    /* random.choice(cuts) (tree.py:545)                              */
  RC_DRAW_WROOT = 5,  /* Wilson root (tree.py:112)                                      */
  RC_DRAW_ACCEPT = 7, /* accept.py:35 random.random() of cut_edge_accept                 */
  RC_DRAW_WALK = 6    /* Wilson successor draws (tree.py:119): draw t of a tree, t counting the
                         draws in the order the loop of tree.py:116-125 makes them when the
                         start nodes are taken in ascending node index and every walk runs to
                         its end: index = t >> 2, word t & 3; the value picks among the current
                         node's neighbours inside the pair in ascending node index.  (The walk
                         is a sequential chain on any machine - the device runs one walk per
                         lane - so the draw need not be a function of the node, which keeps the
                         generator off the walk's dependency chain.)                    */
};

typedef struct RcPhilox { uint32_t v[4]; } RcPhilox;

RC_HD uint32_t rc_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

/* Philox4x32-10 (Salmon et al., SC'11), the Random123 constants. */
RC_HD RcPhilox rc_philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                uint32_t c2, uint32_t c3) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = rc_mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = rc_mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += W0; k1 += W1;
  }
  RcPhilox o;
  o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

RC_HD RcPhilox rc_draw(uint64_t seed, uint32_t chain, uint32_t purpose, uint32_t index,
                       uint32_t tree_no, uint32_t proposal_no, uint32_t sub) {
  return rc_philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32) ^ chain, index, tree_no,
                          proposal_no, (purpose << 28) | (sub & 0x0FFFFFFFu));
}

/* Uniform integer in [0, n) from 64 random bits (bias < n * 2^-64). */
RC_HD uint32_t rc_below(uint32_t w0, uint32_t w1, uint32_t n) {
  uint64_t r = ((uint64_t)w0 << 32) | (uint64_t)w1;
#if defined(__CUDA_ARCH__)
  return (uint32_t)__umul64hi(r, (uint64_t)n);
#else
  return (uint32_t)(((unsigned __int128)r * (unsigned __int128)n) >> 64);
#endif
}

/* Uniform integer in [0, n) from 32 random bits (Wilson walk; n = degree). */
RC_HD uint32_t rc_below32(uint32_t w, uint32_t n) { return rc_mulhi32(w, n); }

/* CPython's random.random(): (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53. */
RC_HD double rc_unit_double(uint32_t a, uint32_t b) {
  uint64_t x = ((uint64_t)(a >> 5) << 26) | (uint64_t)(b >> 6);
  return (double)x * (1.0 / 9007199254740992.0);
}

/*
 * Edge weights.  The reference draws random.random() (53 bits) per edge and adds the
 * surcharges of the region columns the edge crosses (tree.py:78-89); only the ORDER of
 * the weights matters (Kruskal's sort, and the max-weight cut choice).  Free-running mode
 * uses 32-bit fixed point: raw = one Philox word (u = raw / 2^32), surcharge s_r as
 * sfix_r = round(s_r * 2^32), sum in 64-bit integers, rescaled monotonically into 32 bits;
 * the total order is (key32, edge id).  Exact ties of key32 (p ~ m^2 / 2^33 per tree) fall
 * to the lower edge id - as Kruskal's stable sort would do with tied floats.
 */
RC_HD uint64_t rc_surcharge_fix(double s) { return (uint64_t)(s * 4294967296.0 + 0.5); }

/* mult = 0 when there are no region columns (key32 = raw). */
RC_HD uint64_t rc_key_mult(uint64_t sfix_total, int n_region_cols) {
  if (n_region_cols <= 0) return 0;
  return 0xFFFFFFFFFFFFFFFFull / ((1ull << 32) + sfix_total);
}

RC_HD uint32_t rc_key_finish(uint32_t raw, uint64_t sadd, uint64_t mult) {
  if (mult == 0) return raw;
  return (uint32_t)((((uint64_t)raw + sadd) * mult) >> 32);
}

/* raw 32-bit weight of local edge j for tree `tree_no` of proposal `proposal_no`. */
RC_HD uint32_t rc_edge_raw(uint64_t seed, uint32_t chain, uint32_t j, uint32_t tree_no,
                           uint32_t proposal_no) {
  RcPhilox p = rc_draw(seed, chain, RC_DRAW_WEIGHT, j >> 2, tree_no, proposal_no, 0);
  return p.v[j & 3u];
}

#endif /* RECOM_RNG_H */
