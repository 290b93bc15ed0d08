// Saddle-point solve for the mixed network system: exact per-edge elimination (two segmented
// scans over the cells of every edge) -> SPD graph-level Schur system on the bifurcation
// multipliers, solved by ONE persistent kernel (exact rake of the tree-like parts, level by level or
// in subtree blocks; PCG on the cyclic core -- Jacobi, or two-level for large cores; device-side
// convergence) -> back-substitution.  Replaces ii_convert + LUSolver(A,"mumps").solve
// (graphnics/flowmodels/flow_models.py:344-348, time_stepping.py:187-192).
//
// Edge-local frame: nodes j = 0..m from u to v, cell k between nodes k and k+1.
//   p-rows :  sigma (q_{k+1} - q_k) = b_p[k]              -> q_j = q_0 + cumF_j
//   q-rows :  (a M + k K) q |_j + P_j - P_{j-1} = b_q[j]   (P = sigma p, P_{-1} = P(u), P_m = P(v))
//   sum    :  q_0 = cond (P(u) - P(v) + rsrc),  cond = 1/(a L),  rsrc = sum b_q - a sum_k h_k (cumF_k+cumF_{k+1})/2
//   lam-row:  sum_in q_e(m) - sum_out q_e(0) = b_lam / sigma   (Kirchhoff) -> graph Laplacian in P(b)
#include <stdlib.h>
#include <cooperative_groups.h>
#include "gnx_common.cuh"
#include "gnx_tma.cuh"
#include "gnx_asm_tile.cuh"
namespace cg = cooperative_groups;

struct EdgeFrame {
  int32_t u, v, m;
  bool flip;
  int64_t qb, pb, hb;
  __device__ __forceinline__ int64_t qdof(const int32_t* __restrict__ loc, int32_t j) const {
    return qb + loc[flip ? m - j : j];
  }
  __device__ __forceinline__ int64_t pdof(int32_t k) const {
    return pb + gnx_gray(flip ? m - 1 - k : k);
  }
};

__device__ __forceinline__ EdgeFrame edge_frame(const gnx_network_t& net, const GnxMixedDims& d, int32_t e) {
  EdgeFrame f;
  f.u = net.edges[2 * e];
  f.v = net.edges[2 * e + 1];
  f.m = d.m;
  f.flip = f.u > f.v;
  f.qb = (int64_t)e * (d.m + 1);
  f.pb = d.p_off + (int64_t)e * d.m;
  f.hb = (int64_t)e * d.m;
  return f;
}

// ------------------------------------------------------------------------------------
// phase 1: edge scalars
// ------------------------------------------------------------------------------------
// Where the edge scalars of (local) edge e go.  Single GPU: three local arrays.  Edge-partitioned
// network: straight into the all-gather buffer of EVERY rank -- peer stores over NVLink issued by
// the thread that holds the value (graphnics_b200/engine.py maps the peers' buffers with
// torch symmetric memory), so the collective is fused into the elimination kernel; the Schur kernel
// opens with the flag exchange that tells the ranks the scalars have landed (gnx_exchange_t).
constexpr int GNX_MAX_PEERS = 8;
struct EdgeOut {
  double* base[GNX_MAX_PEERS];        // destination buffers (n = 1: base[0] = cond)
  int32_t n;
  int64_t off_rsrc, off_ftot;         // offsets of the rsrc / ftot arrays relative to base (in doubles)
  // Subtree-partitioned network (gnx_exchange mode 1): base[0] is this rank's buffer indexed by GLOBAL edge id,
  // base[1..n) the other ranks'; egl[e] = global id of local edge e, bit 31 set if an end of the edge is a
  // REPLICATED bifurcation -- only those scalars cross NVLink.  lcond / lrsrc [E_loc]: what the local
  // back-substitution reads.
  const int32_t* egl;
  double *lcond, *lrsrc;
  __device__ __forceinline__ void put(int32_t e, double c, double r, double f) const {
    if (egl) {
      const int32_t g = egl[e];
      const int32_t ge = g & 0x7fffffff;
      lcond[e] = c; lrsrc[e] = r;
      double* p = base[0];
      p[ge] = c; p[off_rsrc + ge] = r; p[off_ftot + ge] = f;
      if (g < 0) {
#pragma unroll
        for (int k = 1; k < GNX_MAX_PEERS; ++k) {
          if (k < n) {
            double* q = base[k];
            q[ge] = c; q[off_rsrc + ge] = r; q[off_ftot + ge] = f;
          }
        }
        __threadfence_system();       // (a handful of edges per rank: the fence is off every hot path)
      }
      return;
    }
#pragma unroll
    for (int k = 0; k < GNX_MAX_PEERS; ++k) {
      if (k < n) {
        double* p = base[k];
        p[e] = c; p[off_rsrc + e] = r; p[off_ftot + e] = f;
      }
    }
  }
};

// One thread per cell; a segment of `seg` = min(m, blockDim) consecutive threads owns one edge
// (several edges per block when m < 256, several chunks per edge when m > 1024).
struct EdgeTiling { int threads, seg, edges_per_block; };
static inline EdgeTiling edge_tiling(int32_t n) {
  const int64_t m = 1ll << n;
  EdgeTiling t;
  t.threads = m >= 1024 ? 1024 : (m >= 256 ? (int)m : 256);
  t.seg = m < t.threads ? (int)m : t.threads;
  t.edges_per_block = t.threads / t.seg;
  return t;
}

__global__ void __launch_bounds__(1024)
edge_eliminate_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                      double inv_sigma, const double* __restrict__ b, int seg, EdgeOut out) {
  __shared__ double sh[3 * 32];
  const int epb = blockDim.x / seg;
  const int32_t e = blockIdx.x * epb + threadIdx.x / seg;
  const int gl = threadIdx.x & (seg - 1);
  const bool active = e < net.E;
  const int32_t ee = active ? e : net.E - 1;      // keep whole blocks alive for the barriers
  const EdgeFrame f = edge_frame(net, d, ee);
  const int32_t m = d.m;
  double carry = 0.0;
  double acc[3] = {0.0, 0.0, 0.0};                // sum h (cumF_k + cumF_{k+1})/2, sum G, length
  for (int32_t base = 0; base < m; base += seg) {
    const int32_t k = base + gl;                  // cell k and node k
    const double Fk = b[f.pdof(k)] * inv_sigma;
    const double hk = net.h[f.hb + k];
    const double Gk = b[f.qdof(net.loc, k)];
    double tot;
    const double inc = gnx_seg_incl_scan(Fk, seg, sh, &tot);
    const double c1 = carry + inc;                // cumF_{k+1}
    const double c0 = c1 - Fk;                    // cumF_k
    acc[0] += hk * (c0 + c1) * 0.5;
    acc[1] += Gk;
    acc[2] += hk;
    carry += tot;
  }
  if (gl == 0) acc[1] += b[f.qdof(net.loc, m)];
  gnx_seg_sum<3>(acc, seg, sh);
  if (active && gl == 0) {
    const double a = a_edge[e];
    out.put(e, 1.0 / (a * acc[2]), acc[1] - a * acc[0], carry);
  }
}

// ------------------------------------------------------------------------------------
// Edge-tile kernels (m <= 1024): a block owns TILE = 256*CPT cells = TILE/m whole edges.  The
// three contiguous global ranges of the tile (p-block of b, q-block of b, cell lengths) are
// brought into shared memory by the copy engine (cp.async.bulk, gnx_tma.cuh; or by coalesced loads
// of the block) -- that is where all the memory-level parallelism comes from -- and every permuted
// access (Gray-code cell order, first-encounter vertex order, edge orientation) happens in shared
// memory.  A thread owns CPT consecutive cells: serial prefix in registers + block-segmented scan
// of the thread totals.  Outputs leave the same way (staged at their dof slots, one bulk store per range).
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// The per-cell arithmetic of the edge-tile kernels with every rounding spelled out: the chain kernels and the
// persistent step kernel (which recomputes instead of re-reading, and has a Constant-zero-data fast path) must
// produce the same bits, so no expression is left to the compiler's choice of FMA contraction.
// ------------------------------------------------------------------------------------
// right cell of node j: a h_j (2 q_j + q_{j+1}) / 6 with q_{j+1} = q_j + F_j
__device__ __forceinline__ double gnx_mq_right(double a, double h, double q, double F) {
  return __dmul_rn(__dmul_rn(__dmul_rn(a, h), __fma_rn(3.0, q, F)), 1.0 / 6.0);
}
// + left cell: a h_{j-1} (q_{j-1} + 2 q_j) / 6 with q_{j-1} = q_j - F_{j-1}
__device__ __forceinline__ double gnx_mq_left(double mq, double a, double hl, double q, double fl) {
  return __fma_rn(__dmul_rn(__dmul_rn(a, hl), __fma_rn(3.0, q, -fl)), 1.0 / 6.0, mq);
}
// running sum of h_k (cumF_k + cumF_{k+1}) / 2
__device__ __forceinline__ double gnx_s2_step(double s2, double h, double cum, double F) {
  return __fma_rn(__dmul_rn(h, __dadd_rn(__dadd_rn(cum, cum), F)), 0.5, s2);
}
__device__ __forceinline__ double gnx_edge_cond(double a, double L) { return __ddiv_rn(1.0, __dmul_rn(a, L)); }
__device__ __forceinline__ double gnx_edge_rsrc(double a, double sumG, double s2) { return __fma_rn(-a, s2, sumG); }

constexpr int ET_T = 256;

struct TileCtx {
  int32_t e0, nE, le, e, kk0, m;
  bool active, flip;
  int64_t pbase, qbase, hbase;
  int32_t shp, shq, shh;     // staging shifts (parity of the range starts, gnx_tma.cuh); 0 without bulk copies
};

// shared-memory tile of the edge kernels: p-range of b (or x), cell lengths, q-range of b (or x)
template <int CPT> struct TileSmem {
  static constexpr int TILE = 256 * CPT;
  __align__(16) double F[TILE + 2];
  __align__(16) double H[TILE + 2];
  __align__(16) double G[TILE + 256 + 2];
  __align__(8) uint64_t bar, bar_h;
};

template <int CPT>
__device__ __forceinline__ TileCtx tile_ctx_at(const gnx_network_t& net, const GnxMixedDims& d, int32_t tile, int32_t ltid) {
  constexpr int TILE = ET_T * CPT;
  TileCtx c;
  c.m = d.m;
  const int32_t G = TILE >> net.n;
  c.e0 = tile * G;
  c.nE = min(G, net.E - c.e0);
  const int32_t k0 = ltid * CPT;
  int32_t le = k0 >> net.n;
  c.active = le < c.nE;
  c.le = c.active ? le : c.nE - 1;
  c.e = c.e0 + c.le;
  c.kk0 = k0 & (c.m - 1);
  c.flip = net.edges[2 * c.e] > net.edges[2 * c.e + 1];
  c.pbase = d.p_off + (int64_t)c.e0 * c.m;
  c.hbase = (int64_t)c.e0 * c.m;
  c.qbase = (int64_t)c.e0 * (c.m + 1);
  c.shp = c.shq = c.shh = 0;
  return c;
}
template <int CPT, bool BULK>
__device__ __forceinline__ TileCtx tile_ctx_at(const gnx_network_t& net, const GnxMixedDims& d, int32_t tile, int32_t ltid) {
  TileCtx c = tile_ctx_at<CPT>(net, d, tile, ltid);
  if (BULK) { c.shp = (int32_t)(c.pbase & 1); c.shq = (int32_t)(c.qbase & 1); c.shh = (int32_t)(c.hbase & 1); }
  return c;
}
template <int CPT, bool BULK>
__device__ __forceinline__ TileCtx tile_ctx(const gnx_network_t& net, const GnxMixedDims& d) {
  return tile_ctx_at<CPT, BULK>(net, d, (int32_t)blockIdx.x, (int32_t)threadIdx.x);
}

// Tile inputs -> shared memory.  BULK: thread 0 hands the contiguous ranges to the copy engine, completion
// on ts.bar_h (cell lengths) and ts.bar (the two ranges of b); callers run tile_wait() after their next
// __syncthreads().  Otherwise the block loads them itself.
template <int CPT, bool BULK>
__device__ __forceinline__ void tile_load_h(const TileCtx& c, const double* __restrict__ h, TileSmem<CPT>& ts) {
  const int32_t ncell = c.nE * c.m;
  if (BULK) {
    if (threadIdx.x == 0) {
      gnx_mbar_init(&ts.bar_h, 1);
      gnx_mbar_init(&ts.bar, 1);
      gnx_mbar_expect_tx(&ts.bar_h, gnx_bulk_body_bytes(c.hbase, ncell));
      gnx_bulk_stream_in(ts.H, h, c.hbase, ncell, &ts.bar_h);
    }
  } else {
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      const int32_t i = threadIdx.x + r * ET_T;
      if (i < ncell) ts.H[i] = h[c.hbase + i];
    }
  }
}
template <int CPT, bool BULK>
__device__ __forceinline__ void tile_load_b(const TileCtx& c, const double* __restrict__ b, TileSmem<CPT>& ts) {
  const int32_t ncell = c.nE * c.m, nq = c.nE * (c.m + 1);
  if (BULK) {
    if (threadIdx.x == 0) {
      gnx_mbar_expect_tx(&ts.bar, gnx_bulk_body_bytes(c.pbase, ncell) + gnx_bulk_body_bytes(c.qbase, nq));
      gnx_bulk_stream_in(ts.F, b, c.pbase, ncell, &ts.bar);
      gnx_bulk_stream_in(ts.G, b, c.qbase, nq, &ts.bar);
    }
  } else {
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      const int32_t i = threadIdx.x + r * ET_T;
      if (i < ncell) ts.F[i] = b[c.pbase + i];
    }
#pragma unroll
    for (int r = 0; r < CPT + 1; ++r) {
      const int32_t i = threadIdx.x + r * ET_T;
      if (i < nq) ts.G[i] = b[c.qbase + i];
    }
  }
}
// after the __syncthreads() that follows tile_load (it publishes the barrier's initialisation and the
// odd ends thread 0 loaded by hand)
template <int CPT, bool BULK, bool WITH_B>
__device__ __forceinline__ void tile_wait(TileSmem<CPT>& ts) {
  if (BULK) { gnx_mbar_wait(&ts.bar_h, 0); if (WITH_B) gnx_mbar_wait(&ts.bar, 0); }
}

// Tile outputs (staged in ts.F / ts.G at the dof slots) -> out[p range], out[q range].  BULK: callers
// run tile_store_done() before the block exits.
template <int CPT, bool BULK>
__device__ __forceinline__ void tile_store(const TileCtx& c, TileSmem<CPT>& ts, double* __restrict__ out) {
  const int32_t ncell = c.nE * c.m, nq = c.nE * (c.m + 1);
  if (BULK) {
    gnx_fence_async_smem();
    __syncthreads();
    gnx_bulk_stream_out(ts.G, out, c.qbase, nq);
    gnx_bulk_stream_out(ts.F, out, c.pbase, ncell);
    if (threadIdx.x == 0) gnx_bulk_commit();
  } else {
    __syncthreads();
#pragma unroll
    for (int r = 0; r < CPT + 1; ++r) {
      const int32_t i = threadIdx.x + r * ET_T;
      if (i < nq) out[c.qbase + i] = ts.G[i];
    }
#pragma unroll
    for (int r = 0; r < CPT; ++r) {
      const int32_t i = threadIdx.x + r * ET_T;
      if (i < ncell) out[c.pbase + i] = ts.F[i];
    }
  }
}
template <bool BULK> __device__ __forceinline__ void tile_store_done() {
  if (BULK && threadIdx.x == 0) gnx_bulk_wait_read();
}

// Right-hand side with Constant f, g and boundary data only (MixedHydraulicNetwork.L_form :299-332 for
// the steady models): cheap enough to be PRODUCED by the elimination kernel instead of being read by it.
struct ConstRhs {
  double fc, gc;
  const double* pbc_out;     // [E] projected p_bc at the v end (BOUN_OUT term), or NULL
  const double* pbc_node;    // [V] p_bc at the graph nodes (BOUN_IN term), or NULL
  double* b;                 // [N] the right-hand side, WRITTEN by the kernel
};

// FUSED_RHS = false: b is an input, streamed into shared memory.
// FUSED_RHS = true : the tile's part of b is computed in shared memory from `rhs` (same arithmetic as
//                    mixed_rhs_const_tile_kernel), stored to rhs.b with coalesced stores and eliminated
//                    right away -- one launch and one 8 B/dof read less per solve; blocks past the tiles
//                    zero the lambda rows of b.
template <int CPT, bool FUSED_RHS, bool BULK>
__global__ void __launch_bounds__(ET_T)
edge_eliminate_tile_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                           double inv_sigma, const double* __restrict__ b, EdgeOut out, ConstRhs rhs,
                           int32_t nb_tiles, const double* __restrict__ a_cell) {
  __shared__ TileSmem<CPT> ts;
  __shared__ double sh[3 * 32];
  if (FUSED_RHS && (int32_t)blockIdx.x >= nb_tiles) {
    const int64_t r = d.lam_off + (int64_t)(blockIdx.x - nb_tiles) * ET_T + threadIdx.x;
    if (r < d.N) rhs.b[r] = 0.0;
    return;
  }
  gnx_pdl_trigger();                              // the Schur kernel may become resident (it waits for this grid)
  const TileCtx c = tile_ctx<CPT, BULK>(net, d);
  const int32_t m = c.m;
  const int32_t cbF = c.le * m + c.shp, cbH = c.le * m + c.shh, qb = c.le * (m + 1) + c.shq;
  double* const sF = ts.F; double* const sH = ts.H; double* const sG = ts.G;
  tile_load_h<CPT, BULK>(c, net.h, ts);
  if (!FUSED_RHS) tile_load_b<CPT, BULK>(c, b, ts);
  __syncthreads();
  tile_wait<CPT, BULK, !FUSED_RHS>(ts);
  double F[CPT], H[CPT];
  double sumF = 0.0, sumG = 0.0, sumH = 0.0;
  if (FUSED_RHS) {
    // the thread that eliminates cells kk0 .. kk0+CPT-1 also PRODUCES their right-hand side entries: they go to
    // shared memory only to leave as b (bulk store) and stay in registers for the elimination
#pragma unroll
    for (int i = 0; i < CPT; ++i) { H[i] = sH[cbH + c.kk0 + i]; F[i] = 0.0; }
    if (c.active) {
      const int32_t u = net.edges[2 * c.e], v = net.edges[2 * c.e + 1];
      double hl = c.kk0 > 0 ? sH[cbH + c.kk0 - 1] : 0.0;
#pragma unroll
      for (int i = 0; i < CPT; ++i) {                          // edge-local node j (u -> v), cells j-1 | j
        const int32_t j = c.kk0 + i;
        const double hr = H[i];
        // (explicitly rounded: the persistent step kernel recomputes these values and must get the same bits)
        double val = __dmul_rn(__dmul_rn(rhs.gc, __dadd_rn(hl, hr)), 0.5);
        if (j == 0 && net.bix[u] < 0 && rhs.pbc_node) val = __dadd_rn(val, rhs.pbc_node[u]);     // edge OUT of a boundary node: BOUN_IN
        sG[qb + __ldg(net.loc + (c.flip ? m - j : j))] = val;
        const double fv = __dmul_rn(rhs.fc, hr);
        sF[cbF + gnx_gray(c.flip ? m - 1 - j : j)] = fv;
        F[i] = __dmul_rn(fv, inv_sigma);
        sumG += val;
        hl = hr;
      }
      if (c.kk0 + CPT == m) {                                  // the edge's last thread also owns node m
        double val = __dmul_rn(__dmul_rn(rhs.gc, __dadd_rn(hl, 0.0)), 0.5);
        if (net.bix[v] < 0 && rhs.pbc_out) val = __dsub_rn(val, rhs.pbc_out[c.e]);               // edge INTO a boundary node: BOUN_OUT
        sG[qb + __ldg(net.loc + (c.flip ? 0 : m))] = val;
        sumG += val;
      }
    }
#pragma unroll
    for (int i = 0; i < CPT; ++i) { sumF += F[i]; sumH += H[i]; }
    tile_store<CPT, BULK>(c, ts, rhs.b);
  } else {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int32_t kk = c.kk0 + i;
      F[i] = c.active ? __dmul_rn(sF[cbF + gnx_gray(c.flip ? m - 1 - kk : kk)], inv_sigma) : 0.0;
      H[i] = sH[cbH + kk];
      sumG += sG[qb + __ldg(net.loc + (c.flip ? m - kk : kk))];
      sumF += F[i];
      sumH += H[i];
    }
    if (c.kk0 + CPT == m) sumG += sG[qb + __ldg(net.loc + (c.flip ? 0 : m))];
  }
  const int segT = m / CPT;
  double tot;
  const double incl = gnx_seg_incl_scan(sumF, segT, sh, &tot);
  double cum = incl - sumF;                       // cumF at the first cell of this thread
  double s2 = 0.0;
  if (a_cell) {
    // mass coefficient per cell: the weights a_k h_k take the place of h_k, the edge coefficient is 1
    const double* ac = a_cell + (int64_t)c.e * m + c.kk0;
    sumH = 0.0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const double ah = __dmul_rn(c.active ? ac[i] : 0.0, H[i]);
      s2 = gnx_s2_step(s2, ah, cum, F[i]); cum = __dadd_rn(cum, F[i]);
      sumH += ah;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CPT; ++i) { s2 = gnx_s2_step(s2, H[i], cum, F[i]); cum = __dadd_rn(cum, F[i]); }
  }
  double acc[3] = {s2, sumG, sumH};
  gnx_seg_sum<3>(acc, segT, sh);
  if (c.active && c.kk0 == 0) {
    const double a = a_cell ? 1.0 : a_edge[c.e];
    out.put(c.e, gnx_edge_cond(a, acc[2]), gnx_edge_rsrc(a, acc[1], acc[0]), tot);
  }
  if (FUSED_RHS) tile_store_done<BULK>();
}

// tuning knob: tile inputs / outputs through the copy engine (default) or by the block's own loads / stores.
// Bulk copies need 16-byte aligned arrays (torch allocations are 512-byte aligned; checked per call).
static bool edge_bulk(const void* p0, const void* p1 = nullptr, const void* p2 = nullptr) {
  static int bulk = -1;
  if (bulk < 0) { const char* e = getenv("GNX_EDGE_BULK"); bulk = e ? atoi(e) : 1; }
  return bulk && (((uintptr_t)p0 | (uintptr_t)p1 | (uintptr_t)p2) & 15) == 0;
}

template <int CPT, bool BULK>
static void launch_eliminate_tile2(const gnx_network_t* net, const GnxMixedDims& d, const gnx_mixed_coeffs_t* co,
                                   const double* b, const EdgeOut& out, const ConstRhs* rhs, cudaStream_t st) {
  const int G = (ET_T * CPT) >> net->n;
  const int nb_tiles = gnx_blocks(net->E, G);
  if (rhs) {
    const int nb_lam = net->B > 0 ? gnx_blocks(net->B, ET_T) : 0;
    edge_eliminate_tile_kernel<CPT, true, BULK><<<nb_tiles + nb_lam, ET_T, 0, st>>>(
        *net, d, co->a_edge, 1.0 / co->sigma, nullptr, out, *rhs, nb_tiles, co->a_cell);
  } else {
    ConstRhs none = {0.0, 0.0, nullptr, nullptr, nullptr};
    edge_eliminate_tile_kernel<CPT, false, BULK><<<nb_tiles, ET_T, 0, st>>>(*net, d, co->a_edge, 1.0 / co->sigma, b,
                                                                             out, none, nb_tiles, co->a_cell);
  }
}
template <int CPT>
static void launch_eliminate_tile(const gnx_network_t* net, const GnxMixedDims& d, const gnx_mixed_coeffs_t* co,
                                  const double* b, const EdgeOut& out, const ConstRhs* rhs, cudaStream_t st) {
  if (edge_bulk(net->h, rhs ? rhs->b : b)) launch_eliminate_tile2<CPT, true>(net, d, co, b, out, rhs, st);
  else launch_eliminate_tile2<CPT, false>(net, d, co, b, out, rhs, st);
}

static int edge_cpt() {           // tuning knob: cells per thread of the edge-tile kernels (default 4)
  static int cpt = -1;
  if (cpt < 0) { const char* e = getenv("GNX_EDGE_CPT"); cpt = e ? atoi(e) : 4; }
  return cpt;
}

static bool const_tiles_apply(const gnx_network_t* net, const void* b_or_x);
static int launch_eliminate_const(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const EdgeOut& out,
                                  const ConstRhs& rhs, cudaStream_t st);

static int launch_eliminate(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const double* b,
                            const EdgeOut& out, cudaStream_t st, const ConstRhs* rhs = nullptr) {
  const GnxMixedDims d = gnx_mixed_dims(*net);
  if (rhs && !co->a_cell && const_tiles_apply(net, rhs->b)) return launch_eliminate_const(net, co, out, *rhs, st);
  if (co->a_cell && net->n > 10) { gnx_set_error("per-cell mass coefficients need n <= 10"); return GNX_ERR_ARG; }
  if (net->n == 0) launch_eliminate_tile<1>(net, d, co, b, out, rhs, st);
  else if (net->n == 1 || (edge_cpt() == 2 && net->n <= 9)) launch_eliminate_tile<2>(net, d, co, b, out, rhs, st);
  else if (net->n <= 10) launch_eliminate_tile<4>(net, d, co, b, out, rhs, st);
  else if (rhs) { gnx_set_error("fused rhs + elimination needs n <= 10"); return GNX_ERR_ARG; }
  else {
    const EdgeTiling t = edge_tiling(net->n);
    edge_eliminate_kernel<<<gnx_blocks(net->E, t.edges_per_block), t.threads, 0, st>>>(
        *net, d, co->a_edge, 1.0 / co->sigma, b, t.seg, out);
  }
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

static void edge_out_local(EdgeOut& out, double* cond, double* rsrc, double* ftot) {
  for (int k = 0; k < GNX_MAX_PEERS; ++k) out.base[k] = nullptr;
  out.base[0] = cond; out.n = 1;
  out.off_rsrc = rsrc - cond; out.off_ftot = ftot - cond;
  out.egl = nullptr; out.lcond = out.lrsrc = nullptr;
}

extern "C" int gnx_edge_eliminate_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                                        const double* b, double* cond, double* rsrc, double* ftot,
                                        void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && b && cond && rsrc && ftot && net->h && co->sigma != 0.0);
  EdgeOut out;
  edge_out_local(out, cond, rsrc, ftot);
  return launch_eliminate(net, co, b, out, (cudaStream_t)stream);
}

// MixedHydraulicNetwork.L_form with Constant f / g (flow_models.py:299-332) + elimination in ONE kernel:
// b [N] is an OUTPUT (identical to gnx_assemble_rhs_mixed with f_nodal = g_nodal = NULL).  n <= 10.
// ex != NULL: edge-partitioned variant, scalars stored to every rank.

// Edge-partitioned network (gnx_exchange): every edge scalar is stored straight into the gather buffer of
// every rank; block of rank q at q*3*chunk, the world flag words behind the world blocks.
static int edge_out_exchange(EdgeOut& out, const gnx_exchange_t* ex, int32_t E, double* lcond = nullptr, double* lrsrc = nullptr) {
  GNX_CHECK_ARG(ex->peers_host && ex->world >= 1 && ex->world <= GNX_MAX_PEERS && ex->rank >= 0 &&
                ex->rank < ex->world && ex->epoch >= 0);
  for (int k = 0; k < GNX_MAX_PEERS; ++k) out.base[k] = nullptr;
  out.egl = nullptr; out.lcond = out.lrsrc = nullptr;
  out.n = ex->world;
  if (ex->mode == 1) {               // global-indexed buffers, own first
    GNX_CHECK_ARG(ex->egl && lcond && lrsrc && ex->E_glob >= E && ex->B_glob >= 0);
    int k = 1;
    for (int q = 0; q < ex->world; ++q) {
      GNX_CHECK_ARG(ex->peers_host[q] != nullptr);
      if (q == ex->rank) out.base[0] = ex->peers_host[q]; else out.base[k++] = ex->peers_host[q];
    }
    out.off_rsrc = ex->E_glob; out.off_ftot = 2 * (int64_t)ex->E_glob;
    out.egl = ex->egl; out.lcond = lcond; out.lrsrc = lrsrc;
    return GNX_OK;
  }
  GNX_CHECK_ARG(ex->chunk >= E);
  for (int k = 0; k < ex->world; ++k) {
    GNX_CHECK_ARG(ex->peers_host[k] != nullptr);
    out.base[k] = ex->peers_host[k] + (int64_t)ex->rank * 3 * ex->chunk;
  }
  out.off_rsrc = ex->chunk; out.off_ftot = 2 * (int64_t)ex->chunk;
  return GNX_OK;
}

extern "C" int gnx_rhs_eliminate_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const,
                                       double g_const, const double* pbc_out, const double* pbc_node, double* b,
                                       double* cond, double* rsrc, double* ftot, const gnx_exchange_t* ex,
                                       void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && b && net->h && co->sigma != 0.0 && net->n <= 10);
  EdgeOut out;
  if (ex) {
    if (int rc = edge_out_exchange(out, ex, net->E, cond, rsrc)) return rc;
  } else {
    GNX_CHECK_ARG(cond && rsrc && ftot);
    edge_out_local(out, cond, rsrc, ftot);
  }
  ConstRhs rhs = {f_const, g_const, pbc_out, pbc_node, b};
  return launch_eliminate(net, co, nullptr, out, (cudaStream_t)stream, &rhs);
}

extern "C" int gnx_edge_eliminate_mixed_p2p(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                                            const double* b, const gnx_exchange_t* ex, double* cond, double* rsrc,
                                            void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && b && net->h && co->sigma != 0.0 && ex);
  EdgeOut out;
  if (int rc = edge_out_exchange(out, ex, net->E, cond, rsrc)) return rc;
  return launch_eliminate(net, co, b, out, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------
// Schur system: one thread per bifurcation gathers its incident edges (no atomics)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
schur_build_kernel(gnx_network_t net, double inv_sigma, const double* __restrict__ cond,
                   const double* __restrict__ rsrc, const double* __restrict__ ftot,
                   const double* __restrict__ b_lam, double* __restrict__ sdiag,
                   double* __restrict__ soff, int32_t* __restrict__ scol, double* __restrict__ srhs) {
  const int32_t bi = blockIdx.x * blockDim.x + threadIdx.x;
  if (bi >= net.B) return;
  const int32_t node = net.bif_nodes[bi];
  double diag = 0.0, rhs = b_lam ? -b_lam[bi] * inv_sigma : 0.0;
  for (int32_t k = net.inc_ptr[node]; k < net.inc_ptr[node + 1]; ++k) {
    const int32_t code = net.inc_edge[k];
    const int32_t e = code >> 1, end = code & 1;
    const double c = cond[e];
    const int32_t other = end ? net.edges[2 * e] : net.edges[2 * e + 1];
    const int32_t ob = net.bix[other];
    diag += c;
    rhs += end ? (c * rsrc[e] + ftot[e]) : (-c * rsrc[e]);
    soff[k] = ob >= 0 ? -c : 0.0;
    scol[k] = ob;
  }
  sdiag[bi] = diag;
  srhs[bi] = rhs;
}

extern "C" int gnx_schur_build(const gnx_network_t* net, double sigma, const double* cond,
                               const double* rsrc, const double* ftot, const double* b_lam,
                               double* sdiag, double* soff, int32_t* scol, double* srhs, void* stream) {
  GNX_CHECK_ARG(net && cond && rsrc && ftot && sdiag && soff && scol && srhs && sigma != 0.0);
  if (net->B == 0) return GNX_OK;
  schur_build_kernel<<<gnx_blocks(net->B, 128), 128, 0, (cudaStream_t)stream>>>(
      *net, 1.0 / sigma, cond, rsrc, ftot, b_lam, sdiag, soff, scol, srhs);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// Schur solve: ONE persistent kernel per solve.
//   phase 0  build   : every bifurcation gathers its incident edges -> d (diagonal), r (rhs)
//   phase 1  rake up : tree-like parts are eliminated exactly, level by level (plan built once on
//                      the host, graphnics_b200/schur_plan.py); a parent gathers its children, no atomics
//   phase 2  core    : Jacobi-PCG on what is left (the cyclic core; empty for a tree), fused
//                      p-update+SpMV+<p,Sp> | x,r,z-update+<r,z>,<r,r>, deterministic reductions
//                      that are bitwise identical in every block (uniform control flow)
//   phase 3  rake dn : back-substitution through the levels, writes P and lambda = P / sigma
// Barriers are __syncthreads (one CTA, small systems) or cooperative grid syncs; levels that fit
// one CTA are swept by block 0 alone so a deep tree costs two grid barriers per LARGE level only.
// Nothing returns to the host unless the caller asks for the iteration count / residual.
// ------------------------------------------------------------------------------------
constexpr int SCH_T = 1024;

struct SchurArgs {
  gnx_network_t net;
  gnx_schur_plan_t plan;
  double inv_sigma;
  const double *cond, *rsrc, *ftot, *b_lam, *b_node;
  double *P, *lam_out;
  double *d, *r;                              // [B]
  double *xc, *rc, *zc, *wc, *p0, *p1, *dg;   // [nc]
  double *cval;                               // [nnzc]
  double *Eg, *sE;                            // two-level PCG: coarse matrix [na*na], aggregate sums of r [na]
  double *s1, *dE;                            // fine aggregates [na*ns]: sums of r, diagonal of W1^T S W1
  double *part[5];                            // [nblk] each
  double *scal;                               // [8]: tol2, converged, iterations, rr, bb
  double rtol;
  int32_t max_iter, warm;
  // edge scalar e lives at arr[(e / chunk) * rank_stride + e % chunk]: the layout an all-gather of
  // per-rank [cond | rsrc | ftot] blocks produces (one collective instead of three).  A single
  // GPU passes chunk = INT32_MAX, i.e. plain arr[e].
  int32_t chunk, rank_stride;
  // edge-partitioned solve with the flag protocol (gnx_exchange): tell every rank that my scalars have landed
  // (my elimination kernel is complete when this kernel runs), then wait until every rank has told me
  const long long* wait_flags;                // [wait_world] flag words of my gather buffer, or NULL
  long long* sig_flags[GNX_MAX_PEERS];        // my flag word in every rank's gather buffer
  long long wait_epoch;
  int32_t wait_world;
  // subtree-partitioned network (gnx_exchange mode 1, plan.ncls != NULL): the flag exchange happens INSIDE the
  // solve, after the rank-local sweeps (plan.l_sync); xbuf[k] = rank k's buffer (NULL for me), where the final
  // (d, r) of my exported subtree roots go (d at xd_off + node, r at xr_off + node, like a.d / a.r in mine)
  double* xbuf[GNX_MAX_PEERS];
  int32_t nx;
  int64_t xd_off, xr_off;
  int32_t phase;                              // 0: whole solve, flags in the kernel; 1: up to the exchange; 2: after it
  int32_t xpoll;                              // 1: exported roots validate themselves (GNX_XSENT); 0: flag after the data
  int32_t pcg_grid_sync;                      // multilevel PCG: 0 tagged all-reduce (tl_allreduce), 1 grid.sync() + partials
  __device__ __forceinline__ int64_t eidx(int32_t e) const {
    if (chunk == 2147483647) return e;          // (uniform: everything but the gathered layout -- no integer division)
    const int32_t r = e / chunk;
    return (int64_t)r * rank_stride + (e - r * chunk);
  }
};

extern "C" int64_t gnx_schur_work_doubles(const gnx_schur_plan_t* plan) {
  if (!plan) return 0;
  const int64_t nval = plan->ell_w > 0 ? (int64_t)plan->ell_w * plan->nc : 0;
  const int64_t coarse = plan->na > 0 ? (int64_t)plan->na * plan->na + plan->na + 2 * (int64_t)GNX_SUB_MAX : 0;
  return 2 * (int64_t)plan->B + 7 * (int64_t)plan->nc + (plan->nnzc > nval ? plan->nnzc : nval) + 5 * 2048 + 32 + coarse;
}

// All-reduce of K values over the block with ONE barrier.  `sh` holds K*32 doubles; callers
// alternate between two buffers so that a fast warp never overwrites partials still being read.
template <int K>
__device__ __forceinline__ void block_allreduce(double (&v)[K], double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(GNX_FULL, v[k], o);
    if (lane == 0) sh[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = (lane < nw) ? sh[k * 32 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(GNX_FULL, t, o);
    v[k] = t;
  }
}

// How the CTAs of one solve synchronise:
//   SCH_SINGLE   one CTA, __syncthreads
//   SCH_CLUSTER  one thread-block cluster (<= 16 CTAs on one GPC), hardware cluster barrier (~0.2 us)
//   SCH_GRID     cooperative grid, software grid barrier (~2-3 us) -- only for phases that need > 16 CTAs
enum { SCH_SINGLE = 0, SCH_CLUSTER = 1, SCH_GRID = 2, SCH_GRID_TL = 3 };   // 3: grid + two-level PCG on the core

template <int MODE>
__device__ __forceinline__ void sch_sync() {
  if (MODE == SCH_SINGLE) __syncthreads();
  else if (MODE == SCH_CLUSTER) cg::this_cluster().sync();
  else cg::this_grid().sync();                  // SCH_GRID, SCH_GRID_TL
}

// partials of every block are re-reduced by every block in the same order -> identical bits
template <int MODE, int K>
__device__ __forceinline__ void grid_allreduce(double (&v)[K], double* const* part, int nblk, double* sh) {
  block_allreduce<K>(v, sh);
  if (MODE == SCH_SINGLE) return;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) part[k][blockIdx.x] = v[k];
  }
  sch_sync<MODE>();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) t += __ldcg(part[k] + i);
    v[k] = t;
  }
  __syncthreads();                            // sh reuse
  block_allreduce<K>(v, sh);
}

// 1/x to within an ulp or two: hardware seed (MUFU.RCP64H, ~20 bits) + two Newton steps = 4 dependent FMAs, where the
// correctly rounded division is a ~30-instruction dependent chain -- and one reciprocal sits on the critical path of
// EVERY level of the one-CTA rake.  Deterministic; x is a positive, finite Schur diagonal.
__device__ __forceinline__ double gnx_rcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = __fma_rn(-x, y, 1.0);
  y = __fma_rn(y, e, y);
  e = __fma_rn(-x, y, 1.0);
  return __fma_rn(y, e, y);
}

// Diagonal and right-hand side of the Schur row of bifurcation bi (gather over its incident edges, no atomics)
__device__ __forceinline__ void schur_row(const SchurArgs& a, int32_t bi, double& diag, double& rhs) {
  const gnx_network_t& net = a.net;
  const int32_t node = net.bif_nodes[bi];
  diag = 0.0;
  rhs = a.b_lam ? -a.b_lam[bi] * a.inv_sigma : (a.b_node ? a.b_node[node] * a.inv_sigma : 0.0);
  for (int32_t k = net.inc_ptr[node]; k < net.inc_ptr[node + 1]; ++k) {
    const int32_t code = net.inc_edge[k];
    const int32_t e = code >> 1, end = code & 1;
    const int64_t ei = a.eidx(e);
    const double c = a.cond[ei];
    diag += c;
    rhs += end ? (c * a.rsrc[ei] + a.ftot[ei]) : (-c * a.rsrc[ei]);
  }
}

// The index data schur_row(bi) walks (never written during a solve), touched ahead of time: a replicated node's row
// can only be built after the exchange, and by then these dependent loads are L1 / L2 hits instead of a chain of
// cold misses on the solve's critical path.  Returns a checksum (the caller keeps it alive).
__device__ __forceinline__ int32_t schur_row_warm(const SchurArgs& a, int32_t bi) {
  const gnx_network_t& net = a.net;
  const int32_t node = net.bif_nodes[bi];
  int32_t acc = node;
  for (int32_t k = net.inc_ptr[node]; k < net.inc_ptr[node + 1]; ++k) acc ^= net.inc_edge[k];
  return acc;
}

// Flag exchange between the ranks of a partitioned solve, run by the first `wait_world` threads of ONE CTA: tell every
// rank that what I export has landed (the stores were followed by system-scope fences and a barrier), wait for all.
__device__ __forceinline__ void sch_flag_exchange(const SchurArgs& a) {
  if ((int)threadIdx.x < a.wait_world) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" :: "l"(a.sig_flags[threadIdx.x]), "l"(a.wait_epoch) : "memory");
    const long long* f = a.wait_flags + threadIdx.x;
    long long v;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= a.wait_epoch) break;
      __nanosleep(spins < 4096 ? 32 : 2000);    // ranks run in step: poll fast first, then patiently
      if (++spins > (1u << 25)) __trap();       // about a minute: a rank is missing -- fail instead of hanging the GPU
    }
  }
}

// In-kernel exchange of the exported roots without a flag round trip: a root's final (d, r) travel as two plain
// 8-byte peer stores and VALIDATE THEMSELVES -- the slot holds GNX_XSENT (a NaN no computation produces) until the
// value lands, the reader polls the word itself and puts the sentinel back once it has the value.  (A flag would need
// a release after the data: the storing SM waits for the NVLink acknowledgement of its data stores -- 2-4 us measured
// -- before the flag may even leave.)  The two exchange buffers alternate per solve and a rank cannot run two solves
// ahead of a peer, so a slot is never rewritten before its reader has reset it; engine.py fills the (d, r) region
// with the sentinel when it creates the buffers.
#define GNX_XSENT 0x7FF85EA71E550001ull
__device__ __forceinline__ double sch_poll_word(double* p) {
  unsigned long long v;
  unsigned spins = 0;
  for (;;) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (v != GNX_XSENT) break;
    __nanosleep(spins < 4096 ? 20 : 2000);
    if (++spins > (1u << 25)) __trap();         // about a minute: a rank is missing -- fail instead of hanging the GPU
  }
  return __longlong_as_double((long long)v);
}
__device__ __forceinline__ void sch_take_root(const SchurArgs& a, int32_t c, double& dc, double& rc) {
  dc = sch_poll_word(a.d + c);
  rc = sch_poll_word(a.r + c);
  *reinterpret_cast<volatile unsigned long long*>(a.d + c) = GNX_XSENT;
  *reinterpret_cast<volatile unsigned long long*>(a.r + c) = GNX_XSENT;
}

__device__ __forceinline__ void sch_wait_flags(const SchurArgs& a) {
  if ((int)threadIdx.x < a.wait_world) {
    const long long* f = a.wait_flags + threadIdx.x;
    long long v;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= a.wait_epoch) break;
      __nanosleep(spins < 4096 ? 32 : 2000);
      if (++spins > (1u << 25)) __trap();
    }
  }
}

// The exchange of a subtree-partitioned solve, run by the first `world` threads of ONE CTA after a barrier that
// made my roots' final (d, r) visible in MY global arrays: thread q stores them into rank q's buffer and then
// RELEASES its flag word there (st.release.sys orders the thread's own stores before the flag: no separate
// system-scope fence; the edge scalars other threads sent earlier were fenced by their senders, EdgeOut::put),
// then waits for rank q's flag in my buffer.  do_send / do_flags: phases of gnx_exchange_t.
__device__ __forceinline__ void sch_export_roots(const SchurArgs& a, bool do_send, bool do_flags) {
  const gnx_schur_plan_t& pl = a.plan;
  const int q = threadIdx.x;
  if (q >= a.wait_world) return;
  // diagnostic stamps of the thread that talks to the first OTHER rank: scal[18] roots sent | [19] flag released | [20] flag seen
  const bool dbg = a.xbuf[q] && (q == 0 || (q == 1 && !a.xbuf[0]));
  auto stamp = [&](int k) { if (dbg) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.scal[k] = (double)(t & 0xffffffffffull); } };
  if (do_send && a.xbuf[q]) {
    for (int32_t k = 0; k < pl.nxl; ++k) {
      const int32_t node = pl.xlist[k];
      a.xbuf[q][a.xd_off + node] = __ldcg(a.d + node);
      a.xbuf[q][a.xr_off + node] = __ldcg(a.r + node);
    }
    if (!do_flags) __threadfence_system();
  }
  stamp(18);
  if (do_flags) {
    asm volatile("st.release.sys.global.s64 [%0], %1;" :: "l"(a.sig_flags[q]), "l"(a.wait_epoch) : "memory");
    stamp(19);
    const long long* f = a.wait_flags + q;
    long long v;
    unsigned spins = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
      if (v >= a.wait_epoch) break;
      __nanosleep(spins < 4096 ? 32 : 2000);
      if (++spins > (1u << 25)) __trap();
    }
    stamp(20);
  }
}

// The TOP part of the rake (levels >= l_split, <= GNX_TOP_MAX nodes) swept by ONE CTA: a thread
// owns up to TOP_ROWS nodes, all plan data are fetched up front (independent loads), the running
// diagonal / rhs / link weight live in shared memory, and a level costs one block barrier instead
// of a chain of dependent global loads.  Children below the top part are final in global memory.
constexpr int TOP_ROWS = GNX_TOP_MAX / SCH_T;
constexpr int CL_CTAS = 16;          // CTAs of the one cluster a cluster-mode solve runs as
constexpr int CL_T = 512;            // threads per CTA in cluster mode (128 registers per thread available)
constexpr size_t SCH_SMEM = 3 * GNX_TOP_MAX * sizeof(double);

// CLUSTER = false: one CTA of SCH_T threads, top-local index ti = tid + q*SCH_T.
// CLUSTER = true : the top part is spread over the CTAs of one cluster, CTA `rank` owns the slots
//                  [rank*S, (rank+1)*S) (S = plan.top_S, a power of two); a child / parent in another
//                  CTA is read over distributed shared memory and a level costs one hardware cluster
//                  barrier -- a whole tree of up to 16*2048 bifurcations then needs no grid barrier.
// Subtree-partitioned network (plan.ncls != NULL, CLUSTER = false): the top part holds this rank's upper local nodes
// (levels < l_sync) and ALL replicated raked nodes (levels >= l_sync).  Between the two the ranks exchange: my
// exported roots' final (d, r) go to everybody, flags are swapped, and only then are the rows of the replicated
// nodes built (their edges' scalars and their local children's pairs have arrived from the owners).  A replicated
// node folds its local children -- mine or another rank's, all the same -- from global memory and keeps only
// replicated children in registers, so every rank performs the same operations on the same bits.
// Returns true when the solve stops at the exchange (a.phase == 1).
// wait_fn(): called once, after the rows' STATIC plan data are in registers and before anything the step produces
// is read -- the persistent step kernel's solver CTA waits for the workers' edge scalars there (its plan loads then
// overlap the elimination phase instead of following it); a no-op everywhere else.
struct SchNoWait { __device__ __forceinline__ void operator()() const {} };
template <bool CLUSTER, int ROWS, class WaitFn = SchNoWait>
__device__ __forceinline__ bool top_solve(const SchurArgs& a, double* sd, double* sr, double* sw, long long* tick = nullptr,
                                          WaitFn wait_fn = WaitFn()) {
  const gnx_schur_plan_t& pl = a.plan;
  const int tid = threadIdx.x, T = blockDim.x;
  cg::cluster_group cluster = cg::this_cluster();
  const int32_t S = CLUSTER ? pl.top_S : GNX_TOP_MAX;
  const int sshift = 31 - __clz(S);
  const int32_t base = CLUSTER ? (int32_t)cluster.block_rank() * S : 0;
  const int32_t* const ncls = CLUSTER ? nullptr : pl.ncls;
  const int l_sync = ncls ? pl.l_sync : pl.nl;
  auto rd = [&](double* arr, int32_t tc) -> double {
    if (!CLUSTER) return arr[tc];
    const double* p = cluster.map_shared_rank(arr, tc >> sshift);
    return p[tc & (S - 1)];
  };
  auto lsync = [&]() { if (CLUSTER) cluster.sync(); else __syncthreads(); };
  // sd[] holds 1/d of a finished node (all a parent and the down-sweep need), sr[] its r, then its solution;
  // sw[] the link weight to the parent.  The first two top children of a node sit in registers (a binary
  // tree has no more), further ones are re-read from the plan: a level of the up-sweep then costs shared
  // memory reads, a handful of FMAs, ONE reciprocal and a barrier -- no global load, no second division.
  // (zero-initialised for the compiler's sake: the lambdas below capture them before the set-up loop fills them)
  int32_t node[ROWS] = {}, lev[ROWS] = {}, ptop[ROWS] = {}, k0[ROWS] = {}, k1[ROWS] = {}, c0[ROWS] = {}, c1[ROWS] = {}, nct[ROWS] = {};
  double d[ROWS] = {}, r[ROWS] = {}, w[ROWS] = {};
  bool repl[ROWS] = {};
  // is child c of row q kept in registers / shared memory (else: final in global memory)?
  auto top_child = [&](int q, int32_t c) -> int32_t {
    const int32_t tc = pl.top_of[c];
    if (tc >= 0 && repl[q] && !(ncls[c] & 1)) return -1;
    return tc;
  };
  // numeric part of a row's set-up: row (from global memory, or built here for a replicated node), link weight,
  // children below the top part folded in
  const int4* const trec = CLUSTER ? nullptr : reinterpret_cast<const int4*>(pl.top_rec);
  const bool xin = !CLUSTER && ncls && a.phase == 0 && a.wait_flags && a.xpoll;   // partitioned solve, exchange in this kernel
  if (!CLUSTER && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.scal[16] = (double)(t & 0xffffffffffull); }
  // the same from the row's packed record (plan.top_rec): one 64-byte load, then every scalar it names at once
  int4 keep[ROWS == 1 ? 4 : 1] = {};                           // one row per thread: its record stays in registers
  auto numeric_rec = [&](int q) {
    const int32_t li = tid + q * T;
    int4 r0, r1, r2, r3;
    if (ROWS == 1) { r0 = keep[0]; r1 = keep[1 % (ROWS == 1 ? 4 : 1)]; r2 = keep[2 % (ROWS == 1 ? 4 : 1)]; r3 = keep[3 % (ROWS == 1 ? 4 : 1)]; }
    else {
      r0 = __ldg(trec + 4 * (base + li)); r1 = __ldg(trec + 4 * (base + li) + 1);
      r2 = __ldg(trec + 4 * (base + li) + 2); r3 = __ldg(trec + 4 * (base + li) + 3);
    }
    const int32_t i = r0.x;
    w[q] = r0.w >= 0 ? -a.cond[a.eidx(r0.w)] : 0.0;
    double di = 0.0;
    double ri = a.b_lam ? -a.b_lam[i] * a.inv_sigma : (a.b_node ? a.b_node[r2.w] * a.inv_sigma : 0.0);
    const int32_t inc[4] = {r3.x, r3.y, r3.z, r3.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (inc[k] >= 0) {
        const int64_t ei = a.eidx(inc[k] >> 1);
        const double c = a.cond[ei];
        di += c;
        ri += (inc[k] & 1) ? (c * a.rsrc[ei] + a.ftot[ei]) : (-c * a.rsrc[ei]);
      }
    }
    const int32_t bc[2] = {r1.w, r2.x}, be[2] = {r2.y, r2.z};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (bc[k] >= 0) {                                     // child below the top part: final in global memory
        const double wc = -a.cond[a.eidx(be[k])];
        double dc, rc;
        if (xin && repl[q] && (ncls[bc[k]] & 2)) sch_take_root(a, bc[k], dc, rc);     // another rank's subtree root
        else { dc = __ldcg(a.d + bc[k]); rc = __ldcg(a.r + bc[k]); }
        const double t = wc / dc;
        di -= wc * t;
        ri -= t * rc;
      }
    }
    c0[q] = r1.x; c1[q] = r1.y; nct[q] = r1.z;
    d[q] = di; r[q] = ri;
    sr[li] = ri; sw[li] = w[q];
    if (nct[q] == 0) sd[li] = gnx_rcp(di);
  };
  auto numeric = [&](int q) {
    if (trec && !(nct[q] & (1 << 30))) { numeric_rec(q); return; }
    const int32_t li = tid + q * T;
    const int32_t i = node[q];
    const int32_t par = pl.parent[i];
    w[q] = par >= 0 ? -a.cond[a.eidx(pl.pedge[i])] : 0.0;
    double di, ri;
    if (repl[q] || trec) schur_row(a, i, di, ri);          // (with records nobody has built the top rows beforehand)
    else { di = __ldcg(a.d + i); ri = __ldcg(a.r + i); }
    nct[q] = 0; c0[q] = c1[q] = -1;
    for (int32_t k = k0[q]; k < k1[q]; ++k) {
      const int32_t c = pl.ch_idx[k];
      const int32_t tc = top_child(q, c);
      if (tc < 0) {                                        // child below the top part: final in global memory
        const double wc = -a.cond[a.eidx(pl.pedge[c])];
        double dc, rc;
        if (xin && repl[q] && (ncls[c] & 2)) sch_take_root(a, c, dc, rc);
        else { dc = __ldcg(a.d + c); rc = __ldcg(a.r + c); }
        const double t = wc / dc;
        di -= wc * t;
        ri -= t * rc;
      } else {
        if (nct[q] == 0) c0[q] = tc; else if (nct[q] == 1) c1[q] = tc;
        ++nct[q];
      }
    }
    d[q] = di; r[q] = ri;
    sr[li] = ri; sw[li] = w[q];
    if (nct[q] == 0) sd[li] = gnx_rcp(di);                    // no top children: final already
  };
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const int32_t li = tid + q * T;
    const int32_t ti = base + li;
    lev[q] = -1; node[q] = 0; ptop[q] = -1; k0[q] = k1[q] = 0; d[q] = 1.0; r[q] = 0.0; w[q] = 0.0;
    c0[q] = c1[q] = -1; nct[q] = 0; repl[q] = false;
    if (li < S && ti < pl.nt) {
      if (trec) {
        const int4 r0 = __ldg(trec + 4 * ti), r1 = __ldg(trec + 4 * ti + 1);
        if (ROWS == 1) {
          keep[0] = r0; keep[1 % (ROWS == 1 ? 4 : 1)] = r1;
          keep[2 % (ROWS == 1 ? 4 : 1)] = __ldg(trec + 4 * ti + 2); keep[3 % (ROWS == 1 ? 4 : 1)] = __ldg(trec + 4 * ti + 3);
        }
        node[q] = r0.x; lev[q] = r0.y; ptop[q] = r0.z;
        nct[q] = r1.z;                                       // (bit 30: numeric() takes the generic path for this row)
        repl[q] = ncls && (ncls[r0.x] & 1);
        if (r1.z & (1 << 30)) { k0[q] = pl.ch_ptr[r0.x]; k1[q] = pl.ch_ptr[r0.x + 1]; }
      } else {
        const int32_t i = pl.top_nodes[ti];
        node[q] = i;
        lev[q] = pl.level_of[i];
        repl[q] = ncls && (ncls[i] & 1);
        const int32_t par = pl.parent[i];
        if (par >= 0) ptop[q] = pl.top_of[par];
        k0[q] = pl.ch_ptr[i]; k1[q] = pl.ch_ptr[i + 1];
        if (repl[q]) {                                     // warm the index data of the row built after the exchange
          int32_t acc = schur_row_warm(a, i) ^ (par >= 0 ? pl.pedge[i] : 0);
          for (int32_t k = k0[q]; k < k1[q]; ++k) { const int32_t c = pl.ch_idx[k]; acc ^= pl.top_of[c] ^ ncls[c] ^ pl.pedge[c]; }
          if (acc == 0x7fffffff) sw[li] = 0.0;             // (never true in practice: keeps the loads alive)
        }
      }
    }
  }
  wait_fn();                                               // ---- from here on: data of this step
  // my edge scalars have landed in every rank's buffer (their senders fenced them at system scope, and this CTA runs
  // after they all arrived): tell the ranks NOW -- the flags then cross NVLink under the local sweep
  if (xin && tid < a.wait_world)
    asm volatile("st.release.sys.global.s64 [%0], %1;" :: "l"(a.sig_flags[tid]), "l"(a.wait_epoch) : "memory");
  if (!CLUSTER && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.scal[25] = (double)(t & 0xffffffffffull); }
#pragma unroll
  for (int q = 0; q < ROWS; ++q)
    if (lev[q] >= 0 && !repl[q]) numeric(q);
  if (!CLUSTER && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.scal[17] = (double)(t & 0xffffffffffull); }
  lsync();
  if (tick) tick[0] = clock64();
  auto stamp = [&](int k) { if (!CLUSTER && tid == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.scal[8 + k] = (double)(t & 0xffffffffffull); } };
  stamp(1);
  // the exchange, between the last local and the first replicated level
  auto exchange = [&]() -> bool {
    stamp(2);
    // my exported roots' final (d, r): into my own arrays and -- in-kernel exchange -- straight into every other
    // rank's, by the thread that holds them (the flag threads' release below is cumulative over the barrier)
    const bool direct = a.phase == 0;
#pragma unroll
    for (int q = 0; q < ROWS; ++q)
      if (lev[q] >= 0 && !repl[q] && (ncls[node[q]] & 4)) {
        a.d[node[q]] = d[q]; a.r[node[q]] = r[q];
        if (direct) {
#pragma unroll
          for (int k = 0; k < GNX_MAX_PEERS; ++k)
            if (k < a.wait_world && a.xbuf[k]) { a.xbuf[k][a.xd_off + node[q]] = d[q]; a.xbuf[k][a.xr_off + node[q]] = r[q]; }
        }
      }
    lsync();
    stamp(3);
    if (xin) sch_wait_flags(a);                                      // (released long ago: the edge scalars are here)
    else sch_export_roots(a, a.phase == 1, a.phase == 0);            // (phase 0: the flag threads' release covers the roots)
    lsync();
    if (a.phase == 1) return true;
    stamp(4);
#pragma unroll
    for (int q = 0; q < ROWS; ++q)
      if (lev[q] >= 0 && repl[q]) {
        // (diagnostic: scal[22] / [23] around the first replicated row's set-up, read after the loads it depends on)
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0) : "r"(lev[q]));
        numeric(q);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1) : "d"(d[q]), "d"(r[q]));
        if (!CLUSTER && lev[q] == l_sync) { a.scal[22] = (double)(t0 & 0xffffffffffull); a.scal[23] = (double)(t1 & 0xffffffffffull); }
      }
    lsync();
    stamp(5);
    return false;
  };
  // contribution of top child tc to its parent: d -= w (w/d_c), r -= (w/d_c) r_c
  auto fold = [&](int q, int32_t tc) {
    const double wc = rd(sw, tc);
    const double t = wc * rd(sd, tc);
    d[q] -= wc * t;
    r[q] -= t * rd(sr, tc);
  };
  for (int l = pl.l_split; l < pl.nl; ++l) {               // up
    if (ncls && l == l_sync) { if (exchange()) return true; }
    if (l == pl.l_split) continue;                          // the lowest top level has no top children
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      if (lev[q] == l && nct[q] > 0) {
        fold(q, c0[q]);
        if (nct[q] > 1) fold(q, c1[q]);
        if (nct[q] > 2) {                                    // rare: third and later top children, from the plan
          int seen = 0;
          for (int32_t k = k0[q]; k < k1[q]; ++k) {
            const int32_t tc = ncls ? top_child(q, pl.ch_idx[k]) : pl.top_of[pl.ch_idx[k]];
            if (tc >= 0 && ++seen > 2) fold(q, tc);
          }
        }
        const int32_t li = tid + q * T;
        sd[li] = gnx_rcp(d[q]); sr[li] = r[q];
      }
    }
    if (ncls && l == l_sync) stamp(13);                     // (diagnostic: scal[21], first replicated level folded)
    lsync();
  }
  if (ncls && l_sync >= pl.nl) { if (exchange()) return true; }     // (no replicated raked node)
  if (tick) tick[1] = clock64();
  stamp(6);
  for (int l = pl.nl - 1; l >= pl.l_split; --l) {           // down: sr[] turns into the solution
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      if (lev[q] == l) {
        const int32_t li = tid + q * T;
        double ri = r[q];
        if (ptop[q] >= 0) ri -= w[q] * rd(sr, ptop[q]);
        const double xi = ri * sd[li];
        sr[li] = xi;
        a.P[node[q]] = xi;
        if (a.lam_out) a.lam_out[node[q]] = xi * a.inv_sigma;
      }
    }
    lsync();
  }
  stamp(7);
  return false;
}

// ------------------------------------------------------------------------------------
// Blocked rake (plan.nb > 0): a block is a complete subtree of <= GNX_BLK_MAX raked nodes, one per thread.
// blk_up builds the Schur rows of its nodes, sweeps the block leaf-to-root out of shared memory (block
// barriers only) and leaves the FINAL diagonal / right-hand side of every node in a.d / a.r (the block root's are what the top part folds in); blk_down,
// two grid barriers later, turns them into the solution top-down, the root's parent read from a.P.
// sd / sr / sw: GNX_BLK_MAX doubles each.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void blk_up(const SchurArgs& a, int32_t b, double* sd, double* sr, double* sw) {
  const gnx_schur_plan_t& pl = a.plan;
  const int tid = threadIdx.x;
  const int32_t p0 = pl.blk_ptr[b], nn = pl.blk_ptr[b + 1] - p0;
  const int32_t lmin = pl.blk_lv[2 * b], lmax = pl.blk_lv[2 * b + 1];
  int32_t node = 0, lev = -1, k0 = 0, k1 = 0, c0 = -1, c1 = -1, nct = 0;
  double d = 1.0, r = 0.0;
  if (tid < nn) {
    node = pl.blk_nodes[p0 + tid];
    lev = pl.level_of[node];
    sw[tid] = pl.parent[node] >= 0 ? -a.cond[a.eidx(pl.pedge[node])] : 0.0;
    k0 = pl.ch_ptr[node]; k1 = pl.ch_ptr[node + 1];
    nct = k1 - k0;                                           // a block is a complete subtree: all children inside
    if (nct > 0) c0 = pl.blk_pos[pl.ch_idx[k0]];
    if (nct > 1) c1 = pl.blk_pos[pl.ch_idx[k0 + 1]];
    schur_row(a, node, d, r);
    sr[tid] = r;
    if (nct == 0) sd[tid] = gnx_rcp(d);
  }
  __syncthreads();
  auto fold = [&](int32_t c) {
    const double wc = sw[c];
    const double t = wc * sd[c];
    d -= wc * t;
    r -= t * sr[c];
  };
  for (int l = lmin + 1; l <= lmax; ++l) {
    if (lev == l && nct > 0) {
      fold(c0);
      if (nct > 1) fold(c1);
      for (int32_t k = k0 + 2; k < k1; ++k) fold(pl.blk_pos[pl.ch_idx[k]]);
      sd[tid] = gnx_rcp(d); sr[tid] = r;
    }
    __syncthreads();
  }
  if (tid < nn) { a.d[node] = d; a.r[node] = r; }
  __syncthreads();                                           // shared memory is reused by the next block
}

__device__ __forceinline__ void blk_down(const SchurArgs& a, int32_t b, double* sr) {
  const gnx_schur_plan_t& pl = a.plan;
  const int tid = threadIdx.x;
  const int32_t p0 = pl.blk_ptr[b], nn = pl.blk_ptr[b + 1] - p0;
  const int32_t lmin = pl.blk_lv[2 * b], lmax = pl.blk_lv[2 * b + 1];
  int32_t node = 0, lev = -1, par = -1, ppos = -1;
  double d = 1.0, r = 0.0, w = 0.0;
  if (tid < nn) {
    node = pl.blk_nodes[p0 + tid];
    lev = pl.level_of[node];
    par = pl.parent[node];
    if (par >= 0) { ppos = pl.blk_pos[par]; w = -a.cond[a.eidx(pl.pedge[node])]; }
    d = __ldcg(a.d + node); r = __ldcg(a.r + node);
  }
  for (int l = lmax; l >= lmin; --l) {
    if (lev == l) {
      double ri = r;
      if (par >= 0) ri -= w * (ppos >= 0 ? sr[ppos] : __ldcg(a.P + par));
      const double xi = ri / d;
      sr[tid] = xi;
      a.P[node] = xi;
      if (a.lam_out) a.lam_out[node] = xi * a.inv_sigma;
    }
    __syncthreads();
  }
}

// Stopping rule of the core PCG: ||r||^2 <= rtol^2 ||b||^2 -- relative to the INITIAL residual when b = 0 (a warm start
// from a non-zero state with a zero right-hand side must still come down to x = 0, and rtol^2 * 0 could never be met) --
// and a stagnation exit: no 10 % gain in 2000 iterations means the recurrence has hit its floor; spinning on to
// max_iter would only burn grid barriers (the caller sees "not converged").  All CTAs evaluate it on identical bits.
struct PcgStop {
  double tol2, best;
  int it_best;
  __device__ __forceinline__ PcgStop(double rtol, double bb, double rr0) : tol2(rtol * rtol * (bb > 0.0 ? bb : rr0)), best(rr0), it_best(0) {}
  __device__ __forceinline__ bool done(double rr, int it) {
    if (rr <= tol2) return true;
    if (rr < 0.9 * best) { best = rr; it_best = it; }
    return it - it_best > 2000;
  }
};

// ------------------------------------------------------------------------------------
// Cluster-resident Jacobi-PCG on the core (SCH_CLUSTER with an ELL plan): the CG of a cyclic
// network is a chain of hundreds of tiny dependent steps, so what matters is the latency of one
// iteration, not bandwidth.  The solve runs as ONE thread-block cluster of 16 CTAs; the search
// direction p and the preconditioned residual z -- the only vectors a row reads from its
// neighbours -- live in distributed shared memory (a CTA owns ell_R consecutive rows; a neighbour
// is addressed as (CTA rank, offset) and read over DSMEM), the matrix is in ELL layout (fixed
// width, independent coalesced loads, no row-pointer chain), dot products are combined through
// per-CTA shared-memory slots read by every CTA in the same order (bitwise identical results), and
// the two barriers of an iteration are hardware cluster barriers.
// Shared memory: z[R] | p0[R] | p1[R] | slots[8].
// ------------------------------------------------------------------------------------
constexpr int CL_RMAX = 5120;
constexpr size_t CL_SMEM = (3 * CL_RMAX + 8) * sizeof(double);

template <int K>
__device__ __forceinline__ void cluster_allreduce(double (&v)[K], double* slot, double* sh, cg::cluster_group& cluster) {
  block_allreduce<K>(v, sh);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) slot[k] = v[k];
  }
  cluster.sync();
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = 0.0;
  for (int c = 0; c < CL_CTAS; ++c) {
    const double* rs = cluster.map_shared_rank(slot, c);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] += rs[k];
  }
}

template <class Gather>
__device__ __forceinline__ void core_pcg_cluster(const SchurArgs& a, double* sm, double (*red)[3 * 32], Gather gather,
                                                 int& it, double& rr_out, double& bb_out) {
  cg::cluster_group cluster = cg::this_cluster();
  const gnx_schur_plan_t& pl = a.plan;
  const int32_t R = pl.ell_R, W = pl.ell_w, nc = pl.nc;
  const int rank = (int)cluster.block_rank();
  double* zs = sm;
  double* pbuf[2] = {sm + R, sm + 2 * R};
  double* slotA = sm + 3 * R;          // 3 values
  double* slotB = sm + 3 * R + 4;      // 1-2 values
  const int32_t row0 = rank * R;
  const int32_t nloc = max(0, min(R, nc - row0));
  double acc[3] = {0.0, 0.0, 0.0};
  for (int32_t li = threadIdx.x; li < nloc; li += blockDim.x) {
    const int32_t i = row0 + li;
    const int32_t node = pl.core_nodes[i];
    gather(node);
    const double di = a.d[node];
    const double xi = a.warm ? a.P[node] : 0.0;
    double ax = di * xi;
    for (int k = 0; k < W; ++k) {
      const int32_t e = pl.ell_edge[(int64_t)k * nc + i];
      double v = 0.0;
      if (e >= 0) {
        v = -a.cond[a.eidx(e)];
        if (a.warm) {
          const int32_t pc = pl.ell_col[(int64_t)k * nc + i];
          ax += v * a.P[pl.core_nodes[(pc >> 16) * R + (pc & 0xffff)]];
        }
      }
      a.cval[(int64_t)k * nc + i] = v;
    }
    const double bi = a.r[node];
    const double ri = bi - ax;
    const double zi = ri / di;
    a.dg[i] = di; a.xc[i] = xi; a.rc[i] = ri;
    zs[li] = zi; pbuf[0][li] = 0.0; pbuf[1][li] = 0.0;
    acc[0] += ri * zi; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  cluster_allreduce<3>(acc, slotA, red[0], cluster);
  double rzc = acc[0], rzp = acc[0], rr = acc[1];
  const double bb = acc[2];
  PcgStop stop(a.rtol, bb, rr);
  for (; it < a.max_iter; ++it) {
    if (stop.done(rr, it)) break;
    const double beta = (it == 0 || rzp == 0.0) ? 0.0 : rzc / rzp;
    const double* p_old = pbuf[it & 1];
    double* p_new = pbuf[(it & 1) ^ 1];
    double pw[1] = {0.0};
    for (int32_t li = threadIdx.x; li < nloc; li += blockDim.x) {
      const int32_t i = row0 + li;
      const double pi = zs[li] + beta * p_old[li];
      double wi = a.dg[i] * pi;
      for (int k = 0; k < W; ++k) {
        const int32_t pc = pl.ell_col[(int64_t)k * nc + i];
        if (pc >= 0) {
          const double v = a.cval[(int64_t)k * nc + i];
          const int rk = pc >> 16, off = pc & 0xffff;
          const double* zr = cluster.map_shared_rank(zs, rk);
          const double* pr = cluster.map_shared_rank(p_old, rk);
          wi += v * (zr[off] + beta * pr[off]);
        }
      }
      p_new[li] = pi;
      a.wc[i] = wi;
      pw[0] += pi * wi;
    }
    cluster_allreduce<1>(pw, slotB, red[1], cluster);
    const double alpha = (pw[0] != 0.0) ? rzc / pw[0] : 0.0;
    double a2[2] = {0.0, 0.0};
    for (int32_t li = threadIdx.x; li < nloc; li += blockDim.x) {
      const int32_t i = row0 + li;
      a.xc[i] += alpha * p_new[li];
      const double ri = a.rc[i] - alpha * a.wc[i];
      const double zi = ri / a.dg[i];
      a.rc[i] = ri; zs[li] = zi;
      a2[0] += ri * zi; a2[1] += ri * ri;
    }
    cluster_allreduce<2>(a2, slotA, red[0], cluster);
    rzp = rzc; rzc = a2[0]; rr = a2[1];
  }
  rr_out = rr; bb_out = bb;
  for (int32_t li = threadIdx.x; li < nloc; li += blockDim.x) {
    const int32_t i = row0 + li;
    const int32_t node = pl.core_nodes[i];
    const double xi = a.xc[i];
    a.P[node] = xi;
    if (a.lam_out) a.lam_out[node] = xi * a.inv_sigma;
  }
  cluster.sync();      // nobody leaves (or reuses shared memory) while a neighbour may still read it
}

// ------------------------------------------------------------------------------------
// Multilevel-additive PCG on a large core (cooperative grid, plan.na CTAs, CTA J owns the rows of coarse
// aggregate J).  Jacobi alone needs O(diameter) iterations on a lattice-like core (2476 on the 200 x 200
// honeycomb); the preconditioner
//     M^-1 = D^-1 + W1 diag(W1^T S W1)^-1 W1^T + W2 (W2^T S W2)^-1 W2^T
// over na * ns fine and na <= 128 coarse aggregates (W = their indicator vectors; a coarse aggregate is the
// union of ns fine ones) removes the smooth error Jacobi cannot reach: 579 iterations with the coarse level
// alone, ~340 with both.  It costs no extra grid barrier:
//  * the coarse matrix E = W2^T S W2 (na x na) is summed once per solve over fixed lists (row J by CTA J),
//    and EVERY CTA inverts it by Gauss-Jordan in its own shared memory -- the same arithmetic in the same
//    order, so all CTAs hold the same bits and the inverse never travels; the fine level only needs the
//    diagonal of W1^T S W1 (one number per fine aggregate, summed by one warp);
//  * per iteration CTA J publishes the sums of r over its fine aggregates (one warp each) and their total
//    together with its dot-product partials (the barrier of the existing all-reduce), every CTA then forms
//    y1 = s1 / dE, y2 = E^-1 s2 and s1.y1 + s2.y2 redundantly, and the preconditioned residual is never
//    stored whole: z_j = r_j/d_j + y1[fine(j)] + y2[coarse(j)] is rebuilt by whoever reads it.
// Shared memory: E^-1 [na][na+1] | s2 | y2 | pivot row | pivot column | 8 x na partials | y1 | 1/dE.
// ------------------------------------------------------------------------------------
constexpr int AGG_LD = GNX_AGG_MAX + 1;          // odd leading dimension: column walks are conflict-free
constexpr int TL_RSH = 2048;                     // rows of an aggregate whose residual is also kept in shared memory
constexpr size_t TL_SMEM = ((size_t)GNX_AGG_MAX * AGG_LD + 12 * GNX_AGG_MAX + 2 * GNX_SUB_MAX + TL_RSH) * sizeof(double);

// All-reduce over the co-resident CTAs of the multilevel PCG WITHOUT a grid barrier: every CTA publishes its K partials
// in its own 128-byte line and reads everybody's; a word says by itself whether it is the one a reader is waiting
// for -- the two low mantissa bits of partial 0 carry a tag (call / 2 mod 3; 3 = never published) -- so there is no
// flag, no counter and no reset, and the data arrive with the synchronisation instead of one L2 round trip after it
// (grid.sync() + reading the partials: 2.8 us of a 13 us iteration at CTA 0).  Two sets of lines alternate:
// to publish call c + 2 a CTA must have finished call c + 1, which needs everyone's c + 1 partial, published after
// its owner finished reading call c.  The tag perturbs a dot product by <= 3 ulp, the same in every CTA (all read the
// same words: control flow stays grid-uniform).  RELEASE: everything the CTA wrote before (vectors its neighbours read
// next, partials 1..K-1, aggregate sums) is visible to whoever sees partial 0 -- the barrier semantics the second
// reduction point of an iteration needs; the first one publishes nothing but its number.  Measured gain: 2 % (honeycomb
// 200x200: 4.65 against 4.74 ms) -- most of the "barrier time" is waiting for the slowest CTA.  (An epoch-flag barrier in
// place of grid.sync(), partials read afterwards, was tried first: no gain -- and 1.5x SLOWER with the 128 flags packed
// into eight cache lines, the polls serialise on a few L2 slices; hence one line per CTA here.)
template <int K, bool RELEASE>
__device__ __forceinline__ void tl_allreduce(double (&v)[K], double* arena, int nblk, unsigned& call, double* sh) {
  static_assert(RELEASE || K == 1, "without release semantics only the tagged word itself is ordered");
  block_allreduce<K>(v, sh);
  const unsigned set = call & 1u;
  const unsigned long long tag = (call >> 1) % 3u;
  if (threadIdx.x == 0) {
    double* const p = arena + ((size_t)set * GNX_AGG_MAX + blockIdx.x) * 16;
#pragma unroll
    for (int k = 1; k < K; ++k) p[k] = v[k];
    const unsigned long long w = ((unsigned long long)__double_as_longlong(v[0]) & ~3ull) | tag;
    if (RELEASE) asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
    else asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(w) : "memory");
  }
  double t[K];
#pragma unroll
  for (int k = 0; k < K; ++k) t[k] = 0.0;
  for (int i = threadIdx.x; i < nblk; i += blockDim.x) {
    double* const p = arena + ((size_t)set * GNX_AGG_MAX + i) * 16;
    unsigned long long w;
    unsigned spins = 0;
    for (;;) {
      if (RELEASE) asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
      else asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
      if ((w & 3ull) == tag) break;
      if (++spins > (1u << 28)) __trap();       // a CTA is missing: fail instead of hanging the GPU
    }
    t[0] += __longlong_as_double((long long)(w & ~3ull));
#pragma unroll
    for (int k = 1; k < K; ++k) t[k] += __ldcg(p + k);
  }
  __syncthreads();                              // sh reuse; and: the whole CTA is behind thread i's acquire
  block_allreduce<K>(t, sh);
#pragma unroll
  for (int k = 0; k < K; ++k) v[k] = t[k];
  ++call;
}

template <class Gather>
__device__ __forceinline__ void core_pcg_twolevel(const SchurArgs& a, double* sm, double (*red)[3 * 32], Gather gather,
                                                  int& it, double& rr_final, double& bb_final) {
  const gnx_schur_plan_t& pl = a.plan;
  const int na = pl.na, T = blockDim.x, tid = threadIdx.x, J = blockIdx.x;     // gridDim.x == na
  double* const Ei = sm;
  double* const s_sh = Ei + GNX_AGG_MAX * AGG_LD;
  double* const y_sh = s_sh + GNX_AGG_MAX;
  double* const rk = y_sh + GNX_AGG_MAX;
  double* const ck = rk + GNX_AGG_MAX;
  double* const part8 = ck + GNX_AGG_MAX;
  double* const y1_sh = part8 + 8 * GNX_AGG_MAX;
  double* const dEi_sh = y1_sh + GNX_SUB_MAX;
  double* const rsh = dEi_sh + GNX_SUB_MAX;       // [TL_RSH] my rows' residual (when they fit: no re-read from L2)
  const int32_t r0 = pl.agg_ptr[J], r1 = pl.agg_ptr[J + 1];
  const bool r_in_smem = r1 - r0 <= TL_RSH;
  // tagged all-reduce (tl_allreduce): my two lines start as "never published"; the grid barrier of the first
  // (cooperative-groups) all-reduce below orders this before anybody's first poll
  double* const arena = a.part[3];                // part[3], part[4]: 2 sets x GNX_AGG_MAX lines of 16 doubles
  const bool tagged = a.pcg_grid_sync == 0 && na <= GNX_AGG_MAX;
  if (tagged && tid < 2) {
    const unsigned long long never = 3ull;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(arena + ((size_t)tid * GNX_AGG_MAX + J) * 16), "l"(never) : "memory");
  }
  unsigned arc = 0;
  const int ns = pl.ns > 1 ? pl.ns : 1, nf = na * ns, lg = __ffs(ns) - 1;       // fine aggregates per coarse one
  cg::grid_group grid = cg::this_grid();
  // sums of r over my fine aggregates and their total -> a.s1, a.sE; with_dE: also the fine diagonal of
  // W1^T S W1.  All 32 warps work: 32/ns warps share a fine aggregate (lane- and warp-strided rows), partials
  // are combined in a fixed order.  Block-wide.
  auto sub_sums = [&](bool with_dE) {
    __syncthreads();                               // this CTA's rc (dg, cval) are written
    const int w = tid >> 5, lane = tid & 31, nw = T >> 5;
    const int wps = nw / ns;                       // warps per fine aggregate (ns <= 32)
    if (w < wps * ns) {
      const int32_t f = J * ns + w / wps;
      const int32_t i0 = ns > 1 ? pl.sub_ptr[f] : r0, i1 = ns > 1 ? pl.sub_ptr[f + 1] : r1;
      double sr = 0.0, sd = 0.0;
      for (int32_t i = i0 + (w % wps) * 32 + lane; i < i1; i += 32 * wps) {
        sr += (r_in_smem && !with_dE) ? rsh[i - r0] : __ldcg(a.rc + i);
        if (with_dE) {
          double t = __ldcg(a.dg + i);
          for (int32_t k = pl.crp[i]; k < pl.crp[i + 1]; ++k)
            if (pl.col_agg[k] == f) t += __ldcg(a.cval + k);
          sd += t;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sr += __shfl_xor_sync(GNX_FULL, sr, o);
        sd += __shfl_xor_sync(GNX_FULL, sd, o);
      }
      if (lane == 0) { part8[w] = sr; part8[64 + w] = sd; }
    }
    __syncthreads();
    if (tid < ns) {
      double sr = 0.0, sd = 0.0;
      for (int q = 0; q < wps; ++q) { sr += part8[tid * wps + q]; sd += part8[64 + tid * wps + q]; }
      part8[128 + tid] = sr;
      a.s1[J * ns + tid] = sr;
      if (with_dE) a.dE[J * ns + tid] = sd;
    }
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w2 = 0; w2 < ns; ++w2) t += part8[128 + w2];
      a.sE[J] = t;
    }
  };

  // ---- raked children, matrix values, residual of the initial guess; row sums for the coarse diagonal
  double acc[3] = {0.0, 0.0, 0.0};
  double v2[2] = {0.0, 0.0};                     // sum of r, sum of row sums over my aggregate
  for (int32_t i = r0 + tid; i < r1; i += T) {
    const int32_t node = pl.core_nodes[i];
    gather(node);
    const double di = a.d[node];
    const double xi = a.warm ? a.P[node] : 0.0;
    double ax = di * xi, rs = di;
    for (int32_t k = pl.crp[i]; k < pl.crp[i + 1]; ++k) {
      const double v = -a.cond[a.eidx(pl.cedge[k])];
      a.cval[k] = v;
      rs += v;
      if (a.warm) ax += v * a.P[pl.core_nodes[pl.ccol[k]]];
    }
    const double bi = a.r[node];
    const double ri = bi - ax;
    const double zi = ri / di;
    a.dg[i] = di; a.xc[i] = xi; a.rc[i] = ri; a.zc[i] = zi; a.p0[i] = 0.0; a.p1[i] = 0.0;
    acc[0] += ri * zi; acc[1] += ri * ri; acc[2] += bi * bi;
    v2[0] += ri; v2[1] += rs;
  }
  block_allreduce<2>(v2, red[1]);                // (its barrier also publishes cval inside the CTA)
  // row J of E: off-diagonal = sum of the entries that couple J to K, diagonal = sum of my row sums - those
  for (int32_t K = tid; K < na; K += T) {
    double e = 0.0;
    if (K != J)
      for (int32_t c = pl.cut_ptr[J * na + K]; c < pl.cut_ptr[J * na + K + 1]; ++c) e += a.cval[pl.cut_k[c]];
    rk[K] = e;
  }
  __syncthreads();
  if (tid == 0) {
    double off = 0.0;
    for (int32_t K = 0; K < na; ++K) if (K != J) off += rk[K];
    rk[J] = v2[1] - off;
  }
  __syncthreads();
  for (int32_t K = tid; K < na; K += T) a.Eg[J * na + K] = rk[K];
  sub_sums(true);
  {
    double* const part[3] = {a.part[0], a.part[1], a.part[2]};
    grid_allreduce<SCH_GRID, 3>(acc, part, na, red[0]);          // its grid barrier publishes Eg, sE, s1 and dE
  }
  // ---- every CTA: E -> E^-1 in shared memory (in-place Gauss-Jordan; E is SPD, no pivoting).
  // Thread t owns column t & 127 of rows (t >> 7) + 8 u.
  const int cj = tid & (GNX_AGG_MAX - 1), ci0 = tid >> 7;
  constexpr int RSTEP = SCH_T / GNX_AGG_MAX;
  if (cj < na)
    for (int i = ci0; i < na; i += RSTEP) Ei[i * AGG_LD + cj] = __ldcg(a.Eg + i * na + cj);
  if (ns > 1)
    for (int f = tid; f < nf; f += T) dEi_sh[f] = 1.0 / __ldcg(a.dE + f);
  __syncthreads();
  for (int k = 0; k < na; ++k) {
    if (tid < na) { rk[tid] = Ei[k * AGG_LD + tid]; ck[tid] = Ei[tid * AGG_LD + k]; }
    __syncthreads();
    const double ip = 1.0 / rk[k];
    if (cj < na) {
      const double rkj = rk[cj] * ip;
      for (int i = ci0; i < na; i += RSTEP) {
        double v;
        if (i == k) v = (cj == k) ? ip : rkj;
        else if (cj == k) v = -ck[i] * ip;
        else v = Ei[i * AGG_LD + cj] - ck[i] * rkj;
        Ei[i * AGG_LD + cj] = v;
      }
    }
    __syncthreads();
  }
  // y = E^-1 s (all of it, every CTA), returns s.y
  auto coarse = [&]() -> double {
    for (int32_t K = tid; K < na; K += T) s_sh[K] = __ldcg(a.sE + K);
    double fine = 0.0;                            // s1 . y1 (my share)
    if (ns > 1)
      for (int f = tid; f < nf; f += T) {
        const double sf = __ldcg(a.s1 + f);
        const double yf = sf * dEi_sh[f];
        y1_sh[f] = yf;
        fine += sf * yf;
      }
    __syncthreads();
    double pv = 0.0;
    if (cj < na) {
      const int c1 = min(na, ci0 * 16 + 16);
      for (int c = ci0 * 16; c < c1; ++c) pv += Ei[cj * AGG_LD + c] * s_sh[c];
    }
    part8[ci0 * GNX_AGG_MAX + cj] = pv;
    __syncthreads();
    double sg[1] = {fine};
    if (tid < na) {
      double y = 0.0;
#pragma unroll
      for (int q = 0; q < RSTEP; ++q) y += part8[q * GNX_AGG_MAX + tid];
      y_sh[tid] = y;
      sg[0] += y * s_sh[tid];
    }
    block_allreduce<1>(sg, red[1]);
    __syncthreads();                              // red[1] is the next reduction's buffer, too
    return sg[0];
  };
  double rr = acc[1];
  const double bb = acc[2];
  PcgStop stop(a.rtol, bb, rr);
  double rzc = acc[0] + coarse(), rzp = rzc;
  for (; it < a.max_iter; ++it) {
    if (stop.done(rr, it)) break;
    const double beta = (it == 0 || rzp == 0.0) ? 0.0 : rzc / rzp;
    const double* p_old = (it & 1) ? a.p1 : a.p0;
    double* p_new = (it & 1) ? a.p0 : a.p1;
    const double yJ = y_sh[J];
    double pw[1] = {0.0};
    for (int32_t i = r0 + tid; i < r1; i += T) {
      const double yi = ns > 1 ? yJ + y1_sh[pl.row_sub[i]] : yJ;
      const double pi = (a.zc[i] + yi) + beta * p_old[i];
      double wi = a.dg[i] * pi;
      for (int32_t k = pl.crp[i]; k < pl.crp[i + 1]; ++k) {
        const int32_t j = pl.ccol[k];
        const int32_t cs = pl.col_agg[k];                     // fine aggregate of the column
        const double yj = ns > 1 ? y_sh[cs >> lg] + y1_sh[cs] : y_sh[cs];
        wi += a.cval[k] * ((__ldcg(a.zc + j) + yj) + beta * __ldcg(p_old + j));
      }
      p_new[i] = pi;
      a.wc[i] = wi;
      pw[0] += pi * wi;
    }
    if (tagged) tl_allreduce<1, false>(pw, arena, na, arc, red[1]);
    else {
      double* const part[1] = {a.part[3]};      // (the arena of the tagged path: unused in this mode)
      grid_allreduce<SCH_GRID, 1>(pw, part, na, red[1]);
    }
    const double alpha = (pw[0] != 0.0) ? rzc / pw[0] : 0.0;
    double v3[2] = {0.0, 0.0};                    // r.D^-1 r, r.r
    for (int32_t i = r0 + tid; i < r1; i += T) {
      a.xc[i] += alpha * p_new[i];
      const double ri = a.rc[i] - alpha * a.wc[i];
      const double zi = ri / a.dg[i];
      a.rc[i] = ri; a.zc[i] = zi;
      if (r_in_smem) rsh[i - r0] = ri;
      v3[0] += ri * zi; v3[1] += ri * ri;
    }
    double t2[2] = {v3[0], v3[1]};
    if (tagged) {
      sub_sums(false);                            // (a.s1, a.sE: published by the release of the all-reduce below)
      tl_allreduce<2, true>(t2, arena, na, arc, red[0]);
    } else {
      block_allreduce<2>(v3, red[0]);
      if (tid == 0) { a.part[0][J] = v3[0]; a.part[1][J] = v3[1]; }
      sub_sums(false);
      grid.sync();
      t2[0] = t2[1] = 0.0;
      for (int i = tid; i < na; i += T) { t2[0] += __ldcg(a.part[0] + i); t2[1] += __ldcg(a.part[1] + i); }
      __syncthreads();                            // red[0] reuse
      block_allreduce<2>(t2, red[0]);
    }
    rzp = rzc; rzc = t2[0] + coarse(); rr = t2[1];
  }
  rr_final = rr; bb_final = bb;
  for (int32_t i = r0 + tid; i < r1; i += T) {
    const int32_t node = pl.core_nodes[i];
    const double xi = a.xc[i];
    a.P[node] = xi;
    if (a.lam_out) a.lam_out[node] = xi * a.inv_sigma;
  }
}

template <int MODE> struct SchThreads { static constexpr int value = SCH_T; };
template <> struct SchThreads<SCH_CLUSTER> { static constexpr int value = CL_T; };

// The whole Schur solve as a device function: the body of schur_solve_kernel<MODE>, and -- MODE == SCH_SINGLE,
// executed by CTA 0 -- the middle phase of the persistent step kernel (step_resident_kernel below).
// top_sh: 3 * GNX_TOP_MAX doubles of shared memory, red: 2 x 96 doubles.
template <int MODE, class WaitFn = SchNoWait>
__device__ __forceinline__ void schur_body(const SchurArgs& a, double* top_sh, double (*red)[3 * 32], WaitFn wait_fn = WaitFn()) {
  const gnx_schur_plan_t& pl = a.plan;
  // wait_fn (see top_solve): deferred into the top sweep when nothing before it reads data of this step -- a tree
  // that fits the one-CTA sweep whole -- and called right here otherwise
  const bool defer_wait = MODE == SCH_SINGLE && pl.top_rec != nullptr && pl.nc == 0 && pl.nb == 0 && pl.l_split == 0 &&
                          pl.nt > 0 && a.phase != 2 && !(a.wait_flags && !pl.ncls);
  if (!defer_wait) wait_fn();
  const int nblk = MODE == SCH_SINGLE ? 1 : gridDim.x;       // (one CTA of a larger grid in the persistent step kernel)
  const int32_t stride = nblk * blockDim.x;
  const int32_t t0 = (MODE == SCH_SINGLE ? 0 : blockIdx.x * blockDim.x) + threadIdx.x;
  auto gsync = [&]() { sch_sync<MODE>(); };
  long long tick[4] = {0, 0, 0, 0};               // SM cycles of block 0: phase 0 | top prologue | up | down (diagnostic, scal[5..7])
  const long long tick0 = clock64();
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); a.scal[8] = (double)(t & 0xffffffffffull); }
  // subtree-partitioned network: local sweeps first, ONE exchange at level l_sync, replicated nodes after it
  const int32_t* const ncls = pl.ncls;
  const bool part = ncls != nullptr;
  const int l_sync = part ? pl.l_sync : pl.nl;
  if (a.wait_flags && !part) {                    // replicated solve: the edge scalars of ALL ranks, over NVLink
    if (blockIdx.x == 0) {
      sch_flag_exchange(a);
    } else if ((int)threadIdx.x < a.wait_world) {
      const long long* f = a.wait_flags + threadIdx.x;
      long long v;
      unsigned spins = 0;
      for (;;) {
        asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
        if (v >= a.wait_epoch) break;
        __nanosleep(spins < 4096 ? 32 : 2000);
        if (++spins > (1u << 25)) __trap();
      }
    }
    __syncthreads();
  }

  // parent i absorbs its raked children: d_i -= w^2/d_c, r_i -= w r_c/d_c  (w = -cond of the link)
  auto gather = [&](int32_t i) {
    const int32_t c0 = pl.ch_ptr[i], c1 = pl.ch_ptr[i + 1];
    if (c0 == c1) return;
    double di = a.d[i], ri = a.r[i];
    for (int32_t k = c0; k < c1; ++k) {
      const int32_t c = pl.ch_idx[k];
      const double w = -a.cond[a.eidx(pl.pedge[c])];
      const double t = w / __ldcg(a.d + c);
      di -= w * t;
      ri -= t * __ldcg(a.r + c);
    }
    a.d[i] = di; a.r[i] = ri;
  };
  auto up = [&](int l, int32_t first, int32_t step) {
    for (int32_t idx = pl.lvl_ptr[l] + first; idx < pl.lvl_ptr[l + 1]; idx += step) gather(pl.lvl_nodes[idx]);
  };
  auto down = [&](int l, int32_t first, int32_t step) {
    for (int32_t idx = pl.lvl_ptr[l] + first; idx < pl.lvl_ptr[l + 1]; idx += step) {
      const int32_t i = pl.lvl_nodes[idx];
      const int32_t p = pl.parent[i];
      double ri = a.r[i];
      if (p >= 0) ri += a.cond[a.eidx(pl.pedge[i])] * __ldcg(a.P + p);
      const double xi = ri / a.d[i];
      a.P[i] = xi;
      if (a.lam_out) a.lam_out[i] = xi * a.inv_sigma;
    }
  };

  auto row_to_global = [&](int32_t bi) {
    double diag, rhs;
    schur_row(a, bi, diag, rhs);
    a.d[bi] = diag; a.r[bi] = rhs;
  };
  // ---- phase 0: Schur diagonal and right-hand side (one thread per bifurcation, gather)
  // (blocked rake: the blocks build their own rows, CTA 0 those of the top part -- no barrier in between;
  //  partitioned: only my local nodes now, the replicated ones after the exchange)
  const bool top_rows_later = pl.top_rec != nullptr && pl.nc == 0 && MODE != SCH_CLUSTER;   // top_solve builds them from the records
  if (pl.nb > 0 || a.phase == 2) {
    if (blockIdx.x == 0 && pl.nc == 0 && !top_rows_later)
      for (int32_t ti = threadIdx.x; ti < pl.nt; ti += blockDim.x) {
        const int32_t bi = pl.top_nodes[ti];
        if (!part || !(ncls[bi] & 1)) row_to_global(bi);
      }
  } else if (part || top_rows_later) {
    // (levels are contiguous in lvl_nodes: everything below the exchange level / below the top part)
    const int32_t lim = pl.lvl_ptr[top_rows_later ? (pl.l_split < l_sync ? pl.l_split : l_sync) : l_sync];
    for (int32_t idx = t0; idx < lim; idx += stride) row_to_global(pl.lvl_nodes[idx]);
    if (lim > 0) gsync();
  } else {
    for (int32_t bi = t0; bi < pl.B; bi += stride) row_to_global(bi);
    gsync();
  }

  // ---- phase 1: rake up (level 0 has no children by construction)
  if (pl.nb > 0) {                                 // blocked rake: whole subtrees per CTA, ONE grid barrier
    if (a.phase != 2)
      for (int32_t b = blockIdx.x; b < pl.nb; b += nblk) blk_up(a, b, top_sh, top_sh + GNX_BLK_MAX, top_sh + 2 * GNX_BLK_MAX);
    gsync();
  } else if (a.phase != 2) {
    const int lend = pl.l_split < l_sync ? pl.l_split : l_sync;       // (a forest's top part starts at or below l_sync)
    for (int l = 1; l < lend; ++l) { up(l, t0, stride); gsync(); }
  }
  if (part && pl.nc > 0) {
    // a core exists: no top part; the exchange sits between the local and the replicated levels of the grid sweep
    if (blockIdx.x == 0) sch_export_roots(a, a.phase != 2, a.phase == 0);
    if (a.phase == 1) return;
    if (a.phase == 0) gsync();
    for (int32_t idx = pl.lvl_ptr[l_sync] + t0; idx < pl.lvl_ptr[pl.nl]; idx += stride) row_to_global(pl.lvl_nodes[idx]);
    for (int32_t i = t0; i < pl.nc; i += stride) row_to_global(pl.core_nodes[i]);
    gsync();
    for (int l = l_sync; l < pl.nl; ++l) { up(l, t0, stride); gsync(); }
  }
  double rr_final = 0.0, bb_final = 0.0;
  int it = 0;
  if (pl.nc == 0) {
    if (MODE == SCH_CLUSTER) {                  // whole tree in one cluster's distributed shared memory
      if (pl.nt > 0) top_solve<true, GNX_TOP_MAX / CL_T>(a, top_sh, top_sh + GNX_TOP_MAX, top_sh + 2 * GNX_TOP_MAX);
    } else {
      tick[0] = clock64();
      // (one row per thread when the top part fits: half the per-thread state of the general two-row sweep --
      //  at 64 registers per thread that is the difference between spilling and not)
      if (blockIdx.x == 0 && pl.nt > 0) {
        if (defer_wait) {
          if (pl.nt <= SCH_T) top_solve<false, 1>(a, top_sh, top_sh + GNX_TOP_MAX, top_sh + 2 * GNX_TOP_MAX, tick + 1, wait_fn);
          else top_solve<false, TOP_ROWS>(a, top_sh, top_sh + GNX_TOP_MAX, top_sh + 2 * GNX_TOP_MAX, tick + 1, wait_fn);
        } else {
          if (pl.nt <= SCH_T) top_solve<false, 1>(a, top_sh, top_sh + GNX_TOP_MAX, top_sh + 2 * GNX_TOP_MAX, tick + 1);
          else top_solve<false, TOP_ROWS>(a, top_sh, top_sh + GNX_TOP_MAX, top_sh + 2 * GNX_TOP_MAX, tick + 1);
        }
      }
      tick[3] = clock64();
      if (a.phase == 1) return;                  // (partitioned, emulated ranks / NCCL: stop at the exchange)
      if (pl.l_split > 0) gsync();
    }
  } else {                                       // a core exists: l_split == nl, no top part
    // ---- phase 2: Jacobi-PCG on the core
    if (MODE == SCH_CLUSTER) {                   // launched only with an ELL plan (gnx_schur_solve)
      core_pcg_cluster(a, top_sh, red, gather, it, rr_final, bb_final);
    } else if (MODE == SCH_GRID_TL) {            // launched with na CTAs and TL_SMEM (gnx_schur_solve)
      core_pcg_twolevel(a, top_sh, red, gather, it, rr_final, bb_final);
    } else {
    double acc[3] = {0.0, 0.0, 0.0};
    for (int32_t i = t0; i < pl.nc; i += stride) {
      const int32_t node = pl.core_nodes[i];
      gather(node);
      const double di = a.d[node];
      const double xi = a.warm ? a.P[node] : 0.0;
      double ax = di * xi;
      for (int32_t k = pl.crp[i]; k < pl.crp[i + 1]; ++k) {
        const double v = -a.cond[a.eidx(pl.cedge[k])];
        a.cval[k] = v;
        if (a.warm) ax += v * a.P[pl.core_nodes[pl.ccol[k]]];
      }
      const double bi = a.r[node];
      const double ri = bi - ax;
      const double zi = ri / di;
      a.dg[i] = di; a.xc[i] = xi; a.rc[i] = ri; a.zc[i] = zi; a.p0[i] = 0.0; a.p1[i] = 0.0;
      acc[0] += ri * zi; acc[1] += ri * ri; acc[2] += bi * bi;
    }
    {
      double* const part[3] = {a.part[0], a.part[1], a.part[2]};
      grid_allreduce<MODE, 3>(acc, part, nblk, red[0]);
    }
    double rzc = acc[0], rzp = acc[0], rr = acc[1];
    const double bb = acc[2];
    PcgStop stop(a.rtol, bb, rr);
    for (; it < a.max_iter; ++it) {
      if (stop.done(rr, it)) break;
      const double beta = (it == 0 || rzp == 0.0) ? 0.0 : rzc / rzp;
      const double* p_old = (it & 1) ? a.p1 : a.p0;
      double* p_new = (it & 1) ? a.p0 : a.p1;
      double pw[1] = {0.0};
      for (int32_t i = t0; i < pl.nc; i += stride) {
        const double pi = a.zc[i] + beta * p_old[i];
        double wi = a.dg[i] * pi;
        for (int32_t k = pl.crp[i]; k < pl.crp[i + 1]; ++k) {
          const int32_t j = pl.ccol[k];
          wi += a.cval[k] * (__ldcg(a.zc + j) + beta * __ldcg(p_old + j));
        }
        p_new[i] = pi;
        a.wc[i] = wi;
        pw[0] += pi * wi;
      }
      {
        double* const part[1] = {a.part[3]};
        grid_allreduce<MODE, 1>(pw, part, nblk, red[1]);
      }
      const double alpha = (pw[0] != 0.0) ? rzc / pw[0] : 0.0;
      double a2[2] = {0.0, 0.0};
      for (int32_t i = t0; i < pl.nc; i += stride) {
        a.xc[i] += alpha * p_new[i];
        const double ri = a.rc[i] - alpha * a.wc[i];
        const double zi = ri / a.dg[i];
        a.rc[i] = ri; a.zc[i] = zi;
        a2[0] += ri * zi; a2[1] += ri * ri;
      }
      {
        // one buffer per reduction point is enough: consecutive uses of a buffer are separated by
        // the grid barrier of the other reduction point
        double* const part[2] = {a.part[0], a.part[1]};
        grid_allreduce<MODE, 2>(a2, part, nblk, red[0]);
      }
      rzp = rzc; rzc = a2[0]; rr = a2[1];
    }
    rr_final = rr; bb_final = bb;
    for (int32_t i = t0; i < pl.nc; i += stride) {
      const int32_t node = pl.core_nodes[i];
      const double xi = a.xc[i];
      a.P[node] = xi;
      if (a.lam_out) a.lam_out[node] = xi * a.inv_sigma;
    }
    }
    if (pl.nl > 0) gsync();
  }
  // ---- phase 3: rake down
  if (pl.nb > 0) {
    for (int32_t b = blockIdx.x; b < pl.nb; b += nblk) { blk_down(a, b, top_sh); __syncthreads(); }
  } else {
    for (int l = pl.l_split - 1; l >= 0; --l) { down(l, t0, stride); if (l > 0) gsync(); }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double tol2 = a.rtol * a.rtol * bb_final;
    const bool conv = rr_final <= tol2 || (bb_final == 0.0 && it < a.max_iter);
    a.scal[0] = tol2; a.scal[1] = conv ? 1.0 : 0.0; a.scal[2] = (double)it;
    if (!conv) a.scal[24] += 1.0;                 // sticky: solves since the last gnx_schur_unconverged() that did not converge
    a.scal[3] = rr_final; a.scal[4] = bb_final;
    a.scal[5] = (double)(tick[0] - tick0);        // wait for peers + phase 0 + wide rake levels
    a.scal[6] = (double)(tick[1] - tick[0]);      // top part: prologue (plan + edge scalars -> shared memory)
    a.scal[7] = (double)(tick[3] - tick[1]);      // top part: up + down sweeps
  }
}

template <int MODE>
__global__ void __launch_bounds__(SchThreads<MODE>::value, 1) schur_solve_kernel(SchurArgs a) {
  extern __shared__ double top_sh[];             // 3 * GNX_TOP_MAX doubles (dynamic, opt-in > 48 KB)
  __shared__ double red[2][3 * 32];
  gnx_pdl_wait();                                 // the edge scalars come from the elimination kernel
  schur_body<MODE>(a, top_sh, red);
}

static int g_sch_blocks_per_sm = -1, g_sch_sms = 0, g_sch_cluster_max = 0;

// Cooperative launch that is ALSO a programmatic dependent launch (the grid may become resident while the
// elimination kernel drains; it starts with gnx_pdl_wait()).  Falls back to a plain cooperative launch, once
// and for good, if the driver refuses the combination.
template <int MODE>
static cudaError_t launch_coop(const SchurArgs& a, int blocks, size_t smem, cudaStream_t st) {
  static int pdl_ok = -1;
  if (pdl_ok < 0) pdl_ok = gnx_pdl_enabled() ? 1 : 0;
  if (pdl_ok) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks); cfg.blockDim = dim3(SCH_T); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative; attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, schur_solve_kernel<MODE>, a);
    if (err == cudaSuccess) return err;
    cudaGetLastError();
    pdl_ok = 0;
  }
  SchurArgs copy = a;
  void* args[] = {&copy};
  return cudaLaunchCooperativeKernel((void*)schur_solve_kernel<MODE>, dim3(blocks), dim3(SCH_T), args, smem, st);
}

static int sch_init() {
  if (g_sch_blocks_per_sm >= 0) return GNX_OK;
  int dev = 0, coop = 0, per_sm = 0;
  GNX_CUDA(cudaGetDevice(&dev));
  GNX_CUDA(cudaDeviceGetAttribute(&g_sch_sms, cudaDevAttrMultiProcessorCount, dev));
  GNX_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  GNX_CUDA(cudaFuncSetAttribute(schur_solve_kernel<SCH_GRID>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCH_SMEM));
  GNX_CUDA(cudaFuncSetAttribute(schur_solve_kernel<SCH_GRID_TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_SMEM));
  GNX_CUDA(cudaFuncSetAttribute(schur_solve_kernel<SCH_SINGLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SCH_SMEM));
  GNX_CUDA(cudaFuncSetAttribute(schur_solve_kernel<SCH_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, CL_SMEM));
  static_assert(TL_SMEM >= SCH_SMEM, "the two-level kernel also holds the top part");
  GNX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, schur_solve_kernel<SCH_GRID>, SCH_T, SCH_SMEM));
  if (!coop || per_sm < 1) { gnx_set_error("cooperative launch unsupported on this device"); return GNX_ERR_CUDA; }
  // largest cluster this kernel can run as (16 needs the non-portable opt-in)
  g_sch_cluster_max = 0;
  if (cudaFuncSetAttribute(schur_solve_kernel<SCH_CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(16); cfg.blockDim = dim3(CL_T); cfg.dynamicSmemBytes = CL_SMEM;
    int mc = 0;
    if (cudaOccupancyMaxPotentialClusterSize(&mc, schur_solve_kernel<SCH_CLUSTER>, &cfg) == cudaSuccess) g_sch_cluster_max = mc;
  }
  cudaGetLastError();
  if (const char* e = getenv("GNX_SCHUR_NO_CLUSTER")) { if (atoi(e)) g_sch_cluster_max = 0; }   // tuning / debugging knob
  g_sch_blocks_per_sm = per_sm;
  return GNX_OK;
}

// argument checks + kernel argument block of a Schur solve (shared by gnx_schur_solve and the persistent step)
static int schur_args_fill(SchurArgs& a, const gnx_network_t* net, const gnx_schur_plan_t* plan, double sigma,
                           const double* cond, const double* rsrc, const double* ftot, const double* b_lam,
                           const double* b_node, double* P, double* lam_out, double* work, double rtol, int32_t max_iter,
                           int32_t warm, int32_t chunk, int32_t rank_stride, const gnx_exchange_t* ex) {
  GNX_CHECK_ARG(net && plan && cond && rsrc && ftot && P && work && sigma != 0.0 && rtol > 0 && max_iter > 0);
  GNX_CHECK_ARG(plan->B == net->B && plan->B > 0 && plan->nl >= 0 && plan->nl <= GNX_MAX_LEVELS);
  GNX_CHECK_ARG(plan->l_split >= 0 && plan->l_split <= plan->nl && plan->nc >= 0 && plan->nc <= plan->B);
  const bool cluster_top = plan->top_S > 0;       // whole tree swept by one cluster
  GNX_CHECK_ARG(plan->nt >= 0 && plan->nt <= (cluster_top ? CL_CTAS * plan->top_S : GNX_TOP_MAX));
  GNX_CHECK_ARG(!cluster_top || (plan->nc == 0 && plan->l_split == 0 && plan->top_S <= GNX_TOP_MAX &&
                                 (plan->top_S & (plan->top_S - 1)) == 0 && plan->top_S >= CL_T));
  GNX_CHECK_ARG(plan->nc == 0 || plan->l_split == plan->nl);
  GNX_CHECK_ARG(plan->nt == 0 || (plan->top_nodes && plan->top_of && plan->level_of));
  GNX_CHECK_ARG(plan->nb > 0 || plan->nt == plan->lvl_ptr[plan->nl] - plan->lvl_ptr[plan->l_split]);
  GNX_CHECK_ARG(plan->nb == 0 || (plan->nb > 0 && plan->nc == 0 && plan->top_S == 0 && plan->blk_ptr && plan->blk_nodes &&
                                  plan->blk_pos && plan->blk_lv && plan->level_of && plan->B > GNX_TOP_MAX));
  GNX_CHECK_ARG(plan->ncls == nullptr || plan->nc == 0 || plan->l_split == plan->nl);
  GNX_CHECK_ARG(plan->ell_w == 0 || (plan->ell_w <= 8 && plan->ell_col && plan->ell_edge && plan->ell_R > 0 &&
                                    plan->ell_R <= CL_RMAX && (int64_t)plan->ell_R * CL_CTAS >= plan->nc));
  GNX_CHECK_ARG(plan->nl == 0 || (plan->lvl_nodes && plan->parent && plan->pedge && plan->ch_ptr && plan->ch_idx));
  GNX_CHECK_ARG(plan->nc == 0 || (plan->core_nodes && plan->crp && plan->ccol && plan->cedge && plan->ch_ptr));
  GNX_CHECK_ARG(plan->na == 0 || (plan->na >= 2 && plan->na <= GNX_AGG_MAX && plan->na <= plan->nc && plan->agg_ptr &&
                                  plan->col_agg && plan->cut_ptr && plan->cut_k));
  GNX_CHECK_ARG(plan->na == 0 || plan->ns <= 1 ||
                (plan->ns <= 32 && (plan->ns & (plan->ns - 1)) == 0 && (int64_t)plan->na * plan->ns <= GNX_SUB_MAX &&
                 plan->sub_ptr && plan->row_sub));
  if (int rc = sch_init()) return rc;
  a.net = *net; a.plan = *plan;
  a.inv_sigma = 1.0 / sigma;
  a.cond = cond; a.rsrc = rsrc; a.ftot = ftot; a.b_lam = b_lam; a.b_node = b_node; a.P = P; a.lam_out = lam_out;
  const int64_t B = plan->B, nc = plan->nc;
  double* w = work;
  a.d = w; w += B; a.r = w; w += B;
  a.xc = w; w += nc; a.rc = w; w += nc; a.zc = w; w += nc; a.wc = w; w += nc;
  a.p0 = w; w += nc; a.p1 = w; w += nc; a.dg = w; w += nc;
  a.cval = w; w += plan->nnzc;
  for (int k = 0; k < 5; ++k) { a.part[k] = w; w += 2048; }
  a.scal = w; w += 32;                          // [0..7] status, [8..] %globaltimer stamps of CTA 0 (diagnostic)
  a.Eg = w; w += (int64_t)plan->na * plan->na; a.sE = w; w += plan->na;
  a.s1 = w; w += GNX_SUB_MAX; a.dE = w;
  a.rtol = rtol; a.max_iter = max_iter; a.warm = warm;
  a.chunk = chunk > 0 ? chunk : 2147483647; a.rank_stride = chunk > 0 ? rank_stride : 0;
  a.wait_flags = nullptr; a.wait_epoch = 0; a.wait_world = 0;
  a.nx = 0; a.xd_off = a.xr_off = 0; a.phase = 0;
  static int xpoll = -1;                        // A-B knob: GNX_XCHG_FLAGS=1 keeps the data-then-flag exchange of the roots
  if (xpoll < 0) { const char* e = getenv("GNX_XCHG_FLAGS"); xpoll = (e && atoi(e)) ? 0 : 1; }
  a.xpoll = xpoll;
  static int pcg_gs = -1;                       // A-B knob: GNX_PCG_GRID_SYNC=1 keeps cooperative_groups' grid barrier
  if (pcg_gs < 0) { const char* e = getenv("GNX_PCG_GRID_SYNC"); pcg_gs = (e && atoi(e)) ? 1 : 0; }
  a.pcg_grid_sync = pcg_gs;
  for (int k = 0; k < GNX_MAX_PEERS; ++k) a.xbuf[k] = nullptr;
  const bool subtree = ex && ex->mode == 1;
  GNX_CHECK_ARG(subtree == (plan->ncls != nullptr));
  if (subtree) {
    // global-indexed buffers [cond | rsrc | ftot (E_glob each) | d | r (B_glob each) | world flag words]; the
    // caller passes cond / rsrc / ftot = my buffer's arrays; d and r of the sweep live in it as well
    GNX_CHECK_ARG(ex->peers_host && ex->world >= 1 && ex->world <= GNX_MAX_PEERS && ex->rank >= 0 && ex->rank < ex->world &&
                  ex->peers_host[ex->rank] && ex->B_glob == plan->B && ex->E_glob == net->E && chunk == 0 &&
                  ex->phase >= 0 && ex->phase <= 2 && (ex->phase != 0 || ex->epoch > 0));
    GNX_CHECK_ARG(plan->l_sync >= 0 && plan->l_sync <= plan->nl && plan->nxl >= 0 && (plan->nxl == 0 || plan->xlist) &&
                  plan->top_S == 0 && (plan->nc > 0 || plan->l_split <= plan->l_sync));
    GNX_CHECK_ARG(cond == ex->peers_host[ex->rank]);
    const int64_t flag_off = 3 * (int64_t)ex->E_glob + 2 * (int64_t)ex->B_glob;
    a.xd_off = 3 * (int64_t)ex->E_glob; a.xr_off = a.xd_off + ex->B_glob;
    a.d = ex->peers_host[ex->rank] + a.xd_off; a.r = ex->peers_host[ex->rank] + a.xr_off;
    a.phase = ex->phase;
    for (int k = 0; k < ex->world; ++k) {               // (indexed by rank; my own slot stays NULL)
      GNX_CHECK_ARG(ex->peers_host[k] != nullptr);
      if (k != ex->rank) a.xbuf[k] = ex->peers_host[k];
      a.sig_flags[k] = reinterpret_cast<long long*>(ex->peers_host[k] + flag_off) + ex->rank;
    }
    a.nx = ex->world;
    a.wait_flags = reinterpret_cast<const long long*>(ex->peers_host[ex->rank] + flag_off);
    a.wait_epoch = ex->epoch; a.wait_world = ex->world;
  } else if (ex && ex->epoch > 0) {
    GNX_CHECK_ARG(ex->peers_host && ex->world >= 1 && ex->world <= GNX_MAX_PEERS && ex->rank >= 0 &&
                  ex->rank < ex->world && ex->peers_host[ex->rank] && ex->chunk == chunk);
    a.wait_flags = reinterpret_cast<const long long*>(ex->peers_host[ex->rank] + (int64_t)ex->world * 3 * ex->chunk);
    a.wait_epoch = ex->epoch; a.wait_world = ex->world;
    for (int k = 0; k < ex->world; ++k) {
      GNX_CHECK_ARG(ex->peers_host[k] != nullptr);
      a.sig_flags[k] = reinterpret_cast<long long*>(ex->peers_host[k] + (int64_t)ex->world * 3 * ex->chunk) + ex->rank;
    }
  }
  return GNX_OK;
}

// resid_host[0] = ||r||/||rhs|| of the core system (recurrence residual; 0 for a pure tree),
// info_host[3] = {CG iterations, converged, grid blocks (< 0: one cluster)}.  Asynchronous when both are NULL.
static int schur_report(const SchurArgs& a, int want, double* resid_host, int32_t* info_host, cudaStream_t st) {
  if (!resid_host && !info_host) return GNX_OK;              // fully asynchronous
  double scal[8];
  GNX_CUDA(cudaMemcpyAsync(scal, a.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
  GNX_CUDA(cudaStreamSynchronize(st));
  const double bb = scal[4];
  const double rel = bb > 0 ? sqrt(scal[3] / bb) : 0.0;
  if (resid_host) resid_host[0] = rel;
  const int done = scal[1] != 0.0;
  if (info_host) { info_host[0] = (int32_t)scal[2]; info_host[1] = done; info_host[2] = want; }
  if (!done) {
    gnx_set_error("schur CG: no convergence in %d iterations (rel. resid %.3e)", (int)scal[2], rel);
    return GNX_ERR_NOCONV;
  }
  return GNX_OK;
}

// number of solves on this work buffer that ended without convergence since the last call (asynchronous solves do not
// report per call); reads one double back and resets it.  work: the buffer of gnx_schur_solve, sized for `plan`.
extern "C" int gnx_schur_unconverged(const gnx_schur_plan_t* plan, double* work, void* stream) {
  if (!plan || !work) return 0;
  const int64_t nval = plan->ell_w > 0 ? (int64_t)plan->ell_w * plan->nc : 0;
  double* scal = work + 2 * (int64_t)plan->B + 7 * (int64_t)plan->nc + (plan->nnzc > nval ? plan->nnzc : nval) + 5 * 2048;
  double v = 0.0, zero = 0.0;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemcpyAsync(&v, scal + 24, sizeof(double), cudaMemcpyDeviceToHost, st) != cudaSuccess) return -1;
  if (cudaStreamSynchronize(st) != cudaSuccess) return -1;
  if (v != 0.0) cudaMemcpyAsync(scal + 24, &zero, sizeof(double), cudaMemcpyHostToDevice, st);
  return (int)v;
}

// P holds the initial guess of the core unknowns on entry when warm != 0.
extern "C" int gnx_schur_solve(const gnx_network_t* net, const gnx_schur_plan_t* plan, double sigma,
                               const double* cond, const double* rsrc, const double* ftot,
                               const double* b_lam, const double* b_node, double* P, double* lam_out, double* work,
                               double rtol, int32_t max_iter, int32_t warm, int32_t chunk, int32_t rank_stride,
                               const gnx_exchange_t* ex, double* resid_host, int32_t* info_host, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SchurArgs a;
  if (int rc = schur_args_fill(a, net, plan, sigma, cond, rsrc, ftot, b_lam, b_node, P, lam_out, work, rtol, max_iter,
                               warm, chunk, rank_stride, ex)) return rc;
  const bool cluster_top = plan->top_S > 0;
  const int64_t B = plan->B, nc = plan->nc;
  // CTAs: enough threads for the widest phase.  <= 2 CTAs' worth runs as one CTA; a system with a
  // Krylov core (hundreds of barriers) runs as ONE CLUSTER when that gives every thread <= 32 rows;
  // everything else (wide rake levels of a large tree: few barriers, lots of rows) as a cooperative grid.
  const int64_t nrows = plan->ncls ? (int64_t)plan->lvl_ptr[plan->nl] + nc : B;     // partitioned: what THIS rank sweeps
  int want = (int)((nrows + SCH_T - 1) / SCH_T);
  const int cap = g_sch_sms * g_sch_blocks_per_sm;
  if (want > cap) want = cap;
  if (want > 2048) want = 2048;
  if (cluster_top && g_sch_cluster_max < CL_CTAS) {
    gnx_set_error("plan asks for a cluster-resident tree sweep but 16-CTA clusters are unavailable");
    return GNX_ERR_CUDA;
  }
  static_assert(SCH_T == 1024 && GNX_AGG_MAX == 128 && SCH_T / GNX_AGG_MAX * 16 == GNX_AGG_MAX,
                "core_pcg_twolevel: RSTEP, the 16-column chunks of coarse() and the part8 offsets assume these sizes");
  if (plan->na > 0 && plan->na <= g_sch_sms) {
    // two-level PCG: one CTA per aggregate on the cooperative grid (every phase before it is grid-stride)
    want = plan->na;
    GNX_CUDA(launch_coop<SCH_GRID_TL>(a, want, TL_SMEM, st));
  } else if (plan->na > 0) {
    // fewer SMs than aggregates (a partitioned GPU): the aggregates are ignored, plain Jacobi-PCG on the grid
    GNX_CUDA(launch_coop<SCH_GRID>(a, want, SCH_SMEM, st));
  } else if (want <= 2 && !cluster_top) {
    GNX_CUDA(gnx_launch_dependent(schur_solve_kernel<SCH_SINGLE>, dim3(1), dim3(SCH_T), SCH_SMEM, st, a));
    GNX_LAUNCH_CHECK();
    want = 1;
  } else if (cluster_top) {
    want = CL_CTAS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(want); cfg.blockDim = dim3(CL_T); cfg.dynamicSmemBytes = CL_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = want; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GNX_CUDA(cudaLaunchKernelEx(&cfg, schur_solve_kernel<SCH_CLUSTER>, a));
    want = -want;
  } else if (g_sch_cluster_max >= CL_CTAS && nc > 0 && plan->ell_w > 0 && nc <= CL_CTAS * CL_T) {
    // a Krylov core with at most one row per thread of ONE cluster: cluster-resident PCG.  Measured on
    // B200 (tools/time_schur.py): 7.3 us/iteration against 7.8 for the cooperative grid at 7.4 k unknowns;
    // with more rows per thread the 16 SMs of a cluster lose to the full grid (80 k unknowns: 25.6 vs 7.9 us),
    // so larger cores stay on the grid path.  Either way an iteration is latency-, not bandwidth-bound.
    want = CL_CTAS;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(want); cfg.blockDim = dim3(CL_T); cfg.dynamicSmemBytes = CL_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = want; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    GNX_CUDA(cudaLaunchKernelEx(&cfg, schur_solve_kernel<SCH_CLUSTER>, a));
    want = -want;                                  // reported as negative: cluster
  } else {
    GNX_CUDA(launch_coop<SCH_GRID>(a, want, SCH_SMEM, st));
  }
  return schur_report(a, want, resid_host, info_host, st);
}

// ------------------------------------------------------------------------------------
// phase 2: back-substitution
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
edge_backsub_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                    const double* __restrict__ k_edge, double inv_sigma, const double* __restrict__ b,
                    const double* __restrict__ cond, const double* __restrict__ rsrc,
                    const double* __restrict__ P, int seg, double* __restrict__ x) {
  __shared__ double sh[32];
  const int epb = blockDim.x / seg;
  const int32_t e = blockIdx.x * epb + threadIdx.x / seg;
  const int gl = threadIdx.x & (seg - 1);
  const bool active = e < net.E;
  const int32_t ee = active ? e : net.E - 1;
  const EdgeFrame f = edge_frame(net, d, ee);
  const int32_t m = d.m;
  const int32_t bu = net.bix[f.u], bv = net.bix[f.v];
  const double Pu = bu >= 0 ? P[bu] : 0.0;
  const double Pv = bv >= 0 ? P[bv] : 0.0;
  const double a = a_edge[ee];
  const double kv = k_edge ? k_edge[ee] : 0.0;
  const double q0 = cond[ee] * (Pu - Pv + rsrc[ee]);
  double carryF = 0.0, carryD = 0.0;      // cumF at node `base`, sum of D over nodes < base
  for (int32_t base = 0; base < m; base += seg) {
    const int32_t j = base + gl;          // node j, right cell j
    const double Fj = b[f.pdof(j)] * inv_sigma;
    const double hj = net.h[f.hb + j];
    const double Gj = b[f.qdof(net.loc, j)];
    double totF, totD;
    const double inc = gnx_seg_incl_scan(Fj, seg, sh, &totF);
    const double qj = q0 + carryF + (inc - Fj);
    double mq = a * hj * (3.0 * qj + Fj) * (1.0 / 6.0) - kv * Fj / hj;           // h_j (2 q_j + q_{j+1})/6
    if (j > 0) {                          // left cell: served by L1, the neighbour thread loads the same words
      const double Fl = b[f.pdof(j - 1)] * inv_sigma;
      const double hl = net.h[f.hb + j - 1];
      mq += a * hl * (3.0 * qj - Fl) * (1.0 / 6.0) + kv * Fl / hl;               // h_{j-1}(q_{j-1} + 2 q_j)/6
    }
    const double Dj = Gj - mq;
    const double incD = gnx_seg_incl_scan(Dj, seg, sh, &totD);
    const double pj = Pu + carryD + incD;
    if (active) {
      x[f.qdof(net.loc, j)] = qj;
      x[f.pdof(j)] = pj * inv_sigma;
    }
    carryF += totF;
    carryD += totD;
  }
  if (active && gl == 0) x[f.qdof(net.loc, m)] = q0 + carryF;
}

template <int CPT, bool BULK>
__global__ void __launch_bounds__(ET_T, 4)
edge_backsub_tile_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                         const double* __restrict__ k_edge, double inv_sigma, const double* __restrict__ b,
                         const double* __restrict__ cond, const double* __restrict__ rsrc,
                         const double* __restrict__ P, double* __restrict__ x, const double* __restrict__ a_cell) {
  __shared__ TileSmem<CPT> ts;
  __shared__ double sh[32];
  const TileCtx c = tile_ctx<CPT, BULK>(net, d);
  tile_load_h<CPT, BULK>(c, net.h, ts);           // static data: in flight while the kernels before finish
  gnx_pdl_wait();                                 // P (and, transitively, b) come from the kernels before
  tile_load_b<CPT, BULK>(c, b, ts);
  double* const sF = ts.F; double* const sH = ts.H; double* const sG = ts.G;
  const int32_t m = c.m;
  const int32_t eu = net.edges[2 * c.e], ev = net.edges[2 * c.e + 1];
  const int32_t bu = net.bix[eu], bv = net.bix[ev];
  const double Pu = bu >= 0 ? P[bu] : 0.0;
  const double Pv = bv >= 0 ? P[bv] : 0.0;
  const double a = a_edge[c.e];
  const double kv = k_edge ? k_edge[c.e] : 0.0;
  const double q0 = __dmul_rn(cond[c.e], __dadd_rn(__dsub_rn(Pu, Pv), rsrc[c.e]));
  __syncthreads();
  tile_wait<CPT, BULK, true>(ts);
  const int32_t cbF = c.le * m + c.shp, cbH = c.le * m + c.shh, qb = c.le * (m + 1) + c.shq;
  int32_t ps[CPT], qs[CPT];                        // shared-memory slots of my p / q dofs
  double F[CPT], H[CPT], Gq[CPT];
  double sumF = 0.0;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int32_t kk = c.kk0 + i;
    ps[i] = cbF + gnx_gray(c.flip ? m - 1 - kk : kk);
    qs[i] = qb + __ldg(net.loc + (c.flip ? m - kk : kk));
    F[i] = c.active ? __dmul_rn(sF[ps[i]], inv_sigma) : 0.0;
    H[i] = sH[cbH + kk];
    Gq[i] = c.active ? sG[qs[i]] : 0.0;
    sumF += F[i];
  }
  double Fl = 0.0, Hl = 1.0;                       // cell left of my first node
  if (c.kk0 > 0) {
    const int32_t kl = c.kk0 - 1;
    Fl = c.active ? __dmul_rn(sF[cbF + gnx_gray(c.flip ? m - 1 - kl : kl)], inv_sigma) : 0.0;
    Hl = sH[cbH + kl];
  }
  const int segT = m / CPT;
  double totF, totD;
  const double inclF = gnx_seg_incl_scan(sumF, segT, sh, &totF);
  double cum = inclF - sumF;
  double q[CPT], D[CPT];
  double sumD = 0.0;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    q[i] = __dadd_rn(q0, cum);
    const bool has_left = (i > 0) || (c.kk0 > 0);
    const double fl = i > 0 ? F[i - 1] : Fl, hl = i > 0 ? H[i - 1] : Hl;
    double ar = a, al = a;                                                         // mass coefficient of the right / left cell
    if (a_cell) {
      const double* ac = a_cell + (int64_t)c.e * m + c.kk0 + i;
      ar = c.active ? ac[0] : 0.0;
      al = (c.active && has_left) ? ac[-1] : 0.0;
    }
    double mq = gnx_mq_right(ar, H[i], q[i], F[i]);                                // h_j (2 q_j + q_{j+1})/6
    if (has_left) mq = gnx_mq_left(mq, al, hl, q[i], fl);                          // h_{j-1}(q_{j-1} + 2 q_j)/6
    if (kv != 0.0) {                                                               // stiffness (NetworkStokes)
      mq = __dsub_rn(mq, __ddiv_rn(__dmul_rn(kv, F[i]), H[i]));
      if (has_left) mq = __dadd_rn(mq, __ddiv_rn(__dmul_rn(kv, fl), hl));
    }
    D[i] = __dsub_rn(Gq[i], mq);
    sumD += D[i];
    cum = __dadd_rn(cum, F[i]);
  }
  const double inclD = gnx_seg_incl_scan(sumD, segT, sh, &totD);
  double cd = Pu + (inclD - sumD);
  __syncthreads();                                  // every read of sF / sG is done: reuse them for x
  if (c.active) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      cd += D[i];
      sG[qs[i]] = q[i];
      sF[ps[i]] = __dmul_rn(cd, inv_sigma);
    }
    if (c.kk0 + CPT == m) sG[qb + __ldg(net.loc + (c.flip ? 0 : m))] = __dadd_rn(q0, totF);
  }
  tile_store<CPT, BULK>(c, ts, x);
  tile_store_done<BULK>();
}

template <int CPT>
static void launch_backsub_tile(const gnx_network_t* net, const GnxMixedDims& d, const gnx_mixed_coeffs_t* co,
                                const double* b, const double* cond, const double* rsrc, const double* P,
                                double* x, cudaStream_t st) {
  const int G = (ET_T * CPT) >> net->n;
  if (edge_bulk(net->h, b, x))
    gnx_launch_dependent(edge_backsub_tile_kernel<CPT, true>, dim3(gnx_blocks(net->E, G)), dim3(ET_T), 0, st,
                         *net, d, co->a_edge, co->k_edge, 1.0 / co->sigma, b, cond, rsrc, P, x, co->a_cell);
  else
    edge_backsub_tile_kernel<CPT, false><<<gnx_blocks(net->E, G), ET_T, 0, st>>>(
        *net, d, co->a_edge, co->k_edge, 1.0 / co->sigma, b, cond, rsrc, P, x, co->a_cell);
}

extern "C" int gnx_edge_backsub_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                                      const double* b, const double* cond, const double* rsrc,
                                      const double* P, double* x, void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && b && cond && rsrc && x && net->h && co->sigma != 0.0);
  GNX_CHECK_ARG(net->B == 0 || P);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  cudaStream_t st = (cudaStream_t)stream;
  if (net->n == 0) launch_backsub_tile<1>(net, d, co, b, cond, rsrc, P, x, st);
  else if (net->n == 1 || (edge_cpt() == 2 && net->n <= 9)) launch_backsub_tile<2>(net, d, co, b, cond, rsrc, P, x, st);
  else if (net->n <= 10) launch_backsub_tile<4>(net, d, co, b, cond, rsrc, P, x, st);
  else if (co->a_cell) { gnx_set_error("per-cell mass coefficients need n <= 10"); return GNX_ERR_ARG; }
  else {
    const EdgeTiling t = edge_tiling(net->n);
    edge_backsub_kernel<<<gnx_blocks(net->E, t.edges_per_block), t.threads, 0, st>>>(
        *net, d, co->a_edge, co->k_edge, 1.0 / co->sigma, b, cond, rsrc, P, t.seg, x);
  }
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// Persistent step kernel: right-hand side + elimination -> Schur solve -> back-substitution (and the CSR assembly)
// in ONE cooperative launch, one CTA per SM, for networks whose cell lengths fit the GPU's shared memory
// (<= RS_MAX_TILES tiles of 1024 cells per worker SM: 1.2 M cells on 148 SMs -- the C2 benchmark size).
//
// The three-kernel chain above is latency-bound at that size: every kernel boundary costs a launch and a
// fresh round of global loads, and the back-substitution re-reads what the elimination just had in shared
// memory.  Here CTA 0 is the SOLVER and does nothing but the graph-level solve; CTAs 1.. are WORKERS (4 sub-blocks
// of 256 threads, each working on one tile at a time with its own named barrier) that keep the cell lengths of
// THEIR tiles resident in shared memory for the whole step:
//   phase 1  worker: all tiles' lengths requested from the copy engine at once (one mbarrier per tile), what the
//            tiles need to know about their edges gathered into shared-memory records; per tile ONLY the reductions
//            that give the three scalars per edge (nothing is staged or stored: the graph-level solve waits for
//            nothing else), then the CTA arrives on a global counter (release);
//            solver: meanwhile loads the STATIC part of its rows (plan records -> registers; prefetches, argument
//            lines) and only then waits for the counter (acquire) -- schur_body's wait_fn;
//   phase 2  solver: the Schur solve (schur_body<SCH_SINGLE>: rake sweeps out of shared memory, the exchange with the
//            other ranks of a partitioned network, multipliers), then releases an epoch flag;
//            worker: ASSEMBLES the CSR values of its own tiles' edges (the solver never reads them: a_form /
//            ii_assemble of the step, flow_models.py:343) out of the resident lengths and records -- no global load
//            at all; two groups of 512 threads, each double-buffering its bulk stores (rs_assemble) -- then produces
//            and stores its tiles of b, then parks on the epoch flag: the latency-bound graph-level solve hides under
//            the step's largest stream of bytes;
//   phase 3  worker: back-substitution out of the resident lengths (Constant data: the right-hand side entries are
//            recomputed, not re-read), the solution tile leaves as two bulk stores.
// Per-thread arithmetic is that of edge_eliminate_tile_kernel<4,true,true>, edge_backsub_tile_kernel<4,true>
// and mixed_values_tile_kernel<2,true,STIFF>, operation for operation: b, x and the CSR values are
// bit-identical to the chain's.  HBM traffic of the step: h once (8 B/cell) + b, x (16 B/dof) and the values
// (64 B/cell) written, nothing re-read.
// ------------------------------------------------------------------------------------
constexpr int RS_T = 1024;                       // threads per CTA
constexpr int RS_SUB = RS_T / ET_T;              // sub-blocks per CTA
constexpr int RS_CPT = 4;
constexpr int RS_TILE = ET_T * RS_CPT;           // cells per tile
constexpr int RS_MIN_N = 6;                      // 2^6 .. 2^9 cells per edge: <= 16 edges per solve tile,
constexpr int RS_MAX_N = 9;                      // whole edges in a 512-cell assembly tile
constexpr int RS_MAX_G = RS_TILE >> RS_MIN_N;    // edges per tile
constexpr int RS_MAX_TILES = 8;                  // resident tiles per CTA
constexpr int RS_HSLOT = RS_TILE + 2;            // doubles per resident tile (+ parity shift)
constexpr int RS_FSLOT = RS_TILE + 2;
constexpr int RS_GSLOT = RS_TILE + ET_T + 2;
constexpr int RS_ACPT = 2;                       // cells per thread of the in-kernel assembly tiles
constexpr int RS_ATILE = ASM_T * RS_ACPT;
constexpr int RS_AQ = 5 * RS_ATILE + 3 * (RS_ATILE >> RS_MIN_N) + 8;      // staged q-row range of an assembly tile (n >= RS_MIN_N)
constexpr int RS_STAGE = RS_AQ + 3 * RS_ATILE + 4;                        // + its p-row range; or the p | q range of b / x
static_assert(ASM_T == ET_T, "assembly and solve tiles share the 256-thread sub-blocks");
static_assert(RS_STAGE >= RS_FSLOT + RS_GSLOT, "a solve tile fits the staging area of a sub-block");
static_assert(RS_SUB * RS_STAGE >= 3 * GNX_TOP_MAX, "the Schur phase of CTA 0 reuses the staging area");
static_assert((RS_HSLOT % 2) == 0 && (RS_FSLOT % 2) == 0 && (RS_STAGE % 2) == 0 && (RS_AQ % 2) == 0, "16-byte aligned slots");

// what a tile's threads need to know about an edge, gathered ONCE per step into shared memory (the chain of
// dependent global loads edges -> bix -> p_bc costs a DRAM round trip per hop when the L2 is cold)
struct __align__(16) RsEdge {
  double a, pin, pout;                           // mass coefficient; boundary data at the u / v end (flags say whether they apply)
  double q0, Pu;                                 // phase 3: flux at node 0, pressure at node u
  int32_t bu, bv, flags, ebp;                    // bifurcation index of the ends (-1: boundary); bit 0 flip, 1 pin, 2 pout;
                                                 // ebp: bifurcation ends of the edges before this one (net.ebif_prefix)
};

// bu, bv, flip bit, ebp and the node u of edge e: from the edge's 16-byte record when the network has one
// (gnx_network_t.erec), else by walking edges -> bix (one more dependent trip to device memory)
__device__ __forceinline__ int32_t rs_edge_static(const gnx_network_t& net, int32_t e, RsEdge& r) {
  int32_t u;
  if (net.erec) {
    const int4 q = __ldg(reinterpret_cast<const int4*>(net.erec) + e);
    r.bu = q.x; r.bv = q.y; r.ebp = q.z;
    u = q.w & 0x7fffffff;
    r.flags = q.w < 0 ? 1 : 0;
  } else {
    u = net.edges[2 * e];
    const int32_t v = net.edges[2 * e + 1];
    r.bu = net.bix[u]; r.bv = net.bix[v];
    r.ebp = net.ebif_prefix[e];
    r.flags = u > v ? 1 : 0;
  }
  return u;
}

struct StepArgs {
  gnx_network_t net;
  GnxMixedDims d;
  const double* a_edge;
  const double* k_edge;
  double inv_sigma;
  ConstRhs rhs;
  EdgeOut out;
  const double *cond, *rsrc;                     // this rank's edge scalars where the back-substitution finds them
  const double *q0_in, *Pu_in;                   // or, streaming back-substitution: the edges' end flux and pressure ready
                                                 // made (gnx_edge_flux_mixed); NULL: gathered from P by every block
  const double* P;
  double* x;
  double* vals;                                  // CSR values (NULL: no assembly)
  double sigma;
  int32_t nb_tiles, tiles_per_cta;
  unsigned long long* sync;                      // [0] arrival counter (reset by CTA 0), [1] epoch of the last finished Schur phase,
  unsigned long long epoch;                      // [2..9] globaltimer stamps of CTA 0 (diagnostic)
};

__device__ __forceinline__ unsigned long long rs_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int32_t gnx_blocks_dev(int64_t n, int t) { return (int32_t)((n + t - 1) / t); }
// [p, p + bytes) into L2, one request per 128-byte line, by the CTA's threads (p == NULL: nothing)
__device__ __forceinline__ void rs_prefetch_l2(const void* p, int64_t bytes, int tid) {
  if (!p) return;
  const char* c = static_cast<const char*>(p);
  for (int64_t o = (int64_t)tid * 128; o < bytes; o += (int64_t)RS_T * 128) asm volatile("prefetch.global.L2 [%0];" :: "l"(c + o));
}
__device__ __forceinline__ void rs_bar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(ET_T) : "memory"); }

// gnx_seg_incl_scan / gnx_seg_sum for a 256-thread sub-block (same combination order, named barrier)
__device__ __forceinline__ double rs_seg_incl_scan(double v, int seg, double* sh, double* total, int ltid, int bid) {
  const int lane = ltid & 31;
  const int w = seg < 32 ? seg : 32;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double t = __shfl_up_sync(GNX_FULL, v, o, w);
    if (o < w && (lane & (w - 1)) >= o) v += t;
  }
  if (seg <= 32) {
    *total = __shfl_sync(GNX_FULL, v, w - 1, w);
    return v;
  }
  const int wid = ltid >> 5;
  const int wps = seg >> 5;
  rs_bar(bid);
  if (lane == 31) sh[wid] = v;
  rs_bar(bid);
  const int w0 = wid & ~(wps - 1);
  double t = (lane < wps) ? sh[w0 + lane] : 0.0;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double u = __shfl_up_sync(GNX_FULL, t, o);
    if (o < wps && lane >= o) t += u;
  }
  *total = __shfl_sync(GNX_FULL, t, wps - 1);
  const int me = wid - w0;
  const double before = __shfl_sync(GNX_FULL, t, me > 0 ? me - 1 : 0);
  return me > 0 ? v + before : v;
}
template <int K>
__device__ __forceinline__ void rs_seg_sum(double (&v)[K], int seg, double* sh, int ltid, int bid) {
  const int lane = ltid & 31;
  const int w = seg < 32 ? seg : 32;
#pragma unroll
  for (int k = 0; k < K; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double t = __shfl_xor_sync(GNX_FULL, v[k], o, w);
      if (o < w) v[k] += t;
    }
  }
  if (seg <= 32) return;
  const int wid = ltid >> 5;
  const int wps = seg >> 5;
  const int w0 = wid & ~(wps - 1);
  rs_bar(bid);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) sh[k * 32 + wid] = v[k];
  }
  rs_bar(bid);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = (lane < wps) ? sh[k * 32 + w0 + lane] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(GNX_FULL, t, o);
    v[k] = t;
  }
}

// tile geometry without touching global memory (the orientation comes from the edge record)
__device__ __forceinline__ TileCtx rs_tile_ctx(const gnx_network_t& net, const GnxMixedDims& d, int32_t tile, int32_t ltid,
                                               const RsEdge* rec) {
  TileCtx c;
  c.m = d.m;
  const int32_t G = RS_TILE >> net.n;
  c.e0 = tile * G;
  c.nE = min(G, net.E - c.e0);
  const int32_t k0 = ltid * RS_CPT;
  const int32_t le = k0 >> net.n;
  c.active = le < c.nE;
  c.le = c.active ? le : c.nE - 1;
  c.e = c.e0 + c.le;
  c.kk0 = k0 & (c.m - 1);
  c.flip = rec[c.le].flags & 1;
  c.pbase = d.p_off + (int64_t)c.e0 * c.m;
  c.hbase = (int64_t)c.e0 * c.m;
  c.qbase = (int64_t)c.e0 * (c.m + 1);
  c.shp = (int32_t)(c.pbase & 1); c.shq = (int32_t)(c.qbase & 1); c.shh = (int32_t)(c.hbase & 1);
  return c;
}

// staged range -> out[base .. base+len) by the sub-block's first thread (odd ends: its second warp)
__device__ __forceinline__ void rs_put(const double* st, double* __restrict__ out, int64_t b0, int32_t len, int ltid) {
  const int32_t sh = (int32_t)(b0 & 1);
  const int32_t rest = len - sh;
  if (ltid == 0) {
    const int32_t body = rest > 0 ? (rest & ~1) : 0;
    if (body > 0) gnx_bulk_s2g(out + b0 + sh, st + 2 * sh, (uint32_t)body * 8u);
  }
  if (ltid == 32) {
    if (sh && len > 0) out[b0] = st[1];
    if (rest > 0 && (rest & 1)) out[b0 + len - 1] = st[sh + len - 1];
  }
}
__device__ __forceinline__ void rs_tile_store(const TileCtx& c, const double* sF, const double* sG, double* __restrict__ out,
                                              int ltid, int bid) {
  gnx_fence_async_smem();
  rs_bar(bid);
  rs_put(sG, out, c.qbase, c.nE * (c.m + 1), ltid);
  rs_put(sF, out, c.pbase, c.nE * c.m, ltid);
  if (ltid == 0) gnx_bulk_commit();
}

// right-hand side entry of edge-local node j (cells j-1 | j) for Constant data: what edge_eliminate_tile_kernel
// <FUSED_RHS> stages at q dof j -- recomputed with the same, explicitly rounded, operations wherever it is needed
// (the chain rounds these values by storing them; a recomputation must not be contracted into its consumer)
__device__ __forceinline__ double rs_rhs_node(const StepArgs& s, const RsEdge& r, int32_t j, double hl, double hr) {
  double val = __dmul_rn(__dmul_rn(s.rhs.gc, __dadd_rn(hl, hr)), 0.5);
  if (j == 0 && (r.flags & 2)) val = __dadd_rn(val, r.pin);
  return val;
}

// ZFG: f = g = 0 (boundary data only, the common steady case): every F is +0, every q-row entry of b is +0 except
// at boundary ends; the general code then adds and multiplies zeros -- the fast path skips exactly those
// operations (x + 0 = x, x * 0 = +0 for the positive lengths and finite data involved), so it yields the same bits.
// PART: bit 0 the edge scalars (reductions over the tile -> EdgeOut), bit 1 the tile of b (staged, bulk-stored).  The
// streaming kernel does both at once; the persistent kernel computes the scalars in phase 1 -- the graph-level solve
// waits for nothing else -- and writes b in phase 2, out of the same resident lengths (same operations, same bits).
// ENDS: `sloc` holds only loc[0] and loc[m] (all the f = g = 0 path looks up -- the streaming kernel then needs no
// 4 KB numbering table in shared memory: 6 -> 8 resident blocks per SM).
template <bool ZFG, int PART = 3, bool ENDS = false>
__device__ __forceinline__ void rs_eliminate_tile(const StepArgs& s, int32_t tile, int ltid, int bid, const double* sH,
                                                  double* sF, double* sG, double* sh, const RsEdge* rec, const int32_t* sloc) {
  constexpr int CPT = RS_CPT;
  constexpr bool RHS = (PART & 2) != 0, SCAL = (PART & 1) != 0;
  const TileCtx c = rs_tile_ctx(s.net, s.d, tile, ltid, rec);
  const RsEdge& r = rec[c.le];
  const int32_t m = c.m;
  const int32_t cbF = c.le * m + c.shp, cbH = c.le * m + c.shh, qb = c.le * (m + 1) + c.shq;
  double F[CPT], H[CPT];
  double sumF = 0.0, sumG = 0.0, sumH = 0.0;
#pragma unroll
  for (int i = 0; i < CPT; ++i) { H[i] = (SCAL || !ZFG) ? sH[cbH + c.kk0 + i] : 0.0; F[i] = 0.0; }
  if (ZFG) {
    // b tile: zeros, then the boundary ends
    if (RHS) {
#pragma unroll
      for (int i = 0; i < CPT; ++i) sF[RS_FSLOT - 1 - (ltid + i * ET_T)] = 0.0;     // (covers [2, RS_FSLOT); slots 0, 1 below)
#pragma unroll
      for (int i = 0; i < CPT + 1; ++i) sG[ltid + i * ET_T] = 0.0;
      if (ltid < 2) { sF[ltid] = 0.0; sG[RS_GSLOT - 1 - ltid] = 0.0; }
      rs_bar(bid);
    }
    if (c.active && c.kk0 == 0 && (r.flags & 2)) {
      const double v0 = __dadd_rn(0.0, r.pin);
      if (RHS) sG[qb + (ENDS ? sloc[c.flip ? 1 : 0] : sloc[c.flip ? m : 0])] = v0;
      sumG += v0;
    }
    if (c.active && c.kk0 + CPT == m && (r.flags & 4)) {
      const double vm = __dsub_rn(0.0, r.pout);
      if (RHS) sG[qb + (ENDS ? sloc[c.flip ? 0 : 1] : sloc[c.flip ? 0 : m])] = vm;
      sumG += vm;
    }
  } else if (c.active) {
    double hl = c.kk0 > 0 ? sH[cbH + c.kk0 - 1] : 0.0;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int32_t j = c.kk0 + i;
      const double hr = H[i];
      const double val = rs_rhs_node(s, r, j, hl, hr);
      if (RHS) sG[qb + sloc[c.flip ? m - j : j]] = val;
      const double fv = __dmul_rn(s.rhs.fc, hr);
      if (RHS) sF[cbF + gnx_gray(c.flip ? m - 1 - j : j)] = fv;
      F[i] = __dmul_rn(fv, s.inv_sigma);
      sumG += val;
      hl = hr;
    }
    if (c.kk0 + CPT == m) {
      double val = __dmul_rn(__dmul_rn(s.rhs.gc, __dadd_rn(hl, 0.0)), 0.5);
      if (r.flags & 4) val = __dsub_rn(val, r.pout);
      if (RHS) sG[qb + sloc[c.flip ? 0 : m]] = val;
      sumG += val;
    }
  }
#pragma unroll
  for (int i = 0; i < CPT; ++i) { sumF += F[i]; sumH += H[i]; }
  if (RHS) rs_tile_store(c, sF, sG, s.rhs.b, ltid, bid);
  if (!SCAL) return;
  const int segT = m / CPT;
  if (ZFG) {
    double acc[2] = {sumG, sumH};
    rs_seg_sum<2>(acc, segT, sh, ltid, bid);
    if (c.active && c.kk0 == 0) s.out.put(c.e, gnx_edge_cond(r.a, acc[1]), gnx_edge_rsrc(r.a, acc[0], 0.0), 0.0);
    return;
  }
  double tot;
  const double incl = rs_seg_incl_scan(sumF, segT, sh, &tot, ltid, bid);
  double cum = incl - sumF;
  double s2 = 0.0;
#pragma unroll
  for (int i = 0; i < CPT; ++i) { s2 = gnx_s2_step(s2, H[i], cum, F[i]); cum = __dadd_rn(cum, F[i]); }
  double acc[3] = {s2, sumG, sumH};
  rs_seg_sum<3>(acc, segT, sh, ltid, bid);
  if (c.active && c.kk0 == 0) s.out.put(c.e, gnx_edge_cond(r.a, acc[2]), gnx_edge_rsrc(r.a, acc[1], acc[0]), tot);
}

template <bool ZFG>
__device__ __forceinline__ void rs_backsub_tile(const StepArgs& s, int32_t tile, int ltid, int bid, const double* sH,
                                                double* sF, double* sG, double* sh, const RsEdge* rec, const int32_t* sloc) {
  constexpr int CPT = RS_CPT;
  const TileCtx c = rs_tile_ctx(s.net, s.d, tile, ltid, rec);
  const RsEdge& r = rec[c.le];
  const int32_t m = c.m;
  const double Pu = r.Pu, a = r.a, q0 = r.q0;
  const double kv = s.k_edge ? s.k_edge[c.e] : 0.0;
  const double inv_sigma = s.inv_sigma;
  const int32_t cbF = c.le * m + c.shp, cbH = c.le * m + c.shh, qb = c.le * (m + 1) + c.shq;
  int32_t ps[CPT];
  double F[CPT], H[CPT], Gq[CPT];
  double sumF = 0.0;
  double Fl = 0.0, Hl = 1.0;
  double hl = 0.0;
  if (c.kk0 > 0) {
    Hl = sH[cbH + c.kk0 - 1];
    hl = Hl;
    if (!ZFG) Fl = c.active ? __dmul_rn(__dmul_rn(s.rhs.fc, Hl), inv_sigma) : 0.0;
  }
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int32_t kk = c.kk0 + i;
    ps[i] = cbF + gnx_gray(c.flip ? m - 1 - kk : kk);
    H[i] = sH[cbH + kk];
    if (ZFG) {
      F[i] = 0.0;
      Gq[i] = (c.active && kk == 0 && (r.flags & 2)) ? __dadd_rn(0.0, r.pin) : 0.0;
    } else {
      F[i] = c.active ? __dmul_rn(__dmul_rn(s.rhs.fc, H[i]), inv_sigma) : 0.0;
      Gq[i] = c.active ? rs_rhs_node(s, r, kk, hl, H[i]) : 0.0;
      hl = H[i];
      sumF += F[i];
    }
  }
  const int segT = m / CPT;
  double totF = 0.0, totD;
  double cum = 0.0;
  if (!ZFG) {
    const double inclF = rs_seg_incl_scan(sumF, segT, sh, &totF, ltid, bid);
    cum = inclF - sumF;
  }
  double q[CPT], D[CPT];
  double sumD = 0.0;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    q[i] = ZFG ? q0 : __dadd_rn(q0, cum);
    double mq = gnx_mq_right(a, H[i], q[i], F[i]);
    const bool has_left = (i > 0) || (c.kk0 > 0);
    const double fl = i > 0 ? F[i - 1] : Fl, hlft = i > 0 ? H[i - 1] : Hl;
    if (has_left) mq = gnx_mq_left(mq, a, hlft, q[i], fl);
    if (!ZFG && kv != 0.0) {
      mq = __dsub_rn(mq, __ddiv_rn(__dmul_rn(kv, F[i]), H[i]));
      if (has_left) mq = __dadd_rn(mq, __ddiv_rn(__dmul_rn(kv, fl), hlft));
    }
    D[i] = __dsub_rn(Gq[i], mq);
    sumD += D[i];
    if (!ZFG) cum = __dadd_rn(cum, F[i]);
  }
  const double inclD = rs_seg_incl_scan(sumD, segT, sh, &totD, ltid, bid);
  double cd = Pu + (inclD - sumD);
  if (c.active) {
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      cd += D[i];
      // (f = 0: the flux is the same number at every node of the edge -- any slot of the edge's range will do,
      //  no numbering lookup)
      if (ZFG) sG[qb + c.kk0 + i] = q0;
      else sG[qb + sloc[c.flip ? m - (c.kk0 + i) : c.kk0 + i]] = q[i];
      sF[ps[i]] = __dmul_rn(cd, inv_sigma);
    }
    if (c.kk0 + CPT == m) {
      if (ZFG) sG[qb + m] = q0;
      else sG[qb + sloc[c.flip ? 0 : m]] = __dadd_rn(q0, totF);
    }
  }
  rs_tile_store(c, sF, sG, s.x, ltid, bid);
}

// CSR values of one half (hsel = 0 / 1: RS_ATILE cells = whole edges) of the CTA's resident solve tile `stile`:
// the row code of mixed_values_tile_kernel for a group of RS_AT threads (one cell per thread), fed from shared memory
// only -- the cell lengths are the resident ones (Hslot), what it needs to know about an edge sits in the tile's
// records (orientation, multiplier ends, position of its rows in the CSR arrays), the row table is a shared-memory
// copy (stab).  The stand-alone kernel starts every tile with three dependent trips to device memory (ebif_prefix ->
// edges -> bix, then the lengths: 60 % of its stall samples, profiles/r2j); here a tile starts computing at once.
// The caller owns the staging buffer's life cycle (rs_assemble): this function stages the two ranges and commits
// their bulk stores as one group of the group's first thread.
constexpr int RS_AT = 512;                       // threads of an assembling group (two per CTA)
constexpr int RS_AGRP = RS_T / RS_AT;
static_assert(RS_AT == RS_ATILE && RS_AGRP * 2 == RS_SUB, "one cell per thread; a group uses the staging areas of two sub-blocks");
__device__ __forceinline__ void rs_abar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(RS_AT) : "memory"); }

template <bool STIFF>
__device__ __forceinline__ void rs_assemble_tile(const StepArgs& s, int32_t stile, int32_t hsel, const double* Hslot,
                                                 const RsEdge* rec, const int2* stab, int gtid, int bid, double* stage) {
  const gnx_network_t& net = s.net;
  double* const stage_p = stage + RS_AQ;
  const int32_t m = s.d.m, n = net.n;
  const int32_t Ga = RS_ATILE >> n;
  const int32_t e0s = stile * (RS_TILE >> n);
  const int32_t le0 = hsel * Ga;
  const int32_t e0 = e0s + le0;
  if (e0 >= net.E) return;                          // (the network's last solve tile may be half empty)
  const int32_t nE = min(Ga, net.E - e0);
  const RsEdge& rl = rec[le0 + nE - 1];
  const int64_t base = (int64_t)e0 * (5 * (int64_t)m + 1) + rec[le0].ebp;
  const int64_t end = (int64_t)(e0 + nE) * (5 * (int64_t)m + 1) + rl.ebp + (rl.bu >= 0 ? 1 : 0) + (rl.bv >= 0 ? 1 : 0);
  const int64_t pbase = s.d.p_nnz_off + 3 * (int64_t)e0 * m;
  const int32_t psh = (int32_t)(pbase & 1);
  const int32_t shh = (int32_t)(((int64_t)e0s * m) & 1);             // parity shift of the resident tile (rs_tile_ctx)
  const double sg = s.sigma;
  const int32_t k = gtid;                            // cell of the tile
  const int32_t le = k >> n;
  if (le < nE) {
    const int32_t t = k & (m - 1);
    const RsEdge& r = rec[le0 + le];
    const int32_t e = e0 + le;
    AsmEdge E;
    E.flip = r.flags & 1;
    E.blo1 = (E.flip ? r.bv : r.bu) >= 0;
    E.bhi1 = (E.flip ? r.bu : r.bv) >= 0;
    E.a = r.a;
    E.kk = (STIFF && s.k_edge) ? s.k_edge[e] : 0.0;
    E.sgn = E.flip ? -sg : sg;
    E.he = Hslot + shh + (le0 + le) * m;
    E.ac = nullptr;
    E.ebase = (int32_t)((int64_t)e * (5 * (int64_t)m + 1) + r.ebp - base) + (int32_t)(base & 1);
    const int32_t f = asm_qrow<STIFF, false>(stage, stab, E, t, m, sg);
    if (t == m - 1) asm_qrow<STIFF, false>(stage, stab, E, m, m, sg);
    const int32_t slot = psh + 3 * (le * m + gnx_gray(t));
    const bool up = (f >> 17) & 1;
    stage_p[slot] = up ? -E.sgn : E.sgn;
    stage_p[slot + 1] = up ? E.sgn : -E.sgn;
    stage_p[slot + 2] = 0.0;
  }
  gnx_fence_async_smem();
  rs_abar(bid);
  rs_put(stage, s.vals, base, (int32_t)(end - base), gtid);
  rs_put(stage_p, s.vals, pbase, 3 * nE * m, gtid);
  if (gtid == 0) gnx_bulk_commit();
}

// phase 2 of a worker: the CSR values of ITS tiles' edges -- two assembly tiles per resident solve tile, dealt to two
// groups of 512 threads.  A group alternates between the staging areas of its two sub-blocks: while the copy engine
// reads one tile out of shared memory (2 us for 32 KB with four stores in flight per SM -- longer than computing it,
// 1.6 us), the group computes the next into the other.  Then the worker's share of the lambda rows (blocks of ASM_T
// rows, dealt round-robin over the grid's sub-blocks).
__device__ __forceinline__ void rs_assemble(const StepArgs& s, int32_t w, int32_t W, int32_t nmine, int tid,
                                            const double* Hres, const RsEdge* recs, const int2* stab, double* stage) {
  const int grp = tid / RS_AT, gtid = tid % RS_AT, bid = 5 + grp;
  int32_t it = 0;
  for (int32_t j = grp; j < 2 * nmine; j += RS_AGRP, ++it) {
    const int32_t k = j >> 1, hsel = j & 1;
    double* const buf = stage + (size_t)(2 * grp + (it & 1)) * RS_STAGE;
    if (it >= 2) {                                   // the store that last used this buffer has read it
      if (gtid == 0) gnx_bulk_wait_read_1();
      rs_abar(bid);
    }
    if (s.k_edge) rs_assemble_tile<true>(s, w + k * W, hsel, Hres + (size_t)k * RS_HSLOT, recs + k * RS_MAX_G, stab, gtid, bid, buf);
    else rs_assemble_tile<false>(s, w + k * W, hsel, Hres + (size_t)k * RS_HSLOT, recs + k * RS_MAX_G, stab, gtid, bid, buf);
  }
  if (gtid == 0) gnx_bulk_wait_read();
  rs_abar(bid);                                      // both staging areas of the group are free again
  const int sub = tid / ET_T, ltid = tid % ET_T;
  const int32_t nb_lam = s.net.B > 0 ? gnx_blocks_dev(s.net.B, ASM_T) : 0;
  for (int32_t t = w * RS_SUB + sub; t < nb_lam; t += W * RS_SUB) {
    const int64_t r = s.d.lam_off + (int64_t)t * ASM_T + ltid;
    if (r < s.d.N) gnx_mixed_row_emit<GNX_MODE_VALUES>(s.net, s.d, r, s.a_edge, s.k_edge, 1.0, s.sigma, nullptr, nullptr, s.vals);
  }
}

__device__ __forceinline__ void rs_spin_ge(const unsigned long long* p, unsigned long long want) {
  unsigned long long v;
  unsigned spins = 0;
  for (;;) {
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    if (v >= want) break;
    __nanosleep(spins < 64 ? 20 : 200);
    if (++spins > (1u << 24)) __trap();            // seconds: a CTA is missing -- fail instead of hanging the GPU
  }
}

// a few words of every 64-byte line of the Schur arguments, read through the constant bank: CTA 0 first needs them on
// the step's critical path, and a parameter line nobody has touched yet costs a trip to device memory
__device__ __forceinline__ int32_t rs_touch_args(const SchurArgs& a) {
  const gnx_schur_plan_t& pl = a.plan;
  int32_t acc = a.net.E ^ a.net.B ^ (int32_t)(intptr_t)a.net.inc_ptr ^ (int32_t)(intptr_t)a.net.bif_nodes;
  acc ^= pl.B ^ pl.nl ^ pl.l_split ^ pl.nt ^ pl.nb ^ pl.l_sync ^ pl.nxl ^ pl.lvl_ptr[0] ^ pl.lvl_ptr[16] ^ pl.lvl_ptr[32];
  acc ^= (int32_t)(intptr_t)pl.top_rec ^ (int32_t)(intptr_t)pl.ncls ^ (int32_t)(intptr_t)pl.xlist ^ (int32_t)(intptr_t)pl.parent ^
         (int32_t)(intptr_t)pl.ch_ptr ^ (int32_t)(intptr_t)pl.top_of ^ (int32_t)(intptr_t)pl.pedge;
  acc ^= (int32_t)(intptr_t)a.cond ^ (int32_t)(intptr_t)a.rsrc ^ (int32_t)(intptr_t)a.ftot ^ (int32_t)(intptr_t)a.b_lam ^
         (int32_t)(intptr_t)a.P ^ (int32_t)(intptr_t)a.lam_out ^ (int32_t)(intptr_t)a.d ^ (int32_t)(intptr_t)a.r ^
         (int32_t)(intptr_t)a.scal ^ a.max_iter ^ a.chunk ^ (int32_t)(intptr_t)a.wait_flags ^ (int32_t)a.wait_epoch ^
         a.wait_world ^ (int32_t)(intptr_t)a.sig_flags[0] ^ (int32_t)(intptr_t)a.sig_flags[GNX_MAX_PEERS - 1] ^
         (int32_t)(intptr_t)a.xbuf[0] ^ (int32_t)(intptr_t)a.xbuf[GNX_MAX_PEERS - 1] ^ (int32_t)a.xd_off ^ a.phase ^
         (int32_t)__double2loint(a.inv_sigma) ^ (int32_t)__double2loint(a.rtol);
  return acc;
}

// CTA 0 runs the graph-level solve and nothing else (gridDim.x >= 2); CTAs 1.. are the workers: worker w owns the
// tiles w, w + W, w + 2W, ... (W = gridDim.x - 1) for the whole kernel.
__global__ void __launch_bounds__(RS_T, 1)
step_resident_kernel(const __grid_constant__ StepArgs s, const __grid_constant__ SchurArgs a) {
  extern __shared__ __align__(16) double rs_sm[];
  __shared__ double red[2][3 * 32];
  __shared__ double sh_sub[RS_SUB][3 * 32];
  __shared__ __align__(8) uint64_t hbar[RS_MAX_TILES];
  const int tid = threadIdx.x, sub = tid / ET_T, ltid = tid % ET_T, bid = 1 + sub;
  const gnx_network_t& net = s.net;
  const int32_t m = s.d.m;
  // dynamic shared memory: resident cell lengths | staging (4 sub-blocks) | edge records | SubMesh numbering table
  double* const Hres = rs_sm;
  double* const stage = rs_sm + (size_t)s.tiles_per_cta * RS_HSLOT;
  RsEdge* const recs = reinterpret_cast<RsEdge*>(stage + RS_SUB * RS_STAGE);
  int2* const stab = reinterpret_cast<int2*>(recs + s.tiles_per_cta * RS_MAX_G);      // (RsEdge is 16-byte aligned)
  int32_t* const sloc = reinterpret_cast<int32_t*>(stab + (m + 2));
  double* const sF = stage + sub * RS_STAGE;
  double* const sG = sF + RS_FSLOT;
  const int32_t G = RS_TILE >> net.n;
  const bool zfg = s.rhs.fc == 0.0 && s.rhs.gc == 0.0 && !s.k_edge;
  const int32_t W = (int32_t)gridDim.x - 1, w = (int32_t)blockIdx.x - 1;
  // lambda rows of b (every CTA helps)
  for (int64_t r = s.d.lam_off + (int64_t)blockIdx.x * RS_T + tid; r < s.d.N; r += (int64_t)gridDim.x * RS_T) s.rhs.b[r] = 0.0;

  if (blockIdx.x == 0) {
    // ---- the solver CTA.  While the workers eliminate: the rows' plan records (64 B per row) and node classes --
    // static data, cold in a flushed L2 -- are asked for now and are L2 hits when the solve needs them; so are the
    // kernel arguments it will read.
    if (tid == 0) s.sync[2] = rs_now();
    rs_prefetch_l2(a.plan.top_rec, (int64_t)a.plan.nt * 64, tid);
    rs_prefetch_l2(a.plan.ncls, (int64_t)a.plan.B * 4, tid);
    // (the workers' second dependent trip: boundary data at the node an edge record names)
    if (net.V <= (1 << 16)) rs_prefetch_l2(s.rhs.pbc_node, (int64_t)net.V * 8, tid);
    if (tid == 32) { const int32_t t = rs_touch_args(a); if (t == 0x5a5a5a5b) s.sync[31] = (unsigned long long)t; }
    auto wait_workers = [&]() {                    // all workers' edge scalars are in global memory
      if (tid == 0) {
        s.sync[8] = s.sync[3] = rs_now();
        rs_spin_ge(s.sync, (unsigned long long)W);
        s.sync[0] = 0ull;                          // nobody touches the counter again in this launch
        s.sync[4] = rs_now();
      }
      __syncthreads();
    };
    schur_body<SCH_SINGLE>(a, stage, red, wait_workers);
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(s.sync + 1), "l"(s.epoch) : "memory");
      s.sync[5] = s.sync[6] = s.sync[7] = rs_now();
    }
    return;
  }

  // ---- the workers, phase 1: every tile's cell lengths in flight at once ...
  int32_t nmine = 0;
  for (int32_t t = w; t < s.nb_tiles; t += W) ++nmine;
  if (tid == 0) {
    for (int32_t k = 0; k < nmine; ++k) {
      const int32_t tile = w + k * W;
      const int32_t e0 = tile * G, nE = min(G, net.E - e0);
      const int64_t hbase = (int64_t)e0 * m;
      gnx_mbar_init(&hbar[k], 1);
      gnx_mbar_expect_tx(&hbar[k], gnx_bulk_body_bytes(hbase, nE * m));
      gnx_bulk_stream_in(Hres + (size_t)k * RS_HSLOT, net.h, hbase, nE * m, &hbar[k]);
    }
  }
  // ... and, beside them, what the tiles need to know about their edges (one thread per edge: the dependent
  // loads edges -> bix -> boundary data run for all edges of the CTA at once) and the numbering table
  for (int32_t idx = tid; idx < nmine * G; idx += RS_T) {
    const int32_t k = idx / G, le = idx - k * G;
    const int32_t e = (w + k * W) * G + le;
    if (e < net.E) {
      RsEdge r;
      const int32_t u = rs_edge_static(net, e, r);
      r.flags |= ((r.bu < 0 && s.rhs.pbc_node) ? 2 : 0) | ((r.bv < 0 && s.rhs.pbc_out) ? 4 : 0);
      r.a = s.a_edge[e];
      r.pin = (r.flags & 2) ? s.rhs.pbc_node[u] : 0.0;
      r.pout = (r.flags & 4) ? s.rhs.pbc_out[e] : 0.0;
      r.q0 = 0.0; r.Pu = 0.0;
      recs[k * RS_MAX_G + le] = r;
    }
  }
  for (int32_t i = tid; i <= m; i += RS_T) sloc[i] = net.loc[i];
  if (s.vals) for (int32_t i = tid; i <= m; i += RS_T) stab[i] = reinterpret_cast<const int2*>(net.rowtab)[i];
  __syncthreads();
  const bool dbg = blockIdx.x == 1 && tid == 0;    // diagnostic stamps of a worker: sync[16..]
  if (dbg) s.sync[16] = rs_now();
  // the edge scalars: all the graph-level solve waits for (the tile of b is written in phase 2)
  for (int32_t k = sub; k < nmine; k += RS_SUB) {
    gnx_mbar_wait(&hbar[k], 0);
    if (dbg) s.sync[17 + (k >= RS_SUB ? 2 : 0)] = rs_now();
    if (zfg) rs_eliminate_tile<true, 1>(s, w + k * W, ltid, bid, Hres + (size_t)k * RS_HSLOT, sF, sG, sh_sub[sub], recs + k * RS_MAX_G, sloc);
    else rs_eliminate_tile<false, 1>(s, w + k * W, ltid, bid, Hres + (size_t)k * RS_HSLOT, sF, sG, sh_sub[sub], recs + k * RS_MAX_G, sloc);
    if (dbg) s.sync[18 + (k >= RS_SUB ? 2 : 0)] = rs_now();
  }
  if (dbg) s.sync[21] = rs_now();
  // (scalars stored into other ranks' buffers were followed by a system-scope fence of the storing thread, EdgeOut::put;
  //  the gathered layout of mode 0 stores ALL of them remotely: one fence per thread there)
  if (s.out.n > 1 && !s.out.egl) __threadfence_system();
  __syncthreads();
  if (dbg) s.sync[29] = rs_now();
  if (tid == 0) {
    __threadfence();
    atomicAdd(s.sync, 1ull);
  }
  // ---- phase 2, while CTA 0 solves: the CSR values (the solver never reads them: a_form / ii_assemble of the step,
  // flow_models.py:343), then the tiles of b
  if (dbg) s.sync[22] = rs_now();
  if (s.vals) rs_assemble(s, w, W, nmine, tid, Hres, recs, stab, stage);
  if (dbg) s.sync[23] = rs_now();
  for (int32_t k = sub; k < nmine; k += RS_SUB) {
    if (k >= RS_SUB) {                             // the staging area still feeds the bulk stores of my previous tile
      if (ltid == 0) gnx_bulk_wait_read();
      rs_bar(bid);
    }
    if (zfg) rs_eliminate_tile<true, 2>(s, w + k * W, ltid, bid, Hres + (size_t)k * RS_HSLOT, sF, sG, sh_sub[sub], recs + k * RS_MAX_G, sloc);
    else rs_eliminate_tile<false, 2>(s, w + k * W, ltid, bid, Hres + (size_t)k * RS_HSLOT, sF, sG, sh_sub[sub], recs + k * RS_MAX_G, sloc);
  }
  if (ltid == 0) gnx_bulk_wait_read();
  if (dbg) s.sync[30] = rs_now();
  if (tid == 0) rs_spin_ge(s.sync + 1, s.epoch);
  if (dbg) s.sync[24] = rs_now();
  __syncthreads();
  // ---- phase 3: back-substitution out of the resident lengths; first the edges' end pressures (one thread per edge)
  for (int32_t idx = tid; idx < nmine * G; idx += RS_T) {
    const int32_t k = idx / G, le = idx - k * G;
    const int32_t e = (w + k * W) * G + le;
    if (e < net.E) {
      RsEdge& r = recs[k * RS_MAX_G + le];
      const double Pu = r.bu >= 0 ? __ldcg(s.P + r.bu) : 0.0;
      const double Pv = r.bv >= 0 ? __ldcg(s.P + r.bv) : 0.0;
      r.Pu = Pu;
      r.q0 = __dmul_rn(__ldcg(s.cond + e), __dadd_rn(__dsub_rn(Pu, Pv), __ldcg(s.rsrc + e)));
    }
  }
  __syncthreads();
  if (dbg) s.sync[25] = rs_now();
  for (int32_t k = sub; k < nmine; k += RS_SUB) {
    if (k >= RS_SUB) {
      if (ltid == 0) gnx_bulk_wait_read();
      rs_bar(bid);
    }
    if (zfg) rs_backsub_tile<true>(s, w + k * W, ltid, bid, Hres + (size_t)k * RS_HSLOT, sF, sG, sh_sub[sub],
                                   recs + k * RS_MAX_G, sloc);
    else rs_backsub_tile<false>(s, w + k * W, ltid, bid, Hres + (size_t)k * RS_HSLOT, sF, sG, sh_sub[sub],
                                recs + k * RS_MAX_G, sloc);
    if (dbg) s.sync[26 + (k >= RS_SUB ? 1 : 0)] = rs_now();
  }
  if (ltid == 0) gnx_bulk_wait_read();
  if (dbg) s.sync[28] = rs_now();
}

static size_t rs_smem_bytes(int tiles_per_cta, int m) {
  return ((size_t)tiles_per_cta * RS_HSLOT + (size_t)RS_SUB * RS_STAGE) * sizeof(double) +
         (size_t)tiles_per_cta * RS_MAX_G * sizeof(RsEdge) + ((size_t)m + 2) * sizeof(int2) +
         (((size_t)m + 1 + 3) & ~(size_t)3) * sizeof(int32_t);
}

// ------------------------------------------------------------------------------------
// The same per-tile functions as STREAMING kernels (one 1024-cell tile per 256-thread block) for networks too large
// to stay resident: with Constant data the back-substitution RECOMPUTES the right-hand side entries instead of
// re-reading b (16 B/dof of traffic less than edge_backsub_tile_kernel), and f = g = 0 takes the fast path.
// What a tile needs to know about its edges is gathered into shared memory by one thread per edge first.
// ------------------------------------------------------------------------------------
// TPB tiles per block: the cell lengths and edge records of ALL of them are requested when the block starts, so only the
// first tile waits for device memory -- the second one's data land while the first is swept, and its sweep overlaps
// the first one's bulk stores.  (One tile per block spends more time waiting -- two dependent trips for the records
// beside the bulk load -- than computing, and 6-8 resident blocks per SM do not cover that.)
template <bool ZFG, bool BACKSUB, int TPB>
__global__ void __launch_bounds__(ET_T, (ZFG && !BACKSUB) ? 6 : 4)
edge_const_tile_kernel(const __grid_constant__ StepArgs s, int32_t nb_blocks) {
  __shared__ __align__(16) double sH[TPB][RS_HSLOT];
  __shared__ __align__(16) double sFG[RS_FSLOT + RS_GSLOT];
  __shared__ double sh[3 * 32];
  __shared__ RsEdge rec[TPB][RS_MAX_G];
  __shared__ int32_t sloc[ZFG ? 4 : 1024 + 4];                        // (f = g = 0: loc[0] and loc[m] only)
  __shared__ __align__(8) uint64_t hbar[TPB];
  const gnx_network_t& net = s.net;
  const int tid = threadIdx.x;
  const int32_t m = s.d.m;
  if (!BACKSUB && (int32_t)blockIdx.x >= nb_blocks) {               // blocks past the tiles zero the lambda rows of b
    const int64_t r = s.d.lam_off + (int64_t)(blockIdx.x - nb_blocks) * ET_T + tid;
    if (r < s.d.N) s.rhs.b[r] = 0.0;
    return;
  }
  if (!BACKSUB) gnx_pdl_trigger();
  const int32_t G = RS_TILE >> net.n;
  const int32_t t0 = blockIdx.x * TPB;
  const int32_t nt = min(TPB, s.nb_tiles - t0);
  if (tid == 0) {                                                   // static data: in flight before the dependency wait
    for (int j = 0; j < nt; ++j) {
      const int32_t e0 = (t0 + j) * G, nE = min(G, net.E - e0);
      const int64_t hbase = (int64_t)e0 * m;
      gnx_mbar_init(&hbar[j], 1);
      gnx_mbar_expect_tx(&hbar[j], gnx_bulk_body_bytes(hbase, nE * m));
      gnx_bulk_stream_in(sH[j], net.h, hbase, nE * m, &hbar[j]);
    }
  }
  if (ZFG) { if (tid < 2) sloc[tid] = net.loc[tid ? m : 0]; }
  else for (int32_t i = tid; i <= m; i += ET_T) sloc[i] = net.loc[i];
  if (BACKSUB) gnx_pdl_wait();                                      // P and the edge scalars come from the kernels before
  for (int32_t idx = tid; idx < nt * G; idx += ET_T) {
    const int32_t j = idx / G, le = idx - j * G;
    const int32_t e = (t0 + j) * G + le;
    if (e < net.E) {
      RsEdge r;
      const int32_t u = rs_edge_static(net, e, r);
      r.flags |= ((r.bu < 0 && s.rhs.pbc_node) ? 2 : 0) | ((r.bv < 0 && s.rhs.pbc_out) ? 4 : 0);
      r.a = s.a_edge[e];
      r.pin = (r.flags & 2) ? s.rhs.pbc_node[u] : 0.0;
      r.pout = (r.flags & 4) ? s.rhs.pbc_out[e] : 0.0;
      r.q0 = 0.0; r.Pu = 0.0;
      if (BACKSUB) {
        if (s.q0_in) {                                              // (no second dependent trip: P gathered once, per edge)
          r.Pu = s.Pu_in[e];
          r.q0 = s.q0_in[e];
        } else {
          const double Pu = r.bu >= 0 ? s.P[r.bu] : 0.0;
          const double Pv = r.bv >= 0 ? s.P[r.bv] : 0.0;
          r.Pu = Pu;
          r.q0 = __dmul_rn(s.cond[e], __dadd_rn(__dsub_rn(Pu, Pv), s.rsrc[e]));
        }
      }
      rec[j][le] = r;
    }
  }
  __syncthreads();
  for (int j = 0; j < nt; ++j) {
    if (j > 0) {                                                    // the staging area still feeds the previous tile's stores
      if (tid == 0) gnx_bulk_wait_read();
      __syncthreads();
    }
    gnx_mbar_wait(&hbar[j], 0);
    if (BACKSUB) rs_backsub_tile<ZFG>(s, t0 + j, tid, 0, sH[j], sFG, sFG + RS_FSLOT, sh, rec[j], sloc);
    else rs_eliminate_tile<ZFG, 3, ZFG>(s, t0 + j, tid, 0, sH[j], sFG, sFG + RS_FSLOT, sh, rec[j], sloc);
  }
  if (tid == 0) gnx_bulk_wait_read();
}

static int const_tpb() {              // tuning knob: tiles per block of the streaming Constant-data kernels (1 or 2)
  static int v = -1;
  if (v < 0) { const char* e = getenv("GNX_CONST_TPB"); v = e ? atoi(e) : 2; if (v != 1) v = 2; }
  return v;
}

static bool const_tiles_apply(const gnx_network_t* net, const void* b_or_x) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("GNX_CONST_TILES"); on = e ? atoi(e) : 1; }      // A-B knob
  return on && net->n >= RS_MIN_N && net->n <= 10 && edge_bulk(net->h, b_or_x);
}

static void const_tiles_args(StepArgs& s, const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const ConstRhs& rhs) {
  s.net = *net; s.d = gnx_mixed_dims(*net);
  s.a_edge = co->a_edge; s.k_edge = co->k_edge; s.inv_sigma = 1.0 / co->sigma; s.sigma = co->sigma;
  s.rhs = rhs;
  s.cond = s.rsrc = nullptr; s.q0_in = s.Pu_in = nullptr; s.P = nullptr; s.x = nullptr; s.vals = nullptr;
  s.nb_tiles = gnx_blocks(net->E, RS_TILE >> net->n);
  s.tiles_per_cta = 1; s.sync = nullptr; s.epoch = 0;
}

// fused right-hand side + elimination for Constant data through the tile functions of the persistent kernel
static int launch_eliminate_const(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const EdgeOut& out,
                                  const ConstRhs& rhs, cudaStream_t st) {
  StepArgs s;
  const_tiles_args(s, net, co, rhs);
  s.out = out;
  const int nb_lam = net->B > 0 ? gnx_blocks(net->B, ET_T) : 0;
  const bool zfg = rhs.fc == 0.0 && rhs.gc == 0.0 && !co->k_edge;
  const int tpb = const_tpb();
  const int nbk = gnx_blocks(s.nb_tiles, tpb);
#define GNX_ELIM_LAUNCH(Z, T) edge_const_tile_kernel<Z, false, T><<<nbk + nb_lam, ET_T, 0, st>>>(s, nbk)
  if (zfg) { if (tpb == 1) GNX_ELIM_LAUNCH(true, 1); else GNX_ELIM_LAUNCH(true, 2); }
  else { if (tpb == 1) GNX_ELIM_LAUNCH(false, 1); else GNX_ELIM_LAUNCH(false, 2); }
#undef GNX_ELIM_LAUNCH
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// Back-substitution for a right-hand side of Constant data (what gnx_rhs_eliminate_mixed produced): b is not read,
// its entries are recomputed from f_const, g_const and the boundary data; x bit-identical to gnx_edge_backsub_mixed.
extern "C" int gnx_edge_backsub_mixed_const(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const,
                                            double g_const, const double* pbc_out, const double* pbc_node, const double* b,
                                            const double* cond, const double* rsrc, const double* P, double* x,
                                            void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && cond && rsrc && x && net->h && co->sigma != 0.0);
  GNX_CHECK_ARG(net->B == 0 || P);
  if (net->B == 0 || co->a_cell || !const_tiles_apply(net, x)) {
    GNX_CHECK_ARG(b != nullptr);
    return gnx_edge_backsub_mixed(net, co, b, cond, rsrc, P, x, stream);
  }
  StepArgs s;
  const_tiles_args(s, net, co, ConstRhs{f_const, g_const, pbc_out, pbc_node, nullptr});
  edge_out_local(s.out, nullptr, nullptr, nullptr);
  s.cond = cond; s.rsrc = rsrc; s.P = P; s.x = x;
  const bool zfg = f_const == 0.0 && g_const == 0.0 && !co->k_edge;
  cudaStream_t st = (cudaStream_t)stream;
  const int tpb = const_tpb();
  const int32_t nbk = gnx_blocks(s.nb_tiles, tpb);
#define GNX_BS_LAUNCH(Z, T) GNX_CUDA(gnx_launch_dependent(edge_const_tile_kernel<Z, true, T>, dim3(nbk), dim3(ET_T), 0, st, s, nbk))
  if (zfg) { if (tpb == 1) GNX_BS_LAUNCH(true, 1); else GNX_BS_LAUNCH(true, 2); }
  else { if (tpb == 1) GNX_BS_LAUNCH(false, 1); else GNX_BS_LAUNCH(false, 2); }
#undef GNX_BS_LAUNCH
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// End flux and end pressure of every edge from the multipliers, ONCE per solve (one thread per edge):
//   Pu[e] = P at the tail node (0 at a boundary node),  q0[e] = cond[e] (P_u - P_v + rsrc[e])
// -- what every block of the streaming back-substitution otherwise gathers for its own edges through a second dependent
// trip to device memory (edge record -> P), behind which its eight warps wait at a barrier (ncu: 10 warps stalled per
// issue).  q0 may alias ftot and Pu may alias rsrc (a thread reads its edge's entries before it writes them).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
edge_flux_kernel(gnx_network_t net, const double* __restrict__ cond, const double* rsrc, const double* __restrict__ P,
                 double* q0, double* Pu_out) {
  gnx_pdl_trigger();
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  int32_t bu = -1, bv = -1;
  if (e < net.E) {                                  // static data before the dependency wait
    if (net.erec) { const int4 q = __ldg(reinterpret_cast<const int4*>(net.erec) + e); bu = q.x; bv = q.y; }
    else { bu = net.bix[net.edges[2 * e]]; bv = net.bix[net.edges[2 * e + 1]]; }
  }
  gnx_pdl_wait();                                   // P comes from the Schur kernel
  if (e >= net.E) return;
  const double Pu = bu >= 0 ? P[bu] : 0.0;
  const double Pv = bv >= 0 ? P[bv] : 0.0;
  const double q = __dmul_rn(cond[e], __dadd_rn(__dsub_rn(Pu, Pv), rsrc[e]));
  q0[e] = q;
  Pu_out[e] = Pu;
}

extern "C" int gnx_edge_flux_mixed(const gnx_network_t* net, const double* cond, const double* rsrc, const double* P,
                                   double* q0, double* Pu, void* stream) {
  GNX_CHECK_ARG(net && cond && rsrc && q0 && Pu && (net->B == 0 || P) && net->edges && net->bix);
  GNX_CUDA(gnx_launch_dependent(edge_flux_kernel, dim3(gnx_blocks(net->E, 256)), dim3(256), 0, (cudaStream_t)stream, *net, cond,
                                rsrc, P, q0, Pu));
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// gnx_edge_backsub_mixed_const with the edges' end flux q0 [E] and end pressure Pu [E] given (gnx_edge_flux_mixed)
// instead of cond / rsrc / P: x bit-identical.  Needs the tile path (2^6 .. 2^10 cells per edge, per-edge coefficients).
extern "C" int gnx_edge_backsub_mixed_flux(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const,
                                           double g_const, const double* pbc_out, const double* pbc_node, const double* q0,
                                           const double* Pu, double* x, void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && q0 && Pu && x && net->h && co->sigma != 0.0);
  if (co->a_cell || !const_tiles_apply(net, x)) {
    gnx_set_error("gnx_edge_backsub_mixed_flux needs the edge-tile path (2^6..2^10 cells per edge, per-edge coefficients)");
    return GNX_ERR_ARG;
  }
  StepArgs s;
  const_tiles_args(s, net, co, ConstRhs{f_const, g_const, pbc_out, pbc_node, nullptr});
  edge_out_local(s.out, nullptr, nullptr, nullptr);
  s.q0_in = q0; s.Pu_in = Pu; s.x = x;
  const bool zfg = f_const == 0.0 && g_const == 0.0 && !co->k_edge;
  cudaStream_t st = (cudaStream_t)stream;
  const int tpb = const_tpb();
  const int32_t nbk = gnx_blocks(s.nb_tiles, tpb);
#define GNX_BS_LAUNCH(Z, T) GNX_CUDA(gnx_launch_dependent(edge_const_tile_kernel<Z, true, T>, dim3(nbk), dim3(ET_T), 0, st, s, nbk))
  if (zfg) { if (tpb == 1) GNX_BS_LAUNCH(true, 1); else GNX_BS_LAUNCH(true, 2); }
  else { if (tpb == 1) GNX_BS_LAUNCH(false, 1); else GNX_BS_LAUNCH(false, 2); }
#undef GNX_BS_LAUNCH
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

static int rs_enabled() {             // tuning / A-B knob: GNX_STEP_RESIDENT=0 keeps the three-kernel chain
  static int v = -1;
  if (v < 0) { const char* e = getenv("GNX_STEP_RESIDENT"); v = e ? atoi(e) : 1; }
  return v;
}

// 1 if the persistent kernel can run this step (else the caller uses the kernel chain)
int gnx_step_resident_applies(const gnx_network_t* net, const gnx_schur_plan_t* plan, const double* b, const double* x,
                              const unsigned long long* sync) {
  if (!rs_enabled() || !sync || !plan || net->B <= 0 || net->n < RS_MIN_N || net->n > RS_MAX_N) return 0;
  const int64_t nrows = plan->ncls ? (int64_t)plan->lvl_ptr[plan->nl] + plan->nc : plan->B;
  if (nrows > 2 * SCH_T || plan->na > 0 || plan->top_S > 0 || plan->nb > 0) return 0;       // one-CTA Schur solve only
  if (!edge_bulk(net->h, b, x)) return 0;
  if (sch_init() != GNX_OK) return 0;
  const int G = RS_TILE >> net->n;
  const int nb_tiles = gnx_blocks(net->E, G);
  const int grid = nb_tiles + 1 < g_sch_sms ? nb_tiles + 1 : g_sch_sms;     // CTA 0 solves, the others own the tiles
  return grid >= 2 && gnx_blocks(nb_tiles, grid - 1) <= RS_MAX_TILES;
}

int gnx_step_resident_launch(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const, double g_const,
                             const double* pbc_out, const double* pbc_node, double* b, double* x,
                             const gnx_network_t* schur_net, const gnx_schur_plan_t* plan, double* cond, double* rsrc,
                             double* ftot, double* P, double* work, double rtol, int32_t max_iter, int32_t warm,
                             const gnx_exchange_t* ex, unsigned long long* sync, unsigned long long epoch,
                             double* vals, double* resid_host, int32_t* info_host, cudaStream_t st) {
  GNX_CHECK_ARG(!vals || (net->rowtab && ((uintptr_t)vals & 15) == 0));
  GNX_CHECK_ARG(net && co && co->a_edge && b && x && plan && P && work && sync && epoch > 0 && co->sigma != 0.0);
  StepArgs s;
  s.net = *net; s.d = gnx_mixed_dims(*net);
  s.a_edge = co->a_edge; s.k_edge = co->k_edge; s.inv_sigma = 1.0 / co->sigma;
  s.rhs = ConstRhs{f_const, g_const, pbc_out, pbc_node, b};
  int32_t chunk = 0, stride = 0;
  const double *gc = cond, *gr = rsrc, *gf = ftot;
  if (ex && ex->mode == 1) {                     // subtree partition: global-indexed buffer, local copies for phase 3
    GNX_CHECK_ARG(cond && rsrc && ex->phase == 0);
    if (int rc = edge_out_exchange(s.out, ex, net->E, cond, rsrc)) return rc;
    double* g = ex->peers_host[ex->rank];
    s.cond = cond; s.rsrc = rsrc;
    gc = g; gr = g + ex->E_glob; gf = g + 2 * (int64_t)ex->E_glob;
  } else if (ex) {
    if (int rc = edge_out_exchange(s.out, ex, net->E)) return rc;
    double* g = ex->peers_host[ex->rank];
    chunk = ex->chunk; stride = 3 * ex->chunk;
    s.cond = g + (int64_t)ex->rank * stride; s.rsrc = s.cond + chunk;
    gc = g; gr = g + chunk; gf = g + 2 * (int64_t)chunk;
  } else {
    GNX_CHECK_ARG(cond && rsrc && ftot);
    edge_out_local(s.out, cond, rsrc, ftot);
    s.cond = cond; s.rsrc = rsrc;
  }
  s.P = P; s.x = x; s.vals = vals; s.sigma = co->sigma;
  s.q0_in = s.Pu_in = nullptr;
  const int G = RS_TILE >> net->n;
  s.nb_tiles = gnx_blocks(net->E, G);
  const int grid = s.nb_tiles + 1 < g_sch_sms ? s.nb_tiles + 1 : g_sch_sms;
  GNX_CHECK_ARG(grid >= 2);
  s.tiles_per_cta = gnx_blocks(s.nb_tiles, grid - 1);
  GNX_CHECK_ARG(s.tiles_per_cta <= RS_MAX_TILES);
  s.sync = sync; s.epoch = epoch;
  SchurArgs a;
  const GnxMixedDims d = s.d;
  if (int rc = schur_args_fill(a, schur_net, plan, co->sigma, gc, gr, gf, b + d.lam_off, nullptr, P, x + d.lam_off, work,
                               rtol, max_iter, warm, chunk, stride, ex)) return rc;
  const size_t smem = rs_smem_bytes(s.tiles_per_cta, 1 << net->n);
  static size_t smem_set = 0;
  if (smem > smem_set) {
    GNX_CUDA(cudaFuncSetAttribute(step_resident_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  // one CTA per SM and at most as many CTAs as SMs: the grid is co-resident either way; the cooperative launch
  // makes the driver check it (GNX_STEP_COOP=0: ordinary launch, A-B knob)
  static int coop = -1;
  if (coop < 0) { const char* e = getenv("GNX_STEP_COOP"); coop = e ? atoi(e) : 1; }
  if (coop) {
    void* args[] = {(void*)&s, (void*)&a};
    GNX_CUDA(cudaLaunchCooperativeKernel((void*)step_resident_kernel, dim3(grid), dim3(RS_T), args, smem, st));
  } else {
    step_resident_kernel<<<grid, RS_T, smem, st>>>(s, a);
    GNX_LAUNCH_CHECK();
  }
  return schur_report(a, 1, resid_host, info_host, st);
}
