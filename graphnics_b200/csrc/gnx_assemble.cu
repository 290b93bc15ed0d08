// Sparse assembly of the mixed (q per edge, p, lambda) system: CSR pattern built once in closed
// form, values and right-hand side by row gather (every CSR slot = fixed-order sum of its <= 2
// element contributions; deterministic, no atomics).
// Reference behaviour replaced: graphnics/flowmodels/flow_models.py:200-332 (a_form, L_form,
// init_a_form) + ii_assemble/ii_convert (:343-344); time_stepping.py:75-101 (mass_matrix_q).
#include <stdlib.h>
#include "gnx_common.cuh"
#include "gnx_tma.cuh"
#include "gnx_asm_tile.cuh"

extern "C" int64_t gnx_mixed_ndofs(const gnx_network_t* net) { return gnx_mixed_dims(*net).N; }
extern "C" int64_t gnx_mixed_nnz(const gnx_network_t* net) { return gnx_mixed_dims(*net).nnz; }

constexpr int MODE_PATTERN = GNX_MODE_PATTERN;

// One thread per matrix row (row logic in gnx_index.h, shared with the host index self-check).
template <int MODE>
__global__ void __launch_bounds__(256)
mixed_rows_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                  const double* __restrict__ k_edge, double a_const, double sg,
                  int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx,
                  double* __restrict__ vals) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.N) return;
  gnx_mixed_row_emit<MODE>(net, d, r, a_edge, k_edge, a_const, sg, rowptr, colidx, vals);
}

// Values only, staged: the q/p rows of a block occupy one contiguous CSR range; every thread
// drops its <= 6 values into shared memory and the block streams the range out with fully
// coalesced stores.  Blocks past the q/p rows handle the lambda rows (variable length) directly.
__global__ void __launch_bounds__(ASM_T)
mixed_values_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                    const double* __restrict__ k_edge, double a_const, double sg, int32_t nb_qp,
                    double* __restrict__ vals) {
  __shared__ double stage[ASM_T * 6];
  __shared__ int64_t range[2];
  if ((int32_t)blockIdx.x >= nb_qp) {
    const int64_t r = d.lam_off + (int64_t)(blockIdx.x - nb_qp) * ASM_T + threadIdx.x;
    if (r < d.N) gnx_mixed_row_emit<GNX_MODE_VALUES>(net, d, r, a_edge, k_edge, a_const, sg, nullptr, nullptr, vals);
    return;
  }
  const int64_t r0 = (int64_t)blockIdx.x * ASM_T;
  const int64_t r = r0 + threadIdx.x;
  const int64_t rlast = (r0 + ASM_T <= d.lam_off ? r0 + ASM_T : d.lam_off) - 1;
  GnxRow row;
  row.nnz = 0; row.start = 0;
  if (r <= rlast) gnx_mixed_row_qp<GNX_MODE_VALUES>(net, d, r, a_edge, k_edge, a_const, sg, row);
  if (threadIdx.x == 0) range[0] = row.start;
  if (r == rlast) range[1] = row.start + row.nnz;
  __syncthreads();
  const int64_t base = range[0];
  const int32_t off = (int32_t)(row.start - base);
#pragma unroll
  for (int i = 0; i < 6; ++i) if (i < row.nnz) stage[off + i] = row.val[i];
  __syncthreads();
  const int32_t len = (int32_t)(range[1] - base);
  for (int32_t i = threadIdx.x; i < len; i += ASM_T) vals[base + i] = stage[i];
}

// ------------------------------------------------------------------------------------
// Edge-tile value kernel (m <= 1024): a block owns TILE = 256*CPT cells = TILE/m whole edges.
// A thread computes the q-rows of CPT consecutive positions along the edge (neighbour cells,
// orientation and coefficients are per-thread, not per-row, work), ranks the <= 6 entries of each
// row without local arrays and drops them into the block's contiguous CSR range in shared memory;
// the range and the p-rows (pure sign patterns) are then streamed out with coalesced stores.
// ------------------------------------------------------------------------------------
// vals[base .. base+len) = stage[sh .. sh+len), sh = base & 1: 128-bit copies on the 16-byte aligned
// interior, scalar head / tail.  `stage` is 16-byte aligned and `vals` at least 16-byte aligned.
__device__ __forceinline__ void stream_out(const double* stage, double* __restrict__ vals, int64_t base, int32_t len) {
  const int32_t sh = (int32_t)(base & 1);
  double* gp = vals + (base - sh);                     // 16-byte aligned
  const int32_t total = len + sh;
  const int32_t npairs = total >> 1;
  const double2* s2 = reinterpret_cast<const double2*>(stage);
  double2* g2 = reinterpret_cast<double2*>(gp);
#pragma unroll 4
  for (int32_t p = threadIdx.x; p < npairs; p += ASM_T) {
    const double2 v = s2[p];
    if (p == 0 && sh) gp[1] = v.y; else g2[p] = v;
  }
  if (threadIdx.x == 0 && (total & 1) && total - 1 >= sh) gp[total - 1] = stage[total - 1];
}

// Cells are dealt to the lanes of a warp round-robin (cell = 32 * CPT * warp + 32 * round + lane): in every
// round a warp touches 32 neighbouring rows, i.e. shared-memory addresses 5 (q) / 3 (p) doubles apart --
// conflict-free, where a thread owning CPT consecutive cells hit every bank 4 times.
// FAST (bulk copies, net.erec present, 16-byte aligned cell lengths): a tile used to open with three dependent trips to
// device memory -- ebif_prefix[e0], then edges -> bix per edge, then the cell lengths: 60 % of the kernel's stall
// samples (profiles/r2j, ncu source view) -- while a block's useful work is about one such trip long.  Now thread 0
// hands the tile's range of h to the copy engine first thing (completion on an mbarrier), every thread fetches its
// edge's 16-byte record (net.erec) and coefficient beside it, and the row code reads the lengths from shared memory:
// ONE round trip before the arithmetic starts.  Same operations on the same numbers: bit-identical values.
template <int CPT, bool BULK, bool STIFF, bool FAST>
__global__ void __launch_bounds__(ASM_T)
mixed_values_tile_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                         const double* __restrict__ k_edge, const double* __restrict__ a_cell, double a_const, double sg,
                         int32_t nb_tiles, double* __restrict__ vals) {
  constexpr int TILE = ASM_T * CPT;
  if (FAST) {
    extern __shared__ __align__(16) double stage[];
    double* const stage_p = stage + 5 * TILE + 3 * ASM_T + 4;
    double* const sH = stage_p + 3 * TILE + 4;                          // [TILE + 2] cell lengths of the tile (parity-shifted)
    __shared__ __align__(8) uint64_t hbar;
    if ((int32_t)blockIdx.x >= nb_tiles) {
      const int64_t r = d.lam_off + (int64_t)(blockIdx.x - nb_tiles) * ASM_T + threadIdx.x;
      if (r < d.N) gnx_mixed_row_emit<GNX_MODE_VALUES>(net, d, r, a_edge, k_edge, a_const, sg, nullptr, nullptr, vals);
      return;
    }
    const int32_t m = d.m, n = net.n;
    const int32_t G = TILE >> n;
    const int32_t e0 = blockIdx.x * G;
    const int32_t nE = min(G, net.E - e0);
    const int64_t hbase = (int64_t)e0 * m;
    if (threadIdx.x == 0) {
      gnx_mbar_init(&hbar, 1);
      gnx_mbar_expect_tx(&hbar, gnx_bulk_body_bytes(hbase, nE * m));
      gnx_bulk_stream_in(sH, net.h, hbase, nE * m, &hbar);
    }
    const int4* __restrict__ erec = reinterpret_cast<const int4*>(net.erec);
    const int32_t k0 = (threadIdx.x >> 5) * (32 * CPT) + (threadIdx.x & 31);
    // everything a thread needs from device memory, requested at once
    const int4 rf = __ldg(erec + e0), rl = __ldg(erec + e0 + nE - 1);
    int4 rc[CPT];
    double ac[CPT], kc[CPT];
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int32_t le = min((k0 + 32 * i) >> n, nE - 1);
      rc[i] = __ldg(erec + e0 + le);
      ac[i] = a_edge ? a_edge[e0 + le] : a_const;
      kc[i] = (STIFF && k_edge) ? k_edge[e0 + le] : 0.0;
    }
    const int64_t base = (int64_t)e0 * (5 * (int64_t)m + 1) + rf.z;
    const int64_t end = (int64_t)(e0 + nE) * (5 * (int64_t)m + 1) + rl.z + (rl.x >= 0 ? 1 : 0) + (rl.y >= 0 ? 1 : 0);
    const int64_t pbase = d.p_nnz_off + 3 * (int64_t)e0 * m;
    const int32_t psh = (int32_t)(pbase & 1), shh = (int32_t)(hbase & 1);
    const int2* __restrict__ tab = reinterpret_cast<const int2*>(net.rowtab);
    __syncthreads();                                                    // (the barrier's initialisation, thread 0's odd ends)
    gnx_mbar_wait(&hbar, 0);
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int32_t k = k0 + 32 * i;
      const int32_t le = k >> n;
      if (le >= nE) continue;
      const int32_t t = k & (m - 1);
      const int32_t e = e0 + le;
      AsmEdge E;
      E.flip = rc[i].w < 0;
      E.blo1 = (E.flip ? rc[i].y : rc[i].x) >= 0;
      E.bhi1 = (E.flip ? rc[i].x : rc[i].y) >= 0;
      E.a = ac[i];
      E.kk = kc[i];
      E.sgn = E.flip ? -sg : sg;
      E.he = sH + shh + le * m;
      E.ac = nullptr;
      E.ebase = (int32_t)((int64_t)e * (5 * (int64_t)m + 1) + rc[i].z - base) + (int32_t)(base & 1);
      const int32_t f = asm_qrow<STIFF>(stage, tab, E, t, m, sg);
      if (t == m - 1) asm_qrow<STIFF>(stage, tab, E, m, m, sg);
      const int32_t slot = psh + 3 * (le * m + gnx_gray(t));
      const bool up = (f >> 17) & 1;
      stage_p[slot] = up ? -E.sgn : E.sgn;
      stage_p[slot + 1] = up ? E.sgn : -E.sgn;
      stage_p[slot + 2] = 0.0;
    }
    gnx_fence_async_smem();
    __syncthreads();
    gnx_bulk_stream_out(stage, vals, base, (int32_t)(end - base));
    gnx_bulk_stream_out(stage_p, vals, pbase, 3 * nE * m);
    if (threadIdx.x == 0) { gnx_bulk_commit(); gnx_bulk_wait_read(); }
    return;
  }
  // staging buffer, 16-byte aligned; the block's CSR range is staged SHIFTED by the parity of its first
  // slot so that shared and global addresses are 16-byte congruent and the stream-out uses bulk copies
  // or 128-bit loads/stores
  extern __shared__ __align__(16) double stage[];
  double* const stage_p = stage + 5 * TILE + 3 * ASM_T + 4;          // q range: <= 5 TILE + 3 (TILE/m) + 1
  if ((int32_t)blockIdx.x >= nb_tiles) {
    const int64_t r = d.lam_off + (int64_t)(blockIdx.x - nb_tiles) * ASM_T + threadIdx.x;
    if (r < d.N) gnx_mixed_row_emit<GNX_MODE_VALUES>(net, d, r, a_edge, k_edge, a_const, sg, nullptr, nullptr, vals);
    return;
  }
  const int32_t m = d.m, n = net.n;
  const int32_t G = TILE >> n;
  const int32_t e0 = blockIdx.x * G;
  const int32_t nE = min(G, net.E - e0);
  const int64_t base = (int64_t)e0 * (5 * (int64_t)m + 1) + net.ebif_prefix[e0];
  const int64_t end = (int64_t)(e0 + nE) * (5 * (int64_t)m + 1) + net.ebif_prefix[e0 + nE];
  const int64_t pbase = d.p_nnz_off + 3 * (int64_t)e0 * m;
  const int32_t psh = (int32_t)(pbase & 1);
  const int2* __restrict__ tab = reinterpret_cast<const int2*>(net.rowtab);
  const int32_t k0 = (threadIdx.x >> 5) * (32 * CPT) + (threadIdx.x & 31);
  AsmEdge E;
  int32_t cur = -1;
#pragma unroll
  for (int i = 0; i < CPT; ++i) {
    const int32_t k = k0 + 32 * i;                   // cell of the tile
    const int32_t le = k >> n;
    if (le >= nE) break;
    const int32_t t = k & (m - 1);                   // lo-frame position along the edge
    if (le != cur) {                                 // (warp-uniform for m >= 32)
      cur = le;
      const int32_t e = e0 + le;
      const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
      E.flip = u > v;
      E.blo1 = net.bix[E.flip ? v : u] >= 0;
      E.bhi1 = net.bix[E.flip ? u : v] >= 0;
      E.a = a_edge ? a_edge[e] : a_const;
      E.kk = (STIFF && k_edge) ? k_edge[e] : 0.0;
      E.sgn = E.flip ? -sg : sg;
      E.he = net.h + (int64_t)e * m;
      E.ac = a_cell ? a_cell + (int64_t)e * m : nullptr;
      E.ebase = (int32_t)((int64_t)e * (5 * (int64_t)m + 1) + net.ebif_prefix[e] - base) + (int32_t)(base & 1);
    }
    const int32_t f = asm_qrow<STIFF>(stage, tab, E, t, m, sg);
    if (t == m - 1) asm_qrow<STIFF>(stage, tab, E, m, m, sg);       // the edge's last cell also owns node m
    // p-row of lo-frame cell t: (-sgn, +sgn) ordered by the q-dof number, explicit zero diagonal
    const int32_t slot = psh + 3 * (le * m + gnx_gray(t));
    const bool up = (f >> 17) & 1;
    stage_p[slot] = up ? -E.sgn : E.sgn;
    stage_p[slot + 1] = up ? E.sgn : -E.sgn;
    stage_p[slot + 2] = 0.0;
  }
  if (BULK) {
    gnx_fence_async_smem();
    __syncthreads();
    gnx_bulk_stream_out(stage, vals, base, (int32_t)(end - base));
    gnx_bulk_stream_out(stage_p, vals, pbase, 3 * nE * m);
    if (threadIdx.x == 0) { gnx_bulk_commit(); gnx_bulk_wait_read(); }
  } else {
    __syncthreads();
    stream_out(stage, vals, base, (int32_t)(end - base));
    stream_out(stage_p, vals, pbase, 3 * nE * m);
  }
}

template <int CPT, bool BULK, bool STIFF>
static void launch_values_tile3(const gnx_network_t* net, const GnxMixedDims& d, const double* a_edge,
                                const double* k_edge, const double* a_cell, double a_const, double sg, double* vals,
                                cudaStream_t st) {
  const int G = (ASM_T * CPT) >> net->n;
  const int nb_tiles = gnx_blocks(net->E, G);
  const int nb_lam = net->B > 0 ? gnx_blocks(net->B, ASM_T) : 0;
  static int fast_on = -1;
  if (fast_on < 0) { const char* e = getenv("GNX_ASM_FAST"); fast_on = e ? atoi(e) : 1; }                // A-B knob
  if (BULK && fast_on && net->erec && !a_cell && ((uintptr_t)net->h & 15) == 0) {
    constexpr int smem = (asm_stage_doubles<CPT>() + ASM_T * CPT + 2) * (int)sizeof(double);
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
      cudaFuncSetAttribute(mixed_values_tile_kernel<CPT, BULK, STIFF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr = true;
    }
    mixed_values_tile_kernel<CPT, BULK, STIFF, true><<<nb_tiles + nb_lam, ASM_T, smem, st>>>(*net, d, a_edge, k_edge, a_cell,
                                                                                            a_const, sg, nb_tiles, vals);
    return;
  }
  constexpr int smem = asm_stage_doubles<CPT>() * (int)sizeof(double);
  static bool attr = false;
  if (!attr && smem > 48 * 1024) {
    cudaFuncSetAttribute(mixed_values_tile_kernel<CPT, BULK, STIFF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr = true;
  }
  mixed_values_tile_kernel<CPT, BULK, STIFF, false><<<nb_tiles + nb_lam, ASM_T, smem, st>>>(*net, d, a_edge, k_edge, a_cell,
                                                                                           a_const, sg, nb_tiles, vals);
}

template <int CPT>
static void launch_values_tile(const gnx_network_t* net, const GnxMixedDims& d, const double* a_edge,
                               const double* k_edge, const double* a_cell, double a_const, double sg, double* vals,
                               cudaStream_t st) {
  static int bulk = -1;
  if (bulk < 0) { const char* e = getenv("GNX_ASM_BULK"); bulk = e ? atoi(e) : 1; }    // tuning knob
  if (bulk) {
    if (k_edge) launch_values_tile3<CPT, true, true>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
    else launch_values_tile3<CPT, true, false>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
  } else {
    if (k_edge) launch_values_tile3<CPT, false, true>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
    else launch_values_tile3<CPT, false, false>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
  }
}

static int launch_values(const gnx_network_t* net, const GnxMixedDims& d, const double* a_edge,
                         const double* k_edge, double a_const, double sg, double* vals, cudaStream_t st,
                         const double* a_cell = nullptr) {
  if (net->n <= 10 && net->rowtab && ((uintptr_t)vals & 15) == 0) {
    static int cpt = -1;
    if (cpt < 0) { const char* e = getenv("GNX_ASM_CPT"); cpt = e ? atoi(e) : 2; }     // tuning knob
    // a tile (256 * CPT cells) holds whole edges: CPT >= m / 256
    const int c = net->n >= 10 ? 4 : (net->n == 9 && cpt < 2) ? 2 : cpt;
    if (net->n == 0 || c == 1) launch_values_tile<1>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
    else if (net->n == 1 || c == 2) launch_values_tile<2>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
    else launch_values_tile<4>(net, d, a_edge, k_edge, a_cell, a_const, sg, vals, st);
    GNX_LAUNCH_CHECK();
    return GNX_OK;
  }
  if (a_cell) { gnx_set_error("per-cell mass coefficients need the edge-tile assembly (n <= 10, 16-byte aligned values)"); return GNX_ERR_ARG; }
  const int nb_qp = gnx_blocks(d.lam_off, ASM_T);
  const int nb_lam = net->B > 0 ? gnx_blocks(net->B, ASM_T) : 0;
  mixed_values_kernel<<<nb_qp + nb_lam, ASM_T, 0, st>>>(*net, d, a_edge, k_edge, a_const, sg, nb_qp, vals);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

static int check_net(const gnx_network_t* net) {
  GNX_CHECK_ARG(net != nullptr);
  GNX_CHECK_ARG(net->V > 0 && net->E > 0 && net->n >= 0 && net->n <= 24 && net->B >= 0);
  GNX_CHECK_ARG(net->edges && net->bix && net->inc_ptr && net->inc_edge && net->lam_ptr &&
                net->ebif_prefix && net->loc && net->tloc);
  GNX_CHECK_ARG(net->B == 0 || net->bif_nodes);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  GNX_CHECK_ARG(d.N < (1ll << 31) && d.nnz < (1ll << 31));
  return GNX_OK;
}

extern "C" int gnx_build_pattern_mixed(const gnx_network_t* net, int32_t* rowptr, int32_t* colidx,
                                       void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(rowptr && colidx);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  mixed_rows_kernel<MODE_PATTERN><<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      *net, d, nullptr, nullptr, 1.0, 1.0, rowptr, colidx, nullptr);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

extern "C" int gnx_assemble_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                                  const int32_t* rowptr, double* vals, void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(co && co->a_edge && vals && net->h);
  (void)rowptr;  // row starts are closed-form; kept in the ABI for pattern-agnostic callers
  const GnxMixedDims d = gnx_mixed_dims(*net);
  return launch_values(net, d, co->a_edge, co->k_edge, 1.0, co->sigma, vals, (cudaStream_t)stream, co->a_cell);
}

extern "C" int gnx_assemble_mass_q_mixed(const gnx_network_t* net, double coeff,
                                         const int32_t* rowptr, double* vals, void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(vals && net->h);
  (void)rowptr;
  const GnxMixedDims d = gnx_mixed_dims(*net);
  return launch_values(net, d, nullptr, nullptr, coeff, 0.0, vals, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------
// right-hand side, MixedHydraulicNetwork.L_form (flow_models.py:299-332)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mixed_rhs_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ f_nodal,
                 const double* __restrict__ g_nodal, double fc, double gc,
                 const double* __restrict__ pbc_out, const double* __restrict__ pbc_node,
                 double* __restrict__ b) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.N) return;
  b[r] = gnx_mixed_rhs_row(net, d, r, f_nodal, g_nodal, fc, gc, pbc_out, pbc_node);
}

// Constant f / g (the steady models with boundary data only, and every Constant source): an edge-tile
// kernel -- thread per CPT positions, values staged at their dof slots in shared memory, coalesced
// stream-out; with f = g = 0 it reads nothing but the edge ends.  lambda rows are zero.
template <int CPT>
__global__ void __launch_bounds__(ASM_T)
mixed_rhs_const_tile_kernel(gnx_network_t net, GnxMixedDims d, double fc, double gc,
                            const double* __restrict__ pbc_out, const double* __restrict__ pbc_node,
                            int32_t nb_tiles, double* __restrict__ b) {
  constexpr int TILE = ASM_T * CPT;
  __shared__ double sq[TILE + ASM_T], sp[TILE];
  if ((int32_t)blockIdx.x >= nb_tiles) {
    const int64_t r = d.lam_off + (int64_t)(blockIdx.x - nb_tiles) * ASM_T + threadIdx.x;
    if (r < d.N) b[r] = 0.0;
    return;
  }
  const int32_t m = d.m, n = net.n;
  const int32_t G = TILE >> n;
  const int32_t e0 = blockIdx.x * G;
  const int32_t nE = min(G, net.E - e0);
  const int32_t k0 = threadIdx.x * CPT;
  const int32_t le = k0 >> n;
  if (le < nE) {
    const int32_t e = e0 + le;
    const int32_t t0 = k0 & (m - 1);
    const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
    const bool flip = u > v;
    const bool need_h = (fc != 0.0) || (gc != 0.0);
    const double* he = net.h + (int64_t)e * m;
    const int32_t qb = le * (m + 1), cb = le * m;
    const int32_t tlast = (t0 + CPT == m) ? m : t0 + CPT - 1;
    double hl = (need_h && t0 > 0) ? he[flip ? m - t0 : t0 - 1] : 0.0;
    for (int32_t t = t0; t <= tlast; ++t) {
      const double hr = (need_h && t < m) ? he[flip ? m - 1 - t : t] : 0.0;
      double out = gc * (hl + hr) * 0.5;
      if (t == 0 || t == m) {                       // L[i] -= p_bcs[i] v ds(BOUN_OUT) - p_bc v ds(BOUN_IN)  (:324)
        const int32_t node = (t == 0) ? gnx_imin(u, v) : gnx_imax(u, v);
        if (net.bix[node] < 0) {
          if (node == v) { if (pbc_out) out -= pbc_out[e]; }
          else           { if (pbc_node) out += pbc_node[node]; }
        }
      }
      sq[qb + __ldg(net.loc + t)] = out;
      if (t < m) sp[cb + gnx_gray(t)] = fc * hr;   // L[E] += f phi dx_i  (:327)
      hl = hr;
    }
  }
  __syncthreads();
  const int32_t ncell = nE * m, nq = nE * (m + 1);
  double* bq = b + (int64_t)e0 * (m + 1);
  double* bp = b + d.p_off + (int64_t)e0 * m;
  for (int32_t i = threadIdx.x; i < nq; i += ASM_T) bq[i] = sq[i];
  for (int32_t i = threadIdx.x; i < ncell; i += ASM_T) bp[i] = sp[i];
}

template <int CPT>
static void launch_rhs_const_tile(const gnx_network_t* net, const GnxMixedDims& d, double fc, double gc,
                                  const double* pbc_out, const double* pbc_node, double* b, cudaStream_t st) {
  const int G = (ASM_T * CPT) >> net->n;
  const int nb_tiles = gnx_blocks(net->E, G);
  const int nb_lam = net->B > 0 ? gnx_blocks(net->B, ASM_T) : 0;
  mixed_rhs_const_tile_kernel<CPT><<<nb_tiles + nb_lam, ASM_T, 0, st>>>(*net, d, fc, gc, pbc_out, pbc_node, nb_tiles, b);
}

extern "C" int gnx_assemble_rhs_mixed(const gnx_network_t* net, const double* f_nodal,
                                      const double* g_nodal, double f_const, double g_const,
                                      const double* pbc_out, const double* pbc_node, double* b,
                                      void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(b && net->h);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  if (!f_nodal && !g_nodal && net->n <= 10) {
    cudaStream_t st = (cudaStream_t)stream;
    if (net->n == 0) launch_rhs_const_tile<1>(net, d, f_const, g_const, pbc_out, pbc_node, b, st);
    else if (net->n == 1) launch_rhs_const_tile<2>(net, d, f_const, g_const, pbc_out, pbc_node, b, st);
    else launch_rhs_const_tile<4>(net, d, f_const, g_const, pbc_out, pbc_node, b, st);
    GNX_LAUNCH_CHECK();
    return GNX_OK;
  }
  mixed_rhs_kernel<<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      *net, d, f_nodal, g_nodal, f_const, g_const, pbc_out, pbc_node, b);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// coordinates of q-dofs and of cell midpoints (inputs of coefficient evaluation)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dof_coords_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ coords,
                  const int32_t* __restrict__ cells, double* __restrict__ xq, double* __restrict__ xmid) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.lam_off) return;
  const int32_t m = d.m, gd = net.gdim;
  if (r < d.p_off) {
    if (!xq) return;
    const int32_t e = (int32_t)(r / (m + 1));
    const int32_t l = (int32_t)(r - (int64_t)e * (m + 1));
    const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
    const int32_t vid = gnx_vertex_id(net.V, net.E, net.n, e, min(u, v), max(u, v), net.tloc[l]);
    for (int i = 0; i < gd; ++i) xq[r * gd + i] = coords[(int64_t)vid * gd + i];
  } else {
    if (!xmid) return;
    const int64_t c = r - d.p_off;
    const int32_t va = cells[2 * c], vb = cells[2 * c + 1];
    for (int i = 0; i < gd; ++i)
      xmid[c * gd + i] = 0.5 * (coords[(int64_t)va * gd + i] + coords[(int64_t)vb * gd + i]);
  }
}

extern "C" int gnx_dof_coordinates_mixed(const gnx_network_t* net, const double* coords,
                                         const int32_t* cells, double* xq, double* xmid, void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(coords && cells);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  dof_coords_kernel<<<gnx_blocks(d.lam_off, 256), 256, 0, (cudaStream_t)stream>>>(*net, d, coords, cells, xq, xmid);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// y = coeff * blockdiag(M_e) x  on the q rows, 0 on the p / lambda rows: the action of
// TimeDepMixedHydraulicNetwork.mass_matrix_q / mass_vector_q (time_stepping.py:75-123) without
// going through the CSR values (row gather over the <= 2 cells of a vertex).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mass_apply_q_kernel(gnx_network_t net, GnxMixedDims d, double coeff, const double* __restrict__ x,
                    double* __restrict__ y) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.N) return;
  double out = 0.0;
  if (r < d.p_off) {
    const int32_t m = d.m;
    const int32_t e = (int32_t)(r / (m + 1));
    const int32_t l = (int32_t)(r - (int64_t)e * (m + 1));
    const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
    const int32_t t = net.tloc[l];
    const int64_t qb = (int64_t)e * (m + 1);
    const double* he = net.h + (int64_t)e * m;
    const double xt = x[r];
    if (t > 0) out += he[flip ? m - t : t - 1] * (x[qb + net.loc[t - 1]] + 2.0 * xt) * (1.0 / 6.0);
    if (t < m) out += he[flip ? m - 1 - t : t] * (2.0 * xt + x[qb + net.loc[t + 1]]) * (1.0 / 6.0);
    out *= coeff;
  }
  y[r] = out;
}

extern "C" int gnx_mass_apply_q_mixed(const gnx_network_t* net, double coeff, const double* x, double* y,
                                      void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(x && y && x != y && net->h);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  mass_apply_q_kernel<<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(*net, d, coeff, x, y);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}
