// Primal hydraulic model (HydraulicNetwork / TimeDepHydraulicNetwork,
// graphnics/flowmodels/flow_models.py:53-126, time_stepping.py:20-64): q in DG0(mesh), p in
// CG1(mesh), strong Dirichlet data at the boundary nodes.
//   [ diag(a_c h_c)   s G ] [q]   [b_q]      G[c,w] = +1 (w downstream vertex of cell c), -1 upstream
//   [ -s G^T          0   ] [p] = [b_p]
// Pattern / values / rhs in closed form (gnx_primal.h), Dirichlet elimination as in xii.apply_bc
// (flow_models.py:117), and the same exact per-edge elimination as the mixed model:
//   vertex rows :  q_k - q_{k-1} = F_k = b_p[vertex k]/s          -> q_k = q_0 + cumF_k
//   cell rows   :  P_{k+1} - P_k = b_q[k] - a_k h_k q_k   (P = s p) -> second prefix sum
//   edge sum    :  q_0 = cond (P_u - P_v + rsrc),  cond = 1/sum a_k h_k,  rsrc = sum b_q - sum a_k h_k cumF_k
//   node rows   :  Kirchhoff at every graph node -> graph Laplacian in P(bifurcation nodes),
// solved by gnx_schur_solve_nodes (rake + core PCG).  Boundary nodes are Dirichlet: apply_bc has
// moved their value into b, so they enter the scans as P = 0 and come out as x = b.
#include "gnx_common.cuh"
#include "gnx_primal.h"

extern "C" int64_t gnx_primal_ndofs(const gnx_network_t* net) { return gnx_primal_dims(*net).N; }
extern "C" int64_t gnx_primal_nnz(const gnx_network_t* net) { return gnx_primal_dims(*net).nnz; }

static int check_net_primal(const gnx_network_t* net) {
  GNX_CHECK_ARG(net != nullptr);
  GNX_CHECK_ARG(net->V > 0 && net->E > 0 && net->n >= 0 && net->n <= 24);
  GNX_CHECK_ARG(net->edges && net->bix && net->inc_ptr && net->inc_edge && net->h);
  const GnxPrimalDims d = gnx_primal_dims(*net);
  GNX_CHECK_ARG(d.N < (1ll << 31) && d.nnz < (1ll << 31));
  return GNX_OK;
}

template <int MODE>
__global__ void __launch_bounds__(256)
primal_rows_kernel(gnx_network_t net, GnxPrimalDims d, double a_const, const double* __restrict__ a_cell,
                   double sg, int32_t* __restrict__ rowptr, int32_t* __restrict__ colidx,
                   double* __restrict__ vals) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.N) return;
  gnx_primal_row_emit<MODE>(net, d, r, a_const, a_cell, sg, rowptr, colidx, vals);
}

extern "C" int gnx_build_pattern_primal(const gnx_network_t* net, int32_t* rowptr, int32_t* colidx, void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(rowptr && colidx);
  const GnxPrimalDims d = gnx_primal_dims(*net);
  primal_rows_kernel<GNX_MODE_PATTERN><<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      *net, d, 1.0, nullptr, 1.0, rowptr, colidx, nullptr);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

extern "C" int gnx_assemble_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, double* vals,
                                   void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(co && vals);
  const GnxPrimalDims d = gnx_primal_dims(*net);
  primal_rows_kernel<GNX_MODE_VALUES><<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      *net, d, co->a_const, co->a_cell, co->sigma, nullptr, nullptr, vals);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

__global__ void __launch_bounds__(256)
primal_rhs_kernel(gnx_network_t net, GnxPrimalDims d, const int32_t* __restrict__ cells,
                  const double* __restrict__ f_cell, int f_lin, double fc,
                  const double* __restrict__ g_cell, int g_lin, double gc, double* __restrict__ b) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.N) return;
  b[r] = gnx_primal_rhs_row(net, d, cells, r, f_cell, f_lin, fc, g_cell, g_lin, gc);
}

extern "C" int gnx_assemble_rhs_primal(const gnx_network_t* net, const int32_t* cells, const double* f_cell,
                                       int32_t f_lin, double f_const, const double* g_cell, int32_t g_lin,
                                       double g_const, double* b, void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(b && cells);
  const GnxPrimalDims d = gnx_primal_dims(*net);
  primal_rhs_kernel<<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(
      *net, d, cells, f_cell, f_lin, f_const, g_cell, g_lin, g_const, b);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// y[q rows] = (a_c h_c) x_c, y[p rows] = 0: mass_matrix_q / mass_vector_q (time_stepping.py:28-64)
__global__ void __launch_bounds__(256)
primal_mass_apply_kernel(gnx_network_t net, GnxPrimalDims d, double a_const, const double* __restrict__ a_cell,
                         const double* __restrict__ x, double* __restrict__ y) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= d.N) return;
  double out = 0.0;
  if (r < d.NC) {
    const int32_t m = d.m;
    const int32_t e = (int32_t)(r >> net.n);
    const int32_t s = gnx_inv_gray((int32_t)(r & (m - 1)));
    const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
    const int64_t kl = (int64_t)e * m + (flip ? m - 1 - s : s);
    out = (a_cell ? a_cell[kl] : a_const) * net.h[kl] * x[r];
  }
  y[r] = out;
}

extern "C" int gnx_mass_apply_q_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const double* x,
                                       double* y, void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(co && x && y && x != y);
  const GnxPrimalDims d = gnx_primal_dims(*net);
  primal_mass_apply_kernel<<<gnx_blocks(d.N, 256), 256, 0, (cudaStream_t)stream>>>(*net, d, co->a_const, co->a_cell, x, y);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// Dirichlet data on the boundary nodes, symmetric elimination (xii.apply_bc, flow_models.py:117):
// the node's row becomes the identity, its column is moved to the right-hand side.
// One thread per graph node; a boundary node has exactly one incident cell.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
primal_bc_kernel(gnx_network_t net, GnxPrimalDims d, double sg, const double* __restrict__ bc_node,
                 double* __restrict__ vals, double* __restrict__ b) {
  const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= net.V) return;
  // degree-1 nodes of the WHOLE network only ("on_boundary"); on an edge-partitioned network a
  // bifurcation may have a single LOCAL edge, and a boundary node's edge may live on another rank
  if (net.bix[w] >= 0 || net.inc_ptr[w + 1] - net.inc_ptr[w] != 1) return;
  const int32_t m = d.m;
  const int32_t code = net.inc_edge[net.inc_ptr[w]];
  const int32_t e = code >> 1, end = code & 1;
  const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
  const int32_t kk = end ? m - 1 : 0;                             // edge-local end cell
  const int64_t c = gnx_cell_of(m, e, flip ? m - 1 - kk : kk);
  const double val = end ? sg : -sg;                              // entry (q_c, p_w)
  const double g = bc_node[w];
  if (b) {
    atomicAdd(b + c, -val * g);                                  // n = 0: both ends may share the cell
    b[d.p_off + w] = g;
  }
  if (vals) {
    const int64_t prow = d.node_nnz_off + net.inc_ptr[w] + w;     // [cell entry, diagonal]
    vals[prow] = 0.0;
    vals[prow + 1] = 1.0;
    // q-row of c: [diag, p_c0, p_c1] with sorted vertex columns; w is a graph node (id < V <= any midpoint id)
    const int32_t s = flip ? m - 1 - kk : kk;
    const int32_t lo = gnx_imin(net.edges[2 * e], net.edges[2 * e + 1]), hi = gnx_imax(net.edges[2 * e], net.edges[2 * e + 1]);
    const int32_t c0 = gnx_vertex_id(net.V, net.E, net.n, e, lo, hi, s);
    const int32_t c1 = gnx_vertex_id(net.V, net.E, net.n, e, lo, hi, s + 1);
    const int32_t other = (c0 == w) ? c1 : c0;
    vals[3 * c + ((w < other) ? 1 : 2)] = 0.0;
  }
}

extern "C" int gnx_apply_bc_primal(const gnx_network_t* net, double sigma, const double* bc_node, double* vals,
                                   double* b, void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(bc_node && (vals || b));
  const GnxPrimalDims d = gnx_primal_dims(*net);
  primal_bc_kernel<<<gnx_blocks(net->V, 128), 128, 0, (cudaStream_t)stream>>>(*net, d, sigma, bc_node, vals, b);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// per-edge elimination: one thread per cell, block-segmented scans (several edges per block when
// m < 256, several chunks per edge when m > 1024)
// ------------------------------------------------------------------------------------
struct PTiling { int threads, seg, edges_per_block; };
static inline PTiling ptiling(int32_t n) {
  const int64_t m = 1ll << n;
  PTiling t;
  t.threads = m >= 1024 ? 1024 : (m >= 256 ? (int)m : 256);
  t.seg = m < t.threads ? (int)m : t.threads;
  t.edges_per_block = t.threads / t.seg;
  return t;
}

struct PEdge {
  int32_t u, v, lo, hi, m;
  bool flip;
  int64_t cb;
  // global cell (q-dof) of edge-local cell k, p-dof of edge-local interior node j (1..m-1)
  __device__ __forceinline__ int64_t qdof(int32_t k) const { return cb + gnx_gray(flip ? m - 1 - k : k); }
};

__device__ __forceinline__ PEdge pedge(const gnx_network_t& net, int32_t m, int32_t e) {
  PEdge f;
  f.u = net.edges[2 * e]; f.v = net.edges[2 * e + 1];
  f.lo = gnx_imin(f.u, f.v); f.hi = gnx_imax(f.u, f.v);
  f.flip = f.u > f.v; f.m = m;
  f.cb = (int64_t)e * m;
  return f;
}

__global__ void __launch_bounds__(1024)
primal_eliminate_kernel(gnx_network_t net, GnxPrimalDims d, double a_const, const double* __restrict__ a_cell,
                        double inv_sigma, const double* __restrict__ b, int seg,
                        double* __restrict__ cond, double* __restrict__ rsrc, double* __restrict__ ftot) {
  __shared__ double sh[3 * 32];
  const int epb = blockDim.x / seg;
  const int32_t e = blockIdx.x * epb + threadIdx.x / seg;
  const int gl = threadIdx.x & (seg - 1);
  const bool active = e < net.E;
  const int32_t ee = active ? e : net.E - 1;
  const int32_t m = d.m;
  const PEdge f = pedge(net, m, ee);
  double carry = 0.0;
  double acc[3] = {0.0, 0.0, 0.0};                 // sum a h cumF, sum b_q, sum a h
  for (int32_t base = 0; base < m; base += seg) {
    const int32_t k = base + gl;                   // cell k, its upstream node k
    double Fk = 0.0;                               // jump of q across node k (interior nodes only)
    if (k > 0) {
      const int32_t vid = gnx_vertex_id(net.V, net.E, net.n, ee, f.lo, f.hi, f.flip ? m - k : k);
      Fk = b[d.p_off + vid] * inv_sigma;
    }
    const int64_t kl = f.cb + k;
    const double ah = (a_cell ? a_cell[kl] : a_const) * net.h[kl];
    double tot;
    const double inc = gnx_seg_incl_scan(Fk, seg, sh, &tot);
    acc[0] += ah * (carry + inc);                  // cumF_k includes F_k
    acc[1] += b[f.qdof(k)];
    acc[2] += ah;
    carry += tot;
  }
  gnx_seg_sum<3>(acc, seg, sh);
  if (active && gl == 0) {
    cond[e] = 1.0 / acc[2];
    rsrc[e] = acc[1] - acc[0];
    ftot[e] = carry;
  }
}

extern "C" int gnx_edge_eliminate_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const double* b,
                                         double* cond, double* rsrc, double* ftot, void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(co && b && cond && rsrc && ftot && co->sigma != 0.0);
  const GnxPrimalDims d = gnx_primal_dims(*net);
  const PTiling t = ptiling(net->n);
  primal_eliminate_kernel<<<gnx_blocks(net->E, t.edges_per_block), t.threads, 0, (cudaStream_t)stream>>>(
      *net, d, co->a_const, co->a_cell, 1.0 / co->sigma, b, t.seg, cond, rsrc, ftot);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

__global__ void __launch_bounds__(1024)
primal_backsub_kernel(gnx_network_t net, GnxPrimalDims d, double a_const, const double* __restrict__ a_cell,
                      double inv_sigma, const double* __restrict__ b, const double* __restrict__ cond,
                      const double* __restrict__ rsrc, const double* __restrict__ P, int seg,
                      double* __restrict__ x) {
  __shared__ double sh[32];
  const int epb = blockDim.x / seg;
  const int32_t e = blockIdx.x * epb + threadIdx.x / seg;
  const int gl = threadIdx.x & (seg - 1);
  const bool active = e < net.E;
  const int32_t ee = active ? e : net.E - 1;
  const int32_t m = d.m;
  const PEdge f = pedge(net, m, ee);
  const int32_t bu = net.bix[f.u], bv = net.bix[f.v];
  const double Pu = bu >= 0 ? P[bu] : 0.0;         // Dirichlet nodes: value already moved into b
  const double Pv = bv >= 0 ? P[bv] : 0.0;
  const double q0 = cond[ee] * (Pu - Pv + rsrc[ee]);
  double carryF = 0.0, carryD = 0.0;
  for (int32_t base = 0; base < m; base += seg) {
    const int32_t k = base + gl;
    double Fk = 0.0;
    int32_t vid = 0;
    if (k > 0) {
      vid = gnx_vertex_id(net.V, net.E, net.n, ee, f.lo, f.hi, f.flip ? m - k : k);
      Fk = b[d.p_off + vid] * inv_sigma;
    }
    const int64_t kl = f.cb + k;
    const double ah = (a_cell ? a_cell[kl] : a_const) * net.h[kl];
    const int64_t qd = f.qdof(k);
    double totF, totD;
    const double inc = gnx_seg_incl_scan(Fk, seg, sh, &totF);
    const double qk = q0 + carryF + inc;
    const double Dk = b[qd] - ah * qk;             // P_{k+1} - P_k
    const double incD = gnx_seg_incl_scan(Dk, seg, sh, &totD);
    if (active) {
      x[qd] = qk;
      if (k > 0) x[d.p_off + vid] = (Pu + carryD + incD - Dk) * inv_sigma;   // P at node k
    }
    carryF += totF;
    carryD += totD;
  }
}

// p at the graph nodes: bifurcations from the Schur solve, Dirichlet nodes from b
__global__ void __launch_bounds__(128)
primal_nodes_kernel(gnx_network_t net, GnxPrimalDims d, double inv_sigma, const double* __restrict__ b,
                    const double* __restrict__ P, double* __restrict__ x) {
  const int32_t w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= net.V) return;
  const int32_t bi = net.bix[w];
  x[d.p_off + w] = bi >= 0 ? P[bi] * inv_sigma : b[d.p_off + w];
}

extern "C" int gnx_edge_backsub_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const double* b,
                                       const double* cond, const double* rsrc, const double* P, double* x,
                                       void* stream) {
  if (int rc = check_net_primal(net)) return rc;
  GNX_CHECK_ARG(co && b && cond && rsrc && x && co->sigma != 0.0 && (net->B == 0 || P));
  const GnxPrimalDims d = gnx_primal_dims(*net);
  const PTiling t = ptiling(net->n);
  cudaStream_t st = (cudaStream_t)stream;
  primal_backsub_kernel<<<gnx_blocks(net->E, t.edges_per_block), t.threads, 0, st>>>(
      *net, d, co->a_const, co->a_cell, 1.0 / co->sigma, b, cond, rsrc, P, t.seg, x);
  primal_nodes_kernel<<<gnx_blocks(net->V, 128), 128, 0, st>>>(*net, d, 1.0 / co->sigma, b, P, x);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}
