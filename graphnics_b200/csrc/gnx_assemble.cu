// Sparse assembly of the mixed (q per edge, p, lambda) system: CSR pattern built once in closed
// form, values and right-hand side by row gather (every CSR slot = fixed-order sum of its <= 2
// element contributions; deterministic, no atomics).
// Reference behaviour replaced: graphnics/flowmodels/flow_models.py:200-332 (a_form, L_form,
// init_a_form) + ii_assemble/ii_convert (:343-344); time_stepping.py:75-101 (mass_matrix_q).
#include "gnx_common.cuh"

extern "C" int64_t gnx_mixed_ndofs(const gnx_network_t* net) { return gnx_mixed_dims(*net).N; }
extern "C" int64_t gnx_mixed_nnz(const gnx_network_t* net) { return gnx_mixed_dims(*net).nnz; }

constexpr int MODE_PATTERN = GNX_MODE_PATTERN, MODE_VALUES = GNX_MODE_VALUES;

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
constexpr int ASM_T = 256;
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

static int launch_values(const gnx_network_t* net, const GnxMixedDims& d, const double* a_edge,
                         const double* k_edge, double a_const, double sg, double* vals, cudaStream_t st) {
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
  return launch_values(net, d, co->a_edge, co->k_edge, 1.0, co->sigma, vals, (cudaStream_t)stream);
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

extern "C" int gnx_assemble_rhs_mixed(const gnx_network_t* net, const double* f_nodal,
                                      const double* g_nodal, double f_const, double g_const,
                                      const double* pbc_out, const double* pbc_node, double* b,
                                      void* stream) {
  if (int rc = check_net(net)) return rc;
  GNX_CHECK_ARG(b && net->h);
  const GnxMixedDims d = gnx_mixed_dims(*net);
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
