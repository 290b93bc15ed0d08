// TEST-ONLY host build of the closed-form index/row math in graphnics_b200/csrc/gnx_index.h.
// The CUDA kernels call exactly these __host__ __device__ functions, one thread per row/cell;
// here they are looped on the CPU so that the numbering can be checked against the oracle in the
// GPU-less container (`pytest -m "not gpu"`).  The product never loads this file.
#include <cmath>
#include "../../graphnics_b200/csrc/gnx_index.h"
#include "../../graphnics_b200/csrc/gnx_rpn.h"
#include "../../graphnics_b200/csrc/gnx_primal.h"

extern "C" {

void hc_refine(const double* pos, const int32_t* edges, int32_t V, int32_t E, int32_t gd, int32_t n,
               double* coords, int32_t* cells, int32_t* mf, double* h) {
  for (int64_t i = 0; i < (int64_t)V * gd; ++i) coords[i] = pos[i];
  const int64_t nc = (int64_t)E << n;
  for (int64_t gid = 0; gid < nc; ++gid) {
    if (gd == 1) gnx_refine_cell<1>(pos, edges, V, E, n, gid, coords, cells, mf, h);
    else if (gd == 2) gnx_refine_cell<2>(pos, edges, V, E, n, gid, coords, cells, mf, h);
    else gnx_refine_cell<3>(pos, edges, V, E, n, gid, coords, cells, mf, h);
  }
}

void hc_tangents(const double* pos, const int32_t* edges, int32_t E, int32_t gd, double* out) {
  for (int32_t e = 0; e < E; ++e) {
    const int32_t u = edges[2 * e], v = edges[2 * e + 1];
    double t[3];
    for (int i = 0; i < gd; ++i) t[i] = pos[(int64_t)v * gd + i] - pos[(int64_t)u * gd + i];
    double s = t[0] * t[0];
    for (int i = 1; i < gd; ++i) s = std::fma(t[i], t[i], s);
    const double inv = 1.0 / std::sqrt(s);
    for (int i = 0; i < gd; ++i) out[(int64_t)e * gd + i] = t[i] * inv;
  }
}

void hc_pattern(const gnx_network_t* net, int32_t* rowptr, int32_t* colidx) {
  const GnxMixedDims d = gnx_mixed_dims(*net);
  for (int64_t r = 0; r < d.N; ++r)
    gnx_mixed_row_emit<GNX_MODE_PATTERN>(*net, d, r, nullptr, nullptr, 1.0, 1.0, rowptr, colidx, nullptr);
}

void hc_values(const gnx_network_t* net, const double* a_edge, const double* k_edge, double a_const,
               double sg, double* vals) {
  const GnxMixedDims d = gnx_mixed_dims(*net);
  for (int64_t r = 0; r < d.N; ++r)
    gnx_mixed_row_emit<GNX_MODE_VALUES>(*net, d, r, a_edge, k_edge, a_const, sg, nullptr, nullptr, vals);
}

void hc_rhs(const gnx_network_t* net, const double* f, const double* g, double fc, double gc,
            const double* pbc_out, const double* pbc_node, double* b) {
  const GnxMixedDims d = gnx_mixed_dims(*net);
  for (int64_t r = 0; r < d.N; ++r) b[r] = gnx_mixed_rhs_row(*net, d, r, f, g, fc, gc, pbc_out, pbc_node);
}

double hc_rpn(const gnx_program_t* prog, const double* x, int gd, const double* tau, double ea) {
  return gnx_rpn_eval(*prog, x, gd, tau, ea);
}

// ---- primal model (gnx_primal.h)
int64_t hc_primal_ndofs(const gnx_network_t* net) { return gnx_primal_dims(*net).N; }
int64_t hc_primal_nnz(const gnx_network_t* net) { return gnx_primal_dims(*net).nnz; }

void hc_primal_pattern(const gnx_network_t* net, int32_t* rowptr, int32_t* colidx) {
  const GnxPrimalDims d = gnx_primal_dims(*net);
  for (int64_t r = 0; r < d.N; ++r)
    gnx_primal_row_emit<GNX_MODE_PATTERN>(*net, d, r, 1.0, nullptr, 1.0, rowptr, colidx, nullptr);
}

void hc_primal_values(const gnx_network_t* net, double a_const, const double* a_cell, double sg, double* vals) {
  const GnxPrimalDims d = gnx_primal_dims(*net);
  for (int64_t r = 0; r < d.N; ++r)
    gnx_primal_row_emit<GNX_MODE_VALUES>(*net, d, r, a_const, a_cell, sg, nullptr, nullptr, vals);
}

void hc_primal_rhs(const gnx_network_t* net, const int32_t* cells, const double* f_cell, int f_lin, double fc,
                   const double* g_cell, int g_lin, double gc, double* b) {
  const GnxPrimalDims d = gnx_primal_dims(*net);
  for (int64_t r = 0; r < d.N; ++r) b[r] = gnx_primal_rhs_row(*net, d, cells, r, f_cell, f_lin, fc, g_cell, g_lin, gc);
}

int64_t hc_ndofs(const gnx_network_t* net) { return gnx_mixed_dims(*net).N; }
int64_t hc_nnz(const gnx_network_t* net) { return gnx_mixed_dims(*net).nnz; }
}
