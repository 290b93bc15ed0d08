// Closed-form rows of the primal hydraulic system (HydraulicNetwork, flow_models.py:53-126):
// q in DG0(mesh) (dof = global cell index), p in CG1(mesh) (dof = NC + global vertex id).
//   [ a diag(h)   s G ] [q]     G[c, w] = +1 if w is the downstream vertex of cell c, -1 upstream
//   [ -s G^T      0   ] [p]     (a01 = dds(p) v, a10 = -dds(phi) q, :87-89)
// Canonical sparsity: structural entries + an explicit diagonal on the p block (what apply_bc
// needs for the unit diagonal; time_stepping.py:45 adds it explicitly).  Device + TEST-ONLY host.
#ifndef GNX_PRIMAL_H
#define GNX_PRIMAL_H
#include "gnx_index.h"

struct GnxPrimalDims {
  int32_t m;
  int64_t NC, NV, N, p_off;
  int64_t node_nnz_off, mid_nnz_off, nnz;
};

GNX_HD GnxPrimalDims gnx_primal_dims(const gnx_network_t& net) {
  GnxPrimalDims d;
  d.m = 1 << net.n;
  d.NC = (int64_t)net.E * d.m;
  d.NV = net.V + (int64_t)net.E * (d.m - 1);
  d.N = d.NC + d.NV;
  d.p_off = d.NC;
  d.node_nnz_off = 3 * d.NC;
  d.mid_nnz_off = d.node_nnz_off + 2 * (int64_t)net.E + net.V;
  d.nnz = d.mid_nnz_off + 3 * (int64_t)net.E * (d.m - 1);
  return d;
}

// inverse of gnx_vertex_id for midpoint vertices w >= V: edge and lo-frame position
GNX_HD void gnx_vertex_locate(int32_t V, int32_t E, int32_t n, int32_t w, int32_t& e, int32_t& t) {
  const int64_t r0 = (int64_t)(w - V);
  int32_t k = 1;
  while ((int64_t)E * ((1ll << k) - 1) <= r0) ++k;          // level of the vertex, <= n iterations
  const int64_t r = r0 - (int64_t)E * ((1ll << (k - 1)) - 1);
  e = (int32_t)(r >> (k - 1));
  const int32_t g = (int32_t)(r & ((1ll << (k - 1)) - 1));
  t = (2 * gnx_inv_gray(g) + 1) << (n - k);
}

// edge-local helpers: global cell of lo-frame cell s, vertex of lo-frame position t
GNX_HD int64_t gnx_cell_of(int32_t m, int32_t e, int32_t s) { return (int64_t)e * m + gnx_gray(s); }

template <int MODE>
GNX_HD void gnx_primal_row_emit(const gnx_network_t& net, const GnxPrimalDims& d, int64_t r,
                                double a_const, const double* a_cell, double sg, int32_t* rowptr,
                                int32_t* colidx, double* vals) {
  const int32_t m = d.m, n = net.n;
  if (r < d.NC) {                                            // ---- q-row of cell c
    const int32_t e = (int32_t)(r >> n);
    const int32_t s = gnx_inv_gray((int32_t)(r & (m - 1)));
    const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
    const int32_t lo = gnx_imin(u, v), hi = gnx_imax(u, v);
    const bool flip = u > v;
    const double sgn = flip ? -sg : sg;
    int32_t c0 = gnx_vertex_id(net.V, net.E, n, e, lo, hi, s);       // upstream in the lo-frame
    int32_t c1 = gnx_vertex_id(net.V, net.E, n, e, lo, hi, s + 1);
    double v0 = -sgn, v1 = sgn;
    gnx_sort2(c0, v0, c1, v1);
    const int64_t st = 3 * r;
    if (MODE == GNX_MODE_PATTERN) {
      rowptr[r] = (int32_t)st;
      colidx[st] = (int32_t)r; colidx[st + 1] = (int32_t)(d.p_off + c0); colidx[st + 2] = (int32_t)(d.p_off + c1);
    } else {
      const int64_t kl = (int64_t)e * m + (flip ? m - 1 - s : s);      // edge-local (u -> v) cell slot
      vals[st] = (a_cell ? a_cell[kl] : a_const) * net.h[kl];
      vals[st + 1] = v0; vals[st + 2] = v1;
    }
  } else {
    const int32_t w = (int32_t)(r - d.NC);
    if (w < net.V) {                                         // ---- p-row of a graph node
      int64_t pos = d.node_nnz_off + net.inc_ptr[w] + w;
      if (MODE == GNX_MODE_PATTERN) rowptr[r] = (int32_t)pos;
      for (int32_t k = net.inc_ptr[w]; k < net.inc_ptr[w + 1]; ++k, ++pos) {
        const int32_t code = net.inc_edge[k];
        const int32_t e = code >> 1, end = code & 1;
        if (MODE == GNX_MODE_PATTERN) {
          const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
          // in-edge (end 1): last cell of the edge; out-edge: first cell (edge-local frame)
          const int32_t kk = end ? m - 1 : 0;
          colidx[pos] = (int32_t)gnx_cell_of(m, e, flip ? m - 1 - kk : kk);
        } else {
          vals[pos] = end ? -sg : sg;
        }
      }
      if (MODE == GNX_MODE_PATTERN) colidx[pos] = (int32_t)r; else vals[pos] = 0.0;
    } else {                                                 // ---- p-row of an edge-interior vertex
      int32_t e, t;
      gnx_vertex_locate(net.V, net.E, n, w, e, t);
      const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
      const double sgn = flip ? -sg : sg;
      int32_t c0 = (int32_t)gnx_cell_of(m, e, t - 1), c1 = (int32_t)gnx_cell_of(m, e, t);
      double v0 = -sgn, v1 = sgn;
      gnx_sort2(c0, v0, c1, v1);
      const int64_t st = d.mid_nnz_off + 3 * (int64_t)(w - net.V);
      if (MODE == GNX_MODE_PATTERN) {
        rowptr[r] = (int32_t)st;
        colidx[st] = c0; colidx[st + 1] = c1; colidx[st + 2] = (int32_t)r;
      } else {
        vals[st] = v0; vals[st + 1] = v1; vals[st + 2] = 0.0;
      }
    }
  }
  if (MODE == GNX_MODE_PATTERN && r == d.N - 1) rowptr[d.N] = (int32_t)d.nnz;
}

// HydraulicNetwork.L_form (flow_models.py:94-101): L0 = g v dx (per cell), L1 = f phi dx (per vertex).
// g/f are given CELL-WISE as triples (value at the cell's first stored vertex, at the midpoint, at
// the second stored vertex) -- this also covers data that jump between edges (UserExpression with
// eval_cell); NULL -> the constants gc / fc.  *_lin != 0: degree <= 1 data, FEniCS integrates the
// linear interpolant, i.e. midpoint value = average of the ends.
GNX_HD double gnx_primal_rhs_row(const gnx_network_t& net, const GnxPrimalDims& d, const int32_t* cells,
                                 int64_t r, const double* f_cell, int f_lin, double fc,
                                 const double* g_cell, int g_lin, double gc) {
  const int32_t m = d.m, n = net.n;
  auto hcell = [&](int64_t c) {
    const int32_t e = (int32_t)(c >> n);
    const int32_t s = gnx_inv_gray((int32_t)(c & (m - 1)));
    const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
    return net.h[(int64_t)e * m + (flip ? m - 1 - s : s)];
  };
  if (r < d.NC) {
    const double h = hcell(r);
    if (!g_cell) return gc * h;
    const double ga = g_cell[3 * r], gb = g_cell[3 * r + 2];
    const double gm = g_lin ? 0.5 * (ga + gb) : g_cell[3 * r + 1];
    return h * (ga + 4.0 * gm + gb) * (1.0 / 6.0);
  }
  const int32_t w = (int32_t)(r - d.NC);
  if (!f_cell && fc == 0.0) return 0.0;
  auto load = [&](int64_t c) {                     // int f_I phi_w over cell c (w one of its vertices)
    const double h = hcell(c);
    if (!f_cell) return fc * h * 0.5;
    const double fa = f_cell[3 * c], fb = f_cell[3 * c + 2];
    const double fm = f_lin ? 0.5 * (fa + fb) : f_cell[3 * c + 1];
    const double fw = (cells[2 * c] == w) ? fa : fb;
    return h * (fw + 2.0 * fm) * (1.0 / 6.0);
  };
  double out = 0.0;
  if (w < net.V) {
    for (int32_t k = net.inc_ptr[w]; k < net.inc_ptr[w + 1]; ++k) {
      const int32_t code = net.inc_edge[k];
      const int32_t e = code >> 1, end = code & 1;
      const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
      const int32_t kk = end ? m - 1 : 0;
      out += load(gnx_cell_of(m, e, flip ? m - 1 - kk : kk));
    }
  } else {
    int32_t e, t;
    gnx_vertex_locate(net.V, net.E, n, w, e, t);
    const int64_t ca = gnx_cell_of(m, e, t - 1), cb = gnx_cell_of(m, e, t);
    out = load(ca < cb ? ca : cb) + load(ca < cb ? cb : ca);
  }
  return out;
}
#endif  // GNX_PRIMAL_H
