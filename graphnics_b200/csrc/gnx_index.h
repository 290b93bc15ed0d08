// Closed-form numbering of the hierarchically bisected 1-D network mesh and of the mixed
// system's CSR rows.  Pure integer/fp64 arithmetic, usable from device code and from the
// host-compiled index self-check in tests/hostcheck (TEST ONLY; the product never runs it).
//
// Derivation (DOLFIN rules A1-A5, reference call sites
// graphnics/fenics_graph.py:58-99,132-156):
//   * positions along an edge are counted from lo = min(u,v):  t in [0,m], cell s = [s,s+1];
//   * n bisections with "children (v0,new),(v1,new)" make the cell with spatial index s the
//     local cell gray(s) = s ^ (s>>1) of its edge (global cell e*m + gray(s));
//   * the midpoint created at level k (1..n) inside level-(k-1) cell c' gets vertex id
//     V + E*(2^(k-1)-1) + c';  position t has level k = n - ctz(t) and sits in the level-(k-1)
//     cell with spatial index t >> (n-k+1);
//   * SubMesh vertex numbering (first encounter) is the same permutation loc[] of t for every
//     edge; it is tabulated once on the host (m+1 ints).
#ifndef GNX_INDEX_H
#define GNX_INDEX_H

#include <stdint.h>
#include <math.h>
#include "../../include/graphnics_b200.h"

#if defined(__CUDACC__)
#define GNX_HD __host__ __device__ __forceinline__
#else
#define GNX_HD inline
#endif

GNX_HD int32_t gnx_gray(int32_t s) { return s ^ (s >> 1); }

GNX_HD int32_t gnx_inv_gray(int32_t j) {
  j ^= j >> 1; j ^= j >> 2; j ^= j >> 4; j ^= j >> 8; j ^= j >> 16;
  return j;
}

GNX_HD int32_t gnx_ctz(int32_t t) {
#if defined(__CUDA_ARCH__)
  return __ffs(t) - 1;
#else
  return __builtin_ctz((unsigned)t);
#endif
}

// global vertex id of position t (counted from lo) on edge e
GNX_HD int32_t gnx_vertex_id(int32_t V, int32_t E, int32_t n, int32_t e, int32_t lo, int32_t hi,
                             int32_t t) {
  const int32_t m = 1 << n;
  if (t == 0) return lo;
  if (t == m) return hi;
  const int32_t k = n - gnx_ctz(t);                 // level at which the vertex was created
  const int32_t half = 1 << (k - 1);                // #level-(k-1) cells per edge
  return V + E * (half - 1) + e * half + gnx_gray(t >> (n - k + 1));
}

// ---------------------------------------------------------------------------------------
// Mixed system rows.  Layout W = [q_0..q_{E-1} | p | lambda] (flow_models.py:187).
// ---------------------------------------------------------------------------------------
struct GnxMixedDims {
  int32_t m, q_per_edge;
  int64_t p_off, lam_off, N;      // dof offsets
  int64_t q_nnz, p_nnz_off, lam_nnz_off, nnz;
};

GNX_HD GnxMixedDims gnx_mixed_dims(const gnx_network_t& net) {
  GnxMixedDims d;
  d.m = 1 << net.n;
  d.q_per_edge = d.m + 1;
  d.p_off = (int64_t)net.E * (d.m + 1);
  d.lam_off = d.p_off + (int64_t)net.E * d.m;
  d.N = d.lam_off + net.B;
  d.q_nnz = (int64_t)net.E * (5 * (int64_t)d.m + 1) + net.I;
  d.p_nnz_off = d.q_nnz;
  d.lam_nnz_off = d.q_nnz + 3 * (int64_t)net.E * d.m;
  d.nnz = d.lam_nnz_off + net.I + net.B;
  return d;
}

// One q- or p-row: up to 6 sorted (col,val) pairs.
struct GnxRow {
  int64_t start;
  int32_t nnz;
  int32_t col[6];
  double val[6];
};

GNX_HD void gnx_sort2(int32_t& c0, double& v0, int32_t& c1, double& v1) {
  if (c1 < c0) { int32_t c = c0; c0 = c1; c1 = c; double v = v0; v0 = v1; v1 = v; }
}

// q-row of edge e, SubMesh-local vertex l.
//   hl/hr : lengths of the cells left/right of the vertex in the lo-frame (unused side ignored)
//   a,k   : mass / stiffness coefficients of the edge; sg = sigma
// (flow_models.py:233 mass, :236 -B^T, :257-259 C^T, :399 Stokes stiffness)
GNX_HD void gnx_mixed_q_row(const gnx_network_t& net, const GnxMixedDims& d, int32_t e, int32_t l,
                            const int32_t* __restrict__ loc, int32_t t, int32_t u, int32_t v,
                            int32_t bif_lo, int32_t bif_hi, int32_t ebif_before, double hl, double hr,
                            double a, double k, double sg, GnxRow& r) {
  const int32_t m = d.m;
  const bool flip = u > v;
  const double sgn = flip ? -sg : sg;
  const int32_t qb = e * (m + 1);
  const int64_t pb = d.p_off + (int64_t)e * m;
  const int32_t lm = loc[m];
  r.start = (int64_t)e * (5 * (int64_t)m + 1) + ebif_before + 5 * (int64_t)l -
            (l > 0 ? 2 - (bif_lo >= 0) : 0) - (l > lm ? 2 - (bif_hi >= 0) : 0);
  int32_t nq = 0;
  int32_t qc[3]; double qv[3];
  double diag = 0.0;
  if (t > 0) {
    qc[nq] = qb + loc[t - 1]; qv[nq] = a * hl * (1.0 / 6.0) - k / hl; ++nq;
    diag += a * hl * (2.0 / 6.0) + k / hl;
  }
  if (t < m) {
    qc[nq] = qb + loc[t + 1]; qv[nq] = a * hr * (1.0 / 6.0) - k / hr; ++nq;
    diag += a * hr * (2.0 / 6.0) + k / hr;
  }
  qc[nq] = qb + l; qv[nq] = diag; ++nq;
  // sort the (<=3) q columns
  if (nq == 3) gnx_sort2(qc[0], qv[0], qc[1], qv[1]);
  gnx_sort2(qc[nq - 2], qv[nq - 2], qc[nq - 1], qv[nq - 1]);
  if (nq == 3) gnx_sort2(qc[0], qv[0], qc[1], qv[1]);
  int32_t c = 0;
  for (int32_t i = 0; i < nq; ++i) { r.col[c] = qc[i]; r.val[c] = qv[i]; ++c; }
  // p columns: cell s=t-1 (coefficient -sgn) and s=t (+sgn)
  int32_t pc0 = -1, pc1 = -1; double pv0 = 0, pv1 = 0;
  if (t > 0) { pc0 = (int32_t)(pb + gnx_gray(t - 1)); pv0 = -sgn; }
  if (t < m) { pc1 = (int32_t)(pb + gnx_gray(t)); pv1 = sgn; }
  if (pc0 >= 0 && pc1 >= 0) {
    gnx_sort2(pc0, pv0, pc1, pv1);
    r.col[c] = pc0; r.val[c] = pv0; ++c;
    r.col[c] = pc1; r.val[c] = pv1; ++c;
  } else if (pc0 >= 0) {
    r.col[c] = pc0; r.val[c] = pv0; ++c;
  } else {
    r.col[c] = pc1; r.val[c] = pv1; ++c;
  }
  // lambda column at a bifurcation end: +sigma if the edge points INTO the node (node == v)
  if (t == 0 && bif_lo >= 0) {
    r.col[c] = (int32_t)(d.lam_off + bif_lo); r.val[c] = flip ? sg : -sg; ++c;
  } else if (t == m && bif_hi >= 0) {
    r.col[c] = (int32_t)(d.lam_off + bif_hi); r.val[c] = flip ? -sg : sg; ++c;
  }
  r.nnz = c;
}

// p-row of global cell c (flow_models.py:235): +sigma at the downstream q vertex, -sigma upstream,
// plus the explicit-zero diagonal of init_a_form (:275-278).
GNX_HD void gnx_mixed_p_row(const gnx_network_t& net, const GnxMixedDims& d, int32_t c,
                            const int32_t* __restrict__ loc, int32_t u, int32_t v, double sg,
                            GnxRow& r) {
  const int32_t m = d.m;
  const int32_t e = c >> net.n;
  const int32_t s = gnx_inv_gray(c & (m - 1));
  const double sgn = (u > v) ? -sg : sg;
  const int32_t qb = e * (m + 1);
  int32_t c0 = qb + loc[s], c1 = qb + loc[s + 1];
  double v0 = -sgn, v1 = sgn;
  gnx_sort2(c0, v0, c1, v1);
  r.start = d.p_nnz_off + 3 * (int64_t)c;
  r.nnz = 3;
  r.col[0] = c0; r.val[0] = v0;
  r.col[1] = c1; r.val[1] = v1;
  r.col[2] = (int32_t)(d.p_off + c); r.val[2] = 0.0;
}

// position (lo-frame) of the edge end described by an incidence code
GNX_HD int32_t gnx_end_t(int32_t m, int32_t u, int32_t v, int32_t end) {
  const bool flip = u > v;
  // end 0: node == u ; end 1: node == v
  return (end == 0) ? (flip ? m : 0) : (flip ? 0 : m);
}

GNX_HD int32_t gnx_imin(int32_t a, int32_t b) { return a < b ? a : b; }
GNX_HD int32_t gnx_imax(int32_t a, int32_t b) { return a > b ? a : b; }

enum { GNX_MODE_PATTERN = 0, GNX_MODE_VALUES = 1 };

// q- or p-row r (< lam_off) of the mixed matrix: start slot, sorted columns, values.
template <int MODE>
GNX_HD void gnx_mixed_row_qp(const gnx_network_t& net, const GnxMixedDims& d, int64_t r,
                             const double* a_edge, const double* k_edge, double a_const, double sg,
                             GnxRow& row) {
  const int32_t m = d.m;
  if (r < d.p_off) {
    const int32_t e = (int32_t)(r / (m + 1));
    const int32_t l = (int32_t)(r - (int64_t)e * (m + 1));
    const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
    const int32_t lo = gnx_imin(u, v), hi = gnx_imax(u, v);
    const int32_t t = net.tloc[l];
    double hl = 1.0, hr = 1.0, a = 1.0, k = 0.0;
    if (MODE == GNX_MODE_VALUES) {
      const bool flip = u > v;
      const double* he = net.h + (int64_t)e * m;
      if (t > 0) hl = he[flip ? m - t : t - 1];        // lo-frame cell s=t-1
      if (t < m) hr = he[flip ? m - 1 - t : t];        // lo-frame cell s=t
      a = a_edge ? a_edge[e] : a_const;
      k = k_edge ? k_edge[e] : 0.0;
    }
    gnx_mixed_q_row(net, d, e, l, net.loc, t, u, v, net.bix[lo], net.bix[hi],
                    net.ebif_prefix[e], hl, hr, a, k, sg, row);
  } else {
    const int32_t c = (int32_t)(r - d.p_off);
    const int32_t e = c >> net.n;
    gnx_mixed_p_row(net, d, c, net.loc, net.edges[2 * e], net.edges[2 * e + 1], sg, row);
  }
}

// Whole-row emitter used by the pattern / value kernels (one thread per matrix row r).
template <int MODE>
GNX_HD void gnx_mixed_row_emit(const gnx_network_t& net, const GnxMixedDims& d, int64_t r,
                               const double* a_edge, const double* k_edge, double a_const, double sg,
                               int32_t* rowptr, int32_t* colidx, double* vals) {
  const int32_t m = d.m;
  if (r < d.lam_off) {
    GnxRow row;
    gnx_mixed_row_qp<MODE>(net, d, r, a_edge, k_edge, a_const, sg, row);
    if (MODE == GNX_MODE_PATTERN) {
      rowptr[r] = (int32_t)row.start;
      for (int i = 0; i < 6; ++i) if (i < row.nnz) colidx[row.start + i] = row.col[i];
    } else {
      for (int i = 0; i < 6; ++i) if (i < row.nnz) vals[row.start + i] = row.val[i];
    }
  } else {
    // lambda row (flow_models.py:260-262): sign * q_e(end at the bifurcation), explicit-zero diagonal
    const int32_t bi = (int32_t)(r - d.lam_off);
    const int32_t node = net.bif_nodes[bi];
    const int32_t k0 = net.inc_ptr[node], k1 = net.inc_ptr[node + 1];
    int64_t pos = d.lam_nnz_off + net.lam_ptr[bi] + bi;
    if (MODE == GNX_MODE_PATTERN) rowptr[r] = (int32_t)pos;
    for (int32_t k = k0; k < k1; ++k, ++pos) {
      const int32_t code = net.inc_edge[k];
      const int32_t e = code >> 1, end = code & 1;
      if (MODE == GNX_MODE_PATTERN) {
        const int32_t t = gnx_end_t(m, net.edges[2 * e], net.edges[2 * e + 1], end);
        colidx[pos] = e * (m + 1) + net.loc[t];
      } else {
        vals[pos] = end ? sg : -sg;
      }
    }
    if (MODE == GNX_MODE_PATTERN) colidx[pos] = (int32_t)r; else vals[pos] = 0.0;
  }
  if (MODE == GNX_MODE_PATTERN && r == d.N - 1) rowptr[d.N] = (int32_t)d.nnz;
}

// One entry of MixedHydraulicNetwork.L_form (flow_models.py:299-332), row r.
// f/g: nodal arrays in q-dof layout, or NULL -> the constants fc / gc.
GNX_HD double gnx_mixed_rhs_row(const gnx_network_t& net, const GnxMixedDims& d, int64_t r,
                                const double* f_nodal, const double* g_nodal, double fc, double gc,
                                const double* pbc_out, const double* pbc_node) {
  const int32_t m = d.m;
  double out = 0.0;
  if (r < d.p_off) {
    const int32_t e = (int32_t)(r / (m + 1));
    const int32_t l = (int32_t)(r - (int64_t)e * (m + 1));
    const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
    const bool flip = u > v;
    const int32_t t = net.tloc[l];
    const int64_t qb = (int64_t)e * (m + 1);
    // L[i] -= p_bcs[i] v ds(BOUN_OUT) - p_bc v ds(BOUN_IN)     (:324)
    if (t == 0 || t == m) {
      const int32_t node = (t == 0) ? gnx_imin(u, v) : gnx_imax(u, v);
      if (net.bix[node] < 0) {                // degree-1 node == boundary
        if (node == v) { if (pbc_out) out -= pbc_out[e]; }              // edge INTO node: BOUN_OUT
        else           { if (pbc_node) out += pbc_node[node]; }         // edge OUT of node: BOUN_IN
      }
    }
    // L[i] += gs[i] v dx_i  (:325)  P1 x P1 mass action on the projected nodal g
    const double* he = net.h + (int64_t)e * m;
    if (g_nodal) {
      const double gt = g_nodal[r];
      if (t > 0) out += he[flip ? m - t : t - 1] * (g_nodal[qb + net.loc[t - 1]] + 2.0 * gt) * (1.0 / 6.0);
      if (t < m) out += he[flip ? m - 1 - t : t] * (2.0 * gt + g_nodal[qb + net.loc[t + 1]]) * (1.0 / 6.0);
    } else if (gc != 0.0) {
      if (t > 0) out += he[flip ? m - t : t - 1] * gc * 0.5;
      if (t < m) out += he[flip ? m - 1 - t : t] * gc * 0.5;
    }
  } else if (r < d.lam_off) {
    // L[E] += fs[i] phi dx_i  (:327)
    if (f_nodal || fc != 0.0) {
      const int32_t c = (int32_t)(r - d.p_off);
      const int32_t e = c >> net.n;
      const int32_t s = gnx_inv_gray(c & (m - 1));
      const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
      const int64_t qb = (int64_t)e * (m + 1);
      const double hc = net.h[(int64_t)e * m + (flip ? m - 1 - s : s)];
      out = f_nodal ? hc * (f_nodal[qb + net.loc[s]] + f_nodal[qb + net.loc[s + 1]]) * 0.5 : hc * fc;
    }
  }
  return out;
}

// FenicsGraph.get_mesh (fenics_graph.py:58-99) for one (edge, spatial cell): walks the bisection
// tree (A2/A3: midpoint = (x_a + x_b)/2, recursively -> bit-exact DOLFIN coordinates).
template <int GD>
GNX_HD void gnx_refine_cell(const double* pos, const int32_t* edges, int32_t V, int32_t E, int32_t n,
                            int64_t gid, double* coords, int32_t* cells, int32_t* mf, double* h) {
  const int32_t m = 1 << n;
  const int32_t e = (int32_t)(gid >> n);
  const int32_t s = (int32_t)(gid & (m - 1));
  const int32_t u = edges[2 * e], v = edges[2 * e + 1];
  const int32_t lo = gnx_imin(u, v), hi = gnx_imax(u, v);
  double x0[GD], x1[GD];
  for (int i = 0; i < GD; ++i) { x0[i] = pos[(int64_t)lo * GD + i]; x1[i] = pos[(int64_t)hi * GD + i]; }
  int32_t t0 = 0, t1 = m;
  for (int32_t lvl = 0; lvl < n; ++lvl) {
    const int32_t tm = (t0 + t1) >> 1;
    double xm[GD];
    for (int i = 0; i < GD; ++i) xm[i] = (x0[i] + x1[i]) * 0.5;   // (a+b)*c cannot contract to an fma
    if (s < tm) { t1 = tm; for (int i = 0; i < GD; ++i) x1[i] = xm[i]; }
    else        { t0 = tm; for (int i = 0; i < GD; ++i) x0[i] = xm[i]; }
  }
  double ss = 0.0;
  for (int i = 0; i < GD; ++i) { const double dd = x1[i] - x0[i]; ss += dd * dd; }
  h[(int64_t)e * m + (u > v ? m - 1 - s : s)] = sqrt(ss);
  const int32_t va = gnx_vertex_id(V, E, n, e, lo, hi, s);
  const int32_t vb = gnx_vertex_id(V, E, n, e, lo, hi, s + 1);
  const int64_t c = (int64_t)e * m + gnx_gray(s);
  cells[2 * c] = gnx_imin(va, vb);
  cells[2 * c + 1] = gnx_imax(va, vb);
  if (mf) mf[c] = e;
  if (s + 1 < m) for (int i = 0; i < GD; ++i) coords[(int64_t)vb * GD + i] = x1[i];
}

#endif  // GNX_INDEX_H
