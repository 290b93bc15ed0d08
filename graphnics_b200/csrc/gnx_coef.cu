// Coefficient kernels: expression evaluation (RPN), per-edge CG1 L2 projection, projected
// outlet values.  Reference behaviour replaced: the 3E `project(., CG1(submesh_i))` calls of
// MixedHydraulicNetwork.L_form (graphnics/flowmodels/flow_models.py:315-317), each an assemble +
// LU solve in DOLFIN, repeated every time step (time_stepping.py:219).
#include "gnx_common.cuh"
#include "gnx_rpn.h"

__global__ void __launch_bounds__(256)
eval_expr_kernel(gnx_program_t P, const double* __restrict__ x, int gd, int64_t npts, int pts_per_edge,
                 const double* __restrict__ tangents, const double* __restrict__ edge_attr,
                 double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npts) return;
  double xi[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < gd; ++k) xi[k] = x[i * gd + k];
  const double* tau = nullptr;
  double ea = 0.0;
  if (pts_per_edge > 0) {
    const int64_t e = i / pts_per_edge;
    if (tangents) tau = tangents + e * gd;
    if (edge_attr) ea = edge_attr[e];
  }
  out[i] = gnx_rpn_eval(P, xi, gd, tau, ea);
}

static int check_prog(const gnx_program_t* p) {
  GNX_CHECK_ARG(p != nullptr && p->len > 0 && p->len <= GNX_PROG_MAX);
  // validate stack discipline on the host so the device interpreter needs no checks
  int sp = 0;
  for (int i = 0; i < p->len; ++i) {
    const int op = p->op[i];
    if (op < 10) { ++sp; }
    else if (op == GNX_OP_SELECT) { GNX_CHECK_ARG(sp >= 3); sp -= 2; }
    else if ((op >= GNX_OP_ADD && op <= GNX_OP_POW) || op == GNX_OP_MIN || op == GNX_OP_MAX ||
             op == GNX_OP_FMOD || op == GNX_OP_ATAN2 || (op >= GNX_OP_LT && op <= GNX_OP_GE)) {
      GNX_CHECK_ARG(sp >= 2); --sp;
    } else { GNX_CHECK_ARG(sp >= 1); }
    GNX_CHECK_ARG(sp <= GNX_RPN_STACK);
    if (op == GNX_OP_PARAM) GNX_CHECK_ARG(p->arg[i] >= 0 && p->arg[i] < GNX_PROG_PARAMS);
  }
  GNX_CHECK_ARG(sp == 1);
  return GNX_OK;
}

extern "C" int gnx_eval_expr(const gnx_program_t* prog, const double* x, int32_t gdim, int64_t npts,
                             int32_t pts_per_edge, const double* tangents, const double* edge_attr,
                             double* out, void* stream) {
  if (int rc = check_prog(prog)) return rc;
  GNX_CHECK_ARG(x && out && npts > 0 && gdim >= 1 && gdim <= 3);
  eval_expr_kernel<<<gnx_blocks(npts, 256), 256, 0, (cudaStream_t)stream>>>(
      *prog, x, gdim, npts, pts_per_edge, tangents, edge_attr, out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// L2 projection onto CG1 of every edge by parallel cyclic reduction, spatial (lo-frame) layout.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pcr_setup_kernel(gnx_network_t net, const double* __restrict__ c_nodal, const double* __restrict__ c_mid,
                 double* __restrict__ A, double* __restrict__ Bd, double* __restrict__ Cc,
                 double* __restrict__ D) {
  const int32_t m = 1 << net.n, m1 = m + 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)net.E * m1) return;
  const int32_t e = (int32_t)(i / m1), t = (int32_t)(i - (int64_t)e * m1);
  const bool flip = net.edges[2 * e] > net.edges[2 * e + 1];
  const double* he = net.h + (int64_t)e * m;
  const double* fn = c_nodal + (int64_t)e * m1;
  const double ft = fn[net.loc[t]];
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  if (t > 0) {
    const double hl = he[flip ? m - t : t - 1];
    const double fl = fn[net.loc[t - 1]];
    const double fm = c_mid ? c_mid[(int64_t)e * m + gnx_gray(t - 1)] : 0.5 * (fl + ft);
    a = hl * (1.0 / 6.0); b += hl * (1.0 / 3.0);
    d += hl * (2.0 * fm + ft) * (1.0 / 6.0);
  }
  if (t < m) {
    const double hr = he[flip ? m - 1 - t : t];
    const double fr = fn[net.loc[t + 1]];
    const double fm = c_mid ? c_mid[(int64_t)e * m + gnx_gray(t)] : 0.5 * (ft + fr);
    c = hr * (1.0 / 6.0); b += hr * (1.0 / 3.0);
    d += hr * (ft + 2.0 * fm) * (1.0 / 6.0);
  }
  A[i] = a; Bd[i] = b; Cc[i] = c; D[i] = d;
}

__global__ void __launch_bounds__(256)
pcr_step_kernel(int32_t E, int32_t m1, int32_t s, const double* __restrict__ A, const double* __restrict__ Bd,
                const double* __restrict__ Cc, const double* __restrict__ D, double* __restrict__ A2,
                double* __restrict__ B2, double* __restrict__ C2, double* __restrict__ D2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)E * m1) return;
  const int32_t t = (int32_t)(i % m1);
  double a = A[i], b = Bd[i], c = Cc[i], d = D[i];
  double na = 0.0, nc = 0.0;
  if (t - s >= 0) {
    const double al = -a / Bd[i - s];
    na = al * A[i - s]; b += al * Cc[i - s]; d += al * D[i - s];
  }
  if (t + s < m1) {
    const double ga = -c / Bd[i + s];
    nc = ga * Cc[i + s]; b += ga * A[i + s]; d += ga * D[i + s];
  }
  A2[i] = na; B2[i] = b; C2[i] = nc; D2[i] = d;
}

__global__ void __launch_bounds__(256)
pcr_final_kernel(gnx_network_t net, const double* __restrict__ Bd, const double* __restrict__ D,
                 double* __restrict__ out) {
  const int32_t m1 = (1 << net.n) + 1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)net.E * m1) return;
  const int32_t e = (int32_t)(i / m1), t = (int32_t)(i - (int64_t)e * m1);
  out[(int64_t)e * m1 + net.loc[t]] = D[i] / Bd[i];
}

extern "C" int gnx_project_p1_edges(const gnx_network_t* net, const double* c_nodal, const double* c_mid,
                                    double* out, double* work, void* stream) {
  GNX_CHECK_ARG(net && c_nodal && out && work && net->h && net->loc && net->edges);
  cudaStream_t st = (cudaStream_t)stream;
  const int32_t m1 = (1 << net->n) + 1;
  const int64_t n = (int64_t)net->E * m1;
  double* buf[8];
  for (int k = 0; k < 8; ++k) buf[k] = work + k * n;
  const int bl = gnx_blocks(n, 256);
  pcr_setup_kernel<<<bl, 256, 0, st>>>(*net, c_nodal, c_mid, buf[0], buf[1], buf[2], buf[3]);
  int cur = 0;
  int steps = 0;
  for (int32_t s = 1; s < m1 && steps < 6; s <<= 1, ++steps) {
    double** in = buf + 4 * cur;
    double** ot = buf + 4 * (cur ^ 1);
    pcr_step_kernel<<<bl, 256, 0, st>>>(net->E, m1, s, in[0], in[1], in[2], in[3], ot[0], ot[1], ot[2], ot[3]);
    cur ^= 1;
  }
  pcr_final_kernel<<<bl, 256, 0, st>>>(*net, buf[4 * cur + 1], buf[4 * cur + 3], out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// projected value at the v-end of every edge (only edges whose head is a boundary node matter)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
project_end_kernel(gnx_network_t net, const double* __restrict__ coords, gnx_program_t P, int degree,
                   int window, double* __restrict__ out) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= net.E) return;
  const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
  if (net.bix[v] >= 0) { out[e] = 0.0; return; }           // head is a bifurcation: value unused
  const int32_t m = 1 << net.n, gd = net.gdim;
  const int32_t lo = min(u, v), hi = max(u, v);
  const bool flip = u > v;
  const int32_t K = min(window, m);
  const double* he = net.h + (int64_t)e * m;
  // walk edge-local nodes j = m-K .. m (node m is the head v)
  auto node_x = [&](int32_t j, double* xo) {
    const int32_t t = flip ? m - j : j;
    const int32_t vid = gnx_vertex_id(net.V, net.E, net.n, e, lo, hi, t);
    for (int k = 0; k < 3; ++k) xo[k] = k < gd ? coords[(int64_t)vid * gd + k] : 0.0;
  };
  auto mid_f = [&](const double* xa, const double* xb, double fa, double fb) {
    if (degree < 2) return 0.5 * (fa + fb);
    double xm[3];
    for (int k = 0; k < 3; ++k) xm[k] = 0.5 * (xa[k] + xb[k]);
    return gnx_rpn_eval(P, xm, gd, nullptr, 0.0);
  };
  const int32_t j0 = m - K;
  double xp[3], xc[3], xn[3];
  double fp = 0.0, fc, fn_;
  node_x(j0, xc);
  fc = gnx_rpn_eval(P, xc, gd, nullptr, 0.0);
  double hl = 0.0, fml = 0.0;
  if (j0 > 0) {                                              // left cell of the first window node
    node_x(j0 - 1, xp);
    fp = gnx_rpn_eval(P, xp, gd, nullptr, 0.0);
    hl = he[j0 - 1];
    fml = mid_f(xp, xc, fp, fc);
  }
  double cprev = 0.0, dprev = 0.0;                           // Thomas forward sweep
  for (int32_t j = j0; j <= m; ++j) {
    double hr = 0.0, fmr = 0.0;
    if (j < m) {
      node_x(j + 1, xn);
      fn_ = gnx_rpn_eval(P, xn, gd, nullptr, 0.0);
      hr = he[j];
      fmr = mid_f(xc, xn, fc, fn_);
    }
    const double a = (j > j0) ? hl * (1.0 / 6.0) : 0.0;     // coupling to j0-1 dropped (decays 0.268^K)
    const double b = (hl + hr) * (1.0 / 3.0);
    const double c = hr * (1.0 / 6.0);
    const double d = hl * (2.0 * fml + fc) * (1.0 / 6.0) + hr * (fc + 2.0 * fmr) * (1.0 / 6.0);
    const double den = b - a * cprev;
    cprev = c / den;
    dprev = (d - a * dprev) / den;
    hl = hr; fml = fmr; fc = fn_;
    for (int k = 0; k < 3; ++k) xc[k] = xn[k];
  }
  out[e] = dprev;
}

extern "C" int gnx_project_end_values(const gnx_network_t* net, const double* coords,
                                      const gnx_program_t* prog, int32_t degree, int32_t window,
                                      double* out, void* stream) {
  if (int rc = check_prog(prog)) return rc;
  GNX_CHECK_ARG(net && coords && out && window >= 1 && net->h && net->bix && net->edges);
  project_end_kernel<<<gnx_blocks(net->E, 128), 128, 0, (cudaStream_t)stream>>>(*net, coords, *prog, degree, window, out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}
