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
// One WARP per edge.  (One thread per edge walked the window serially: 65 nodes x (vertex id -> coordinates from
// device memory -> two expression evaluations) -- a chain of dependent loads with E threads on the whole GPU, 230 us
// on the C2 tree and 3/4 of a pulsatile time step.)  The lanes fetch the window's node coordinates and evaluate the
// expression at nodes and midpoints in parallel (two rounds of independent work through shared memory); lane 0 then
// runs the Thomas forward sweep -- the same recurrence in the same order -- out of shared memory.
constexpr int PE_WARPS = 4;
constexpr int PE_MAXW = 64;                                  // window nodes (the host passes <= 64)
__global__ void __launch_bounds__(32 * PE_WARPS)
project_end_kernel(gnx_network_t net, const double* __restrict__ coords, gnx_program_t P, int degree,
                   int window, double* __restrict__ out) {
  __shared__ double sx[PE_WARPS][PE_MAXW + 2][3];
  __shared__ double sf[PE_WARPS][PE_MAXW + 2], sfm[PE_WARPS][PE_MAXW + 2], shh[PE_WARPS][PE_MAXW + 2];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t e = blockIdx.x * PE_WARPS + w;
  if (e >= net.E) return;
  const int32_t u = net.edges[2 * e], v = net.edges[2 * e + 1];
  if (net.bix[v] >= 0) { if (lane == 0) out[e] = 0.0; return; }   // head is a bifurcation: value unused
  const int32_t m = 1 << net.n, gd = net.gdim;
  const int32_t lo = min(u, v), hi = max(u, v);
  const bool flip = u > v;
  const int32_t K = min(min(window, PE_MAXW), m);
  const double* he = net.h + (int64_t)e * m;
  const int32_t j0 = m - K;
  const int32_t jb = j0 > 0 ? j0 - 1 : j0;                   // first node needed (left neighbour of the window)
  const int32_t nn = m - jb + 1;                             // nodes jb .. m  (<= K + 2)
  // round 1: node coordinates and values
  for (int32_t i = lane; i < nn; i += 32) {
    const int32_t j = jb + i;
    const int32_t t = flip ? m - j : j;
    const int32_t vid = gnx_vertex_id(net.V, net.E, net.n, e, lo, hi, t);
    double x[3];
    for (int k = 0; k < 3; ++k) x[k] = k < gd ? coords[(int64_t)vid * gd + k] : 0.0;
    for (int k = 0; k < 3; ++k) sx[w][i][k] = x[k];
    sf[w][i] = gnx_rpn_eval(P, x, gd, nullptr, 0.0);
  }
  __syncwarp();
  // round 2: cells jb .. m-1 (between nodes i and i+1): length and midpoint value
  for (int32_t i = lane; i < nn - 1; i += 32) {
    shh[w][i] = he[jb + i];
    double fm;
    if (degree < 2) fm = 0.5 * (sf[w][i] + sf[w][i + 1]);
    else {
      double xm[3];
      for (int k = 0; k < 3; ++k) xm[k] = 0.5 * (sx[w][i][k] + sx[w][i + 1][k]);
      fm = gnx_rpn_eval(P, xm, gd, nullptr, 0.0);
    }
    sfm[w][i] = fm;
  }
  __syncwarp();
  if (lane != 0) return;
  // Thomas forward sweep over the window nodes j = j0 .. m
  const int32_t o = j0 - jb;                                 // index of node j0
  double hl = 0.0, fml = 0.0;
  if (j0 > 0) { hl = shh[w][0]; fml = sfm[w][0]; }
  double cprev = 0.0, dprev = 0.0;
  for (int32_t j = j0; j <= m; ++j) {
    const int32_t i = o + (j - j0);
    const double fc = sf[w][i];
    double hr = 0.0, fmr = 0.0;
    if (j < m) { hr = shh[w][i]; fmr = sfm[w][i]; }
    const double a = (j > j0) ? hl * (1.0 / 6.0) : 0.0;     // coupling to j0-1 dropped (decays 0.268^K)
    const double b = (hl + hr) * (1.0 / 3.0);
    const double c = hr * (1.0 / 6.0);
    const double d = hl * (2.0 * fml + fc) * (1.0 / 6.0) + hr * (fc + 2.0 * fmr) * (1.0 / 6.0);
    const double inv = 1.0 / (b - a * cprev);               // (one division per node on the serial chain)
    cprev = c * inv;
    dprev = (d - a * dprev) * inv;
    hl = hr; fml = fmr;
  }
  out[e] = dprev;
}

extern "C" int gnx_project_end_values(const gnx_network_t* net, const double* coords,
                                      const gnx_program_t* prog, int32_t degree, int32_t window,
                                      double* out, void* stream) {
  if (int rc = check_prog(prog)) return rc;
  GNX_CHECK_ARG(net && coords && out && window >= 1 && net->h && net->bix && net->edges);
  GNX_CHECK_ARG(window <= PE_MAXW);
  project_end_kernel<<<gnx_blocks(net->E, PE_WARPS), 32 * PE_WARPS, 0, (cudaStream_t)stream>>>(*net, coords, *prog, degree, window, out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}
