// Saddle-point solve for the mixed network system: exact per-edge elimination (two segmented
// scans over the cells of every edge) -> SPD graph-level Schur system on the bifurcation
// multipliers -> Jacobi-PCG (2 fused kernels / iteration, CUDA-graph replay, device-side
// convergence) -> back-substitution.  Replaces ii_convert + LUSolver(A,"mumps").solve
// (graphnics/flowmodels/flow_models.py:344-348, time_stepping.py:187-192).
//
// Edge-local frame: nodes j = 0..m from u to v, cell k between nodes k and k+1.
//   p-rows :  sigma (q_{k+1} - q_k) = b_p[k]              -> q_j = q_0 + cumF_j
//   q-rows :  (a M + k K) q |_j + P_j - P_{j-1} = b_q[j]   (P = sigma p, P_{-1} = P(u), P_m = P(v))
//   sum    :  q_0 = cond (P(u) - P(v) + rsrc),  cond = 1/(a L),  rsrc = sum b_q - a sum_k h_k (cumF_k+cumF_{k+1})/2
//   lam-row:  sum_in q_e(m) - sum_out q_e(0) = b_lam / sigma   (Kirchhoff) -> graph Laplacian in P(b)
#include <cooperative_groups.h>
#include "gnx_common.cuh"
namespace cg = cooperative_groups;

// lane-group width for one edge: min(32, m) (m is a power of two)
static inline int group_width(int32_t n) { return n >= 5 ? 32 : (1 << n); }

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
__global__ void __launch_bounds__(128)
edge_eliminate_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                      double inv_sigma, const double* __restrict__ b, int gw,
                      double* __restrict__ cond, double* __restrict__ rsrc, double* __restrict__ ftot) {
  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t e = (int32_t)(gthread / gw);
  const int gl = threadIdx.x & (gw - 1);
  const bool active = e < net.E;
  const int32_t ee = active ? e : net.E - 1;      // keep whole warps alive for the shuffles
  const EdgeFrame f = edge_frame(net, d, ee);
  const int32_t m = d.m;
  double carry = 0.0, s2 = 0.0, sg = 0.0, len = 0.0;
  for (int32_t base = 0; base < m; base += gw) {
    const int32_t k = base + gl;                 // cell k and node k
    const double Fk = b[f.pdof(k)] * inv_sigma;
    const double hk = net.h[f.hb + k];
    const double Gk = b[f.qdof(net.loc, k)];
    const double inc = gnx_group_incl_scan(Fk, gw);
    const double c1 = carry + inc;               // cumF_{k+1}
    const double c0 = c1 - Fk;                   // cumF_k
    s2 += hk * (c0 + c1) * 0.5;
    sg += Gk;
    len += hk;
    carry += __shfl_sync(GNX_FULL, inc, gw - 1, gw);
  }
  if (gl == 0) sg += b[f.qdof(net.loc, m)];
  s2 = gnx_group_sum(s2, gw);
  sg = gnx_group_sum(sg, gw);
  len = gnx_group_sum(len, gw);
  if (active && gl == 0) {
    const double a = a_edge[e];
    cond[e] = 1.0 / (a * len);
    rsrc[e] = sg - a * s2;
    ftot[e] = carry;
  }
}

extern "C" int gnx_edge_eliminate_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                                        const double* b, double* cond, double* rsrc, double* ftot,
                                        void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && b && cond && rsrc && ftot && net->h && co->sigma != 0.0);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  const int gw = group_width(net->n);
  const int64_t threads = (int64_t)net->E * gw;
  edge_eliminate_kernel<<<gnx_blocks(threads, 128), 128, 0, (cudaStream_t)stream>>>(
      *net, d, co->a_edge, 1.0 / co->sigma, b, gw, cond, rsrc, ftot);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
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
// Jacobi-PCG on S P = r.  Work layout (doubles):
//   r[B] z[B] w[B] p0[B] p1[B] | part: pw[NB] rz0[NB] rz1[NB] rr0[NB] rr1[NB] | scal[8]
//   scal: 0 tol2, 1 converged(0/1), 2 iterations, 3 last rr, 4 bb
// ------------------------------------------------------------------------------------
constexpr int CG_T = 256;

struct CgPtrs {
  int32_t B, nb;
  const int32_t* bif_nodes; const int32_t* inc_ptr; const int32_t* scol;
  const double* sdiag; const double* soff; const double* rhs;
  double *x, *r, *z, *w, *p0, *p1, *pw, *rz0, *rz1, *rr0, *rr1, *scal;
};

extern "C" int64_t gnx_cg_work_doubles(int32_t B) {
  const int64_t nb = (B + CG_T - 1) / CG_T;
  return 5 * (int64_t)B + 5 * nb + 8;
}

__global__ void __launch_bounds__(CG_T) cg_init_kernel(CgPtrs c) {
  __shared__ double sh[32];
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0.0, rr = 0.0, bb = 0.0;
  if (i < c.B) {
    const int32_t node = c.bif_nodes[i];
    double ax = c.sdiag[i] * c.x[i];
    for (int32_t k = c.inc_ptr[node]; k < c.inc_ptr[node + 1]; ++k) {
      const int32_t j = c.scol[k];
      if (j >= 0) ax += c.soff[k] * c.x[j];
    }
    const double ri = c.rhs[i] - ax;
    const double zi = ri / c.sdiag[i];
    c.r[i] = ri; c.z[i] = zi; c.p0[i] = 0.0; c.p1[i] = 0.0; c.w[i] = 0.0;
    rz = ri * zi; rr = ri * ri; bb = c.rhs[i] * c.rhs[i];
  }
  rz = gnx_block_sum<false>(rz, sh);
  rr = gnx_block_sum<false>(rr, sh);
  bb = gnx_block_sum<false>(bb, sh);
  if (threadIdx.x == 0) {
    c.rz0[blockIdx.x] = rz; c.rz1[blockIdx.x] = rz;   // beta = 1 with p_old = 0 on iteration 0
    c.rr0[blockIdx.x] = rr; c.rr1[blockIdx.x] = rr;
    c.pw[blockIdx.x] = bb;                             // borrowed for ||rhs||^2 until finalize
  }
}

__global__ void __launch_bounds__(CG_T) cg_init_finalize_kernel(CgPtrs c, double rtol) {
  __shared__ double sh[32];
  const double bb = gnx_reduce_partials(c.pw, c.nb, sh);
  const double rr = gnx_reduce_partials(c.rr0, c.nb, sh);
  if (threadIdx.x == 0) {
    c.scal[0] = rtol * rtol * bb;
    c.scal[1] = (rr <= rtol * rtol * bb) ? 1.0 : 0.0;
    c.scal[2] = 0.0;
    c.scal[3] = rr;
    c.scal[4] = bb;
  }
}

// K1: p_new = z + beta p_old ; w = S p_new ; partial <p_new, w>
template <int PAR>
__global__ void __launch_bounds__(CG_T) cg_spmv_kernel(CgPtrs c) {
  __shared__ double sh[32];
  const double* rz_cur = PAR ? c.rz1 : c.rz0;
  const double* rz_prev = PAR ? c.rz0 : c.rz1;
  const double* rr_cur = PAR ? c.rr1 : c.rr0;
  const double* p_old = PAR ? c.p1 : c.p0;
  double* p_new = PAR ? c.p0 : c.p1;
  const double rr = gnx_reduce_partials(rr_cur, c.nb, sh);
  if (rr <= c.scal[0]) return;                       // converged: uniform across the grid
  const double rzc = gnx_reduce_partials(rz_cur, c.nb, sh);
  const double rzp = gnx_reduce_partials(rz_prev, c.nb, sh);
  const double beta = (rzp != 0.0) ? rzc / rzp : 0.0;
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double pw = 0.0;
  if (i < c.B) {
    const int32_t node = c.bif_nodes[i];
    const double pi = c.z[i] + beta * p_old[i];
    double wi = c.sdiag[i] * pi;
    for (int32_t k = c.inc_ptr[node]; k < c.inc_ptr[node + 1]; ++k) {
      const int32_t j = c.scol[k];
      if (j >= 0) wi += c.soff[k] * (c.z[j] + beta * p_old[j]);
    }
    p_new[i] = pi;
    c.w[i] = wi;
    pw = pi * wi;
  }
  pw = gnx_block_sum<false>(pw, sh);
  if (threadIdx.x == 0) c.pw[blockIdx.x] = pw;
}

// K2: alpha = rz/pw ; x += alpha p ; r -= alpha w ; z = r/diag ; partials <r,z>, <r,r>
template <int PAR>
__global__ void __launch_bounds__(CG_T) cg_update_kernel(CgPtrs c) {
  __shared__ double sh[32];
  const double* rz_cur = PAR ? c.rz1 : c.rz0;
  double* rz_next = PAR ? c.rz0 : c.rz1;
  const double* rr_cur = PAR ? c.rr1 : c.rr0;
  double* rr_next = PAR ? c.rr0 : c.rr1;
  const double* p_new = PAR ? c.p0 : c.p1;
  const double rr = gnx_reduce_partials(rr_cur, c.nb, sh);
  if (rr <= c.scal[0]) {
    if (blockIdx.x == 0 && threadIdx.x == 0) { c.scal[1] = 1.0; c.scal[3] = rr; }
    // keep both parities of the partials equal so later (no-op) launches still see convergence
    if (threadIdx.x == 0) rr_next[blockIdx.x] = rr_cur[blockIdx.x];
    return;
  }
  const double rzc = gnx_reduce_partials(rz_cur, c.nb, sh);
  const double pw = gnx_reduce_partials(c.pw, c.nb, sh);
  const double alpha = (pw != 0.0) ? rzc / pw : 0.0;
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  double rz = 0.0, r2 = 0.0;
  if (i < c.B) {
    c.x[i] += alpha * p_new[i];
    const double ri = c.r[i] - alpha * c.w[i];
    const double zi = ri / c.sdiag[i];
    c.r[i] = ri; c.z[i] = zi;
    rz = ri * zi; r2 = ri * ri;
  }
  rz = gnx_block_sum<false>(rz, sh);
  r2 = gnx_block_sum<false>(r2, sh);
  if (threadIdx.x == 0) {
    rz_next[blockIdx.x] = rz;
    rr_next[blockIdx.x] = r2;
    if (blockIdx.x == 0) c.scal[2] += 1.0;
  }
}

// final bookkeeping: scal[3] = current ||r||^2, scal[1] = converged
__global__ void __launch_bounds__(CG_T) cg_status_kernel(CgPtrs c, int par) {
  __shared__ double sh[32];
  const double rr = gnx_reduce_partials(par ? c.rr1 : c.rr0, c.nb, sh);
  if (threadIdx.x == 0) { c.scal[3] = rr; c.scal[1] = (rr <= c.scal[0]) ? 1.0 : 0.0; }
}


// ------------------------------------------------------------------------------------
// Persistent PCG: the whole solve is ONE kernel (block / grid barriers instead of launches,
// convergence decided on device, nothing returns to the host).  Same fused math as the
// two-kernel version; reductions are fixed-order butterflies, bitwise identical in every
// thread, so every control decision is grid-uniform.
// ------------------------------------------------------------------------------------
constexpr int CGP_T = 1024;

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

// Single CTA, B <= ROWS*1024: p lives in shared memory, all per-row state and the first
// CGS_DEG neighbours of every row in registers; 3 barriers per iteration.
constexpr int CGS_DEG = 4;

template <int ROWS>
__global__ void __launch_bounds__(CGP_T)
cg_small_kernel(CgPtrs c, double rtol, int max_iter) {
  extern __shared__ double sp[];             // [B] search direction
  __shared__ double red[2][3 * 32];
  const int tid = threadIdx.x;
  int32_t ncol[ROWS][CGS_DEG];
  double nval[ROWS][CGS_DEG];
  int32_t k_over0[ROWS], k_over1[ROWS];      // neighbours beyond CGS_DEG stay in global memory
  double dg[ROWS], xr[ROWS], rr_[ROWS], zr[ROWS], pr[ROWS];
  bool act[ROWS];
  double acc[3] = {0.0, 0.0, 0.0};
  // x must be visible block-wide for the initial residual: stage it in sp
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const int32_t i = tid + q * CGP_T;
    act[q] = i < c.B;
    if (act[q]) sp[i] = c.x[i];
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < ROWS; ++q) {
    const int32_t i = tid + q * CGP_T;
    dg[q] = 1.0; xr[q] = 0.0; rr_[q] = 0.0; zr[q] = 0.0; pr[q] = 0.0;
    k_over0[q] = k_over1[q] = 0;
#pragma unroll
    for (int d = 0; d < CGS_DEG; ++d) { ncol[q][d] = 0; nval[q][d] = 0.0; }
    if (act[q]) {
      const int32_t node = c.bif_nodes[i];
      const int32_t k0 = c.inc_ptr[node], k1 = c.inc_ptr[node + 1];
      dg[q] = c.sdiag[i];
      xr[q] = sp[i];
      double ax = dg[q] * xr[q];
#pragma unroll
      for (int d = 0; d < CGS_DEG; ++d) {
        if (k0 + d < k1) {
          const int32_t j = c.scol[k0 + d];
          if (j >= 0) { ncol[q][d] = j; nval[q][d] = c.soff[k0 + d]; ax += nval[q][d] * sp[j]; }
        }
      }
      k_over0[q] = min(k0 + CGS_DEG, k1); k_over1[q] = k1;
      for (int32_t k = k_over0[q]; k < k_over1[q]; ++k) {
        const int32_t j = c.scol[k];
        if (j >= 0) ax += c.soff[k] * sp[j];
      }
      const double bi = c.rhs[i];
      rr_[q] = bi - ax;
      zr[q] = rr_[q] / dg[q];
      acc[0] += rr_[q] * zr[q]; acc[1] += rr_[q] * rr_[q]; acc[2] += bi * bi;
    }
  }
  __syncthreads();                           // everyone is done reading x from sp
  block_allreduce<3>(acc, red[0]);
  double rzc = acc[0], rzp = acc[0], rr = acc[1];
  const double bb = acc[2];
  const double tol2 = rtol * rtol * bb;
  int it = 0;
  for (; it < max_iter; ++it) {
    if (rr <= tol2) break;
    const double beta = (it == 0 || rzp == 0.0) ? 0.0 : rzc / rzp;
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      pr[q] = zr[q] + beta * pr[q];
      if (act[q]) sp[tid + q * CGP_T] = pr[q];
    }
    __syncthreads();
    double pw[1] = {0.0};
    double wr[ROWS];
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      double wi = dg[q] * pr[q];
#pragma unroll
      for (int d = 0; d < CGS_DEG; ++d) wi += nval[q][d] * sp[ncol[q][d]];
      for (int32_t k = k_over0[q]; k < k_over1[q]; ++k) {
        const int32_t j = c.scol[k];
        if (j >= 0) wi += c.soff[k] * sp[j];
      }
      wr[q] = act[q] ? wi : 0.0;
      pw[0] += pr[q] * wr[q];
    }
    block_allreduce<1>(pw, red[1]);          // its barrier also fences the sp reads above
    const double alpha = (pw[0] != 0.0) ? rzc / pw[0] : 0.0;
    double a2[2] = {0.0, 0.0};
#pragma unroll
    for (int q = 0; q < ROWS; ++q) {
      xr[q] += alpha * pr[q];
      rr_[q] -= alpha * wr[q];
      zr[q] = rr_[q] / dg[q];
      a2[0] += rr_[q] * zr[q]; a2[1] += rr_[q] * rr_[q];
    }
    block_allreduce<2>(a2, red[0]);
    rzp = rzc; rzc = a2[0]; rr = a2[1];
  }
#pragma unroll
  for (int q = 0; q < ROWS; ++q) if (act[q]) c.x[tid + q * CGP_T] = xr[q];
  if (tid == 0) {
    c.scal[0] = tol2; c.scal[1] = (rr <= tol2) ? 1.0 : 0.0; c.scal[2] = (double)it;
    c.scal[3] = rr; c.scal[4] = bb;
  }
}

// Multi-CTA cooperative version: vectors stay in global memory (L2 resident), two grid barriers
// per iteration; partials of every block are re-reduced by every block in the same order.
template <int K>
__device__ __forceinline__ void grid_allreduce(double (&v)[K], double* const (&part)[K], int nblk,
                                               cg::grid_group& grid, double* sh) {
  block_allreduce<K>(v, sh);
  if (nblk == 1) return;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) part[k][blockIdx.x] = v[k];
  }
  grid.sync();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    double t = 0.0;
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) t += __ldcg(part[k] + i);
    v[k] = t;
  }
  __syncthreads();                            // sh reuse
  block_allreduce<K>(v, sh);
}

__global__ void __launch_bounds__(CGP_T)
cg_persistent_kernel(CgPtrs c, double rtol, int max_iter) {
  __shared__ double red[2][3 * 32];
  cg::grid_group grid = cg::this_grid();
  const int nblk = gridDim.x;
  const int32_t stride = nblk * blockDim.x;
  const int32_t t0 = blockIdx.x * blockDim.x + threadIdx.x;
  double acc[3] = {0.0, 0.0, 0.0};
  for (int32_t i = t0; i < c.B; i += stride) {
    const int32_t node = c.bif_nodes[i];
    double ax = c.sdiag[i] * c.x[i];
    for (int32_t k = c.inc_ptr[node]; k < c.inc_ptr[node + 1]; ++k) {
      const int32_t j = c.scol[k];
      if (j >= 0) ax += c.soff[k] * c.x[j];
    }
    const double bi = c.rhs[i];
    const double ri = bi - ax;
    const double zi = ri / c.sdiag[i];
    c.r[i] = ri; c.z[i] = zi; c.p0[i] = 0.0; c.p1[i] = 0.0;
    acc[0] += ri * zi; acc[1] += ri * ri; acc[2] += bi * bi;
  }
  {
    double* const part[3] = {c.rz0, c.rr0, c.rz1};
    grid_allreduce<3>(acc, part, nblk, grid, red[0]);
  }
  double rzc = acc[0], rzp = acc[0], rr = acc[1];
  const double bb = acc[2];
  const double tol2 = rtol * rtol * bb;
  int it = 0;
  for (; it < max_iter; ++it) {
    if (rr <= tol2) break;
    const double beta = (it == 0 || rzp == 0.0) ? 0.0 : rzc / rzp;
    const double* p_old = (it & 1) ? c.p1 : c.p0;
    double* p_new = (it & 1) ? c.p0 : c.p1;
    // ---- phase A: p = z + beta p_old ; w = S p ; <p,w>
    double pw[1] = {0.0};
    for (int32_t i = t0; i < c.B; i += stride) {
      const int32_t node = c.bif_nodes[i];
      const double pi = c.z[i] + beta * p_old[i];
      double wi = c.sdiag[i] * pi;
      for (int32_t k = c.inc_ptr[node]; k < c.inc_ptr[node + 1]; ++k) {
        const int32_t j = c.scol[k];
        if (j >= 0) wi += c.soff[k] * (__ldcg(c.z + j) + beta * __ldcg(p_old + j));
      }
      p_new[i] = pi;
      c.w[i] = wi;
      pw[0] += pi * wi;
    }
    {
      double* const part[1] = {c.pw};
      grid_allreduce<1>(pw, part, nblk, grid, red[1]);
    }
    const double alpha = (pw[0] != 0.0) ? rzc / pw[0] : 0.0;
    // ---- phase B: x += alpha p ; r -= alpha w ; z = r / diag ; <r,z>, <r,r>
    double a2[2] = {0.0, 0.0};
    for (int32_t i = t0; i < c.B; i += stride) {
      c.x[i] += alpha * p_new[i];
      const double ri = c.r[i] - alpha * c.w[i];
      const double zi = ri / c.sdiag[i];
      c.r[i] = ri; c.z[i] = zi;
      a2[0] += ri * zi; a2[1] += ri * ri;
    }
    {
      double* const part[2] = {(it & 1) ? c.rz1 : c.rz0, (it & 1) ? c.rr1 : c.rr0};
      grid_allreduce<2>(a2, part, nblk, grid, red[0]);
    }
    rzp = rzc; rzc = a2[0]; rr = a2[1];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    c.scal[0] = tol2; c.scal[1] = (rr <= tol2) ? 1.0 : 0.0; c.scal[2] = (double)it;
    c.scal[3] = rr; c.scal[4] = bb;
  }
}

struct gnx_cg {
  CgPtrs c;
  int batch;                 // iterations per graph launch (even)
  int mode;                  // 0: kernel pairs, 1: CUDA-graph batches, 2: persistent cooperative kernel
  int pgrid;                 // grid of the persistent kernel
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  double* scal_host;         // pinned [8]
};

static void cg_enqueue_iters(const CgPtrs& c, int iters, cudaStream_t st) {
  for (int it = 0; it < iters; ++it) {
    if ((it & 1) == 0) {
      cg_spmv_kernel<0><<<c.nb, CG_T, 0, st>>>(c);
      cg_update_kernel<0><<<c.nb, CG_T, 0, st>>>(c);
    } else {
      cg_spmv_kernel<1><<<c.nb, CG_T, 0, st>>>(c);
      cg_update_kernel<1><<<c.nb, CG_T, 0, st>>>(c);
    }
  }
}

extern "C" int gnx_cg_create(const gnx_network_t* net, const double* sdiag, const double* soff,
                             const int32_t* scol, const double* srhs, double* x, double* work,
                             int32_t batch, int32_t mode, void* stream, gnx_cg** out) {
  GNX_CHECK_ARG(net && sdiag && soff && scol && srhs && x && work && out && net->B > 0);
  gnx_cg* h = new gnx_cg();
  const int32_t B = net->B;
  const int32_t nb = (B + CG_T - 1) / CG_T;
  CgPtrs& c = h->c;
  c.B = B; c.nb = nb;
  c.bif_nodes = net->bif_nodes; c.inc_ptr = net->inc_ptr; c.scol = scol;
  c.sdiag = sdiag; c.soff = soff; c.rhs = srhs; c.x = x;
  double* w = work;
  c.r = w; w += B; c.z = w; w += B; c.w = w; w += B; c.p0 = w; w += B; c.p1 = w; w += B;
  c.pw = w; w += nb; c.rz0 = w; w += nb; c.rz1 = w; w += nb; c.rr0 = w; w += nb; c.rr1 = w; w += nb;
  c.scal = w;
  (void)stream;
  h->batch = (batch < 2 ? 2 : batch) & ~1;
  h->graph = nullptr; h->exec = nullptr; h->scal_host = nullptr;
  h->mode = mode; h->pgrid = 0;
  const int use_graph = (mode == 1);
  if (mode == 2) {
    int dev = 0, nsm = 0, per_sm = 0, coop = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cg_persistent_kernel, CGP_T, 0);
    if (!coop || per_sm < 1) { gnx_set_error("cooperative launch unsupported"); delete h; return GNX_ERR_CUDA; }
    int want = (B + CGP_T - 1) / CGP_T;
    if (want > nsm * per_sm) want = nsm * per_sm;
    if (want > nb) want = nb;
    h->pgrid = want < 1 ? 1 : want;
  }
  cudaError_t err = cudaMallocHost(&h->scal_host, 8 * sizeof(double));
  if (err != cudaSuccess) { gnx_set_error("cudaMallocHost: %s", cudaGetErrorString(err)); delete h; return GNX_ERR_CUDA; }
  if (use_graph) {
    // capture on a private stream: the caller's stream may be the legacy default stream, which
    // cannot be captured.  The instantiated graph is launched on the caller's stream.
    cudaStream_t cs = nullptr;
    err = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    if (err == cudaSuccess) {
      cg_enqueue_iters(c, h->batch, cs);
      err = cudaStreamEndCapture(cs, &h->graph);
    }
    if (err == cudaSuccess) err = cudaGraphInstantiate(&h->exec, h->graph, 0);
    if (cs) cudaStreamDestroy(cs);
    if (err != cudaSuccess) {
      gnx_set_error("CUDA graph capture of the CG batch failed: %s", cudaGetErrorString(err));
      cudaGetLastError();
      if (h->graph) cudaGraphDestroy(h->graph);
      cudaFreeHost(h->scal_host);
      delete h;
      return GNX_ERR_CUDA;
    }
  }
  *out = h;
  return GNX_OK;
}

extern "C" void gnx_cg_destroy(gnx_cg* h) {
  if (!h) return;
  if (h->exec) cudaGraphExecDestroy(h->exec);
  if (h->graph) cudaGraphDestroy(h->graph);
  if (h->scal_host) cudaFreeHost(h->scal_host);
  delete h;
}

// x holds the initial guess on entry.  Synchronises the stream.
// resid_host[0] = ||r||/||rhs|| (recurrence residual); info_host = {iterations, converged, graph launches}
extern "C" int gnx_cg_solve(gnx_cg* h, double rtol, int32_t max_iter, double* resid_host,
                            int32_t* info_host, void* stream) {
  GNX_CHECK_ARG(h && rtol > 0 && max_iter > 0);
  const CgPtrs& c = h->c;
  cudaStream_t st = (cudaStream_t)stream;
  if (h->mode == 2) {
    CgPtrs cc = c;
    double rt = rtol;
    int mi = max_iter;
    if (c.B <= CGP_T) {
      cg_small_kernel<1><<<1, CGP_T, sizeof(double) * c.B, st>>>(cc, rt, mi);
      GNX_LAUNCH_CHECK();
    } else if (c.B <= 2 * CGP_T) {
      cg_small_kernel<2><<<1, CGP_T, sizeof(double) * c.B, st>>>(cc, rt, mi);
      GNX_LAUNCH_CHECK();
    } else {
      void* args[] = {&cc, &rt, &mi};
      GNX_CUDA(cudaLaunchCooperativeKernel((void*)cg_persistent_kernel, dim3(h->pgrid), dim3(CGP_T), args, 0, st));
    }
    if (!resid_host && !info_host) return GNX_OK;            // fully asynchronous
    GNX_CUDA(cudaMemcpyAsync(h->scal_host, c.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    GNX_CUDA(cudaStreamSynchronize(st));
    const double bb = h->scal_host[4];
    if (resid_host) resid_host[0] = bb > 0 ? sqrt(h->scal_host[3] / bb) : 0.0;
    const int done = h->scal_host[1] != 0.0;
    if (info_host) { info_host[0] = (int32_t)h->scal_host[2]; info_host[1] = done; info_host[2] = 1; }
    if (!done) { gnx_set_error("schur CG: no convergence in %d iterations (rel. resid %.3e)", (int)h->scal_host[2],
                               bb > 0 ? sqrt(h->scal_host[3] / bb) : 0.0); return GNX_ERR_NOCONV; }
    return GNX_OK;
  }
  cg_init_kernel<<<c.nb, CG_T, 0, st>>>(c);
  cg_init_finalize_kernel<<<1, CG_T, 0, st>>>(c, rtol);
  GNX_LAUNCH_CHECK();
  int launches = 0, done = 0;
  while (true) {
    GNX_CUDA(cudaMemcpyAsync(h->scal_host, c.scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));
    GNX_CUDA(cudaStreamSynchronize(st));
    if (h->scal_host[1] != 0.0 || h->scal_host[4] == 0.0) { done = 1; break; }
    if ((int)h->scal_host[2] >= max_iter) break;
    if (h->exec) {
      GNX_CUDA(cudaGraphLaunch(h->exec, st));
    } else {
      cg_enqueue_iters(c, h->batch, st);
      GNX_LAUNCH_CHECK();
    }
    cg_status_kernel<<<1, CG_T, 0, st>>>(c, 0);      // batch is even -> parity 0 is current
    ++launches;
  }
  if (resid_host) {
    const double bb = h->scal_host[4];
    resid_host[0] = bb > 0 ? sqrt(h->scal_host[3] / bb) : 0.0;
  }
  if (info_host) { info_host[0] = (int32_t)h->scal_host[2]; info_host[1] = done; info_host[2] = launches; }
  if (!done) { gnx_set_error("schur CG: no convergence in %d iterations (rel. resid %.3e)", (int)h->scal_host[2],
                             sqrt(h->scal_host[3] / h->scal_host[4])); return GNX_ERR_NOCONV; }
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// phase 2: back-substitution
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
edge_backsub_kernel(gnx_network_t net, GnxMixedDims d, const double* __restrict__ a_edge,
                    const double* __restrict__ k_edge, double inv_sigma, const double* __restrict__ b,
                    const double* __restrict__ cond, const double* __restrict__ rsrc,
                    const double* __restrict__ P, int gw, double* __restrict__ x) {
  const int64_t gthread = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t e = (int32_t)(gthread / gw);
  const int gl = threadIdx.x & (gw - 1);
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
  double prevF = 0.0, prevH = 1.0;        // cell left of node `base`
  for (int32_t base = 0; base < m; base += gw) {
    const int32_t j = base + gl;          // node j, right cell j
    const double Fj = b[f.pdof(j)] * inv_sigma;
    const double hj = net.h[f.hb + j];
    const double Gj = b[f.qdof(net.loc, j)];
    const double inc = gnx_group_incl_scan(Fj, gw);
    const double qj = q0 + carryF + (inc - Fj);
    double Fl = __shfl_up_sync(GNX_FULL, Fj, 1, gw);
    double hl = __shfl_up_sync(GNX_FULL, hj, 1, gw);
    if (gl == 0) { Fl = prevF; hl = prevH; }
    double mq = a * hj * (3.0 * qj + Fj) * (1.0 / 6.0) - kv * Fj / hj;           // h_j (2 q_j + q_{j+1})/6
    if (j > 0) mq += a * hl * (3.0 * qj - Fl) * (1.0 / 6.0) + kv * Fl / hl;      // h_{j-1}(q_{j-1} + 2 q_j)/6
    const double Dj = Gj - mq;
    const double incD = gnx_group_incl_scan(Dj, gw);
    const double pj = Pu + carryD + incD;
    if (active) {
      x[f.qdof(net.loc, j)] = qj;
      x[f.pdof(j)] = pj * inv_sigma;
    }
    carryF += __shfl_sync(GNX_FULL, inc, gw - 1, gw);
    carryD += __shfl_sync(GNX_FULL, incD, gw - 1, gw);
    prevF = __shfl_sync(GNX_FULL, Fj, gw - 1, gw);
    prevH = __shfl_sync(GNX_FULL, hj, gw - 1, gw);
  }
  if (active && gl == 0) x[f.qdof(net.loc, m)] = q0 + carryF;
}

__global__ void lambda_store_kernel(const double* __restrict__ P, int32_t B, double inv_sigma, double* __restrict__ xl) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) xl[i] = P[i] * inv_sigma;
}

extern "C" int gnx_edge_backsub_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                                      const double* b, const double* cond, const double* rsrc,
                                      const double* P, double* x, void* stream) {
  GNX_CHECK_ARG(net && co && co->a_edge && b && cond && rsrc && x && net->h && co->sigma != 0.0);
  GNX_CHECK_ARG(net->B == 0 || P);
  const GnxMixedDims d = gnx_mixed_dims(*net);
  const int gw = group_width(net->n);
  const int64_t threads = (int64_t)net->E * gw;
  cudaStream_t st = (cudaStream_t)stream;
  edge_backsub_kernel<<<gnx_blocks(threads, 128), 128, 0, st>>>(
      *net, d, co->a_edge, co->k_edge, 1.0 / co->sigma, b, cond, rsrc, P, gw, x);
  if (net->B > 0)
    lambda_store_kernel<<<gnx_blocks(net->B, 256), 256, 0, st>>>(P, net->B, 1.0 / co->sigma, x + d.lam_off);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}
