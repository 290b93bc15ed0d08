// Graph -> mesh kernels: hierarchical bisection, tangents, bifurcation/boundary vertex tables.
// Reference behaviour replaced: graphnics/fenics_graph.py:58-99 (get_mesh), :161-184
// (record_bifurcation_and_boundary_nodes), :230-253 (assign_tangents), :132-156,280-311 (tags).
#include <stdarg.h>
#include "gnx_common.cuh"

// ------------------------------------------------------------------------------------
// error plumbing (shared by all translation units)
// ------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void gnx_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* gnx_last_error(void) { return g_err; }
extern "C" int gnx_version(void) { return 100; }
extern "C" int gnx_device_ok(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { cudaGetLastError(); return 0; }
  return n > 0 ? 1 : 0;
}

// ------------------------------------------------------------------------------------
// refine: one thread per (edge, spatial cell).  Bit-exact recursive midpoints.
// ------------------------------------------------------------------------------------
template <int GD>
__global__ void __launch_bounds__(256)
refine_kernel(const double* __restrict__ pos, const int32_t* __restrict__ edges, int32_t V,
              int32_t E, int32_t n, double* __restrict__ coords, int32_t* __restrict__ cells,
              int32_t* __restrict__ mf, double* __restrict__ h) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t m = 1 << n;
  if (gid < (int64_t)V * GD) coords[gid] = pos[gid];  // graph nodes keep ids 0..V-1
  if (gid >= (int64_t)E * m) return;
  gnx_refine_cell<GD>(pos, edges, V, E, n, gid, coords, cells, mf, h);
}

extern "C" int gnx_refine_1d(const double* pos, const int32_t* edges, int32_t V, int32_t E,
                             int32_t gdim, int32_t n, double* coords, int32_t* cells, int32_t* mf,
                             double* h, void* stream) {
  GNX_CHECK_ARG(pos && edges && coords && cells && h);
  GNX_CHECK_ARG(V > 0 && E > 0 && (gdim == 1 || gdim == 2 || gdim == 3) && n >= 0 && n <= 24);
  const int64_t ncell = (int64_t)E << n;
  GNX_CHECK_ARG(ncell < (1ll << 31) && (int64_t)V + ncell < (1ll << 31));
  const int64_t work = ncell > (int64_t)V * gdim ? ncell : (int64_t)V * gdim;
  cudaStream_t st = (cudaStream_t)stream;
  const int th = 256, bl = gnx_blocks(work, th);
  if (gdim == 1) refine_kernel<1><<<bl, th, 0, st>>>(pos, edges, V, E, n, coords, cells, mf, h);
  else if (gdim == 2) refine_kernel<2><<<bl, th, 0, st>>>(pos, edges, V, E, n, coords, cells, mf, h);
  else refine_kernel<3><<<bl, th, 0, st>>>(pos, edges, V, E, n, coords, cells, mf, h);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// tangents: t = pos[v]-pos[u]; t *= 1.0/||t||   (fenics_graph.py:239-246)
// ||t|| reproduces numpy/BLAS on the golden vectors: s = t0*t0; s = fma(t_i,t_i,s); sqrt(s).
// ------------------------------------------------------------------------------------
template <int GD>
__global__ void tangent_kernel(const double* __restrict__ pos, const int32_t* __restrict__ edges,
                               int32_t E, double* __restrict__ tang, double* __restrict__ elen) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t u = edges[2 * e], v = edges[2 * e + 1];
  double t[GD];
#pragma unroll
  for (int i = 0; i < GD; ++i) t[i] = __dsub_rn(pos[(int64_t)v * GD + i], pos[(int64_t)u * GD + i]);
  double s = __dmul_rn(t[0], t[0]);
#pragma unroll
  for (int i = 1; i < GD; ++i) s = __fma_rn(t[i], t[i], s);
  const double nrm = __dsqrt_rn(s);
  const double inv = __ddiv_rn(1.0, nrm);
#pragma unroll
  for (int i = 0; i < GD; ++i) tang[(int64_t)e * GD + i] = __dmul_rn(t[i], inv);
  if (elen) elen[e] = nrm;
}

extern "C" int gnx_edge_tangents(const double* pos, const int32_t* edges, int32_t E, int32_t gdim,
                                 double* tangents, double* elen, void* stream) {
  GNX_CHECK_ARG(pos && edges && tangents && E > 0 && gdim >= 1 && gdim <= 3);
  cudaStream_t st = (cudaStream_t)stream;
  const int th = 128, bl = gnx_blocks(E, th);
  if (gdim == 1) tangent_kernel<1><<<bl, th, 0, st>>>(pos, edges, E, tangents, elen);
  else if (gdim == 2) tangent_kernel<2><<<bl, th, 0, st>>>(pos, edges, E, tangents, elen);
  else tangent_kernel<3><<<bl, th, 0, st>>>(pos, edges, E, tangents, elen);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// int32 exclusive scan (3 passes: block sums, scan of block sums, rescan with offsets)
// ------------------------------------------------------------------------------------
constexpr int SCAN_T = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_T * SCAN_ITEMS;

__device__ __forceinline__ int32_t block_excl_scan_i32(int32_t v, int32_t* sh, int32_t& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(GNX_FULL, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) sh[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int32_t w = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : 0;
    int32_t winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int32_t t = __shfl_up_sync(GNX_FULL, winc, o);
      if (lane >= o) winc += t;
    }
    sh[lane] = winc - w;           // exclusive warp offsets
    if (lane == 31) sh[32] = winc; // block total
  }
  __syncthreads();
  total = sh[32];
  return inc - v + sh[wid];
}

__global__ void __launch_bounds__(SCAN_T)
scan_block_sums(const int32_t* __restrict__ in, int64_t n, int32_t* __restrict__ bsum) {
  __shared__ int32_t sh[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) if (base + i < n) s += in[base + i];
  int32_t tot;
  block_excl_scan_i32(s, sh, tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_T)
scan_of_sums(int32_t* __restrict__ bsum, int32_t nb, int32_t* __restrict__ total_out) {
  __shared__ int32_t sh[33];
  int32_t carry = 0;
  for (int32_t base = 0; base < nb; base += SCAN_T) {
    const int32_t i = base + threadIdx.x;
    const int32_t v = (i < nb) ? bsum[i] : 0;
    int32_t tot;
    const int32_t ex = block_excl_scan_i32(v, sh, tot);
    if (i < nb) bsum[i] = carry + ex;
    carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_T)
scan_apply(const int32_t* __restrict__ in, int64_t n, const int32_t* __restrict__ bsum,
           int32_t* __restrict__ out) {
  __shared__ int32_t sh[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
  int32_t v[SCAN_ITEMS];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) { v[i] = (base + i < n) ? in[base + i] : 0; s += v[i]; }
  int32_t tot;
  int32_t ex = block_excl_scan_i32(s, sh, tot) + bsum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_ITEMS; ++i) {
    if (base + i < n) out[base + i] = ex;
    ex += v[i];
  }
}

// out[0..n) = exclusive scan of in; *total_out (device) = sum.  in/out may alias.
static int scan_i32(const int32_t* in, int32_t* out, int64_t n, int32_t* total_out, int32_t* bsum,
                    cudaStream_t st) {
  const int nb = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
  scan_block_sums<<<nb, SCAN_T, 0, st>>>(in, n, bsum);
  scan_of_sums<<<1, SCAN_T, 0, st>>>(bsum, nb, total_out);
  scan_apply<<<nb, SCAN_T, 0, st>>>(in, n, bsum, out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// vertex tables
// ------------------------------------------------------------------------------------
__global__ void degree_kernel(const int32_t* __restrict__ edges, int32_t E, int32_t* __restrict__ deg) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  atomicAdd(&deg[edges[2 * e]], 1);      // integer atomics: order-independent result
  atomicAdd(&deg[edges[2 * e + 1]], 1);
}

// `cls` (classification degree) = the degree in the WHOLE graph; `deg` = the degree in this edge
// list.  They differ only for a partition of a larger network (cls_deg != NULL).
__global__ void node_flags_kernel(const int32_t* __restrict__ deg, const int32_t* __restrict__ cls, int32_t V,
                                  int32_t* __restrict__ fbif, int32_t* __restrict__ fbnd,
                                  int32_t* __restrict__ fdegbif, int32_t* __restrict__ flonely) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int32_t d = cls[v];
  fbif[v] = d > 1;
  fbnd[v] = d == 1;
  fdegbif[v] = d > 1 ? deg[v] : 0;
  if (d == 0) atomicAdd(flonely, 1);
}

__global__ void node_compact_kernel(const int32_t* __restrict__ deg, int32_t V,
                                    const int32_t* __restrict__ sbif, const int32_t* __restrict__ sbnd,
                                    const int32_t* __restrict__ sdegbif, int32_t* __restrict__ bix,
                                    int32_t* __restrict__ bif_nodes, int32_t* __restrict__ bnd_nodes,
                                    int32_t* __restrict__ lam_ptr) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int32_t d = deg[v];          // classification degree
  if (d > 1) {
    const int32_t b = sbif[v];
    bix[v] = b;
    bif_nodes[b] = v;
    lam_ptr[b] = sdegbif[v];
  } else {
    bix[v] = -1;
    if (d == 1) bnd_nodes[sbnd[v]] = v;
  }
}

__global__ void set_tail_kernel(int32_t* __restrict__ arr, const int32_t* __restrict__ idx,
                                const int32_t* __restrict__ val) {
  arr[*idx] = *val;  // arr[B] = I
}

__global__ void inc_fill_kernel(const int32_t* __restrict__ edges, int32_t E, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ inc_edge) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t u = edges[2 * e], v = edges[2 * e + 1];
  inc_edge[atomicAdd(&cursor[u], 1)] = (e << 1);
  inc_edge[atomicAdd(&cursor[v], 1)] = (e << 1) | 1;
}

// arrival order of the atomics is arbitrary -> sort every node's (short) list: deterministic result
__global__ void inc_sort_kernel(const int32_t* __restrict__ inc_ptr, int32_t V, int32_t* __restrict__ inc_edge) {
  const int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const int32_t a = inc_ptr[v], b = inc_ptr[v + 1];
  for (int32_t i = a + 1; i < b; ++i) {
    const int32_t key = inc_edge[i];
    int32_t j = i - 1;
    while (j >= a && inc_edge[j] > key) { inc_edge[j + 1] = inc_edge[j]; --j; }
    inc_edge[j + 1] = key;
  }
}

__global__ void edge_tags_kernel(const int32_t* __restrict__ edges, int32_t E, const int32_t* __restrict__ deg,
                                 int32_t* __restrict__ tag_u, int32_t* __restrict__ tag_v,
                                 int32_t* __restrict__ febif) {
  const int32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t du = deg[edges[2 * e]], dv = deg[edges[2 * e + 1]];
  // mark_edge_vfs(bifurcation_ixs, BIF_IN, BIF_OUT); mark_edge_vfs(boundary_ixs, BOUN_OUT, BOUN_IN)
  tag_u[e] = du > 1 ? GNX_BIF_OUT : (du == 1 ? GNX_BOUN_IN : 0);   // edge goes OUT of u
  tag_v[e] = dv > 1 ? GNX_BIF_IN : (dv == 1 ? GNX_BOUN_OUT : 0);   // edge goes IN to v
  febif[e] = (du > 1) + (dv > 1);
}

extern "C" int64_t gnx_tables_work_ints(int32_t V, int32_t E) {
  const int64_t mx = (V > E ? V : E) + 1;
  return 4 * mx + (mx / SCAN_TILE + 2) + 16;
}

extern "C" int gnx_vertex_tables(const int32_t* edges, int32_t V, int32_t E, int32_t* deg,
                                 int32_t* bix, int32_t* bif_nodes, int32_t* bnd_nodes,
                                 int32_t* inc_ptr, int32_t* inc_edge, int32_t* lam_ptr,
                                 int32_t* ebif_prefix, int32_t* tag_u, int32_t* tag_v,
                                 int32_t* counts, int32_t* work, const int32_t* cls_deg, void* stream) {
  GNX_CHECK_ARG(edges && deg && bix && bif_nodes && bnd_nodes && inc_ptr && inc_edge && lam_ptr &&
                ebif_prefix && tag_u && tag_v && counts && work && V > 0 && E > 0);
  GNX_CHECK_ARG((int64_t)E < (1ll << 30));
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t mx = (V > E ? V : E) + 1;
  int32_t* w0 = work;            // flags / scans
  int32_t* w1 = work + mx;
  int32_t* w2 = work + 2 * mx;
  int32_t* w3 = work + 3 * mx;   // cursor
  int32_t* bsum = work + 4 * mx;
  const int th = 256;
  GNX_CUDA(cudaMemsetAsync(deg, 0, sizeof(int32_t) * V, st));
  GNX_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * 4, st));
  degree_kernel<<<gnx_blocks(E, th), th, 0, st>>>(edges, E, deg);
  const int32_t* cls = cls_deg ? cls_deg : deg;
  node_flags_kernel<<<gnx_blocks(V, th), th, 0, st>>>(deg, cls, V, w0, w1, w2, counts + 3);
  GNX_LAUNCH_CHECK();
  if (scan_i32(w0, w0, V, counts + 0, bsum, st)) return GNX_ERR_CUDA;   // B
  if (scan_i32(w1, w1, V, counts + 1, bsum, st)) return GNX_ERR_CUDA;   // n_bnd
  if (scan_i32(w2, w2, V, counts + 2, bsum, st)) return GNX_ERR_CUDA;   // I
  node_compact_kernel<<<gnx_blocks(V, th), th, 0, st>>>(cls, V, w0, w1, w2, bix, bif_nodes, bnd_nodes, lam_ptr);
  set_tail_kernel<<<1, 1, 0, st>>>(lam_ptr, counts + 0, counts + 2);
  // incidence CSR over all nodes
  if (scan_i32(deg, inc_ptr, V, inc_ptr + V, bsum, st)) return GNX_ERR_CUDA;
  GNX_CUDA(cudaMemcpyAsync(w3, inc_ptr, sizeof(int32_t) * V, cudaMemcpyDeviceToDevice, st));
  inc_fill_kernel<<<gnx_blocks(E, th), th, 0, st>>>(edges, E, w3, inc_edge);
  inc_sort_kernel<<<gnx_blocks(V, th), th, 0, st>>>(inc_ptr, V, inc_edge);
  // per-edge tags and the prefix of bifurcation ends
  edge_tags_kernel<<<gnx_blocks(E, th), th, 0, st>>>(edges, E, cls, tag_u, tag_v, w0);
  GNX_LAUNCH_CHECK();
  if (scan_i32(w0, ebif_prefix, E, ebif_prefix + E, bsum, st)) return GNX_ERR_CUDA;
  return GNX_OK;
}
