// Krylov building blocks: bandwidth-bound fp64 CSR SpMV (row-block streaming through shared
// memory), fused residual + norms, fused axpby + dot.  Deterministic two-stage reductions.
#include "gnx_common.cuh"

constexpr int SPMV_T = 256;          // threads == rows per block
constexpr int SPMV_TILE = 2048;      // products staged per pass (16 KB)
constexpr int SPMV_U = SPMV_TILE / SPMV_T;

// acc = sum_j A[r,j] x[j] for the row owned by this thread (r may be >= nrows: acc = 0)
__device__ __forceinline__ double spmv_row_block(int32_t nrows, const int32_t* __restrict__ rowptr,
                                                 const int32_t* __restrict__ colidx,
                                                 const double* __restrict__ vals,
                                                 const double* __restrict__ x, double* prod) {
  const int32_t r0 = blockIdx.x * SPMV_T;
  const int32_t r1 = min(r0 + SPMV_T, nrows);
  const int32_t r = r0 + threadIdx.x;
  const int32_t nz0 = rowptr[r0], nz1 = rowptr[r1];
  int32_t a = 0, b = 0;
  if (r < nrows) { a = rowptr[r]; b = rowptr[r + 1]; }
  double acc = 0.0;
  for (int32_t ts = nz0; ts < nz1; ts += SPMV_TILE) {
    // coalesced stream of (val, col), gather x, stage the products
    int32_t c[SPMV_U];
    double v[SPMV_U];
#pragma unroll
    for (int k = 0; k < SPMV_U; ++k) {
      const int32_t i = ts + k * SPMV_T + threadIdx.x;
      const bool ok = i < nz1;
      c[k] = ok ? __ldg(colidx + i) : -1;
      v[k] = ok ? __ldg(vals + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < SPMV_U; ++k) {
      const double xv = (c[k] >= 0) ? __ldg(x + c[k]) : 0.0;
      prod[k * SPMV_T + threadIdx.x] = v[k] * xv;
    }
    __syncthreads();
    const int32_t lo = max(a, ts), hi = min(b, ts + SPMV_TILE);
    for (int32_t i = lo; i < hi; ++i) acc += prod[i - ts];
    __syncthreads();
  }
  return acc;
}

__global__ void __launch_bounds__(SPMV_T)
spmv_kernel(int32_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
            const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double prod[SPMV_TILE];
  const double acc = spmv_row_block(nrows, rowptr, colidx, vals, x, prod);
  const int32_t r = blockIdx.x * SPMV_T + threadIdx.x;
  if (r < nrows) y[r] = acc;
}

extern "C" int gnx_spmv_csr_f64(int32_t nrows, const int32_t* rowptr, const int32_t* colidx,
                                const double* vals, const double* x, double* y, void* stream) {
  GNX_CHECK_ARG(nrows > 0 && rowptr && colidx && vals && x && y);
  spmv_kernel<<<gnx_blocks(nrows, SPMV_T), SPMV_T, 0, (cudaStream_t)stream>>>(nrows, rowptr, colidx, vals, x, y);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

extern "C" int64_t gnx_reduce_work_doubles(int64_t n) { return 2 * ((n + SPMV_T - 1) / SPMV_T) + 2; }

__global__ void __launch_bounds__(SPMV_T)
residual_kernel(int32_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ vals, const double* __restrict__ x,
                const double* __restrict__ b, double* __restrict__ res, double* __restrict__ part) {
  __shared__ double prod[SPMV_TILE];
  __shared__ double sh[32];
  const double acc = spmv_row_block(nrows, rowptr, colidx, vals, x, prod);
  const int32_t r = blockIdx.x * SPMV_T + threadIdx.x;
  double rr = 0.0, bb = 0.0;
  if (r < nrows) {
    const double bi = b[r];
    const double ri = bi - acc;
    if (res) res[r] = ri;
    rr = ri * ri; bb = bi * bi;
  }
  rr = gnx_block_sum<false>(rr, sh);
  bb = gnx_block_sum<false>(bb, sh);
  if (threadIdx.x == 0) { part[2 * blockIdx.x] = rr; part[2 * blockIdx.x + 1] = bb; }
}

// out[c] = sum_i part[ncomp*i + c]   (single block, fixed order)
__global__ void __launch_bounds__(256)
final_reduce_kernel(const double* __restrict__ part, int32_t nb, int ncomp, double* __restrict__ out) {
  __shared__ double sh[32];
  for (int c = 0; c < ncomp; ++c) {
    double v = 0.0;
    for (int32_t i = threadIdx.x; i < nb; i += blockDim.x) v += part[(int64_t)ncomp * i + c];
    v = gnx_block_sum<false>(v, sh);
    if (threadIdx.x == 0) out[c] = v;
    __syncthreads();
  }
}

extern "C" int gnx_residual_csr_f64(int32_t nrows, const int32_t* rowptr, const int32_t* colidx,
                                    const double* vals, const double* x, const double* b, double* r,
                                    double* out2, double* work, void* stream) {
  GNX_CHECK_ARG(nrows > 0 && rowptr && colidx && vals && x && b && out2 && work);
  cudaStream_t st = (cudaStream_t)stream;
  const int nb = gnx_blocks(nrows, SPMV_T);
  residual_kernel<<<nb, SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, b, r, work);
  final_reduce_kernel<<<1, 256, 0, st>>>(work, nb, 2, out2);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

__global__ void __launch_bounds__(256)
axpby_dot_kernel(int64_t n, double alpha, const double* __restrict__ x, double beta,
                 double* __restrict__ y, double* __restrict__ part) {
  __shared__ double sh[32];
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double xi = x[i];
    const double yi = alpha * xi + (beta != 0.0 ? beta * y[i] : 0.0);
    y[i] = yi;
    acc += xi * yi;
  }
  if (part) {
    acc = gnx_block_sum<false>(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = acc;
  }
}

extern "C" int gnx_axpby_dot(int64_t n, double alpha, const double* x, double beta, double* y,
                             double* out, double* work, void* stream) {
  GNX_CHECK_ARG(n > 0 && x && y && (!out || work));
  cudaStream_t st = (cudaStream_t)stream;
  int nb = gnx_blocks(n, 256 * 4);
  if (nb > 148 * 8) nb = 148 * 8;
  axpby_dot_kernel<<<nb, 256, 0, st>>>(n, alpha, x, beta, y, out ? work : nullptr);
  if (out) final_reduce_kernel<<<1, 256, 0, st>>>(work, nb, 1, out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

// out = sum_k c[k] * v[k]  (k < nterms <= 4; out may alias v[0]) -- the theta-scheme right-hand side
//   b = DDn + cn1 dt Ln1 - dt cn An x0 + dt cn Ln      (time_stepping.py:185)
struct Lincomb { double c[4]; const double* v[4]; int n; };

__global__ void __launch_bounds__(256)
lincomb_kernel(int64_t n, Lincomb L, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double acc = L.c[0] * L.v[0][i];
#pragma unroll
    for (int k = 1; k < 4; ++k) if (k < L.n) acc += L.c[k] * L.v[k][i];
    out[i] = acc;
  }
}

extern "C" int gnx_lincomb(int64_t n, int32_t nterms, const double* coefs_host, const double* const* vecs_host,
                           double* out, void* stream) {
  GNX_CHECK_ARG(n > 0 && nterms >= 1 && nterms <= 4 && coefs_host && vecs_host && out);
  Lincomb L;
  L.n = nterms;
  for (int k = 0; k < 4; ++k) {
    L.c[k] = k < nterms ? coefs_host[k] : 0.0;
    L.v[k] = k < nterms ? vecs_host[k] : nullptr;
    GNX_CHECK_ARG(k >= nterms || L.v[k]);
  }
  int nb = gnx_blocks(n, 256 * 2);
  if (nb > 148 * 16) nb = 148 * 16;
  lincomb_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(n, L, out);
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}
