// Krylov building blocks: bandwidth-bound fp64 CSR SpMV (row-block streaming through shared
// memory), fused residual + norms, fused axpby + dot.  Deterministic two-stage reductions.
#include <stdlib.h>
#include "gnx_common.cuh"

constexpr int SPMV_T = 256;          // threads per block

// acc[j] = sum_k A[r,k] x[k] for row r = r0 + threadIdx.x + j*SPMV_T (0 beyond nrows); a block owns
// SPMV_T*RPT consecutive rows.  The block's contiguous (col, val) range is streamed with coalesced
// loads, U pairs per thread and pass (2U independent loads in flight), x is gathered, the products are
// staged in shared memory and every thread sums the short segments of its rows.
template <int RPT, int U>
__device__ __forceinline__ void spmv_row_block(int32_t nrows, const int32_t* __restrict__ rowptr,
                                               const int32_t* __restrict__ colidx,
                                               const double* __restrict__ vals,
                                               const double* __restrict__ x, double* prod,
                                               double (&acc)[RPT]) {
  constexpr int TILE = SPMV_T * U;
  const int32_t r0 = blockIdx.x * (SPMV_T * RPT);
  const int32_t r1 = min(r0 + SPMV_T * RPT, nrows);
  const int32_t nz0 = rowptr[r0], nz1 = rowptr[r1];
  int32_t a[RPT], b[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int32_t r = r0 + threadIdx.x + j * SPMV_T;
    a[j] = b[j] = 0;
    if (r < nrows) { a[j] = rowptr[r]; b[j] = rowptr[r + 1]; }
    acc[j] = 0.0;
  }
  for (int32_t ts = nz0; ts < nz1; ts += TILE) {
    int32_t c[U];
    double v[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const int32_t i = ts + k * SPMV_T + threadIdx.x;
      const bool ok = i < nz1;
      c[k] = ok ? __ldg(colidx + i) : -1;
      v[k] = ok ? __ldg(vals + i) : 0.0;
    }
#pragma unroll
    for (int k = 0; k < U; ++k) {
      const double xv = (c[k] >= 0) ? __ldg(x + c[k]) : 0.0;
      prod[k * SPMV_T + threadIdx.x] = v[k] * xv;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const int32_t lo = max(a[j], ts), hi = min(b[j], ts + TILE);
      for (int32_t i = lo; i < hi; ++i) acc[j] += prod[i - ts];
    }
    __syncthreads();
  }
}

template <int RPT, int U>
__global__ void __launch_bounds__(SPMV_T)
spmv_kernel(int32_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
            const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y) {
  __shared__ double prod[SPMV_T * U];
  double acc[RPT];
  spmv_row_block<RPT, U>(nrows, rowptr, colidx, vals, x, prod, acc);
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int32_t r = blockIdx.x * (SPMV_T * RPT) + threadIdx.x + j * SPMV_T;
    if (r < nrows) y[r] = acc[j];
  }
}

static int spmv_variant() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("GNX_SPMV_VARIANT"); v = e ? atoi(e) : 0; }   // tuning knob
  return v;
}

extern "C" int gnx_spmv_csr_f64(int32_t nrows, const int32_t* rowptr, const int32_t* colidx,
                                const double* vals, const double* x, double* y, void* stream) {
  GNX_CHECK_ARG(nrows > 0 && rowptr && colidx && vals && x && y);
  cudaStream_t st = (cudaStream_t)stream;
  switch (spmv_variant()) {
    case 1: spmv_kernel<1, 8><<<gnx_blocks(nrows, SPMV_T), SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, y); break;
    case 2: spmv_kernel<2, 8><<<gnx_blocks(nrows, 2 * SPMV_T), SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, y); break;
    case 3: spmv_kernel<4, 16><<<gnx_blocks(nrows, 4 * SPMV_T), SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, y); break;
    case 4: spmv_kernel<4, 8><<<gnx_blocks(nrows, 4 * SPMV_T), SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, y); break;
    case 5: spmv_kernel<2, 4><<<gnx_blocks(nrows, 2 * SPMV_T), SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, y); break;
    // default: 256 rows x ~4 nnz = one 1024-product pass, 32 registers, 8 KB -> 8 CTAs/SM.  Measured on
    // B200 (tools/tune_spmv.py, 135 M rows): 5.14 TB/s = 79 % of the measured HBM peak; the deeper-unrolled
    // variants lose more to occupancy than they gain in loads in flight (4.1-5.1 TB/s).
    default: spmv_kernel<1, 4><<<gnx_blocks(nrows, SPMV_T), SPMV_T, 0, st>>>(nrows, rowptr, colidx, vals, x, y); break;
  }
  GNX_LAUNCH_CHECK();
  return GNX_OK;
}

constexpr int RES_RPT = 1, RES_U = 4;
constexpr int SPMV_ROWS = SPMV_T * RES_RPT;
constexpr int SPMV_RPT = RES_RPT;
constexpr int SPMV_TILE = SPMV_T * RES_U;

extern "C" int64_t gnx_reduce_work_doubles(int64_t n) { return 2 * ((n + SPMV_ROWS - 1) / SPMV_ROWS) + 2; }

__global__ void __launch_bounds__(SPMV_T)
residual_kernel(int32_t nrows, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                const double* __restrict__ vals, const double* __restrict__ x,
                const double* __restrict__ b, double* __restrict__ res, double* __restrict__ part) {
  __shared__ double prod[SPMV_TILE];
  __shared__ double sh[32];
  double acc[SPMV_RPT];
  spmv_row_block<RES_RPT, RES_U>(nrows, rowptr, colidx, vals, x, prod, acc);
  double rr = 0.0, bb = 0.0;
#pragma unroll
  for (int j = 0; j < SPMV_RPT; ++j) {
    const int32_t r = blockIdx.x * SPMV_ROWS + threadIdx.x + j * SPMV_T;
    if (r < nrows) {
      const double bi = b[r];
      const double ri = bi - acc[j];
      if (res) res[r] = ri;
      rr += ri * ri; bb += bi * bi;
    }
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
  const int nb = gnx_blocks(nrows, SPMV_ROWS);
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
