// Edge-tile assembly of the mixed CSR values: the per-row device code shared by mixed_values_tile_kernel
// (gnx_assemble.cu) and the persistent step kernel (gnx_solve.cu).
#ifndef GNX_ASM_TILE_CUH
#define GNX_ASM_TILE_CUH

#include "gnx_common.cuh"

constexpr int ASM_T = 256;

// The q-row range and the p-row range of a tile are staged side by side; BULK: each leaves the SM as ONE
// bulk copy issued by thread 0 (gnx_tma.cuh); otherwise the block streams them out itself.
template <int CPT> constexpr int asm_stage_doubles() { return (5 * ASM_T * CPT + 3 * ASM_T + 4) + (3 * ASM_T * CPT + 4); }

// q-row of lo-frame position t of one edge -> its slots in the staged CSR range; returns the row's table flags.
// STIFF = false compiles the stiffness term k/h (NetworkStokes; two fp64 divisions per row) out.
struct AsmEdge {
  const double* he;          // cell lengths of the edge (u -> v order)
  const double* ac;          // per-cell mass coefficient of the edge (same order), or NULL: `a` for every cell
  double a, kk, sgn;         // mass coefficient, stiffness coefficient, +-sigma (orientation of the lo frame)
  int32_t ebase, blo1, bhi1; // staged offset of the edge's range, 1 if the lo / hi end carries a multiplier
  bool flip;                 // u > v: the lo frame runs against the edge
};
// TAB_GLOBAL = false: `tab` (and usually E.he) point into shared memory (persistent step kernel).
template <bool STIFF, bool TAB_GLOBAL = true>
__device__ __forceinline__ int32_t asm_qrow(double* stage, const int2* __restrict__ tab, const AsmEdge& E, int32_t t,
                                            int32_t m, double sg) {
  const bool hasl = t > 0, hasr = t < m;
  const double hl = hasl ? E.he[E.flip ? m - t : t - 1] : 0.0;       // lo-frame cell t-1
  const double hr = hasr ? E.he[E.flip ? m - 1 - t : t] : 0.0;       // lo-frame cell t
  const int2 tb = TAB_GLOBAL ? __ldg(tab + t) : tab[t];              // row offset + packed slot order (device.row_table)
  const int32_t f = tb.y;
  const int32_t off = E.ebase + tb.x + ((f & 1) ? E.blo1 : 0) + ((f & 2) ? E.bhi1 : 0);
  double kl = 0.0, kr = 0.0;
  if (STIFF) { if (hasl) kl = E.kk / hl; if (hasr) kr = E.kk / hr; }
  double diag = 0.0;
  const double al = (E.ac && hasl) ? E.ac[E.flip ? m - t : t - 1] : E.a;
  const double ar = (E.ac && hasr) ? E.ac[E.flip ? m - 1 - t : t] : E.a;
  if (hasl) {
    diag += al * hl * (2.0 / 6.0) + kl;
    stage[off + ((f >> 2) & 7)] = al * hl * (1.0 / 6.0) - kl;
    stage[off + ((f >> 11) & 7)] = -E.sgn;                           // p column of cell t-1
  }
  if (hasr) {
    diag += ar * hr * (2.0 / 6.0) + kr;
    stage[off + ((f >> 8) & 7)] = ar * hr * (1.0 / 6.0) - kr;
    stage[off + ((f >> 14) & 7)] = E.sgn;                            // p column of cell t
  }
  stage[off + ((f >> 5) & 7)] = diag;
  // lambda column at a bifurcation end (last slot of the row): +sigma if the edge points INTO the node
  if (t == 0 && E.blo1) stage[off + 3] = E.flip ? sg : -sg;
  if (t == m && E.bhi1) stage[off + 3] = E.flip ? -sg : sg;
  return f;
}

#endif  // GNX_ASM_TILE_CUH
