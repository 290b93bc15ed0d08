/* graphnics_b200 -- C ABI of the B200-native network-flow hot path.
 *
 * The reference (IngeborgGjerde/graphnics) has NO native/FFI boundary: its hot path is Python
 * calling DOLFIN / fenics_ii / PETSc.  This header is the boundary a maintainer would bind
 * instead (ctypes stub in INTEGRATION.md); every entry point names the reference call it
 * replaces (paths relative to the reference repo).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns all buffers (the Python host allocates them as torch CUDA tensors);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing syncs
 *     unless stated;
 *   - return value: 0 = ok, <0 = error (message via gnx_last_error());
 *   - int32 indices: #cells, #dofs and #non-zeros must be < 2^31;
 *   - no CPU fallback exists: without a CUDA device every compute entry point fails.
 *
 * Numbering (canonical): mesh vertices/cells follow DOLFIN's hierarchical
 * bisection; mixed dofs W = [q_0..q_{E-1} | p | lambda], q_i numbered like the vertices of
 * SubMesh i (first encounter), p like global cells, lambda like bifurcation_ixs.
 */
#ifndef GRAPHNICS_B200_H
#define GRAPHNICS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GNX_OK 0
#define GNX_ERR_ARG -1
#define GNX_ERR_CUDA -2
#define GNX_ERR_NOCONV -3

/* Marker tags, graphnics/fenics_graph.py:22-25 */
#define GNX_BIF_IN 1
#define GNX_BIF_OUT 2
#define GNX_BOUN_IN 3
#define GNX_BOUN_OUT 4

/* Network descriptor (host struct holding device pointers).  Filled by the host from the
 * outputs of gnx_refine_1d / gnx_vertex_tables; read-only for all later calls. */
typedef struct gnx_network {
  int32_t V, E, gdim, n;      /* nodes, edges, geometric dim (2|3), refinements (m = 2^n cells/edge) */
  int32_t B, n_bnd;           /* #bifurcation nodes (degree>=2), #boundary nodes (degree==1) */
  int64_t I;                  /* #(edge end, bifurcation node) incidences */
  const double*  pos;         /* [V*gdim]  node positions */
  const int32_t* edges;       /* [E*2]     (u,v) in G.edges() order */
  const double*  h;           /* [E*m]     cell lengths, edge-local order u -> v */
  const int32_t* bix;         /* [V]       bifurcation index of node, or -1 */
  const int32_t* bif_nodes;   /* [B]       node of bifurcation index (== G.bifurcation_ixs) */
  const int32_t* inc_ptr;     /* [V+1]     node -> incident edge ends (CSR) */
  const int32_t* inc_edge;    /* [2E]      (edge<<1)|end, end 0: node==u (out-edge), 1: node==v (in-edge); ascending */
  const int32_t* lam_ptr;     /* [B+1]     exclusive scan of bifurcation degrees */
  const int32_t* ebif_prefix; /* [E+1]     exclusive scan over edges of #ends that are bifurcations */
  const int32_t* loc;         /* [m+1]     SubMesh-local vertex id of position t counted from min(u,v) */
  const int32_t* tloc;        /* [m+1]     inverse of loc */
  const int32_t* rowtab;      /* [2*(m+1)] per position t: CSR row offset inside an edge and the packed slot order of
                                 its q-row (same for every edge; built once by the host, device.row_table); the
                                 edge-tile assembly kernel needs it, every other entry point ignores it (may be NULL
                                 for n > 10) */
  const int32_t* erec;        /* [E*4] or NULL: what the edge-tile kernels need to know about an edge in ONE 16-byte
                                 load instead of the chain edges -> bix: bix[u], bix[v], ebif_prefix[e],
                                 u | (u > v ? 1 << 31 : 0).  Built once by the host (device.py) from the tables above;
                                 kernels walk the tables themselves when it is NULL */
} gnx_network_t;

/* Edge-wise operator coefficients of the (generalised) mixed system
 *   [ a M + k K   -s B^T   s C^T ] [q]     M = P1 mass, K = P1 stiffness (per edge),
 *   [ s B          0        0    ] [p]     B = d/ds coupling, C = bifurcation incidence.
 *   [ s C          0        0    ] [lam]
 * steady MixedHydraulicNetwork: a = Res_e, k = 0, s = 1      (flow_models.py:233-236)
 * NetworkStokes:                k = mu/Ainv_e                (flow_models.py:399)
 * theta-scheme step:            a = Ainv + theta*dt*Res_e, s = theta*dt  (time_stepping.py:184) */
typedef struct gnx_mixed_coeffs {
  const double* a_edge;       /* [E] mass coefficient */
  const double* k_edge;       /* [E] stiffness coefficient, or NULL */
  double sigma;               /* coupling scale s */
  /* [E*m] mass coefficient per CELL in edge-local order (cell k of edge e at e*m + k, u -> v, like gnx_network_t.h), or
   * NULL: edge attribute "Res" given as a function of position / time (flow_models.py:228-233,
   * demo/Vasomotion in arterial trees.ipynb:222-289), sampled at the cell midpoints -- the element mass matrix is
   * a_cell h/6 [[2,1],[1,2]], exact for cell-wise constant data.  Overrides a_edge where given; the chain kernels of
   * the solve handle it (2^0..2^10 cells per edge), the persistent / Constant-data fast paths step aside. */
  const double* a_cell;
} gnx_mixed_coeffs_t;

const char* gnx_last_error(void);
int gnx_version(void);
/* 1 if a CUDA device is usable, else 0 (never throws) */
int gnx_device_ok(void);

/* ---------------- graph -> mesh (graphnics/fenics_graph.py) ---------------- */

/* FenicsGraph.get_mesh :58-99 (MeshEditor + n x refine/adapt).  One thread per cell walks the
 * bisection tree; writes hierarchical vertex coordinates (recursive midpoints, bit-exact),
 * sorted cell->vertex pairs, the cell->edge marker and the cell lengths.
 *   coords [ (V+E(m-1))*gdim ], cells [E*m*2], mf [E*m] (may be NULL), h [E*m] */
int gnx_refine_1d(const double* pos, const int32_t* edges, int32_t V, int32_t E, int32_t gdim,
                  int32_t n, double* coords, int32_t* cells, int32_t* mf, double* h, void* stream);

/* FenicsGraph.assign_tangents :230-253 (+ compute_edge_lengths :187-198).
 *   tangents [E*gdim] = (pos[v]-pos[u]) * (1/||.||), elen [E] (may be NULL) */
int gnx_edge_tangents(const double* pos, const int32_t* edges, int32_t E, int32_t gdim,
                      double* tangents, double* elen, void* stream);

/* record_bifurcation_and_boundary_nodes :161-184 and the tag logic of make_submeshes /
 * mark_edge_vfs :132-156,280-311.
 *   deg[V], bix[V], bif_nodes[V] (first B valid), bnd_nodes[V] (first n_bnd valid),
 *   inc_ptr[V+1], inc_edge[2E], lam_ptr[V+1] (first B+1 valid), ebif_prefix[E+1],
 *   tag_u[E], tag_v[E] (vf tag at the u / v end of every edge),
 *   counts[4] = {B, n_bnd, I, n_lonely}, work[>= gnx_tables_work_ints(V,E)] int32 scratch;
 *   cls_deg [V] or NULL: when `edges` is one rank's PART of a larger network, the node degrees in the
 *   whole network -- nodes are classified (bifurcation / boundary) by cls_deg while the incidence
 *   lists, lam_ptr, ebif_prefix and I count this part's edges only (subdomain tables). */
int64_t gnx_tables_work_ints(int32_t V, int32_t E);
int gnx_vertex_tables(const int32_t* edges, int32_t V, int32_t E, int32_t* deg, int32_t* bix,
                      int32_t* bif_nodes, int32_t* bnd_nodes, int32_t* inc_ptr, int32_t* inc_edge,
                      int32_t* lam_ptr, int32_t* ebif_prefix, int32_t* tag_u, int32_t* tag_v,
                      int32_t* counts, int32_t* work, const int32_t* cls_deg, void* stream);

/* ---------------- sparse assembly (graphnics/flowmodels/flow_models.py) ---------------- */

/* sizes of the monolithic mixed system (host arithmetic) */
int64_t gnx_mixed_ndofs(const gnx_network_t* net);
int64_t gnx_mixed_nnz(const gnx_network_t* net);

/* CSR pattern of MixedHydraulicNetwork.a_form :200-280 after ii_assemble/ii_convert :343-344
 * (canonical sparsity: structural entries + the explicit-zero p / lambda diagonals of
 * init_a_form).  Built once; one thread per row, closed form, sorted columns.
 *   rowptr [N+1], colidx [nnz] */
int gnx_build_pattern_mixed(const gnx_network_t* net, int32_t* rowptr, int32_t* colidx, void* stream);

/* Values of the same matrix (row gather: every CSR slot is the fixed-order sum of its <=2
 * element contributions; no atomics).  vals [nnz] */
int gnx_assemble_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                       const int32_t* rowptr, double* vals, void* stream);

/* TimeDepMixedHydraulicNetwork.mass_matrix_q (time_stepping.py:75-101) on the SAME pattern:
 * vals = coeff * blockdiag(M_e), zero elsewhere. */
int gnx_assemble_mass_q_mixed(const gnx_network_t* net, double coeff, const int32_t* rowptr,
                              double* vals, void* stream);

/* y = coeff * blockdiag(M_e) x on the q rows, 0 elsewhere: the action of mass_matrix_q /
 * mass_vector_q (time_stepping.py:103-123, `DDn = DDn1*qp0.block_vec()` :216) without CSR values. */
int gnx_mass_apply_q_mixed(const gnx_network_t* net, double coeff, const double* x, double* y,
                           void* stream);

/* Coefficient expressions (f, g, p_bc: dolfin Constant / Expression("c-string") / UFL in the
 * reference, tests/test_flow_models.py:94-119) are compiled on the host to a small RPN program
 * and interpreted on the device, one point per thread. */
#define GNX_PROG_MAX 96
#define GNX_PROG_PARAMS 8
typedef struct gnx_program {
  int32_t len;
  int32_t reserved;
  int16_t op[GNX_PROG_MAX];     /* opcode, see GNX_OP_* */
  int16_t arg[GNX_PROG_MAX];    /* integer operand (coordinate / parameter index, integer power) */
  double val[GNX_PROG_MAX];     /* constant operand */
  double params[GNX_PROG_PARAMS]; /* run-time scalars (params[0] = t by convention) */
} gnx_program_t;
enum {
  GNX_OP_CONST = 0, GNX_OP_X = 1, GNX_OP_PARAM = 2, GNX_OP_TAU = 3, GNX_OP_EDGE = 4,
  GNX_OP_ADD = 10, GNX_OP_SUB, GNX_OP_MUL, GNX_OP_DIV, GNX_OP_POW, GNX_OP_NEG, GNX_OP_MIN, GNX_OP_MAX,
  GNX_OP_SIN = 20, GNX_OP_COS, GNX_OP_TAN, GNX_OP_EXP, GNX_OP_LOG, GNX_OP_SQRT, GNX_OP_ABS, GNX_OP_TANH,
  GNX_OP_SINH, GNX_OP_COSH, GNX_OP_ATAN, GNX_OP_ASIN, GNX_OP_ACOS, GNX_OP_FLOOR, GNX_OP_FMOD,
  GNX_OP_POWI, GNX_OP_ATAN2, GNX_OP_SIGN, GNX_OP_LT, GNX_OP_GT, GNX_OP_LE, GNX_OP_GE, GNX_OP_SELECT,
  GNX_OP_CEIL
};
/* out[i] = prog(x[i,:]) for i < npts.  If pts_per_edge > 0, point i lies on edge i / pts_per_edge
 * and GNX_OP_TAU reads tangents[edge*gdim + arg], GNX_OP_EDGE reads edge_attr[edge]. */
int gnx_eval_expr(const gnx_program_t* prog_host, const double* x, int32_t gdim, int64_t npts,
                  int32_t pts_per_edge, const double* tangents, const double* edge_attr, double* out,
                  void* stream);

/* MixedHydraulicNetwork.L_form :299-332.
 *   f_nodal, g_nodal [E*(m+1)] = the per-edge CG1 projections (:315-317) in q-dof layout, or NULL
 *                                 to use the constants f_const / g_const (dolfin Constant data);
 *   pbc_out  [E] = projected p_bc at the v-end of every edge (used where v is a boundary node:
 *                  the `p_bcs[i]*v*ds(BOUN_OUT)` term :324), NULL = 0;
 *   pbc_node [V] = p_bc evaluated at graph nodes (the un-projected BOUN_IN term :324), NULL = 0.
 *   b [N] */
int gnx_assemble_rhs_mixed(const gnx_network_t* net, const double* f_nodal, const double* g_nodal,
                           double f_const, double g_const, const double* pbc_out,
                           const double* pbc_node, double* b, void* stream);

/* project(c, CG1(submesh_i)) for all edges (:315-317): consistent-mass L2 projection by parallel
 * cyclic reduction (diagonally dominant tridiagonal systems: couplings fall below 1 ulp after 7
 * reduction steps, so the step count is independent of m).
 *   c_nodal [E*(m+1)] coefficient at the q-dof vertices, c_mid [E*m] at cell midpoints in p-dof
 *   order (NULL => degree<=1 data: midpoint = average).  out [E*(m+1)] q-dof layout.
 *   work [>= 8*E*(m+1)] doubles */
int gnx_project_p1_edges(const gnx_network_t* net, const double* c_nodal, const double* c_mid,
                         double* out, double* work, void* stream);

/* Value at the v-end of every edge of the same projection of `prog` (degree = FEniCS degree of
 * the Expression), computed from a window of the last `window` cells only (the inverse mass
 * matrix decays like 0.268^k).  out [E]. */
int gnx_project_end_values(const gnx_network_t* net, const double* coords,
                           const gnx_program_t* prog_host, int32_t degree, int32_t window,
                           double* out, void* stream);

/* coordinates of the q-dof vertices [E*(m+1)*gdim] and of cell midpoints [E*m*gdim] (p-dof order),
 * gathered from the mesh coordinates -- inputs of coefficient evaluation. */
int gnx_dof_coordinates_mixed(const gnx_network_t* net, const double* coords, const int32_t* cells,
                              double* xq, double* xmid, void* stream);

/* ---------------- saddle-point solve (replaces ii_convert + LUSolver("mumps")) ---------------- */

/* Exact per-edge elimination -> graph-level SPD system on the bifurcation multipliers.
 * Phase 1: two segmented scans over the cells of every edge.
 *   b [N] rhs (dof order);  cond[E], rsrc[E], ftot[E] edge scalars out */
int gnx_edge_eliminate_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co,
                             const double* b, double* cond, double* rsrc, double* ftot, void* stream);

/* Edge-partitioned network (one process per GPU): where the edge scalars of a solve go and how the ranks
 * tell each other that they have arrived.  Every rank owns one GATHER BUFFER per solve in flight (callers
 * alternate two): world blocks [cond | rsrc | ftot] of `chunk` doubles each (block of rank q at
 * q*3*chunk -- the layout gnx_schur_solve reads with chunk, rank_stride = 3*chunk), followed by `world`
 * 64-bit flag words.  peers_host[r] is the device address, valid on THIS GPU, of rank r's buffer (NVLink
 * peer mapping, e.g. torch symmetric memory).  The elimination kernel stores its scalars straight into all
 * buffers.  With epoch > 0 the Schur kernel (gnx_schur_solve given the same descriptor) starts by writing
 * `epoch` into flag word `rank` of every buffer (st.release.sys; its predecessor on the stream, the
 * elimination kernel, is complete, so this rank's scalars have landed everywhere) and then waits until all
 * `world` flag words of its own buffer are >= epoch (ld.acquire.sys) -- no collective call and no barrier
 * kernel is left.  epoch must grow from solve to solve on a buffer; epoch = 0: no flags, the caller
 * separates the two kernels with its own cross-GPU barrier.  world <= 8. */
typedef struct gnx_exchange {
  int32_t world, rank, chunk, mode;   /* mode 0: gathered blocks as described above; mode 1: see below */
  double* const* peers_host;   /* [world], host array of device addresses */
  int64_t epoch;
  /* mode 1 -- SUBTREE-ALIGNED partition (graphnics_b200/partition.py; north_star: "edge-balanced subtrees"): a rank owns
   * an arbitrary set of edges, egl[e] = global id of its local edge e with bit 31 set when an end of the edge is a
   * REPLICATED bifurcation.  Every rank's buffer is indexed by GLOBAL ids:
   *   [cond | rsrc | ftot (E_glob doubles each) | d | r (B_glob each) | world 64-bit flag words].
   * A rank eliminates, builds the Schur rows of and rakes the bifurcations whose whole subtree it owns WITHOUT any
   * communication; only the scalars of edges at replicated bifurcations (bit 31) and the final (d, r) pair of every
   * local subtree root under a replicated parent (gnx_schur_plan_t.xlist) are stored into all buffers over NVLink.  The
   * ONE exchange of a solve then happens inside the Schur kernel at rake level plan.l_sync; the replicated
   * bifurcations (the top of the tree: 2^k - 1 for a binary tree on 2^k ranks) are processed redundantly, bit-identically.
   * phase: 0 = the whole solve in one kernel (epoch > 0): the flag words tell that a rank's EDGE SCALARS have landed
   * (released as soon as its elimination is complete), the (d, r) words of the subtree roots VALIDATE THEMSELVES -- the
   * caller creates both buffers with the whole (d, r) region filled with the 64-bit pattern 0x7FF85EA71E550001 (a NaN
   * no computation produces), the owner of a root stores its pair with plain peer stores, a reader polls the word until
   * it differs from the pattern and writes the pattern back (GNX_XCHG_FLAGS=1: data first, flag after, no pattern
   * needed); 1 = up to the exchange (stores to the peers done, no flags); 2 = from the exchange on -- 1 and 2 let a
   * caller put its own collective / barrier in between (NCCL fallback, ranks emulated in one process). */
  const int32_t* egl;
  int32_t E_glob, B_glob, phase, reserved;
} gnx_exchange_t;

/* L_form with Constant f / g (what the steady models use: boundary data + constants) and the elimination
 * in ONE kernel: b [N] is an OUTPUT, bit-identical to gnx_assemble_rhs_mixed(net, NULL, NULL, f_const,
 * g_const, pbc_out, pbc_node, b).  Saves a launch and the 8 B/dof read of b.  n <= 10.
 * ex != NULL selects the edge-partitioned output (mode 0: cond / rsrc / ftot are ignored; mode 1: cond / rsrc [E] receive
 * the local copies the back-substitution reads). */
int gnx_rhs_eliminate_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const,
                            double g_const, const double* pbc_out, const double* pbc_node, double* b,
                            double* cond, double* rsrc, double* ftot, const gnx_exchange_t* ex, void* stream);

/* Edge-partitioned network: gnx_edge_eliminate_mixed with the collective fused in (see gnx_exchange; cond / rsrc:
 * mode 1 only, may be NULL in mode 0). */
int gnx_edge_eliminate_mixed_p2p(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const double* b,
                                 const gnx_exchange_t* ex, double* cond, double* rsrc, void* stream);

/* Graph-level Schur system S P = r on the B bifurcation unknowns (node gather, no atomics),
 * as explicit arrays -- inspection / tests; gnx_schur_solve builds the same quantities itself.
 *   b_lam [B] (lambda rows of b, may be NULL), sdiag [B], soff [2E] (aligned with inc_edge),
 *   scol [2E] (bifurcation index of the other end, -1 = boundary), srhs [B] */
int gnx_schur_build(const gnx_network_t* net, double sigma, const double* cond, const double* rsrc,
                    const double* ftot, const double* b_lam, double* sdiag, double* soff,
                    int32_t* scol, double* srhs, void* stream);

/* Symbolic plan of the Schur solve (index data only, built once per network on the host --
 * graphnics_b200/schur_plan.py): tree-like parts of the bifurcation graph are eliminated exactly
 * level by level ("rake": a node with one remaining neighbour folds into it without fill), the
 * cyclic core that is left is solved by Jacobi-PCG.  All pointers are device pointers, indices are
 * bifurcation indices unless stated. */
#define GNX_MAX_LEVELS 64
#define GNX_TOP_MAX 2048
#define GNX_AGG_MAX 128
#define GNX_BLK_MAX 1024
#define GNX_SUB_MAX 1024
typedef struct gnx_schur_plan {
  int32_t B, nl, l_split, nc;          /* unknowns, rake levels, first level of the TOP part, core size */
  int32_t nnzc, nt;                    /* off-diagonal entries of the core matrix; #nodes in levels >= l_split
                                          (<= GNX_TOP_MAX: swept by one CTA out of shared memory) */
  int32_t lvl_ptr[GNX_MAX_LEVELS + 1]; /* level l = lvl_nodes[lvl_ptr[l] .. lvl_ptr[l+1]) */
  const int32_t* lvl_nodes;            /* [lvl_ptr[nl]] */
  const int32_t* parent;               /* [B] node a raked node folds into, -1 = none */
  const int32_t* pedge;                /* [B] graph edge linking a raked node to its parent, or -1 */
  const int32_t* ch_ptr;               /* [B+1] raked children of every node (CSR) */
  const int32_t* ch_idx;
  const int32_t* core_nodes;           /* [nc] */
  const int32_t* crp;                  /* [nc+1] core matrix rows (CSR, core-local columns) */
  const int32_t* ccol;                 /* [nnzc] */
  const int32_t* cedge;                /* [nnzc] graph edge whose -cond is the entry */
  const int32_t* top_nodes;            /* [nt] nodes of the top part */
  const int32_t* top_of;               /* [B] position in top_nodes, or -1 */
  const int32_t* level_of;             /* [B] rake level of a node (nl for core nodes) */
  /* ELL copy of the core matrix for the cluster-resident PCG (ell_w = 0: not available -- core wider
   * than 8 entries per row or larger than 16 x 5120 rows): entry k of core row i at [k*nc + i];
   * ell_col packs (owning CTA << 16 | offset) for ell_R rows per CTA, -1 = padding */
  int32_t ell_w, ell_R;
  int32_t top_S;                       /* > 0: a pure tree (nc = 0, l_split = 0) of <= 16*top_S nodes is swept by ONE
                                          16-CTA cluster, top_S slots per CTA (power of two, 512..2048) */
  int32_t na;                          /* > 0: two-level preconditioner, the core rows are grouped into na aggregates */
  const int32_t* ell_col;              /* [ell_w*nc] */
  const int32_t* ell_edge;             /* [ell_w*nc] graph edge whose -cond is the entry, -1 = padding */
  /* Multilevel-additive PCG on a large core (na > 0, 2 <= na <= GNX_AGG_MAX; cooperative grid of na CTAs, one
   * per coarse aggregate): preconditioner  M^-1 = D^-1 + W1 diag(W1^T S W1)^-1 W1^T + W2 (W2^T S W2)^-1 W2^T
   * with W1 / W2 the piecewise-constant indicator vectors of the na*ns fine / na coarse aggregates
   * (na*ns <= GNX_SUB_MAX, ns a power of two <= 32).  Core rows are numbered fine aggregate by fine aggregate. */
  const int32_t* agg_ptr;              /* [na+1] core rows of aggregate J = agg_ptr[J] .. agg_ptr[J+1]) */
  const int32_t* col_agg;              /* [nnzc] FINE aggregate of the column of every core entry (coarse = fine / ns) */
  const int32_t* cut_ptr;              /* [na*na+1] cut_k[cut_ptr[I*na+J] ..): core entries coupling aggregate I to J != I */
  const int32_t* cut_k;
  /* Blocked rake of a large forest (nb > 0, nc == 0): the raked nodes outside the top part form nb BLOCKS,
   * maximal complete subtrees of <= GNX_BLK_MAX nodes, each swept by one CTA out of shared memory; the top
   * part (top_nodes, <= GNX_TOP_MAX nodes) is then the set of nodes with larger subtrees and l_split its
   * lowest rake level.  Two grid barriers replace one per wide level. */
  int32_t nb;
  int32_t ns;                          /* multilevel PCG: every coarse aggregate is the union of ns fine ones (1: coarse level only) */
  const int32_t* blk_ptr;              /* [nb+1] block b = blk_nodes[blk_ptr[b] .. blk_ptr[b+1]) */
  const int32_t* blk_nodes;
  const int32_t* blk_pos;              /* [B] position of a node inside its block, -1 for top-part nodes */
  const int32_t* blk_lv;               /* [2*nb] lowest and highest rake level of a block */
  const int32_t* sub_ptr;              /* [na*ns+1] core rows of the fine aggregates (ns > 1) */
  const int32_t* row_sub;              /* [nc] fine aggregate of a core row (ns > 1) */
  /* One rank's view of the plan of a subtree-partitioned network (gnx_exchange mode 1; schur_plan.RankPlan), ncls != NULL:
   * lvl_nodes / top_nodes / blocks hold the nodes THIS rank sweeps -- its local ones on levels < l_sync, the
   * replicated raked ones on levels >= l_sync (l_split <= l_sync when nc == 0); parent / pedge / children / core arrays
   * stay global.  ncls[B]: bit 0 replicated, bit 1 another rank's local node, bit 2 local root whose final (d, r) is
   * exported; xlist[nxl]: those roots. */
  int32_t l_sync, nxl;
  const int32_t* ncls;
  const int32_t* xlist;
  /* [nt][16] packed records of the top part (schur_plan.top_records; NULL: the sweep walks the generic arrays): node,
   * level, parent's top position, parent edge, two shared-memory children, their count (bit 30: row does not fit the
   * record), two global-memory children and their edges, graph node, four incident edges (edge << 1 | end) */
  const int32_t* top_rec;
} gnx_schur_plan_t;

/* The whole Schur solve as ONE persistent kernel (a single CTA for small systems, a cooperative
 * grid otherwise): Schur diagonal/rhs gather, rake up-sweep, Jacobi-PCG on the core (two fused
 * phases per iteration, deterministic reductions, convergence decided on the device), rake
 * down-sweep.  Replaces ii_convert + LUSolver("mumps") (flow_models.py:344-348) together with
 * gnx_edge_eliminate_mixed / gnx_edge_backsub_mixed.
 *   b_lam [B] (mixed model: lambda rows of b, enters as -b_lam/sigma) or b_node [V] (primal model:
 *   p rows of b at the graph nodes, enters as +b_node/sigma); at most one of them non-NULL;
 *   P [B] out: sigma*lambda (in: initial guess of the core unknowns when warm != 0);
 *   lam_out [B] (may be NULL) = P / sigma, i.e. the lambda block of the solution vector;
 *   work [>= gnx_schur_work_doubles(plan)] doubles;
 *   chunk > 0 (edge-partitioned network): cond / rsrc / ftot are read through the all-gather layout
 *   arr[(e / chunk) * rank_stride + e % chunk], so one collective of the per-rank [cond|rsrc|ftot]
 *   blocks feeds the kernel directly; chunk = 0: plain arr[e].
 *   ex != NULL with ex->epoch > 0: wait for the flag words of all ranks before reading the edge scalars
 *   (see gnx_exchange); NULL otherwise.
 * Fully asynchronous (and CUDA-graph capturable) when resid_host and info_host are NULL; otherwise
 * synchronises and returns resid_host[0] = ||r||/||rhs|| of the core CG, info_host[3] =
 * {CG iterations, converged, CTAs}; GNX_ERR_NOCONV if max_iter was hit. */
int64_t gnx_schur_work_doubles(const gnx_schur_plan_t* plan);
int gnx_schur_solve(const gnx_network_t* net, const gnx_schur_plan_t* plan, double sigma,
                    const double* cond, const double* rsrc, const double* ftot, const double* b_lam,
                    const double* b_node, double* P, double* lam_out, double* work, double rtol, int32_t max_iter,
                    int32_t warm, int32_t chunk, int32_t rank_stride, const gnx_exchange_t* ex, double* resid_host, int32_t* info_host,
                    void* stream);

/* How many solves on `work` ended without convergence since the last call of this function (asynchronous solves --
 * resid_host = info_host = NULL, e.g. the steps of a time loop -- cannot say so themselves); resets the count.
 * Synchronises the stream.  Negative: CUDA error. */
int gnx_schur_unconverged(const gnx_schur_plan_t* plan, double* work, void* stream);

/* Phase 2: back-substitution, the q and p blocks of x [N] (dof order) from the multipliers P [B]
 * (= sigma*lambda; the lambda block itself is written by gnx_schur_solve). */
int gnx_edge_backsub_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const double* b,
                           const double* cond, const double* rsrc, const double* P, double* x,
                           void* stream);

/* gnx_edge_backsub_mixed for a right-hand side of Constant data (the b that gnx_rhs_eliminate_mixed produced from the
 * same f_const, g_const, pbc_out, pbc_node): b is NOT read -- its entries are recomputed (16 B/dof of traffic less),
 * f = g = 0 takes a fast path; x is bit-identical.  2^6 .. 2^10 cells per edge, else (or when b != NULL and the tile
 * path does not apply) it forwards to gnx_edge_backsub_mixed. */
int gnx_edge_backsub_mixed_const(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const, double g_const,
                                 const double* pbc_out, const double* pbc_node, const double* b, const double* cond,
                                 const double* rsrc, const double* P, double* x, void* stream);

/* End flux and end pressure of every edge from the multipliers (one thread per edge; the last step of
 * LUSolver.solve, flow_models.py:347-350, seen from an edge):  Pu [E] = P at the tail node (0 at a boundary node),
 * q0 [E] = cond (P_u - P_v + rsrc).  q0 may alias ftot and Pu may alias rsrc of the elimination. */
int gnx_edge_flux_mixed(const gnx_network_t* net, const double* cond, const double* rsrc, const double* P,
                        double* q0, double* Pu, void* stream);

/* gnx_edge_backsub_mixed_const given q0 / Pu (gnx_edge_flux_mixed) instead of cond / rsrc / P: the blocks of a large
 * network then reach their edge data in ONE trip to device memory; x is bit-identical.  Edge-tile path only
 * (2^6 .. 2^10 cells per edge, per-edge coefficients), GNX_ERR_ARG otherwise. */
int gnx_edge_backsub_mixed_flux(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const, double g_const,
                                const double* pbc_out, const double* pbc_node, const double* q0, const double* Pu,
                                double* x, void* stream);

/* One MixedHydraulicNetwork.solve() (flow_models.py:337-350: ii_assemble(a), ii_assemble(L) with Constant f, g,
 * ii_convert, LUSolver.solve) behind ONE call -- the same kernels as gnx_assemble_mixed +
 * gnx_rhs_eliminate_mixed + gnx_schur_solve + gnx_edge_backsub_mixed with bit-identical results; it exists
 * because four host round trips cost more than the GPU needs for a 2 M-dof step.
 *   vals [nnz] out (NULL: no assembly);  b [N] out, x [N] out;  cond / rsrc / ftot [E], P [B], work: scratch
 *   as for the separate calls;  schur_net: the network the Schur plan was built for (== net on one GPU);
 *   side_stream != NULL: the solve chain runs on it and overlaps the assembly on `stream`; ev_fork / ev_join
 *   are two CUDA events of the caller used to fork and join the streams (everything is ordered on `stream`
 *   when the call returns);  resid_host / info_host as in gnx_schur_solve (non-NULL synchronises);
 *   ex != NULL (edge-partitioned network, ex->epoch > 0): cond / rsrc / ftot are ignored, the scalars travel
 *   through the gather buffers and the Schur kernel reads ex->peers_host[ex->rank].  n <= 10.
 *   sync != NULL (2 device words, zeroed ONCE by the caller) and epoch >= 1 (strictly increasing from call to
 *   call on the same `sync`): networks whose cell lengths fit the shared memory of the GPU (<= 8 tiles of 1024
 *   cells per SM, Schur system of <= 2048 unknowns, 2 <= n <= 10) run right-hand side + elimination -> Schur
 *   solve -> back-substitution as ONE persistent cooperative kernel (step_resident_kernel, gnx_solve.cu) that keeps
 *   the lengths resident between the phases; b and x are bit-identical to the kernel chain's. */
int gnx_step_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const int32_t* rowptr, double* vals,
                   double f_const, double g_const, const double* pbc_out, const double* pbc_node, double* b,
                   double* x, const gnx_network_t* schur_net, const gnx_schur_plan_t* plan, double* cond,
                   double* rsrc, double* ftot, double* P, double* work, double rtol, int32_t max_iter,
                   int32_t warm, const gnx_exchange_t* ex, double* resid_host, int32_t* info_host, void* ev_fork, void* ev_join,
                   void* side_stream, int64_t* sync, int64_t epoch, void* stream);

/* 1 if the last gnx_step_mixed call of this process ran as the persistent kernel, 0 if as the kernel chain (reporting) */
int gnx_step_is_resident(void);

/* ---------------- primal model: HydraulicNetwork (flow_models.py:53-126) ---------------- */

/* Operator  [ diag(a_c h_c)  s G ; -s G^T  0 ]  on  W = [DG0(mesh) | CG1(mesh)]  (q dof = global
 * cell, p dof = NC + global vertex).  steady: a = Res (:87), s = 1; theta-scheme step:
 * a = Ainv + theta*dt*Res, s = theta*dt (time_stepping.py:184).  a_cell (may be NULL -> a_const)
 * is per cell in EDGE-LOCAL order (cell k of edge e at e*m + k, u -> v), like gnx_network_t.h. */
typedef struct gnx_primal_coeffs {
  double a_const;
  const double* a_cell;
  double sigma;
} gnx_primal_coeffs_t;

int64_t gnx_primal_ndofs(const gnx_network_t* net);
int64_t gnx_primal_nnz(const gnx_network_t* net);
/* a_form :79-91 after ii_assemble/ii_convert: closed-form CSR rows, sorted columns, explicit p diagonal */
int gnx_build_pattern_primal(const gnx_network_t* net, int32_t* rowptr, int32_t* colidx, void* stream);
int gnx_assemble_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, double* vals, void* stream);
/* L_form :94-101.  f_cell / g_cell [NC*3]: the coefficient at (first stored vertex, midpoint, second
 * stored vertex) of every cell, or NULL -> the constants; *_lin != 0: degree <= 1 data (midpoint :=
 * average of the ends, what FEniCS integrates for a degree-1 Expression). */
int gnx_assemble_rhs_primal(const gnx_network_t* net, const int32_t* cells, const double* f_cell,
                            int32_t f_lin, double f_const, const double* g_cell, int32_t g_lin,
                            double g_const, double* b, void* stream);
/* y = diag(a_c h_c) x on the q rows, 0 on the p rows: mass_matrix_q / mass_vector_q (time_stepping.py:28-64) */
int gnx_mass_apply_q_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const double* x, double* y,
                            void* stream);
/* xii.apply_bc with DirichletBC(W[1], p_bc, "on_boundary") (:103-105,117): symmetric elimination of the
 * degree-1 graph nodes.  bc_node [V] = p_bc at the graph nodes; vals and/or b may be NULL. */
int gnx_apply_bc_primal(const gnx_network_t* net, double sigma, const double* bc_node, double* vals, double* b,
                        void* stream);
/* per-edge elimination / back-substitution of the primal system; b is the right-hand side AFTER
 * gnx_apply_bc_primal.  Between the two: gnx_schur_solve(..., b_lam = NULL, b_node = b + NC, ...). */
int gnx_edge_eliminate_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const double* b,
                              double* cond, double* rsrc, double* ftot, void* stream);
int gnx_edge_backsub_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const double* b,
                            const double* cond, const double* rsrc, const double* P, double* x, void* stream);

/* ---------------- Krylov building blocks ---------------- */

/* y = A x, fp64 CSR with int32 indices (row-block streaming through shared memory). */
int gnx_spmv_csr_f64(int32_t nrows, const int32_t* rowptr, const int32_t* colidx,
                     const double* vals, const double* x, double* y, void* stream);
/* r = b - A x and out2[0] = ||r||^2, out2[1] = ||b||^2 (deterministic two-stage reduction).
 * work [>= gnx_reduce_work_doubles(nrows)] */
int64_t gnx_reduce_work_doubles(int64_t n);
int gnx_residual_csr_f64(int32_t nrows, const int32_t* rowptr, const int32_t* colidx,
                         const double* vals, const double* x, const double* b, double* r,
                         double* out2, double* work, void* stream);
/* y = alpha*x + beta*y ; out[0] = <x,y> (if out != NULL, computed on the UPDATED y) */
int gnx_axpby_dot(int64_t n, double alpha, const double* x, double beta, double* y, double* out,
                  double* work, void* stream);

/* out = sum_{k<nterms} coefs_host[k] * vecs_host[k]  (nterms <= 4; out may alias vecs_host[0]).
 * The theta-scheme right-hand side b = DDn + cn1 dt Ln1 - dt cn An x0 + dt cn Ln
 * (graphnics/flowmodels/time_stepping.py:185) in one pass. */
int gnx_lincomb(int64_t n, int32_t nterms, const double* coefs_host, const double* const* vecs_host,
                double* out, void* stream);

/* One theta-scheme time step of time_stepping_stokes (graphnics/flowmodels/time_stepping.py:179-216) behind ONE call:
 *   Ln1 = L^{n+1};  b = DDn + w1 Ln1 - w0 A x0 + w0 Ln  (w1 = theta dt, w0 = (1-theta) dt; w0 = 0: implicit Euler -- vals_n,
 *   x0, Ln, ax, rowptr, colidx may then be NULL);  x1 = (D + w1 A)^-1 b  (co: a = Ainv + w1 Res, sigma = w1);  DDn = D x1.
 * Same kernels and results as gnx_assemble_rhs_* + gnx_spmv_csr_f64 + gnx_lincomb + the solve + gnx_mass_apply_q_*; it
 * exists because seven host round trips per step cost more than the GPU needs on a tree.  Single GPU.
 *   mixed: f_nodal / g_nodal / f_const / g_const / pbc_out / pbc_node as in gnx_assemble_rhs_mixed; DDn [N] in/out;
 *   Ln1, b, ax, x1 [N] out / scratch; cond, rsrc, ftot, P, work as in the separate solve calls.
 *   primal: f_cell / g_cell as in gnx_assemble_rhs_primal, bc_node [V] Dirichlet data (NULL: none), co_mass the
 *   coefficient of the q mass term. */
int gnx_timestep_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const gnx_schur_plan_t* plan,
                       const double* f_nodal, const double* g_nodal, double f_const, double g_const,
                       const double* pbc_out, const double* pbc_node, double w1, double w0, const int32_t* rowptr,
                       const int32_t* colidx, const double* vals_n, const double* x0, const double* Ln, double mass_coeff,
                       double* DDn, double* Ln1, double* b, double* ax, double* x1, double* cond, double* rsrc,
                       double* ftot, double* P, double* work, double rtol, int32_t max_iter, int32_t warm,
                       double* resid_host, int32_t* info_host, void* stream);
int gnx_timestep_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const gnx_schur_plan_t* plan,
                        const int32_t* cells, const double* f_cell, int32_t f_lin, double f_const, const double* g_cell,
                        int32_t g_lin, double g_const, const double* bc_node, double w1, double w0, const int32_t* rowptr,
                        const int32_t* colidx, const double* vals_n, const double* x0, const double* Ln,
                        const gnx_primal_coeffs_t* co_mass, double* DDn, double* Ln1, double* b, double* ax, double* x1,
                        double* cond, double* rsrc, double* ftot, double* P, double* work, double rtol, int32_t max_iter,
                        int32_t warm, double* resid_host, int32_t* info_host, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GRAPHNICS_B200_H */
