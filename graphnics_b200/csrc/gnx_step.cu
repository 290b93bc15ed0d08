// One MixedHydraulicNetwork.solve() step (flow_models.py:337-350) behind ONE call: CSR values on the
// caller's stream, right-hand side + elimination -> Schur kernel -> back-substitution on a side stream,
// forked and joined with the two events the caller provides.  The kernels are the ones of the separate
// entry points (bit-identical results); what this saves is host time -- four Python/ctypes round trips,
// stream-context switches and event calls cost more than the 55 us the GPU needs for the C2 step.
#include "gnx_common.cuh"

// gnx_solve.cu: the persistent (resident-tile) step kernel
int gnx_step_resident_applies(const gnx_network_t* net, const gnx_schur_plan_t* plan, const double* b, const double* x,
                              const unsigned long long* sync);
int gnx_step_resident_launch(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, double f_const, double g_const,
                             const double* pbc_out, const double* pbc_node, double* b, double* x,
                             const gnx_network_t* schur_net, const gnx_schur_plan_t* plan, double* cond, double* rsrc,
                             double* ftot, double* P, double* work, double rtol, int32_t max_iter, int32_t warm,
                             const gnx_exchange_t* ex, unsigned long long* sync, unsigned long long epoch,
                             double* vals, double* resid_host, int32_t* info_host, cudaStream_t st);

static int g_last_step_resident = 0;
extern "C" int gnx_step_is_resident(void) { return g_last_step_resident; }

extern "C" int gnx_step_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const int32_t* rowptr,
                              double* vals, double f_const, double g_const, const double* pbc_out,
                              const double* pbc_node, double* b, double* x, const gnx_network_t* schur_net,
                              const gnx_schur_plan_t* plan, double* cond, double* rsrc, double* ftot, double* P,
                              double* work, double rtol, int32_t max_iter, int32_t warm, const gnx_exchange_t* ex,
                              double* resid_host, int32_t* info_host, void* ev_fork, void* ev_join, void* side_stream,
                              int64_t* sync, int64_t epoch, void* stream) {
  GNX_CHECK_ARG(net && co && b && x && (ex || (cond && rsrc && ftot)));
  GNX_CHECK_ARG(!ex || (ex->epoch > 0 && ex->peers_host && ex->rank >= 0 && ex->rank < ex->world));
  GNX_CHECK_ARG(!ex || ex->mode == 0 || (cond && rsrc));
  GNX_CHECK_ARG(net->B == 0 || (schur_net && plan && P && work));
  cudaStream_t main_st = (cudaStream_t)stream;
  if (sync && epoch > 0 && net->B > 0 && !co->a_cell && (!vals || (net->rowtab && ((uintptr_t)vals & 15) == 0)) &&
      gnx_step_resident_applies(net, plan, b, x, reinterpret_cast<const unsigned long long*>(sync))) {
    // small enough for the cell lengths to stay in shared memory: the whole step -- assembly included -- is ONE
    // persistent kernel on the caller's stream (no second stream, no events)
    g_last_step_resident = 1;
    return gnx_step_resident_launch(net, co, f_const, g_const, pbc_out, pbc_node, b, x, schur_net, plan, cond, rsrc,
                                    ftot, P, work, rtol, max_iter, warm, ex, reinterpret_cast<unsigned long long*>(sync),
                                    (unsigned long long)epoch, vals, resid_host, info_host, main_st);
  }
  g_last_step_resident = 0;
  cudaStream_t solve_st = side_stream ? (cudaStream_t)side_stream : main_st;
  const bool fork = side_stream != nullptr && vals != nullptr;
  if (fork) {
    GNX_CHECK_ARG(ev_fork && ev_join);
    GNX_CUDA(cudaEventRecord((cudaEvent_t)ev_fork, main_st));
    GNX_CUDA(cudaStreamWaitEvent(solve_st, (cudaEvent_t)ev_fork, 0));
  } else {
    solve_st = main_st;
    if (vals) if (int rc = gnx_assemble_mixed(net, co, rowptr, vals, main_st)) return rc;
  }
  if (int rc = gnx_rhs_eliminate_mixed(net, co, f_const, g_const, pbc_out, pbc_node, b, cond, rsrc, ftot, ex,
                                       solve_st)) return rc;
  int32_t chunk = 0, stride = 0;
  if (ex && ex->mode == 1) {                    // subtree partition: local sweeps, one exchange inside the Schur kernel
    GNX_CHECK_ARG(cond && rsrc && ex->phase == 0);
    if (net->B > 0) {
      double* g = ex->peers_host[ex->rank];
      double* lam = x + gnx_mixed_dims(*net).lam_off;
      if (int rc = gnx_schur_solve(schur_net, plan, co->sigma, g, g + ex->E_glob, g + 2 * (int64_t)ex->E_glob,
                                   b + gnx_mixed_dims(*net).lam_off, nullptr, P, lam, work, rtol, max_iter, warm, 0, 0, ex,
                                   resid_host, info_host, solve_st)) return rc;
    }
  } else if (ex) {                              // the scalars of ALL edges, as the ranks stored them into my buffer
    double* g = ex->peers_host[ex->rank];
    chunk = ex->chunk; stride = 3 * ex->chunk;
    cond = g + (int64_t)ex->rank * stride;      // my own block: what the back-substitution reads
    rsrc = cond + chunk;
    if (net->B > 0) {
      double* lam = x + gnx_mixed_dims(*net).lam_off;
      if (int rc = gnx_schur_solve(schur_net, plan, co->sigma, g, g + chunk, g + 2 * (int64_t)chunk,
                                   b + gnx_mixed_dims(*net).lam_off, nullptr, P, lam, work, rtol, max_iter, warm, chunk,
                                   stride, ex, resid_host, info_host, solve_st)) return rc;
    }
  } else if (net->B > 0) {
    double* lam = x + gnx_mixed_dims(*net).lam_off;
    if (int rc = gnx_schur_solve(schur_net, plan, co->sigma, cond, rsrc, ftot, b + gnx_mixed_dims(*net).lam_off, nullptr,
                                 P, lam, work, rtol, max_iter, warm, 0, 0, nullptr, resid_host, info_host, solve_st)) return rc;
  }
  // Large networks: the edges' end flux and end pressure are gathered from the multipliers ONCE (one thread per edge,
  // written over the scratch arrays ftot / rsrc, dead by now) instead of by every back-substitution block for itself.
  static int flux_on = -1;
  if (flux_on < 0) { const char* e = getenv("GNX_EDGE_FLUX"); flux_on = e ? atoi(e) : 1; }              // A-B knob
  if (flux_on && net->B > 0 && ftot && (!ex || ex->mode == 1) && !co->a_cell && net->n >= 6 && net->n <= 10 &&
      net->E >= 4096 && ((uintptr_t)net->h & 15) == 0 && ((uintptr_t)x & 15) == 0) {
    double* q0 = ftot;
    double* Pu = const_cast<double*>(rsrc);
    if (int rc = gnx_edge_flux_mixed(net, cond, rsrc, P, q0, Pu, solve_st)) return rc;
    if (int rc = gnx_edge_backsub_mixed_flux(net, co, f_const, g_const, pbc_out, pbc_node, q0, Pu, x, solve_st)) return rc;
  } else if (int rc = gnx_edge_backsub_mixed_const(net, co, f_const, g_const, pbc_out, pbc_node, b, cond, rsrc,
                                                   net->B > 0 ? P : nullptr, x, solve_st)) return rc;
  if (fork) {
    GNX_CUDA(cudaEventRecord((cudaEvent_t)ev_join, solve_st));
    if (int rc = gnx_assemble_mixed(net, co, rowptr, vals, main_st)) return rc;
    GNX_CUDA(cudaStreamWaitEvent(main_st, (cudaEvent_t)ev_join, 0));
  }
  return GNX_OK;
}

// ------------------------------------------------------------------------------------
// One theta-scheme time step (graphnics/flowmodels/time_stepping.py:179-216) behind ONE call:
//   Ln1 = L^{n+1}                                    right-hand-side kernel
//   b   = DDn + w1 Ln1 - w0 A x^n + w0 Ln             w1 = theta dt, w0 = (1 - theta) dt; implicit Euler: w0 = 0, no SpMV
//   x1  = (D + w1 A)^-1 b                             elimination -> Schur solve -> back-substitution
//   DDn = D x1                                        mass apply (the next step's first term)
// The kernels are those of the separate entry points (identical results); what this saves is the host: seven
// Python / ctypes round trips per step cost more than the ~70 us the GPU needs for a step on a tree.
// ------------------------------------------------------------------------------------
static int theta_rhs(int64_t N, double w1, double w0, const int32_t* rowptr, const int32_t* colidx, const double* vals_n,
                     const double* x0, const double* DDn, const double* Ln1, const double* Ln, double* ax, double* b,
                     void* stream) {
  double coefs[4] = {1.0, w1, 0.0, 0.0};
  const double* vecs[4] = {DDn, Ln1, nullptr, nullptr};
  int k = 2;
  if (w0 != 0.0) {
    GNX_CHECK_ARG(rowptr && colidx && vals_n && x0 && Ln && ax);
    if (int rc = gnx_spmv_csr_f64((int32_t)N, rowptr, colidx, vals_n, x0, ax, stream)) return rc;
    coefs[2] = -w0; vecs[2] = ax;
    coefs[3] = w0; vecs[3] = Ln;
    k = 4;
  }
  return gnx_lincomb(N, k, coefs, vecs, b, stream);
}

extern "C" int gnx_timestep_mixed(const gnx_network_t* net, const gnx_mixed_coeffs_t* co, const gnx_schur_plan_t* plan,
                                  const double* f_nodal, const double* g_nodal, double f_const, double g_const,
                                  const double* pbc_out, const double* pbc_node, double w1, double w0,
                                  const int32_t* rowptr, const int32_t* colidx, const double* vals_n, const double* x0,
                                  const double* Ln, double mass_coeff, double* DDn, double* Ln1, double* b, double* ax,
                                  double* x1, double* cond, double* rsrc, double* ftot, double* P, double* work,
                                  double rtol, int32_t max_iter, int32_t warm, double* resid_host, int32_t* info_host,
                                  void* stream) {
  GNX_CHECK_ARG(net && co && DDn && Ln1 && b && x1 && cond && rsrc && ftot);
  GNX_CHECK_ARG(net->B == 0 || (plan && P && work));
  const GnxMixedDims d = gnx_mixed_dims(*net);
  if (int rc = gnx_assemble_rhs_mixed(net, f_nodal, g_nodal, f_const, g_const, pbc_out, pbc_node, Ln1, stream)) return rc;
  if (int rc = theta_rhs(d.N, w1, w0, rowptr, colidx, vals_n, x0, DDn, Ln1, Ln, ax, b, stream)) return rc;
  if (int rc = gnx_edge_eliminate_mixed(net, co, b, cond, rsrc, ftot, stream)) return rc;
  if (net->B > 0)
    if (int rc = gnx_schur_solve(net, plan, co->sigma, cond, rsrc, ftot, b + d.lam_off, nullptr, P, x1 + d.lam_off, work,
                                 rtol, max_iter, warm, 0, 0, nullptr, resid_host, info_host, stream)) return rc;
  if (int rc = gnx_edge_backsub_mixed(net, co, b, cond, rsrc, net->B > 0 ? P : nullptr, x1, stream)) return rc;
  return gnx_mass_apply_q_mixed(net, mass_coeff, x1, DDn, stream);
}

// The same for the primal model (TimeDepHydraulicNetwork): f_cell / g_cell as in gnx_assemble_rhs_primal, the Dirichlet
// data bc_node [V] applied to b (xii.apply_bc, time_stepping.py:188), co_mass the coefficient of the q mass term.
extern "C" int gnx_timestep_primal(const gnx_network_t* net, const gnx_primal_coeffs_t* co, const gnx_schur_plan_t* plan,
                                   const int32_t* cells, const double* f_cell, int32_t f_lin, double f_const,
                                   const double* g_cell, int32_t g_lin, double g_const, const double* bc_node, double w1,
                                   double w0, const int32_t* rowptr, const int32_t* colidx, const double* vals_n,
                                   const double* x0, const double* Ln, const gnx_primal_coeffs_t* co_mass, double* DDn,
                                   double* Ln1, double* b, double* ax, double* x1, double* cond, double* rsrc, double* ftot,
                                   double* P, double* work, double rtol, int32_t max_iter, int32_t warm,
                                   double* resid_host, int32_t* info_host, void* stream) {
  GNX_CHECK_ARG(net && co && co_mass && cells && DDn && Ln1 && b && x1 && cond && rsrc && ftot);
  GNX_CHECK_ARG(net->B == 0 || (plan && P && work));
  const int64_t N = gnx_primal_ndofs(net);
  const int64_t NC = (int64_t)net->E << net->n;
  if (int rc = gnx_assemble_rhs_primal(net, cells, f_cell, f_lin, f_const, g_cell, g_lin, g_const, Ln1, stream)) return rc;
  if (int rc = theta_rhs(N, w1, w0, rowptr, colidx, vals_n, x0, DDn, Ln1, Ln, ax, b, stream)) return rc;
  if (bc_node)
    if (int rc = gnx_apply_bc_primal(net, co->sigma, bc_node, nullptr, b, stream)) return rc;
  if (int rc = gnx_edge_eliminate_primal(net, co, b, cond, rsrc, ftot, stream)) return rc;
  if (net->B > 0)
    if (int rc = gnx_schur_solve(net, plan, co->sigma, cond, rsrc, ftot, nullptr, b + NC, P, nullptr, work, rtol, max_iter,
                                 warm, 0, 0, nullptr, resid_host, info_host, stream)) return rc;
  if (int rc = gnx_edge_backsub_primal(net, co, b, cond, rsrc, net->B > 0 ? P : nullptr, x1, stream)) return rc;
  return gnx_mass_apply_q_primal(net, co_mass, x1, DDn, stream);
}
