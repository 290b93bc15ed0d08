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
  if (int rc = gnx_edge_backsub_mixed_const(net, co, f_const, g_const, pbc_out, pbc_node, b, cond, rsrc,
                                            net->B > 0 ? P : nullptr, x, solve_st)) return rc;
  if (fork) {
    GNX_CUDA(cudaEventRecord((cudaEvent_t)ev_join, solve_st));
    if (int rc = gnx_assemble_mixed(net, co, rowptr, vals, main_st)) return rc;
    GNX_CUDA(cudaStreamWaitEvent(main_st, (cudaEvent_t)ev_join, 0));
  }
  return GNX_OK;
}
