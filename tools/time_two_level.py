"""Two-level PCG against plain Jacobi-PCG on cyclic cores (GNX_SCHUR_TWO_LEVEL = 1 | 0): iterations and
microseconds of the whole solve (eliminate + Schur kernel + back-substitute), honeycomb lattices."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    from graphnics_b200.generators import honeycomb
    k, n = int(sys.argv[2]), int(sys.argv[3])
    pos, edges = honeycomb(k, k, mesh=False)._graph_arrays()
    dn = DeviceNetwork(pos, edges, n); ms = MixedSystem(dn); ms.build_pattern()
    import time
    t0 = time.perf_counter(); ms.schur_plan(); t_plan = time.perf_counter() - t0
    rng = np.random.default_rng(0)
    co = ms.coeffs(torch.from_numpy(rng.uniform(0.5, 2.0, dn.E)).cuda())
    vals = ms.assemble(co)
    xq, _ = ms.dof_coordinates(False)
    b = ms.rhs(None, None, ms.end_values(xq[:, 1].contiguous()), dn.pos[:, 1].contiguous())
    x = torch.empty_like(b)
    _, info = ms.solve(co, b, x=x)
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ms.solve(co, b, x=x, sync=False); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    r = torch.empty_like(x); ms.residual(vals, x, b, r)
    print(json.dumps({"two_level": os.environ.get("GNX_SCHUR_TWO_LEVEL"), "honeycomb": k, "n": n, "B": dn.B, "N": ms.N,
                      "plan_s": round(t_plan, 3), "iterations": info.cg_iterations, "ctas": info.cg_batches,
                      "solve_ms": float(np.median(ts)), "rel_resid": float(r.norm() / b.norm())}))
else:
    for k in (70, 100, 200):
        for tl in ("0", "1"):
            r = subprocess.run([sys.executable, __file__, "child", str(k), "1"], env=dict(os.environ, GNX_SCHUR_TWO_LEVEL=tl),
                               capture_output=True, text=True, timeout=170)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-1500:], flush=True)
