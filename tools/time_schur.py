"""Time the Schur kernel alone (CUDA events) on the tumour-like graph and the honeycomb; prints
iterations and microseconds per iteration.  GNX_SCHUR_NO_CLUSTER=1 selects the cooperative-grid path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.synthetic import tumour_like_graph

def run(name, pos, edges, n):
    dn = DeviceNetwork(pos, edges, n); ms = MixedSystem(dn)
    co = ms.coeffs(np.ones(dn.E))
    xq, _ = ms.dof_coordinates(False)
    b = ms.rhs(None, None, ms.end_values(xq[:, 0].contiguous()), dn.pos[:, 0].contiguous(), f_const=0.1)
    x = torch.empty(ms.N, dtype=torch.float64, device="cuda")
    _, info = ms.solve(co, b, x=x)
    d = ms._bufs(); plan, _, work, hp = ms.schur_plan()
    lam = x[dn.lam_off:]
    def schur(max_iter):
        ms.lib.gnx_schur_solve(dn.schur_net_ref(), C.byref(plan), 1.0, d["cond"].data_ptr(), d["rsrc"].data_ptr(), d["ftot"].data_ptr(),
                               b[dn.lam_off:].data_ptr(), None, d["P"].data_ptr(), lam.data_ptr(), work.data_ptr(), 1e-13, max_iter, 0, 0, 0, None,
                               None, None, torch.cuda.current_stream().cuda_stream)
    out = {"case": name, "B": dn.B, "core": hp.nc, "iterations": info.cg_iterations, "ctas": info.cg_batches}
    for mi in (1, 101, 100000):
        for _ in range(2): schur(mi)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); schur(mi); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        out[f"ms_maxiter_{mi}"] = min(ts)
    out["us_per_iteration"] = (out["ms_maxiter_101"] - out["ms_maxiter_1"]) * 10.0
    print(json.dumps(out), flush=True)

pos, edges, _ = tumour_like_graph()
run("tumour", pos, edges, 2)
from graphnics_b200.generators import honeycomb
G = honeycomb(200, 200, mesh=False); pos, edges = G._graph_arrays()
run("honeycomb200", pos, edges, 1)
G = honeycomb(60, 60, mesh=False); pos, edges = G._graph_arrays()
run("honeycomb60", pos, edges, 1)
