"""Schur kernel alone on binary trees of 12..15 generations (B = 2047..16383): microseconds per solve.
GNX_SCHUR_NO_CLUSTER=1 selects the cooperative-grid sweep instead of the cluster-resident one."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.synthetic import murray_tree
for levels in (11, 12, 13, 14, 15):
    pos, edges, _, _ = murray_tree(levels, gdim=2)
    dn = DeviceNetwork(pos, edges, 2); ms = MixedSystem(dn)
    co = ms.coeffs(np.ones(dn.E))
    xq, _ = ms.dof_coordinates(False)
    b = ms.rhs(None, None, ms.end_values(xq[:, 1].contiguous()), dn.pos[:, 1].contiguous())
    x = torch.empty(ms.N, dtype=torch.float64, device="cuda")
    _, info = ms.solve(co, b, x=x)
    d = ms._bufs(); plan, _, work, hp = ms.schur_plan(); lam = x[dn.lam_off:]
    def schur():
        ms.lib.gnx_schur_solve(dn.schur_net_ref(), C.byref(plan), 1.0, d["cond"].data_ptr(), d["rsrc"].data_ptr(), d["ftot"].data_ptr(),
                               b[dn.lam_off:].data_ptr(), None, d["P"].data_ptr(), lam.data_ptr(), work.data_ptr(), 1e-13, 10, 0, 0, 0, None,
                               None, None, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): schur()
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); schur(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    torch.cuda.synchronize()
    scal = work[work.numel() - 8:].cpu().numpy()          # block 0's SM cycles (gnx_solve.cu: scal[5..7])
    print(json.dumps({"levels": levels, "B": dn.B, "ctas": info.cg_batches, "top_S": plan.top_S, "l_split": plan.l_split,
                      "us": min(ts) * 1e3, "cycles_phase0_and_wide_levels": scal[5], "cycles_top_prologue": scal[6],
                      "cycles_top_sweeps": scal[7]}), flush=True)
