"""Development timing of the mixed hot path on one GPU (CUDA events).  Not the contract bench."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.synthetic import murray_tree

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 11
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9
pos, edges, radius, level = murray_tree(levels)
print("E", len(edges), "n", n)
t0 = time.time(); dn = DeviceNetwork(pos, edges, n); torch.cuda.synchronize(); print("mesh+tables s", time.time() - t0)
ms = MixedSystem(dn)
print("N", ms.N, "nnz", ms.nnz, "B", dn.B)

def timed(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

ms.build_pattern()
co = ms.coeffs(1.0)
vals = ms.assemble(co)
pbc_node = torch.from_numpy(np.ascontiguousarray(pos[:, 1])).cuda()
xq, _ = ms.dof_coordinates(False)
pbc_nodal = ms.end_values(xq[:, 1].contiguous())
b = ms.rhs(None, None, pbc_nodal, pbc_node)
x = torch.empty(ms.N, dtype=torch.float64, device="cuda")
def pat():
    ms.rowptr = None; ms.build_pattern()
print("pattern ms", timed(pat))
print("assemble ms", timed(lambda: ms.assemble(co, out=vals)))
print("rhs ms", timed(lambda: ms.rhs(None, None, pbc_nodal, pbc_node, out=b)))
print("eliminate ms", timed(lambda: ms.eliminate(co, b)))
info = None
def solve():
    global info
    _, info = ms.solve(co, b, x=x)
print("solve ms", timed(solve), info)
y = torch.empty_like(x)
t = timed(lambda: ms.spmv(vals, x, y))
bytes_spmv = 12 * ms.nnz + 4 * (ms.N + 1) + 16 * ms.N
print("spmv ms", t, "GB/s", bytes_spmv / t[0] / 1e6)
print("residual", ms.residual(vals, x, b).cpu().tolist())
def step():
    ms.assemble(co, out=vals); ms.rhs(None, None, pbc_nodal, pbc_node, out=b); ms.solve(co, b, x=x, sync=False)
t = timed(step)
print("assemble+solve (async) ms", t, "MDOF/s", ms.N / t[0] / 1e3)
print("solve async:", timed(lambda: ms.solve(co, b, x=x, sync=False)))
print("solve async max_iter=1:", timed(lambda: ms.solve(co, b, x=x, sync=False, max_iter=1)))
try:
    g = torch.cuda.CUDAGraph()
    s_ = torch.cuda.Stream()
    s_.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s_):
        step(); step()
    torch.cuda.current_stream().wait_stream(s_)
    with torch.cuda.graph(g):
        step()
    t = timed(lambda: g.replay())
    print("graph step ms", t, "MDOF/s", ms.N / t[0] / 1e3)
    print("residual after graph", ms.residual(vals, x, b).cpu().tolist())
except Exception as ex:
    print("whole-step graph capture failed:", repr(ex))
