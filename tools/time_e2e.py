"""Where the end-to-end call spends its wall time (C2): perf_counter around the parts of
MixedHydraulicNetwork(G, p_bc).solve() + qp.vector().get_local(), medians over 30 calls."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphnics_b200 as gn
from graphnics_b200.synthetic import murray_tree
from graphnics_b200.function import ii_Function
pos, edges, _, _ = murray_tree(11, gdim=2)
G = gn.FenicsGraph()
G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
G.add_edges_from(map(tuple, edges.tolist()))
G.make_mesh(9)
res = gn.Constant(1.0)
for e in G.edges(): G.edges[e]["Res"] = res
p_bc = gn.Expression("x[1]", degree=2)
T = {}
def tick(name, t0):
    t = time.perf_counter(); T.setdefault(name, []).append((t - t0) * 1e6); return t
def step():
    t = time.perf_counter()
    model = gn.MixedHydraulicNetwork(G, p_bc=p_bc); t = tick("model init", t)
    ms = model._system(); W = model.W
    L = model.L_form(); t = tick("L_form (projections of p_bc)", t)
    co = model._coeffs(1.0, 0.0); t = tick("_coeffs (edge attributes -> device)", t)
    ms.build_pattern()
    vals = torch.empty(ms.nnz, dtype=torch.float64, device=ms.dev)
    b = torch.empty(ms.N, dtype=torch.float64, device=ms.dev)
    qp = ii_Function(W, data=torch.empty(ms.N, dtype=torch.float64, device=ms.dev)); t = tick("allocations + ii_Function", t)
    kw = L.rhs_kwargs(); t = tick("rhs_kwargs (p_bc at the nodes)", t)
    ms.assemble_rhs_solve(co, vals, b, qp.tensor(), kw, sync=False); t = tick("assemble_rhs_solve (enqueue)", t)
    qp.touch()
    torch.cuda.current_stream().synchronize(); t = tick("wait for the GPU", t)
    x = qp.vector().get_local(); t = tick("get_local (D2H 16.8 MB)", t)
    return x
for _ in range(5): step()
T.clear()
t0 = time.perf_counter()
for _ in range(30): step()
tot = (time.perf_counter() - t0) / 30 * 1e6
for k, v in T.items(): print(f"{k:45s} {np.median(v):8.1f} us")
print(f"{'total per call':45s} {tot:8.1f} us")
