import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import graphnics_b200 as gn
from graphnics_b200.synthetic import murray_tree
pos, edges, _, _ = murray_tree(11, gdim=2)
G = gn.FenicsGraph()
G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
G.add_edges_from(map(tuple, edges.tolist()))
G.make_mesh(9)
res = gn.Constant(1.0)
for e in G.edges(): G.edges[e]["Res"] = res
p_bc = gn.Expression("x[1]", degree=2)
def step():
    model = gn.MixedHydraulicNetwork(G, p_bc=p_bc)
    qp = model.solve()
    return qp.vector().get_local()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(10): step()
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
