import cProfile, pstats, sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import graphnics_b200 as gn
from graphnics_b200.synthetic import tumour_like_graph
pos, edges, radius = tumour_like_graph()
G = gn.FenicsGraph()
G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
G.add_edges_from(map(tuple, edges.tolist()))
G.make_mesh(4)
t = gn.Constant(0)
f = gn.sin(2 * np.pi * t) * (1.0 + 0.0 * gn.SpatialCoordinate(G.mesh)[0])
model = gn.TimeDepHydraulicNetwork(G, f=f, p_bc=gn.Constant(0), Res=gn.Constant(1.0), Ainv=gn.Constant(1.0))
ps = model._system()
hp = ps.schur_plan()[3]
print("B", ps.dn.B, "rake levels", len(hp.levels), "core", hp.nc, "nnzc", hp.nnzc)
sols = gn.time_stepping_stokes(model, t=t, t_steps=20, T=0.02, reassemble_lhs=False)
print("last info", model.__dict__.get("solve_info"))
pr = cProfile.Profile(); pr.enable()
sols = gn.time_stepping_stokes(model, t=t, t_steps=200, T=0.2, reassemble_lhs=False)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(30)
