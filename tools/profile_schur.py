"""One Schur solve of a 60x60 honeycomb (7440 core unknowns, cluster-resident PCG): ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.generators import honeycomb
G = honeycomb(60, 60, mesh=False); pos, edges = G._graph_arrays()
dn = DeviceNetwork(pos, edges, 1); ms = MixedSystem(dn)
co = ms.coeffs(np.ones(dn.E))
xq, _ = ms.dof_coordinates(False)
b = ms.rhs(None, None, ms.end_values(xq[:, 0].contiguous()), dn.pos[:, 0].contiguous(), f_const=0.1)
for _ in range(3):
    x, info = ms.solve(co, b)
print(info)
