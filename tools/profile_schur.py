"""the one-CTA Schur solve of the C2 tree, a few launches (for `ncu -k regex:schur_solve_kernel`)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.synthetic import murray_tree
pos, edges, _, _ = murray_tree(11, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
dn = DeviceNetwork(pos, edges, 9)
ms = MixedSystem(dn)
co = ms.coeffs(np.ones(dn.E))
xq, _ = ms.dof_coordinates(False)
b = ms.rhs(None, None, ms.end_values(xq[:, 1].contiguous()), dn.pos[:, 1].contiguous())
x = torch.empty(ms.N, dtype=torch.float64, device="cuda")
for _ in range(6):
    ms.solve(co, b, x=x, sync=False)
torch.cuda.synchronize()
print("ok")
