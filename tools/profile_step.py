"""The bench step (C2 tree, assemble + rhs + solve) a few times, nothing else: the command
profiled under ncu (launch list / --set full).  Optional args: levels n steps."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.synthetic import murray_tree

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 11
n = int(sys.argv[2]) if len(sys.argv) > 2 else 9
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pos, edges, _, _ = murray_tree(levels, gdim=2)
dn = DeviceNetwork(pos, edges, n)
ms = MixedSystem(dn)
ms.build_pattern()
co = ms.coeffs(np.ones(dn.E))
vals = torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
b = torch.empty(ms.N, dtype=torch.float64, device="cuda")
x = torch.empty(ms.N, dtype=torch.float64, device="cuda")
xq, _ = ms.dof_coordinates(False)
pbc_out = ms.end_values(xq[:, 1].contiguous())
pbc_node = dn.pos[:, 1].contiguous()
rhs_kw = dict(pbc_out=pbc_out, pbc_node=pbc_node)
for _ in range(steps):      # the bench step, kernels back to back on one stream (shares are easier to read)
    ms.assemble_rhs_solve(co, vals, b, x, rhs_kw, overlap=False, sync=False)
y = torch.empty_like(x)
for _ in range(3):
    ms.spmv(vals, x, y)
torch.cuda.synchronize()
rr, bb = ms.residual(vals, x, b).cpu().tolist()
print("N", ms.N, "rel residual", (rr / bb) ** 0.5)
