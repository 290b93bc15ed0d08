"""Where a pulsatile time step on the C2 tree spends its time: CUPTI kernel records (torch.profiler) of 50 steps of
time_stepping_stokes(TimeDepMixedHydraulicNetwork) -- per kernel: launches per step, mean microseconds, share --
and the wall time per step."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import graphnics_b200 as gn
from graphnics_b200.synthetic import murray_tree
pos, edges, _, _ = murray_tree(11, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
G = gn.FenicsGraph()
G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
G.add_edges_from(map(tuple, edges.tolist()))
G.make_mesh(9)
t = gn.Constant(0)
p_bc = gn.Expression("x[1]*(1 + 0.1*sin(6.28*t))", degree=2, t=0.0)
model = gn.TimeDepMixedHydraulicNetwork(G, p_bc=p_bc, Ainv=gn.Constant(1.0))
gn.time_stepping_stokes(model, t=t, t_steps=10, T=0.01, reassemble_lhs=False)
torch.cuda.synchronize()
steps = 50
t0 = time.perf_counter()
gn.time_stepping_stokes(model, t=t, t_steps=steps, T=0.05, reassemble_lhs=False)
torch.cuda.synchronize()
print("wall us/step", (time.perf_counter() - t0) / (steps - 1) * 1e6)
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    gn.time_stepping_stokes(model, t=t, t_steps=steps, T=0.05, reassemble_lhs=False)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        a = agg[ev.name[:70]]; a[0] += 1; a[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print("GPU busy us/step", tot / (steps - 1))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{k:70s} {v[0] / (steps - 1):6.2f}/step {v[1] / v[0]:8.2f} us {100 * v[1] / tot:5.1f}%")
