"""The other BASELINE.json configurations on one GPU (parity-test cases, not bench lines):
C3 honeycomb 200x200 (mixed + primal, n=1 and n=5), C4 pulsatile primal flow on a ~10k-edge
tumour-like graph (1000 implicit-Euler steps, operator reused), C5-size tree (19 generations, n=7,
135 M DOFs) on ONE GPU.  Prints one JSON line per case: sizes, ms, MDOF/s, GB/s of the streaming
kernels, CG iterations, global residual.   python tools/bench_configs.py [c3 c4 c5]"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.engine_primal import PrimalSystem


def timed(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def mixed_case(name, pos, edges, n, pcol):
    t0 = time.perf_counter()
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
    ms.build_pattern(); ms.schur_plan()
    torch.cuda.synchronize()
    setup = time.perf_counter() - t0
    co = ms.coeffs(np.ones(dn.E))
    vals = torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
    b = torch.empty(ms.N, dtype=torch.float64, device="cuda")
    x = torch.empty(ms.N, dtype=torch.float64, device="cuda")
    xq, _ = ms.dof_coordinates(False)
    pbc_out = ms.end_values(xq[:, pcol].contiguous())
    pbc_node = dn.pos[:, pcol].contiguous()
    del xq
    ms._xq = None

    def step():                                   # the one-call step, as bench.py and model.solve() run it
        ms.assemble_rhs_solve(co, vals, b, x, dict(pbc_out=pbc_out, pbc_node=pbc_node), sync=False)
    t = timed(step)
    _, info = ms.solve(co, b, x=x)
    rr, bb = ms.residual(vals, x, b).cpu().tolist()
    N, nnz, NC = ms.N, ms.nnz, dn.num_cells
    ta = timed(lambda: ms.assemble(co, out=vals))
    te = timed(lambda: ms.eliminate(co, b))
    y = torch.empty_like(x)
    tsp = timed(lambda: ms.spmv(vals, x, y))
    hp = ms.schur_plan()[3] if dn.B else None
    print(json.dumps({"case": name, "model": "MixedHydraulicNetwork", "E": dn.E, "n": n, "B": dn.B, "N": N, "nnz": nnz,
                      "setup_s": setup, "ms_per_step": t, "MDOF_per_s": N / t / 1e3,
                      "assembly_GBps": (8 * nnz + 8 * NC) / ta / 1e6, "eliminate_GBps": (8 * N + 8 * NC) / te / 1e6,
                      "spmv_GBps": (12 * nnz + 4 * (N + 1) + 16 * N) / tsp / 1e6, "spmv_ms": tsp,
                      "cg_iterations": info.cg_iterations, "rake_levels": len(hp.levels) if hp else 0,
                      "core": hp.nc if hp else 0, "rel_residual": (rr / bb) ** 0.5}), flush=True)


def primal_case(name, pos, edges, n, pcol):
    dn = DeviceNetwork(pos, edges, n)
    ps = PrimalSystem(dn)
    ps.build_pattern(); ps.schur_plan()
    co = ps.coeffs(1.0)
    vals = torch.empty(ps.nnz, dtype=torch.float64, device="cuda")
    x = torch.empty(ps.N, dtype=torch.float64, device="cuda")
    g_node = dn.pos[:, pcol].contiguous()
    b = torch.empty(ps.N, dtype=torch.float64, device="cuda")

    def step():
        ps.assemble(co, out=vals)
        ps.rhs(None, None, out=b)
        ps.apply_bc_rhs(1.0, g_node, b)
        ps.solve(co, b, x=x, sync=False)
    t = timed(step)
    _, info = ps.solve(co, b, x=x)
    from graphnics_b200._lib import check, ptr
    from graphnics_b200.device import current_stream
    check(ps.lib.gnx_apply_bc_primal(dn.net_ref(), 1.0, ptr(g_node), ptr(vals), None, current_stream()))
    rr, bb = ps.residual(vals, x, b).cpu().tolist()
    print(json.dumps({"case": name, "model": "HydraulicNetwork", "E": dn.E, "n": n, "B": dn.B, "N": ps.N,
                      "ms_per_step": t, "MDOF_per_s": ps.N / t / 1e3, "cg_iterations": info.cg_iterations,
                      "rel_residual": (rr / bb) ** 0.5}), flush=True)


def c3():
    from graphnics_b200.generators import honeycomb
    t0 = time.perf_counter()
    G = honeycomb(200, 200, mesh=False)
    pos, edges = G._graph_arrays()
    print(json.dumps({"case": "C3 graph", "V": len(pos), "E": len(edges), "networkx_s": time.perf_counter() - t0}), flush=True)
    for n in (1, 5):
        mixed_case(f"C3 honeycomb(200,200) n={n}", pos, edges, n, 0)
        primal_case(f"C3 honeycomb(200,200) n={n}", pos, edges, n, 0)


def c4(steps=1000):
    import graphnics_b200 as gn
    from graphnics_b200.synthetic import tumour_like_graph
    pos, edges, radius = tumour_like_graph()
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
    G.add_edges_from(map(tuple, edges.tolist()))
    G.make_mesh(4)
    t = gn.Constant(0)
    f = gn.sin(2 * np.pi * t) * (1.0 + 0.0 * gn.SpatialCoordinate(G.mesh)[0])       # pulsatile source
    model = gn.TimeDepHydraulicNetwork(G, f=f, p_bc=gn.Constant(0), Res=gn.Constant(1.0), Ainv=gn.Constant(1.0))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sols = gn.time_stepping_stokes(model, t=t, t_steps=steps, T=1.0, reassemble_lhs=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    N = sols[-1].vector().size()
    qn = float(torch.linalg.norm(sols[-1].tensor()))
    print(json.dumps({"case": "C4 pulsatile tumour-like graph", "model": "TimeDepHydraulicNetwork", "E": len(edges), "n": 4,
                      "N": N, "steps": steps, "wall_s": dt, "ms_per_step": dt / (steps - 1) * 1e3,
                      "MDOF_per_s": N * (steps - 1) / dt / 1e6, "final_norm": qn}), flush=True)


def ts_tree(steps=200):
    """pulsatile loop on the C2 tree (no Krylov part: a step is ~70 us of kernels, so the HOST decides): wall time per
    step with the one-call step (gnx_timestep_mixed) and with the separate calls (GNX_TIMESTEP_ONE_CALL=0)"""
    import graphnics_b200 as gn
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(11, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
    G.add_edges_from(map(tuple, edges.tolist()))
    G.make_mesh(9)
    for mode in ("1", "0"):
        os.environ["GNX_TIMESTEP_ONE_CALL"] = mode
        t = gn.Constant(0)
        p_bc = gn.Expression("x[1]*(1 + 0.1*sin(6.28*t))", degree=2, t=0.0)
        model = gn.TimeDepMixedHydraulicNetwork(G, p_bc=p_bc, Ainv=gn.Constant(1.0))
        gn.time_stepping_stokes(model, t=t, t_steps=10, T=0.01, reassemble_lhs=False)          # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sols = gn.time_stepping_stokes(model, t=t, t_steps=steps, T=0.2, reassemble_lhs=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        N = sols[-1].vector().size()
        print(json.dumps({"case": "pulsatile loop on the C2 tree (TimeDepMixedHydraulicNetwork, IE, operator reused)",
                          "one_call_step": mode == "1", "N": N, "steps": steps, "ms_per_step": dt / (steps - 1) * 1e3,
                          "MDOF_per_s": N * (steps - 1) / dt / 1e6}), flush=True)


def c5():
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(19, gdim=2)
    mixed_case("C5 tree 19 generations n=7 on ONE GPU", pos, edges, 7, 1)


def c3mg(n=5):
    """C3 edge-partitioned (BASELINE.json configs[2]) under torchrun: honeycomb(200,200), mixed model, N ranks.  The
    lattice has no rakeable bifurcation, so every multiplier is replicated: the ranks split the edge work (assembly,
    elimination, back-substitution) and each runs the whole multilevel PCG -- the time is the PCG's, whatever N."""
    import torch.distributed as dist
    from graphnics_b200.generators import honeycomb
    from graphnics_b200.parallel import PartitionedMixed, global_residual
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    pos, edges = honeycomb(200, 200, mesh=False)._graph_arrays()
    if world > 1:
        pm = PartitionedMixed(pos, edges, n)
        dn, ms = pm.dn, pm.ms
    else:
        dn = DeviceNetwork(pos, edges, n)
        ms = MixedSystem(dn)
    ms.build_pattern(); ms.schur_plan()
    co = ms.coeffs(np.ones(dn.E))
    vals = torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
    b = torch.empty(ms.N, dtype=torch.float64, device="cuda")
    x = torch.zeros(ms.N, dtype=torch.float64, device="cuda")
    xq, _ = ms.dof_coordinates(False)
    kw = dict(pbc_out=ms.end_values(xq[:, 0].contiguous()), pbc_node=dn.pos[:, 0].contiguous())
    del xq
    ms._xq = None

    def step():
        ms.assemble_rhs_solve(co, vals, b, x, kw, sync=False)
    t = timed(step)
    tt = torch.tensor([t], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    _, info = ms.assemble_rhs_solve(co, vals, b, x, kw, sync=True)
    resid = global_residual(ms, vals, x, b) if world > 1 else float(np.sqrt(np.divide(*ms.residual(vals, x, b).cpu().tolist())))
    N = (len(edges) * (2 * (1 << n) + 1) + dn.B)
    if world == 1 or dist.get_rank() == 0:
        print(json.dumps({"case": f"C3 honeycomb(200,200) n={n} edge-partitioned", "world": world, "N": N, "edges_per_rank": dn.E,
                          "ms_per_step_max_over_ranks": float(tt[0]), "MDOF_per_s": N / float(tt[0]) / 1e3,
                          "cg_iterations": info.cg_iterations, "rel_residual": resid,
                          "exchange": ms.exchange_mode if world > 1 else None}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    for w in which:
        {"c3": c3, "c4": c4, "c5": c5, "c3mg": c3mg, "ts_tree": ts_tree}[w]()
        torch.cuda.empty_cache()
