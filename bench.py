#!/usr/bin/env python
"""Headline benchmark: MixedHydraulicNetwork assemble+solve on a synthetic arterial tree
(BASELINE.json configs[1]: binary Murray tree, generations 0..10 = 2047 edges, 2^9 cells per
edge = 1 048 064 cells, N = 2 099 198 mixed DOFs), reported as MDOF/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = what `MixedHydraulicNetwork.solve()` does after the mesh exists
(graphnics/flowmodels/flow_models.py:337-350): assemble the monolithic CSR values, assemble the
right-hand side, solve.  Mesh, CSR pattern and solver workspaces are built once, outside the timed
region, and reported as `built_once_ms` (SURVEY.md 8d "built once").

  value        device-resident inputs, CUDA events around every step (L2 flushed between steps)
  e2e          the user-facing call `MixedHydraulicNetwork(G, p_bc=...).solve()` +
               `qp.vector().get_local()` : host (numpy) coefficient data in, host solution out
  roofline     the step's dominant kernel, algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  parity       the timed step's b / x against the CPU oracle on the same workload
  cpu_baseline the numpy/scipy oracle (oracle/, the CPU restatement of the reference path) on the
               same workload on this box's host cores
  at_scale     the same kernels where HBM, not latency, is the limit (N > 1: BASELINE.json configs[4], the
               ~100 M-DOF tree across the GPUs)
`--impl reference` times that CPU restatement alone (the real reference needs FEniCS + fenics_ii +
PETSc/MUMPS, none of which can be installed offline -- DESIGN.md).
N > 1 (torchrun, one rank per GPU): weak scaling, one more tree generation per doubling, the tree cut into
subtrees (graphnics_b200/partition.py).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "MDOF/s (assemble+solve) MixedHydraulicNetwork on synthetic tree"
TREE_LEVELS, TREE_N = 11, 9
SCALE_LEVELS, SCALE_N = 16, 7          # at_scale at N GPUs: 16 + log2(N) generations, 2^7 cells per edge (N = 8: C5)
# SURVEY.md 8(d) per-unit algorithmic bytes: assembly ~110 B/cell, edge-scan elimination ~32 B/DOF
ASM_B_PER_CELL, ELIM_B_PER_DOF = 110, 32


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def tree_arrays(levels=TREE_LEVELS):
    from graphnics_b200.synthetic import murray_tree
    pos, edges, radius, level = murray_tree(levels, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
    return pos, edges


def config_dict(levels, n, E, N, extra=None):
    d = {"workload": f"C2 synthetic binary arterial tree: {levels} generations, E={E} edges, 2^{n} cells/edge, "
                     f"N={N} DOFs; Res=1, p_bc=x[1], f=g=0 (demo/Tree Profiling.ipynb model); steady "
                     f"MixedHydraulicNetwork assemble+solve",
         "levels": levels, "n": n, "edges": E, "ndofs": N, "l2": "flushed between timed steps (256 MiB write)",
         "timing": "a pair of CUDA events around every step, the flush between the pairs; the host enqueues the K steps without "
                   "waiting in between"}
    d.update(extra or {})
    return d


def step_stats(ts):
    ts = np.asarray(ts, dtype=np.float64)
    return {"min": float(ts.min()), "median": float(np.median(ts)), "max": float(ts.max()), "mean": float(ts.mean())}


def count_launches(fn, reps):
    """kernels launched by `reps` calls of fn, counted by CUPTI (torch.profiler) on a replay outside the timed
    region: (total, {kernel name: launches per call})"""
    try:
        import torch
        from torch.profiler import profile, ProfilerActivity
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
        names = {}
        for ev in prof.events():
            if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and not ev.name.lower().startswith("memcpy") \
                    and not ev.name.lower().startswith("memset"):
                names[ev.name] = names.get(ev.name, 0) + 1
        total = int(sum(names.values()))
        return total, {k[:90]: v / reps for k, v in sorted(names.items())}
    except Exception as ex:                      # no CUPTI on this box: say so instead of inventing a number
        return None, {"unavailable": repr(ex)[:200]}


# --------------------------------------------------------------------------------------------
# CPU restatement (oracle) -- cpu_baseline, parity and --impl reference
# --------------------------------------------------------------------------------------------
def oracle_step_factory(pos, edges, n):
    import oracle
    om = oracle.build_mesh(pos, edges, n)                    # built once, like the GPU arm
    pbc = oracle.Coef(lambda x: x[..., 1], 1)
    N = oracle.mixed_layout(om)["N"]

    def step():
        A = oracle.assemble_mixed(om, 1.0)
        b = oracle.assemble_mixed_rhs(om, p_bc=pbc, project=False)
        return oracle.lu_solve(A, b)
    step.om = om
    return step, N


def cpu_baseline(pos, edges, budget_s=20.0):
    """oracle on the full workload if it fits the budget, else on a coarser refinement.  Returns the baseline
    record and (n, x) of the last oracle solve for the parity check."""
    n = TREE_N
    while True:
        step, N = oracle_step_factory(pos, edges, n)
        t0 = time.perf_counter(); x = step(); t = time.perf_counter() - t0
        if t <= budget_s or n <= 4:
            break
        n -= 2
    reps = max(1, min(3, int(budget_s / max(t, 1e-3)) - 1))
    ts = [t]
    for _ in range(reps):
        t0 = time.perf_counter(); x = step(); ts.append(time.perf_counter() - t0)
    best = min(ts)
    rec = {"value": N / best / 1e6, "unit": "MDOF/s", "cores": 1, "kind": "port",
           "sample": f"numpy/scipy oracle (COO->CSR assembly + SuperLU), same tree, 2^{n} cells/edge, N={N}, "
                     f"best of {len(ts)} runs, {best:.2f} s per assemble+solve",
           "seconds_per_step": best}
    return rec, (n, x, step.om)


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def parity_block(om, x_ref, route, xg, bg=None, rowptr=None, colidx=None, vals=None):
    """the timed step's result against the oracle (tests/test_gpu_bench_configs.py asserts the same quantities)"""
    import oracle
    lay = oracle.mixed_layout(om)
    out = {"oracle": route,
           "rel_l2_q": rel(xg[:lay["p_off"]], x_ref[:lay["p_off"]]),
           "rel_l2_p": rel(xg[lay["p_off"]:lay["lam_off"]], x_ref[lay["p_off"]:lay["lam_off"]]),
           "rel_l2_lam": rel(xg[lay["lam_off"]:], x_ref[lay["lam_off"]:])}
    if bg is not None:
        b_ref = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda x: x[..., 1], 1), project=False)
        out["rhs_max_abs_diff"] = float(np.abs(bg - b_ref).max())
    if vals is not None:
        A = oracle.assemble_mixed(om, 1.0)
        out["pattern_equal"] = bool(np.array_equal(rowptr, A.indptr) and np.array_equal(colidx, A.indices))
        out["values_max_rel_diff"] = float(np.abs(vals - A.data).max() / np.abs(A.data).max())
    out["ok"] = bool(max(out["rel_l2_q"], out["rel_l2_p"], out["rel_l2_lam"]) <= 1e-8 and out.get("pattern_equal", True))
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the arm's own workload: at N GPUs the tree has log2(N) more generations (weak scaling).  The bounded
    # sample keeps that tree and coarsens the cells per edge: it starts from the single-GPU number of unknowns
    grow = int(np.ceil(np.log2(max(args.gpus, 1))))
    levels = TREE_LEVELS + grow
    pos, edges = tree_arrays(levels)
    n = max(TREE_N - grow, 4)
    n = min(n, int(os.environ.get("GNX_BENCH_REF_MAX_N", n)))      # (tests: a smaller sample)
    step, N = oracle_step_factory(pos, edges, n)
    t0 = time.perf_counter(); step(); t = time.perf_counter() - t0
    while t * (args.steps + args.warmup) > 240.0 and n > 4:       # bounded sample
        n -= 1
        step, N = oracle_step_factory(pos, edges, n)
        t0 = time.perf_counter(); step(); t = time.perf_counter() - t0
    for _ in range(args.warmup):
        step()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter(); step(); ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    val = N / dt / 1e6
    E_full = len(edges)
    full_N = E_full * (2 * (1 << TREE_N) + 1) + (E_full - 1) // 2
    sample = (f"numpy/scipy oracle of the reference path (FEniCS/fenics_ii/MUMPS not installable offline), "
              f"same tree ({levels} generations), 2^{n} cells/edge, N={N}")
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "MDOF/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(levels, TREE_N, len(edges), full_N,
                                 {"sample_n": n, "sample_ndofs": N, "same_workload": bool(n == TREE_N),
                                  "note": None if n == TREE_N else
                                  f"bounded sample: the same tree at 2^{n} instead of 2^{TREE_N} cells per edge -- the "
                                  f"per-DOF throughput of a COARSER refinement, not the GPU arm's workload"}),
           "ms_per_step_stats": step_stats(np.asarray(ts) * 1e3),
           "cpu_baseline": {"value": val, "unit": "MDOF/s", "cores": 1, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "MDOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def timed_steps(torch, fn, reps, warm, flush):
    host_us = []
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    # Every step is bracketed by its own pair of CUDA events, the L2 flush sits between the pairs, and the host does
    # NOT wait between steps: it runs ahead of the GPU (a step + flush take 110 us on the device, 30 us to enqueue), so a
    # host hiccup -- the nvidia-smi clock sampler polling the driver, a Python GC pause -- no longer lands between an
    # event and the kernel it brackets (one such 0.2 ms outlier in 20 steps moved the mean by 20 %).
    evs = []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); fn(); host_us.append((time.perf_counter() - t0) * 1e6); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = [e0.elapsed_time(e1) for e0, e1 in evs]
    return ts, host_us


def pin_to_gpu_numa(local):
    """bind this rank's host threads (and with them its pinned staging buffers, first touch) to the CPUs NVML names as
    local to its GPU: eight ranks copying 17 MB each to the host at once otherwise share one memory controller"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        # keep stdout to the one JSON line: whatever native libraries write to file descriptor 1 (the NCCL
        # version banner) goes to stderr, Python's own stdout keeps the original descriptor
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = os.fdopen(keep, "w", buffering=1)
        pin_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=dev)
        return run_partitioned(args, dev)

    pos, edges = tree_arrays()
    n = TREE_N
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dn = DeviceNetwork(pos, edges, n, device=dev)
    torch.cuda.synchronize()
    t_mesh = time.perf_counter() - t0
    ms = MixedSystem(dn)
    t0 = time.perf_counter(); ms.build_pattern(); torch.cuda.synchronize(); t_pat = time.perf_counter() - t0
    t0 = time.perf_counter(); ms.schur_plan(); torch.cuda.synchronize(); t_plan = time.perf_counter() - t0
    built_once = {"mesh_tables_ms": t_mesh * 1e3, "csr_pattern_ms": t_pat * 1e3, "schur_plan_host_ms": t_plan * 1e3,
                  "total_ms": (t_mesh + t_pat + t_plan) * 1e3,
                  "note": "first call in the process: includes CUDA context / module load; not part of the step"}
    N, nnz, NC = ms.N, ms.nnz, dn.num_cells
    Res_host = np.ones(dn.E)
    co = ms.coeffs(Res_host)
    vals = torch.empty(nnz, dtype=torch.float64, device=dev)
    b = torch.empty(N, dtype=torch.float64, device=dev)
    x = torch.empty(N, dtype=torch.float64, device=dev)
    xq, _ = ms.dof_coordinates(False)
    pbc_out = ms.end_values(xq[:, 1].contiguous())
    pbc_node = dn.pos[:, 1].contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rhs_kw = dict(pbc_out=pbc_out, pbc_node=pbc_node)

    def step():
        ms.assemble_rhs_solve(co, vals, b, x, rhs_kw, overlap=not args.no_overlap, sync=False)

    def timed(fn, reps, warm, do_flush=True):
        return timed_steps(torch, fn, reps, warm, flush if do_flush else None)[0]

    # ---- headline: K timed steps
    clocks = ClockSampler(local).start()
    ts, host_us = timed_steps(torch, step, args.steps, max(args.warmup, 3), flush)
    torch.cuda.synchronize()
    ms_per_step = float(np.sum(ts) / args.steps)
    value = N / (ms_per_step * 1e-3) / 1e6
    resident = bool(int(ms.lib.gnx_step_is_resident()))
    launches, launch_names = count_launches(step, args.steps)

    # ---- accuracy of what was timed: monolithic residual + (below) the oracle
    rr, bb = ms.residual(vals, x, b).cpu().tolist()
    rel_resid = float(np.sqrt(rr / bb))
    x_timed, b_timed = x.cpu().numpy(), b.cpu().numpy()
    vals_timed = vals.cpu().numpy()

    # ---- per-kernel breakdown (each kernel alone, L2 flushed) and roofline
    import ctypes
    bufs = ms._bufs()
    kernels = {}

    def k(name, fn, bytes_):
        t = float(np.mean(timed(fn, 10, 3)))
        kernels[name] = {"ms": t, "algorithmic_bytes": int(bytes_), "GBps": bytes_ / t / 1e6}

    step_bytes_survey = ASM_B_PER_CELL * NC + ELIM_B_PER_DOF * N
    step_bytes_moved = 8 * nnz + 8 * NC + 2 * 8 * N                 # values + h once + b, x written (nothing re-read)
    k("step kernel(s) as timed (assemble + rhs + solve)", step, step_bytes_survey)
    k("mixed_values_tile_kernel (assembly alone)", lambda: ms.assemble(co, out=vals), 8 * nnz + 8 * NC)
    k("mixed_rhs_kernel", lambda: ms.rhs(None, None, pbc_out, pbc_node, out=b), 8 * N)
    k("edge_eliminate_kernel", lambda: ms.lib.gnx_edge_eliminate_mixed(
        dn.net_ref(), co.ref(), b.data_ptr(), bufs["cond"].data_ptr(), bufs["rsrc"].data_ptr(),
        bufs["ftot"].data_ptr(), torch.cuda.current_stream().cuda_stream), 8 * N + 8 * NC)
    plan, _, work, hp = ms.schur_plan()
    lam = x[dn.lam_off:]
    k("schur_solve_kernel (rake + core PCG, latency-bound)", lambda: ms.lib.gnx_schur_solve(
        dn.net_ref(), ctypes.byref(plan), 1.0, bufs["cond"].data_ptr(), bufs["rsrc"].data_ptr(),
        bufs["ftot"].data_ptr(), b[dn.lam_off:].data_ptr(), None, bufs["P"].data_ptr(), lam.data_ptr(), work.data_ptr(),
        1e-13, 100000, 0, 0, 0, None, None, None, torch.cuda.current_stream().cuda_stream), 8 * 3 * dn.E + 8 * 6 * dn.B)
    k("edge_backsub_kernel", lambda: ms.lib.gnx_edge_backsub_mixed(
        dn.net_ref(), co.ref(), b.data_ptr(), bufs["cond"].data_ptr(), bufs["rsrc"].data_ptr(),
        bufs["P"].data_ptr(), x.data_ptr(), torch.cuda.current_stream().cuda_stream), 2 * 8 * N + 8 * NC)
    k("eliminate+schur+backsub (kernel chain, 3 launches)", lambda: ms.solve(co, b, x=x, sync=False), 3 * 8 * N + 2 * 8 * NC)
    y = torch.empty_like(x)
    spmv_bytes = 12 * nnz + 4 * (N + 1) + 16 * N
    k("spmv_kernel (not in the step; Krylov building block)", lambda: ms.spmv(vals, x, y), spmv_bytes)
    k("reference point: torch copy_ of N doubles (not in the step)", lambda: y.copy_(x), 16 * N)
    hbm, peak_src = peaks()
    if resident:
        dom = "step kernel(s) as timed (assemble + rhs + solve)"
        dom_name = "step_resident_kernel (ONE persistent kernel: rhs + elimination | Schur solve + CSR assembly | back-substitution)"
    else:
        dom = max((kk for kk in kernels if "not in the step" not in kk and "chain" not in kk and "latency" not in kk
                   and "as timed" not in kk), key=lambda kk: kernels[kk]["ms"])
        dom_name = dom
    # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture (profiles/), if any
    traffic, tsrc = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        for key, rec in tj.get("kernels", {}).items():
            if key in dom_name:
                traffic, tsrc = rec["dram_bytes_read"] + rec["dram_bytes_write"], tj.get("source")
    roof = {"bound": "hbm", "kernel": dom_name, "achieved": kernels[dom]["GBps"], "peak": hbm, "unit": "GB/s",
            "frac": kernels[dom]["GBps"] / hbm, "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src,
            "algorithmic_bytes": kernels[dom]["algorithmic_bytes"],
            "algorithmic_bytes_rule": (f"SURVEY.md 8(d): assembly {ASM_B_PER_CELL} B/cell x {NC} cells + edge-scan elimination "
                                       f"{ELIM_B_PER_DOF} B/DOF x {N} DOFs" if resident else "see DESIGN.md section 4"),
            "bytes_the_kernel_has_to_move": int(step_bytes_moved) if resident else None,
            "ms": kernels[dom]["ms"]}

    # ---- the same kernels where HBM, not launch latency, is the limit: a 17-generation tree, 2^7 cells per
    #      edge (33.7 M DOFs, 1.1 GB of CSR values; every array is many times the 126 MB L2)
    at_scale = kernels_at_scale(hbm, timed)

    # ---- e2e through the drop-in Python API, host data in / host solution out
    e2e = e2e_python_api(pos, edges, n, args)
    # ---- the pulsatile time loop (BASELINE.json configs[3]) through the drop-in API
    time_loop = time_stepping_block(pos, edges, n)
    clk = clocks.stop()

    out = {"metric": METRIC, "value": value, "unit": "MDOF/s", "n_gpus": 1, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": config_dict(TREE_LEVELS, n, dn.E, N, {"nnz": nnz, "bifurcations": dn.B}),
           "ms_per_step_stats": step_stats(ts), "built_once_ms": built_once,
           "rel_residual": rel_resid, "gpu_launches": launches, "gpu_launches_per_step": launch_names,
           "gpu_launches_how": "CUPTI (torch.profiler) kernel records of a replay of the same K steps after the timed region",
           "host_enqueue_us_per_step": float(np.median(host_us)),
           "schur": {"rake_levels": len(hp.levels), "core": hp.nc},
           "step_path": ("ONE persistent cooperative kernel (cell lengths resident in shared memory; CTA 0 runs the graph-level "
                         "solve while the other SMs assemble the CSR values of their own edges)" if resident else
                         ("kernel chain on one stream" if args.no_overlap else
                          "kernel chain: CSR assembly on the caller's stream overlaps rhs -> eliminate -> Schur -> back-substitute "
                          "on a high-priority stream")),
           "kernels": kernels, "roofline": roof, "at_scale": at_scale, "e2e": e2e, "time_stepping": time_loop,
           "clocks": clk}
    if not args.no_cpu:
        out["cpu_baseline"], (n_or, x_or, om) = cpu_baseline(pos, edges)
        import oracle
        if n_or == n:
            out["parity"] = parity_block(om, x_or, "LU (SuperLU) of the oracle's assembled matrix, same tree, same refinement",
                                         x_timed, b_timed, ms.rowptr.cpu().numpy(), ms.colidx.cpu().numpy(), vals_timed)
        else:                                    # the LU was too slow for the budget here: elimination route at full size
            om = oracle.build_mesh(pos, edges, n)
            b_ref = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda xx: xx[..., 1], 1), project=False)
            out["parity"] = parity_block(om, oracle.solve_by_elimination(om, b_ref, 1.0),
                                         "oracle elimination route (oracle/elimination.py)", x_timed, b_timed)
    print(json.dumps(out))


def time_stepping_block(pos, edges, n, tree_steps=100, c4_steps=300):
    """time_stepping_stokes (flowmodels/time_stepping.py:126-221) through the drop-in API, wall clock over the whole
    loop (operator assembled once, right-hand side re-projected every step, every state stored, as the reference does):
    the C2 tree with a time-dependent p_bc (mixed model, no Krylov part) and C4 (primal model on a tumour-like graph
    with cycles: one warm-started multilevel-PCG solve per step).  Parity of both loops against the oracle's literal
    replay: tests/test_gpu_bench_configs.py, tests/test_gpu_api_mixed.py."""
    import torch
    import graphnics_b200 as gn
    from graphnics_b200.synthetic import tumour_like_graph
    out = {}
    G = api_graph(pos, edges, n)
    t = gn.Constant(0)
    p_bc = gn.Expression("x[1]*(1 + 0.1*sin(6.28*t))", degree=2, t=0.0)
    model = gn.TimeDepMixedHydraulicNetwork(G, p_bc=p_bc, Ainv=gn.Constant(1.0))
    gn.time_stepping_stokes(model, t=t, t_steps=10, T=0.01, reassemble_lhs=False)            # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sols = gn.time_stepping_stokes(model, t=t, t_steps=tree_steps, T=0.1, reassemble_lhs=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    N = sols[-1].vector().size()
    out["tree"] = {"workload": "C2 tree, TimeDepMixedHydraulicNetwork, implicit Euler, p_bc = x[1] (1 + 0.1 sin(6.28 t)), "
                               "operator reused, one library call per step (gnx_timestep_mixed)",
                   "ndofs": N, "steps": tree_steps - 1, "ms_per_step": dt / (tree_steps - 1) * 1e3,
                   "MDOF_per_s": N * (tree_steps - 1) / dt / 1e6}
    del sols, model, G
    torch.cuda.empty_cache()
    tp, te, _ = tumour_like_graph()
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(tp.tolist()))
    G.add_edges_from(map(tuple, te.tolist()))
    G.make_mesh(4)
    t = gn.Constant(0)
    f = gn.sin(2 * np.pi * t) * (1.0 + 0.0 * gn.SpatialCoordinate(G.mesh)[0])
    model = gn.TimeDepHydraulicNetwork(G, f=f, p_bc=gn.Constant(0), Res=gn.Constant(1.0), Ainv=gn.Constant(1.0))
    gn.time_stepping_stokes(model, t=t, t_steps=5, T=0.005, reassemble_lhs=False)            # warm-up (plan, aggregates)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sols = gn.time_stepping_stokes(model, t=t, t_steps=c4_steps, T=c4_steps / 1000.0, reassemble_lhs=False)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    N = sols[-1].vector().size()
    out["c4"] = {"workload": f"C4 pulsatile flow on a synthetic tumour-like graph with cycles ({len(te)} edges, 2^4 cells/edge), "
                             "TimeDepHydraulicNetwork, implicit Euler, operator reused, warm-started multilevel PCG per step "
                             f"({c4_steps} of the configuration's 1000 steps)",
                 "ndofs": N, "steps": c4_steps - 1, "ms_per_step": dt / (c4_steps - 1) * 1e3,
                 "MDOF_per_s": N * (c4_steps - 1) / dt / 1e6}
    del sols, model, G
    torch.cuda.empty_cache()
    return out


def kernels_at_scale(hbm, timed, levels=17, n=7):
    import torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(levels, gdim=2)
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
    ms.build_pattern()
    N, nnz, NC = ms.N, ms.nnz, dn.num_cells
    co = ms.coeffs(np.ones(dn.E))
    vals = torch.empty(nnz, dtype=torch.float64, device=dn.device)
    b = torch.empty(N, dtype=torch.float64, device=dn.device)
    x = torch.empty(N, dtype=torch.float64, device=dn.device)
    y = torch.empty(N, dtype=torch.float64, device=dn.device)
    xq, _ = ms.dof_coordinates(False)
    pbc_out = ms.end_values(xq[:, 1].contiguous())
    pbc_node = dn.pos[:, 1].contiguous()
    del xq
    ms._xq = None
    bufs = ms._bufs()
    st = lambda: torch.cuda.current_stream().cuda_stream
    out = {"workload": f"binary tree, {levels} generations, E={dn.E}, 2^{n} cells/edge, N={N}, nnz={nnz} (L2 not flushed: "
                       f"every array is >> L2)", "kernels": {}}

    def k(name, fn, bytes_):
        t = float(np.mean(timed(fn, 5, 2)))
        out["kernels"][name] = {"ms": t, "algorithmic_bytes": int(bytes_), "GBps": bytes_ / t / 1e6,
                                "frac_of_measured_hbm_peak": bytes_ / t / 1e6 / hbm}
    k("mixed_values_tile_kernel (assembly)", lambda: ms.assemble(co, out=vals), 8 * nnz + 8 * NC)
    k("mixed_rhs_kernel", lambda: ms.rhs(None, None, pbc_out, pbc_node, out=b), 8 * N)
    k("edge_eliminate_tile_kernel", lambda: ms.eliminate(co, b), 8 * N + 8 * NC)
    ms.solve(co, b, x=x, sync=False)
    k("edge_backsub_tile_kernel", lambda: ms.lib.gnx_edge_backsub_mixed(
        dn.net_ref(), co.ref(), b.data_ptr(), bufs["cond"].data_ptr(), bufs["rsrc"].data_ptr(),
        bufs["P"].data_ptr(), x.data_ptr(), st()), 2 * 8 * N + 8 * NC)
    # Constant data (what the step runs): b is produced by the elimination kernel and never read again
    import ctypes
    k("edge_const_tile_kernel<eliminate> (rhs + elimination, Constant data)", lambda: ms.lib.gnx_rhs_eliminate_mixed(
        dn.net_ref(), co.ref(), 0.0, 0.0, pbc_out.data_ptr(), pbc_node.data_ptr(), b.data_ptr(), bufs["cond"].data_ptr(),
        bufs["rsrc"].data_ptr(), bufs["ftot"].data_ptr(), None, st()), 8 * N + 8 * NC)
    k("edge_const_tile_kernel<backsub> (Constant data: b recomputed, not read)", lambda: ms.lib.gnx_edge_backsub_mixed_const(
        dn.net_ref(), co.ref(), 0.0, 0.0, pbc_out.data_ptr(), pbc_node.data_ptr(), b.data_ptr(), bufs["cond"].data_ptr(),
        bufs["rsrc"].data_ptr(), bufs["P"].data_ptr(), x.data_ptr(), st()), 8 * N + 8 * NC)
    # ... and what the one-call step runs on a network this large: the edges' end flux / pressure gathered once
    q0e = torch.empty(dn.E, dtype=torch.float64, device=dn.device)
    pue = torch.empty(dn.E, dtype=torch.float64, device=dn.device)
    k("edge_flux_kernel (end flux and pressure of every edge, once per solve)", lambda: ms.lib.gnx_edge_flux_mixed(
        dn.net_ref(), bufs["cond"].data_ptr(), bufs["rsrc"].data_ptr(), bufs["P"].data_ptr(), q0e.data_ptr(), pue.data_ptr(), st()),
      8 * 5 * dn.E + 16 * dn.E)
    k("edge_const_tile_kernel<backsub> given the end fluxes (as in the one-call step)", lambda: ms.lib.gnx_edge_backsub_mixed_flux(
        dn.net_ref(), co.ref(), 0.0, 0.0, pbc_out.data_ptr(), pbc_node.data_ptr(), q0e.data_ptr(), pue.data_ptr(),
        x.data_ptr(), st()), 8 * N + 8 * NC)
    k("spmv_kernel (CSR, fp64 values, int32 indices)", lambda: ms.spmv(vals, x, y), 12 * nnz + 4 * (N + 1) + 16 * N)
    # what plain library kernels reach on the same buffers: a pure write stream is bounded well below the
    # copy figure MEASURED_PEAKS.json quotes, which is the ceiling for the write-dominated assembly / rhs kernels
    k("reference point: torch fill_ of the CSR values (pure write)", lambda: vals.fill_(1.0), 8 * nnz)
    k("reference point: torch copy_ of N doubles (read + write)", lambda: y.copy_(x), 16 * N)
    ms.assemble(co, out=vals)
    t = float(np.mean(timed(lambda: ms.assemble_rhs_solve(co, vals, b, x, dict(pbc_out=pbc_out, pbc_node=pbc_node), sync=False), 5, 2)))
    rr, bb = ms.residual(vals, x, b).cpu().tolist()
    out["assemble_plus_solve"] = {"ms": t, "MDOF_per_s": N / t / 1e3, "rel_residual": float(np.sqrt(rr / bb))}
    del vals, b, x, y, ms, dn
    torch.cuda.empty_cache()
    return out


def api_graph(pos, edges, n, comm=None):
    import graphnics_b200 as gn
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(pos.tolist()))
    G.add_edges_from(map(tuple, edges.tolist()))
    G.make_mesh(n, comm=comm)            # built once (not timed), like the reference benchmark
    res = gn.Constant(1.0)
    for e in G.edges():
        G.edges[e]["Res"] = res
    return G


def e2e_python_api(pos, edges, n, args, comm=None):
    """`MixedHydraulicNetwork(G, p_bc=Expression("x[1]")).solve()` on a FenicsGraph, then the solution
    copied to the host -- wall clock around a synchronised region (max over ranks when partitioned)."""
    import torch
    import graphnics_b200 as gn
    G = api_graph(pos, edges, n, comm)
    p_bc = gn.Expression("x[1]", degree=2)

    def step():
        model = gn.MixedHydraulicNetwork(G, p_bc=p_bc)
        qp = model.solve()
        return qp.vector().get_local()

    for _ in range(max(args.warmup, 3)):
        xh = step()
    torch.cuda.synchronize()
    if comm:
        import torch.distributed as dist
        dist.barrier()
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        xh = step()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    dt = float(np.mean(ts))
    n_local, E_local = xh.size, G._dn.E
    if comm:
        import torch.distributed as dist
        t = torch.tensor([dt, float(n_local), float(E_local)], dtype=torch.float64, device="cuda")
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        dt, n_sum, E_sum = float(tmax[0]), float(tsum[1]), float(tsum[2])
        world = dist.get_world_size()
        m = 1 << n
        N = len(edges) * (2 * m + 1) + G._dn.B
        h2d, d2h = int(8 * E_sum + 2 * 1024 * world), int(8 * n_sum)
    else:
        N, h2d, d2h = n_local, int(8 * E_local + 2 * 1024), int(8 * n_local)
    return {"value": N / dt / 1e6, "unit": "MDOF/s", "ms_per_step": dt * 1e3, "ms_per_step_stats": step_stats(np.asarray(ts) * 1e3),
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "timer": "wall clock (perf_counter) around synchronised region" + (", max over ranks" if comm else ""),
            "call": ("G.make_mesh(n, comm=True) once; per step MixedHydraulicNetwork(G, p_bc=Expression('x[1]')).solve(); "
                     "qp.vector().get_local()  (the rank-local part, like PETSc under MPI)" if comm else
                     "MixedHydraulicNetwork(G, p_bc=Expression('x[1]')).solve(); qp.vector().get_local()")}


# --------------------------------------------------------------------------------------------
# N > 1: one rank per GPU
# --------------------------------------------------------------------------------------------
def assemble_global(pm, pieces, lam_pieces, owner):
    """rank 0: the monolithic global vector from the ranks' local [q | p] blocks and multiplier blocks"""
    world, m, E = pm.world, pm.m, pm.E_glob
    B = pm.dn.B
    x = np.full(pm.N_glob, np.nan)
    epart = pm.dn.epart if pm.dn.edge_ids is not None else None
    for r in range(world):
        if epart is not None:
            e = np.nonzero(epart == r)[0]
        else:
            from graphnics_b200.device import partition_range
            e0, e1, _ = partition_range(E, r, world)
            e = np.arange(e0, e1)
        q = (e[:, None] * (m + 1) + np.arange(m + 1)[None, :]).ravel()
        p = E * (m + 1) + (e[:, None] * m + np.arange(m)[None, :]).ravel()
        loc = pieces[r]
        x[q] = loc[: len(q)]
        x[p] = loc[len(q): len(q) + len(p)]
        lam = lam_pieces[r]
        mine = np.arange(B) if owner is None else np.nonzero((owner == r) | ((owner < 0) & (r == 0)))[0]
        x[E * (2 * m + 1) + mine] = lam[mine]
    return x


def partitioned_case(torch, dist, dev, levels, n, args, flush, reps, warm):
    """one weak-scaling workload: build, time, residual; returns (record, pm, tensors)"""
    from graphnics_b200.parallel import PartitionedMixed, global_residual
    pos, edges = tree_arrays(levels)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    pm = PartitionedMixed(pos, edges, n, device=dev, exchange=os.environ.get("GNX_EXCHANGE", "auto"),
                          partition=os.environ.get("GNX_PARTITION", "subtree"))
    dn, ms = pm.dn, pm.ms
    ms.build_pattern()
    ms.schur_plan()
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    N_loc, nnz_loc = ms.N, ms.nnz
    co = ms.coeffs(torch.ones(dn.E, dtype=torch.float64, device=dev))
    vals = torch.empty(nnz_loc, dtype=torch.float64, device=dev)
    b = torch.empty(N_loc, dtype=torch.float64, device=dev)
    x = torch.zeros(N_loc, dtype=torch.float64, device=dev)
    xq, _ = ms.dof_coordinates(False)
    pbc_out = ms.end_values(xq[:, 1].contiguous())
    pbc_node = dn.pos[:, 1].contiguous()
    del xq
    ms._xq = None
    rhs_kw = dict(pbc_out=pbc_out, pbc_node=pbc_node)
    overlap = not getattr(args, "no_overlap", False)

    def step():
        ms.assemble_rhs_solve(co, vals, b, x, rhs_kw, overlap=overlap, sync=False)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    # All K steps (and the L2 flushes between them) are enqueued without a host wait in between: the ranks meet inside
    # every step (the exchange), so their queues stay aligned on the device.  A host wait per step (r1/r2k) let the
    # ranks' wake-up jitter -- tens of microseconds, more than the exchange itself -- skew the next step's start.
    host_us, evs = [], []
    for _ in range(reps):
        if flush is not None:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); step(); host_us.append((time.perf_counter() - t0) * 1e6); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ts = [e0.elapsed_time(e1) for e0, e1 in evs]
    dist.barrier()
    t = torch.tensor(ts, dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                  # per step: the slowest rank
    ts_max = t.cpu().numpy()
    rel_resid = global_residual(ms, vals, x, b)
    hp = ms.schur_plan()[3]
    rec = {"levels": levels, "n": n, "edges": pm.E_glob, "ndofs": pm.N_glob, "ms_per_step": float(ts_max.mean()),
           "ms_per_step_stats": step_stats(ts_max), "rel_residual": rel_resid, "edges_per_rank": dn.E,
           "ndofs_per_rank": N_loc, "bifurcations": dn.B,
           "replicated_bifurcations": int((hp.owner < 0).sum()) if ms.subtree else dn.B,
           "built_once_ms": t_build * 1e3, "host_enqueue_us_per_step": float(np.median(host_us)),
           "per_rank_bytes": {"survey_8d": int(ASM_B_PER_CELL * dn.num_cells + ELIM_B_PER_DOF * N_loc),
                              # what the kernels actually move: the CSR values written, the cell lengths read by the
                              # assembly, the elimination and the back-substitution, b and x written (Constant data: b
                              # is produced and later recomputed, never read)
                              "moved_by_the_kernel_chain": int(8 * nnz_loc + 3 * 8 * dn.num_cells + 2 * 8 * N_loc)}}
    return rec, pm, (co, vals, b, x, step)


def run_partitioned(args, dev):
    import torch
    import torch.distributed as dist
    from graphnics_b200.parallel import gather_to_root
    world, rank = dist.get_world_size(), dist.get_rank()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    grow = int(np.ceil(np.log2(world)))
    hbm, peak_src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clocks = ClockSampler(local).start() if rank == 0 else None

    # ---- headline: C2 per rank (weak scaling)
    levels, n = TREE_LEVELS + grow, TREE_N
    rec, pm, (co, vals, b, x, step) = partitioned_case(torch, dist, dev, levels, n, args, flush, args.steps, max(args.warmup, 3))
    ms, dn = pm.ms, pm.dn
    launches, launch_names = count_launches(step, args.steps)
    resident = bool(int(ms.lib.gnx_step_is_resident()))
    # parity: rank 0 gathers the solution and compares it with the oracle's elimination route on the WHOLE tree
    pieces = gather_to_root(x[:dn.lam_off])
    lam_pieces = gather_to_root(x[dn.lam_off:])
    parity = None
    if rank == 0 and not args.no_cpu:
        import oracle
        pos, edges = tree_arrays(levels)
        om = oracle.build_mesh(pos, edges, n)
        b_ref = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda xx: xx[..., 1], 1), project=False)
        x_ref = oracle.solve_by_elimination(om, b_ref, 1.0)
        owner = ms.schur_plan()[3].owner if ms.subtree else None
        xg = assemble_global(pm, pieces, lam_pieces, owner)
        parity = parity_block(om, x_ref, "oracle elimination route (oracle/elimination.py) on the whole N-GPU tree; "
                                         "the ranks' pieces gathered on rank 0", xg)
        del om, x_ref, xg
    del pieces, lam_pieces
    dist.barrier()

    # ---- e2e through the drop-in API on the partitioned graph
    pos, edges = tree_arrays(levels)
    e2e = e2e_python_api(pos, edges, n, args, comm=True)
    del co, vals, b, x, step, pm, ms, dn
    torch.cuda.empty_cache()

    # ---- at scale: 16 + log2(world) generations, 2^7 cells per edge (world = 8: BASELINE.json configs[4], 135 M DOFs)
    srec, spm, (sco, svals, sb, sx, sstep) = partitioned_case(torch, dist, dev, SCALE_LEVELS + grow, SCALE_N, args, None, 10, 3)
    sms, sdn = spm.ms, spm.dn
    srec["MDOF_per_s"] = srec["ndofs"] / srec["ms_per_step"] / 1e3
    srec["per_rank_GBps_survey_bytes"] = srec["per_rank_bytes"]["survey_8d"] / srec["ms_per_step"] / 1e6
    srec["per_rank_frac_of_measured_hbm_peak"] = srec["per_rank_GBps_survey_bytes"] / hbm
    srec["per_rank_GBps_moved_bytes"] = srec["per_rank_bytes"]["moved_by_the_kernel_chain"] / srec["ms_per_step"] / 1e6
    srec["per_rank_frac_by_moved_bytes"] = srec["per_rank_GBps_moved_bytes"] / hbm
    srec["note"] = ("per_rank_frac_of_measured_hbm_peak counts SURVEY.md 8(d)'s algorithmic bytes (110 B/cell + 32 B/DOF) and can "
                    "exceed 1: the kernels move fewer bytes than that model (b is produced, never read); "
                    "per_rank_frac_by_moved_bytes counts what they do move")
    srec["workload"] = (f"binary tree, {SCALE_LEVELS + grow} generations, E={srec['edges']}, 2^{SCALE_N} cells/edge, "
                        f"N={srec['ndofs']} across {world} GPUs (L2 not flushed: every array is >> L2)")
    # oracle check that does not depend on the refinement: with f = g = 0 the multipliers and the (constant) flux of
    # every edge are the graph-level Kirchhoff solution, whatever the number of cells -- compare with the oracle
    # on the same tree at 2 cells per edge
    lam_all = gather_to_root(sx[sdn.lam_off:])
    q0_all = gather_to_root(sx[:sdn.p_off].reshape(sdn.E, sdn.m + 1)[:, 0].contiguous())
    if rank == 0 and not args.no_cpu:
        import oracle
        pos2, edges2 = tree_arrays(SCALE_LEVELS + grow)
        om = oracle.build_mesh(pos2, edges2, 1)
        b_ref = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda xx: xx[..., 1], 1), project=False)
        x_ref = oracle.solve_by_elimination(om, b_ref, 1.0)
        lay = oracle.mixed_layout(om)
        owner = sms.schur_plan()[3].owner if sms.subtree else None
        lam = np.full(sdn.B, np.nan)
        qe = np.full(spm.E_glob, np.nan)
        for r in range(world):
            mine = np.arange(sdn.B) if owner is None else np.nonzero((owner == r) | ((owner < 0) & (r == 0)))[0]
            lam[mine] = lam_all[r][mine]
            e = np.nonzero(sdn.epart == r)[0] if sdn.edge_ids is not None else None
            if e is None:
                from graphnics_b200.device import partition_range
                e0, e1, _ = partition_range(spm.E_glob, r, world)
                e = np.arange(e0, e1)
            qe[e] = q0_all[r]
        srec["oracle_sample"] = {"what": "multipliers and per-edge flux vs the oracle on the same tree at 2 cells/edge "
                                         "(refinement-independent for f = g = 0)",
                                 "rel_l2_lam": rel(lam, x_ref[lay["lam_off"]:]),
                                 "rel_l2_q_edge": rel(qe, x_ref[:lay["p_off"]].reshape(len(edges2), 3)[:, 0])}
        srec["oracle_sample"]["ok"] = bool(max(srec["oracle_sample"]["rel_l2_lam"], srec["oracle_sample"]["rel_l2_q_edge"]) <= 1e-8)
    dist.barrier()
    clk = clocks.stop() if clocks is not None else None
    if rank == 0:
        N = rec["ndofs"]
        ms_per_step = rec["ms_per_step"]
        out = {"metric": METRIC, "value": N / (ms_per_step * 1e-3) / 1e6, "unit": "MDOF/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config_dict(levels, n, rec["edges"], N, {
                   "parallelism": (f"edge-partitioned x{world}: one subtree per rank (partition.py), rank-local elimination and "
                                   f"rake, {rec['replicated_bifurcations']} replicated bifurcations, ONE exchange per solve "
                                   f"(edges at replicated bifurcations + local subtree roots)"),
                   "exchange": "peer stores over NVLink + flag words inside the kernel" if sms.exchange_mode == "p2p"
                               else f"NCCL all-gather between two launches ({getattr(sms, '_p2p_error', '')})",
                   "edges_per_rank": rec["edges_per_rank"], "ndofs_per_rank": rec["ndofs_per_rank"],
                   "bifurcations": rec["bifurcations"],
                   "timing": "CUDA events around every step on every rank, per step the max over ranks; the K steps and the "
                             "L2 flushes between them are enqueued back to back (no host wait between steps)"}),
               "ms_per_step_stats": rec["ms_per_step_stats"], "built_once_ms": {"total_ms": rec["built_once_ms"]},
               "rel_residual": rec["rel_residual"], "gpu_launches": launches, "gpu_launches_per_step": launch_names,
               "gpu_launches_how": "CUPTI (torch.profiler) kernel records of a replay of the same K steps on rank 0",
               "host_enqueue_us_per_step": rec["host_enqueue_us_per_step"],
               "step_path": "ONE persistent cooperative kernel per rank, the exchange inside it" if resident else "kernel chain",
               "parity": parity,
               "roofline": {"bound": "hbm", "kernel": "per-rank step (assemble + rhs + eliminate + Schur + back-substitute)",
                            "achieved": rec["per_rank_bytes"]["survey_8d"] / ms_per_step / 1e6, "peak": hbm, "unit": "GB/s",
                            "frac": rec["per_rank_bytes"]["survey_8d"] / ms_per_step / 1e6 / hbm, "traffic": None,
                            "peak_source": peak_src,
                            "algorithmic_bytes_rule": f"SURVEY.md 8(d): {ASM_B_PER_CELL} B/cell + {ELIM_B_PER_DOF} B/DOF, per rank"},
               "at_scale": srec, "e2e": e2e, "clocks": clk}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / oracle parity legs (profiling runs)")
    ap.add_argument("--no-overlap", action="store_true",
                    help="kernel chain only: run the step's kernels back to back on one stream")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    main()
