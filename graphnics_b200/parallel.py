"""Edge-partitioned multi-GPU execution of the mixed model: one process per GPU,
torch.distributed (NCCL) for the plumbing.

A rank owns a contiguous range of edges and with it, exclusively, their q and p unknowns; the
multipliers lambda are replicated.  Assembly, right-hand side, per-edge elimination and
back-substitution run on the local edges only (the rank's SUBDOMAIN matrix / vector: q and p rows
are complete, lambda rows hold the local incidences, the global rows are their sum over ranks).
The one exchange of a solve is an all-gather of the three edge scalars (conductance, source, total
divergence: 24 bytes per edge of the whole network) -- peer stores over NVLink from inside the
elimination kernel, synchronised by flag words inside the Schur kernel (engine.MixedSystem._p2p_setup;
NCCL all_gather_into_tensor as the fallback); the graph-level Schur solve is then replicated
-- identical inputs and a deterministic kernel give bit-identical multipliers on every rank, so no
Krylov dot product or halo ever crosses NVLink.
"""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .device import DeviceNetwork, partition_range
from .engine import MixedSystem

__all__ = ["PartitionedMixed", "run_partitioned_bench", "global_residual"]


class PartitionedMixed:
    """the mixed system of the whole network, this rank's share.

    partition="subtree" (default): edges are dealt by partition.partition_edges -- whole subtrees of the rake forest
    per rank, so a rank eliminates and rakes its part without communication and only the few replicated bifurcations
    at the top of the tree (plus the edges at them) take part in the one exchange of a solve.
    partition="range": contiguous equal index ranges, every edge's scalars gathered, Schur solve replicated (round 1)."""

    def __init__(self, pos, edges, n, rank=None, world=None, device=None, group=None, exchange="auto",
                 partition="subtree"):
        self.rank = dist.get_rank() if rank is None else rank
        self.world = dist.get_world_size() if world is None else world
        epart = plan = None
        if partition == "subtree" and self.world > 1:
            from .partition import partition_edges
            epart, plan = partition_edges(np.asarray(edges), len(pos), self.world)
        self.dn = DeviceNetwork(pos, edges, n, device=device, part=(self.rank, self.world), epart=epart)
        if plan is not None:
            self.dn.host_plan = plan
        self.ms = MixedSystem(self.dn, group=group, exchange=exchange)
        if epart is not None:
            try:
                self.ms.schur_plan()
            except ValueError:                       # replicated set too large for the top part: round-1 layout
                self.dn = DeviceNetwork(pos, edges, n, device=device, part=(self.rank, self.world))
                self.ms = MixedSystem(self.dn, group=group, exchange=exchange)
        self.E_glob, self.m = self.dn.E_glob, self.dn.m
        self.N_glob = self.E_glob * (2 * self.m + 1) + self.dn.B

    def edge_ids(self):
        """global ids of this rank's edges, ascending"""
        dn = self.dn
        return dn.edge_ids if dn.edge_ids is not None else np.arange(dn.e0, dn.e1, dtype=np.int64)

    def local_dofs_of_global(self):
        """indices into the monolithic GLOBAL vector of this rank's local vector [q_loc | p_loc | lambda]"""
        m = self.m
        e = self.edge_ids()
        q = (e[:, None] * (m + 1) + np.arange(m + 1)[None, :]).ravel()
        p = self.E_glob * (m + 1) + (e[:, None] * m + np.arange(m)[None, :]).ravel()
        lam = self.E_glob * (2 * m + 1) + np.arange(self.dn.B)
        return np.concatenate([q, p, lam])

    def valid_multipliers(self):
        """bifurcations whose multiplier this rank computes: its local nodes and the replicated ones (all of them with
        the range partition)"""
        if not self.ms.subtree or self.dn.B == 0:
            return np.arange(self.dn.B)
        v = self.ms.schur_plan()[3]
        return np.nonzero((v.owner == self.rank) | (v.owner < 0))[0]


def global_residual(ms, vals, x, b, group=None):
    """||b - A x|| / ||b|| of the WHOLE system from the subdomain pieces: q/p rows are local, the
    lambda rows of A x are summed over ranks (b's lambda block is replicated)."""
    dn = ms.dn
    r = torch.empty_like(x)
    ms.residual(vals, x, b, r)
    lam_off = dn.lam_off
    world = dist.get_world_size(group)
    # local lambda residual = b_lam - (A_r x_r)_lam ; sum over ranks = world*b_lam - (A x)_lam
    rl = r[lam_off:].clone()
    dist.all_reduce(rl, group=group)
    rl -= (world - 1) * b[lam_off:]
    acc = torch.stack([(r[:lam_off] ** 2).sum(), (b[:lam_off] ** 2).sum()])
    dist.all_reduce(acc, group=group)
    rr = float(acc[0]) + float((rl ** 2).sum())
    bb = float(acc[1]) + float((b[lam_off:] ** 2).sum())
    return (rr / bb) ** 0.5 if bb > 0 else rr ** 0.5


def run_partitioned_bench(args, metric, config_dict, peaks, ClockSampler, tree_levels=11, n=9):
    """bench.py --gpus N (N > 1): weak scaling -- the tree grows by one generation per doubling of
    the GPU count, every rank keeps ~2047 edges x 2^9 cells."""
    from .synthetic import murray_tree
    world, rank = dist.get_world_size(), dist.get_rank()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    levels = tree_levels + int(np.ceil(np.log2(world)))
    pos, edges, _, _ = murray_tree(levels, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
    pm = PartitionedMixed(pos, edges, n, device=dev, exchange=os.environ.get("GNX_EXCHANGE", "auto"))
    dn, ms = pm.dn, pm.ms
    ms.build_pattern()
    ms.schur_plan()
    N_loc, nnz_loc = ms.N, ms.nnz
    res_host = torch.ones(dn.E, dtype=torch.float64).pin_memory()
    co = ms.coeffs(res_host.to(dev))
    vals = torch.empty(nnz_loc, dtype=torch.float64, device=dev)
    b = torch.empty(N_loc, dtype=torch.float64, device=dev)
    x = torch.empty(N_loc, dtype=torch.float64, device=dev)
    xq, _ = ms.dof_coordinates(False)
    pbc_out = ms.end_values(xq[:, 1].contiguous())
    pbc_node = dn.pos[:, 1].contiguous()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    rhs_kw = dict(pbc_out=pbc_out, pbc_node=pbc_node)
    overlap = not getattr(args, "no_overlap", False)

    def step():
        ms.assemble_rhs_solve(co, vals, b, x, rhs_kw, overlap=overlap, sync=False)

    clocks = ClockSampler(local).start() if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    tot = 0.0
    host_us = []
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); t0 = time.perf_counter(); step(); host_us.append((time.perf_counter() - t0) * 1e6); e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    torch.cuda.synchronize()
    dist.barrier()
    t = torch.tensor([tot], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0]) / args.steps
    rel_resid = global_residual(ms, vals, x, b)

    # e2e: host coefficient data in, host solution out, wall clock, max over ranks
    xh = torch.empty(N_loc, dtype=torch.float64).pin_memory()

    def e2e_step():
        co2 = ms.coeffs(res_host.to(dev, non_blocking=True))
        ms.assemble_rhs_solve(co2, vals, b, x, rhs_kw, overlap=overlap, sync=False)
        xh.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
    for _ in range(3):
        e2e_step()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    dt = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    clk = clocks.stop() if clocks is not None else None
    hp = ms.schur_plan()[3]
    if rank == 0:
        hbm, peak_src = peaks()
        N = pm.N_glob
        out = {"metric": metric, "value": N / (ms_per_step * 1e-3) / 1e6, "unit": "MDOF/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": config_dict(levels, n, pm.E_glob, N, {
                   "parallelism": f"edge-partitioned x{world} (contiguous equal edge ranges), replicated Schur solve, "
                                  f"one exchange of 24 B/edge per solve",
                   "exchange": (("peer stores over NVLink fused into the elimination kernel, flag words signalled and "
                                 "awaited inside the Schur kernel (no collective, no barrier kernel)"
                                 if ms._p2p and ms._p2p.get("flags") else
                                 "peer stores over NVLink fused into the elimination kernel + one cross-GPU barrier kernel")
                                if ms.exchange_mode == "p2p" else
                                f"NCCL all_gather_into_tensor ({getattr(ms, '_p2p_error', '')})"),
                   "edges_per_rank": dn.E, "ndofs_per_rank": N_loc, "bifurcations": dn.B}),
               "rel_residual": rel_resid, "gpu_launches": 4 * args.steps,
               "host_enqueue_us_per_step": float(np.median(host_us)),
               "schur": {"rake_levels": len(hp.levels), "core": hp.nc},
               "roofline": {"bound": "hbm", "kernel": "per-rank step (assemble + rhs + eliminate + back-substitute)",
                            "achieved": (8 * (nnz_loc + dn.num_cells) + 8 * N_loc + 2 * 8 * (N_loc + dn.num_cells)
                                         + 8 * N_loc) / ms_per_step / 1e6,
                            "peak": hbm, "unit": "GB/s", "frac": None, "traffic": None, "peak_source": peak_src},
               "e2e": {"value": N / float(dt[0]) / 1e6, "unit": "MDOF/s", "ms_per_step": float(dt[0]) * 1e3,
                       "h2d_bytes_per_step": int(8 * dn.E) * world, "d2h_bytes_per_step": int(8 * N_loc) * world,
                       "timer": "wall clock, max over ranks",
                       "call": "per rank: pinned Res -> device, assemble + rhs + solve, local solution -> pinned host"},
               "clocks": clk}
        out["roofline"]["frac"] = out["roofline"]["achieved"] / hbm
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
