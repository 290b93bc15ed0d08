"""Edge-partitioned multi-GPU execution of the mixed model: one process per GPU,
torch.distributed (NCCL) for the plumbing.

A rank owns a set of edges (partition.py: whole subtrees of the rake forest) and with it, exclusively, their q and
p unknowns.  Assembly, right-hand side, per-edge elimination and back-substitution run on the local edges only
(the rank's SUBDOMAIN matrix / vector: q and p rows are complete, lambda rows hold the local incidences, the
global rows are their sum over ranks).  The graph-level Schur system is raked rank-locally wherever a whole
subtree sits on one rank; the ONE exchange of a solve carries the three scalars of the edges at REPLICATED
bifurcations and the final (diagonal, right-hand side) pair of every local subtree root -- peer stores over NVLink
from inside the kernels, synchronised by flag words inside the Schur kernel (engine.MixedSystem._xch_setup; one
NCCL all-gather between two kernel launches as the fallback).  The replicated bifurcations (2^k - 1 for a binary
tree on 2^k ranks; everything for a network without tree-like parts) are processed redundantly: identical inputs
and a deterministic kernel give bit-identical multipliers on every rank, so no Krylov dot product or halo crosses
NVLink.  partition="range" keeps round 1's layout (contiguous index ranges, all edge scalars gathered, whole Schur
solve replicated).
"""
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from .device import DeviceNetwork, partition_range
from .engine import MixedSystem

__all__ = ["PartitionedMixed", "global_residual", "gather_to_root"]


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


def gather_to_root(t, group=None):
    """the (differently sized) 1-D device tensors of all ranks on rank 0, as a list of host arrays (None elsewhere)"""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s_[0]) for s_ in sizes]
    pad = torch.zeros(max(sizes), dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, out, dst=0, group=group)
    if rank != 0:
        return None
    return [o[:k].cpu().numpy() for o, k in zip(out, sizes)]
