"""Run under torchrun (one rank per GPU, NCCL): the edge-partitioned mixed solve of a synthetic tree
and of a honeycomb against the oracle's monolithic LU solution.  Prints one OK line per case on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/check_partitioned.py
"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import oracle
from graphnics_b200.parallel import PartitionedMixed, global_residual
from graphnics_b200.synthetic import murray_tree


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    cases = []
    pos, edges, _, _ = murray_tree(10, gdim=2)
    cases.append(("tree10", pos, edges, 4, "subtree"))
    cases.append(("tree10", pos, edges, 7, "subtree"))       # n >= 6: the persistent step kernel runs the exchange
    pos, edges, _, _ = murray_tree(14, gdim=2)
    cases.append(("tree14", pos, edges, 1, "subtree"))       # blocked local rake + replicated top
    cases.append(("tree14", pos, edges, 1, "range"))         # round-1 layout: everything gathered, Schur solve replicated
    from graphnics_b200.generators import honeycomb
    G = honeycomb(12, 12, mesh=False)
    pos, edges = G._graph_arrays()
    cases.append(("honeycomb12", pos, edges, 3, "subtree"))  # no rakeable node: every bifurcation replicated
    ok = True
    for name, pos, edges, n, how in cases:
        pm = PartitionedMixed(pos, edges, n, exchange=os.environ.get("GNX_EXCHANGE", "auto"), partition=how)
        dn, ms = pm.dn, pm.ms
        rng = np.random.default_rng(1)
        Res = rng.uniform(0.5, 2.0, size=len(edges))
        co = ms.coeffs(Res[pm.edge_ids()])
        xq, _ = ms.dof_coordinates(False)
        pbc = lambda x: 1.0 + 0.3 * x[..., 0] - 0.2 * x[..., 1]
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
        b = ms.rhs(None, None, ms.end_values(dev(pbc(xq.cpu().numpy()))), dev(pbc(pos)), f_const=0.3)
        vals = ms.assemble(co)
        x = torch.zeros(ms.N, dtype=torch.float64, device="cuda")
        for _ in range(4):                      # repeated solves alternate the two exchange buffers
            x, info = ms.solve(co, b, x=x)
        # ... and the one-call step (Constant data: fused right-hand side; persistent kernel when it applies)
        x2 = torch.zeros(ms.N, dtype=torch.float64, device="cuda")
        b2 = torch.empty_like(b)
        vals2 = torch.empty_like(vals)
        for _ in range(3):
            ms.assemble_rhs_solve(co, vals2, b2, x2, dict(pbc_out=ms.end_values(dev(pbc(xq.cpu().numpy()))), pbc_node=dev(pbc(pos)),
                                                           f_const=0.3), sync=False)
        torch.cuda.synchronize()
        same = bool(torch.equal(x2[:dn.lam_off], x[:dn.lam_off]) and torch.equal(b2, b) and torch.equal(vals2, vals))
        resid = global_residual(ms, vals, x, b)
        om = oracle.build_mesh(pos, edges, n)
        A = oracle.assemble_mixed(om, Res)
        bo = oracle.assemble_mixed_rhs(om, f=0.3, p_bc=oracle.Coef(pbc, 1), project=False)
        xo = oracle.lu_solve(A, bo)
        idx = pm.local_dofs_of_global()
        keep = np.concatenate([np.arange(dn.lam_off), dn.lam_off + pm.valid_multipliers()])
        xh = x.cpu().numpy()
        err = np.linalg.norm(xh[keep] - xo[idx][keep]) / np.linalg.norm(xo[idx][keep])
        t = torch.tensor([err, 0.0 if same else 1.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        good = float(t[0]) < 1e-8 and resid < 1e-10 and float(t[1]) == 0.0
        ok = ok and good
        if rank == 0:
            plan = ms.schur_plan()[3]
            print(f"{'OK' if good else 'FAIL'} {name} n={n} {how}: world={world} N={pm.N_glob} max rel err vs oracle LU "
                  f"{float(t[0]):.2e}, global residual {resid:.2e}, step == solve bitwise: {float(t[1]) == 0.0}, "
                  f"replicated bifurcations {int((plan.owner < 0).sum()) if ms.subtree else dn.B} of {dn.B}, "
                  f"cg iterations {info.cg_iterations}, exchange {ms.exchange_mode} {getattr(ms, '_p2p_error', '')}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
