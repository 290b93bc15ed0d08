"""Persistent (resident-tile) step kernel on the C2 tree: CUDA-event time per step with the L2 flushed, the
phase stamps of CTA 0 (%globaltimer: start | own tiles eliminated | all CTAs arrived | Schur done | parked CTAs
released | end), and the same step through the three-kernel chain (GNX_STEP_RESIDENT=0 in a second process).
  python tools/time_resident.py [levels] [n] [--no-vals]"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from graphnics_b200.device import DeviceNetwork
from graphnics_b200.engine import MixedSystem
from graphnics_b200.synthetic import murray_tree

args = [a for a in sys.argv[1:] if not a.startswith("--")]
levels = int(args[0]) if args else 11
n = int(args[1]) if len(args) > 1 else 9
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:                      # torchrun: `levels` is the single-GPU tree, one more generation per doubling
    import torch.distributed as dist
    from graphnics_b200.parallel import PartitionedMixed
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    levels += int(np.ceil(np.log2(world)))
pos, edges, _, _ = murray_tree(levels, radius0=0.1, gam=0.9, L0=10.0, gdim=2)
if world > 1:
    pm = PartitionedMixed(pos, edges, n, partition=os.environ.get("GNX_PARTITION", "subtree"))
    dn, ms = pm.dn, pm.ms
else:
    dn = DeviceNetwork(pos, edges, n)
    ms = MixedSystem(dn)
ms.build_pattern()
co = ms.coeffs(np.ones(dn.E))
vals = None if "--no-vals" in sys.argv else torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
b = torch.empty(ms.N, dtype=torch.float64, device="cuda")
x = torch.zeros(ms.N, dtype=torch.float64, device="cuda")
xq, _ = ms.dof_coordinates(False)
kw = dict(pbc_out=ms.end_values(xq[:, 1].contiguous()), pbc_node=dn.pos[:, 1].contiguous())
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def step():
    if vals is None:
        ms._step_native(co, None, b, x, kw, True, sync=False)
    else:
        ms.assemble_rhs_solve(co, vals, b, x, kw, sync=False)


for _ in range(5):
    step()
torch.cuda.synchronize()
ts, stamps, stamps1, stamps2, stamps3 = [], [], [], [], []
for _ in range(30):
    if "--read-flush" in sys.argv:                  # L2 full of CLEAN foreign lines (a write flush leaves dirty ones)
        flush_sink = flush.view(torch.int64).sum()
    elif "--no-flush" not in sys.argv:
        flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); e1.record(); e1.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
    s = ms._step_sync.cpu().numpy()
    stamps.append((s[[8, 3, 4, 5, 6, 7]] - s[2]) / 1e3)
    stamps1.append((s[16:31] - s[2]) / 1e3)
    if ms.dn.B > 0:
        w = ms.schur_plan()[2]
        sc = w[-32 - ((ms.schur_plan()[0].na ** 2 + ms.schur_plan()[0].na + 2 * 1024) if ms.schur_plan()[0].na else 0):][:32].cpu().numpy()
        stamps2.append((np.concatenate([sc[16:18], sc[9:16], sc[18:24], sc[25:26]]) - sc[8]) / 1e3)
out = {"levels": levels, "n": n, "N": ms.N, "resident": os.environ.get("GNX_STEP_RESIDENT", "1"), "vals": vals is not None,
       "us_per_step": {"min": float(np.min(ts)), "median": float(np.median(ts)), "max": float(np.max(ts))}}
if os.environ.get("GNX_STEP_RESIDENT", "1") != "0":
    st = np.median(np.asarray(stamps), axis=0)
    out["cta0_stamps_us"] = dict(zip(["prefetched", "own_tiles_done", "all_arrived", "schur_done", "released", "end"], st.round(2).tolist()))
    st1 = np.median(np.asarray(stamps1), axis=0)
    out["cta1_stamps_us"] = dict(zip(["prefetched", "H0_arrived", "elim0_done", "H4_arrived", "elim4_done", "b_stores_read", "asm_start",
                                      "asm_done", "released", "p3_prefetched", "bs0_done", "bs4_done", "end", "phase1_barrier", "b_tiles_done"], st1.round(2).tolist()))
if stamps2:
    out["schur_stamps_us"] = dict(zip(["top_solve_entry", "rows_ready_thread0", "prologue", "local_up", "exported", "flags", "repl_rows", "up_done", "down_done",
                                       "x_roots_sent", "x_flag_released", "x_flag_seen", "first_repl_level_folded", "repl_row_begin", "repl_row_end", "workers_arrived"],
                                      np.median(np.asarray(stamps2), axis=0).round(2).tolist()))
if vals is not None and world == 1:
    rr, bb = ms.residual(vals, x, b).cpu().tolist()
    out["rel_residual"] = float(np.sqrt(rr / bb))
if world > 1:
    out["rank"] = dist.get_rank()
    for r in range(world):
        if r == dist.get_rank():
            print(json.dumps(out), flush=True)
        dist.barrier()
    dist.destroy_process_group()
else:
    print(json.dumps(out))
