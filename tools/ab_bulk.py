"""A/B of the copy-engine (cp.async.bulk) tile I/O against the block's own load/store loops:
GNX_ASM_BULK / GNX_EDGE_BULK = 0 | 1, C2 size and at scale; per-kernel and whole-step times (CUDA events,
L2 flushed), plus a digest of the outputs (must be identical across variants)."""
import os, sys, subprocess, json, hashlib
KEYS = ("GNX_ASM_BULK", "GNX_EDGE_BULK", "GNX_PDL")
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    from graphnics_b200.synthetic import murray_tree
    lv, n = int(sys.argv[2]), int(sys.argv[3])
    pos, edges, _, _ = murray_tree(lv, gdim=2)
    dn = DeviceNetwork(pos, edges, n); ms = MixedSystem(dn); ms.build_pattern(); ms.schur_plan()
    co = ms.coeffs(torch.linspace(0.5, 2.0, dn.E, dtype=torch.float64, device="cuda"))
    vals = torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
    b = torch.empty(ms.N, dtype=torch.float64, device="cuda"); x = torch.empty_like(b)
    xq, _ = ms.dof_coordinates(False)
    kw = dict(pbc_out=ms.end_values(xq[:, 1].contiguous()), pbc_node=dn.pos[:, 1].contiguous())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed(fn, reps=12):
        for _ in range(3): fn()
        torch.cuda.synchronize(); ts = []
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
        return float(np.median(ts)) * 1e3

    out = {",".join(KEYS): ",".join(os.environ.get(k, "-") for k in KEYS), "N": ms.N}
    out["asm_us"] = timed(lambda: ms.assemble(co, out=vals))
    out["step_us"] = timed(lambda: ms.assemble_rhs_solve(co, vals, b, x, kw, overlap=True, sync=False))
    out["step_serial_us"] = timed(lambda: ms.assemble_rhs_solve(co, vals, b, x, kw, overlap=False, sync=False))
    out["solve_us"] = timed(lambda: ms.solve(co, b, x, const_rhs=kw, sync=False))
    b2 = b.clone()
    out["solve_given_b_us"] = timed(lambda: ms.solve(co, b2, x, sync=False))
    torch.cuda.synchronize()
    r = torch.empty_like(x); ms.residual(vals, x, b, r)
    out["rel_resid"] = float(r.norm() / b.norm())
    out["digest"] = hashlib.sha1(vals.cpu().numpy().tobytes() + b.cpu().numpy().tobytes() + x.cpu().numpy().tobytes()).hexdigest()[:16]
    print(json.dumps(out))
else:
    for lv, n in ((11, 9), (14, 6), (17, 7)):
        for vals in (("1", "1", "0"), ("1", "1", "1")):
            r = subprocess.run([sys.executable, __file__, "child", str(lv), str(n)],
                               env=dict(os.environ, **dict(zip(KEYS, vals))), capture_output=True, text=True)
            print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "FAILED " + r.stderr[-800:], flush=True)
