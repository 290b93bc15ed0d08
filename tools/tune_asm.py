"""Assembly tile kernel with CPT = 1, 2, 4 cells per thread (GNX_ASM_CPT): GB/s at C2 size and at scale."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(int(sys.argv[2]), gdim=2)
    dn = DeviceNetwork(pos, edges, int(sys.argv[3])); ms = MixedSystem(dn); ms.build_pattern()
    co = ms.coeffs(1.0); vals = torch.empty(ms.nnz, dtype=torch.float64, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): ms.assemble(co, out=vals)
    torch.cuda.synchronize(); ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ms.assemble(co, out=vals); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    byt = 8 * ms.nnz + 8 * dn.num_cells
    print(json.dumps({"cpt": os.environ.get("GNX_ASM_CPT"), "N": ms.N, "ms": float(np.median(ts)), "GBps": byt / float(np.median(ts)) / 1e6}))
else:
    for v in ("4",):
        for lv, n in ((11, 9), (17, 7)):
            print(subprocess.run([sys.executable, __file__, "child", str(lv), str(n)], env=dict(os.environ, GNX_ASM_CPT=v), capture_output=True, text=True).stdout.strip().splitlines()[-1])
