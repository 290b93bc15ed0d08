"""SpMV variants (GNX_SPMV_VARIANT) on a large tree matrix: GB/s each.  Tuning aid."""
import os, sys, subprocess, json
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np, torch
    from graphnics_b200.device import DeviceNetwork
    from graphnics_b200.engine import MixedSystem
    from graphnics_b200.synthetic import murray_tree
    pos, edges, _, _ = murray_tree(int(sys.argv[2]), gdim=2)
    dn = DeviceNetwork(pos, edges, int(sys.argv[3])); ms = MixedSystem(dn); ms.build_pattern()
    vals = ms.assemble(ms.coeffs(1.0))
    x = torch.randn(ms.N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    for _ in range(3): ms.spmv(vals, x, y)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ms.spmv(vals, x, y); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    byt = 12 * ms.nnz + 4 * (ms.N + 1) + 16 * ms.N
    import scipy.sparse as sp
    print(json.dumps({"variant": os.environ.get("GNX_SPMV_VARIANT", "0"), "N": ms.N, "nnz": ms.nnz, "ms": min(ts), "GBps": byt / min(ts) / 1e6}))
else:
    for v in ("0", "1", "2", "3", "4", "5"):
        env = dict(os.environ, GNX_SPMV_VARIANT=v)
        for lv, n in ((19, 7), (11, 9)):
            print(subprocess.run([sys.executable, __file__, "child", str(lv), str(n)], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1])
