"""FE-level pinning against dumps of the REAL reference (tools/dump_reference.py -> tests/golden/
reference_dump_*.npz, produced in the reference's FEniCS image).  While those files are absent the
real-dump cases skip; the comparator itself (tests/refdump.py) is always exercised on a synthetic dump: the
oracle's C1 system under a random per-block DOF permutation with extra explicit zeros, which it must accept,
and the same dump with one matrix entry off by 1e-10 / one swapped DOF, which it must reject.
"""
import glob
import os

import numpy as np
import pytest

import oracle
import refdump
from _util import y_graph

HERE = os.path.dirname(os.path.abspath(__file__))
REAL = sorted(glob.glob(os.path.join(HERE, "golden", "reference_dump_*.npz")))


@pytest.mark.parametrize("kind,n", [("mixed", 4), ("mixed", 10), ("primal", 6)])
def test_comparator_accepts_permuted_oracle_dump(kind, n, tmp_path):
    pos, edges = y_graph()
    d = refdump.synthetic_dump(kind, pos, edges, n, seed=n)
    path = tmp_path / "reference_dump_synth.npz"
    np.savez_compressed(path, **d)                       # through the file format the tool writes
    dump = dict(np.load(path))
    rep = refdump.compare(dump, refdump.oracle_side(dump))
    assert rep["mesh"] == "bit-exact" and not rep["identity_permutation"]
    assert rep["reference_explicit_zeros"] > 0 and rep["our_entries_not_stored_by_reference"] == 0
    assert max(v for k, v in rep.items() if k.startswith("rel_l2_")) < 1e-12


def test_comparator_rejects_wrong_entries_and_wrong_numbering():
    pos, edges = y_graph()
    bad = refdump.synthetic_dump("mixed", pos, edges, 4, seed=1, perturb=(17, 1e-10))
    with pytest.raises(AssertionError):
        refdump.compare(bad, refdump.oracle_side(bad))
    bad = refdump.synthetic_dump("mixed", pos, edges, 4, seed=2)
    x = bad["x"].copy()
    k = int(bad["block_sizes"][:3].sum())                # first pressure dof (p varies along an edge; q does not)
    x[[k + 3, k + 4]] = x[[k + 4, k + 3]]                # two pressure dofs swapped in the solution
    bad["x"] = x
    with pytest.raises(AssertionError):
        refdump.compare(bad, refdump.oracle_side(bad))
    bad = refdump.synthetic_dump("mixed", pos, edges, 4, seed=3)
    c = bad["coordinates"].copy()
    c[7, 1] = np.nextafter(c[7, 1], 1.0)                 # one ulp in one mesh coordinate
    bad["coordinates"] = c
    with pytest.raises(AssertionError):
        refdump.compare(bad, refdump.oracle_side(bad))


@pytest.mark.skipif(not REAL, reason="no tests/golden/reference_dump_*.npz (run tools/dump_reference.py in the "
                                     "reference's FEniCS image: ingeborggjerde/graphnics, docker/Dockerfile)")
@pytest.mark.parametrize("path", REAL or ["-"])
def test_oracle_against_real_reference_dump(path):
    dump = dict(np.load(path))
    rep = refdump.compare(dump, refdump.oracle_side(dump))
    print(os.path.basename(path), rep)


# ---------------------------------------------------------------------------------------------
# the CUDA path against the same dumps (synthetic always, real ones when present)
# ---------------------------------------------------------------------------------------------
class _DeviceMesh:
    """the mesh-level arrays of a DeviceNetwork in the shape the comparator reads (host copies)"""

    def __init__(self, dn, om_ref):
        self.coords = dn.coords.cpu().numpy()
        self.cells = dn.cells.cpu().numpy().astype(np.int64)
        self.mf = dn.mf.cpu().numpy().astype(np.int64)
        self.bif_ixs, self.bnd_ixs = dn.bifurcation_ixs(), dn.boundary_ixs()
        self.tangents = dn.tangents.cpu().numpy()
        self.parent_vertex = dn.parent_vertex_host()
        self.vf = dn.vf_host()
        self.E, self.m, self.B, self.V = dn.E, dn.m, dn.B, dn.V
        self.pos, self.edges = dn.pos_host(), dn.edges_host()
        self.sub_cells = np.broadcast_to(dn.sub_cells_host(), (dn.E, dn.m, 2))
        self.sub_coords = self.coords[self.parent_vertex]
        # SubMesh cell j is parent cell e*m + j (cells of an edge are contiguous, A5)
        self.parent_cell = np.arange(dn.E * dn.m, dtype=np.int64).reshape(dn.E, dn.m)
        self.n = dn.n


def gpu_side(dump):
    import graphnics_b200 as gn
    kind, case = str(dump["kind"]), str(dump["case"])
    pos, edges, n = dump["pos"], dump["edges"], int(dump["n"])
    G = gn.FenicsGraph()
    G.add_nodes_from((i, {"pos": p}) for i, p in enumerate(np.asarray(pos).tolist()))
    G.add_edges_from(map(tuple, np.asarray(edges).tolist()))
    G.make_mesh(n)
    G.make_submeshes()
    out = {"om": _DeviceMesh(G._dn, None), "kind": kind}
    csr = lambda M: M.tocsr()
    if case in ("c1_mixed", "synthetic_mixed"):
        model = gn.MixedHydraulicNetwork(G, p_bc=gn.Expression("x[1]", degree=2))
        qp = model.solve()
        out.update(A=csr(model.A), b=model.b.tensor.cpu().numpy(), x=qp.vector().get_local())
    elif case in ("c1_primal", "synthetic_primal"):
        model = gn.HydraulicNetwork(G, p_bc=gn.Expression("-x[1]", degree=2))
        A, b = map(gn.ii_assemble, (model.a_form(), model.L_form()))
        A, b = gn.apply_bc(A, b, model.get_bc())
        qp = gn.ii_Function(model.W)
        gn.LUSolver(gn.ii_convert(A), "mumps").solve(qp.vector(), gn.ii_convert(b))
        out.update(A=csr(A), b=b.tensor.cpu().numpy(), x=qp.vector().get_local())
    elif case == "ts_mixed":
        import sympy as sym
        Ainv, Res, T, t_steps = (float(v) for v in dump["params"])
        for e in G.edges():
            G.edges()[e]["Ainv"] = Ainv
            G.edges()[e]["Res"] = Res
        x, t = sym.symbols("x[0] t")
        q = t * sym.sin(2 * 3.14 * x)
        p = sym.cos(2 * 3.14 * x)
        g = Ainv * sym.diff(q, t) + Res * q + sym.diff(p, x)
        f = sym.diff(q, x)
        p, f, g = [gn.Expression(sym.printing.ccode(e), degree=2, t=0) for e in (p, f, g)]
        model = gn.TimeDepMixedHydraulicNetwork(G, f=f, g=g, p_bc=p, Ainv=Ainv)
        qps = gn.time_stepping_stokes(model, gn.Constant(0), t_steps=int(t_steps), T=T, t_step_scheme="IE")
        out["qps"] = np.asarray([s.vector().get_local() for s in qps])
    elif case.startswith("ts_primal"):
        Ainv, Res, T, t_steps = (float(v) for v in dump["params"])
        t = gn.Constant(0)
        xx = gn.SpatialCoordinate(G.mesh)
        q = t * gn.sin(2 * 3.14 * xx[0])
        p = t * gn.cos(2 * 3.14 * xx[0])
        g = Ainv * gn.diff(q, t) + Res * q + G.dds(p)
        f = G.dds(q)
        model = gn.TimeDepHydraulicNetwork(G, f=f, g=g, p_bc=p, Res=Res, Ainv=Ainv)
        qps = gn.time_stepping_stokes(model, t, t_steps=int(t_steps), T=T, t_step_scheme=case.rsplit("_", 1)[1])
        out["qps"] = np.asarray([s.vector().get_local() for s in qps])
    else:
        raise KeyError(case)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("kind,n", [("mixed", 10), ("primal", 6)])
def test_gpu_against_permuted_oracle_dump(kind, n):
    """C1 (Y_bifurcation, n = 10: BASELINE.json configs[0]) through the drop-in API vs a dump in foreign numbering"""
    pos, edges = y_graph()
    dump = refdump.synthetic_dump(kind, pos, edges, n, seed=7)
    rep = refdump.compare(dump, gpu_side(dump))
    assert rep["mesh"] == "bit-exact"


@pytest.mark.gpu
@pytest.mark.skipif(not REAL, reason="no tests/golden/reference_dump_*.npz")
@pytest.mark.parametrize("path", REAL or ["-"])
def test_gpu_against_real_reference_dump(path):
    dump = dict(np.load(path))
    print(os.path.basename(path), refdump.compare(dump, gpu_side(dump)))
