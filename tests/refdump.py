"""Comparator for dumps of the REAL reference (tools/dump_reference.py, run in the reference's FEniCS image)
against this repository's canonical numbering -- TEST INFRASTRUCTURE.

The reference's monolithic system lives in DOLFIN's (re-ordered) dof numbering; ours in the canonical one
(DG0 dof = cell, CG1 dof = vertex of the space's own mesh, blocks in W order; SURVEY.md 8c rule 2).  The
comparator
  1. checks the mesh level bit for bit (coordinates, sorted cell pairs, cell->edge marker, SubMesh maps, tags);
  2. recovers the dof permutation of every block from `tabulate_dof_coordinates()` (nearest canonical dof of
     the same block; must be a bijection);
  3. brings A, b, x (and every stored time state) into canonical numbering, drops explicit zeros on both
     sides (DOLFIN stores more of them than the forms need: SURVEY.md App. A9) and asserts identical
     non-zero patterns, entries to 1e-12, right-hand side to 1e-12, fields to 1e-8 relative L2.
`synthetic_dump` writes the same file format from the oracle under a random per-block dof permutation with
extra explicit zeros sprinkled in, so the comparator itself is exercised without FEniCS.
"""
import numpy as np
import scipy.sparse as sp
from scipy.spatial import cKDTree

import oracle


# ---------------------------------------------------------------------------------------------
# canonical side
# ---------------------------------------------------------------------------------------------
def canonical_blocks(om, kind):
    """(block sizes, dof coordinates per block or None for Real blocks) in W order"""
    if kind == "mixed":
        mid = om.coords[om.cells].mean(axis=1)
        sizes = [om.m + 1] * om.E + [om.E * om.m] + [1] * om.B
        coords = [om.sub_coords[i] for i in range(om.E)] + [mid] + [None] * om.B
    else:
        mid = om.coords[om.cells].mean(axis=1)
        sizes = [len(om.cells), len(om.coords)]
        coords = [mid, om.coords]
    return sizes, coords


def oracle_side(dump):
    """what the oracle computes for the dumped case (canonical numbering)"""
    kind = str(dump["kind"])
    om = oracle.build_mesh(dump["pos"], dump["edges"], int(dump["n"]))
    out = {"om": om, "kind": kind}
    case = str(dump["case"])
    if case == "c1_mixed" or case.startswith("synthetic_mixed"):
        A = oracle.assemble_mixed(om, 1.0)
        b = oracle.assemble_mixed_rhs(om, p_bc=oracle.Coef(lambda x: x[..., 1], 2), project=True)
        out.update(A=A, b=b, x=oracle.lu_solve(A, b))
    elif case == "c1_primal" or case.startswith("synthetic_primal"):
        A = oracle.assemble_primal(om, 1.0)
        b = oracle.assemble_primal_rhs(om)
        dofs, vals = oracle.primal_bc(om, oracle.Coef(lambda x: -x[..., 1], 2))
        A, b = oracle.apply_bc(A, b, dofs, vals)
        out.update(A=A, b=b, x=oracle.lu_solve(A, b))
    elif case == "ts_mixed":
        Ainv, Res, T, t_steps = (float(v) for v in dump["params"])
        w = 2 * 3.14
        TC = lambda fn: oracle.TimeCoef(fn, degree=2, clock="attr")
        out["qps"] = np.asarray(oracle.time_stepping(
            "mixed", om, f=TC(lambda x, t: t * w * np.cos(w * x[..., 0])),
            g=TC(lambda x, t: Ainv * np.sin(w * x[..., 0]) + Res * t * np.sin(w * x[..., 0]) - w * np.sin(w * x[..., 0])),
            p_bc=TC(lambda x, t: np.cos(w * x[..., 0])), Res=Res, Ainv=Ainv, t_steps=int(t_steps), T=T, scheme="IE"))
    elif case.startswith("ts_primal"):
        Ainv, Res, T, t_steps = (float(v) for v in dump["params"])
        w = 2 * 3.14
        TC = lambda fn: oracle.TimeCoef(fn, degree=2, clock="const")
        out["qps"] = np.asarray(oracle.time_stepping(
            "primal", om, f=TC(lambda x, t: t * w * np.cos(w * x[..., 0])),
            g=TC(lambda x, t: Ainv * np.sin(w * x[..., 0]) + Res * t * np.sin(w * x[..., 0]) - t * w * np.sin(w * x[..., 0])),
            p_bc=TC(lambda x, t: t * np.cos(w * x[..., 0])), Res=Res, Ainv=Ainv, t_steps=int(t_steps), T=T,
            scheme=case.rsplit("_", 1)[1]))
    else:
        raise KeyError(case)
    return out


# ---------------------------------------------------------------------------------------------
# permutation recovery and comparison
# ---------------------------------------------------------------------------------------------
def recover_permutation(dump, om, kind):
    """can2ref [N]: reference dof index (monolithic) of every canonical dof"""
    sizes, coords = canonical_blocks(om, kind)
    ref_sizes = [int(s) for s in dump["block_sizes"]]
    assert ref_sizes == sizes, ("block sizes differ", ref_sizes[:4], sizes[:4])
    can2ref = np.empty(sum(sizes), dtype=np.int64)
    off = 0
    for k, (sz, xc) in enumerate(zip(sizes, coords)):
        if xc is None:
            assert sz == 1
            can2ref[off] = off
        else:
            xr = np.asarray(dump["dofc_%d" % k], dtype=np.float64).reshape(sz, -1)
            xc = np.asarray(xc, dtype=np.float64).reshape(sz, -1)
            scale = max(float(np.abs(xc).max()), 1e-300)
            dist, idx = cKDTree(xr).query(xc)                  # reference dof nearest to every canonical dof
            assert dist.max() <= 1e-9 * scale, ("dof coordinates of block %d do not match" % k, float(dist.max()))
            assert len(np.unique(idx)) == sz, "dof coordinate match of block %d is not a bijection" % k
            can2ref[off:off + sz] = off + idx
        off += sz
    return can2ref


def _nonzero_csr(A):
    A = sp.csr_matrix(A).copy()
    A.eliminate_zeros()
    A.sort_indices()
    return A


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def field_slices(om, kind):
    if kind == "mixed":
        lay = oracle.mixed_layout(om)
        s = [("q", 0, lay["p_off"]), ("p", lay["p_off"], lay["lam_off"])]
        if om.B:
            s.append(("lam", lay["lam_off"], lay["N"]))
        return s
    lay = oracle.primal_layout(om)
    return [("q", 0, lay["p_off"]), ("p", lay["p_off"], lay["N"])]


def compare(dump, ours):
    """dump: mapping of the arrays tools/dump_reference.py writes; ours: dict(om=OracleMesh-like with the
    canonical mesh arrays, kind=..., A=csr?, b=?, x=?, qps=?) in canonical numbering.  Raises AssertionError
    on the first mismatch; returns a report dict."""
    om, kind = ours["om"], ours["kind"]
    rep = {}
    # ---- 1. mesh level: bit for bit
    assert np.array_equal(np.asarray(dump["coordinates"]).view(np.int64), om.coords.view(np.int64)), "mesh coordinates"
    assert np.array_equal(np.sort(np.asarray(dump["cells"]), axis=1), om.cells), "cells"
    assert np.array_equal(np.asarray(dump["mf"]), om.mf), "cell -> edge marker"
    assert list(dump["bifurcation_ixs"]) == list(om.bif_ixs) and list(dump["boundary_ixs"]) == list(om.bnd_ixs)
    assert np.array_equal(np.asarray(dump["tangents"]).view(np.int64), om.tangents.view(np.int64)), "tangents"
    if "parent_vertex" in dump:
        assert np.array_equal(np.asarray(dump["parent_vertex"]), om.parent_vertex), "SubMesh vertex maps"
        assert np.array_equal(np.asarray(dump["parent_cell"]), om.parent_cell), "SubMesh cell maps"
        assert np.array_equal(np.sort(np.asarray(dump["sub_cells"]), axis=2), om.sub_cells), "SubMesh cells"
        assert np.array_equal(np.asarray(dump["sub_coords"]).view(np.int64), om.sub_coords.view(np.int64))
        assert np.array_equal(np.asarray(dump["vf"]), om.vf), "vertex tags"
    rep["mesh"] = "bit-exact"
    # ---- 2. dof permutation
    can2ref = recover_permutation(dump, om, kind)
    rep["identity_permutation"] = bool(np.array_equal(can2ref, np.arange(len(can2ref))))
    # ---- 3. operator, right-hand side, fields
    if "A_data" in dump and ours.get("A") is not None:
        N = len(can2ref)
        Aref = sp.csr_matrix((np.asarray(dump["A_data"]), np.asarray(dump["A_indices"]), np.asarray(dump["A_indptr"])),
                             shape=(N, N))
        stored = Aref.copy()
        stored.data[:] = 1.0
        Aref_c = _nonzero_csr(Aref[can2ref][:, can2ref])
        Aour = sp.csr_matrix(ours["A"])
        Aour_c = _nonzero_csr(Aour)
        assert np.array_equal(Aref_c.indptr, Aour_c.indptr) and np.array_equal(Aref_c.indices, Aour_c.indices), \
            "non-zero pattern differs"
        np.testing.assert_allclose(Aour_c.data, Aref_c.data, rtol=1e-12, atol=0.0)
        rep["nnz_nonzero"] = int(Aour_c.nnz)
        # our explicit zeros (p / lambda diagonals of init_a_form): are they stored by the reference, too?
        ours_pat = Aour.copy(); ours_pat.data[:] = 1.0
        ours_pat = ours_pat.tocsr()
        stored_c = stored[can2ref][:, can2ref].tocsr()
        extra = (ours_pat - ours_pat.multiply(stored_c))
        extra.eliminate_zeros()
        rep["our_entries_not_stored_by_reference"] = int(extra.nnz)
        rep["reference_explicit_zeros"] = int(Aref.nnz - Aref_c.nnz)
    if "b" in dump and ours.get("b") is not None:
        bref = np.asarray(dump["b"])[can2ref]
        np.testing.assert_allclose(ours["b"], bref, rtol=1e-12, atol=1e-14 * max(np.abs(bref).max(), 1.0))
    if "x" in dump and ours.get("x") is not None:
        xref = np.asarray(dump["x"])[can2ref]
        for name, lo, hi in field_slices(om, kind):
            rep["rel_l2_" + name] = rel(ours["x"][lo:hi], xref[lo:hi])
            assert rep["rel_l2_" + name] < 1e-8, (name, rep)
    if "qps" in dump and ours.get("qps") is not None:
        qref = np.asarray(dump["qps"])[:, can2ref]
        assert qref.shape == np.asarray(ours["qps"]).shape
        worst = 0.0
        for k in range(len(qref)):
            for name, lo, hi in field_slices(om, kind):
                ref = qref[k, lo:hi]
                err = np.linalg.norm(ours["qps"][k][lo:hi] - ref) / max(np.linalg.norm(ref), 1e-30)
                worst = max(worst, float(err)) if np.linalg.norm(ref) > 0 else worst
                assert np.linalg.norm(ours["qps"][k][lo:hi] - ref) <= 1e-8 * max(np.linalg.norm(ref), 1e-30), (k, name)
        rep["worst_state_rel_l2"] = worst
    return rep


# ---------------------------------------------------------------------------------------------
# synthetic dump (oracle under a random dof permutation): exercises the comparator without FEniCS
# ---------------------------------------------------------------------------------------------
def synthetic_dump(kind, pos, edges, n, seed=0, extra_zeros=50, perturb=None):
    rng = np.random.default_rng(seed)
    om = oracle.build_mesh(pos, edges, n)
    case = "synthetic_" + kind
    side = oracle_side({"kind": kind, "case": case, "pos": pos, "edges": edges, "n": n})
    A, b, x = side["A"], side["b"], side["x"]
    sizes, coords = canonical_blocks(om, kind)
    N = sum(sizes)
    ref2can = np.empty(N, dtype=np.int64)         # canonical dof sitting at reference position r
    d = {"case": case, "kind": kind, "n": n, "pos": np.asarray(pos, dtype=np.float64), "edges": np.asarray(edges),
         "coordinates": om.coords.copy(), "cells": om.cells[:, ::-1].copy(),      # unsorted pairs: comparator sorts
         "mf": om.mf.copy(), "bifurcation_ixs": np.asarray(om.bif_ixs), "boundary_ixs": np.asarray(om.bnd_ixs),
         "tangents": om.tangents.copy(), "block_sizes": np.asarray(sizes)}
    if kind == "mixed":
        d.update(sub_coords=om.sub_coords.copy(), sub_cells=om.sub_cells.copy(), parent_vertex=om.parent_vertex.copy(),
                 parent_cell=om.parent_cell.copy(), vf=om.vf.copy())
    off = 0
    for k, (sz, xc) in enumerate(zip(sizes, coords)):
        p = rng.permutation(sz)
        ref2can[off:off + sz] = off + p
        if xc is not None:
            d["dofc_%d" % k] = np.asarray(xc).reshape(sz, -1)[p].copy()
        off += sz
    Aref = sp.csr_matrix(A)[ref2can][:, ref2can].tocoo()
    rows, cols, data = Aref.row, Aref.col, Aref.data.copy()
    if perturb is not None:
        data[perturb[0] % len(data)] *= 1.0 + perturb[1]
    # DOLFIN keeps explicit zeros the forms do not need (App. A9): sprinkle some in
    zr, zc = rng.integers(0, N, size=extra_zeros), rng.integers(0, N, size=extra_zeros)
    have = set(zip(rows.tolist(), cols.tolist()))
    keep = [i for i in range(extra_zeros) if (int(zr[i]), int(zc[i])) not in have]
    rows = np.concatenate([rows, zr[keep]]); cols = np.concatenate([cols, zc[keep]])
    data = np.concatenate([data, np.zeros(len(keep))])
    order = np.lexsort((cols, rows))
    rows, cols, data = rows[order], cols[order], data[order]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=N))])
    d.update(A_indptr=indptr.astype(np.int64), A_indices=cols.astype(np.int64), A_data=data,
             b=b[ref2can].copy(), x=x[ref2can].copy())
    return d
