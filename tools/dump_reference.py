#!/usr/bin/env python3
"""Dump what the REAL reference (graphnics on FEniCS / fenics_ii / PETSc-MUMPS) computes, so that the
mesh / DOF / CSR / solution level of the parity contract can be pinned (last paragraph).

Run INSIDE the reference's image (docker/Dockerfile: `ingeborggjerde/graphnics`, base
`ceciledc/fenics_mixed_dimensional:21-06-22` + fenics_ii + cbc.block), from a checkout of the reference:

    docker run --rm -v $PWD:/work -w /work ingeborggjerde/graphnics:main \
        python3 /work/dump_reference.py --out /work/dumps

and copy the resulting `reference_dump_*.npz` files to `tests/golden/` of this repository;
`tests/test_reference_dump.py` picks them up (it skips the real-dump cases while they are absent and always
exercises the comparator on a synthetic dump made from the oracle under a random DOF permutation).

Nothing in this file imports graphnics_b200; it uses only the reference's public API:
  C1  Y_bifurcation + make_mesh(10) + make_submeshes, MixedHydraulicNetwork(G, p_bc = x[1]), Res = 1
      (generators.py:88-111, flowmodels/flow_models.py:155-196, :337-350)       -> reference_dump_c1_mixed.npz
      same graph, HydraulicNetwork(G, p_bc = -x[1]) (:53-126)                    -> reference_dump_c1_primal.npz
  T1  tests/test_time_stepping.py:98-136 (mixed, IE, 10 steps, n = 5)            -> reference_dump_ts_mixed.npz
  T2  tests/test_time_stepping.py:63-95 (primal, IE and CN, 100 steps, n = 8)    -> reference_dump_ts_primal_{IE,CN}.npz

Arrays written per case (all in the REFERENCE's own numbering; the comparator recovers the permutations):
  pos, edges                      graph nodes (iteration order) and G.edges() order
  coordinates, cells, mf          G.mesh.coordinates(), G.mesh.cells(), G.mf.array()
  sub_coords[E,m+1,g], sub_cells[E,m,2], parent_vertex[E,m+1], parent_cell[E,m], vf[E,m+1]    (mixed only)
  block_sizes                     dimension of every space in W
  dofc_<k>                        W[k].tabulate_dof_coordinates() for the first E+1 blocks (q_i, p)
  A_indptr, A_indices, A_data     monolithic matrix after ii_convert, PETSc getValuesCSR (explicit zeros kept)
  b, x                            monolithic right-hand side and solution
  qps                             [t_steps, N] every stored state of time_stepping_stokes (time cases)
"""
import argparse
import os

import numpy as np


def _csr(A):
    from fenics import as_backend_type
    m = as_backend_type(A).mat()
    indptr, indices, data = m.getValuesCSR()
    return np.asarray(indptr, dtype=np.int64), np.asarray(indices, dtype=np.int64), np.asarray(data, dtype=np.float64)


def _mesh_arrays(G):
    import networkx as nx
    out = {
        "pos": np.asarray([G.nodes[v]["pos"] for v in G.nodes()], dtype=np.float64),
        "edges": np.asarray([[u, v] for u, v in G.edges()], dtype=np.int64),
        "coordinates": np.array(G.mesh.coordinates(), dtype=np.float64),
        "cells": np.array(G.mesh.cells(), dtype=np.int64),
        "mf": np.array(G.mf.array(), dtype=np.int64),
        "bifurcation_ixs": np.asarray(G.bifurcation_ixs, dtype=np.int64),
        "boundary_ixs": np.asarray(G.boundary_ixs, dtype=np.int64),
        "tangents": np.asarray([t for _, t in G.tangents], dtype=np.float64),
    }
    subs = list(nx.get_edge_attributes(G, "submesh").values())
    if subs:
        vfs = list(nx.get_edge_attributes(G, "vf").values())
        out["sub_coords"] = np.asarray([s.coordinates() for s in subs], dtype=np.float64)
        out["sub_cells"] = np.asarray([s.cells() for s in subs], dtype=np.int64)
        pv, pc = [], []
        for s in subs:
            # xii.EmbeddedMesh keeps the maps to the parent mesh as {parent mesh id: {tdim: {child: parent}}}
            pmap = list(s.parent_entity_map.values())[0]
            v = pmap[0]
            c = pmap[s.topology().dim()]
            pv.append([v[i] for i in range(s.num_vertices())])
            pc.append([c[i] for i in range(s.num_cells())])
        out["parent_vertex"] = np.asarray(pv, dtype=np.int64)
        out["parent_cell"] = np.asarray(pc, dtype=np.int64)
        out["vf"] = np.asarray([f.array() for f in vfs], dtype=np.int64)
    return out


def _space_arrays(W, n_coord_blocks):
    out = {"block_sizes": np.asarray([V.dim() for V in W], dtype=np.int64)}
    for k in range(n_coord_blocks):
        out["dofc_%d" % k] = np.array(W[k].tabulate_dof_coordinates(), dtype=np.float64)
    return out


def _steady(model, G, W, n_coord_blocks, a=None, L=None, bcs=None):
    from fenics import LUSolver
    from xii import ii_assemble, ii_convert, ii_Function, apply_bc
    a = model.a_form() if a is None else a
    L = model.L_form() if L is None else L
    A, b = map(ii_assemble, (a, L))
    if bcs is not None:
        A, b = apply_bc(A, b, bcs)
    A, b = map(ii_convert, (A, b))
    qp = ii_Function(W)
    LUSolver(A, "mumps").solve(qp.vector(), b)
    out = _mesh_arrays(G)
    out.update(_space_arrays(W, n_coord_blocks))
    out["A_indptr"], out["A_indices"], out["A_data"] = _csr(A)
    out["b"] = np.array(b.get_local(), dtype=np.float64)
    out["x"] = np.array(qp.vector().get_local(), dtype=np.float64)
    return out


def dump_c1(outdir, n=10):
    from fenics import Expression, Constant
    from graphnics import Y_bifurcation, MixedHydraulicNetwork, HydraulicNetwork
    G = Y_bifurcation()
    G.make_mesh(n)
    G.make_submeshes()
    for e in G.edges():
        G.edges()[e]["Res"] = Constant(1)
        G.edges()[e]["Ainv"] = Constant(1)
    model = MixedHydraulicNetwork(G, p_bc=Expression("x[1]", degree=2))
    out = _steady(model, G, model.W, G.number_of_edges() + 1)
    out["case"] = "c1_mixed"; out["n"] = n; out["kind"] = "mixed"
    np.savez_compressed(os.path.join(outdir, "reference_dump_c1_mixed.npz"), **out)
    model = HydraulicNetwork(G, p_bc=Expression("-x[1]", degree=2))
    out = _steady(model, G, model.W, 2, bcs=model.get_bc())
    out["case"] = "c1_primal"; out["n"] = n; out["kind"] = "primal"
    np.savez_compressed(os.path.join(outdir, "reference_dump_c1_primal.npz"), **out)


def _states(qps):
    return np.asarray([np.concatenate([np.array(f.vector().get_local()) for f in qp]) for qp in qps], dtype=np.float64)


def dump_time_mixed(outdir):
    """tests/test_time_stepping.py:98-136, verbatim set-up"""
    import sympy as sym
    from fenics import Expression, Constant
    from graphnics import line_graph, TimeDepMixedHydraulicNetwork, time_stepping_stokes
    G = line_graph(2)
    G.make_mesh(5)
    G.make_submeshes()
    Ainv, Res, T, t_steps = 2, 10, 1, 10
    for e in G.edges():
        G.edges()[e]["Ainv"] = Ainv
        G.edges()[e]["Res"] = Res
    x, t = sym.symbols("x[0] t")
    q = t * sym.sin(2 * 3.14 * x)
    p = sym.cos(2 * 3.14 * x)
    g = Ainv * sym.diff(q, t) + Res * q + sym.diff(p, x)
    f = sym.diff(q, x)
    q, p, f, g = [Expression(sym.printing.ccode(e), degree=2, t=0) for e in (q, p, f, g)]
    tc = Constant(0)
    model = TimeDepMixedHydraulicNetwork(G, f=f, g=g, p_bc=p, Ainv=Ainv)
    qps = time_stepping_stokes(model, tc, t_steps=t_steps, T=T, t_step_scheme="IE")
    out = _mesh_arrays(G)
    out.update(_space_arrays(model.W, G.number_of_edges() + 1))
    out["qps"] = _states(qps)
    out["case"] = "ts_mixed"; out["n"] = 5; out["kind"] = "mixed"
    out["params"] = np.asarray([Ainv, Res, T, t_steps], dtype=np.float64)
    np.savez_compressed(os.path.join(outdir, "reference_dump_ts_mixed.npz"), **out)


def dump_time_primal(outdir):
    """tests/test_time_stepping.py:63-95, verbatim set-up"""
    import ufl
    from fenics import Constant, SpatialCoordinate, sin, cos
    from graphnics import line_graph, TimeDepHydraulicNetwork, time_stepping_stokes
    G = line_graph(2)
    G.make_mesh(8)
    Ainv, Res, T, t_steps = 3, 10, 1, 100
    t = Constant(0)
    xx = SpatialCoordinate(G.mesh)
    q = t * sin(2 * 3.14 * xx[0])
    p = t * cos(2 * 3.14 * xx[0])
    g = Ainv * ufl.diff(q, t) + Res * q + G.dds(p)
    f = G.dds(q)
    for scheme in ("IE", "CN"):
        t.assign(0)
        model = TimeDepHydraulicNetwork(G, f=f, g=g, p_bc=p, Res=Res, Ainv=Ainv)
        qps = time_stepping_stokes(model, t, t_steps=t_steps, T=T, t_step_scheme=scheme)
        out = _mesh_arrays(G)
        out.update(_space_arrays(model.W, 2))
        out["qps"] = _states(qps)
        out["case"] = "ts_primal_" + scheme; out["n"] = 8; out["kind"] = "primal"
        out["params"] = np.asarray([Ainv, Res, T, t_steps], dtype=np.float64)
        np.savez_compressed(os.path.join(outdir, "reference_dump_ts_primal_%s.npz" % scheme), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=".")
    ap.add_argument("--n", type=int, default=10, help="refinement of the C1 case (BASELINE.json configs[0]: 10)")
    ap.add_argument("--cases", default="c1,ts_mixed,ts_primal")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    # the permutation is recoverable either way (dof coordinates are dumped); switching DOLFIN's serial
    # re-ordering off makes the canonical numbering directly visible as well
    try:
        from fenics import parameters
        parameters["reorder_dofs_serial"] = os.environ.get("GNX_DUMP_REORDER", "1") != "0"
    except Exception:
        pass
    which = args.cases.split(",")
    if "c1" in which:
        dump_c1(args.out, args.n)
    if "ts_mixed" in which:
        dump_time_mixed(args.out)
    if "ts_primal" in which:
        dump_time_primal(args.out)


if __name__ == "__main__":
    main()
