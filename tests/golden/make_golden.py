#!/usr/bin/env python
"""Generate tests/golden/graphs.json by RUNNING THE REFERENCE'S OWN PYTHON (read-only, from
/root/reference) with its un-installable third-party imports (`fenics`, `xii`) stubbed out.

What this pins (bit-exact, floats stored as float.hex()):
  * node order / positions and `G.edges()` order of every reference generator
    (graphnics/generators.py:14-151, data/generate_arterial_tree.py:34-188),
  * `bifurcation_ixs`, `boundary_ixs` (fenics_graph.py:161-184),
  * per-edge unit tangents (fenics_graph.py:239-247),
  * weighted vertex degrees of the Y bifurcation (fenics_graph.py:200-226).
What it cannot pin: anything that needs DOLFIN (mesh, dofs, matrices) -- `get_mesh`,
`interpolate` and `EmbeddedMesh` are replaced by no-ops below.

Run here (container with /root/reference):   python tests/golden/make_golden.py
The output travels with the repo; nothing on the GPU box reads /root/reference.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    fen = types.ModuleType("fenics")

    class UserExpression:                      # base class only; never evaluated
        def __init__(self, *a, **k):
            pass

    class Constant:
        def __init__(self, v=0):
            self.v = v

    def _noop(*a, **k):
        return None

    for name in ("Mesh", "MeshEditor", "MeshFunction", "refine", "adapt", "interpolate",
                 "VectorFunctionSpace", "FunctionSpace", "Function", "dot", "grad",
                 "set_log_level", "Expression", "Measure", "TrialFunction", "TestFunction",
                 "project", "DirichletBC", "LUSolver"):
        setattr(fen, name, _noop)
    fen.UserExpression = UserExpression
    fen.Constant = Constant
    fen.__all__ = [n for n in dir(fen) if not n.startswith("_")]
    xii = types.ModuleType("xii")
    xii.EmbeddedMesh = _noop
    xii.__all__ = ["EmbeddedMesh"]
    sys.modules["fenics"] = fen
    sys.modules["xii"] = xii


def _load(name, path, package=None):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    _stub_modules()
    pkg = types.ModuleType("graphnics")
    pkg.__path__ = [os.path.join(REF, "graphnics")]
    sys.modules["graphnics"] = pkg
    fg = _load("graphnics.fenics_graph", os.path.join(REF, "graphnics/fenics_graph.py"))
    # DOLFIN-dependent halves become no-ops; everything else is the reference's code
    fg.FenicsGraph.get_mesh = lambda self, n=1: (None, None)
    gen = _load("graphnics.generators", os.path.join(REF, "graphnics/generators.py"))
    for mod in (fg, gen):
        for k, v in vars(mod).items():
            if not k.startswith("_"):
                setattr(pkg, k, v)
    tree = _load("generate_arterial_tree", os.path.join(REF, "data/generate_arterial_tree.py"))
    return fg, gen, tree


def hexlist(a):
    return [[float(x).hex() for x in row] for row in np.asarray(a, dtype=np.float64).reshape(len(a), -1)]


def dump(G):
    nodes = list(G.nodes())
    out = {
        "nodes": [int(v) for v in nodes],
        "pos": hexlist([G.nodes[v]["pos"] for v in nodes]),
        "edges": [[int(u), int(v)] for u, v in G.edges()],
        "bifurcation_ixs": [int(v) for v in G.bifurcation_ixs],
        "boundary_ixs": [int(v) for v in G.boundary_ixs],
        "tangents": hexlist([t for _, t in G.tangents]),
        "tangent_keys": [[int(u), int(v)] for (u, v), _ in G.tangents],
    }
    rad = [G.edges[e].get("radius") for e in G.edges()]
    if all(r is not None for r in rad):
        out["radius"] = [float(r).hex() for r in rad]
    return out


def main():
    fg, gen, tree = load_reference()
    cases = {
        "line_graph_3_dim3": gen.line_graph(3, dim=3),
        "line_graph_2_dx2": gen.line_graph(2, dx=2),
        "Y_bifurcation": gen.Y_bifurcation(),
        "Y_bifurcation_dim3": gen.Y_bifurcation(dim=3),
        "YY_bifurcation": gen.YY_bifurcation(),
        "honeycomb_1_1": gen.honeycomb(1, 1),
        "honeycomb_2_2": gen.honeycomb(2, 2),
        "honeycomb_4_4": gen.honeycomb(4, 4),
        "honeycomb_3_5": gen.honeycomb(3, 5),
    }
    for N in range(1, 8):
        signs = np.tile([-1, 1], 200).tolist()
        cases[f"arterial_tree_N{N}"] = tree.make_arterial_tree(N, directions=signs, gam=0.9, radius0=0.1)
    out = {k: dump(G) for k, G in cases.items()}
    # vertex degrees (tests/test_fenics_graph.py:29-35)
    G = gen.Y_bifurcation()
    G.compute_vertex_degrees()
    out["Y_bifurcation"]["degree"] = [float(G.nodes[v]["degree"]).hex() for v in G.nodes()]
    import networkx as nx, scipy
    out["_meta"] = {"networkx": nx.__version__, "numpy": np.__version__, "scipy": scipy.__version__,
                    "generated_by": "tests/golden/make_golden.py"}
    with open(os.path.join(HERE, "graphs.json"), "w") as fh:
        json.dump(out, fh, indent=0, separators=(",", ":"))
    print({k: (len(v["nodes"]), len(v["edges"])) for k, v in out.items() if k != "_meta"})


if __name__ == "__main__":
    main()
