"""CPU oracle for the graphnics network-flow hot path.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy *restatement* of what the reference
(IngeborgGjerde/graphnics, mounted read-only at /root/reference when available)
computes on its hot path:

    graph -> 1-D mesh (fenics_graph.py:58-99)  -> submeshes/tags (:132-156,280-311)
          -> tangents (:230-253)               -> mixed / primal block forms
          (flowmodels/flow_models.py:53-126,140-350,353-430)
          -> ii_assemble / ii_convert / LUSolver("mumps")   (:337-350)
          -> theta-scheme time loop (flowmodels/time_stepping.py:126-221)

Every floating point operation of that path lives in third-party packages that are
neither vendored under /root/reference nor installable here (FEniCS/DOLFIN 2019.1
mixed-dimensional fork `ceciledc/fenics_mixed_dimensional:21-06-22`, fenics_ii master,
cbc.block master, PETSc+MUMPS: docker/Dockerfile:6,27-36).  The oracle therefore restates
their *published/recalled* behaviour (rules A1-A14); each function
cites the reference call site that relies on the rule.

PARITY STATUS
  * graph level (node/edge order, bifurcation & boundary tables, tangents, generators):
    PINNED against the reference's own Python run here with `fenics`/`xii` stubbed out
    (tests/golden/make_golden.py -> tests/golden/*.json).
  * physics level: PINNED against the reference tests' known answers
    (mass conservation, BC values, manufactured-solution error bounds, Kirchhoff closed
    forms of SURVEY.md Appendix C).
  * mesh numbering / DOF numbering / sparsity / matrix entries vs. real DOLFIN output:
    **PARITY UNPINNED** -- no FEniCS dump exists and DOLFIN's serial dof re-ordering is
    not reproducible offline.  The contract is "canonical (entity-index) numbering".

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  The product (graphnics_b200) never does.
"""

from .graph import *        # noqa: F401,F403
from .mesh import *         # noqa: F401,F403
from .fem import *          # noqa: F401,F403
from .timeloop import *     # noqa: F401,F403
from .elimination import *  # noqa: F401,F403
