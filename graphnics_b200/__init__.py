"""graphnics_b200 -- B200-native drop-in for the network-flow hot path of graphnics.

Same flat namespace as graphnics/__init__.py:1-5 (`from graphnics_b200 import *`): the graph
class and tags, the generators, the flow models and time stepping, plus the handful of
`fenics` / `xii` names graphnics user code touches (Constant, Expression, ii_assemble, ...).
There is no CPU fallback: every compute call goes to libgraphnics_b200.so (sm_100a) or raises.
"""
import numpy as np              # noqa: F401  (the reference namespace re-exports np and nx)
import networkx as nx           # noqa: F401

from ._lib import GnxError, LIB_PATH  # noqa: F401
from .coefficients import *     # noqa: F401,F403
from .function import *         # noqa: F401,F403
from .la import *               # noqa: F401,F403
from .fenics_graph import *     # noqa: F401,F403
from .generators import *       # noqa: F401,F403
from .flowmodels import *       # noqa: F401,F403
from .graph_utils import *      # noqa: F401,F403
from .plot import *             # noqa: F401,F403

DOLFIN_EPS = 3.0e-16


def near(a, b, eps=DOLFIN_EPS):
    return abs(a - b) < eps


def set_log_level(level):
    return None
