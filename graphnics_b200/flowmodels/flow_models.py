"""Flow models, drop-in for graphnics/flowmodels/flow_models.py:53-430.

Same classes, constructor signatures, attributes and methods (`a_form`, `L_form`, `get_bc`,
`solve`, `init_a_form`, `init_L_form`); the forms are opaque handles that `ii_assemble` turns into
device CSR / vectors (graphnics_b200.la), and `solve()` is assembly kernels + the Schur-complement
solver instead of ii_assemble -> ii_convert -> MUMPS.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib
from ..coefficients import Constant, as_coefficient
from ..device import current_stream
from ..engine import MixedSystem
from ..fenics_graph import VaryingEdgeAttribute
from ..function import FunctionSpace, TrialFunction, TestFunction, ii_Function
from ..la import AssembledMatrix, AssembledVector, ii_assemble, apply_bc
from ..mesh import SubMesh

__all__ = ["HydraulicNetwork", "MixedHydraulicNetwork", "NetworkStokes", "NetworkPoisson", "RT"]

# Network Raviart-Thomas type space (flow_models.py:132-137)
RT = {
    "flux_space": "CG",
    "flux_degree": 1,
    "pressure_space": "DG",
    "pressure_degree": 0,
}


def _system(G, kind):
    """One engine object per (graph mesh, model family): pattern and solver state are built once."""
    if G._dn is None:
        raise RuntimeError("call G.make_mesh() before building a model")
    if kind not in G._systems:
        if kind == "mixed":
            G._systems[kind] = MixedSystem(G._dn, group=getattr(G, "_group", None))
        else:
            from ..engine_primal import PrimalSystem
            G._systems[kind] = PrimalSystem(G._dn)
    return G._systems[kind]


class _MixedSpaces:
    """W = [CG1(submesh_0..E-1), DG0(mesh), R x B] (flow_models.py:175-187), created lazily."""

    def __init__(self, G):
        self.G, self.net = G, G._dn
        self.E, self.B, self.m = G._dn.E, G._dn.B, G._dn.m
        self._cache = {}

    def __len__(self):
        return self.E + 1 + self.B

    def sizes(self):
        return [self.m + 1] * self.E + [self.E * self.m] + [1] * self.B

    _offsets = {}

    def offsets(self):
        """block offsets into the monolithic vector ([len+1] int64), cached per (E, m, B)"""
        key = (self.E, self.m, self.B)
        off = _MixedSpaces._offsets.get(key)
        if off is None:
            E, m, B = key
            q = np.arange(E + 1, dtype=np.int64) * (m + 1)
            p_end = q[-1] + E * m
            off = _MixedSpaces._offsets[key] = np.concatenate([q, p_end + np.arange(B + 1, dtype=np.int64)])
        return off

    def role(self, i):
        if i < self.E:
            return "q_edge", i
        return ("cell", None) if i == self.E else ("real", None)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if i not in self._cache:
            if i < self.E:
                V = FunctionSpace(SubMesh(self.net, i), RT["flux_space"], RT["flux_degree"])
            elif i == self.E:
                V = FunctionSpace(self.G.mesh, RT["pressure_space"], RT["pressure_degree"])
            else:
                V = FunctionSpace(self.G.mesh, "R", 0)
            self._cache[i] = V
        return self._cache[i]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _LazyArgs:
    def __init__(self, W, fn):
        self.W, self.fn = W, fn

    def __len__(self):
        return len(self.W)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self.fn(V) for V in self.W[i]]
        return self.fn(self.W[i])

    def __iter__(self):
        return (self.fn(V) for V in self.W)


def _coef_key(cc):
    """what a compiled coefficient evaluates to depends on its program and the CURRENT parameter values only"""
    if cc.user is not None:
        return None
    if cc.is_constant:
        return ("const", cc.const_value())
    return (tuple(cc._code), tuple(float(cc.ufl.params[s]) for s in cc._pslots), cc.degree)


def _cached(ms, slot, cc, make):
    """boundary data derived from an Expression (projected end values, node values) are kept on the system while the
    expression and its parameters (e.g. `.t`) stand: a steady solve repeated on the same graph, or a time loop with
    time-independent p_bc, does not launch the projection again"""
    key = _coef_key(cc)
    hit = ms.__dict__.get(slot)
    if key is not None and hit is not None and hit[0] == key:
        return hit[1]
    val = make()
    if key is not None:
        ms.__dict__[slot] = (key, val)
    return val


# ------------------------------------------------------------------------------------------
# form handles
# ------------------------------------------------------------------------------------------
class _MixedAForm:
    """a_form handle; zero=True is init_a_form (flow_models.py:266-280): the zero-weighted diagonal forms that only
    fix the shape of the block system -- it assembles to the pattern with all-zero values"""

    def __init__(self, model, sigma=1.0, mass=0.0, zero=False):
        self.model, self.sigma, self.mass, self.zero = model, sigma, mass, zero

    def assemble(self):
        ms = self.model._system()
        if self.zero:
            ms.build_pattern()
            return AssembledMatrix(ms, torch.zeros(ms.nnz, dtype=torch.float64, device=ms.dev), None, "mixed")
        co = self.model._coeffs(self.sigma, self.mass)
        return AssembledMatrix(ms, ms.assemble(co), co, "mixed")


class _MixedMassForm:
    """TimeDepMixedHydraulicNetwork.mass_matrix_q(coeff) (time_stepping.py:75-101)."""

    def __init__(self, model, coeff):
        self.model, self.coeff = model, coeff

    def assemble(self):
        ms = self.model._system()
        return AssembledMatrix(ms, ms.assemble_mass_q(float(self.coeff)), None, "mass")


class _MassVectorForm:
    def __init__(self, model, u, coeff):
        self.model, self.u, self.coeff = model, u, coeff

    def assemble(self):
        ms = self.model._system()
        x = self.u.tensor() if isinstance(self.u, ii_Function) else self.u
        return AssembledVector(ms, ms.mass_apply(self.model._mass_coeff(self.coeff), x))


class _MixedLForm:
    """MixedHydraulicNetwork.L_form (flow_models.py:299-332).  Like the reference, f, g and the
    outlet p_bc are L2-projected onto CG1 of every edge when the form is BUILT (:315-317); the inlet
    p_bc term is evaluated when the form is ASSEMBLED (:324)."""

    def __init__(self, model, zero=False):
        self.model = model
        ms = model._system()
        dn = ms.dn
        self.f_nodal = self.g_nodal = self.pbc_out = None
        self.f_const = self.g_const = 0.0
        self.zero = zero
        if zero:
            return
        f, g, pbc = (as_coefficient(c) for c in (model.f, model.g, model.p_bc))
        if f.is_constant:
            self.f_const = f.const_value()
        else:
            self.f_nodal = ms.project_coefficient(f)
        if g.is_constant:
            self.g_const = g.const_value()
        else:
            self.g_nodal = ms.project_coefficient(g)
        if pbc.is_constant:
            c = pbc.const_value()
            self.pbc_out = None if c == 0.0 else torch.full((dn.E,), c, dtype=torch.float64, device=dn.device)
        else:
            self.pbc_out = _cached(ms, "_pbc_out_cache", pbc, lambda: ms.project_end_values(pbc))
        self._pbc = pbc

    def rhs_kwargs(self):
        """keyword arguments of MixedSystem.rhs for this form (the inlet p_bc term is evaluated NOW)"""
        ms = self.model._system()
        pbc = as_coefficient(self.model.p_bc)
        pbc_node = None
        if not (pbc.is_constant and pbc.const_value() == 0.0):
            pbc_node = _cached(ms, "_pbc_node_cache", pbc, lambda: pbc.eval_device(ms.dn.pos))
        return dict(f_nodal=self.f_nodal, g_nodal=self.g_nodal, pbc_out=self.pbc_out, pbc_node=pbc_node,
                    f_const=self.f_const, g_const=self.g_const)

    def assemble(self):
        ms = self.model._system()
        if self.zero:
            return AssembledVector(ms, torch.zeros(ms.N, dtype=torch.float64, device=ms.dev))
        return AssembledVector(ms, ms.rhs(**self.rhs_kwargs()))


# ------------------------------------------------------------------------------------------
# models
# ------------------------------------------------------------------------------------------
class MixedHydraulicNetwork:
    """Dual mixed form of the hydraulic equations  Res*q + d/ds p = g,  d/ds q = f  on graph G,
    bifurcation conditions q_in = q_out and continuous normal stress (flow_models.py:140-350).

    Edge attribute "Res" (default 1) is the resistance (:228-231)."""

    def __init__(self, G, f=Constant(0), g=Constant(0), p_bc=Constant(0), space=RT):
        self.G = G
        self.f, self.g, self.p_bc = f, g, p_bc
        if space is not RT and space != RT:
            raise NotImplementedError("only the network Raviart-Thomas pair (CG1 flux, DG0 pressure) is implemented")
        self.W = _MixedSpaces(G)
        self.meshes = _LazyArgs(self.W, lambda V: V.mesh())
        self.qp = _LazyArgs(self.W, TrialFunction)
        self.vphi = _LazyArgs(self.W, TestFunction)

    # -- engine hooks ----------------------------------------------------------------------------
    def _system(self):
        return _system(self.G, "mixed")

    def _visc_edge(self):
        return None

    @staticmethod
    def _mass_coeff(c):
        return float(as_coefficient(c).const_value())

    def _coeffs(self, sigma=1.0, mass=0.0):
        """operator  mass*M + sigma*(Res M + visc K)  with couplings scaled by sigma"""
        ms = self._system()
        visc = self._visc_edge()
        kv = None if visc is None else sigma * visc
        try:
            res = self.G.edge_attribute_array("Res", 1.0)
        except VaryingEdgeAttribute as var:
            # "Res" as a function of position / time (flow_models.py:228-233, the vasomotion demo): sampled per cell
            a_cell = self._mass_coeff(mass) + sigma * ms.cell_values(var.per_edge)
            return ms.coeffs(np.ones(ms.dn.E), kv, sigma, a_cell=a_cell)
        return ms.coeffs(self._mass_coeff(mass) + sigma * res, kv, sigma)

    # -- reference API ---------------------------------------------------------------------------
    def a_form(self, a=None):
        """the bilinear form (flow_models.py:200-264).  `a`: the form to add to -- the reference passes init_a_form()'s
        zero form or nothing; any other handle is refused rather than silently dropped"""
        if a is not None and not (isinstance(a, _MixedAForm) and a.zero):
            raise NotImplementedError("a_form(a): `a` must be None or the zero form returned by init_a_form()")
        return _MixedAForm(self)

    def init_a_form(self):
        return _MixedAForm(self, zero=True)

    def init_L_form(self):
        return _MixedLForm(self, zero=True)

    def L_form(self):
        return _MixedLForm(self)

    def get_bc(self):
        return [[], []]

    def solve(self):
        """a_form + L_form + ii_assemble + ii_convert + LUSolver("mumps") (flow_models.py:337-350) as one
        step: the CSR values (kept on `self.A`) are assembled while the solve runs."""
        W = self.W
        ms = self._system()
        L = self.L_form()
        co = self._coeffs(1.0, 0.0)
        ms.build_pattern()
        vals = torch.empty(ms.nnz, dtype=torch.float64, device=ms.dev)
        b = torch.empty(ms.N, dtype=torch.float64, device=ms.dev)
        # (partitioned network: the multipliers of other ranks' local bifurcations are not computed here -- zero)
        alloc = torch.empty if ms.dn.part is None else torch.zeros
        qp = ii_Function(W, data=alloc(ms.N, dtype=torch.float64, device=ms.dev))
        # a network without cycles has no Krylov part: nothing to report, nothing to wait for
        tree = ms.dn.B == 0 or ms.schur_plan()[3].nc == 0
        _, self.solve_info = ms.assemble_rhs_solve(co, vals, b, qp.tensor(), L.rhs_kwargs(), sync=not tree)
        self.A, self.b = AssembledMatrix(ms, vals, co, "mixed"), AssembledVector(ms, b)
        qp.touch()
        return qp


class NetworkStokes(MixedHydraulicNetwork):
    """MixedHydraulicNetwork + viscous term (mu/Ainv) dds(q) dds(v) (flow_models.py:353-430).
    The reference stores `nu` but reads `self.mu` (AttributeError as shipped); here mu := nu."""

    def __init__(self, G, f=Constant(0), g=Constant(0), p_bc=Constant(0), nu=Constant(1), space=RT):
        self.nu = nu
        self.mu = nu
        super().__init__(G=G, f=f, g=g, p_bc=p_bc, space=space)

    def _visc_edge(self):
        try:
            return float(self.mu) / self.G.edge_attribute_array("Ainv", 1.0)
        except VaryingEdgeAttribute:
            raise NotImplementedError('NetworkStokes: the edge attribute "Ainv" must be one number per edge '
                                      '(a position-dependent "Res" is supported)')


class HydraulicNetwork:
    """Primal mixed form: q in DG0(mesh), p in CG1(mesh), strong Dirichlet p = p_bc on boundary
    nodes (flow_models.py:53-126)."""

    def __init__(self, G, f=Constant(0), g=Constant(0), p_bc=Constant(0), Res=Constant(1), degree=1):
        if degree != 1:
            raise NotImplementedError("only degree=1 (DG0 flux, CG1 pressure) is implemented")
        self.G = G
        self.f, self.g, self.p_bc, self.Res = f, g, p_bc, Res
        self.W = [FunctionSpace(G.mesh, "DG", degree - 1), FunctionSpace(G.mesh, "CG", degree)]
        self.qp = list(map(TrialFunction, self.W))
        self.vphi = list(map(TestFunction, self.W))

    def _system(self):
        return _system(self.G, "primal")

    def _cell_coeff(self, c):
        """float for constants, else the per-cell mean (edge-local order) of a varying coefficient"""
        cc = as_coefficient(c)
        if cc.is_constant:
            return float(cc.const_value())
        return self._system().cell_average(cc)

    _mass_coeff = _cell_coeff

    def _coeffs(self, sigma=1.0, mass=0.0):
        """operator  diag((mass + sigma*Res) h)  with couplings scaled by sigma"""
        a = self._cell_coeff(self.Res)
        mc = self._cell_coeff(mass) if not isinstance(mass, (int, float)) else float(mass)
        return self._system().coeffs(mc + sigma * a, sigma)

    def a_form(self):
        from ..engine_primal import PrimalAForm
        return PrimalAForm(self)

    def L_form(self):
        from ..engine_primal import PrimalLForm
        return PrimalLForm(self)

    def get_bc(self):
        from ..engine_primal import DirichletBC
        return [[], [DirichletBC(self.W[1], self.p_bc, "on_boundary")]]

    def solve(self):
        a = self.a_form()
        L = self.L_form()
        W_bcs = self.get_bc()
        A, b = map(ii_assemble, (a, L))
        A, b = apply_bc(A, b, W_bcs)
        qp = ii_Function(self.W)
        ps = self._system()
        _, self.solve_info = ps.solve(A.coeffs, b.tensor, x=qp.tensor())
        qp.touch()
        return qp


class NetworkPoisson:
    """(kappa d/ds u, d/ds v) = (f, v) on the network, CG1 (flow_models.py:15-50).  Not on the
    north-star path; provided through the primal engine's Schur system (q eliminated)."""

    def __init__(self, G, f=Constant(0), p_bc=Constant(0), kappa=Constant(1), degree=1):
        self.G, self.f, self.p_bc, self.kappa = G, f, p_bc, kappa
        self.V = FunctionSpace(G.mesh, "CG", degree)

    def solve(self):
        # -(kappa u')' = f  <=>  primal hydraulic system with Res = 1/kappa, g = 0
        model = HydraulicNetwork(self.G, f=self.f, g=Constant(0), p_bc=self.p_bc,
                                 Res=Constant(1.0 / float(self.kappa)))
        return model.solve()[1]
