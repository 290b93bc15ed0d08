"""Function-space / function façades over device vectors (SURVEY.md 8b "Result objects").

The spaces graphnics builds are fixed (flow_models.py:72-73,175-187): CG1 on an edge SubMesh
(q_i), DG0 on the global mesh (p of the mixed model, q of the primal model), CG1 on the global
mesh (p of the primal model) and Real (lambda_b).  Canonical dof numbering: DG0 dof = cell index,
CG1 dof = vertex index of the space's mesh, Real = one dof.
"""
import numpy as np
import torch

from .mesh import Mesh1D, SubMesh
from .coefficients import as_coefficient, CompiledCoefficient

__all__ = ["FunctionSpace", "VectorFunctionSpace", "Function", "Vector", "ii_Function", "TrialFunction",
           "TestFunction", "project", "interpolate", "errornorm", "ExactFunction"]


class FunctionSpace:
    def __init__(self, mesh, family, degree=0, **kw):
        fam = {"Lagrange": "CG", "P": "CG", "Discontinuous Lagrange": "DG", "Real": "R"}.get(family, family)
        self._mesh, self.family, self.degree = mesh, fam, int(degree)

    def mesh(self):
        return self._mesh

    def dim(self):
        m = self._mesh
        if self.family == "R":
            return 1
        if self.family == "DG" and self.degree == 0:
            return m.num_cells()
        if self.family == "CG" and self.degree == 1:
            return m.num_vertices()
        if self.family == "CG":
            return m.num_vertices() + (self.degree - 1) * m.num_cells()
        if self.family == "DG":
            return (self.degree + 1) * m.num_cells()
        raise NotImplementedError(self.family)

    def tabulate_dof_coordinates(self):
        m = self._mesh
        if self.family == "CG" and self.degree == 1:
            return m.coordinates()
        if self.family == "DG" and self.degree == 0:
            c, x = m.cells(), m.coordinates()
            return 0.5 * (x[c[:, 0]] + x[c[:, 1]])
        if self.family == "R":
            return np.zeros((1, m.geometry().dim()))
        raise NotImplementedError("dof coordinates only for CG1 / DG0 / R")

    def ufl_element(self):
        return (self.family, self.degree)

    def __repr__(self):
        return f"FunctionSpace({self.family}{self.degree}, dim={self.dim()})"


def VectorFunctionSpace(mesh, family, degree, dim=None):
    V = FunctionSpace(mesh, family, degree)
    V.value_dim = dim or mesh.geometry().dim()
    return V


class _Arg:
    def __init__(self, V, kind):
        self.V, self.kind = V, kind

    def function_space(self):
        return self.V


def TrialFunction(V):
    return _Arg(V, "trial")


def TestFunction(V):
    return _Arg(V, "test")


class Vector:
    """dolfin.GenericVector stand-in: a view into a float64 CUDA tensor.  Host copies are cached
    and shared-version invalidated (all views of one monolithic vector share `store`)."""

    def __init__(self, data, store=None):
        self._t = data
        self._store = store if store is not None else {"version": 0}
        self._host = None
        self._host_version = -1

    @property
    def tensor(self):
        return self._t

    def touch(self):
        """call after writing to the underlying device memory"""
        self._store["version"] += 1

    def _to_host(self):
        """device -> PINNED host buffer (torch's caching host allocator) -> numpy view that owns it"""
        t = self._t.detach()
        if not t.is_cuda:
            return t.numpy().copy()
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h.numpy()

    def host_cached(self):
        if self._host is None or self._host_version != self._store["version"]:
            self._host = self._to_host()
            self._host_version = self._store["version"]
        return self._host

    def get_local(self):
        """a fresh host copy (dolfin semantics: the caller owns it)"""
        return self._to_host()

    def set_local(self, arr):
        self._t.copy_(torch.as_tensor(np.ascontiguousarray(arr, dtype=np.float64)).to(self._t.device))
        self.touch()

    def array(self):
        return self.get_local()

    def size(self):
        return int(self._t.numel())

    def __len__(self):
        return self.size()

    def __getitem__(self, k):
        return self.host_cached()[k]

    def __setitem__(self, k, v):
        if isinstance(k, slice) and k == slice(None):
            if np.isscalar(v):
                self._t.fill_(float(v))
                self.touch()
            else:
                self.set_local(np.asarray(v.get_local() if isinstance(v, Vector) else v))
        else:
            a = self.get_local()
            a[k] = v
            self.set_local(a)

    def norm(self, kind="l2"):
        return float(np.linalg.norm(self.host_cached()))

    def copy(self):
        return Vector(self._t.clone())


class Function:
    """dolfin.Function stand-in on one of the spaces above.  Callable at a point."""

    def __init__(self, V, data=None, net=None, role=None, edge=None, store=None):
        self.V = V
        if net is None:
            net = getattr(V.mesh(), "_dn", None)
        if role is None:
            if isinstance(V.mesh(), SubMesh):
                role, edge = "q_edge", V.mesh().i
            else:
                role = {"DG": "cell", "CG": "vertex", "R": "real"}[V.family]
        if data is None:
            dev = net.device if net is not None else "cuda"
            data = torch.zeros(V.dim(), dtype=torch.float64, device=dev)
        self._vec = Vector(data, store) if not isinstance(data, Vector) else data
        self._net = net            # DeviceNetwork (for point evaluation)
        self._role = role          # "q_edge" | "cell" | "vertex" | "real"
        self._edge = edge
        self._name = "f"

    def function_space(self):
        return self.V

    def vector(self):
        return self._vec

    def rename(self, name, label=""):
        self._name = name

    def name(self):
        return self._name

    def assign(self, other):
        self._vec.tensor.copy_(other.vector().tensor if isinstance(other, Function) else other)
        self._vec.touch()

    def copy(self, deepcopy=True):
        return Function(self.V, self._vec.tensor.clone(), self._net, self._role, self._edge)

    def compute_vertex_values(self, mesh=None):
        if self._role in ("q_edge", "vertex"):
            return self.vector().get_local()
        raise NotImplementedError

    # ---- point evaluation ----------------------------------------------------------------------
    def __call__(self, *x):
        x = np.asarray(x[0] if len(x) == 1 and np.ndim(x[0]) > 0 else x, dtype=np.float64).ravel()
        dn = self._net
        vals = self._vec.host_cached()
        if self._role == "real":
            return float(vals[0])
        pos, edges = dn.pos_host(), dn.edges_host()
        if x.size < dn.gdim:
            x = np.concatenate([x, np.zeros(dn.gdim - x.size)])
        x = x[: dn.gdim]
        m = dn.m
        if self._role == "q_edge":
            e = self._edge
        else:
            a, b = pos[edges[:, 0]], pos[edges[:, 1]]
            ab = b - a
            s = np.clip(((x - a) * ab).sum(1) / (ab * ab).sum(1), 0.0, 1.0)
            e = int(np.argmin(np.linalg.norm(a + s[:, None] * ab - x, axis=1)))
        u, v = edges[e]
        ab = pos[v] - pos[u]
        s = float(np.clip(((x - pos[u]) * ab).sum() / (ab * ab).sum(), 0.0, 1.0)) * m
        k = min(int(np.floor(s)), m - 1)            # edge-local cell
        w = s - k
        flip = u > v
        loc = dn.loc_host
        if self._role == "q_edge":
            la, lb = loc[m - k if flip else k], loc[m - k - 1 if flip else k + 1]
            return float((1 - w) * vals[la] + w * vals[lb])
        s_lo = m - 1 - k if flip else k
        if self._role == "cell":
            return float(vals[e * m + (s_lo ^ (s_lo >> 1))])
        # CG1 on the global mesh: vertex ids of edge-local nodes k, k+1
        pv = dn.parent_vertex_host()[e]
        va, vb = pv[loc[m - k if flip else k]], pv[loc[m - k - 1 if flip else k + 1]]
        return float((1 - w) * vals[va] + w * vals[vb])


class ii_Function(list):
    """xii.ii_Function(W): list of component Functions viewing ONE monolithic device vector
    (block order of W, flow_models.py:187).  Components are created lazily for big W."""

    def __init__(self, W, data=None, net=None):
        net = net if net is not None else getattr(W, "net", None)
        if hasattr(W, "offsets"):
            self._offsets = W.offsets()
        else:
            sizes = W.sizes() if hasattr(W, "sizes") else [V.dim() for V in W]
            self._offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        N = int(self._offsets[-1])
        if net is None and len(W):
            net = getattr(W[0].mesh(), "_dn", None)
        dev = net.device if net is not None else "cuda"
        self._data = data if data is not None else torch.zeros(N, dtype=torch.float64, device=dev)
        self._store = {"version": 0}
        self._vector = Vector(self._data, self._store)
        self._W, self._net = W, net
        super().__init__([None] * (len(self._offsets) - 1))

    def _make(self, i):
        V = self._W[i]
        o0, o1 = int(self._offsets[i]), int(self._offsets[i + 1])
        role, edge = self._W.role(i) if hasattr(self._W, "role") else (None, None)
        return Function(V, self._data[o0:o1], self._net, role, edge, store=self._store)

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        f = list.__getitem__(self, i)
        if f is None:
            f = self._make(i)
            list.__setitem__(self, i, f)
        return f

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def vector(self):
        return self._vector

    def block_vec(self):
        return self

    def tensor(self):
        return self._data

    def touch(self):
        self._store["version"] += 1

    def copy(self):
        out = ii_Function(self._W, self._data.clone(), self._net)
        return out


# ----------------------------------------------------------------------------------------
# project / interpolate / errornorm (enough for the reference's manufactured-solution tests)
# ----------------------------------------------------------------------------------------
class ExactFunction:
    """Result of project(expr, V) / interpolate(expr, V) for spaces we never solve in (CG2, CG3 of
    the tests): keeps the expression and evaluates it exactly instead of building the interpolant."""

    def __init__(self, coef, V):
        self.coef, self.V = as_coefficient(coef), V
        self.degree = max(getattr(self.coef, "degree", 2), V.degree)

    def as_coefficient(self):
        c = self.coef
        return CompiledCoefficient(ufl=c.ufl, constant=c.constant, user=c.user, degree=min(self.degree, 2))

    def function_space(self):
        return self.V

    def __call__(self, *x):
        x = np.asarray(x[0] if len(x) == 1 and np.ndim(x[0]) > 0 else x, dtype=np.float64)
        return float(self.coef.eval_host(x.reshape(1, -1))[0])


def project(v, V, **kw):
    """dolfin.project(v, V) for the spaces of the network models:
      * DG0 on the network mesh: the L2 projection = cell means (Simpson per cell), on the device;
      * CG1 on the submesh of one edge (the flux spaces W[i] of the mixed model): the consistent-mass L2 projection
        the model's L_form uses for f, g, p_bc (flow_models.py:315-317), computed by the same kernel;
      * CG1 on the whole network mesh: the nodal interpolant (the L2 projection there couples the edges through the
        junction vertices and is not implemented; the two agree to O(h^2));
      * anything of higher order (CG2 / CG3 in the reference's tests hold exact solutions): the expression itself,
        evaluated exactly where it is used (ExactFunction)."""
    if isinstance(v, Function) and v.function_space() is V:
        return v
    mesh = V.mesh()
    dn = getattr(mesh, "_dn", None)
    if dn is None or isinstance(v, (Function, ExactFunction)):
        return ExactFunction(v, V)
    if V.family == "DG" and V.degree == 0 and isinstance(mesh, Mesh1D):
        from .flowmodels.time_stepping import _cell_means
        return Function(V, _cell_means(dn, as_coefficient(v)).contiguous(), net=dn, role="cell")
    if V.family == "CG" and V.degree == 1 and isinstance(mesh, SubMesh):
        from .engine import MixedSystem
        ms = dn.__dict__.get("_proj_system")
        if ms is None:
            ms = dn.__dict__["_proj_system"] = MixedSystem(dn)
        i, m = mesh.i, dn.m
        whole = ms.project_coefficient(as_coefficient(v))
        return Function(V, whole[i * (m + 1):(i + 1) * (m + 1)].clone(), net=dn, role="q_edge", edge=i)
    if V.family == "CG" and V.degree == 1 and isinstance(mesh, Mesh1D):
        return interpolate(v, V)
    return ExactFunction(v, V)


def interpolate(v, V):
    """dolfin.interpolate.  Into CG1 / DG0 on the global network mesh: a real Function (nodal / midpoint
    values, evaluated on the device for expressions, on the host for UserExpression callbacks); into the
    higher-order spaces the reference's tests use for exact solutions: the expression itself."""
    mesh = V.mesh()
    dn = getattr(mesh, "_dn", None)
    lowest = (V.family == "CG" and V.degree == 1) or (V.family == "DG" and V.degree == 0)
    if dn is not None and isinstance(mesh, Mesh1D) and lowest and not isinstance(v, (Function, ExactFunction)):
        coef = as_coefficient(v)
        if V.family == "CG":
            cells_h = mesh.cells()
            adj = np.zeros(dn.num_vertices, dtype=np.int64)          # one adjacent cell per vertex (eval_cell data)
            adj[cells_h[:, 1]] = np.arange(len(cells_h))
            adj[cells_h[:, 0]] = np.arange(len(cells_h))
            vals = coef.eval_device(dn.coords, cell_index=adj if coef.user is not None else None)
            role = "vertex"
        else:
            c = dn.cells.long()
            mid = (0.5 * (dn.coords[c[:, 0]] + dn.coords[c[:, 1]])).contiguous()
            vals = coef.eval_device(mid, pts_per_edge=dn.m, tangents=dn.tangents, cell_index=np.arange(dn.num_cells))
            role = "cell"
        return Function(V, vals.to(torch.float64).contiguous(), net=dn, role=role)
    return ExactFunction(v, V)


_GL3 = (np.array([-np.sqrt(3.0 / 5.0), 0.0, np.sqrt(3.0 / 5.0)]) * 0.5 + 0.5, np.array([5.0, 8.0, 5.0]) / 18.0)


def errornorm(uh, u, norm_type="L2", **kw):
    """dolfin.errornorm(uh, u) (L2) for uh one of our Functions and u an expression-like
    (tests/test_flow_models.py:155-159).  3-point Gauss quadrature per cell, on the host."""
    if isinstance(u, Function) and not isinstance(uh, Function):
        uh, u = u, uh
    dn = uh._net
    mesh = dn._host_mesh()
    x, cells = mesh.coordinates(), mesh.cells()
    vals = uh.vector().get_local()
    exact = as_coefficient(u if not isinstance(u, ExactFunction) else u.coef)
    if uh._role == "q_edge":
        pv = dn.parent_vertex_host()[uh._edge]
        sc = dn.sub_cells_host()
        xa, xb = x[pv[sc[:, 0]]], x[pv[sc[:, 1]]]
        ua, ub = vals[sc[:, 0]], vals[sc[:, 1]]
        tan = np.repeat(dn.tangents[uh._edge:uh._edge + 1].cpu().numpy(), len(sc), 0)
    elif uh._role == "cell":
        xa, xb = x[cells[:, 0]], x[cells[:, 1]]
        ua = ub = vals
        tan = dn.tangents.cpu().numpy()[np.arange(len(cells)) >> dn.n]
    else:
        xa, xb = x[cells[:, 0]], x[cells[:, 1]]
        ua, ub = vals[cells[:, 0]], vals[cells[:, 1]]
        tan = dn.tangents.cpu().numpy()[np.arange(len(cells)) >> dn.n]
    h = np.linalg.norm(xb - xa, axis=1)
    err2 = 0.0
    for xi, wi in zip(*_GL3):
        xq = xa + xi * (xb - xa)
        e = (1 - xi) * ua + xi * ub - exact.eval_host(xq, tangents=tan)
        err2 += float((wi * h * e * e).sum())
    return float(np.sqrt(err2))
