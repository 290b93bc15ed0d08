"""Oracle, finite-element level: DOF layout, element-by-element assembly (COO -> canonical CSR),
L2 projections, right-hand sides, Dirichlet BC application, sparse LU solve.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy/scipy only.

Follows the reference forms literally, one element at a time with the geometry computed the
way a UFC kernel does (Jacobian J = x_b - x_a, detJ = |J|, pseudo-inverse gradient), then
scatter-adds into COO and lets scipy sum duplicates -- i.e. the opposite strategy from the
CUDA path (closed-form row gather), so agreement is a real check.

Canonical numbering (SURVEY.md 8c rule 2; PARITY UNPINNED vs. DOLFIN's re-ordered dofs):
  DG0 dof = cell index, CG1 dof = vertex index of the space's own mesh, Real = 1 dof,
  monolithic offset = cumulative block sizes in W order.
Canonical sparsity (rule 3): structural non-zeros + explicit-zero diagonals that
`init_a_form` (flow_models.py:266-280) puts on every diagonal block; CSR, sorted columns.
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from .graph import BIF_IN, BIF_OUT, BOUN_IN, BOUN_OUT

__all__ = [
    "Coef", "as_coef", "mixed_layout", "primal_layout",
    "element_geometry", "assemble_mixed", "assemble_mixed_rhs", "assemble_mass_q_mixed",
    "project_p1_edges", "assemble_primal", "assemble_primal_rhs", "assemble_mass_q_primal",
    "primal_bc", "apply_bc", "lu_solve", "solve_mixed", "solve_primal",
    "eval_q_edge", "mixed_components",
]


# --------------------------------------------------------------------------------------
# coefficients
# --------------------------------------------------------------------------------------
class Coef:
    """A scalar coefficient: `fn(x[...,gdim]) -> [...]` with a FEniCS-style `degree`.

    degree 0 == Constant; degree 1/2 == Expression(..., degree=d): FEniCS interpolates the
    expression into CG_d on every cell before integrating (that is what `degree` means).
    """

    def __init__(self, fn, degree=2):
        self.fn = fn
        self.degree = degree

    def __call__(self, x):
        x = np.asarray(x, dtype=np.float64)
        out = self.fn(x)
        return np.broadcast_to(np.asarray(out, dtype=np.float64), x.shape[:-1]).copy()


def as_coef(c):
    if isinstance(c, Coef):
        return c
    if callable(c):
        return Coef(c, getattr(c, "degree", 2))
    val = float(c)
    return Coef(lambda x: np.full(x.shape[:-1], val), 0)


# --------------------------------------------------------------------------------------
# layouts
# --------------------------------------------------------------------------------------
def mixed_layout(om):
    """W = [q_0..q_{E-1}, p, lam_0..lam_{B-1}]  (flow_models.py:175-187)."""
    E, m, B = om.E, om.m, om.B
    q_off = np.arange(E, dtype=np.int64) * (m + 1)
    p_off = E * (m + 1)
    lam_off = p_off + E * m
    return dict(q_off=q_off, p_off=p_off, lam_off=lam_off, N=lam_off + B)


def primal_layout(om):
    """W = [DG0(mesh), CG1(mesh)]  (flow_models.py:72-73)."""
    nc, nv = len(om.cells), len(om.coords)
    return dict(q_off=0, p_off=nc, N=nc + nv)


# --------------------------------------------------------------------------------------
# element geometry
# --------------------------------------------------------------------------------------
def element_geometry(xa, xb, tau):
    """Per interval cell with vertices (a,b) embedded in R^gdim and unit tangent tau.

    J = x_b - x_a, h = |J| (= detJ of the pseudo-determinant), physical gradient of the P1
    basis phi_b is J/h^2 (pseudo-inverse), so  d/ds phi_b = (J . tau)/h^2 = -d/ds phi_a.
    Returns (h, dsb).   (G.dds_i: fenics_graph.py:261-266)
    """
    J = xb - xa
    h = np.sqrt((J * J).sum(axis=-1))
    dsb = (J * tau).sum(axis=-1) / (h * h)
    return h, dsb


def _csr(rows, cols, vals, N):
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                      shape=(N, N)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A


def _edge_geom(om):
    """h[E,m], dsb[E,m] for every submesh cell (local vertices a<b) w.r.t. its edge tangent."""
    a = om.sub_cells[:, :, 0]
    b = om.sub_cells[:, :, 1]
    ar = np.arange(om.E)[:, None]
    xa = om.sub_coords[ar, a]
    xb = om.sub_coords[ar, b]
    return element_geometry(xa, xb, om.tangents[:, None, :])


def _bif_incidence(om):
    """[(b_ix, edge_ix, sign, tag)] in the reference's loop order (flow_models.py:242-262):
    for each bifurcation, in-edges (sign +1, BIF_IN) then out-edges (sign -1, BIF_OUT)."""
    out = []
    for b_ix, b in enumerate(om.bif_ixs):
        for i, (u, v) in enumerate(om.edges):
            if v == b:
                out.append((b_ix, i, 1.0, BIF_IN))
        for i, (u, v) in enumerate(om.edges):
            if u == b:
                out.append((b_ix, i, -1.0, BIF_OUT))
    return out


def _bif_incidence_fast(om):
    bpos = {b: k for k, b in enumerate(om.bif_ixs)}
    out = []
    for i, (u, v) in enumerate(om.edges):
        if int(v) in bpos:
            out.append((bpos[int(v)], i, 1.0, BIF_IN))
        if int(u) in bpos:
            out.append((bpos[int(u)], i, -1.0, BIF_OUT))
    return out


# --------------------------------------------------------------------------------------
# mixed model  (MixedHydraulicNetwork / NetworkStokes / TimeDepMixedHydraulicNetwork)
# --------------------------------------------------------------------------------------
def assemble_mixed(om, Res=1.0, visc=None):
    """Monolithic matrix of MixedHydraulicNetwork.a_form (flow_models.py:200-264),
    ii_assemble + ii_convert (:343-344), canonical numbering/sparsity.

    Res  : scalar or [E] (edge attribute "Res", default Constant(1), :228-231); or [E, m]: one value per cell of
           every submesh (submesh cell order) -- a resistance given as a function of position, sampled cell-wise
    visc : None, or scalar/[E] = mu/Ainv_e -> adds NetworkStokes' (mu/Ainv) dds(q) dds(v)
           (flow_models.py:399; the reference reads self.mu which its ctor never sets)
    """
    lay = mixed_layout(om)
    E, m, N = om.E, om.m, lay["N"]
    Res = np.asarray(Res, dtype=np.float64)
    h, dsb = _edge_geom(om)
    a = om.sub_cells[:, :, 0] + lay["q_off"][:, None]
    b = om.sub_cells[:, :, 1] + lay["q_off"][:, None]
    pc = lay["p_off"] + om.parent_cell          # Restriction(p, submesh_i): parent cell dof
    R = Res if Res.ndim == 2 else np.broadcast_to(Res, (E,))[:, None]
    rows, cols, vals = [], [], []

    def add(r, c, v):
        rows.append(np.asarray(r).ravel()); cols.append(np.asarray(c).ravel())
        vals.append(np.broadcast_to(np.asarray(v, dtype=np.float64), np.shape(r)).ravel())

    # a[i][i] += Res q v dx_i  (:233)
    add(a, a, R * h * (2.0 / 6.0)); add(a, b, R * h * (1.0 / 6.0))
    add(b, a, R * h * (1.0 / 6.0)); add(b, b, R * h * (2.0 / 6.0))
    if visc is not None:
        k = np.broadcast_to(np.asarray(visc, dtype=np.float64), (E,))[:, None]
        dsa = -dsb
        add(a, a, k * h * dsa * dsa); add(a, b, k * h * dsa * dsb)
        add(b, a, k * h * dsb * dsa); add(b, b, k * h * dsb * dsb)
    # a[E][i] += + dds_i(q) phi dx_i  (:235)
    add(pc, a, -h * dsb); add(pc, b, h * dsb)
    # a[i][E] += - p dds_i(v) dx_i    (:236)
    add(a, pc, h * dsb); add(b, pc, -h * dsb)
    # vertex terms (:240-262)
    for b_ix, i, sign, tag in _bif_incidence_fast(om):
        for l in np.nonzero(om.vf[i] == tag)[0]:
            qd = lay["q_off"][i] + l
            ld = lay["lam_off"] + b_ix
            add([qd], [ld], [sign]); add([ld], [qd], [sign])
    # explicit-zero diagonal blocks for p and lambda (init_a_form :275-278)
    d = np.arange(lay["p_off"], N)
    add(d, d, np.zeros(len(d)))
    return _csr(rows, cols, vals, N)


def assemble_mass_q_mixed(om, coeff=1.0):
    """TimeDepMixedHydraulicNetwork.mass_matrix_q (time_stepping.py:75-101): coeff*(q_i,v_i)
    on every edge block, explicit-zero diagonal blocks for p and lambda."""
    lay = mixed_layout(om)
    N = lay["N"]
    h, _ = _edge_geom(om)
    a = om.sub_cells[:, :, 0] + lay["q_off"][:, None]
    b = om.sub_cells[:, :, 1] + lay["q_off"][:, None]
    c = float(coeff)
    d = np.arange(lay["p_off"], N)
    rows = [a.ravel(), a.ravel(), b.ravel(), b.ravel(), d]
    cols = [a.ravel(), b.ravel(), a.ravel(), b.ravel(), d]
    vals = [(c * h * (2.0 / 6.0)).ravel(), (c * h / 6.0).ravel(), (c * h / 6.0).ravel(),
            (c * h * (2.0 / 6.0)).ravel(), np.zeros(len(d))]
    return _csr(rows, cols, vals, N)


def _p1_load(h, fa, fm, fb):
    """int f_I phi_a, int f_I phi_b with f_I the quadratic through (fa,fm,fb) (Simpson, exact)."""
    return h * (fa + 2.0 * fm) / 6.0, h * (2.0 * fm + fb) / 6.0


def _nodal_and_mid(om, coef):
    ar = np.arange(om.E)[:, None]
    xa = om.sub_coords[ar, om.sub_cells[:, :, 0]]
    xb = om.sub_coords[ar, om.sub_cells[:, :, 1]]
    nod = coef(om.sub_coords)                      # [E, m+1]
    fa = nod[ar, om.sub_cells[:, :, 0]]
    fb = nod[ar, om.sub_cells[:, :, 1]]
    if coef.degree >= 2:
        fm = coef(0.5 * (xa + xb))
    else:
        fm = 0.5 * (fa + fb)
    return nod, fa, fm, fb


def project_p1_edges(om, coef):
    """`project(coef, FunctionSpace(submesh_i,'CG',1))` for every edge (flow_models.py:315-317).

    A12: consistent-mass L2 projection, direct solve.  Returns nodal values [E, m+1] in
    submesh-local vertex numbering.
    """
    coef = as_coef(coef)
    E, m = om.E, om.m
    h, _ = _edge_geom(om)
    _, fa, fm, fb = _nodal_and_mid(om, coef)
    la, lb = _p1_load(h, fa, fm, fb)
    out = np.empty((E, m + 1))
    for i in range(E):
        a, b = om.sub_cells[i, :, 0], om.sub_cells[i, :, 1]
        r = np.zeros(m + 1)
        np.add.at(r, a, la[i]); np.add.at(r, b, lb[i])
        M = sp.coo_matrix((np.concatenate([h[i] / 3, h[i] / 6, h[i] / 6, h[i] / 3]),
                           (np.concatenate([a, a, b, b]), np.concatenate([a, b, a, b]))),
                          shape=(m + 1, m + 1)).tocsc()
        out[i] = spla.splu(M).solve(r)
    return out


def assemble_mixed_rhs(om, f=0.0, g=0.0, p_bc=0.0, project=True):
    """Monolithic vector of MixedHydraulicNetwork.L_form (flow_models.py:299-332).

    project=True follows the reference: f, g, p_bc are first L2-projected onto CG1 of every
    edge (:315-317); project=False uses nodal interpolation (identical for Constants and for
    linear data up to rounding) and is only there so big cases stay cheap.
    """
    lay = mixed_layout(om)
    E, m, N = om.E, om.m, lay["N"]
    f, g, p_bc = as_coef(f), as_coef(g), as_coef(p_bc)
    h, _ = _edge_geom(om)
    ar = np.arange(E)[:, None]
    a, b = om.sub_cells[:, :, 0], om.sub_cells[:, :, 1]

    def nodal(c):
        if c.degree == 0 or not project:
            return c(om.sub_coords)
        return project_p1_edges(om, c)

    fs, gs, pbs = nodal(f), nodal(g), nodal(p_bc)
    rhs = np.zeros(N)
    qa, qb = a + lay["q_off"][:, None], b + lay["q_off"][:, None]
    # L[i] -= p_bcs[i] v ds(BOUN_OUT) - p_bc v ds(BOUN_IN)      (:324)
    exact = p_bc(om.sub_coords)
    e_out, l_out = np.nonzero(om.vf == BOUN_OUT)
    np.add.at(rhs, lay["q_off"][e_out] + l_out, -pbs[e_out, l_out])
    e_in, l_in = np.nonzero(om.vf == BOUN_IN)
    np.add.at(rhs, lay["q_off"][e_in] + l_in, exact[e_in, l_in])
    # L[i] += gs[i] v dx_i   (:325)   P1 x P1 mass
    ga, gb = gs[ar, a], gs[ar, b]
    np.add.at(rhs, qa.ravel(), (h * (2 * ga + gb) / 6.0).ravel())
    np.add.at(rhs, qb.ravel(), (h * (ga + 2 * gb) / 6.0).ravel())
    # L[E] += fs[i] phi dx_i (:327)
    fa, fb = fs[ar, a], fs[ar, b]
    np.add.at(rhs, (lay["p_off"] + om.parent_cell).ravel(), (h * (fa + fb) / 2.0).ravel())
    return rhs


# --------------------------------------------------------------------------------------
# primal model  (HydraulicNetwork / TimeDepHydraulicNetwork)
# --------------------------------------------------------------------------------------
def _global_geom(om):
    xa = om.coords[om.cells[:, 0]]
    xb = om.coords[om.cells[:, 1]]
    return element_geometry(xa, xb, om.tangents[om.mf])


def assemble_primal(om, Res=1.0):
    """HydraulicNetwork.a_form (flow_models.py:79-91): q in DG0(mesh), p in CG1(mesh).
    Canonical sparsity keeps an explicit (zero) diagonal on the p block so that apply_bc
    has somewhere to put the unit diagonal (time_stepping.py:45 does the same explicitly)."""
    lay = primal_layout(om)
    nc, N = len(om.cells), lay["N"]
    h, dsb = _global_geom(om)
    Res = np.broadcast_to(np.asarray(Res, dtype=np.float64), (nc,))
    qc = np.arange(nc)
    pa = lay["p_off"] + om.cells[:, 0]
    pb = lay["p_off"] + om.cells[:, 1]
    d = np.arange(lay["p_off"], N)
    rows = [qc, qc, qc, pa, pb, d]
    cols = [qc, pa, pb, qc, qc, d]
    vals = [Res * h,                 # a00 = Res q v dx            (:87)
            -h * dsb, h * dsb,       # a01 = dds(p) v dx           (:88)
            h * dsb, -h * dsb,       # a10 = - dds(phi) q dx       (:89)
            np.zeros(len(d))]
    return _csr(rows, cols, vals, N)


def assemble_mass_q_primal(om, coeff=1.0):
    """TimeDepHydraulicNetwork.mass_matrix_q (time_stepping.py:27-47)."""
    lay = primal_layout(om)
    nc, N = len(om.cells), lay["N"]
    h, _ = _global_geom(om)
    d = np.arange(N)
    v = np.zeros(N)
    v[:nc] = float(coeff) * h
    return sp.csr_matrix((v, (d, d)), shape=(N, N))


def assemble_primal_rhs(om, f=0.0, g=0.0):
    """HydraulicNetwork.L_form (flow_models.py:94-101): L0 = g v dx, L1 = f phi dx."""
    lay = primal_layout(om)
    nc, N = len(om.cells), lay["N"]
    f, g = as_coef(f), as_coef(g)
    h, _ = _global_geom(om)
    xa, xb = om.coords[om.cells[:, 0]], om.coords[om.cells[:, 1]]
    xm = 0.5 * (xa + xb)

    def abm(c):
        ca, cb = c(xa), c(xb)
        cm = c(xm) if c.degree >= 2 else 0.5 * (ca + cb)
        return ca, cm, cb

    rhs = np.zeros(N)
    ga, gm, gb = abm(g)
    rhs[:nc] = h * (ga + 4.0 * gm + gb) / 6.0
    fa, fm, fb = abm(f)
    la, lb = _p1_load(h, fa, fm, fb)
    np.add.at(rhs, lay["p_off"] + om.cells[:, 0], la)
    np.add.at(rhs, lay["p_off"] + om.cells[:, 1], lb)
    return rhs


def primal_bc(om, p_bc=0.0):
    """DirichletBC(W[1], p_bc, 'on_boundary') (flow_models.py:103-105).  A13: on the global
    1-D mesh the boundary facets are the vertices with exactly one incident cell."""
    lay = primal_layout(om)
    cnt = np.bincount(om.cells.ravel(), minlength=len(om.coords))
    verts = np.nonzero(cnt == 1)[0]
    vals = as_coef(p_bc)(om.coords[verts])
    return lay["p_off"] + verts, vals


def apply_bc(A, b, dofs, vals):
    """xii.apply_bc (flow_models.py:117): symmetric elimination -- columns moved to the rhs,
    rows and columns zeroed (pattern kept), unit diagonal, b[dof] = value.  A11."""
    A = A.tocsr(copy=True)
    b = b.copy()
    N = A.shape[0]
    g = np.zeros(N)
    g[dofs] = vals
    is_bc = np.zeros(N, dtype=bool)
    is_bc[dofs] = True
    b -= A @ g
    rows = np.repeat(np.arange(N), np.diff(A.indptr))
    kill = is_bc[rows] | is_bc[A.indices]
    A.data[kill] = 0.0
    diag = kill & (rows == A.indices) & is_bc[rows]
    A.data[diag] = 1.0
    # a caller's sum of blocks (time loop: D + theta*dt*A) may have pruned the explicit-zero p
    # diagonal: insert the missing unit entries without touching the other explicit zeros
    has_diag = np.zeros(N, dtype=bool)
    has_diag[rows[diag]] = True
    miss = np.asarray(dofs)[~has_diag[dofs]]
    if len(miss):
        A = sp.coo_matrix((np.concatenate([A.data, np.ones(len(miss))]),
                           (np.concatenate([rows, miss]), np.concatenate([A.indices, miss]))), shape=A.shape).tocsr()
        A.sort_indices()
    b[dofs] = vals
    return A, b


# --------------------------------------------------------------------------------------
# solve + helpers
# --------------------------------------------------------------------------------------
def lu_solve(A, b):
    """LUSolver(A,'mumps').solve(x,b) (flow_models.py:347-348) with SuperLU standing in."""
    return spla.splu(sp.csc_matrix(A)).solve(np.asarray(b, dtype=np.float64))


def solve_mixed(om, Res=1.0, f=0.0, g=0.0, p_bc=0.0, visc=None, project=True):
    """MixedHydraulicNetwork.solve (flow_models.py:337-350) -> (A, b, x)."""
    A = assemble_mixed(om, Res, visc)
    b = assemble_mixed_rhs(om, f, g, p_bc, project=project)
    return A, b, lu_solve(A, b)


def solve_primal(om, Res=1.0, f=0.0, g=0.0, p_bc=0.0):
    """HydraulicNetwork.solve (flow_models.py:108-126) -> (A, b, x) after apply_bc."""
    A = assemble_primal(om, Res)
    b = assemble_primal_rhs(om, f, g)
    dofs, vals = primal_bc(om, p_bc)
    A, b = apply_bc(A, b, dofs, vals)
    return A, b, lu_solve(A, b)


def mixed_components(om, x):
    """Split a monolithic mixed vector: (q[E,m+1] submesh-local, p[cells], lam[B])."""
    lay = mixed_layout(om)
    q = x[:lay["p_off"]].reshape(om.E, om.m + 1)
    return q, x[lay["p_off"]:lay["lam_off"]], x[lay["lam_off"]:]


def eval_q_edge(om, q, i, point):
    """q_i(point): P1 evaluation on submesh i (what `qp[i](pos)` does, tests/test_flow_models.py:67)."""
    point = np.asarray(point, dtype=np.float64)
    xa = om.sub_coords[i, om.sub_cells[i, :, 0]]
    xb = om.sub_coords[i, om.sub_cells[i, :, 1]]
    J = xb - xa
    t = ((point - xa) * J).sum(axis=1) / (J * J).sum(axis=1)
    d = np.linalg.norm(xa + t[:, None] * J - point, axis=1)
    ok = (t > -1e-10) & (t < 1 + 1e-10)
    d = np.where(ok, d, np.inf)
    c = int(np.argmin(d))
    a, b = om.sub_cells[i, c]
    return (1 - t[c]) * q[i, a] + t[c] * q[i, b]
