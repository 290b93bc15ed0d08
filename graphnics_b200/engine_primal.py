"""Host driver of the primal system (q in DG0(mesh) | p in CG1(mesh)) on one GPU, and the form
handles of HydraulicNetwork / TimeDepHydraulicNetwork (graphnics/flowmodels/flow_models.py:53-126,
time_stepping.py:20-64).  Thin orchestration over the C ABI; no arithmetic happens here.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ptr, check
from .coefficients import as_coefficient
from .device import DeviceNetwork, current_stream
from .engine import SolveInfo, lincomb
from .la import AssembledMatrix, AssembledVector

__all__ = ["PrimalSystem", "PrimalAForm", "PrimalLForm", "PrimalMassForm", "DirichletBC"]


class _PCoeffs:
    """gnx_primal_coeffs_t + the tensor it points to."""

    def __init__(self, a_const, a_cell, sigma):
        self.a_const, self.a_cell, self.sigma = float(a_const), a_cell, float(sigma)
        self.c = _lib.gnx_primal_coeffs_t()
        self.c.a_const = float(a_const)
        self.c.a_cell = a_cell.data_ptr() if a_cell is not None else None
        self.c.sigma = float(sigma)

    def ref(self):
        return C.byref(self.c)


class DirichletBC:
    """DirichletBC(V, value, "on_boundary") (flow_models.py:104): boundary = degree-1 graph nodes."""

    def __init__(self, V, value, where="on_boundary"):
        self.V, self.value, self.where = V, value, where


class PrimalSystem:
    """Edge-partitioned networks (DeviceNetwork(part=...)): q and the edge-interior p unknowns are owned
    by the rank that owns the edge, the p unknowns at the graph nodes are replicated (like lambda in the
    mixed model).  b's graph-node block must hold the GLOBAL values on every rank (f = 0, or summed by
    the caller); the solve gathers the edge scalars with one all-gather and replicates the Schur solve."""

    def __init__(self, dn: DeviceNetwork, group=None):
        self.dn = dn
        self.group = group
        self.lib = _lib.require_device()
        self.dev = dn.device
        self.NC, self.NV = dn.num_cells, dn.num_vertices
        self.N = self.NC + self.NV
        self.nnz = int(self.lib.gnx_primal_nnz(dn.net_ref()))
        self.p_off = self.NC
        self.rowptr = self.colidx = None
        self._plan = None
        self._bufs_ = None
        self._red_work = None
        self._pts = None
        self._slot = None

    def _f64(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.dev)

    # ---- coefficients ---------------------------------------------------------------------------
    def cell_points(self):
        """[NC, 3, gdim]: first stored vertex, midpoint, second stored vertex of every cell"""
        if self._pts is None:
            dn = self.dn
            c = dn.cells.long()
            xa, xb = dn.coords[c[:, 0]], dn.coords[c[:, 1]]
            self._pts = torch.stack([xa, 0.5 * (xa + xb), xb], dim=1).contiguous()
        return self._pts

    def cell_slot(self):
        """global cell -> edge-local slot (the order of gnx_network_t.h and of a_cell)"""
        if self._slot is None:
            dn = self.dn
            m, n = dn.m, dn.n
            c = np.arange(self.NC, dtype=np.int64)
            e, j = c >> n, c & (m - 1)
            s = j.copy()
            sh = 1
            while sh < 64:
                s ^= s >> sh
                sh <<= 1
            edges = dn.edges_host()
            flip = (edges[:, 0] > edges[:, 1])[e]
            self._slot = torch.from_numpy(e * m + np.where(flip, m - 1 - s, s)).to(self.dev)
        return self._slot

    def eval_cellwise(self, coef, slot=None):
        """coefficient at the 3 points of every cell -> [NC*3] (cell-wise: may jump between edges).
        `slot` names a persistent output buffer (the time loop re-evaluates f and g every step)."""
        dn = self.dn
        out = None
        if slot is not None:
            if getattr(self, "_cw", None) is None:
                self._cw = {}
            if slot not in self._cw:
                self._cw[slot] = self._f64(3 * self.NC)
            out = self._cw[slot]
        pts = self.cell_points().reshape(-1, dn.gdim)
        cell_index = None
        if coef.user is not None:
            cell_index = np.repeat(np.arange(self.NC), 3)
        tang = dn.tangents
        return coef.eval_device(pts, pts_per_edge=3 * dn.m, tangents=tang, cell_index=cell_index, out=out)

    def cell_average(self, coef):
        """per-cell mean of a coefficient in EDGE-LOCAL order (Simpson; midpoint rule for degree <= 1)"""
        v = self.eval_cellwise(coef).reshape(self.NC, 3)
        if coef.degree <= 1:
            avg = v[:, 1] if coef.degree == 0 else 0.5 * (v[:, 0] + v[:, 2])
        else:
            avg = (v[:, 0] + 4.0 * v[:, 1] + v[:, 2]) / 6.0
        out = self._f64(self.NC)
        out[self.cell_slot()] = avg
        return out

    def coeffs(self, a, sigma=1.0):
        """a: float, or a per-cell tensor in edge-local order"""
        if torch.is_tensor(a):
            return _PCoeffs(0.0, a.contiguous(), sigma)
        return _PCoeffs(float(a), None, sigma)

    # ---- pattern / assembly ---------------------------------------------------------------------
    def build_pattern(self):
        if self.rowptr is None:
            self.rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=self.dev)
            self.colidx = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
            check(self.lib.gnx_build_pattern_primal(self.dn.net_ref(), ptr(self.rowptr), ptr(self.colidx),
                                                    current_stream()))
        return self.rowptr, self.colidx

    def assemble(self, co, out=None):
        self.build_pattern()
        vals = out if out is not None else self._f64(self.nnz)
        check(self.lib.gnx_assemble_primal(self.dn.net_ref(), co.ref(), ptr(vals), current_stream()))
        return vals

    def rhs(self, f=None, g=None, out=None):
        """HydraulicNetwork.L_form; f, g CompiledCoefficients (or None = 0)"""
        b = out if out is not None else self._f64(self.N)

        def prep(c, slot):
            if c is None:
                return None, 0, 0.0
            if c.is_constant:
                return None, 0, c.const_value()
            return self.eval_cellwise(c, slot), int(c.degree <= 1), 0.0
        fc, fl, f0 = prep(f, "f")
        gc, gl, g0 = prep(g, "g")
        check(self.lib.gnx_assemble_rhs_primal(self.dn.net_ref(), ptr(self.dn.cells), ptr(fc), fl, f0, ptr(gc), gl, g0,
                                               ptr(b), current_stream()))
        self._keep = (fc, gc)
        return b

    def bc_node_values(self, p_bc):
        """p_bc at the graph nodes [V] (persistent buffer)"""
        c = as_coefficient(p_bc)
        if getattr(self, "_bcv", None) is None:
            self._bcv = self._f64(self.dn.V)
        return c.eval_device(self.dn.pos, out=self._bcv)

    def apply_bc(self, A, b, bcs):
        """xii.apply_bc(A, b, bcs) (flow_models.py:117), in place on copies"""
        bc = bcs[0]
        g = self.bc_node_values(bc.value)
        sigma = A.coeffs.sigma if A is not None and A.coeffs is not None else 1.0
        vals = A.vals.clone() if A is not None else None
        bt = b.tensor.clone() if b is not None else None
        check(self.lib.gnx_apply_bc_primal(self.dn.net_ref(), sigma, ptr(g), ptr(vals), ptr(bt), current_stream()))
        A2 = AssembledMatrix(self, vals, A.coeffs, A.kind, bcs=bcs) if A is not None else None
        b2 = AssembledVector(self, bt) if b is not None else None
        return A2, b2

    def apply_bc_rhs(self, sigma, g_node, b):
        """Dirichlet data into a right-hand side, in place"""
        check(self.lib.gnx_apply_bc_primal(self.dn.net_ref(), float(sigma), ptr(g_node), None, ptr(b), current_stream()))
        return b

    # ---- building blocks ------------------------------------------------------------------------
    def spmv(self, vals, x, y=None):
        self.build_pattern()
        y = y if y is not None else self._f64(self.N)
        check(self.lib.gnx_spmv_csr_f64(self.N, ptr(self.rowptr), ptr(self.colidx), ptr(vals), ptr(x), ptr(y),
                                        current_stream()))
        return y

    def residual(self, vals, x, b, r=None):
        self.build_pattern()
        if self._red_work is None:
            self._red_work = self._f64(int(self.lib.gnx_reduce_work_doubles(self.N)))
        out2 = self._f64(2)
        check(self.lib.gnx_residual_csr_f64(self.N, ptr(self.rowptr), ptr(self.colidx), ptr(vals), ptr(x), ptr(b),
                                            ptr(r), ptr(out2), ptr(self._red_work), current_stream()))
        return out2

    def mass_apply(self, coeff, x, out=None):
        """y = diag(coeff_c h_c) x on the q rows (coeff: float or _PCoeffs)"""
        y = out if out is not None else self._f64(self.N)
        co = coeff if isinstance(coeff, _PCoeffs) else self.coeffs(coeff, 1.0)
        check(self.lib.gnx_mass_apply_q_primal(self.dn.net_ref(), co.ref(), ptr(x), ptr(y), current_stream()))
        return y

    def lincomb(self, coefs, vecs, out=None):
        return lincomb(self.lib, coefs, vecs, out)

    # ---- solve ----------------------------------------------------------------------------------
    def _bufs(self):
        if self._bufs_ is None:
            dn = self.dn
            E, B = dn.E, max(dn.B, 1)
            d = dict(P=torch.zeros(B, dtype=torch.float64, device=self.dev))
            if dn.part is None:
                d.update(cond=self._f64(E), rsrc=self._f64(E), ftot=self._f64(E))
            else:                       # same packed all-gather layout as the mixed engine
                d["pack"] = torch.zeros(3 * dn.chunk, dtype=torch.float64, device=self.dev)
                for i, k in enumerate(("cond", "rsrc", "ftot")):
                    d[k] = d["pack"][i * dn.chunk: i * dn.chunk + E]
                d["gathered"] = self._f64(dn.part[1] * 3 * dn.chunk)
            self._bufs_ = d
        return self._bufs_

    def schur_plan(self):
        if self._plan is None and self.dn.B > 0:
            from .engine import MixedSystem
            # the bifurcation graph is the same for both model families
            self._plan = MixedSystem.schur_plan(self)
        return self._plan

    def unconverged(self):
        from .engine import MixedSystem
        return MixedSystem.unconverged(self)

    def eliminate(self, co, b):
        d = self._bufs()
        check(self.lib.gnx_edge_eliminate_primal(self.dn.net_ref(), co.ref(), ptr(b), ptr(d["cond"]), ptr(d["rsrc"]),
                                                 ptr(d["ftot"]), current_stream()))

    def solve(self, co, b, x=None, rtol=1e-13, max_iter=100000, warm_start=False, info=None, sync=True, bcs=None,
              exchange=None, eliminated=False):
        """x = A^{-1} b; b is the right-hand side AFTER apply_bc (Dirichlet values sit in b)."""
        dn = self.dn
        d = self._bufs()
        info = info if info is not None else SolveInfo()
        x = x if x is not None else self._f64(self.N)
        st = current_stream()
        if not eliminated:
            self.eliminate(co, b)
        if dn.B > 0:
            plan, _, work, _ = self.schur_plan()
            cond, rsrc, ftot = d["cond"], d["rsrc"], d["ftot"]
            chunk = stride = 0
            if dn.part is not None:
                if exchange is not None:          # single-process emulation of the ranks (tests)
                    exchange(d)
                else:
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(d["gathered"], d["pack"], group=self.group)
                g = d["gathered"]
                chunk, stride = dn.chunk, 3 * dn.chunk
                cond, rsrc, ftot = g, g[chunk:], g[2 * chunk:]
            args = (dn.schur_net_ref(), C.byref(plan), co.sigma, ptr(cond), ptr(rsrc), ptr(ftot),
                    None, ptr(b[self.p_off:]), ptr(d["P"]), None, ptr(work), float(rtol), int(max_iter),
                    1 if warm_start else 0, chunk, stride, None)
            if sync:
                resid = C.c_double(0.0)
                inf = (C.c_int32 * 3)()
                check(self.lib.gnx_schur_solve(*args, C.byref(resid), inf, st))
                info.cg_iterations += int(inf[0])
                info.cg_converged = bool(inf[1])
                info.cg_batches += int(inf[2])
                info.cg_rel_resid = float(resid.value)
            else:
                check(self.lib.gnx_schur_solve(*args, None, None, st))
        check(self.lib.gnx_edge_backsub_primal(dn.net_ref(), co.ref(), ptr(b), ptr(d["cond"]), ptr(d["rsrc"]),
                                               ptr(d["P"]) if dn.B > 0 else None, ptr(x), st))
        return x, info

    def timestep(self, model, L, co, w1, w0, a_n, x0, Ln, mass_coeff, DDn, Ln1, b, ax, x1, bcs=None, warm_start=False,
                 sync=False, info=None, rtol=1e-13, max_iter=100000):
        """one theta-scheme step behind ONE library call (gnx_timestep_primal); f, g and the Dirichlet data are
        evaluated (device kernels) here, at assembly time, like PrimalLForm.assemble / solve_bc do"""
        dn = self.dn
        d = self._bufs()
        info = info if info is not None else SolveInfo()
        f, g = as_coefficient(model.f), as_coefficient(model.g)

        def prep(c, slot):
            if c.is_constant:
                return None, 0, c.const_value()
            return self.eval_cellwise(c, slot), int(c.degree <= 1), 0.0
        fc, fl, f0 = prep(f, "f")
        gc, gl, g0 = prep(g, "g")
        flat = [bc for row in (bcs or []) for bc in (row if isinstance(row, (list, tuple)) else [row])]
        g_node = self.bc_node_values(flat[0].value) if flat else None
        co_mass = mass_coeff if isinstance(mass_coeff, _PCoeffs) else self.coeffs(mass_coeff, 1.0)
        plan = work = None
        if dn.B > 0:
            plan, _, work, _ = self.schur_plan()
        resid = inf = None
        if sync and dn.B > 0:
            resid = C.c_double(0.0)
            inf = (C.c_int32 * 3)()
        cn = w0 != 0.0
        if cn:
            self.build_pattern()
        check(self.lib.gnx_timestep_primal(
            dn.net_ref(), co.ref(), C.byref(plan) if plan is not None else None, ptr(dn.cells), ptr(fc), fl, float(f0),
            ptr(gc), gl, float(g0), ptr(g_node), float(w1), float(w0), ptr(self.rowptr) if cn else None,
            ptr(self.colidx) if cn else None, ptr(a_n.vals) if cn else None, ptr(x0) if cn else None,
            ptr(Ln) if cn else None, co_mass.ref(), ptr(DDn), ptr(Ln1), ptr(b), ptr(ax) if cn else None, ptr(x1),
            ptr(d["cond"]), ptr(d["rsrc"]), ptr(d["ftot"]), ptr(d["P"]) if dn.B > 0 else None, ptr(work), float(rtol),
            int(max_iter), 1 if warm_start else 0, C.byref(resid) if resid is not None else None, inf, current_stream()))
        self._keep = (fc, gc, co_mass)
        if resid is not None:
            info.cg_iterations += int(inf[0])
            info.cg_converged = bool(inf[1])
            info.cg_batches += int(inf[2])
            info.cg_rel_resid = float(resid.value)
        return info

    def solve_bc(self, co, b, x, bcs=None, **kw):
        """apply the model's Dirichlet data to b (in place), then solve (time_stepping.py:188-192)"""
        flat = [bc for row in (bcs or []) for bc in (row if isinstance(row, (list, tuple)) else [row])]
        if flat:
            self.apply_bc_rhs(co.sigma, self.bc_node_values(flat[0].value), b)
        return self.solve(co, b, x=x, **kw)

    def solve_checked(self, co, vals_bc, b, rtol=1e-13, resid_tol=1e-10):
        x, info = self.solve(co, b, rtol=rtol)
        rr, bb = self.residual(vals_bc, x, b).cpu().tolist()
        info.rel_resid = float(np.sqrt(rr / bb)) if bb > 0 else float(np.sqrt(rr))
        return x, info


# ---------------------------------------------------------------------------------------------
# form handles
# ---------------------------------------------------------------------------------------------
class PrimalAForm:
    def __init__(self, model, sigma=1.0, mass=0.0):
        self.model, self.sigma, self.mass = model, sigma, mass

    def assemble(self):
        ps = self.model._system()
        co = self.model._coeffs(self.sigma, self.mass)
        return AssembledMatrix(ps, ps.assemble(co), co, "primal")


class PrimalLForm:
    """HydraulicNetwork.L_form (flow_models.py:94-101); evaluated when ASSEMBLED (no projections)."""

    def __init__(self, model):
        self.model = model

    def assemble(self):
        ps = self.model._system()
        f, g = as_coefficient(self.model.f), as_coefficient(self.model.g)
        return AssembledVector(ps, ps.rhs(f, g))


class PrimalMassForm:
    """TimeDepHydraulicNetwork.mass_matrix_q(coeff) (time_stepping.py:28-47)"""

    def __init__(self, model, coeff):
        self.model, self.coeff = model, coeff

    def assemble(self):
        ps = self.model._system()
        co = ps.coeffs(self.model._cell_coeff(self.coeff), 0.0)
        return AssembledMatrix(ps, ps.assemble(co), None, "mass")
