"""Host driver of the mixed (q per edge | p | lambda) system on one GPU: pattern (once),
assembly, right-hand side, Schur-complement solve, residual.  Thin orchestration over the
C ABI (include/graphnics_b200.h); no arithmetic happens here.

Replaces, for MixedHydraulicNetwork / NetworkStokes / TimeDepMixedHydraulicNetwork:
ii_assemble + ii_convert + LUSolver("mumps") (graphnics/flowmodels/flow_models.py:343-348).
"""
import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from ._lib import ptr, check
from .device import DeviceNetwork, current_stream

__all__ = ["MixedSystem", "SolveInfo"]


class SolveInfo:
    def __init__(self):
        self.cg_iterations = 0
        self.cg_converged = True
        self.cg_batches = 0
        self.cg_rel_resid = 0.0
        self.rel_resid = None      # ||b - A x|| / ||b|| of the monolithic system (if checked)
        self.refinements = 0

    def __repr__(self):
        return (f"SolveInfo(cg_iterations={self.cg_iterations}, cg_rel_resid={self.cg_rel_resid:.2e}, "
                f"rel_resid={self.rel_resid}, refinements={self.refinements})")


class _Coeffs:
    """gnx_mixed_coeffs_t + the tensors it points to."""

    def __init__(self, a_edge, k_edge, sigma, a_cell=None):
        self.a_edge, self.k_edge, self.sigma, self.a_cell = a_edge, k_edge, float(sigma), a_cell
        self.c = _lib.gnx_mixed_coeffs_t()
        self.c.a_edge = a_edge.data_ptr()
        self.c.k_edge = k_edge.data_ptr() if k_edge is not None else None
        self.c.sigma = float(sigma)
        self.c.a_cell = a_cell.data_ptr() if a_cell is not None else None

    def ref(self):
        return C.byref(self.c)


def lincomb(lib, coefs, vecs, out=None):
    terms = [(float(c), v) for c, v in zip(coefs, vecs) if v is not None and float(c) != 0.0]
    if not terms:
        terms = [(0.0, vecs[0])]
    k = len(terms)
    out = out if out is not None else torch.empty_like(terms[0][1])
    cs = (C.c_double * k)(*[c for c, _ in terms])
    ps = (C.c_void_p * k)(*[v.data_ptr() for _, v in terms])
    check(lib.gnx_lincomb(int(out.numel()), k, cs, ps, ptr(out), current_stream()))
    return out


class MixedSystem:
    def __init__(self, dn: DeviceNetwork, group=None, exchange="auto"):
        self.dn = dn
        self.group = group           # torch.distributed process group of a partitioned network
        # how a partitioned solve gathers the edge scalars: "p2p" = peer stores fused into the elimination
        # kernel (torch symmetric memory maps the peers' buffers), synchronised by flag words inside the Schur
        # kernel (gnx_exchange_t); "nccl" = one
        # all_gather_into_tensor; "auto" = p2p when symmetric memory can be set up, else nccl
        self.exchange_mode = exchange
        self._p2p = None
        self.lib = _lib.require_device()
        self.N = dn.mixed_ndofs
        self.nnz = dn.mixed_nnz
        self.dev = dn.device
        self.rowptr = None
        self.colidx = None
        self._plan = None
        self._solve_bufs = None
        self._red_work = None
        self.subtree = dn.part is not None and dn.edge_ids is not None     # gnx_exchange mode 1
        self._xch = None             # subtree mode: exchange buffers of all ranks + epoch

    # ---- helpers -------------------------------------------------------------------------
    def _f64(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.dev)

    def edge_array(self, value):
        """scalar | [E] array-like | tensor -> float64 device tensor [E]"""
        E = self.dn.E
        if torch.is_tensor(value):
            return value.to(device=self.dev, dtype=torch.float64).contiguous().reshape(E)
        # host data: uploaded on EVERY call (a user may have changed the attributes; and an end-to-end step includes
        # the host -> device copy of its inputs) through a pinned staging buffer, asynchronous on the current
        # stream (two buffers alternate; a buffer is reused only once the upload that last read it has run)
        st = self.__dict__.get("_edge_stage")
        if st is None:
            st = self._edge_stage = {"bufs": [torch.empty(E, dtype=torch.float64).pin_memory() for _ in range(2)],
                                     "evs": [torch.cuda.Event(), torch.cuda.Event()], "k": 0}
        k = st["k"] = st["k"] ^ 1
        st["evs"][k].synchronize()
        buf = st["bufs"][k]
        buf.numpy()[:] = np.asarray(value, dtype=np.float64)
        out = buf.to(self.dev, non_blocking=True)
        st["evs"][k].record()
        return out

    def coeffs(self, a_edge, k_edge=None, sigma=1.0, a_cell=None):
        """a_cell: [E*m] device tensor, mass coefficient per cell in edge-local order (overrides a_edge)"""
        a = self.edge_array(a_edge)
        k = None if k_edge is None else self.edge_array(k_edge)
        if a_cell is not None:
            a_cell = a_cell.to(device=self.dev, dtype=torch.float64).contiguous().reshape(self.dn.E * self.dn.m)
        return _Coeffs(a, k, sigma, a_cell)

    def cell_midpoints_edge_local(self):
        """[E*m, gdim] midpoints of the cells in EDGE-LOCAL order (cell k of edge e at e*m + k, u -> v) -- where a
        position-dependent edge coefficient is sampled"""
        if getattr(self, "_xmid_el", None) is None:
            dn = self.dn
            m, n = dn.m, dn.n
            k = np.arange(m, dtype=np.int64)
            edges = dn.edges_host()
            flip = edges[:, 0] > edges[:, 1]
            s_lo = np.where(flip[:, None], m - 1 - k[None, :], k[None, :])           # lo-frame position of cell k
            gcell = (np.arange(dn.E, dtype=np.int64)[:, None] * m + (s_lo ^ (s_lo >> 1))).ravel()
            c = dn.cells.long()[torch.from_numpy(gcell).to(self.dev)]
            self._xmid_el = (0.5 * (dn.coords[c[:, 0]] + dn.coords[c[:, 1]])).contiguous()
        return self._xmid_el

    def cell_values(self, per_edge):
        """per-edge coefficient objects (numbers, Constants, Expressions / UFL in x and parameters) -> [E*m] tensor of
        their values at the cell midpoints, edge-local order; one device evaluation per DISTINCT object"""
        from .coefficients import as_coefficient
        dn = self.dn
        out = self._f64(dn.E, dn.m)
        groups = {}
        for e, obj in enumerate(per_edge):
            groups.setdefault(id(obj), (obj, []))[1].append(e)
        xm = None
        for obj, es in groups.values():
            cc = as_coefficient(obj)
            idx = torch.as_tensor(es, device=self.dev)
            if cc.is_constant:
                out[idx] = cc.const_value()
                continue
            xm = self.cell_midpoints_edge_local() if xm is None else xm
            vals = cc.eval_device(xm, pts_per_edge=dn.m, tangents=dn.tangents).reshape(dn.E, dn.m)
            out[idx] = vals[idx]
        return out.reshape(-1)

    # ---- pattern (built once) ----------------------------------------------------------------
    def build_pattern(self):
        if self.rowptr is None:
            self.rowptr = torch.empty(self.N + 1, dtype=torch.int32, device=self.dev)
            self.colidx = torch.empty(self.nnz, dtype=torch.int32, device=self.dev)
            check(self.lib.gnx_build_pattern_mixed(self.dn.net_ref(), ptr(self.rowptr), ptr(self.colidx),
                                                   current_stream()))
        return self.rowptr, self.colidx

    # ---- assembly ----------------------------------------------------------------------------
    def assemble(self, co, out=None):
        self.build_pattern()
        vals = out if out is not None else self._f64(self.nnz)
        check(self.lib.gnx_assemble_mixed(self.dn.net_ref(), co.ref(), ptr(self.rowptr), ptr(vals),
                                          current_stream()))
        return vals

    def assemble_mass_q(self, coeff, out=None):
        self.build_pattern()
        vals = out if out is not None else self._f64(self.nnz)
        check(self.lib.gnx_assemble_mass_q_mixed(self.dn.net_ref(), float(coeff), ptr(self.rowptr),
                                                 ptr(vals), current_stream()))
        return vals

    def mass_vals(self, coeff):
        """values of coeff * blockdiag(M_e) on the pattern, cached per coefficient"""
        key = float(coeff)
        if getattr(self, "_mass_cache", None) is None or self._mass_cache[0] != key:
            self._mass_cache = (key, self.assemble_mass_q(key))
        return self._mass_cache[1]

    def rhs(self, f_nodal=None, g_nodal=None, pbc_out=None, pbc_node=None, f_const=0.0, g_const=0.0, out=None):
        """MixedHydraulicNetwork.L_form; see gnx_assemble_rhs_mixed."""
        b = out if out is not None else self._f64(self.N)
        check(self.lib.gnx_assemble_rhs_mixed(self.dn.net_ref(), ptr(f_nodal), ptr(g_nodal), float(f_const),
                                              float(g_const), ptr(pbc_out), ptr(pbc_node), ptr(b),
                                              current_stream()))
        return b

    def dof_coordinates(self, want_mid=True):
        """(xq [E(m+1), gdim], xmid [E m, gdim]) -- cached."""
        dn = self.dn
        if getattr(self, "_xq", None) is None or (want_mid and getattr(self, "_xmid", None) is None):
            xq = self._f64(dn.E * (dn.m + 1), dn.gdim)
            xm = self._f64(dn.E * dn.m, dn.gdim) if want_mid else None
            check(self.lib.gnx_dof_coordinates_mixed(dn.net_ref(), ptr(dn.coords), ptr(dn.cells), ptr(xq), ptr(xm),
                                                     current_stream()))
            self._xq = xq
            if want_mid:
                self._xmid = xm
        return self._xq, (self._xmid if want_mid else None)

    # ---- coefficients ---------------------------------------------------------------------------
    def project_coefficient(self, coef):
        """project(coef, CG1(submesh_i)) for all edges (flow_models.py:315-317) -> [E(m+1)] q-dof layout"""
        dn = self.dn
        xq, xm = self.dof_coordinates(want_mid=coef.degree >= 2)
        cn = coef.eval_device(xq, pts_per_edge=dn.m + 1, tangents=dn.tangents,
                              cell_index=None)
        cm = coef.eval_device(xm, pts_per_edge=dn.m, tangents=dn.tangents) if coef.degree >= 2 else None
        out = self._f64(dn.E * (dn.m + 1))
        work = self._f64(8 * dn.E * (dn.m + 1))
        check(self.lib.gnx_project_p1_edges(dn.net_ref(), ptr(cn), ptr(cm), ptr(out), ptr(work), current_stream()))
        return out

    def end_values(self, nodal):
        """value at the v-end of every edge of a q-dof-layout nodal field [E(m+1)] -> [E]"""
        dn = self.dn
        if getattr(self, "_lv", None) is None:
            edges = dn.edges_host()
            flip = edges[:, 0] > edges[:, 1]
            lv = np.where(flip, dn.loc_host[0], dn.loc_host[dn.m]).astype(np.int64)
            self._lv = torch.from_numpy(lv + np.arange(dn.E, dtype=np.int64) * (dn.m + 1)).to(self.dev)
        return nodal.reshape(-1)[self._lv].contiguous()

    def project_end_values(self, coef, window=40):
        """projected coefficient at the v-end of every edge (the BOUN_OUT term, flow_models.py:324)"""
        dn = self.dn
        out = self._f64(dn.E)
        if coef.user is not None:          # python callback: project everything, pick the ends
            return self.end_values(self.project_coefficient(coef))
        prog = coef.program()
        check(self.lib.gnx_project_end_values(dn.net_ref(), ptr(dn.coords), C.byref(prog), int(coef.degree),
                                              int(window), ptr(out), current_stream()))
        return out

    # ---- Krylov building blocks ----------------------------------------------------------------
    def spmv(self, vals, x, y=None):
        y = y if y is not None else self._f64(self.N)
        check(self.lib.gnx_spmv_csr_f64(self.N, ptr(self.rowptr), ptr(self.colidx), ptr(vals), ptr(x), ptr(y),
                                        current_stream()))
        return y

    def mass_apply(self, coeff, x, out=None):
        """y = coeff * blockdiag(M_e) x on the q rows, 0 elsewhere (mass_vector_q, time_stepping.py:103-123)"""
        y = out if out is not None else self._f64(self.N)
        check(self.lib.gnx_mass_apply_q_mixed(self.dn.net_ref(), float(coeff), ptr(x), ptr(y), current_stream()))
        return y

    def lincomb(self, coefs, vecs, out=None):
        """out = sum_k coefs[k] * vecs[k]  (<= 4 terms, one pass)"""
        return lincomb(self.lib, coefs, vecs, out)

    def residual(self, vals, x, b, r=None):
        """(||b - A x||^2, ||b||^2) as a 2-element device tensor; r optional output."""
        if self._red_work is None:
            self._red_work = self._f64(int(self.lib.gnx_reduce_work_doubles(self.N)))
        out2 = self._f64(2)
        check(self.lib.gnx_residual_csr_f64(self.N, ptr(self.rowptr), ptr(self.colidx), ptr(vals), ptr(x), ptr(b),
                                            ptr(r), ptr(out2), ptr(self._red_work), current_stream()))
        return out2

    # ---- solve ---------------------------------------------------------------------------------
    def _bufs(self):
        if self._solve_bufs is None:
            dn = self.dn
            E, B = dn.E, max(dn.B, 1)
            d = {}
            d["cond"], d["rsrc"], d["ftot"] = self._f64(E), self._f64(E), self._f64(E)
            d["P"] = torch.zeros(B, dtype=torch.float64, device=self.dev)
            if dn.part is not None and not self.subtree:
                # edge scalars of the WHOLE network, gathered from all ranks (equal chunks; the tail of
                # the last chunk is padding).  The per-rank source buffers are chunk-sized too.
                # Layout: per rank one block [cond | rsrc | ftot] of 3*chunk doubles -> ONE collective.
                world = dn.part[1]
                d["pack"] = torch.zeros(3 * dn.chunk, dtype=torch.float64, device=self.dev)
                for i, k in enumerate(("cond", "rsrc", "ftot")):
                    d[k] = d["pack"][i * dn.chunk: i * dn.chunk + E]
                d["gathered"] = self._f64(world * 3 * dn.chunk)
            self._solve_bufs = d
        return self._solve_bufs

    def schur_plan(self):
        """symbolic rake/core plan of the Schur system (host analysis once, arrays on the device)"""
        if self._plan is None and self.dn.B > 0:
            from .schur_plan import SchurPlan
            dn = self.dn
            bix_h = dn.bix.cpu().numpy()
            hp = getattr(dn, "host_plan", None)            # (the partitioner's plan of the whole network, if it made one)
            if hp is None:
                hp = SchurPlan(dn.schur_edges_host(), bix_h, dn.B, max_levels=_lib.MAX_LEVELS)
            view = None
            if getattr(self, "subtree", False):
                from .schur_plan import RankPlan, node_owners
                owner = node_owners(hp, dn.schur_edges_host(), bix_h, dn.epart)
                hp = view = RankPlan(hp, owner, dn.part[0], max_levels=_lib.MAX_LEVELS)
            agg = None
            if hp.nc > _lib.TWO_LEVEL_MIN_CORE and os.environ.get("GNX_SCHUR_TWO_LEVEL", "1") != "0":
                # large cyclic core: two-level PCG, aggregates cut by coordinate bisection of the node positions
                node_of = np.empty(dn.B, dtype=np.int64)
                node_of[bix_h[bix_h >= 0]] = np.nonzero(bix_h >= 0)[0]
                agg = hp.aggregate_core(dn.pos.cpu().numpy()[node_of], na=_lib.AGG_MAX, ns=_lib.SUB_PER_AGG)
            lvl_ptr, lvl_nodes = hp.level_arrays()
            keep = {}
            dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(self.dev)
            s = _lib.gnx_schur_plan_t()
            l_split = hp.split_level(_lib.TOP_MAX) if hp.nc == 0 else len(hp.levels)
            top_S = 0
            if hp.nc == 0 and _lib.TOP_MAX < dn.B <= 16 * _lib.TOP_MAX and os.environ.get("GNX_SCHUR_CLUSTER_TOP", "0") == "1":
                # EXPERIMENT (off by default): sweep every level of a pure tree in one cluster's distributed
                # shared memory.  Measured on B200 (tools/time_schur_tree.py) it LOSES to the default -- wide
                # levels on the cooperative grid, the last <= 2048 nodes in one CTA's shared memory -- 59.8 vs
                # 46.0 us at 16 k bifurcations: a DSMEM read (~215 cycles) + cluster barrier per level costs more
                # than the three grid barriers it saves, because the one-CTA top part is so cheap per level.
                l_split = 0
                top_S = 512
                while 16 * top_S < dn.B:
                    top_S *= 2
            top, top_of, level_of = hp.top_arrays(l_split)
            blocks = None
            n_swept = dn.B if view is None else view.n_mine
            if hp.nc == 0 and top_S == 0 and n_swept > _lib.TOP_MAX and dn.B > _lib.TOP_MAX and \
                    os.environ.get("GNX_SCHUR_BLOCKS", "1") != "0":
                # big forest: whole subtrees of <= 1024 nodes per CTA instead of one grid barrier per wide level
                blocks = (hp.block_arrays(cap=_lib.BLK_MAX, top_max=_lib.TOP_MAX) if view is None else
                          view.rank_block_arrays(cap=_lib.BLK_MAX, top_max=_lib.TOP_MAX))
                if blocks is not None:
                    top, top_of, level_of, l_split = blocks["top_nodes"], blocks["top_of"], blocks["level_of"], blocks["l_top"]
            s.B, s.nl, s.l_split, s.nc, s.nnzc, s.nt = dn.B, len(hp.levels), l_split, hp.nc, hp.nnzc, len(top)
            for i, v in enumerate(lvl_ptr):
                s.lvl_ptr[i] = int(v)
            ell_w, ell_R, ell_col, ell_edge = hp.ell_arrays()
            s.ell_w, s.ell_R, s.top_S = ell_w, ell_R, top_S
            s.na = agg["na"] if agg is not None else 0
            s.ns = agg["ns"] if agg is not None else 0
            extra = () if agg is None else (("agg_ptr", agg["agg_ptr"]), ("col_agg", agg["col_agg"]),
                                            ("cut_ptr", agg["cut_ptr"]), ("cut_k", agg["cut_k"]),
                                            ("sub_ptr", agg["sub_ptr"]), ("row_sub", agg["row_sub"]))
            s.nb = blocks["nb"] if blocks is not None else 0
            if blocks is not None:
                extra += tuple((k, blocks[k]) for k in ("blk_ptr", "blk_nodes", "blk_pos", "blk_lv"))
            hp.blocks = blocks
            s.l_sync, s.nxl = len(hp.levels), 0
            if len(top) and top_S == 0 and hp.nc == 0:
                from .schur_plan import top_records
                g = dn.glob if dn.glob is not None else dn
                extra += (("top_rec", top_records(hp, top, top_of, level_of, g.inc_ptr.cpu().numpy(), g.inc_edge.cpu().numpy(),
                                                  g._bif_nodes.cpu().numpy(), None if view is None else view.ncls).ravel()),)
            if view is not None:
                if hp.nc == 0 and l_split > view.l_sync:
                    raise ValueError("subtree partition: %d replicated bifurcations do not fit the top part (%d)"
                                     % (int(view.repl_raked.sum()), _lib.TOP_MAX))
                xlist = np.nonzero(view.ncls & 4)[0]
                s.l_sync, s.nxl = view.l_sync, len(xlist)
                extra += (("ncls", view.ncls), ("xlist", xlist))
                # local edge -> global id, bit 31: an end is a replicated bifurcation (the scalars cross NVLink)
                eg = dn.schur_edges_host()[dn.edge_ids]
                bu, bv = bix_h[eg[:, 0]], bix_h[eg[:, 1]]
                rep = (np.where(bu >= 0, view.owner[np.maximum(bu, 0)] < 0, False) |
                       np.where(bv >= 0, view.owner[np.maximum(bv, 0)] < 0, False))
                egl = dn.edge_ids.astype(np.int64) | np.where(rep, 1 << 31, 0)
                self._egl = torch.from_numpy(egl.astype(np.uint32).view(np.int32)).to(self.dev)
                self._xedges = dn.edge_ids[rep]            # what this rank exports (NCCL fallback packs these)
            for name, arr in extra + (("lvl_nodes", lvl_nodes), ("parent", hp.parent), ("pedge", hp.pedge),
                              ("ch_ptr", hp.ch_ptr), ("ch_idx", hp.ch_idx), ("core_nodes", hp.core_nodes),
                              ("crp", hp.crp), ("ccol", hp.ccol), ("cedge", hp.cedge),
                              ("top_nodes", top), ("top_of", top_of), ("level_of", level_of),
                              ("ell_col", ell_col), ("ell_edge", ell_edge)):
                t = dev(arr if len(arr) else np.zeros(1, dtype=np.int32))
                keep[name] = t
                setattr(s, name, t.data_ptr())
            work = torch.zeros(int(self.lib.gnx_schur_work_doubles(C.byref(s))), dtype=torch.float64, device=self.dev)
            self._plan = (s, keep, work, hp)
        return self._plan

    def unconverged(self):
        """solves since the last call whose Krylov part did not converge (asynchronous solves -- sync=False -- cannot
        raise themselves): one small device -> host read"""
        if self._plan is None or self.dn.B == 0:
            return 0
        plan, _, work, _ = self._plan
        return int(self.lib.gnx_schur_unconverged(C.byref(plan), ptr(work), current_stream()))

    def _p2p_setup(self):
        """two symmetric all-gather buffers (alternating per solve: a fast rank may already be writing
        the next solve's scalars into a peer that is still reading this solve's)"""
        if self._p2p is not None or self.exchange_mode == "nccl":
            return self._p2p
        try:
            import torch.distributed as dist
            import torch.distributed._symmetric_memory as symm
            dn = self.dn
            world = dn.part[1]
            if world > 8:
                raise RuntimeError("more than 8 peers")
            group = self.group if self.group is not None else dist.group.WORLD
            bufs, hdls, ptrs = [], [], []
            for _ in range(2):
                # world blocks [cond | rsrc | ftot] + world 64-bit flag words (include/graphnics_b200.h: gnx_exchange)
                t = symm.empty(world * 3 * dn.chunk + world, dtype=torch.float64, device=self.dev)
                t.zero_()
                torch.cuda.synchronize(self.dev)           # flags are zero before any peer can reach them
                h = symm.rendezvous(t, group)
                bufs.append(t); hdls.append(h)
                ptrs.append((C.c_void_p * world)(*[int(p) for p in h.buffer_ptrs]))
            for h in hdls:
                h.barrier(channel=0)
            torch.cuda.synchronize(self.dev)
            self._p2p = dict(bufs=bufs, hdls=hdls, ptrs=ptrs, step=0,
                             flags=os.environ.get("GNX_P2P_BARRIER", "0") != "1")
            self.exchange_mode = "p2p"
        except Exception as ex:          # no peer access / symmetric memory: the NCCL collective still works
            if self.exchange_mode == "p2p":
                raise
            self._p2p_error = repr(ex)
            self.exchange_mode = "nccl"
            self._p2p = None
        return self._p2p

    def _exchange(self, p2p):
        """descriptor of this solve's exchange (gnx_exchange_t): alternating gather buffer, next epoch"""
        k = p2p["step"] & 1
        p2p["step"] += 1
        rank, world = self.dn.part
        ex = _lib.gnx_exchange_t()
        ex.world, ex.rank, ex.chunk = world, rank, self.dn.chunk
        ex.peers_host = p2p["ptrs"][k]
        ex.epoch = p2p["step"] if p2p["flags"] else 0      # 0: separate barrier kernel (GNX_P2P_BARRIER=1)
        return k, ex

    # ---- subtree-partitioned solve (gnx_exchange mode 1) --------------------------------------
    def xch_words(self):
        """doubles of one exchange buffer: [cond | rsrc | ftot (E_glob each) | d | r (B each) | world flag words]"""
        dn = self.dn
        return 3 * dn.E_glob + 2 * dn.B + dn.part[1]

    def xch_attach(self, slots):
        """ranks emulated in ONE process (tests): slots[k][r] = exchange buffer k of rank r on this device"""
        world = self.dn.part[1]
        ptrs = [(C.c_void_p * world)(*[int(t.data_ptr()) for t in slot]) for slot in slots]
        self._xch = dict(bufs=[slot[self.dn.part[0]] for slot in slots], ptrs=ptrs, step=0, mode="emulated", keep=slots)

    def _xch_setup(self):
        """two exchange buffers (alternating per solve) mapped into every rank: torch symmetric memory; without
        peer access each rank keeps plain buffers and the exported values travel by one NCCL all-gather
        between the two halves of the Schur solve"""
        if self._xch is not None:
            return self._xch
        import torch.distributed as dist
        dn = self.dn
        rank, world = dn.part
        group = self.group if self.group is not None else dist.group.WORLD
        nw = self.xch_words()
        try:
            if self.exchange_mode == "nccl" or world > 8:
                raise RuntimeError("NCCL exchange requested")
            import torch.distributed._symmetric_memory as symm
            bufs, hdls, ptrs = [], [], []
            for _ in range(2):
                t = symm.empty(nw, dtype=torch.float64, device=self.dev)
                t.zero_()
                # the (d, r) slots of exported subtree roots validate themselves (GNX_XSENT, gnx_solve.cu)
                t[3 * dn.E_glob: 3 * dn.E_glob + 2 * dn.B].view(torch.int64).fill_(0x7FF85EA71E550001)
                torch.cuda.synchronize(self.dev)           # flags are zero before any peer can reach them
                h = symm.rendezvous(t, group)
                bufs.append(t); hdls.append(h)
                ptrs.append((C.c_void_p * world)(*[int(p) for p in h.buffer_ptrs]))
            for h in hdls:
                h.barrier(channel=0)
            torch.cuda.synchronize(self.dev)
            self._xch = dict(bufs=bufs, hdls=hdls, ptrs=ptrs, step=0, mode="p2p")
            self.exchange_mode = "p2p"
        except Exception as ex:
            if self.exchange_mode == "p2p":
                raise
            self._p2p_error = repr(ex)
            self.exchange_mode = "nccl"
            bufs = [torch.zeros(nw, dtype=torch.float64, device=self.dev) for _ in range(2)]
            ptrs = [(C.c_void_p * 1)(int(t.data_ptr())) for t in bufs]
            # positions (in a buffer) of what every rank exports: its edges at replicated bifurcations, its roots
            plan = self.schur_plan()[3]
            owner, E, B = plan.owner, dn.E_glob, dn.B
            eg = dn.schur_edges_host()
            bix_h = dn.bix.cpu().numpy()
            bu, bv = bix_h[eg[:, 0]], bix_h[eg[:, 1]]
            rep = (np.where(bu >= 0, owner[np.maximum(bu, 0)] < 0, False) | np.where(bv >= 0, owner[np.maximum(bv, 0)] < 0, False))
            par = plan.parent.astype(np.int64)
            root = (owner >= 0) & np.where(par >= 0, owner[np.maximum(par, 0)] < 0, False)
            pos = []
            for q in range(world):
                xe = np.nonzero(rep & (dn.epart == q))[0]
                xn = np.nonzero(root & (owner == q))[0]
                pos.append(np.concatenate([xe, E + xe, 2 * E + xe, 3 * E + xn, 3 * E + B + xn]).astype(np.int64))
            L = max(len(p_) for p_ in pos)
            self._xch = dict(bufs=bufs, ptrs=ptrs, step=0, mode="nccl", group=group,
                             pos=[torch.from_numpy(p_).to(self.dev) for p_ in pos], L=max(L, 1),
                             send=torch.zeros(max(L, 1), dtype=torch.float64, device=self.dev),
                             recv=torch.zeros(world * max(L, 1), dtype=torch.float64, device=self.dev))
        return self._xch

    def _xch_desc(self, xc, k, phase):
        dn = self.dn
        ex = _lib.gnx_exchange_t()
        if xc["mode"] == "nccl":
            ex.world, ex.rank = 1, 0
        else:
            ex.rank, ex.world = dn.part
        ex.chunk, ex.mode = 0, 1
        ex.peers_host = xc["ptrs"][k]
        ex.epoch = xc["step"]
        ex.egl = self._egl.data_ptr()
        ex.E_glob, ex.B_glob, ex.phase = dn.E_glob, dn.B, phase
        return ex

    def _xch_nccl(self, xc, buf):
        """the exchange by collective: everybody's exported values into everybody's buffer"""
        import torch.distributed as dist
        rank, world = self.dn.part
        L = xc["L"]
        mine = xc["pos"][rank]
        xc["send"][: len(mine)] = buf[mine]
        dist.all_gather_into_tensor(xc["recv"], xc["send"], group=xc["group"])
        for q in range(world):
            if q != rank and len(xc["pos"][q]):
                buf[xc["pos"][q]] = xc["recv"][q * L: q * L + len(xc["pos"][q])]

    def _solve_subtree(self, co, b, x, rtol, max_iter, warm_start, info, sync, cr, xphase):
        """local elimination + local rake -> ONE exchange -> replicated top -> local down-sweep + back-substitution.
        xphase None: the whole solve (flags inside the Schur kernel over NVLink peer mappings, or NCCL between
        two kernel launches); 1 / 2: the two halves on their own (ranks emulated in one process)."""
        dn = self.dn
        d = self._bufs()
        st = current_stream()
        plan, _, work, _ = self.schur_plan()
        xc = self._xch_setup()
        if xphase in (None, 1):
            xc["step"] += 1
        k = (xc["step"] - 1) & 1
        buf = xc["bufs"][k]
        E_g = dn.E_glob
        lam = x[dn.lam_off:]
        phases = ([0] if xc["mode"] == "p2p" else [1, 2]) if xphase is None else [xphase]
        for ph in phases:
            ex = self._xch_desc(xc, k, ph)
            if ph in (0, 1):
                if cr is not None:
                    check(self.lib.gnx_rhs_eliminate_mixed(dn.net_ref(), co.ref(), *cr, ptr(d["cond"]), ptr(d["rsrc"]),
                                                           ptr(d["ftot"]), C.byref(ex), st))
                else:
                    check(self.lib.gnx_edge_eliminate_mixed_p2p(dn.net_ref(), co.ref(), ptr(b), C.byref(ex), ptr(d["cond"]),
                                                                ptr(d["rsrc"]), st))
            elif xc["mode"] == "nccl" and xphase is None:
                self._xch_nccl(xc, buf)
            args = (dn.schur_net_ref(), C.byref(plan), co.sigma, ptr(buf), ptr(buf[E_g:]), ptr(buf[2 * E_g:]),
                    ptr(b[dn.lam_off:]), None, ptr(d["P"]), ptr(lam), ptr(work), float(rtol), int(max_iter),
                    1 if warm_start else 0, 0, 0, C.byref(ex))
            if sync and ph != 1:
                resid = C.c_double(0.0)
                inf = (C.c_int32 * 3)()
                check(self.lib.gnx_schur_solve(*args, C.byref(resid), inf, st))
                info.cg_iterations += int(inf[0])
                info.cg_converged = bool(inf[1])
                info.cg_batches += int(inf[2])
                info.cg_rel_resid = float(resid.value)
            else:
                check(self.lib.gnx_schur_solve(*args, None, None, st))
            if ph == 1:
                continue
            if cr is not None:
                check(self.lib.gnx_edge_backsub_mixed_const(dn.net_ref(), co.ref(), *cr[:4], ptr(b), ptr(d["cond"]),
                                                            ptr(d["rsrc"]), ptr(d["P"]), ptr(x), st))
            else:
                check(self.lib.gnx_edge_backsub_mixed(dn.net_ref(), co.ref(), ptr(b), ptr(d["cond"]), ptr(d["rsrc"]),
                                                      ptr(d["P"]), ptr(x), st))
        return x, info

    def eliminate(self, co, b):
        """phase 1: per-edge scans -> edge scalars (device only, no sync)."""
        d = self._bufs()
        check(self.lib.gnx_edge_eliminate_mixed(self.dn.net_ref(), co.ref(), ptr(b), ptr(d["cond"]), ptr(d["rsrc"]),
                                                ptr(d["ftot"]), current_stream()))

    def schur_arrays(self, co, b):
        """explicit Schur diagonal / off-diagonals / rhs (inspection and tests)"""
        dn = self.dn
        d = self._bufs()
        self.eliminate(co, b)
        out = dict(sdiag=self._f64(dn.B), srhs=self._f64(dn.B), soff=self._f64(2 * dn.E),
                   scol=torch.empty(2 * dn.E, dtype=torch.int32, device=self.dev))
        check(self.lib.gnx_schur_build(dn.net_ref(), co.sigma, ptr(d["cond"]), ptr(d["rsrc"]), ptr(d["ftot"]),
                                       ptr(b[dn.lam_off:]), ptr(out["sdiag"]), ptr(out["soff"]), ptr(out["scol"]),
                                       ptr(out["srhs"]), current_stream()))
        return out

    def solve(self, co, b, x=None, rtol=1e-13, max_iter=100000, warm_start=False, info=None, sync=True,
              exchange=None, eliminated=False, const_rhs=None, xphase=None):
        """x = A^{-1} b for the operator described by `co` (never forms A): edge elimination ->
        one persistent Schur kernel (rake + core PCG) -> back-substitution.
        sync=False: nothing returns to the host; `info` is not filled.
        Partitioned network: between elimination and the Schur kernel the edge scalars of all ranks
        are all-gathered (`exchange` / `eliminated` let a test play all ranks in one process).
        const_rhs = dict(f_const, g_const, pbc_out, pbc_node): b is PRODUCED (L_form with Constant f, g)
        by the elimination kernel instead of being read by it (gnx_rhs_eliminate_mixed)."""
        dn = self.dn
        d = self._bufs()
        info = info if info is not None else SolveInfo()
        x = x if x is not None else self._f64(self.N)
        st = current_stream()
        if self.subtree and dn.B > 0:
            if const_rhs is not None and dn.n > 10:
                self.rhs(out=b, **const_rhs)
                const_rhs = None
            cr = None
            if const_rhs is not None:
                cr = (float(const_rhs.get("f_const", 0.0)), float(const_rhs.get("g_const", 0.0)),
                      ptr(const_rhs.get("pbc_out")), ptr(const_rhs.get("pbc_node")), ptr(b))
            return self._solve_subtree(co, b, x, rtol, max_iter, warm_start, info, sync, cr, xphase)
        p2p = None
        if dn.part is not None and exchange is None and not eliminated and dn.B > 0:
            p2p = self._p2p_setup()
        loc_cond, loc_rsrc = d["cond"], d["rsrc"]
        if const_rhs is not None and (dn.n > 10 or eliminated):
            self.rhs(out=b, **const_rhs)           # no fused kernel for this case: plain right-hand side first
            const_rhs = None
        cr = None
        if const_rhs is not None:
            cr = (float(const_rhs.get("f_const", 0.0)), float(const_rhs.get("g_const", 0.0)),
                  ptr(const_rhs.get("pbc_out")), ptr(const_rhs.get("pbc_node")), ptr(b))
        ex = None
        if p2p is not None:
            # elimination with the all-gather fused in: peer stores over NVLink + one flag word per rank that
            # the Schur kernel waits for (or, GNX_P2P_BARRIER=1, a separate cross-GPU barrier kernel)
            k, ex = self._exchange(p2p)
            rank, world = dn.part
            if cr is not None:
                check(self.lib.gnx_rhs_eliminate_mixed(dn.net_ref(), co.ref(), *cr, None, None, None, C.byref(ex), st))
            else:
                check(self.lib.gnx_edge_eliminate_mixed_p2p(dn.net_ref(), co.ref(), ptr(b), C.byref(ex), None, None, st))
            if ex.epoch == 0:
                p2p["hdls"][k].barrier(channel=0)
            g = p2p["bufs"][k]
            loc_cond = g[rank * 3 * dn.chunk:]
            loc_rsrc = g[rank * 3 * dn.chunk + dn.chunk:]
        elif cr is not None:
            check(self.lib.gnx_rhs_eliminate_mixed(dn.net_ref(), co.ref(), *cr, ptr(d["cond"]), ptr(d["rsrc"]),
                                                   ptr(d["ftot"]), None, st))
        elif not eliminated:
            self.eliminate(co, b)
        if dn.B > 0:
            plan, _, work, _ = self.schur_plan()
            lam = x[dn.lam_off:]
            cond, rsrc, ftot = d["cond"], d["rsrc"], d["ftot"]
            chunk = stride = 0
            if p2p is not None:
                chunk, stride = dn.chunk, 3 * dn.chunk
                cond, rsrc, ftot = g, g[chunk:], g[2 * chunk:]
            elif dn.part is not None:
                # the only exchange of the solve: every rank needs the edge scalars of all edges
                # (3 doubles per edge); the Schur solve itself is replicated, so no further traffic
                if exchange is not None:          # single-process emulation of the ranks (tests)
                    exchange(d)
                else:
                    import torch.distributed as dist
                    dist.all_gather_into_tensor(d["gathered"], d["pack"], group=self.group)
                g = d["gathered"]
                chunk, stride = dn.chunk, 3 * dn.chunk
                cond, rsrc, ftot = g, g[chunk:], g[2 * chunk:]
            args = (dn.schur_net_ref(), C.byref(plan), co.sigma, ptr(cond), ptr(rsrc), ptr(ftot),
                    ptr(b[dn.lam_off:]), None, ptr(d["P"]), ptr(lam), ptr(work), float(rtol), int(max_iter),
                    1 if warm_start else 0, chunk, stride, C.byref(ex) if ex is not None and ex.epoch > 0 else None)
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
        if cr is not None and not eliminated:
            check(self.lib.gnx_edge_backsub_mixed_const(dn.net_ref(), co.ref(), *cr[:4], ptr(b), ptr(loc_cond), ptr(loc_rsrc),
                                                        ptr(d["P"]) if dn.B > 0 else None, ptr(x), st))
        else:
            check(self.lib.gnx_edge_backsub_mixed(dn.net_ref(), co.ref(), ptr(b), ptr(loc_cond), ptr(loc_rsrc),
                                                  ptr(d["P"]) if dn.B > 0 else None, ptr(x), st))
        return x, info

    def assemble_rhs_solve(self, co, vals, b, x, rhs_kwargs=None, overlap=True, **solve_kw):
        """One whole `MixedHydraulicNetwork.solve()` step (flow_models.py:337-350): CSR values, right-hand
        side, solve.  The solver never reads the assembled values, so with overlap=True the assembly
        kernel runs on the caller's stream while right-hand side -> elimination -> Schur kernel ->
        back-substitution run on a HIGH-PRIORITY side stream: the Schur phase is latency-bound and
        occupies one CTA (or one cluster, or a cross-GPU barrier when partitioned) and the
        bandwidth-bound assembly fills the SMs it leaves idle.  Both streams are joined before return
        (stream-ordered; no host synchronisation unless solve_kw asks for solver info)."""
        rhs_kwargs = rhs_kwargs or {}
        fused = rhs_kwargs.get("f_nodal") is None and rhs_kwargs.get("g_nodal") is None and self.dn.n <= 10
        if fused:                                  # Constant f, g: the elimination kernel produces b itself
            solve_kw = dict(solve_kw, const_rhs={k: v for k, v in rhs_kwargs.items() if k not in ("f_nodal", "g_nodal")})
        if fused and x is not None and not solve_kw.get("eliminated") and solve_kw.get("exchange") is None \
                and solve_kw.get("xphase") is None:
            p2p = None
            if self.subtree and self.dn.B > 0:
                xc = self._xch_setup()
                if xc["mode"] == "p2p":
                    return self._step_native(co, vals, b, x, solve_kw.pop("const_rhs"), overlap, xch=xc, **solve_kw)
            elif self.dn.part is not None and self.dn.B > 0:
                p2p = self._p2p_setup()
            if self.dn.part is None or (p2p is not None and p2p["flags"]):
                return self._step_native(co, vals, b, x, solve_kw.pop("const_rhs"), overlap, p2p=p2p, **solve_kw)
        if not overlap:
            self.assemble(co, out=vals)
            if not fused:
                self.rhs(out=b, **rhs_kwargs)
            return self.solve(co, b, x=x, **solve_kw)
        cur = torch.cuda.current_stream(self.dev)
        if getattr(self, "_hi_stream", None) is None:
            self._hi_stream = torch.cuda.Stream(device=self.dev, priority=-1)
            self._ev = (torch.cuda.Event(), torch.cuda.Event())
        hi, (ev0, ev1) = self._hi_stream, self._ev
        ev0.record(cur)
        hi.wait_event(ev0)
        with torch.cuda.stream(hi):
            if not fused:
                self.rhs(out=b, **rhs_kwargs)
            out = self.solve(co, b, x=x, **solve_kw)
            ev1.record(hi)
        self.assemble(co, out=vals)
        cur.wait_event(ev1)
        return out

    def _step_native(self, co, vals, b, x, const_rhs, overlap, rtol=1e-13, max_iter=100000, warm_start=False,
                     info=None, sync=True, p2p=None, xch=None, **_):
        """the whole step behind ONE library call (gnx_step_mixed): same kernels, streams and events as the
        Python-sequenced path above, a fraction of its host time"""
        dn = self.dn
        d = self._bufs()
        info = info if info is not None else SolveInfo()
        if getattr(self, "_hi_stream", None) is None:
            self._hi_stream = torch.cuda.Stream(device=self.dev, priority=-1)
            self._ev = (torch.cuda.Event(), torch.cuda.Event())
        if getattr(self, "_ev_ptr", None) is None:
            for ev in self._ev:
                ev.record()                                # a torch event owns a CUDA event from its first record on
            self._ev_ptr = tuple(C.c_void_p(ev.cuda_event) for ev in self._ev)
            self._hi_ptr = C.c_void_p(self._hi_stream.cuda_stream)
        plan = work = None
        if dn.B > 0:
            plan, _, work, _ = self.schur_plan()
        if getattr(self, "_step_sync", None) is None:      # arrival counter + epoch word of the persistent step kernel
            self._step_sync = torch.zeros(64, dtype=torch.int64, device=self.dev)
            self._step_epoch = 0
        self._step_epoch += 1
        resid = inf = None
        if sync and dn.B > 0:
            resid = C.c_double(0.0)
            inf = (C.c_int32 * 3)()
        ex = None
        if p2p is not None:
            _, ex = self._exchange(p2p)
        if xch is not None:
            xch["step"] += 1
            ex = self._xch_desc(xch, (xch["step"] - 1) & 1, 0)
        check(self.lib.gnx_step_mixed(
            dn.net_ref(), co.ref(), ptr(self.rowptr), ptr(vals), float(const_rhs.get("f_const", 0.0)),
            float(const_rhs.get("g_const", 0.0)), ptr(const_rhs.get("pbc_out")), ptr(const_rhs.get("pbc_node")), ptr(b),
            ptr(x), dn.schur_net_ref() if dn.B > 0 else None, C.byref(plan) if plan is not None else None,
            ptr(d["cond"]), ptr(d["rsrc"]), ptr(d["ftot"]), ptr(d["P"]) if dn.B > 0 else None, ptr(work), float(rtol),
            int(max_iter), 1 if warm_start else 0, C.byref(ex) if ex is not None else None,
            C.byref(resid) if resid is not None else None, inf,
            self._ev_ptr[0], self._ev_ptr[1], self._hi_ptr if overlap else None,
            ptr(self._step_sync), self._step_epoch, current_stream()))
        if resid is not None:
            info.cg_iterations += int(inf[0])
            info.cg_converged = bool(inf[1])
            info.cg_batches += int(inf[2])
            info.cg_rel_resid = float(resid.value)
        return x, info

    def timestep(self, model, L, co, w1, w0, a_n, x0, Ln, mass_coeff, DDn, Ln1, b, ax, x1, bcs=None, warm_start=False,
                 sync=False, info=None, rtol=1e-13, max_iter=100000):
        """one theta-scheme step behind ONE library call (gnx_timestep_mixed): L^{n+1} -> b -> solve -> D x^{n+1};
        `L`: the model's L_form handle, `a_n`: the assembled operator of step n (Crank-Nicolson) or None"""
        dn = self.dn
        d = self._bufs()
        info = info if info is not None else SolveInfo()
        kw = L.rhs_kwargs()
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
        check(self.lib.gnx_timestep_mixed(
            dn.net_ref(), co.ref(), C.byref(plan) if plan is not None else None, ptr(kw.get("f_nodal")), ptr(kw.get("g_nodal")),
            float(kw.get("f_const", 0.0)), float(kw.get("g_const", 0.0)), ptr(kw.get("pbc_out")), ptr(kw.get("pbc_node")),
            float(w1), float(w0), ptr(self.rowptr) if cn else None, ptr(self.colidx) if cn else None,
            ptr(a_n.vals) if cn else None, ptr(x0) if cn else None, ptr(Ln) if cn else None, float(mass_coeff), ptr(DDn),
            ptr(Ln1), ptr(b), ptr(ax) if cn else None, ptr(x1), ptr(d["cond"]), ptr(d["rsrc"]), ptr(d["ftot"]),
            ptr(d["P"]) if dn.B > 0 else None, ptr(work), float(rtol), int(max_iter), 1 if warm_start else 0,
            C.byref(resid) if resid is not None else None, inf, current_stream()))
        if resid is not None:
            info.cg_iterations += int(inf[0])
            info.cg_converged = bool(inf[1])
            info.cg_batches += int(inf[2])
            info.cg_rel_resid = float(resid.value)
        return info

    def solve_bc(self, co, b, x, bcs=None, **kw):
        """solve with the model's Dirichlet data applied (the mixed model has none, flow_models.py:334)"""
        return self.solve(co, b, x=x, **kw)

    def solve_checked(self, co, vals, b, rtol=1e-13, resid_tol=1e-10, max_refine=2):
        """solve + monolithic residual check with (optional) iterative refinement."""
        x, info = self.solve(co, b, rtol=rtol)
        r = self._f64(self.N)
        for it in range(max_refine + 1):
            rr, bb = self.residual(vals, x, b, r).cpu().tolist()
            info.rel_resid = float(np.sqrt(rr / bb)) if bb > 0 else float(np.sqrt(rr))
            if info.rel_resid <= resid_tol or it == max_refine:
                break
            dx, _ = self.solve(co, r, rtol=rtol, info=info)
            check(self.lib.gnx_axpby_dot(self.N, 1.0, ptr(dx), 1.0, ptr(x), None, None, current_stream()))
            info.refinements += 1
        return x, info

    def close(self):
        self._plan = None
        self._solve_bufs = None
