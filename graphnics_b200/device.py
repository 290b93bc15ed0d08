"""Device-resident network state: everything `make_mesh(n)` + `make_submeshes()` leave on a
FenicsGraph (graphnics/fenics_graph.py:102-156), as CUDA arrays produced by the C-ABI kernels.

torch is used for device memory and streams only (plumbing); all arithmetic happens in
libgraphnics_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import ptr, check

__all__ = ["submesh_numbering", "DeviceNetwork", "current_stream"]


def current_stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def submesh_numbering(n):
    """SubMesh-local vertex numbering pattern for n bisections (same for every edge).

    Positions t = 0..m are counted from the edge's lower-numbered node.  DOLFIN's SubMesh
    (via xii.EmbeddedMesh, fenics_graph.py:139) numbers vertices by FIRST ENCOUNTER while
    walking the edge's cells in parent-cell order; a cell's stored vertex pair is ascending in
    global id, i.e. the coarser (even-t) endpoint first.  Local cell j is the cell with spatial
    index invgray(j).  Returns (loc[t] -> local id, tloc[local id] -> t) as int32 arrays (m+1).
    """
    m = 1 << n
    j = np.arange(m, dtype=np.int64)
    s = j.copy()
    sh = 1
    while sh < 64:
        s ^= s >> sh
        sh <<= 1
    even = np.where(s % 2 == 0, s, s + 1)
    odd = np.where(s % 2 == 0, s + 1, s)
    walk = np.stack([even, odd], axis=1).ravel()
    uniq, first = np.unique(walk, return_index=True)
    tloc = uniq[np.argsort(first, kind="stable")]
    loc = np.empty(m + 1, dtype=np.int64)
    loc[tloc] = np.arange(m + 1)
    return loc.astype(np.int32), tloc.astype(np.int32)


def row_table(n):
    """Per-position table of the mixed q-rows (the same for every edge), [m+1, 2] int32:
      [t, 0] = CSR offset of the row of position t inside its edge's range, assuming neither end is a
               bifurcation: 5*loc[t] - 2*(loc[t] > 0) - 2*(loc[t] > loc[m]);
      [t, 1] = bit 0: loc[t] > 0, bit 1: loc[t] > loc[m] (add one slot each if the lo / hi end carries a
               multiplier), bits 2-4 / 5-7 / 8-10: slot of the left-neighbour / diagonal / right-neighbour
               q entry, bits 11-13 / 14-16: slot of the p entry of cell t-1 / cell t, bit 17: loc[t] < loc[t+1]
               (order of the two q columns in the p-row of cell t).
    Columns inside a row are sorted (q columns by SubMesh-local number, then p columns by global cell)."""
    m = 1 << n
    loc, _ = submesh_numbering(n)
    loc = loc.astype(np.int64)
    t = np.arange(m + 1)
    lm = loc[m]
    hasl, hasr = t > 0, t < m
    ll = np.where(hasl, loc[np.maximum(t - 1, 0)], -1)
    lr = np.where(hasr, loc[np.minimum(t + 1, m)], -1)
    l = loc
    gray = lambda x: x ^ (x >> 1)
    nq = 1 + hasl.astype(np.int64) + hasr
    pl = (hasl & (ll > l)).astype(np.int64) + (hasl & hasr & (ll > lr))
    pc = (hasl & (l > ll)).astype(np.int64) + (hasr & (l > lr))
    pr = (hasr & (lr > l)).astype(np.int64) + (hasl & hasr & (lr > ll))
    swp = hasl & hasr & (gray(np.maximum(t - 1, 0)) > gray(t))
    ppl = nq + swp
    ppr = nq + (hasl & ~swp)
    up = hasr & (l < lr)
    f0, f1 = l > 0, l > lm
    A = 5 * l - 2 * f0 - 2 * f1
    Bf = (f0.astype(np.int64) | (f1.astype(np.int64) << 1) | (pl << 2) | (pc << 5) | (pr << 8) | (ppl << 11)
          | (ppr << 14) | (up.astype(np.int64) << 17))
    return np.ascontiguousarray(np.stack([A, Bf], axis=1).astype(np.int32))


class _Tables:
    """output of gnx_vertex_tables for one edge list (fenics_graph.py:161-184, :132-156, :280-311)"""

    def __init__(self, lib, edges_t, V, dev, cls_deg=None):
        E = int(edges_t.shape[0])
        i32 = dict(dtype=torch.int32, device=dev)
        self.deg = torch.empty(V, **i32)
        self.bix = torch.empty(V, **i32)
        self._bif_nodes = torch.empty(V, **i32)
        self._bnd_nodes = torch.empty(V, **i32)
        self.inc_ptr = torch.empty(V + 1, **i32)
        self.inc_edge = torch.empty(2 * E, **i32)
        self.lam_ptr = torch.empty(V + 1, **i32)
        self.ebif_prefix = torch.empty(E + 1, **i32)
        self.tag_u = torch.empty(E, **i32)
        self.tag_v = torch.empty(E, **i32)
        counts = torch.empty(4, **i32)
        self._work = torch.empty(int(lib.gnx_tables_work_ints(V, E)), **i32)
        check(lib.gnx_vertex_tables(ptr(edges_t), V, E, ptr(self.deg), ptr(self.bix), ptr(self._bif_nodes),
                                    ptr(self._bnd_nodes), ptr(self.inc_ptr), ptr(self.inc_edge), ptr(self.lam_ptr),
                                    ptr(self.ebif_prefix), ptr(self.tag_u), ptr(self.tag_v), ptr(counts),
                                    ptr(self._work), ptr(cls_deg), current_stream()))
        self._counts = counts

    def counts(self):
        return [int(c) for c in self._counts.cpu().tolist()]      # one small D2H sync


def partition_range(E, rank, world):
    """contiguous, equal-chunk edge range of `rank` (chunk = ceil(E / world); the last ranks may be short)"""
    chunk = -(-E // world)
    e0 = min(E, rank * chunk)
    return e0, min(E, e0 + chunk), chunk


class DeviceNetwork:
    """Graph -> mesh on the GPU.  `pos` [V,gdim] float64, `edges` [E,2] ints (host or device).

    part=(rank, world): this object holds one rank's contiguous EDGE RANGE of a larger network
    (SURVEY.md 8e): mesh, cell lengths, incidence lists and dof layout cover the local edges only
    (the rank's subdomain), nodes are classified by their degree in the whole network, and
    `self.glob` keeps the whole network's tables for the (replicated) Schur solve.
    """

    def __init__(self, pos, edges, n, device=None, with_mf=True, part=None, epart=None):
        lib = _lib.require_device()
        self.device = torch.device(device if device is not None else "cuda")
        dev = self.device
        pos_t = torch.as_tensor(np.ascontiguousarray(pos, dtype=np.float64) if not torch.is_tensor(pos) else pos)
        edg_t = torch.as_tensor(np.ascontiguousarray(edges, dtype=np.int32) if not torch.is_tensor(edges) else edges)
        self.pos = pos_t.to(device=dev, dtype=torch.float64, non_blocking=True).contiguous()
        edges_all = edg_t.to(device=dev, dtype=torch.int32, non_blocking=True).contiguous().reshape(-1, 2)
        self.V, self.gdim = int(self.pos.shape[0]), int(self.pos.shape[1])
        self.n = int(n)
        self.m = 1 << self.n
        self.part = part
        self.glob = None
        with torch.cuda.device(dev):
            self.edge_ids = None
            if part is not None:
                rank, world = part
                self.E_glob = int(edges_all.shape[0])
                self.edges_glob = edges_all
                self.glob = _Tables(lib, edges_all, self.V, dev)
                if epart is None:
                    # contiguous equal index ranges, edge scalars of ALL edges gathered, replicated Schur solve
                    self.e0, self.e1, self.chunk = partition_range(self.E_glob, rank, world)
                    if self.e1 <= self.e0:
                        raise ValueError("empty edge range: more ranks than edges")
                    self.edges = edges_all[self.e0:self.e1].contiguous()
                else:
                    # arbitrary edge set (partition.partition_edges: subtrees): rank-local rake, small exchange
                    self.epart = np.ascontiguousarray(epart, dtype=np.int32)
                    self.edge_ids = np.nonzero(self.epart == rank)[0].astype(np.int64)
                    if len(self.edge_ids) == 0:
                        raise ValueError("rank %d owns no edge" % rank)
                    self.edges = edges_all[torch.from_numpy(self.edge_ids).to(dev)].contiguous()
            else:
                self.edges = edges_all
            self.E = int(self.edges.shape[0])
            V, E, m, gd = self.V, self.E, self.m, self.gdim
            self.num_cells = E * m
            self.num_vertices = V + E * (m - 1)
            st = current_stream()
            # --- mesh (fenics_graph.py:58-99)
            self.coords = torch.empty((self.num_vertices, gd), dtype=torch.float64, device=dev)
            self.cells = torch.empty((self.num_cells, 2), dtype=torch.int32, device=dev)
            self.mf = torch.empty(self.num_cells, dtype=torch.int32, device=dev) if with_mf else None
            self.h = torch.empty(self.num_cells, dtype=torch.float64, device=dev)
            check(lib.gnx_refine_1d(ptr(self.pos), ptr(self.edges), V, E, gd, self.n, ptr(self.coords),
                                    ptr(self.cells), ptr(self.mf), ptr(self.h), st))
            # --- tangents (:230-253)
            self.tangents = torch.empty((E, gd), dtype=torch.float64, device=dev)
            self.elen = torch.empty(E, dtype=torch.float64, device=dev)
            check(lib.gnx_edge_tangents(ptr(self.pos), ptr(self.edges), E, gd, ptr(self.tangents),
                                        ptr(self.elen), st))
            # --- vertex tables (:161-184, :132-156, :280-311)
            t = _Tables(lib, self.edges, V, dev, cls_deg=self.glob.deg if self.glob is not None else None)
            self._tables = t
            for name in ("deg", "bix", "_bif_nodes", "_bnd_nodes", "inc_ptr", "inc_edge", "lam_ptr",
                         "ebif_prefix", "tag_u", "tag_v"):
                setattr(self, name, getattr(t, name))
            loc, tloc = submesh_numbering(self.n)
            self.loc_host, self.tloc_host = loc, tloc
            self.loc = torch.from_numpy(loc).to(dev, non_blocking=True)
            self.tloc = torch.from_numpy(tloc).to(dev, non_blocking=True)
            self.rowtab = torch.from_numpy(row_table(self.n)).to(dev, non_blocking=True) if self.n <= 10 else None
            # per-edge record of the tile kernels (gnx_network_t.erec): one 16-byte load instead of edges -> bix
            eu, ev = self.edges[:, 0].long(), self.edges[:, 1].long()
            flipbit = torch.where(eu > ev, torch.full_like(eu, -(1 << 31)), torch.zeros_like(eu))
            self.erec = torch.stack([self.bix[eu].long(), self.bix[ev].long(), self.ebif_prefix[:E].long(),
                                     eu + flipbit], dim=1).to(torch.int32).contiguous()
            cnt = t.counts()
            if self.glob is not None:
                gcnt = self.glob.counts()
                self.B_glob, self.I_glob = gcnt[0], gcnt[2]
        self.B, self.n_bnd, self.I, self.n_lonely = cnt
        self.bif_nodes = self._bif_nodes[: self.B]
        self.bnd_nodes = self._bnd_nodes[: self.n_bnd]
        self._net = None
        self._gnet = None

    # ---- C struct view ------------------------------------------------------------------
    @property
    def net(self):
        if self._net is None:
            s = _lib.gnx_network_t()
            s.V, s.E, s.gdim, s.n = self.V, self.E, self.gdim, self.n
            s.B, s.n_bnd, s.I = self.B, self.n_bnd, self.I
            for name in ("pos", "edges", "h", "bix", "inc_ptr", "inc_edge", "lam_ptr",
                         "ebif_prefix", "loc", "tloc"):
                setattr(s, name, getattr(self, name).data_ptr())
            s.bif_nodes = self._bif_nodes.data_ptr()
            s.rowtab = self.rowtab.data_ptr() if self.rowtab is not None else None
            s.erec = self.erec.data_ptr() if getattr(self, "erec", None) is not None else None
            self._net = s
        return self._net

    def net_ref(self):
        return C.byref(self.net)

    @property
    def schur_net(self):
        """network descriptor the Schur solve runs on: the WHOLE network (== net when not partitioned)"""
        if self.glob is None:
            return self.net
        if self._gnet is None:
            s = _lib.gnx_network_t()
            g = self.glob
            s.V, s.E, s.gdim, s.n = self.V, self.E_glob, self.gdim, self.n
            s.B, s.n_bnd, s.I = self.B_glob, self.n_bnd, self.I_glob
            s.pos, s.edges = self.pos.data_ptr(), self.edges_glob.data_ptr()
            s.h = None
            s.bix, s.bif_nodes = g.bix.data_ptr(), g._bif_nodes.data_ptr()
            s.inc_ptr, s.inc_edge = g.inc_ptr.data_ptr(), g.inc_edge.data_ptr()
            s.lam_ptr, s.ebif_prefix = g.lam_ptr.data_ptr(), g.ebif_prefix.data_ptr()
            s.loc, s.tloc = self.loc.data_ptr(), self.tloc.data_ptr()
            self._gnet = s
        return self._gnet

    def schur_net_ref(self):
        return C.byref(self.schur_net)

    def schur_edges_host(self):
        """edge list the Schur plan is built from (the whole network)"""
        if self.glob is None:
            return self.edges_host()
        if getattr(self, "_edges_gh", None) is None:
            self._edges_gh = self.edges_glob.cpu().numpy().astype(np.int64)
        return self._edges_gh

    # ---- sizes of the mixed system ---------------------------------------------------------
    @property
    def mixed_ndofs(self):
        return self.E * (self.m + 1) + self.E * self.m + self.B

    @property
    def mixed_nnz(self):
        return self.E * (8 * self.m + 1) + 2 * self.I + self.B

    @property
    def p_off(self):
        return self.E * (self.m + 1)

    @property
    def lam_off(self):
        return self.p_off + self.E * self.m

    # ---- host views (lazy D2H) -----------------------------------------------------------
    def bifurcation_ixs(self):
        return self.bif_nodes.cpu().tolist()

    def boundary_ixs(self):
        return self.bnd_nodes.cpu().tolist()

    def _host_mesh(self):
        if getattr(self, "_mesh_view", None) is None:
            from .mesh import Mesh1D
            self._mesh_view = Mesh1D(self)
        return self._mesh_view

    def edges_host(self):
        if getattr(self, "_edges_h", None) is None:
            self._edges_h = self.edges.cpu().numpy().astype(np.int64)
        return self._edges_h

    def pos_host(self):
        if getattr(self, "_pos_h", None) is None:
            self._pos_h = self.pos.cpu().numpy()
        return self._pos_h

    def sub_cells_host(self):
        """cells of a SubMesh in SubMesh-local vertex numbering [m,2] (same for every edge)."""
        if getattr(self, "_subc", None) is None:
            j = np.arange(self.m, dtype=np.int64)
            s = j.copy()
            sh = 1
            while sh < 64:
                s ^= s >> sh
                sh <<= 1
            loc = self.loc_host.astype(np.int64)
            self._subc = np.sort(np.stack([loc[s], loc[s + 1]], axis=1), axis=1)
        return self._subc

    def vf_host(self):
        """vf[E, m+1]: vertex tags of every SubMesh in SubMesh-local numbering
        (fenics_graph.py:141-156)."""
        if getattr(self, "_vf_h", None) is not None:
            return self._vf_h
        self._vf_h = self._vf_host()
        return self._vf_h

    def _vf_host(self):
        E, m = self.E, self.m
        edges = self.edges_host()
        vf = np.zeros((E, m + 1), dtype=np.int64)
        flip = edges[:, 0] > edges[:, 1]
        lu = np.where(flip, self.loc_host[m], self.loc_host[0])
        lv = np.where(flip, self.loc_host[0], self.loc_host[m])
        ar = np.arange(E)
        vf[ar, lu] = self.tag_u.cpu().numpy()
        vf[ar, lv] = self.tag_v.cpu().numpy()
        return vf

    def parent_vertex_host(self):
        """parent_vertex[E, m+1]: global vertex id of every SubMesh-local vertex (A5)."""
        if getattr(self, "_pv_h", None) is None:
            self._pv_h = self._parent_vertex_host()
        return self._pv_h

    def _parent_vertex_host(self):
        E, m, n, V = self.E, self.m, self.n, self.V
        edges = self.edges.cpu().numpy().astype(np.int64)
        lo, hi = edges.min(axis=1), edges.max(axis=1)
        t = self.tloc_host.astype(np.int64)                # [m+1]
        out = np.empty((E, m + 1), dtype=np.int64)
        e = np.arange(E, dtype=np.int64)[:, None]
        tt = np.broadcast_to(t[None, :], (E, m + 1))
        interior = (tt > 0) & (tt < m)
        ts = np.where(interior, tt, 1)
        ctz = np.zeros_like(ts)
        tmp = ts.copy()
        for _ in range(n):
            z = (tmp & 1) == 0
            ctz += z
            tmp = np.where(z, tmp >> 1, tmp)
        k = n - ctz
        half = 1 << np.maximum(k - 1, 0)
        sp = ts >> (n - k + 1)
        vid = V + E * (half - 1) + e * half + (sp ^ (sp >> 1))
        out[:] = np.where(tt == 0, lo[:, None], np.where(tt == m, hi[:, None], vid))
        return out
