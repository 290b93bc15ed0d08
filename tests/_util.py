"""Shared test helpers (CPU only): golden fixtures, small graphs, host-side network tables."""
import ctypes as C
import json
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def fromhex(rows):
    return np.array([[float.fromhex(x) for x in r] for r in rows], dtype=np.float64)


def golden_graphs():
    with open(os.path.join(HERE, "golden", "graphs.json")) as fh:
        g = json.load(fh)
    g.pop("_meta", None)
    return g


def golden_arrays(name):
    c = golden_graphs()[name]
    return fromhex(c["pos"]), np.asarray(c["edges"], dtype=np.int64)


def y_graph(dim=2):
    pos = np.array([[0, 0], [0, .5], [-.5, 1], [.5, 1]], dtype=np.float64)
    if dim == 3:
        pos = np.hstack([pos, np.zeros((4, 1))])
    return pos, np.array([[0, 1], [1, 2], [1, 3]], dtype=np.int64)


def random_graph(seed, V=12, extra=6, gdim=3, flip_some=True):
    """Connected random graph: random tree + extra edges, random orientations (so u>v occurs)."""
    rng = np.random.default_rng(seed)
    pos = rng.normal(size=(V, gdim))
    edges = []
    for v in range(1, V):
        u = int(rng.integers(0, v))
        edges.append((u, v))
    have = set(edges)
    tries = 0
    while extra > 0 and tries < 1000:
        tries += 1
        u, v = (int(x) for x in rng.integers(0, V, size=2))
        if u == v or (u, v) in have or (v, u) in have:
            continue
        have.add((u, v)); edges.append((u, v)); extra -= 1
    edges = np.asarray(edges, dtype=np.int64)
    if flip_some:
        f = rng.random(len(edges)) < 0.5
        edges[f] = edges[f][:, ::-1]
    # emulate DiGraph.edges() order: by source node (insertion order = label), then insertion
    order = np.argsort(edges[:, 0], kind="stable")
    return pos, edges[order]


# ---------------------------------------------------------------------------------------
# host tables (what gnx_vertex_tables produces) -- numpy, for the host index self-check
# ---------------------------------------------------------------------------------------
def host_tables(V, edges):
    edges = np.asarray(edges, dtype=np.int64)
    E = len(edges)
    deg = np.bincount(edges.ravel(), minlength=V)
    isbif = deg > 1
    bix = np.where(isbif, np.cumsum(isbif) - 1, -1)
    bif_nodes = np.nonzero(isbif)[0]
    bnd_nodes = np.nonzero(deg == 1)[0]
    inc_ptr = np.concatenate([[0], np.cumsum(deg)])
    codes = np.concatenate([2 * np.arange(E), 2 * np.arange(E) + 1])
    nodes = np.concatenate([edges[:, 0], edges[:, 1]])
    order = np.lexsort((codes, nodes))
    inc_edge = codes[order]
    lam_ptr = np.concatenate([[0], np.cumsum(deg[bif_nodes])])
    ebif = isbif[edges[:, 0]].astype(np.int64) + isbif[edges[:, 1]]
    ebif_prefix = np.concatenate([[0], np.cumsum(ebif)])
    tag_u = np.where(deg[edges[:, 0]] > 1, 2, np.where(deg[edges[:, 0]] == 1, 3, 0))
    tag_v = np.where(deg[edges[:, 1]] > 1, 1, np.where(deg[edges[:, 1]] == 1, 4, 0))
    return dict(deg=deg, bix=bix, bif_nodes=bif_nodes, bnd_nodes=bnd_nodes, inc_ptr=inc_ptr,
                inc_edge=inc_edge, lam_ptr=lam_ptr, ebif_prefix=ebif_prefix, tag_u=tag_u, tag_v=tag_v,
                B=len(bif_nodes), n_bnd=len(bnd_nodes), I=int(ebif.sum()))


_HC = None


def hostcheck_lib():
    """g++-compiled TEST-ONLY build of graphnics_b200/csrc/gnx_index.h (tests/hostcheck)."""
    global _HC
    if _HC is None:
        src = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
        out = os.path.join(HERE, "hostcheck", "libhostcheck.so")
        hdrs = [os.path.join(ROOT, "graphnics_b200", "csrc", h) for h in ("gnx_index.h", "gnx_rpn.h", "gnx_primal.h")]
        hdrs.append(os.path.join(ROOT, "include", "graphnics_b200.h"))
        if (not os.path.exists(out) or os.path.getmtime(out) < max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in hdrs])):
            subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src])
        _HC = C.CDLL(out)
        _HC.hc_ndofs.restype = C.c_int64
        _HC.hc_nnz.restype = C.c_int64
        _HC.hc_rpn.restype = C.c_double
        _HC.hc_primal_ndofs.restype = C.c_int64
        _HC.hc_primal_nnz.restype = C.c_int64
    return _HC


class HostNet:
    """gnx_network_t whose pointers are HOST numpy arrays (for tests/hostcheck only)."""

    def __init__(self, pos, edges, n):
        from graphnics_b200._lib import gnx_network_t
        from graphnics_b200.device import submesh_numbering
        self.pos = np.ascontiguousarray(pos, dtype=np.float64)
        self.edges = np.ascontiguousarray(edges, dtype=np.int32)
        self.V, self.gdim = self.pos.shape
        self.E = len(self.edges)
        self.n, self.m = n, 1 << n
        hc = hostcheck_lib()
        nv = self.V + self.E * (self.m - 1)
        self.coords = np.full((nv, self.gdim), np.nan)
        self.cells = np.full((self.E * self.m, 2), -1, dtype=np.int32)
        self.mf = np.full(self.E * self.m, -1, dtype=np.int32)
        self.h = np.full(self.E * self.m, np.nan)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        hc.hc_refine(vp(self.pos), vp(self.edges), self.V, self.E, self.gdim, n, vp(self.coords),
                     vp(self.cells), vp(self.mf), vp(self.h))
        self.tangents = np.empty((self.E, self.gdim))
        hc.hc_tangents(vp(self.pos), vp(self.edges), self.E, self.gdim, vp(self.tangents))
        t = host_tables(self.V, self.edges)
        self.tables = t
        self.loc, self.tloc = submesh_numbering(n)
        self._arr = {k: np.ascontiguousarray(t[k], dtype=np.int32) for k in
                     ("bix", "bif_nodes", "inc_ptr", "inc_edge", "lam_ptr", "ebif_prefix")}
        s = gnx_network_t()
        s.V, s.E, s.gdim, s.n = self.V, self.E, self.gdim, n
        s.B, s.n_bnd, s.I = t["B"], t["n_bnd"], t["I"]
        s.pos, s.edges, s.h = self.pos.ctypes.data, self.edges.ctypes.data, self.h.ctypes.data
        for k, a in self._arr.items():
            setattr(s, k, a.ctypes.data)
        s.loc, s.tloc = self.loc.ctypes.data, self.tloc.ctypes.data
        self.net = s
        self.N = int(hc.hc_ndofs(C.byref(s)))
        self.nnz = int(hc.hc_nnz(C.byref(s)))

    def pattern(self):
        hc = hostcheck_lib()
        rowptr = np.full(self.N + 1, -1, dtype=np.int32)
        colidx = np.full(self.nnz, -1, dtype=np.int32)
        hc.hc_pattern(C.byref(self.net), rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p))
        return rowptr, colidx

    def values(self, a_edge=None, k_edge=None, a_const=1.0, sigma=1.0):
        hc = hostcheck_lib()
        vals = np.full(self.nnz, np.nan)
        a = None if a_edge is None else np.ascontiguousarray(a_edge, dtype=np.float64)
        k = None if k_edge is None else np.ascontiguousarray(k_edge, dtype=np.float64)
        vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
        hc.hc_values(C.byref(self.net), vp(a), vp(k), C.c_double(a_const), C.c_double(sigma), vp(vals))
        return vals

    def rhs(self, f=None, g=None, pbc_out=None, pbc_node=None, f_const=0.0, g_const=0.0):
        hc = hostcheck_lib()
        b = np.full(self.N, np.nan)
        arrs = [None if x is None else np.ascontiguousarray(x, dtype=np.float64) for x in (f, g, pbc_out, pbc_node)]
        vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
        hc.hc_rhs(C.byref(self.net), vp(arrs[0]), vp(arrs[1]), C.c_double(f_const), C.c_double(g_const),
                  vp(arrs[2]), vp(arrs[3]), vp(b))
        return b

    # ---- primal model ------------------------------------------------------------------------
    def primal_pattern(self):
        hc = hostcheck_lib()
        N, nnz = int(hc.hc_primal_ndofs(C.byref(self.net))), int(hc.hc_primal_nnz(C.byref(self.net)))
        rowptr = np.full(N + 1, -1, dtype=np.int32)
        colidx = np.full(nnz, -1, dtype=np.int32)
        hc.hc_primal_pattern(C.byref(self.net), rowptr.ctypes.data_as(C.c_void_p), colidx.ctypes.data_as(C.c_void_p))
        return rowptr, colidx

    def cell_slot(self):
        """global cell -> edge-local slot (order of `h` and of per-cell coefficients)"""
        c = np.arange(self.E * self.m)
        e, j = c >> self.n, c & (self.m - 1)
        s = j.copy()
        sh = 1
        while sh < 64:
            s ^= s >> sh
            sh <<= 1
        flip = (self.edges[:, 0] > self.edges[:, 1])[e]
        return e * self.m + np.where(flip, self.m - 1 - s, s)

    def primal_values(self, a_const=1.0, a_cell=None, sigma=1.0):
        hc = hostcheck_lib()
        vals = np.full(int(hc.hc_primal_nnz(C.byref(self.net))), np.nan)
        a = None if a_cell is None else np.ascontiguousarray(a_cell, dtype=np.float64)
        hc.hc_primal_values(C.byref(self.net), C.c_double(a_const), None if a is None else a.ctypes.data_as(C.c_void_p),
                            C.c_double(sigma), vals.ctypes.data_as(C.c_void_p))
        return vals

    def primal_rhs(self, f_cell=None, f_lin=0, fc=0.0, g_cell=None, g_lin=0, gc=0.0):
        hc = hostcheck_lib()
        b = np.full(int(hc.hc_primal_ndofs(C.byref(self.net))), np.nan)
        fa = None if f_cell is None else np.ascontiguousarray(f_cell, dtype=np.float64)
        ga = None if g_cell is None else np.ascontiguousarray(g_cell, dtype=np.float64)
        vp = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
        hc.hc_primal_rhs(C.byref(self.net), vp(self.cells), vp(fa), int(f_lin), C.c_double(fc), vp(ga), int(g_lin),
                         C.c_double(gc), vp(b))
        return b

    def end_values(self, nodal):
        """value of a q-dof-layout nodal field at the v end of every edge"""
        nodal = np.asarray(nodal).reshape(self.E, self.m + 1)
        flip = self.edges[:, 0] > self.edges[:, 1]
        lv = np.where(flip, self.loc[0], self.loc[self.m])
        return nodal[np.arange(self.E), lv]
