"""CPU-only: the closed-form numbering / row math that the CUDA kernels execute
(graphnics_b200/csrc/gnx_index.h, compiled for the host by tests/hostcheck) against the oracle's
literal level-by-level restatement.  Bit-exact for integers and coordinates."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from _util import HostNet, y_graph, random_graph, golden_arrays

CASES = [("Y", 0), ("Y", 1), ("Y", 2), ("Y", 5), ("Y3", 3), ("rand0", 0), ("rand0", 3),
         ("rand1", 4), ("honeycomb_2_2", 2), ("arterial_tree_N5", 3), ("line_graph_3_dim3", 4)]


def graph(name):
    if name == "Y":
        return y_graph()
    if name == "Y3":
        return y_graph(3)
    if name.startswith("rand"):
        return random_graph(int(name[4:]))
    return golden_arrays(name)


@pytest.mark.parametrize("name,n", CASES)
def test_mesh_bit_exact(name, n):
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    assert np.array_equal(hn.cells, om.cells)
    assert np.array_equal(hn.mf, om.mf)
    assert np.array_equal(hn.coords.view(np.int64), om.coords.view(np.int64)), "coordinates not bit-exact"
    assert np.array_equal(hn.tangents.view(np.int64), om.tangents.view(np.int64)), "tangents not bit-exact"


@pytest.mark.parametrize("name,n", CASES)
def test_submesh_numbering(name, n):
    from graphnics_b200.device import submesh_numbering
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    loc, tloc = submesh_numbering(n)
    m = 1 << n
    # parent vertex of local vertex l on edge e == vertex id of position tloc[l]
    from _util import hostcheck_lib  # noqa: F401  (ensures the shim is built)
    V, E = len(pos), len(edges)
    for e in range(E):
        lo, hi = min(edges[e]), max(edges[e])
        for l in range(m + 1):
            t = int(tloc[l])
            if t == 0:
                vid = lo
            elif t == m:
                vid = hi
            else:
                k = n - (t & -t).bit_length() + 1
                half = 1 << (k - 1)
                sprime = t >> (n - k + 1)
                vid = V + E * (half - 1) + e * half + (sprime ^ (sprime >> 1))
            assert vid == om.parent_vertex[e, l]


@pytest.mark.parametrize("name,n", CASES)
def test_pattern_and_values(name, n):
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    rng = np.random.default_rng(1)
    Res = rng.uniform(0.5, 2.0, size=len(edges))
    A = oracle.assemble_mixed(om, Res)
    rowptr, colidx = hn.pattern()
    assert hn.N == A.shape[0] and hn.nnz == A.nnz
    assert np.array_equal(rowptr, A.indptr)
    assert np.array_equal(colidx, A.indices)
    vals = hn.values(a_edge=Res)
    np.testing.assert_allclose(vals, A.data, rtol=1e-12, atol=1e-300)
    # NetworkStokes viscous term and a sigma-scaled operator
    visc = rng.uniform(0.1, 1.0, size=len(edges))
    A2 = oracle.assemble_mixed(om, Res, visc=visc)
    np.testing.assert_allclose(hn.values(a_edge=Res, k_edge=visc), A2.data, rtol=1e-11, atol=1e-300)
    # mass matrix on the same pattern
    D = oracle.assemble_mass_q_mixed(om, 2.5)
    Dg = sp.csr_matrix((hn.values(a_const=2.5, sigma=0.0), colidx, rowptr), shape=A.shape)
    assert abs(Dg - D).max() <= 1e-12 * abs(D).max()


@pytest.mark.parametrize("name,n", CASES)
def test_rhs(name, n):
    pos, edges = graph(name)
    om = oracle.build_mesh(pos, edges, n)
    hn = HostNet(pos, edges, n)
    lin = oracle.Coef(lambda x: 1.0 + 2.0 * x[..., 0] - 0.5 * x[..., 1], 1)
    g = oracle.Coef(lambda x: np.sin(x[..., 0]) + x[..., 1] ** 2, 1)
    f = oracle.Coef(lambda x: np.cos(x[..., 1]) - x[..., 0], 1)
    b = oracle.assemble_mixed_rhs(om, f=f, g=g, p_bc=lin, project=False)
    nod = lambda c: c(om.sub_coords).ravel()
    bh = hn.rhs(f=nod(f), g=nod(g), pbc_nodal=nod(lin), pbc_node=lin(pos))
    np.testing.assert_allclose(bh, b, rtol=1e-12, atol=1e-14)


def test_vf_tags_match_oracle():
    from _util import host_tables
    from graphnics_b200.device import submesh_numbering
    for name, n in CASES:
        pos, edges = graph(name)
        om = oracle.build_mesh(pos, edges, n)
        t = host_tables(len(pos), edges)
        loc, _ = submesh_numbering(n)
        m = 1 << n
        vf = np.zeros_like(om.vf)
        flip = edges[:, 0] > edges[:, 1]
        ar = np.arange(len(edges))
        vf[ar, np.where(flip, loc[m], loc[0])] = t["tag_u"]
        vf[ar, np.where(flip, loc[0], loc[m])] = t["tag_v"]
        assert np.array_equal(vf, om.vf), name
        assert t["bif_nodes"].tolist() == om.bif_ixs and t["bnd_nodes"].tolist() == om.bnd_ixs
