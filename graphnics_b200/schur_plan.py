"""Symbolic analysis of the graph-level Schur system (built once per network, host side).

After the per-edge elimination (gnx_edge_eliminate_mixed) the unknowns are the bifurcation
multipliers; their matrix is a weighted graph Laplacian  S = sum_e cond_e (e_i - e_j)(e_i - e_j)^T
(+ Dirichlet contributions of edges that end in boundary nodes).  Vertices of the bifurcation graph
that have exactly ONE remaining bifurcation neighbour can be eliminated exactly and without fill
("rake"): the parent only receives a diagonal and a right-hand-side update.  Repeating this level by
level removes every tree-like part of the network -- an arterial tree completely, so the Krylov
iteration of the reference-replacing solver (CG on the Schur complement) only ever sees the cyclic
CORE of the graph (the whole lattice for a honeycomb, nothing for a tree).

The plan is pure index data (it does not depend on conductances), so the time loop and repeated
solves reuse it.  All arrays are int32, bifurcation-index based.
"""
import numpy as np

__all__ = ["SchurPlan", "MAX_LEVELS"]

MAX_LEVELS = 64          # rake levels kept in the kernel's parameter block; deeper parts go to the core


class SchurPlan:
    """
    levels      list of int32 arrays: bifurcation indices eliminated at level l (up-sweep order)
    parent[B]   bifurcation index of the node a raked node is folded into (-1: none / not raked)
    pedge[B]    graph edge that links a raked node to its parent (-1 if parent == -1)
    ch_ptr[B+1], ch_idx   children (raked nodes) of every node, CSR
    core_nodes[nc], core_of[B], crp[nc+1], ccol[nnzc] (core-local), cedge[nnzc] (graph edge)
    """

    def __init__(self, edges, bix, B, max_levels=MAX_LEVELS):
        edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
        bix = np.asarray(bix, dtype=np.int64)
        self.B = int(B)
        self.agg = None                                       # set by aggregate_core (two-level PCG)
        bu, bv = bix[edges[:, 0]], bix[edges[:, 1]]
        inner = np.nonzero((bu >= 0) & (bv >= 0))[0]          # edges joining two bifurcations
        eu, ev = bu[inner], bv[inner]
        # half-edge lists (node, other, edge)
        hn = np.concatenate([eu, ev])
        ho = np.concatenate([ev, eu])
        he = np.concatenate([inner, inner])
        order = np.argsort(hn, kind="stable")
        hn, ho, he = hn[order], ho[order], he[order]
        ptr = np.concatenate([[0], np.cumsum(np.bincount(hn, minlength=B))])
        deg = np.diff(ptr).astype(np.int64)                   # remaining bifurcation-neighbour count
        alive = np.ones(B, dtype=bool)
        half_alive = np.ones(len(hn), dtype=bool)
        parent = -np.ones(B, dtype=np.int64)
        pedge = -np.ones(B, dtype=np.int64)
        levels = []
        # position of the reverse half-edge, to retire both directions of an edge together
        rev = np.empty(len(hn), dtype=np.int64)
        srt = np.argsort(he, kind="stable")
        # every inner edge id appears exactly twice in `he`
        rev[srt[0::2]] = srt[1::2]
        rev[srt[1::2]] = srt[0::2]
        while len(levels) < max_levels:
            cand = np.nonzero(alive & (deg <= 1))[0]
            if cand.size == 0:
                break
            # the single live half-edge of every degree-1 candidate
            par = -np.ones(cand.size, dtype=np.int64)
            pe = -np.ones(cand.size, dtype=np.int64)
            hidx = -np.ones(cand.size, dtype=np.int64)
            one = deg[cand] == 1
            if one.any():
                c1 = cand[one]
                # first live half-edge of each node: vectorised via segmented argmax
                starts, ends = ptr[c1], ptr[c1 + 1]
                maxlen = int((ends - starts).max())
                found = -np.ones(c1.size, dtype=np.int64)
                for off in range(maxlen):
                    idx = starts + off
                    ok = (idx < ends) & (found < 0)
                    okidx = idx[ok]
                    live = half_alive[okidx]
                    tmp = found[ok]
                    tmp[live] = okidx[live]
                    found[ok] = tmp
                hidx[one] = found
                par[one] = ho[found]
                pe[one] = he[found]
            # two mutually-adjacent candidates: keep the higher index for a later level
            cand_mask = np.zeros(B, dtype=bool)
            cand_mask[cand] = True
            clash = one & (par >= 0)
            clash[clash] = cand_mask[par[clash]] & (cand[clash] > par[clash])
            keep = ~clash
            cand, par, pe, hidx = cand[keep], par[keep], pe[keep], hidx[keep]
            if cand.size == 0:
                break
            levels.append(cand.astype(np.int32))
            alive[cand] = False
            parent[cand] = par
            pedge[cand] = pe
            hp = hidx[hidx >= 0]
            half_alive[hp] = False
            half_alive[rev[hp]] = False
            np.subtract.at(deg, par[par >= 0], 1)
            deg[cand] = 0
        self.levels = levels
        self.parent = parent.astype(np.int32)
        self.pedge = pedge.astype(np.int32)
        # children CSR
        kids = np.nonzero(parent >= 0)[0]
        o = np.argsort(parent[kids], kind="stable")
        self.ch_idx = kids[o].astype(np.int32)
        self.ch_ptr = np.concatenate([[0], np.cumsum(np.bincount(parent[kids], minlength=B))]).astype(np.int32)
        # core
        core = np.nonzero(alive)[0]
        self.core_nodes = core.astype(np.int32)
        self.nc = int(core.size)
        core_of = -np.ones(B, dtype=np.int64)
        core_of[core] = np.arange(core.size)
        self.core_of = core_of.astype(np.int32)
        keep = half_alive & alive[hn] & alive[ho]
        cn, co_, ce = core_of[hn[keep]], core_of[ho[keep]], he[keep]
        self.crp = np.concatenate([[0], np.cumsum(np.bincount(cn, minlength=core.size))]).astype(np.int32)
        self.ccol = co_.astype(np.int32)
        self.cedge = ce.astype(np.int32)
        self.nnzc = int(len(ce))

    # level pointer / flat node list as the kernel wants them
    def level_arrays(self):
        sizes = [len(l) for l in self.levels]
        lvl_ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        nodes = np.concatenate(self.levels).astype(np.int32) if self.levels else np.zeros(0, dtype=np.int32)
        return lvl_ptr, nodes

    def split_level(self, top_max=2048):
        """first level of the TOP part: the last levels, holding <= top_max nodes in total, are
        swept by one CTA out of shared memory (dependent global loads cost ~1 us per level)"""
        sizes = [len(l) for l in self.levels]
        l, tot = len(sizes), 0
        while l > 0 and tot + sizes[l - 1] <= top_max:
            tot += sizes[l - 1]
            l -= 1
        return l

    def top_arrays(self, l_split):
        """(top_nodes [nt], top_of [B], level_of [B]) for the levels >= l_split"""
        level_of = np.full(self.B, len(self.levels), dtype=np.int32)
        for l, nodes in enumerate(self.levels):
            level_of[nodes] = l
        top = np.concatenate(self.levels[l_split:]).astype(np.int32) if self.levels[l_split:] \
            else np.zeros(0, dtype=np.int32)
        top_of = -np.ones(self.B, dtype=np.int32)
        top_of[top] = np.arange(len(top), dtype=np.int32)
        return top, top_of, level_of

    def block_arrays(self, cap=1024, top_max=2048, force_upper=None, count=None):
        """Blocked rake of a forest (nc == 0): the raked nodes are split into maximal subtrees of <= cap
        nodes (BLOCKS: complete subtrees, so all of a block's children are inside it; one CTA sweeps a block
        out of shared memory with block barriers only) and the UPPER part (nodes whose subtree is larger),
        which takes the place of the level-defined top part.  Two grid barriers then replace one per wide
        level.  Returns None when it does not apply (a core exists, or the upper part exceeds top_max),
        else dict(nb, blk_ptr [nb+1], blk_nodes, blk_pos [B] (-1: upper), blk_lv [2 nb] (lowest, highest
        rake level in the block), top_nodes, top_of [B], level_of [B], l_top (lowest level of the upper part))."""
        if self.nc != 0 or not self.levels:
            return None
        B, nl = self.B, len(self.levels)
        level_of = np.full(B, nl, dtype=np.int32)
        for l, nodes in enumerate(self.levels):
            level_of[nodes] = l
        parent = self.parent.astype(np.int64)
        size = np.ones(B, dtype=np.int64)
        for nodes in self.levels:                            # children sit on strictly lower levels
            par = parent[nodes]
            has = par >= 0
            np.add.at(size, par[has], size[nodes[has]])
        upper = size > cap
        if force_upper is not None:                          # (rank view: the replicated nodes always sit in the top part)
            upper = upper | force_upper
        in_levels = np.zeros(B, dtype=bool)
        in_levels[np.concatenate(self.levels)] = True
        upper &= in_levels
        top = np.nonzero(upper)[0]
        if len(top) > top_max:
            return None
        # block roots: subtree fits, the parent's does not (or there is no parent)
        par_big = np.where(parent >= 0, upper[np.maximum(parent, 0)], True)
        root = (~upper) & par_big & in_levels
        blk_of = -np.ones(B, dtype=np.int64)
        roots = np.nonzero(root)[0]
        blk_of[roots] = np.arange(len(roots))
        for nodes in reversed(self.levels):                  # top-down: a non-root node inherits its parent's block
            inner = nodes[(~upper[nodes]) & (~root[nodes])]
            blk_of[inner] = blk_of[parent[inner]]
        inb = np.nonzero(blk_of >= 0)[0]
        assert len(inb) + len(top) == (B if count is None else count)
        order = np.lexsort((level_of[inb], blk_of[inb]))     # by block, then by level
        blk_nodes = inb[order]
        counts = np.bincount(blk_of[inb], minlength=len(roots))
        blk_ptr = np.concatenate([[0], np.cumsum(counts)])
        blk_pos = -np.ones(B, dtype=np.int64)
        blk_pos[blk_nodes] = np.arange(len(blk_nodes)) - np.repeat(blk_ptr[:-1], counts)
        lv = level_of[blk_nodes]
        blk_lv = np.stack([lv[blk_ptr[:-1]], lv[blk_ptr[1:] - 1]], axis=1).ravel()
        top = top[np.argsort(level_of[top], kind="stable")]
        top_of = -np.ones(B, dtype=np.int32)
        top_of[top] = np.arange(len(top), dtype=np.int32)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        return dict(nb=len(roots), blk_ptr=i32(blk_ptr), blk_nodes=i32(blk_nodes), blk_pos=i32(blk_pos),
                    blk_lv=i32(blk_lv), top_nodes=i32(top), top_of=top_of, level_of=level_of,
                    l_top=int(level_of[top].min()) if len(top) else nl)

    def aggregate_core(self, pos, na=128, ns=8):
        """Coarse spaces of the multilevel-additive preconditioner of the core PCG
            M^-1 = D^-1 + W1 diag(W1^T S W1)^-1 W1^T + W2 (W2^T S W2)^-1 W2^T
        (W1 / W2: indicator vectors of na*ns fine / na coarse aggregates, every coarse aggregate being the
        union of ns fine ones): the core is cut by recursive coordinate bisection of the node positions
        `pos` [B, gdim] into na*ns compact, equally sized fine aggregates (positions only steer the QUALITY
        of the aggregates; any partition gives a symmetric positive definite preconditioner), and the core
        is RENUMBERED fine aggregate by fine aggregate, so that fine and coarse aggregates are contiguous
        row ranges (one CTA per coarse aggregate in the kernel).  na, ns powers of two.  Sets self.agg =
        dict(na, ns, agg_ptr [na+1], sub_ptr [na*ns+1], row_sub [nc] / col_agg [nnzc]: FINE aggregate of a row
        / of the column of a CSR entry (coarse = fine // ns), cut_ptr [na*na+1], cut_k: the CSR entries that
        couple coarse aggregate I to J != I, grouped by (I, J) in a fixed order -- the coarse matrix is summed
        over these lists, deterministically)."""
        nc = self.nc
        na = int(min(na, nc))
        while na & (na - 1):
            na &= na - 1                                      # power of two
        ns = int(ns)
        while ns > 1 and na * ns > nc:
            ns //= 2
        nf = na * ns
        p = np.asarray(pos, dtype=np.float64)[self.core_nodes]
        owner = np.zeros(nc, dtype=np.int64)
        todo = [(np.arange(nc), nf, 0)]
        while todo:                                           # halving: fine parts [J*ns, (J+1)*ns) make up coarse part J
            idx, k, base = todo.pop()
            if k == 1:
                owner[idx] = base
                continue
            ext = p[idx].max(axis=0) - p[idx].min(axis=0)
            ax = int(np.argmax(ext))
            k0 = k // 2
            cut = len(idx) * k0 // k
            o = np.argsort(p[idx, ax], kind="stable")
            todo.append((idx[o[:cut]], k0, base))
            todo.append((idx[o[cut:]], k - k0, base + k0))
        # renumber the core: aggregates become contiguous ranges (stable inside an aggregate)
        order = np.argsort(owner, kind="stable")
        newpos = np.empty(nc, dtype=np.int64)
        newpos[order] = np.arange(nc)
        deg = np.diff(self.crp).astype(np.int64)
        new_crp = np.concatenate([[0], np.cumsum(deg[order])]).astype(np.int64)
        src = np.repeat(self.crp[order].astype(np.int64), deg[order]) + \
            (np.arange(int(deg.sum())) - np.repeat(new_crp[:-1], deg[order]))
        self.core_nodes = self.core_nodes[order].astype(np.int32)
        self.core_of = -np.ones(self.B, dtype=np.int32)
        self.core_of[self.core_nodes] = np.arange(nc, dtype=np.int32)
        self.ccol = newpos[self.ccol[src]].astype(np.int32)
        self.cedge = self.cedge[src].astype(np.int32)
        self.crp = new_crp.astype(np.int32)
        row_sub = owner[order]
        sub_ptr = np.concatenate([[0], np.cumsum(np.bincount(row_sub, minlength=nf))]).astype(np.int32)
        agg_ptr = sub_ptr[::ns].copy()
        rows = np.repeat(np.arange(nc), np.diff(self.crp))
        col_sub = row_sub[self.ccol]
        I, J = row_sub[rows] // ns, col_sub // ns
        cut = np.nonzero(I != J)[0]
        key = I[cut] * na + J[cut]
        o = np.argsort(key, kind="stable")
        cut_k = cut[o].astype(np.int32)
        cut_ptr = np.concatenate([[0], np.cumsum(np.bincount(key, minlength=na * na))]).astype(np.int32)
        self.agg = dict(na=na, ns=ns, agg_ptr=agg_ptr, sub_ptr=sub_ptr, row_sub=row_sub.astype(np.int32),
                        agg_of=(row_sub // ns).astype(np.int32), col_agg=col_sub.astype(np.int32),
                        cut_ptr=cut_ptr, cut_k=cut_k if len(cut_k) else np.zeros(1, dtype=np.int32))
        return self.agg

    def ell_arrays(self, ctas=16, rmax=5120, wmax=8):
        """ELL copy of the core matrix for the cluster-resident PCG: (w, R, col [w*nc] packed
        (CTA << 16 | offset), edge [w*nc]); w = 0 when the core does not qualify."""
        nc = self.nc
        if nc == 0:
            return 0, 0, np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        deg = np.diff(self.crp)
        w = int(deg.max()) if len(deg) else 0
        R = -(-nc // ctas)
        if w == 0 or w > wmax or R > rmax:
            return 0, 0, np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        col = -np.ones((w, nc), dtype=np.int64)
        edge = -np.ones((w, nc), dtype=np.int64)
        rows = np.repeat(np.arange(nc), deg)
        k = np.arange(len(rows)) - self.crp[rows]
        j = self.ccol.astype(np.int64)
        col[k, rows] = ((j // R) << 16) | (j % R)
        edge[k, rows] = self.cedge
        return w, R, col.ravel().astype(np.int32), edge.ravel().astype(np.int32)


def top_records(hp, top, top_of, level_of, inc_ptr, inc_edge, bif_nodes, ncls=None):
    """Packed per-row records of the top part (gnx_schur_plan_t.top_rec, [nt, 16] int32): everything the one-CTA sweep
    needs to know about a top node in ONE 64-byte load instead of a chain of dependent look-ups --
      0 node (bifurcation index)   1 rake level   2 top position of the parent (-1)   3 edge to the parent (-1)
      4, 5 top positions of the first two children kept in shared memory (-1)         6 number of such children
      7, 8 the first two children folded from global memory (-1)    9, 10 their edges     11 graph node
      12..15 incident edges as (edge << 1 | end) in incidence-list order (-1 unused)
    bit 30 of entry 6 marks a row that does not fit (more than two children of a kind, more than four incident
    edges): the kernel then walks the generic arrays for it.  Rank views (ncls given): a replicated node keeps only
    replicated children in shared memory, local children -- of any rank -- are folded from global memory."""
    nt = len(top)
    rec = -np.ones((nt, 16), dtype=np.int64)
    top = np.asarray(top, dtype=np.int64)
    rec[:, 0] = top
    rec[:, 1] = np.asarray(level_of)[top]
    par = hp.parent[top].astype(np.int64)
    rec[:, 2] = np.where(par >= 0, np.asarray(top_of)[np.maximum(par, 0)], -1)
    rec[:, 3] = np.where(par >= 0, hp.pedge[top], -1)
    rec[:, 6] = 0
    gnode = np.asarray(bif_nodes, dtype=np.int64)[top]
    rec[:, 11] = gnode
    ch_ptr, ch_idx = hp.ch_ptr.astype(np.int64), hp.ch_idx.astype(np.int64)
    slow = np.zeros(nt, dtype=bool)
    nkids = ch_ptr[top + 1] - ch_ptr[top]
    n_in, n_out = np.zeros(nt, dtype=np.int64), np.zeros(nt, dtype=np.int64)
    for j in range(int(nkids.max()) if nt else 0):          # j-th child of every row that has one
        has = nkids > j
        c = ch_idx[np.where(has, ch_ptr[top] + j, 0)]
        tc = np.asarray(top_of)[c].astype(np.int64)
        if ncls is not None:                                 # replicated row, local child: from global memory
            tc = np.where((ncls[top] & 1).astype(bool) & ~(ncls[c] & 1).astype(bool), -1, tc)
        inside = has & (tc >= 0)
        outside = has & (tc < 0)
        for k in (0, 1):
            put = inside & (n_in == k)
            rec[put, 4 + k] = tc[put]
            put = outside & (n_out == k)
            rec[put, 7 + k] = c[put]
            rec[put, 9 + k] = hp.pedge[c[put]]
        slow |= (inside & (n_in >= 2)) | (outside & (n_out >= 2))
        n_in += inside
        n_out += outside
    rec[:, 6] = n_in
    inc_ptr, inc_edge = np.asarray(inc_ptr, dtype=np.int64), np.asarray(inc_edge, dtype=np.int64)
    deg = inc_ptr[gnode + 1] - inc_ptr[gnode]
    slow |= deg > 4
    for j in range(4):
        has = deg > j
        rec[has, 12 + j] = inc_edge[inc_ptr[gnode[has]] + j]
    rec[slow, 6] |= 1 << 30
    return np.ascontiguousarray(rec, dtype=np.int32)


class RankPlan(SchurPlan):
    """One rank's view of the plan of an EDGE-PARTITIONED network.

    `owner[i]` (per bifurcation) says who eliminates a raked node: rank r if every edge at node i and at all of
    its rake descendants belongs to rank r (the whole subtree is rank-local: r builds the rows and sweeps it
    without talking to anybody), -1 otherwise -- the REPLICATED nodes (core nodes and raked nodes that see more
    than one rank), which every rank processes redundantly after ONE exchange: the owner of an edge at a
    replicated node sends its three scalars, the owner of a local subtree whose parent is replicated sends the
    subtree root's final (diagonal, right-hand side) pair.

    The view keeps the global index arrays (parent, pedge, children, core) and re-levels the nodes this rank
    works on: its local nodes keep their rake level, the replicated raked nodes follow on levels >= l_sync
    (their height among themselves) -- so "everything below l_sync" is exactly what can be done before the
    exchange, and the kernel's level loops need one synchronisation point.
      ncls[B]  bit 0: replicated (row built after the exchange), bit 1: foreign (another rank's local node:
               skipped), bit 2: local root whose (d, r) is exported
    """

    def __init__(self, hp, owner, rank, max_levels=MAX_LEVELS):
        self.__dict__.update(hp.__dict__)
        B = hp.B
        owner = np.asarray(owner, dtype=np.int64)
        self.owner, self.rank = owner, int(rank)
        raked = np.zeros(B, dtype=bool)
        hlevel = np.full(B, -1, dtype=np.int64)
        for l, nodes in enumerate(hp.levels):
            raked[nodes] = True
            hlevel[nodes] = l
        local = raked & (owner == rank)
        repl = raked & (owner < 0)
        L_loc = int(hlevel[local].max()) + 1 if local.any() else 0
        # height of the replicated raked nodes among themselves
        tl = np.zeros(B, dtype=np.int64)
        parent = hp.parent.astype(np.int64)
        for nodes in hp.levels:                               # children before parents
            t = nodes[repl[nodes]]
            par = parent[t]
            ok = (par >= 0) & repl[np.maximum(par, 0)]
            np.maximum.at(tl, par[ok], tl[t[ok]] + 1)
        n_t = int(tl[repl].max()) + 1 if repl.any() else 0
        if L_loc + n_t > max_levels:
            raise ValueError("rank view needs %d rake levels (limit %d)" % (L_loc + n_t, max_levels))
        levels = [hp.levels[l][local[hp.levels[l]]] for l in range(L_loc)]
        rn = np.nonzero(repl)[0]
        levels += [rn[tl[rn] == k].astype(np.int32) for k in range(n_t)]
        self.levels = [np.ascontiguousarray(l, dtype=np.int32) for l in levels]
        self.l_sync = L_loc
        self.n_mine = int(local.sum() + repl.sum())          # raked nodes this rank sweeps
        par_repl = np.where(parent >= 0, (owner[np.maximum(parent, 0)] < 0), False)
        ncls = np.zeros(B, dtype=np.int32)
        ncls[owner < 0] |= 1
        ncls[(owner >= 0) & (owner != rank)] |= 2
        ncls[local & par_repl] |= 4
        self.ncls = ncls
        self.repl_raked = repl

    def rank_block_arrays(self, cap=1024, top_max=2048):
        return self.block_arrays(cap=cap, top_max=top_max, force_upper=self.repl_raked, count=self.n_mine)


def node_owners(hp, edges, bix, epart):
    """owner[B] of RankPlan for the edge partition `epart` (rank per edge) of the network `edges`"""
    edges = np.asarray(edges, dtype=np.int64).reshape(-1, 2)
    epart = np.asarray(epart, dtype=np.int64)
    B = hp.B
    bu, bv = np.asarray(bix)[edges[:, 0]], np.asarray(bix)[edges[:, 1]]
    lo = np.full(B, np.iinfo(np.int64).max)
    hi = np.full(B, -1, dtype=np.int64)
    for be in (bu, bv):
        k = be >= 0
        np.minimum.at(lo, be[k], epart[k])
        np.maximum.at(hi, be[k], epart[k])
    owner = np.where(lo == hi, lo, -1)
    owner[hp.core_nodes] = -1
    parent = hp.parent.astype(np.int64)
    for nodes in hp.levels:                                   # a node is only as local as its children
        par = parent[nodes]
        has = par >= 0
        bad = has & (owner[nodes] != owner[np.maximum(par, 0)])
        owner[par[bad]] = -1
    # (a parent demoted above may itself have been visited as a child already only on a LOWER level: levels
    #  are processed bottom-up, so every demotion reaches the ancestors in the same pass)
    return owner
