"""CPU restatement (numpy, fp64) of the reference's GMG / FAC V-cycle hot path.

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
import this module; the product path (pressurepoissonsolver_b200/) never does.

Parity status: PINNED.  Every function here is checked against outputs of the reference's own
sources compiled in this container (oracle/_ref/ref_gmg, see oracle/Makefile and
tests/golden/make_golden.py -> tests/golden/*.npz) by tests/test_oracle_vs_reference.py.

All file:line citations are relative to /root/reference/.

Layout: a level vector is a numpy array of shape [P, n, n] (2D, index [p, y, x]) or
[P, n, n, n] (3D, index [p, z, y, x]); flattened it is the reference's patch-contiguous,
x-fastest storage (src/Thunderegg/PetscVector.h:75-90).  P is ordered by PatchInfo::local_index.
"""
import struct
from collections import deque

import numpy as np

# --------------------------------------------------------------------------------------------
# mesh: Tree<D> (.bin reader + refineLeaves) and per-level domain extraction
# --------------------------------------------------------------------------------------------


class Node:
    __slots__ = ("id", "level", "parent", "lengths", "starts", "nbr_id", "child_id")

    def has_children(self):
        return self.child_id[0] != -1


def orthants_on_side(D, s):
    """Orthant<D>::getValuesOnSide (src/Thunderegg/Side.h:346-362)."""
    bit = s // 2
    set_bit = s % 2
    out = []
    for i in range(1 << (D - 1)):
        lower = i & ((1 << bit) - 1)
        upper = (i >> bit) << (bit + 1)
        out.append(upper | lower | (set_bit << bit))
    return out


class Tree:
    """Tree<D> (src/Thunderegg/OctTree.h:33-213)."""

    def __init__(self, D):
        self.D = D
        self.nodes = {}
        self.levels = {}
        self.root = None
        self.num_levels = 0
        self.max_id = 0

    @classmethod
    def load(cls, path, D):
        """Tree<D>::Tree(std::string) (src/Thunderegg/OctTree.h:90-118); format SURVEY App. B."""
        t = cls(D)
        with open(path, "rb") as fh:
            buf = fh.read()
        num_nodes, _num_trees = struct.unpack_from("<ii", buf, 0)
        off = 8
        ns, no = 2 * D, 1 << D
        for i in range(num_nodes):
            n = Node()
            n.id, n.level, n.parent = struct.unpack_from("<iii", buf, off)
            off += 12
            n.lengths = list(struct.unpack_from("<%dd" % D, buf, off))
            off += 8 * D
            n.starts = list(struct.unpack_from("<%dd" % D, buf, off))
            off += 8 * D
            n.nbr_id = list(struct.unpack_from("<%di" % ns, buf, off))
            off += 4 * ns
            n.child_id = list(struct.unpack_from("<%di" % no, buf, off))
            off += 4 * no
            if i == 0:
                t.root = n.id
            t.max_id = max(t.max_id, n.id)
            t.nodes[n.id] = n
            t.num_levels = max(t.num_levels, n.level)
            t.levels[n.level] = n.id
        return t

    def refine_leaves(self):
        """Tree<D>::refineLeaves (src/Thunderegg/OctTree.h:119-179)."""
        D = self.D
        nodes = self.nodes
        child = nodes[self.root]
        level = 0
        while child.has_children():
            child = nodes[child.child_id[0]]
            level += 1
        q = deque([(level, child.id)])
        qed = {(level, child.id)}
        while q:
            level, nid = q.popleft()
            n = nodes[nid]
            for s in range(2 * D):
                if n.nbr_id[s] == -1 and n.parent != -1 and nodes[n.parent].nbr_id[s] != -1:
                    p = (level - 1, nodes[n.parent].nbr_id[s])
                    if p not in qed:
                        q.append(p)
                        qed.add(p)
                elif n.nbr_id[s] != -1 and nodes[n.nbr_id[s]].has_children():
                    nbr = nodes[n.nbr_id[s]]
                    for o in orthants_on_side(D, s ^ 1):
                        p = (level + 1, nbr.child_id[o])
                        if p not in qed:
                            q.append(p)
                            qed.add(p)
                elif n.nbr_id[s] != -1:
                    p = (level, n.nbr_id[s])
                    if p not in qed:
                        q.append(p)
                        qed.add(p)
        for _lvl, nid in sorted(qed):  # std::set<pair<int,int>> iteration order
            self._refine_node(nodes[nid])
        self.levels[self.num_levels + 1] = nodes[self.levels[self.num_levels]].child_id[0]
        self.num_levels += 1

    def _refine_node(self, n):
        """Tree<D>::refineNode (src/Thunderegg/OctTree.h:180-213), Node(parent, o) (OctNode.h:78-90)."""
        D = self.D
        kids = []
        for o in range(1 << D):
            c = Node()
            c.parent = n.id
            c.level = n.level + 1
            c.nbr_id = [-1] * (2 * D)
            c.child_id = [-1] * (1 << D)
            c.lengths = [n.lengths[i] / 2 for i in range(D)]
            c.starts = [n.starts[i] if not (o >> i) & 1 else n.starts[i] + c.lengths[i] for i in range(D)]
            self.max_id += 1
            c.id = self.max_id
            n.child_id[o] = c.id
            kids.append(c)
        for o in range(1 << D):
            for i in range(D):  # interior sides: upper side on axis i if bit i clear
                s = 2 * i + (0 if (o >> i) & 1 else 1)
                kids[o].nbr_id[s] = kids[o ^ (1 << i)].id
        for s in range(2 * D):
            if n.nbr_id[s] != -1 and self.nodes[n.nbr_id[s]].has_children():
                nbr = self.nodes[n.nbr_id[s]]
                for o in orthants_on_side(D, s):
                    child = kids[o]
                    nbr_child = self.nodes[nbr.child_id[o ^ (1 << (s // 2))]]
                    child.nbr_id[s] = nbr_child.id
                    nbr_child.nbr_id[s ^ 1] = child.id
        for c in kids:
            self.nodes[c.id] = c


NBR_NONE, NBR_NORMAL, NBR_COARSE, NBR_FINE = -1, 0, 1, 2


class Level:
    """Flat per-level metadata in local_index order: the restatement of Domain<D> /
    PatchInfo<D> / NbrInfo (src/Thunderegg/Domain.h:45-279, PatchInfo.h:74-637)."""

    def __init__(self, D, n, P):
        Q = 1 << (D - 1)
        self.D, self.n, self.P, self.Q = D, n, P, Q
        self.ids = np.zeros(P, np.int32)
        self.refine_level = np.zeros(P, np.int32)
        self.parent_id = np.zeros(P, np.int32)
        self.orth_on_parent = np.full(P, -1, np.int32)
        self.parent_idx = np.full(P, -1, np.int32)
        self.neumann = np.zeros(P, np.int32)
        self.starts = np.zeros((P, D))
        self.spacings = np.zeros((P, D))
        self.nbr_type = np.full((P, 2 * D), NBR_NONE, np.int32)
        self.nbr_ids = np.full((P, 2 * D, Q), -1, np.int32)
        self.nbr_idx = np.full((P, 2 * D, Q), -1, np.int32)
        self.orth_on_coarse = np.full((P, 2 * D), -1, np.int32)

    @property
    def shape(self):
        return (self.P,) + (self.n,) * self.D

    @property
    def cells(self):
        return self.P * self.n ** self.D


def extract_levels(tree, n):
    """ThundereggDomGen<D>::extractLevel for curr_level = num_levels..1
    (src/Thunderegg/ThundereggDomGen.h:127-222) followed by Domain<D>::indexDomainsLocal
    (src/Thunderegg/Domain.h:281-376) on one rank.  Returns the levels finest first."""
    D = tree.D
    nodes = tree.nodes
    levels = []
    for curr_level in range(tree.num_levels, 0, -1):
        start = tree.levels[curr_level]
        q = deque([start])
        qed = {start}
        info = {}
        while q:
            nd = nodes[q.popleft()]
            rec = {"id": nd.id, "refine_level": nd.level, "starts": nd.starts,
                   "spacings": [nd.lengths[i] / n for i in range(D)], "orth_on_parent": -1,
                   "nbr": [None] * (2 * D)}
            if nd.level < curr_level:
                rec["parent_id"] = nd.id
            else:
                rec["parent_id"] = nd.parent
                if nd.parent != -1:
                    rec["orth_on_parent"] = nodes[nd.parent].child_id.index(nd.id)
            for s in range(2 * D):
                if nd.nbr_id[s] == -1 and nd.parent != -1 and nodes[nd.parent].nbr_id[s] != -1:
                    parent = nodes[nd.parent]
                    nbr = nodes[parent.nbr_id[s]]
                    octs = orthants_on_side(D, s)
                    quad = [parent.child_id[o] for o in octs].index(nd.id)
                    rec["nbr"][s] = (NBR_COARSE, [nbr.id], quad)
                    new = [nbr.id]
                elif nd.level < curr_level and nd.nbr_id[s] != -1 and nodes[nd.nbr_id[s]].has_children():
                    nbr = nodes[nd.nbr_id[s]]
                    ids = [nbr.child_id[o] for o in orthants_on_side(D, s ^ 1)]
                    rec["nbr"][s] = (NBR_FINE, ids, -1)
                    new = ids
                elif nd.nbr_id[s] != -1:
                    rec["nbr"][s] = (NBR_NORMAL, [nd.nbr_id[s]], -1)
                    new = [nd.nbr_id[s]]
                else:
                    new = []
                for i in new:
                    if i not in qed:
                        q.append(i)
                        qed.add(i)
            info[nd.id] = rec
        # local index = BFS from the lowest id over getNbrIds() order (Domain.h:325-360)
        order = []
        todo = set(info)
        enq = set()
        while todo:
            first = min(todo)
            bq = deque([first])
            enq.add(first)
            while bq:
                i = bq.popleft()
                todo.discard(i)
                order.append(i)
                for nb in info[i]["nbr"]:
                    if nb is None:
                        continue
                    for j in nb[1]:
                        if j not in enq:
                            enq.add(j)
                            bq.append(j)
        rev = {pid: k for k, pid in enumerate(order)}
        L = Level(D, n, len(order))
        for k, pid in enumerate(order):
            r = info[pid]
            L.ids[k] = pid
            L.refine_level[k] = r["refine_level"]
            L.parent_id[k] = r["parent_id"]
            L.orth_on_parent[k] = r["orth_on_parent"]
            L.starts[k] = r["starts"]
            L.spacings[k] = r["spacings"]
            for s in range(2 * D):
                nb = r["nbr"][s]
                if nb is None:
                    continue
                L.nbr_type[k, s] = nb[0]
                L.orth_on_coarse[k, s] = nb[2]
                for qi, j in enumerate(nb[1]):
                    L.nbr_ids[k, s, qi] = j
                    L.nbr_idx[k, s, qi] = rev[j]
        levels.append(L)
    for fine, coarse in zip(levels[:-1], levels[1:]):
        rev = {int(pid): k for k, pid in enumerate(coarse.ids)}
        fine.parent_idx[:] = [rev[int(p)] for p in fine.parent_id]
    return levels


def build_hierarchy(mesh_path, D, n, divide=0, neumann=False):
    """neumann=True: ThundereggDomGen(tree, ns, neumann=true) marks every side without a neighbour
    (ThundereggDomGen.h:216-220, PatchInfo::setNeumann PatchInfo.h:684-697)."""
    t = Tree.load(mesh_path, D)
    for _ in range(divide):
        t.refine_leaves()
    levels = extract_levels(t, n)
    if neumann:
        for L in levels:
            for s in range(2 * D):
                L.neumann[L.nbr_type[:, s] == NBR_NONE] |= 1 << s
    return levels


# --------------------------------------------------------------------------------------------
# face slices and interface ("gamma") values = the reference's ghost fill
# --------------------------------------------------------------------------------------------


def face(u, D, s, offset=0):
    """LocalData<D>::getSliceOnSide(s, offset) (src/Thunderegg/Vector.h:153-177) for all patches:
    drops axis s/2; the remaining axes keep their order (numpy order: slowest first)."""
    n = u.shape[-1]
    ax = s // 2
    idx = offset if s % 2 == 0 else n - 1 - offset
    sl = [slice(None)] * (D + 1)
    sl[D - ax] = idx  # numpy axis of cartesian axis `ax` (x is last)
    return u[tuple(sl)]


def _f2f(sl, D):
    """fine_to_fine: TriLinInterp.cpp:85-98 / BilinearInterpolator.cpp:95-103."""
    out = np.empty_like(sl)
    if D == 2:
        out[..., 0::2] = 5.0 / 6 * sl[..., 0::2] - 1.0 / 6 * sl[..., 1::2]
        out[..., 1::2] = 5.0 / 6 * sl[..., 1::2] - 1.0 / 6 * sl[..., 0::2]
        return out
    a = sl[..., 0::2, 0::2]
    b = sl[..., 0::2, 1::2]  # x+1 (first face axis is the fastest index)
    c = sl[..., 1::2, 0::2]
    d = sl[..., 1::2, 1::2]
    out[..., 0::2, 0::2] = (11 * a - b - c - d) / 12.0
    out[..., 0::2, 1::2] = (-a + 11 * b - c - d) / 12.0
    out[..., 1::2, 0::2] = (-a - b + 11 * c - d) / 12.0
    out[..., 1::2, 1::2] = (-a - b - c + 11 * d) / 12.0
    return out


def _c2f(sl, D, orth):
    """coarse_to_fine(orth): TriLinInterp.cpp:99-131 / BilinearInterpolator.cpp:104-115.
    `sl` is the coarse patch's face; returns the fine-resolution contribution."""
    n = sl.shape[-1]
    i = (np.arange(n) + (orth & 1) * n) // 2
    if D == 2:
        return 2.0 / 6 * sl[..., i]
    j = (np.arange(n) + ((orth >> 1) & 1) * n) // 2
    return 4.0 * sl[..., j[:, None], i[None, :]] / 12.0


def _f2c_add(out, sl, D, orth):
    """fine_to_coarse(orth): TriLinInterp.cpp:139-170 / BilinearInterpolator.cpp:82-94.
    Adds the fine face `sl` into the coarse-resolution interface `out`, in the reference's
    loop order (yi outer, xi inner)."""
    n = sl.shape[-1]
    h = n // 2
    ox = (orth & 1) * h
    if D == 2:
        out[..., ox:ox + h] += 1.0 / 3 * sl[..., 0::2] + 1.0 / 3 * sl[..., 1::2]
        return
    oy = ((orth >> 1) & 1) * h
    tgt = out[..., oy:oy + h, ox:ox + h]
    for dy in (0, 1):
        for dx in (0, 1):
            tgt += 1.0 / 6.0 * sl[..., dy::2, dx::2]


def interface_values(L, u):
    """gamma[p, s] = the interface value aligned with side s of patch p, i.e. what
    StarPatchOp reads via sinfo.getIfaceLocalIndex(s) (src/Thunderegg/StarPatchOp.h:41-42,
    SchurInfo.h:554-557) after SchurHelper has summed every patch's contributions
    (src/Thunderegg/SchurHelper.h:319-327,361-368).  Contribution types and weights per
    SURVEY App. A.2 (IfaceType.h:58-62, SchurInfo.h:253-259,363-370)."""
    D, n, P = L.D, L.n, L.P
    fshape = (n,) * (D - 1)
    gamma = np.zeros((P, 2 * D) + fshape)
    w_normal = 0.5
    w_c2c = 1.0 / 3 if D == 2 else 2.0 / 6.0
    for s in range(2 * D):
        own = face(u, D, s)
        opp = face(u, D, s ^ 1)
        t = L.nbr_type[:, s]
        # Normal <-> Normal
        p = np.nonzero(t == NBR_NORMAL)[0]
        if p.size:
            nb = L.nbr_idx[p, s, 0]
            gamma[p, s] = w_normal * own[p] + w_normal * opp[nb]
        # I am fine, neighbour is coarse: fine_to_fine(own) + coarse_to_fine(coarse nbr)
        p = np.nonzero(t == NBR_COARSE)[0]
        for orth in range(L.Q):
            pp = p[L.orth_on_coarse[p, s] == orth]
            if pp.size:
                nb = L.nbr_idx[pp, s, 0]
                gamma[pp, s] = _f2f(own[pp], D) + _c2f(opp[nb], D, orth)
        # I am coarse, neighbours are fine: coarse_to_coarse(own) + sum fine_to_coarse
        p = np.nonzero(t == NBR_FINE)[0]
        if p.size:
            g = w_c2c * own[p]
            for orth in range(L.Q):
                nb = L.nbr_idx[p, s, orth]
                _f2c_add(g, opp[nb], D, orth)
            gamma[p, s] = g
    return gamma


# --------------------------------------------------------------------------------------------
# operator, patch solver (smoother), transfers
# --------------------------------------------------------------------------------------------


def _is_neumann(L, s):
    return (L.neumann >> s) & 1 == 1


def apply_op(L, u, gamma=None):
    """SchurHelper<D>::apply (src/Thunderegg/SchurHelper.h:361-376) with
    StarPatchOp<D>::applyWithInterface (src/Thunderegg/StarPatchOp.h:28-184)."""
    D, n = L.D, L.n
    if gamma is None:
        gamma = interface_values(L, u)
    out = np.zeros_like(u)
    for ax in range(D):
        h2 = (L.spacings[:, ax] ** 2).reshape((-1,) + (1,) * D)
        npax = D - ax
        term = np.zeros_like(u)
        mid = [slice(None)] * (D + 1)
        lo = list(mid)
        hi = list(mid)
        mid[npax] = slice(1, n - 1)
        lo[npax] = slice(0, n - 2)
        hi[npax] = slice(2, n)
        term[tuple(mid)] = (u[tuple(lo)] - 2 * u[tuple(mid)] + u[tuple(hi)]) / h2
        h2f = (L.spacings[:, ax] ** 2).reshape((-1,) + (1,) * (D - 1))
        for s in (2 * ax, 2 * ax + 1):
            m = face(u, D, s)
            inner = face(u, D, s, 1)
            has = (L.nbr_type[:, s] != NBR_NONE).reshape((-1,) + (1,) * (D - 1))
            neu = _is_neumann(L, s).reshape((-1,) + (1,) * (D - 1))
            if s % 2 == 0:
                with_nbr = (2 * gamma[:, s] - 3 * m + inner) / h2f
                neum = (-m + inner) / h2f
                diri = (-3 * m + inner) / h2f
            else:
                with_nbr = (inner - 3 * m + 2 * gamma[:, s]) / h2f
                neum = (inner - m) / h2f
                diri = (inner - 3 * m) / h2f
            face(term, D, s)[...] = np.where(has, with_nbr, np.where(neu, neum, diri))
        out = term if ax == 0 else out + term
    return out


def _transform_matrices(n):
    """DftPatchSolver<D>::getTransformArray (src/Thunderegg/PatchSolvers/DftPatchSolver.h:227-294).
    M[k, j] such that y_k = sum_j M[k, j] x_j (dgemv 'T' on the column-major array, :339-342)."""
    k = np.arange(n)[:, None]
    j = np.arange(n)[None, :]
    m = {}
    m["DST_II"] = np.sin(np.pi / n * ((k + 1) * (j + 0.5)))
    d3 = np.sin(np.pi / n * ((k + 0.5) * (j + 1)))
    d3[:, n - 1] = np.where(np.arange(n) % 2 == 0, 0.5, -0.5)
    m["DST_III"] = d3
    m["DCT_II"] = np.cos(np.pi / n * (k * (j + 0.5)))
    c3 = np.cos(np.pi / n * ((k + 0.5) * j))
    c3[:, 0] = 0.5
    m["DCT_III"] = c3
    m["DCT_IV"] = np.cos(np.pi / n * ((k + 0.5) * (j + 0.5)))
    m["DST_IV"] = np.sin(np.pi / n * ((k + 0.5) * (j + 0.5)))
    return m


def _axis_kind(neu_lo, neu_hi):
    """transform / eigenvalue choice per axis (DftPatchSolver.h:115-127,150-165)."""
    if neu_lo and neu_hi:
        return "DCT_II", "DCT_III", 0.0
    if neu_lo:
        return "DCT_IV", "DCT_IV", 0.5
    if neu_hi:
        return "DST_IV", "DST_IV", 0.5
    return "DST_II", "DST_III", 1.0


def patch_solve(L, rhs, lam=0.0):
    """DftPatchSolver<D>::solve without the gamma part (DftPatchSolver.h:173-216): forward
    transforms per axis, divide by the eigenvalues, inverse transforms, scale by (2/n)^D."""
    D, n = L.D, L.n
    mats = _transform_matrices(n)
    out = np.empty_like(rhs)
    kk = np.arange(n)
    for bits in np.unique(L.neumann):
        sel = np.nonzero(L.neumann == bits)[0]
        x = rhs[sel]
        kinds = [_axis_kind((bits >> (2 * a)) & 1, (bits >> (2 * a + 1)) & 1) for a in range(D)]
        # spacing: DftPatchSolver uses spacings[0] for every axis (:148); patches are cubic.
        h = L.spacings[sel, 0]
        denom = np.zeros((sel.size,) + (n,) * D)
        for a in range(D):
            lam_a = np.sin((kk + kinds[a][2]) * np.pi / (2 * n)) ** 2
            shp = [1] * (D + 1)
            shp[D - a] = n
            denom = denom - (4 / (h * h)).reshape((-1,) + (1,) * D) * lam_a.reshape(shp)
        denom = denom + lam
        for a in range(D):
            x = np.moveaxis(np.tensordot(x, mats[kinds[a][0]], axes=([D - a], [1])), -1, D - a)
        with np.errstate(divide="ignore", invalid="ignore"):  # zero mode of an all-Neumann patch, reset below
            x = x / denom
        if bits == (1 << (2 * D)) - 1:
            x.reshape(sel.size, -1)[:, 0] = 0
        for a in range(D):
            x = np.moveaxis(np.tensordot(x, mats[kinds[a][1]], axes=([D - a], [1])), -1, D - a)
        out[sel] = x * (2.0 / n) ** D
    return out


def smooth(L, f, u, lam=0.0):
    """SchurHelper<D>::solveWithSolution (src/Thunderegg/SchurHelper.h:319-331): gamma from the
    current u, StarPatchOp::addInterfaceToRHS (StarPatchOp.h:185-203), exact patch solves
    (lam: the patch solver's shift, DftPatchSolver.h:78,168)."""
    D = L.D
    gamma = interface_values(L, u)
    rhs = f.copy()
    for s in range(2 * D):
        has = L.nbr_type[:, s] != NBR_NONE
        h2 = (L.spacings[:, s // 2] ** 2).reshape((-1,) + (1,) * (D - 1))
        fs = face(rhs, D, s)
        fs[has] -= (2.0 / h2 * gamma[:, s])[has]
    return patch_solve(L, rhs, lam)


def jacobi(L, f, u, omega):
    """Weighted point-Jacobi sweep u <- u + omega D^-1 (f - A u) on the ghost-eliminated composite operator (the
    "weighted-Jacobi option" of BASELINE.json's north star; the reference has no such smoother, its operator is
    StarPatchOp.h:28-184).  D = diagonal of A: -(2 D)/h^2 per cell plus, for every patch side the cell touches, the
    coefficient of the cell itself in the ghost value 2 gamma - u (StarPatchOp.h:46-64 with the interface weights of
    SURVEY App. A.2): domain side -1 (Dirichlet) / +1 (Neumann); same-level neighbour 0; coarse neighbour 2 * 11/12 - 1 =
    5/6 (3D), 2 * 5/6 - 1 = 2/3 (2D); fine neighbours 2 * 1/3 - 1 = -1/3."""
    D, n = L.D, L.n
    diag = np.full(L.shape, -2.0 * D)
    for s in range(2 * D):
        ty = L.nbr_type[:, s]
        neu = _is_neumann(L, s)
        add = np.where(ty == NBR_NONE, np.where(neu, 1.0, -1.0), 0.0)
        add = np.where(ty == NBR_COARSE, 5.0 / 6.0 if D == 3 else 2.0 / 3.0, add)
        add = np.where(ty == NBR_FINE, -1.0 / 3.0, add)
        face(diag, D, s)[...] += add.reshape((-1,) + (1,) * (D - 1))
    h2 = (L.spacings[:, 0] ** 2).reshape((-1,) + (1,) * D)
    r = -1 * apply_op(L, u) + f
    return u + omega * r / (diag / h2)


def _child_view(a, D, n, orth):
    """[.., n/2 block] view of the parent's cells covered by child `orth` (bit i -> upper half
    on axis i; GMG/AvgRstr.h:91-93)."""
    h = n // 2
    sl = [slice(None)]
    for npax in range(1, D + 1):
        ax = D - npax
        o = (orth >> ax) & 1
        sl.append(slice(o * h, o * h + h))
    return a[tuple(sl)]


def restrict(fine, coarse, r):
    """GMG::AvgRstr<D>::restrict (src/Thunderegg/GMG/AvgRstr.h:78-113)."""
    D, n = fine.D, fine.n
    out = np.zeros(coarse.shape)
    copy = fine.orth_on_parent < 0
    out[fine.parent_idx[copy]] += r[copy]
    ref = np.nonzero(~copy)[0]
    for orth in range(1 << D):
        pp = ref[fine.orth_on_parent[ref] == orth]
        if not pp.size:
            continue
        tgt = np.zeros((pp.size,) + (n // 2,) * D)
        # reference loop order: coord[D-1] outermost ... coord[0] innermost, each += fine/2^D
        for off in range(1 << D):
            sl = [slice(None)] + [slice((off >> (D - npax)) & 1, None, 2) for npax in range(1, D + 1)]
            # iterate offsets so that x varies fastest: off bit0 = x
            tgt += r[pp][tuple(sl)] / (1 << D)
        view = _child_view(out, D, n, orth)
        view[fine.parent_idx[pp]] += tgt
    return out


def interpolate(fine, coarse, uc, uf):
    """GMG::DrctIntp<D>::interpolate (src/Thunderegg/GMG/DrctIntp.h:80-113): uf += P uc."""
    D, n = fine.D, fine.n
    out = uf.copy()
    copy = fine.orth_on_parent < 0
    out[copy] += uc[fine.parent_idx[copy]]
    ref = np.nonzero(~copy)[0]
    for orth in range(1 << D):
        pp = ref[fine.orth_on_parent[ref] == orth]
        if not pp.size:
            continue
        blk = _child_view(uc, D, n, orth)[fine.parent_idx[pp]]
        for ax in range(1, D + 1):
            blk = np.repeat(blk, 2, axis=ax)
        out[pp] += blk
    return out


# --------------------------------------------------------------------------------------------
# cycle and Krylov
# --------------------------------------------------------------------------------------------


def vcycle(levels, f, pre=1, post=1, coarse_sweeps=1, lvl=0, u=None):
    """GMG::Cycle<D>::apply + VCycle<D>::visit (src/Thunderegg/GMG/Cycle.h:56-126,
    GMG/VCycle.h:44-62): zero initial guess, r = f - A u restricted, DrctIntp added back."""
    L = levels[lvl]
    if u is None:
        u = np.zeros(L.shape)
    if lvl == len(levels) - 1:
        for _ in range(coarse_sweeps):
            u = smooth(L, f, u)
        return u
    for _ in range(pre):
        u = smooth(L, f, u)
    r = apply_op(L, u)
    r = -1 * r + f  # scaleThenAdd(-1, f), Vector.h:253-262
    fc = restrict(L, levels[lvl + 1], r)
    uc = vcycle(levels, fc, pre, post, coarse_sweeps, lvl + 1)
    u = interpolate(L, levels[lvl + 1], uc, u)
    for _ in range(post):
        u = smooth(L, f, u)
    return u


def active_levels(levels, max_levels=0, patches_per_proc=0.0, nranks=1):
    """The level list GMG::CycleFactory3d::getCycle builds (GMG/CycleFactory3d.cpp:99-104): at most max_levels levels
    (0 = all), stopping before a level with fewer than patches_per_proc patches per rank."""
    out = [levels[0]]
    for L in levels[1:]:
        if max_levels > 0 and len(out) >= max_levels:
            break
        if L.P / nranks < patches_per_proc:
            break
        out.append(L)
    return out


def cycle(levels, f, cycle_type="V", pre=1, post=1, mid=1, coarse_sweeps=1, max_levels=0, patches_per_proc=0.0,
          lam=0.0, interp=None):
    """GMG::Cycle<D>::apply (GMG/Cycle.h:116-126) with VCycle::visit (GMG/VCycle.h:44-62) or WCycle::visit
    (GMG/WCycle.h:45-68): prepCoarser = residual + restriction into fresh zero vectors (Cycle.h:56-68), prepFiner =
    interpolation of the coarse correction at the END of the coarser visit (Cycle.h:74-80, WCycle.h:66)."""
    lv = active_levels(levels, max_levels, patches_per_proc)
    interp = interp or interpolate

    def visit(l, fl, u):
        L = lv[l]
        if l == len(lv) - 1:
            for _ in range(coarse_sweeps):
                u = smooth(L, fl, u, lam)
            return u

        def coarse_correction(u):
            r = -1 * apply_op(L, u) + fl
            fc = restrict(L, lv[l + 1], r)
            uc = visit(l + 1, fc, np.zeros(lv[l + 1].shape))
            return interp(L, lv[l + 1], uc, u)

        for _ in range(pre):
            u = smooth(L, fl, u, lam)
        u = coarse_correction(u)
        if cycle_type == "W":
            for _ in range(mid):
                u = smooth(L, fl, u, lam)
            u = coarse_correction(u)
        for _ in range(post):
            u = smooth(L, fl, u, lam)
        return u

    return visit(0, f, np.zeros(lv[0].shape))


def interpolate_trilinear(fine, coarse, uc, uf):
    """Piecewise (bi/tri)linear prolongation uf += P uc: the interpolator the dead GMG/TriLinIntp.cpp:110-181 restates
    with coefficient tables (interior 27/9/9/3/9/3/3/1 over 64; on the parent patch's own boundary the one-sided
    45/15/15/5/-9/-3/-3/-1 over 64, TriLinIntp.cpp:185-190), i.e. the tensor product of the 1-D rule
        fine cell i -> coarse cell c = (i + offset)/2, partner c -/+ 1 towards the fine cell:  3/4 u_c + 1/4 u_partner,
        partner outside the parent patch: linear extrapolation 5/4 u_c - 1/4 u_(c +/- 1 on the other side).
    Patches present on both levels are copied (TriLinIntp.cpp:634-642).  Reproduces linear fields exactly
    (test/GMG.cpp:465-600)."""
    D, n = fine.D, fine.n
    out = uf.copy()
    copy = fine.orth_on_parent < 0
    out[copy] += uc[fine.parent_idx[copy]]
    i = np.arange(n)
    for p in np.nonzero(~copy)[0]:
        src = uc[fine.parent_idx[p]]
        orth = int(fine.orth_on_parent[p])
        val = src
        for ax in range(D):  # ax 0 = x = last numpy axis
            off = ((orth >> ax) & 1) * (n // 2)
            c = off + i // 2
            partner = np.where(i % 2 == 1, c + 1, c - 1)
            inside = (partner >= 0) & (partner < n)
            other = np.where(i % 2 == 1, c - 1, c + 1)
            wc = np.where(inside, 0.75, 1.25)
            wp = np.where(inside, 0.25, -0.25)
            q = np.where(inside, partner, other)
            npax = val.ndim - 1 - ax
            a = np.take(val, c, axis=npax)
            b = np.take(val, q, axis=npax)
            shp = [1] * val.ndim
            shp[npax] = n
            val = a * wc.reshape(shp) + b * wp.reshape(shp)
        out[p] += val
    return out


def vcycle_history(levels, f, ncyc):
    """stationary iteration u += V(f - A u); history of ||f - A u||_2 / ||f||_2."""
    L = levels[0]
    u = np.zeros(L.shape)
    fn = np.sqrt(np.sum(f * f))
    hist = []
    for k in range(ncyc + 1):
        r = -1 * apply_op(L, u) + f
        hist.append(np.sqrt(np.sum(r * r)) / fn)
        if k == ncyc:
            break
        u = u + vcycle(levels, r)
    return u, np.array(hist)


def bicgstab(levels, f, tol=1e-12, max_it=1000, precond=True):
    """BiCGStab<D>::solve with Mr = V-cycle (src/Thunderegg/BiCGStab.h:45-106)."""
    L = levels[0]
    A = lambda v: apply_op(L, v)
    M = (lambda v: vcycle(levels, v)) if precond else (lambda v: v)
    dot = lambda a, b: float(np.sum(a * b))
    x = np.zeros(L.shape)
    resid = -1 * A(x) + f
    r0 = np.sqrt(dot(resid, resid))
    rhat = resid.copy()
    p = resid.copy()
    rho = dot(rhat, resid)
    its = 0
    while np.sqrt(dot(resid, resid)) / r0 > tol and its < max_it:
        mp = M(p)
        ap = A(mp)
        alpha = rho / dot(rhat, ap)
        s = resid + (-alpha) * ap
        ms = M(s)
        as_ = A(ms)
        omega = dot(as_, s) / dot(as_, as_)
        x = x + (mp * alpha + ms * omega)
        resid = resid + (ap * -alpha + as_ * -omega)
        rho_new = dot(resid, rhat)
        beta = rho_new * alpha / (rho * omega)
        p = p + ap * -omega
        p = beta * p + resid
        its += 1
        rho = rho_new
    return x, its


# --------------------------------------------------------------------------------------------
# manufactured problem (apps/3d/steady.cpp:253-265, apps/2d/steady.cpp:314-316, Init.cpp)
# --------------------------------------------------------------------------------------------


def _g3(x, y, z):
    x, y, z = x + .3, y + .3, z + .3
    return np.sin(np.pi * x) * np.cos(2.0 / 3 * np.pi * y) * np.sin(5.0 / 6 * np.pi * z)


def _f3(x, y, z):
    x, y, z = x + .3, y + .3, z + .3
    return -77.0 / 36 * np.pi * np.pi * np.sin(np.pi * x) * np.cos(2.0 / 3 * np.pi * y) * np.sin(5.0 / 6 * np.pi * z)


def _g2(x, y):
    return np.sin(np.pi * y) * np.cos(2 * np.pi * x)


def _f2(x, y):
    return -5 * np.pi * np.pi * np.sin(np.pi * y) * np.cos(2 * np.pi * x)


def trig_rhs(L):
    """Init::initDirichlet / initDirichlet2d (apps/shared/Init.cpp:152-245,305-361): f at cell
    centres, Dirichlet data folded in as f -= 2 g(face)/h^2 on sides without a neighbour.
    Returns (f, exact)."""
    D, n, P = L.D, L.n, L.P
    k = np.arange(n)
    h = L.spacings
    cen = [L.starts[:, a, None] + h[:, a, None] / 2.0 + h[:, a, None] * k[None, :] for a in range(D)]
    lo = [L.starts[:, a] for a in range(D)]
    hi = [L.starts[:, a] + h[:, a] * n for a in range(D)]
    g, ff = (_g2, _f2) if D == 2 else (_g3, _f3)

    def grid(coords):  # coords[a]: [P, n] or [P, 1] -> broadcast to [P, (z), y, x]
        out = []
        for a in range(D):
            shp = [P] + [1] * D
            shp[D - a] = coords[a].shape[1]
            out.append(coords[a].reshape(shp))
        return out

    f = ff(*grid(cen)) * np.ones(L.shape)
    exact = g(*grid(cen)) * np.ones(L.shape)
    for s in range(2 * D):
        a = s // 2
        none = L.nbr_type[:, s] == NBR_NONE
        if not none.any():
            continue
        coords = list(cen)
        coords[a] = (lo[a] if s % 2 == 0 else hi[a])[:, None]
        bshape = [P] + [1 if (D - npax) == a else n for npax in range(1, D + 1)]
        bnd = np.squeeze(np.broadcast_to(g(*grid(coords)), bshape), axis=D - a)
        h2 = (h[:, a] ** 2).reshape((-1,) + (1,) * (D - 1))
        fs = face(f, D, s)
        fs[none] -= (2 * bnd / h2)[none]
    return f, exact


# --------------------------------------------------------------------------------------------
# reference metadata dump reader (format: oracle/ref_driver.cpp dump_meta)
# --------------------------------------------------------------------------------------------


def read_ref_meta(path):
    buf = open(path, "rb").read()
    magic, D, n, nlev = struct.unpack_from("<iiii", buf, 0)
    assert magic == 0x474d4731
    off = 16
    Q = 1 << (D - 1)
    K = 6 + 2 * D * (2 + 2 * Q)
    levels = []
    for _ in range(nlev):
        (P,) = struct.unpack_from("<i", buf, off)
        off += 4
        ints = np.frombuffer(buf, np.int32, P * K, off).reshape(P, K)
        off += 4 * P * K
        reals = np.frombuffer(buf, np.float64, P * 2 * D, off).reshape(P, 2 * D)
        off += 8 * P * 2 * D
        L = Level(D, n, P)
        L.ids[:] = ints[:, 0]
        L.refine_level[:] = ints[:, 1]
        L.parent_id[:] = ints[:, 2]
        L.orth_on_parent[:] = ints[:, 3]
        L.parent_idx[:] = ints[:, 4]
        L.neumann[:] = ints[:, 5]
        side = ints[:, 6:].reshape(P, 2 * D, 2 + 2 * Q)
        L.nbr_type[:] = side[:, :, 0]
        L.orth_on_coarse[:] = side[:, :, 1]
        L.nbr_ids[:] = side[:, :, 2:2 + Q]
        L.nbr_idx[:] = side[:, :, 2 + Q:]
        L.starts[:] = reals[:, :D]
        L.spacings[:] = reals[:, D:]
        levels.append(L)
    return levels
