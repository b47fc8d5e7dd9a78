"""Host-side construction of the reference's *flattened* BVH node list, as arrays.

COMPAT traversal emulates ``intersect_bvh`` (``R/src/bvh_new.py:413-482``)
over exactly the node list the reference's own host code would produce, so this module
restates what ``build_bvh`` (``:148-278``, ``split_method = 1``) and ``flatten_bvh``
(``:281-300``) *effectively compute*, including:

* Q7 — the midpoint "partition" runs on a slice copy, so primitives are never reordered and
  ``mid`` is ``start + #(centroids left of the midpoint)``; leaves are contiguous ranges of
  the input order.
* Q8 — ``second_child_offset`` of an interior node is the index of the *last* node of its
  subtree, not of its second child.

This is setup code that runs once per scene on the host (the reference's is pure Python,
not on the hot path); the GPU LBVH builder replaces it for NEAREST traversal.
Also converts a reference ``typed.List[LinearBVHNode]`` into the same arrays.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class CompatNodes:
    bmin: np.ndarray      # [n,3] f64
    bmax: np.ndarray      # [n,3] f64
    prim_off: np.ndarray  # [n] i32, -1 on interior nodes (reference: None)
    second: np.ndarray    # [n] i32, -1 on leaves
    n_prims: np.ndarray   # [n] i32, 0 on interior nodes
    axis: np.ndarray      # [n] i32, -1 on leaves

    def __len__(self):
        return len(self.n_prims)


def build_compat_nodes(v0, v1, v2) -> CompatNodes:
    """``v0,v1,v2``: [N,3] f64 triangle vertices in primitive order."""
    v0, v1, v2 = (np.asarray(a, dtype=np.float64) for a in (v0, v1, v2))
    lo = np.minimum(np.minimum(v0, v1), v2)           # get_bounds, R/src/bvh_new.py:76-81
    hi = np.maximum(np.maximum(v0, v1), v2)
    cen = (lo + hi) / 2                               # AABB.centroid, R/src/primitives.py:80
    rows = []                                         # (bmin,bmax,prim_off,second,n,axis) preorder

    def rec(start, end):
        idx = len(rows)
        rows.append(None)
        n = end - start
        bmin, bmax = lo[start:end].min(axis=0), hi[start:end].max(axis=0)
        cmin, cmax = cen[start:end].min(axis=0), cen[start:end].max(axis=0)
        ext = np.abs(cmax - cmin)
        dim = 0 if (ext[0] > ext[1] and ext[0] > ext[2]) else (1 if ext[1] > ext[2] else 2)
        if n == 1 or cmax[dim] == cmin[dim]:
            rows[idx] = (bmin, bmax, start, -1, n, -1)
            return idx
        pmid = (cmin[dim] + cmax[dim]) / 2
        mid = start + int(np.count_nonzero(cen[start:end, dim] < pmid))
        if mid == start or mid == end:
            # (cmin + cmax) / 2 rounded onto cmin (adjacent floats): nothing lies left of the midpoint.  The reference's
            # build_bvh would recurse into an empty range here and fail; keep the primitives together in one leaf instead.
            rows[idx] = (bmin, bmax, start, -1, n, -1)
            return idx
        rec(start, mid)
        last = rec(mid, end)
        rows[idx] = (bmin, bmax, -1, last, 0, dim)
        return last

    # rec returns the index of the last node of the subtree it emitted (own index for a leaf)
    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        rec(0, len(v0))
    finally:
        sys.setrecursionlimit(old)
    n = len(rows)
    return CompatNodes(
        bmin=np.array([r[0] for r in rows], dtype=np.float64).reshape(n, 3),
        bmax=np.array([r[1] for r in rows], dtype=np.float64).reshape(n, 3),
        prim_off=np.array([r[2] for r in rows], dtype=np.int32),
        second=np.array([r[3] for r in rows], dtype=np.int32),
        n_prims=np.array([r[4] for r in rows], dtype=np.int32),
        axis=np.array([r[5] for r in rows], dtype=np.int32),
    )


def from_linear_bvh(linear_bvh) -> CompatNodes:
    """Read a reference-style list of ``LinearBVHNode`` (``R/src/bvh_new.py:53-67``)."""
    n = len(linear_bvh)
    out = CompatNodes(np.zeros((n, 3)), np.zeros((n, 3)), np.full(n, -1, np.int32),
                      np.full(n, -1, np.int32), np.zeros(n, np.int32), np.full(n, -1, np.int32))
    for i, nd in enumerate(linear_bvh):
        out.bmin[i] = np.asarray(nd.bounds.min_point)[:3]
        out.bmax[i] = np.asarray(nd.bounds.max_point)[:3]
        out.n_prims[i] = nd.n_primitives if nd.n_primitives is not None else 0
        if out.n_prims[i] > 0:
            out.prim_off[i] = nd.primitives_offset
        else:
            out.second[i] = nd.second_child_offset
            out.axis[i] = nd.axis
    return out


def compat_defined(nodes: CompatNodes) -> bool:
    """Q9b: the reference raises ``IndexError`` when a bounds-hit leaf holds a primitive index
    >= node count; COMPAT traversal is only *defined* when that cannot happen."""
    leaf = nodes.n_prims > 0
    if not leaf.any():
        return False
    return int((nodes.prim_off[leaf] + nodes.n_prims[leaf]).max()) <= len(nodes)
