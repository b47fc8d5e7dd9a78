"""Mesh ingest without per-triangle Python objects.

The reference loads OBJ files through ``pywavefront`` into a list of ``Triangle`` objects
(``R/src/io.py:11-40``, stale constructor arguments) and the notebooks/app triangulate ``pyvista`` meshes
face by face (``R/app.py:218-228``).  Here an OBJ goes straight to a ``[N,3,3]`` vertex array and from
there, via ``from_arrays``, to ``lt_scene_create`` + the GPU LBVH — a 1 M-triangle mesh never becomes
1 M Python objects.
"""
from __future__ import annotations

import numpy as np


def load_obj(path, *, scale=1.0, translate=(0.0, 0.0, 0.0)) -> np.ndarray:
    """Wavefront OBJ → ``[N,3,3]`` float64 triangle vertices.  Polygons are fan-triangulated; negative
    (relative) indices and ``v/vt/vn`` index triples are accepted; everything but ``v`` and ``f`` is ignored."""
    verts, tris = [], []
    with open(path, "r", errors="replace") as fh:
        for line in fh:
            if line.startswith("v "):
                p = line.split()
                verts.append((float(p[1]), float(p[2]), float(p[3])))
            elif line.startswith("f "):
                idx = []
                for tok in line.split()[1:]:
                    k = int(tok.split("/")[0])
                    idx.append(k - 1 if k > 0 else len(verts) + k)
                for a in range(1, len(idx) - 1):
                    tris.append((idx[0], idx[a], idx[a + 1]))
    if not tris:
        raise ValueError(f"{path}: no faces")
    v = np.asarray(verts, dtype=np.float64)
    t = np.asarray(tris, dtype=np.int64)
    if t.min() < 0 or t.max() >= len(v):
        raise ValueError(f"{path}: face index out of range")
    out = v[t] * float(scale) + np.asarray(translate, dtype=np.float64)
    # drop degenerate faces: their normal is undefined (the reference would produce NaN normals)
    n = np.cross(out[:, 1] - out[:, 0], out[:, 2] - out[:, 0])
    return np.ascontiguousarray(out[np.linalg.norm(n, axis=1) > 0])
