"""Synthetic scenes of BASELINE.json's configs, restated from the reference's example driver.

Every builder takes ``ns`` — a namespace providing ``Color``, ``Material``,
``PreComputedTriangle``, ``Light`` and ``generate_area_light_samples``.  Pass
``light_transport_b200.api`` (default) for the plain-Python mirrors, or a namespace of the
reference's own jitclasses (the golden-vector generator does that) — the geometry and the
order of primitives are identical either way.

* ``lts_scene``      C1/C2/C5: ``R/examples/LTS.ipynb`` cells 11-17 (Cornell box of
  ``R/src/cornell_box.py:9-167`` + ``pv.Cone(radius=2, height=5)`` + 2 lamp triangles).
  pyvista is not available, so rectangles ``[a,b,c,d]`` are split ``(a,b,c),(a,c,d)`` and
  the cone is a 6-gon; the triangulation is parity-unpinned but both
  the oracle and the GPU path consume these same triangles.
* ``mirror_box``     C4a: closed box, 4-way split walls (so Q9b cannot fire), one ``Light``.
* ``heightfield``    C3: ~1 M-triangle displaced grid inside the Cornell box, as raw arrays.
"""
from __future__ import annotations

import numpy as np

from . import api as _api


def _tri(ns, a, b, c, mat, is_light=False):
    h = lambda p: np.ascontiguousarray(np.append(np.asarray(p, dtype=np.float64), 1.0))
    return ns.PreComputedTriangle(vertex_1=h(a), vertex_2=h(b), vertex_3=h(c), material=mat,
                                  is_light=is_light)


def _quad(ns, out, q, mat):
    a, b, c, d = q
    out.append(_tri(ns, a, b, c, mat))
    out.append(_tri(ns, a, c, d, mat))


def cone_vertices():
    """``pv.Cone(radius=2, center=0, height=5)`` → 10 triangles; pyvista stores float32 points
    (``LTS.ipynb`` cell 12 prints zmax=1.7320508), so the ring is rounded through float32."""
    apex = np.array([2.5, 0.0, 0.0])
    ring = [np.array([-2.5, np.float32(2 * np.cos(k * np.pi / 3)), np.float32(2 * np.sin(k * np.pi / 3))],
                     dtype=np.float64) for k in range(6)]
    tris = [(apex, ring[k], ring[(k + 1) % 6]) for k in range(6)]
    tris += [(ring[0], ring[k + 1], ring[k]) for k in range(1, 5)]
    return tris


def lts_scene(ns=_api, n_light_samples=1000, light_seed=0, new_list=list):
    """Returns ``(objects, light_sources, camera, depth)``; 28 primitives, 2*n lights."""
    dim = depth = 7.5
    A, B, C, D = [-dim, -dim, -dim], [-dim, -dim, dim], [dim, -dim, dim], [dim, -dim, -dim]
    E, F, G, H = [-dim, dim, -dim], [-dim, dim, dim], [dim, dim, dim], [dim, dim, -dim]
    I, J, K, L = [-1, dim, -dim], [-1, dim, -1], [-1, dim, 1], [-1, dim, dim]
    M, N, O, P = [1, dim, dim], [1, dim, 1], [1, dim, -1], [1, dim, -dim]
    mk = lambda col: ns.Material(color=col, shininess=30, reflection=0.1, ior=1.5210, transmission=1)
    surface, left, right = mk(ns.WHITE_2), mk(ns.RED), mk(ns.GREEN)   # LTS.ipynb cell 14
    objs = new_list()
    _quad(ns, objs, [D, C, G, H], right)
    _quad(ns, objs, [F, B, A, E], left)
    _quad(ns, objs, [E, A, D, H], surface)
    _quad(ns, objs, [A, B, C, D], surface)
    for q in ([H, G, M, P], [N, M, L, K], [P, O, J, I], [I, L, F, E]):
        _quad(ns, objs, q, surface)
    for a, b, c in cone_vertices():
        objs.append(_tri(ns, a, b, c, ns.GLASS_MAT))
    source_mat = ns.Material(color=ns.WHITE, shininess=1, reflection=0.9, ior=1.5, emission=200)
    light_1 = _tri(ns, [-1, depth, -1], [1, depth, 1], [-1, depth, 1], source_mat, True)
    light_2 = _tri(ns, [-1, depth, -1], [1, depth, -1], [1, depth, 1], source_mat, True)
    np.random.seed(light_seed)
    lights = ns.generate_area_light_samples(tri_1=light_1, tri_2=light_2, source_mat=source_mat,
                                            number_of_samples=n_light_samples, total_area=4)
    objs.append(light_1)
    objs.append(light_2)
    camera = np.array([0, 0, depth + 0.5, 1], dtype=np.float64)
    return objs, lights, camera, depth


def _quad4(ns, out, q, mat):
    """Split a rectangle 4-way about its centre → one primitive per BVH leaf (Q9b-safe)."""
    q = [np.asarray(p, dtype=np.float64) for p in q]
    ctr = (q[0] + q[1] + q[2] + q[3]) / 4
    for k in range(4):
        out.append(_tri(ns, q[k], q[(k + 1) % 4], ctr, mat))


def mirror_box(ns=_api, all_mirror=False, new_list=list):
    """C4a: 5 mirror walls + diffuse (or mirror) floor, lamp under the ceiling,
    exactly one ``Light`` so next-event estimation draws no random light index."""
    d = 5.0
    A, B, C, D = [-d, -d, -d], [-d, -d, d], [d, -d, d], [d, -d, -d]
    E, F, G, H = [-d, d, -d], [-d, d, d], [d, d, d], [d, d, -d]
    mirror = ns.Material(color=ns.WHITE_2, shininess=30, reflection=1.0, ior=1.5, is_diffuse=False,
                         is_mirror=True)
    floor = mirror if all_mirror else ns.Material(color=ns.WHITE_2, shininess=30, reflection=0.1, ior=1.5)
    objs = new_list()
    for q in ([D, C, G, H], [F, B, A, E], [E, A, D, H], [B, F, G, C], [H, G, F, E]):
        _quad4(ns, objs, q, mirror)
    _quad4(ns, objs, [A, B, C, D], floor)
    lamp = ns.Material(color=ns.WHITE, shininess=1, reflection=0.9, ior=1.5, emission=200)
    y = d - 0.25
    l1 = _tri(ns, [-1, y, -1], [1, y, 1], [-1, y, 1], lamp, True)
    l2 = _tri(ns, [-1, y, -1], [1, y, -1], [1, y, 1], lamp, True)
    objs.append(l1)
    objs.append(l2)
    src = np.array([0.25, y, 0.35, 1.0])
    lights = new_list()
    lights.append(ns.Light(source=src, material=lamp, normal=l1.normal, total_area=4))
    camera = np.array([0.3, 0.2, 5.0 - 1e-3, 1.0])
    return objs, lights, camera, 4.5


def heightfield(nx=1000, ny=500, seed=3, dim=7.5):
    """C3: ``2*nx*ny`` triangles (1 M at the default) of a displaced grid floating above the
    Cornell floor, all diffuse kd=0.6.  Returned as arrays (never as per-triangle objects):
    ``v[n,3,3]`` float64 vertices."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(-dim * 0.98, dim * 0.98, nx + 1)
    zs = np.linspace(-dim * 0.98, dim * 0.98, ny + 1)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    Y = -5.0 + 0.8 * np.sin(X * 1.3) * np.cos(Z * 0.9) + rng.uniform(-0.02, 0.02, X.shape)
    P = np.stack([X, Y, Z], axis=-1)
    p00, p10, p01, p11 = P[:-1, :-1], P[1:, :-1], P[:-1, 1:], P[1:, 1:]
    t1 = np.stack([p00, p01, p11], axis=-2).reshape(-1, 3, 3)
    t2 = np.stack([p00, p11, p10], axis=-2).reshape(-1, 3, 3)
    v = np.empty((t1.shape[0] * 2, 3, 3), dtype=np.float64)
    v[0::2] = t1
    v[1::2] = t2
    return v
