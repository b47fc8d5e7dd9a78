"""Scene objects → structure-of-arrays (the host side of the drop-in boundary).

Reads, by attribute name only, the fields the reference's hot path reads
(SURVEY.md §8(a) row a11): ``PreComputedTriangle.vertex_1..3/normal/material/is_light``
(``R/src/primitives.py:85-113``), ``Material.color.diffuse/emission/ior/transmission/
is_diffuse/is_mirror`` (``R/src/material.py:18-37``), ``Light.source/normal/material/
total_area`` (``R/src/scene.py:6-17``).  Works on the reference's jitclass instances and on
the plain mirrors in ``api.py`` alike.  Vectors are homogeneous 4-vectors in the reference
(Q11: points w=1, directions w=0); they are packed as 3-vectors after checking w.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MAT_DIFFUSE = 1
MAT_MIRROR = 2


@dataclass
class PackedScene:
    tri_v0: np.ndarray
    tri_v1: np.ndarray
    tri_v2: np.ndarray
    tri_n: np.ndarray
    tri_mat: np.ndarray
    tri_light: np.ndarray
    mat_kd: np.ndarray
    mat_emission: np.ndarray
    mat_ior: np.ndarray
    mat_transmission: np.ndarray
    mat_flags: np.ndarray
    light_pos: np.ndarray
    light_n: np.ndarray
    light_lekd: np.ndarray
    light_area: np.ndarray
    extra: dict = field(default_factory=dict)

    @property
    def n_tris(self):
        return len(self.tri_mat)

    @property
    def n_lights(self):
        return len(self.light_area)

    def to_npz_dict(self, prefix=""):
        keys = ("tri_v0 tri_v1 tri_v2 tri_n tri_mat tri_light mat_kd mat_emission mat_ior "
                "mat_transmission mat_flags light_pos light_n light_lekd light_area").split()
        return {prefix + k: getattr(self, k) for k in keys}

    @staticmethod
    def from_npz_dict(d, prefix=""):
        keys = ("tri_v0 tri_v1 tri_v2 tri_n tri_mat tri_light mat_kd mat_emission mat_ior "
                "mat_transmission mat_flags light_pos light_n light_lekd light_area").split()
        return PackedScene(**{k: np.ascontiguousarray(d[prefix + k]) for k in keys})


def _vec3(v, w_expected, what):
    v = np.asarray(v, dtype=np.float64)
    if v.shape[0] == 4:
        # tolerance: generate_area_light_samples builds w as a barycentric sum (1-√a)+√a(1-b)+b√a,
        # which rounds to 1-1e-16; its effect on the reference's 4-component norms is O(1e-32).
        if abs(v[3] - w_expected) > 1e-12:
            raise ValueError(f"{what}: homogeneous w must be {w_expected}, got {v[3]}")
        return v[:3]
    if v.shape[0] != 3:
        raise ValueError(f"{what}: expected 3 or 4 components")
    return v


def _mat_key(m):
    kd = np.asarray(m.color.diffuse, dtype=np.float64)
    return (float(kd[0]), float(kd[1]), float(kd[2]), float(m.emission), float(m.ior),
            float(m.transmission), bool(m.is_diffuse), bool(m.is_mirror))


def pack_materials_lights(materials_of_tris, lights):
    """Deduplicate materials by value and pack the ``Light`` list."""
    table, index = {}, []
    for m in materials_of_tris:
        k = _mat_key(m)
        index.append(table.setdefault(k, len(table)))
    keys = list(table.keys())
    M = len(keys)
    mat_kd = np.array([k[0:3] for k in keys], dtype=np.float64).reshape(M, 3)
    mat_emission = np.array([k[3] for k in keys], dtype=np.float64)
    mat_ior = np.array([k[4] for k in keys], dtype=np.float64)
    mat_tr = np.array([k[5] for k in keys], dtype=np.float64)
    mat_flags = np.array([(MAT_DIFFUSE if k[6] else 0) | (MAT_MIRROR if k[7] else 0) for k in keys],
                         dtype=np.int32)
    L = len(lights)
    lp, ln, lk, la = np.zeros((L, 3)), np.zeros((L, 3)), np.zeros((L, 3)), np.zeros(L)
    for i, l in enumerate(lights):
        lp[i] = _vec3(l.source, 1.0, "Light.source")
        ln[i] = _vec3(l.normal, 0.0, "Light.normal")
        # R/src/light_samples.py:55: light.material.emission * light.material.color.diffuse
        lk[i] = l.material.emission * np.asarray(l.material.color.diffuse, dtype=np.float64)
        la[i] = l.total_area
    return (np.array(index, dtype=np.int32), mat_kd, mat_emission, mat_ior, mat_tr, mat_flags,
            lp, ln, lk, la)


def pack_scene(primitives, lights) -> PackedScene:
    N = len(primitives)
    v0, v1, v2, nn = (np.zeros((N, 3)) for _ in range(4))
    is_light = np.zeros(N, dtype=np.uint8)
    mats = []
    for i, p in enumerate(primitives):
        v0[i] = _vec3(p.vertex_1, 1.0, "vertex_1")
        v1[i] = _vec3(p.vertex_2, 1.0, "vertex_2")
        v2[i] = _vec3(p.vertex_3, 1.0, "vertex_3")
        nn[i] = _vec3(p.normal, 0.0, "normal")
        is_light[i] = 1 if p.is_light else 0
        mats.append(p.material)
    idx, kd, em, ior, tr, fl, lp, ln, lk, la = pack_materials_lights(mats, lights)
    return PackedScene(v0, v1, v2, nn, idx, is_light, kd, em, ior, tr, fl, lp, ln, lk, la)


def from_arrays(vertices, *, kd=(0.6, 0.6, 0.6), tri_mat=None, materials=None, lights=None,
                base: PackedScene | None = None) -> PackedScene:
    """Mesh ingest without per-triangle Python objects.

    ``vertices``: [N,3,3] f64.  Normals are computed as the reference's constructor does
    (``R/src/primitives.py:108-111``: ``cross(v2-v1, v3-v1)`` normalised).  When ``base`` is
    given its triangles, materials and lights are kept and the mesh is appended with one new
    diffuse material ``kd``."""
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    e1, e2 = v[:, 1] - v[:, 0], v[:, 2] - v[:, 0]
    n = np.cross(e1, e2)
    nl = n.astype(np.longdouble)
    n /= np.sqrt((nl * nl).sum(axis=1, keepdims=True)).astype(np.float64)   # dnrm2-style single rounding
    N = len(v)
    if base is None:
        raise ValueError("from_arrays needs a base scene providing lights (pass base=pack_scene(...))")
    M = len(base.mat_flags)
    return PackedScene(
        np.concatenate([base.tri_v0, v[:, 0]]), np.concatenate([base.tri_v1, v[:, 1]]),
        np.concatenate([base.tri_v2, v[:, 2]]), np.concatenate([base.tri_n, n]),
        np.concatenate([base.tri_mat, np.full(N, M, np.int32)]),
        np.concatenate([base.tri_light, np.zeros(N, np.uint8)]),
        np.concatenate([base.mat_kd, np.array([kd], dtype=np.float64)]),
        np.concatenate([base.mat_emission, [0.0]]), np.concatenate([base.mat_ior, [1.5]]),
        np.concatenate([base.mat_transmission, [0.0]]),
        np.concatenate([base.mat_flags, np.array([MAT_DIFFUSE], np.int32)]),
        base.light_pos, base.light_n, base.light_lekd, base.light_area)
