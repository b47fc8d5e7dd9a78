"""Host-side mirrors of the reference's scene-description types.

The reference describes a scene with numba ``jitclass`` objects
(``R/src/material.py:4-37``, ``R/src/primitives.py:85-113``, ``R/src/scene.py:6-73``;
``R/`` = ``/root/reference/LightTransportSimulator/light_transport/``).  Those classes
need numba + the reference tree, neither of which the GPU path depends on, so this
module offers plain-Python carriers with the *same constructor signatures and field
names*.  ``light_transport_b200.render_scene`` accepts either these or the reference's
own jitclass instances (it only reads attributes, see ``_pack.py``).

Only the fields the hot path reads are kept; see SURVEY.md §8(a) row a11.
"""
from __future__ import annotations

import math

import numpy as np

# R/src/constants.py:7-12
inv_pi = 1 / np.pi
EPSILON = 0.000001


def _f64(v):
    return np.ascontiguousarray(v, dtype=np.float64)


def _norm(v):
    """``np.linalg.norm`` as the reference's jit-compiled constructor evaluates it: numba lowers it to
    BLAS dnrm2, whose x86-64 kernel accumulates in extended precision and rounds once (DESIGN.md,
    "numerics"); NumPy's own norm rounds after every step and differs in the last bit."""
    v = np.asarray(v, dtype=np.longdouble)
    return float(np.sqrt(np.sum(v * v)))


class Color:
    """R/src/material.py:4-13."""

    def __init__(self, ambient, diffuse, specular):
        self.ambient = _f64(ambient)
        self.diffuse = _f64(diffuse)
        self.specular = _f64(specular)


class Material:
    """R/src/material.py:18-37 (same positional order, same defaults)."""

    def __init__(self, color, shininess, reflection, ior, emission=0.0, transmission=0.0,
                 is_diffuse=True, is_mirror=False):
        self.color = color
        self.shininess = float(shininess)
        self.reflection = float(reflection)
        self.ior = float(ior)
        self.emission = float(emission)
        self.transmission = float(transmission)
        self.is_diffuse = bool(is_diffuse)
        self.is_mirror = bool(is_mirror)


class PreComputedTriangle:
    """R/src/primitives.py:99-113: homogeneous (w=1) vertices, geometric normal with w=0.

    The 12-entry ``transformation`` table of the reference (:113-173) feeds only the unused
    ``.intersect`` method (:175-209) and is not reproduced.
    """

    def __init__(self, vertex_1, vertex_2, vertex_3, material, is_light=False):
        self.vertex_1 = _f64(vertex_1)
        self.vertex_2 = _f64(vertex_2)
        self.vertex_3 = _f64(vertex_3)
        self.material = material
        self.is_light = bool(is_light)
        self.centroid = (self.vertex_1 + self.vertex_2 + self.vertex_3) / 3
        self.edge_1 = self.vertex_2 - self.vertex_1
        self.edge_2 = self.vertex_3 - self.vertex_1
        n = np.cross(self.edge_1[:3], self.edge_2[:3])
        self.normal = np.append(n / _norm(n), 0.0)


class Light:
    """R/src/scene.py:12-17."""

    def __init__(self, source, material, normal, total_area):
        self.source = _f64(source)
        self.material = material
        self.normal = _f64(normal)
        self.total_area = float(total_area)


class Scene:
    """R/src/scene.py:53-73.

    Like the reference, the constructor pre-draws ``rand_0``/``rand_1`` of shape
    ``[H, W, spp, max_depth]`` from NumPy's global generator (Q15), their logits and the
    ``bounce_record`` / ``record_log_pdf`` side-effect arrays (``:70-73``) unless
    ``allocate=False`` is given; configs whose arrays would not fit (C2: 704 GB) pass
    ``allocate=False`` and render with ``rng="philox"``.
    """

    def __init__(self, camera, lights, width=400, height=400, max_depth=3, f_distance=5,
                 number_of_samples=8, *, allocate=True):
        self.camera = _f64(camera)
        self.lights = lights
        self.width = int(width)
        self.height = int(height)
        self.max_depth = int(max_depth)
        self.aspect_ratio = width / height
        self.left = -1.0
        self.top = 1 / self.aspect_ratio
        self.right = 1.0
        self.bottom = -1 / self.aspect_ratio
        self.f_distance = float(f_distance)
        self.number_of_samples = int(number_of_samples)
        self.image = np.zeros((height, width, 3), dtype=np.float64)
        if allocate:
            shape = (height, width, number_of_samples, max_depth)
            self.rand_0 = np.random.rand(*shape)
            self.rand_1 = np.random.rand(*shape)
            with np.errstate(divide="ignore"):
                self.rand_0_logit = np.log(self.rand_0 / (1 - self.rand_0))          # R/src/scene.py:70-71
                self.rand_1_logit = np.log(self.rand_1 / (1 - self.rand_1))
            self.bounce_record = np.ones(shape, dtype=np.int8)                       # :72
            self.record_log_pdf = np.zeros(shape, dtype=np.float64)                  # :73
        else:
            self.rand_0 = self.rand_1 = None
            self.rand_0_logit = self.rand_1_logit = None
            self.bounce_record = self.record_log_pdf = None


def generate_area_light_samples(tri_1, tri_2, source_mat, number_of_samples, total_area):
    """R/src/light_samples.py:17-32, including Q14: both lights of a pair sit at the point
    sampled on ``tri_1`` and ``b`` is drawn as ``uniform(1, 0)``.  Returns a Python list."""
    n = number_of_samples
    a = np.random.uniform(0, 1, n)
    b = np.random.uniform(1, 0, n)
    out = []
    for x in range(n):
        sa = math.sqrt(a[x])
        p = tri_1.vertex_1 * (1 - sa) + tri_1.vertex_2 * (sa * (1 - b[x])) + tri_1.vertex_3 * (b[x] * sa)
        out.append(Light(source=p, material=source_mat, normal=tri_1.normal, total_area=total_area))
        out.append(Light(source=p, material=source_mat, normal=tri_2.normal, total_area=total_area))
    return out


# ---- presets used by the example scene (R/src/constants.py:27-37,56-58,82-85) ----
def _c(a, d, s):
    return Color(ambient=np.array(a, dtype=np.float64), diffuse=np.array(d, dtype=np.float64),
                 specular=np.array(s, dtype=np.float64))


WHITE = _c([1, 1, 1], [1, 1, 1], [1, 1, 1])
WHITE_2 = _c([0, 0, 0], [0.55, 0.55, 0.55], [0.7, 0.7, 0.7])
RED = _c([0.1, 0, 0], [0.7, 0, 0], [1, 1, 1])
GREEN = _c([0, 0.1, 0], [0, 0.6, 0], [1, 1, 1])
GLASS = _c([0.0, 0.0, 0.0], [0.588235, 0.670588, 0.729412], [0.9, 0.9, 0.9])
GLASS_MAT = Material(color=GLASS, shininess=96, reflection=0.2, ior=1.5, transmission=1.0,
                     is_diffuse=False, is_mirror=False)
