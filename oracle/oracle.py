"""ctypes wrapper of the CPU oracle (oracle/lt_oracle.c).  TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs — never from the light_transport_b200 package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

PD = C.POINTER(C.c_double)
PI32 = C.POINTER(C.c_int32)
PU8 = C.POINTER(C.c_uint8)


class LtoScene(C.Structure):
    _fields_ = [("n_tris", C.c_int32), ("n_mats", C.c_int32), ("n_lights", C.c_int32), ("n_nodes", C.c_int32),
                ("tri_v0", PD), ("tri_v1", PD), ("tri_v2", PD), ("tri_n", PD), ("tri_mat", PI32), ("tri_light", PU8),
                ("mat_kd", PD), ("mat_emission", PD), ("mat_ior", PD), ("mat_transmission", PD), ("mat_flags", PI32),
                ("light_pos", PD), ("light_n", PD), ("light_lekd", PD), ("light_area", PD),
                ("node_bmin", PD), ("node_bmax", PD), ("node_prim_off", PI32), ("node_second", PI32),
                ("node_n_prims", PI32), ("node_axis", PI32)]


class LtoParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("s0", C.c_int32), ("s1", C.c_int32),
                ("max_depth", C.c_int32), ("semantics", C.c_int32), ("traversal", C.c_int32), ("rr_start", C.c_int32),
                ("n_light_idx", C.c_int32), ("n_threads", C.c_int32), ("pad_", C.c_int32),
                ("rr_floor", C.c_double), ("eps", C.c_double), ("camera", C.c_double * 3), ("f_distance", C.c_double),
                ("seed", C.c_uint64), ("pixel0", C.c_int32), ("pixel1", C.c_int32), ("grad_ungated", C.c_int32), ("pad2_", C.c_int32)]


class LtoCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("paths", "rays_ext", "rays_shadow", "box_tests", "tri_tests", "max_frames",
                                          "light_idx_overflow", "errors")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def build(force=False):
    so = os.path.join(HERE, "liblt_oracle.so")
    src = os.path.join(HERE, "lt_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src),
                                                                    os.path.getmtime(os.path.join(HERE, "lt_oracle.h"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "-B", "liblt_oracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.lto_intersect.restype = C.c_int
        _LIB.lto_render.restype = C.c_int
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


class OracleScene:
    """Holds the numpy arrays alive next to the C struct."""

    def __init__(self, packed, nodes=None):
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.k = dict(tri_v0=f(packed.tri_v0), tri_v1=f(packed.tri_v1), tri_v2=f(packed.tri_v2), tri_n=f(packed.tri_n),
                      tri_mat=i(packed.tri_mat), tri_light=np.ascontiguousarray(packed.tri_light, dtype=np.uint8),
                      mat_kd=f(packed.mat_kd), mat_emission=f(packed.mat_emission), mat_ior=f(packed.mat_ior),
                      mat_transmission=f(packed.mat_transmission), mat_flags=i(packed.mat_flags),
                      light_pos=f(packed.light_pos), light_n=f(packed.light_n), light_lekd=f(packed.light_lekd),
                      light_area=f(packed.light_area))
        s = LtoScene()
        s.n_tris, s.n_mats, s.n_lights = len(self.k["tri_mat"]), len(self.k["mat_flags"]), len(self.k["light_area"])
        for name, typ in LtoScene._fields_:
            if name in self.k:
                setattr(s, name, _p(self.k[name], typ))
        if nodes is not None:
            self.k.update(node_bmin=f(nodes.bmin), node_bmax=f(nodes.bmax), node_prim_off=i(nodes.prim_off),
                          node_second=i(nodes.second), node_n_prims=i(nodes.n_prims), node_axis=i(nodes.axis))
            s.n_nodes = len(nodes)
            for name in ("node_bmin", "node_bmax", "node_prim_off", "node_second", "node_n_prims", "node_axis"):
                setattr(s, name, _p(self.k[name], dict(LtoScene._fields_)[name]))
        self.c = s


def intersect(scene: OracleScene, o, d, traversal="compat"):
    o = np.ascontiguousarray(o, dtype=np.float64)
    d = np.ascontiguousarray(d, dtype=np.float64)
    n = len(o)
    prim = np.empty(n, np.int32)
    t = np.empty(n, np.float64)
    cnt = LtoCounters()
    rc = lib().lto_intersect(C.byref(scene.c), 0 if traversal == "compat" else 1, _p(o, PD), _p(d, PD),
                             C.c_int64(n), _p(prim, PI32), _p(t, PD), C.byref(cnt))
    if rc:
        raise IndexError("setitem out of range (Q9b: the reference's intersect_bvh raises here)")
    return prim, t, cnt.as_dict()


def render(scene: OracleScene, *, width, height, spp, max_depth, camera, f_distance, s0=0, s1=None,
           semantics="reference", traversal="compat", rand0=None, rand1=None, light_idx=None, seed=0,
           rr_start=None, return_samples=False, n_threads=0):
    p = LtoParams()
    p.width, p.height, p.spp, p.s0, p.s1, p.max_depth = width, height, spp, s0, spp if s1 is None else s1, max_depth
    p.semantics = 0 if semantics == "reference" else 1
    p.traversal = 0 if traversal == "compat" else 1
    p.rr_start = (3 if p.semantics == 0 else 5) if rr_start is None else rr_start
    p.rr_floor, p.eps, p.f_distance, p.seed, p.n_threads = 0.05, 0.000001, float(f_distance), int(seed), n_threads
    for k in range(3):
        p.camera[k] = float(camera[k])
    if rand0 is not None:
        rand0 = np.ascontiguousarray(rand0, dtype=np.float64)
        rand1 = np.ascontiguousarray(rand1, dtype=np.float64)
        assert rand0.shape == (height, width, spp, max_depth) == rand1.shape
    if light_idx is not None:
        light_idx = np.ascontiguousarray(light_idx, dtype=np.int32)
        assert light_idx.shape[:3] == (height, width, spp)
        p.n_light_idx = light_idx.shape[3]
    film = np.zeros((height, width, 3), np.float64)
    samples = np.zeros((height, width, p.s1 - p.s0, 3), np.float64) if return_samples else None
    cnt = LtoCounters()
    rc = lib().lto_render(C.byref(scene.c), C.byref(p), _p(rand0, PD), _p(rand1, PD), _p(light_idx, PI32),
                          _p(film, PD), _p(samples, PD), C.byref(cnt))
    if rc:
        raise IndexError("setitem out of range (Q9b)")
    return film, samples, cnt.as_dict()


def render_pixels(scene: OracleScene, pixels, *, width, height, spp, max_depth, camera, f_distance, s0=0, s1=None, semantics="reference",
                  traversal="compat", seed=0, rr_start=None, n_threads=0):
    """(sum[n,3], sumsq[n,3], counters) of the per-sample radiance of the listed pixels (row-major indices) of a W x H film."""
    p = LtoParams()
    p.width, p.height, p.spp, p.s0, p.s1, p.max_depth = width, height, spp, s0, spp if s1 is None else s1, max_depth
    p.semantics = 0 if semantics == "reference" else 1
    p.traversal = 0 if traversal == "compat" else 1
    p.rr_start = (3 if p.semantics == 0 else 5) if rr_start is None else rr_start
    p.rr_floor, p.eps, p.f_distance, p.seed, p.n_threads = 0.05, 0.000001, float(f_distance), int(seed), n_threads
    for k in range(3):
        p.camera[k] = float(camera[k])
    pixels = np.ascontiguousarray(pixels, dtype=np.int32)
    a, b = np.zeros((len(pixels), 3)), np.zeros((len(pixels), 3))
    cnt = LtoCounters()
    L = lib()
    L.lto_render_pixels.restype = C.c_int
    rc = L.lto_render_pixels(C.byref(scene.c), C.byref(p), _p(pixels, PI32), C.c_int32(len(pixels)), _p(a, PD), _p(b, PD), C.byref(cnt))
    if rc:
        raise IndexError("setitem out of range (Q9b)")
    return a, b, cnt.as_dict()


def render_records(scene: OracleScene, *, width, height, spp, max_depth, camera, f_distance, rand0, rand1, light_idx=None,
                   traversal="compat", gradients=False, n_threads=0, pixel_range=None, ungated=False, logits=None):
    """render_scene_samples' per-sample outputs (lto_render_records): dict with samples, record_log_pdf,
    scene_record_log_pdf, bounce_record, direct_list, indirect_list and (gradients=True) rlp_minus/rlp_plus/record_s."""
    p = LtoParams()
    p.width, p.height, p.spp, p.s0, p.s1, p.max_depth = width, height, spp, 0, spp, max_depth
    p.semantics, p.traversal, p.rr_start = 0, 0 if traversal == "compat" else 1, 3
    p.rr_floor, p.eps, p.f_distance, p.n_threads = 0.05, 0.000001, float(f_distance), n_threads
    for k in range(3):
        p.camera[k] = float(camera[k])
    rand0 = np.ascontiguousarray(rand0, dtype=np.float64)
    rand1 = np.ascontiguousarray(rand1, dtype=np.float64)
    lead = (height, width)
    if pixel_range is not None:           # arrays cover the pixels [pixel0, pixel1) only: shape [n_pixels, spp, ...]
        p.pixel0, p.pixel1 = int(pixel_range[0]), int(pixel_range[1])
        lead = (p.pixel1 - p.pixel0,)
    p.grad_ungated = 1 if ungated else 0
    assert rand0.shape == lead + (spp, max_depth)
    if logits is not None:                # the "more samples" block hands the uniforms themselves to calculate_gradients (:347)
        l0, l1 = (np.ascontiguousarray(a, dtype=np.float64) for a in logits)
    else:
        l0 = np.log(rand0 / (1 - rand0))          # R/src/scene.py:70-71
        l1 = np.log(rand1 / (1 - rand1))
    if light_idx is not None:
        light_idx = np.ascontiguousarray(light_idx, dtype=np.int32)
        p.n_light_idx = light_idx.shape[-1]
    D = max_depth
    out = dict(samples=np.zeros(lead + (spp, 3)), record_log_pdf=np.zeros(lead + (spp, D)),
               scene_record_log_pdf=np.zeros(lead + (spp, D)), bounce_record=np.ones(lead + (spp, D), np.int8),
               direct_list=np.zeros(lead + (spp, 3, D)), indirect_list=np.zeros(lead + (spp, 3, D)))
    if gradients:
        out["rlp_minus"] = np.zeros(lead + (spp, 2 * D, D))
        out["rlp_plus"] = np.zeros(lead + (spp, 2 * D, D))
        out["record_s"] = np.full(lead + (spp, 2 * D), 0.01)
    cnt = LtoCounters()
    L = lib()
    L.lto_render_records.restype = C.c_int
    rc = L.lto_render_records(C.byref(scene.c), C.byref(p), _p(rand0, PD), _p(rand1, PD), _p(l0, PD), _p(l1, PD), _p(light_idx, PI32),
                              _p(out["samples"], PD), _p(out["record_log_pdf"], PD), _p(out["scene_record_log_pdf"], PD),
                              out["bounce_record"].ctypes.data_as(C.POINTER(C.c_int8)), _p(out["direct_list"], PD),
                              _p(out["indirect_list"], PD), _p(out.get("rlp_minus"), PD), _p(out.get("rlp_plus"), PD), C.byref(cnt))
    if rc:
        raise RuntimeError(f"lto_render_records failed ({rc})")
    out["counters"] = cnt.as_dict()
    return out


class LtoMediumParams(C.Structure):
    _fields_ = [("albedo", C.c_double), ("g", C.c_double), ("radius", C.c_double), ("max_depth", C.c_int32), ("mode", C.c_int32),
                ("rr_start", C.c_int32), ("n_rbins", C.c_int32), ("rr_floor", C.c_double), ("seed", C.c_uint64), ("first_photon", C.c_uint64)]


class LtoMediumTotals(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("escaped", "absorbed", "truncated", "rr_killed", "rr_gained", "sum_mu")] + \
               [(k, C.c_uint64) for k in ("n_collisions", "n_segments", "n_scatters")]


def medium_walk(n_photons, *, albedo=0.99, g=0.0, radius=0.0, max_depth=256, mode="implicit", rr_start=3, rr_floor=0.05, n_rbins=32,
                seed=0, first_photon=0, n_threads=0):
    """CPU restatement of lt_medium_walk (float64, same Philox streams)."""
    p = LtoMediumParams(float(albedo), float(g), float(radius), int(max_depth), 0 if mode == "analog" else 1, int(rr_start), int(n_rbins),
                        float(rr_floor), int(seed), int(first_photon))
    hist = np.zeros(max_depth + 1, np.uint64)
    rbins = np.zeros(max(n_rbins, 1), np.float64)
    tot = LtoMediumTotals()
    L = lib()
    L.lto_medium_walk.restype = C.c_int
    L.lto_medium_walk(C.byref(p), C.c_uint64(int(n_photons)), hist.ctypes.data_as(C.c_void_p), rbins.ctypes.data_as(C.c_void_p),
                      C.byref(tot), C.c_int(n_threads))
    out = {k: getattr(tot, k) for k, _ in LtoMediumTotals._fields_}
    out.update(hist=hist, rbins=rbins[:n_rbins], n_photons=int(n_photons))
    return out


def henyey_greenstein(cos_theta, g):
    L = lib()
    L.lto_henyey_greenstein.restype = C.c_double
    return L.lto_henyey_greenstein(C.c_double(cos_theta), C.c_double(g))
