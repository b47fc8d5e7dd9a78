"""Python host side of the drop-in: ``render_scene(scene, primitives, bvh)`` on the GPU.

Mirrors the reference's call (``R/src/path_tracing.py:371-387`` /
``R/src/path_tracing_old.py:142-171`` / ``R/src/path_tracing_fix1.py:139-169``;
``R/`` = ``LightTransportSimulator/light_transport/``): same positional arguments, same
side effect on ``scene.image``, plus keyword-only extensions.  Everything numeric happens in
``liblt_b200.so`` (C-ABI: ``include/lt_b200.h``) through ``ctypes``; there is NO CPU path —
a missing library or GPU raises.
"""
from __future__ import annotations

import atexit
import collections
import ctypes as C
import hashlib
import os
import weakref

import numpy as np

from . import _pack, compat_bvh

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblt_b200.so")     # (A/B tools under tools/ point this at a variant build before first use)

TRAVERSAL = {"compat": 0, "nearest": 1, "lbvh": 2, "lbvh2": 3}      # lbvh: 8-wide compressed hierarchy; lbvh2: the binary tree under it
SEMANTICS = {"reference": 0, "iterative": 1}
REAL = {"f32": 0, "f64": 1}
SCHEDULER = {"megakernel": 0, "wavefront": 1}
# measured: megakernel 16.2 (f32) / 4.9 (f64) Grays/s, wavefront scheduler 11.4 / 3.9
AUTO_SCHEDULER = {"f32": "megakernel", "f64": "megakernel"}

PD = C.POINTER(C.c_double)
PI32 = C.POINTER(C.c_int32)
PU8 = C.POINTER(C.c_uint8)


class LtSceneDesc(C.Structure):
    _fields_ = [("n_tris", C.c_int32), ("n_mats", C.c_int32), ("n_lights", C.c_int32), ("reserved", C.c_int32),
                ("tri_v0", PD), ("tri_v1", PD), ("tri_v2", PD), ("tri_n", PD), ("tri_mat", PI32), ("tri_light", PU8),
                ("mat_kd", PD), ("mat_emission", PD), ("mat_ior", PD), ("mat_transmission", PD), ("mat_flags", PI32),
                ("light_pos", PD), ("light_n", PD), ("light_lekd", PD), ("light_area", PD)]


class LtNodeDesc(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("reserved", C.c_int32), ("bmin", PD), ("bmax", PD), ("prim_off", PI32),
                ("second", PI32), ("n_prims", PI32), ("axis", PI32)]


class LtRenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("s0", C.c_int32), ("s1", C.c_int32),
                ("max_depth", C.c_int32), ("semantics", C.c_int32), ("traversal", C.c_int32), ("real", C.c_int32),
                ("scheduler", C.c_int32), ("rr_start", C.c_int32), ("n_light_idx", C.c_int32),
                ("rr_floor", C.c_double), ("eps", C.c_double), ("camera", C.c_double * 3), ("f_distance", C.c_double),
                ("seed", C.c_uint64), ("pixel0", C.c_int32), ("pixel1", C.c_int32), ("paths_per_warp", C.c_int32),
                ("flags", C.c_int32)]


class LtCounters(C.Structure):
    _fields_ = [(k, C.c_uint64) for k in ("paths", "rays_ext", "rays_shadow", "box_tests", "tri_tests", "max_frames",
                                          "light_idx_overflow", "frame_overflow", "lane_steps", "lane_slots",
                                          "kernel_ns", "launches", "rays_memo")]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


class LtBvhStats(C.Structure):
    _fields_ = [("n_prims", C.c_int32), ("n_nodes", C.c_int32), ("max_depth", C.c_int32), ("binary_depth", C.c_int32),
                ("build_ms", C.c_float), ("sort_ms", C.c_float), ("root_bmin", C.c_double * 3),
                ("root_bmax", C.c_double * 3), ("sah_cost", C.c_double), ("collapse_ms", C.c_float), ("reserved", C.c_float),
                ("bytes", C.c_int64), ("binary_bytes", C.c_int64)]


class LtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"lt_b200 error {code}: {msg}")
        self.code = code


_lib = None


def lib():
    """Load liblt_b200.so (built by ``__graft_entry__.build()`` / ``make -C light_transport_b200/csrc``)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `make -C light_transport_b200/csrc` "
                              "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.lt_last_error.restype = C.c_char_p
        L.lt_version.restype = C.c_char_p
        for name in ("lt_init", "lt_set_stream", "lt_device_info", "lt_scene_create", "lt_scene_set_compat_bvh", "lt_bvh_build",
                     "lt_intersect", "lt_render", "lt_render_image", "lt_render_records", "lt_film_finalize", "lt_measure_fp32_peak",
                     "lt_measure_fp64_peak", "lt_measure_l2_bandwidth", "lt_host_register", "lt_host_unregister"):
            getattr(L, name).restype = C.c_int
        L.lt_destroy.restype = None
        L.lt_scene_destroy.restype = None
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise LtError(rc, lib().lt_last_error().decode())


def _ptr(a, typ):
    """numpy array, torch CUDA tensor (anything with data_ptr), raw int address, or None → pointer."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(typ)
    if hasattr(a, "data_ptr"):
        return C.cast(C.c_void_p(a.data_ptr()), typ)
    return C.cast(C.c_void_p(int(a)), typ)


_ctx_cache: dict[int, C.c_void_p] = {}
_live_renderers: "weakref.WeakSet[Renderer]" = None     # created below (needs the class)


def _get_ctx(device: int) -> C.c_void_p:
    """One lt_ctx per device and process: stream, film/arena scratch and the CUDA context set-up are paid
    once, not per render_scene call (creating them per call cost 0.1-1.3 s of cudaMalloc/cudaFree jitter)."""
    ctx = _ctx_cache.get(device)
    if ctx is None:
        ctx = C.c_void_p()
        _check(lib().lt_init(C.c_int(device), C.byref(ctx)))
        _ctx_cache[device] = ctx
    return ctx


@atexit.register
def _destroy_contexts():
    """Scenes first, then their contexts: a Renderer that is still alive at exit (module-level objects, scripts that never
    call close()) must not outlive the lt_ctx its scene points into."""
    try:
        from . import parallel
        parallel.close_comms()
    except Exception:
        pass
    if _live_renderers is not None:
        for r in list(_live_renderers):
            try:
                r.close()
            except Exception:
                pass
    for addr in list(_pinned):
        try:
            _pinned[addr][2].detach()
            _unpin(addr)
        except Exception:
            pass
    for ctx in _ctx_cache.values():
        try:
            lib().lt_destroy(ctx)
        except Exception:
            pass
    _ctx_cache.clear()


class Renderer:
    """One uploaded scene on a device (the 'H2D once' half of the boundary); shares the per-device context."""

    def __init__(self, packed: _pack.PackedScene, nodes: compat_bvh.CompatNodes | None = None, device: int = 0,
                 build_lbvh: bool = False):
        global _live_renderers
        L = lib()
        self._ctx = _get_ctx(device)
        self._device = device
        self._stream = None           # cudaStream_t this renderer's work runs on (None: the context's own stream)
        self._scene = C.c_void_p()
        if _live_renderers is None:
            _live_renderers = weakref.WeakSet()
        _live_renderers.add(self)
        f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self._keep = dict(tri_v0=f(packed.tri_v0), tri_v1=f(packed.tri_v1), tri_v2=f(packed.tri_v2), tri_n=f(packed.tri_n),
                          tri_mat=i(packed.tri_mat), tri_light=np.ascontiguousarray(packed.tri_light, dtype=np.uint8),
                          mat_kd=f(packed.mat_kd), mat_emission=f(packed.mat_emission), mat_ior=f(packed.mat_ior),
                          mat_transmission=f(packed.mat_transmission), mat_flags=i(packed.mat_flags),
                          light_pos=f(packed.light_pos), light_n=f(packed.light_n), light_lekd=f(packed.light_lekd),
                          light_area=f(packed.light_area))
        d = LtSceneDesc()
        d.n_tris, d.n_mats, d.n_lights = packed.n_tris, len(packed.mat_flags), packed.n_lights
        types = dict(LtSceneDesc._fields_)
        for k, v in self._keep.items():
            setattr(d, k, _ptr(v, types[k]))
        _check(L.lt_scene_create(self._ctx, C.byref(d), C.byref(self._scene)))
        self.packed = packed
        self.n_lights = packed.n_lights
        self.has_compat = False
        self.lbvh_stats = None
        if nodes is not None:
            self.set_compat_bvh(nodes)
        if build_lbvh:
            self.build_lbvh(keep_binary=(build_lbvh == "binary"))

    # -- structures --------------------------------------------------------------------------------
    def set_compat_bvh(self, nodes: compat_bvh.CompatNodes):
        self._bind()
        d = LtNodeDesc()
        keep = dict(bmin=np.ascontiguousarray(nodes.bmin, np.float64), bmax=np.ascontiguousarray(nodes.bmax, np.float64),
                    prim_off=np.ascontiguousarray(nodes.prim_off, np.int32), second=np.ascontiguousarray(nodes.second, np.int32),
                    n_prims=np.ascontiguousarray(nodes.n_prims, np.int32), axis=np.ascontiguousarray(nodes.axis, np.int32))
        d.n_nodes = len(nodes)
        types = dict(LtNodeDesc._fields_)
        for k, v in keep.items():
            setattr(d, k, _ptr(v, types[k]))
        _check(lib().lt_scene_set_compat_bvh(self._scene, C.byref(d)))
        self.has_compat = True

    def build_lbvh(self, keep_binary=False, leaf_max=0, binary_leaf=0):
        """GPU LBVH build (Morton, radix sort, Karras, refit) + collapse into the 8-wide compressed hierarchy."""
        self._bind()
        st = LtBvhStats()
        _check(lib().lt_bvh_build(self._scene, C.c_int((1 if keep_binary else 0) | ((int(leaf_max) & 3) << 8) | ((int(binary_leaf) & 3) << 10)), C.byref(st)))
        self.lbvh_stats = dict(n_prims=st.n_prims, n_nodes=st.n_nodes, max_depth=st.max_depth, binary_depth=st.binary_depth,
                               build_ms=st.build_ms, sort_ms=st.sort_ms, collapse_ms=st.collapse_ms, sah_cost=st.sah_cost,
                               bytes=st.bytes, binary_bytes=st.binary_bytes, root_bmin=list(st.root_bmin), root_bmax=list(st.root_bmax))
        return self.lbvh_stats

    # -- queries -----------------------------------------------------------------------------------
    def intersect(self, o, d, traversal="compat", real="f64"):
        """Closest hit of explicit rays: ``(prim[n] int32 (-1 = miss), t[n] float64 (inf = miss), counters)``."""
        o = np.ascontiguousarray(o, dtype=np.float64)
        d = np.ascontiguousarray(d, dtype=np.float64)
        if o.shape != d.shape or o.ndim != 2 or o.shape[1] != 3:
            raise ValueError("o and d must both be [n,3]")
        n = len(o)
        prim = np.empty(n, np.int32)
        t = np.empty(n, np.float64)
        cnt = LtCounters()
        self._bind()
        _check(lib().lt_intersect(self._scene, C.c_int(TRAVERSAL[traversal]), C.c_int(REAL[real]), _ptr(o, PD), _ptr(d, PD),
                                  C.c_int64(n), _ptr(prim, PI32), _ptr(t, PD), C.byref(cnt)))
        return prim, t, cnt.as_dict()

    def render(self, *, width, height, spp, max_depth, camera, f_distance, s0=0, s1=None, semantics="reference",
               traversal="compat", real="f32", rand0=None, rand1=None, light_idx=None, seed=0, rr_start=None,
               film=None, return_samples=False, pixel_range=None, paths_per_warp=0, image=None, scheduler="megakernel", lbvh_chunk=0,
               phased=None, ctas_per_sm=0, l2_window=False):
        """One pass of the walk over samples ``[s0,s1)`` of every pixel.  Returns
        ``(film_sum, samples|None, counters)``; ``film`` may be a preallocated float64 buffer (numpy or a
        CUDA tensor / device address of H*W*3 doubles) that is accumulated into."""
        p = LtRenderParams()
        s1 = spp if s1 is None else s1
        p.width, p.height, p.spp, p.s0, p.s1, p.max_depth = width, height, spp, s0, s1, max_depth
        if scheduler == "auto":
            scheduler = AUTO_SCHEDULER[real]
        p.semantics, p.traversal, p.real, p.scheduler = SEMANTICS[semantics], TRAVERSAL[traversal], REAL[real], SCHEDULER[scheduler]
        p.rr_start = (3 if semantics == "reference" else 5) if rr_start is None else rr_start   # path_tracing.py:148 / fix1:126
        p.rr_floor, p.eps, p.f_distance, p.seed = 0.05, _pack_eps(), float(f_distance), int(seed)
        p.paths_per_warp = int(paths_per_warp)
        p.flags = (int(lbvh_chunk) & 0xff) << 8          # LBVH traversal steps per megakernel iteration (0 = default)
        p.flags |= (int(ctas_per_sm) & 0xf) << 20        # A/B: cap on resident CTAs per SM (0 = what the kernel was built for)
        if l2_window:
            p.flags |= 1 << 19                           # A/B: persisting-L2 window on the hierarchy nodes (measured slower on C3)
        if phased is not None:                           # phase-aligned iterations: None = chosen from the scene, True / False = forced (A/B)
            p.flags |= (1 << 17) if phased else (1 << 16)
        cam = np.asarray(camera, dtype=np.float64)
        for k in range(3):
            p.camera[k] = float(cam[k])
        if pixel_range is not None:
            p.pixel0, p.pixel1 = int(pixel_range[0]), int(pixel_range[1])
        if (rand0 is None) != (rand1 is None):
            raise ValueError("rand0 and rand1 go together")
        if isinstance(rand0, np.ndarray):
            rand0 = np.ascontiguousarray(rand0, dtype=np.float64)
            rand1 = np.ascontiguousarray(rand1, dtype=np.float64)
            if rand0.shape != (height, width, spp, max_depth) or rand1.shape != rand0.shape:
                raise ValueError(f"rand_0/rand_1 must be [H,W,spp,max_depth]={(height, width, spp, max_depth)}, got {rand0.shape}")
        if isinstance(light_idx, np.ndarray):
            light_idx = np.ascontiguousarray(light_idx, dtype=np.int32)
            if light_idx.shape[:3] != (height, width, spp):
                raise ValueError("light_idx must be [H,W,spp,K]")
            if light_idx.size and (light_idx.min() < 0 or light_idx.max() >= self.n_lights):
                raise ValueError("light_idx out of range")
            p.n_light_idx = light_idx.shape[3]
        elif light_idx is not None:
            raise ValueError("light_idx must be a numpy array")
        samples = np.zeros((height, width, s1 - s0, 3), np.float64) if return_samples else None
        cnt = LtCounters()
        self._bind()
        if image is not None:
            # film stays on the device; one D2H of the finished image (lt_render_image)
            _check(lib().lt_render_image(self._scene, C.byref(p), _ptr(rand0, PD), _ptr(rand1, PD), _ptr(light_idx, PI32),
                                         _ptr(image, PD), _ptr(samples, PD), C.byref(cnt)))
            return image, samples, cnt.as_dict()
        film_buf = np.zeros((height, width, 3), np.float64) if film is None else film
        _check(lib().lt_render(self._scene, C.byref(p), _ptr(rand0, PD), _ptr(rand1, PD), _ptr(light_idx, PI32),
                               _ptr(film_buf, PD), _ptr(samples, PD), C.byref(cnt)))
        return film_buf, samples, cnt.as_dict()

    def render_records(self, *, width, height, spp, max_depth, camera, f_distance, rand0, rand1, light_idx=None,
                       traversal="compat", real="f64", gradients=False, lists=True, seed=0, pixel_range=None, ungated=False,
                       logits=None):
        """The per-sample outputs of the reference's ``render_scene_samples`` (R/src/path_tracing.py:251-303) for
        replayed random numbers: dict with ``samples``, ``record_log_pdf`` (the array ``trace_path`` returns),
        ``scene_record_log_pdf`` / ``bounce_record`` (what it writes into ``Scene``), ``direct_list`` /
        ``indirect_list`` (root frame) and, with ``gradients=True``, ``rlp_minus`` / ``rlp_plus`` / ``record_s`` of
        ``calculate_gradients`` (:203-249).  REFERENCE semantics, ``max_depth <= 32``.
        ``pixel_range=(p0, p1)`` walks only those pixels (row-major) of the ``width x height`` film; every array, in and
        out, then has the leading shape ``[p1 - p0, spp, ...]`` instead of ``[H, W, spp, ...]``.  ``ungated=True`` runs the
        gradient replays with ``calculate_grad=False`` and ``logits=(l0, l1)`` overrides the logits derived from the
        uniforms — together they are the "more samples" block of the reference (:310-364)."""
        p = LtRenderParams()
        p.width, p.height, p.spp, p.s0, p.s1, p.max_depth = width, height, spp, 0, spp, max_depth
        p.semantics, p.traversal, p.real, p.scheduler = 0, TRAVERSAL[traversal], REAL[real], 0
        p.rr_start, p.rr_floor, p.eps, p.f_distance, p.seed = 3, 0.05, _pack_eps(), float(f_distance), int(seed)
        cam = np.asarray(camera, dtype=np.float64)
        for k in range(3):
            p.camera[k] = float(cam[k])
        rand0 = np.ascontiguousarray(rand0, dtype=np.float64)
        rand1 = np.ascontiguousarray(rand1, dtype=np.float64)
        lead = (height, width)
        if pixel_range is not None:
            p.pixel0, p.pixel1 = int(pixel_range[0]), int(pixel_range[1])
            if not (0 <= p.pixel0 < p.pixel1 <= width * height):
                raise ValueError("pixel_range must satisfy 0 <= p0 < p1 <= W*H")
            lead = (p.pixel1 - p.pixel0,)
        if ungated:
            p.flags |= 1 << 18
        if rand0.shape != lead + (spp, max_depth) or rand1.shape != rand0.shape:
            raise ValueError(f"rand_0/rand_1 must be {lead + (spp, max_depth)}")
        if light_idx is not None:
            light_idx = np.ascontiguousarray(light_idx, dtype=np.int32)
            if light_idx.shape[:-1] != lead + (spp,):
                raise ValueError("light_idx must be [..., spp, K] like rand_0")
            p.n_light_idx = light_idx.shape[-1]
        D = max_depth
        out = dict(samples=np.zeros(lead + (spp, 3)), record_log_pdf=np.zeros(lead + (spp, D)),
                   scene_record_log_pdf=np.zeros(lead + (spp, D)), bounce_record=np.ones(lead + (spp, D), np.int8))
        if lists:
            out["direct_list"] = np.zeros(lead + (spp, 3, D))
            out["indirect_list"] = np.zeros(lead + (spp, 3, D))
        l0 = l1 = None
        if gradients:
            if logits is not None:
                l0, l1 = (np.ascontiguousarray(a, dtype=np.float64) for a in logits)
                if l0.shape != rand0.shape or l1.shape != rand0.shape:
                    raise ValueError("logits must have the shape of rand_0")
            else:
                with np.errstate(divide="ignore"):
                    l0 = np.log(rand0 / (1 - rand0))              # Scene.rand_0_logit, R/src/scene.py:70-71
                    l1 = np.log(rand1 / (1 - rand1))
            out["rlp_minus"] = np.zeros(lead + (spp, 2 * D, D))
            out["rlp_plus"] = np.zeros(lead + (spp, 2 * D, D))
            out["record_s"] = np.full(lead + (spp, 2 * D), 0.01)      # _s of :210,230
        cnt = LtCounters()
        self._bind()
        _check(lib().lt_render_records(self._scene, C.byref(p), _ptr(rand0, PD), _ptr(rand1, PD), _ptr(l0, PD), _ptr(l1, PD),
                                       _ptr(light_idx, PI32), _ptr(out["samples"], PD), _ptr(out["record_log_pdf"], PD),
                                       _ptr(out["scene_record_log_pdf"], PD), _ptr(out["bounce_record"], C.POINTER(C.c_int8)),
                                       _ptr(out.get("direct_list"), PD), _ptr(out.get("indirect_list"), PD),
                                       _ptr(out.get("rlp_minus"), PD), _ptr(out.get("rlp_plus"), PD), C.byref(cnt)))
        out["counters"] = cnt.as_dict()
        return out

    def finalize(self, film_sum, spp_total, out=None):
        """``image = clip(film_sum / spp, 0, 1)`` (R/src/path_tracing.py:304-305) on the device."""
        if isinstance(film_sum, np.ndarray):
            n = film_sum.size
            out = np.empty_like(film_sum) if out is None else out
        else:
            n = int(film_sum.numel())
            if out is None:
                raise ValueError("pass out= for device buffers")
        self._bind()
        _check(lib().lt_film_finalize(self._ctx, _ptr(film_sum, PD), C.c_int64(n), C.c_int64(spp_total), _ptr(out, PD)))
        return out

    def set_stream(self, cuda_stream):
        """Run this renderer's kernels on ``cuda_stream`` (int handle, e.g.
        ``torch.cuda.current_stream().cuda_stream``); ``None`` = the context's own stream.  The stream belongs to the
        Renderer: it is bound to the (per-device, shared) context at the start of each of its calls, so renderers on one
        device do not disturb one another."""
        self._stream = int(cuda_stream) if cuda_stream else None
        self._bind()

    def _bind(self):
        if self._ctx_gone():
            raise LtError(-1, "the library context of this Renderer was destroyed (interpreter exit)")
        _check(lib().lt_set_stream(self._ctx, C.c_void_p(self._stream) if self._stream else None))

    def _ctx_gone(self):
        return _ctx_cache.get(self._device) is not self._ctx

    def device_info(self):
        sm, khz, cc0, cc1 = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
        mem = C.c_int64()
        _check(lib().lt_device_info(self._ctx, C.byref(sm), C.byref(khz), C.byref(mem), C.byref(cc0), C.byref(cc1)))
        return dict(sm_count=sm.value, clock_khz=khz.value, global_mem=mem.value, cc=(cc0.value, cc1.value))

    def measure_fp32_peak(self):
        self._bind()
        v = C.c_double()
        _check(lib().lt_measure_fp32_peak(self._ctx, C.byref(v)))
        return v.value

    def measure_fp64_peak(self):
        v = C.c_double()
        _check(lib().lt_measure_fp64_peak(self._ctx, C.byref(v)))
        return v.value

    def measure_l2_bandwidth(self, buffer_bytes=48 << 20, granule_bytes=80, mode=1):
        """GB/s of 128-bit L2 gathers (mode 1: one random granule per lane, the pattern of an incoherent traversal)."""
        v = C.c_double()
        _check(lib().lt_measure_l2_bandwidth(self._ctx, C.c_int64(buffer_bytes), C.c_int(granule_bytes), C.c_int(mode), C.byref(v)))
        return v.value

    def close(self):
        """Free the device scene.  Safe to call twice, and a no-op once the context is gone (interpreter exit)."""
        if getattr(self, "_scene", None) and self._scene:
            if not self._ctx_gone():
                lib().lt_scene_destroy(self._scene)
            self._scene = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _pack_eps():
    from .api import EPSILON
    return EPSILON


# ---- uploaded scenes kept across render_scene() calls ---------------------------------------------------------------
# The reference's entry point takes the scene objects on every call; uploading them (and building the hierarchy) every
# time costs more than a small render (C1: 0.17 ms of a 0.9 ms call).  render_scene()/intersect() therefore keep the last
# few Renderers, keyed by the CONTENT of what was uploaded (a hash of the packed arrays and the node list: reference object
# lists are re-packed on every call) for scenes up to RENDERER_CACHE_HASH_BYTES, and by the identity of the PackedScene
# object above that (hashing 100 MB costs more than the upload).  The very same PackedScene/CompatNodes objects are
# recognised by identity without hashing (0.3 ms on the LTS scene); they are value objects — after editing one in place call
# clear_renderer_cache().
RENDERER_CACHE_SIZE = 4
RENDERER_CACHE_HASH_BYTES = 1 << 20
_renderer_cache: "collections.OrderedDict[tuple, tuple]" = collections.OrderedDict()
_renderer_ids: dict[tuple, tuple] = {}          # (device, mode, id(packed), id(nodes)) -> (cache key, weakrefs): no hashing for the same objects
_PACKED_FIELDS = ("tri_v0 tri_v1 tri_v2 tri_n tri_mat tri_light mat_kd mat_emission mat_ior mat_transmission mat_flags "
                  "light_pos light_n light_lekd light_area").split()
_NODE_FIELDS = "bmin bmax prim_off second n_prims axis".split()


def _content_key(packed, nodes):
    arrays = [np.ascontiguousarray(getattr(packed, k)) for k in _PACKED_FIELDS]
    if nodes is not None:
        arrays += [np.ascontiguousarray(getattr(nodes, k)) for k in _NODE_FIELDS]
    if sum(a.nbytes for a in arrays) > RENDERER_CACHE_HASH_BYTES:
        return None
    h = hashlib.sha256()            # hardware-assisted on the hosts this runs on: 3x the throughput of blake2b here
    for a in arrays:
        h.update(str((a.dtype.str, a.shape)).encode())
        h.update(a.view(np.uint8).reshape(-1).data if a.size else b"")
    return h.digest()


def cached_renderer(packed, nodes, device=0, lbvh_mode=False):
    """The Renderer for this scene content on ``device`` (see above); created and uploaded on first use.

    The SAME PackedScene / CompatNodes objects coming in again are recognised without hashing (they are value objects:
    edit a copy, or call clear_renderer_cache()); different objects are compared by content."""
    ikey = (device, lbvh_mode, id(packed), id(nodes))
    hit = _renderer_ids.get(ikey)
    if hit is not None:
        key, pref, nref = hit
        ent = _renderer_cache.get(key)
        if ent is not None and pref() is packed and (nodes is None or nref() is nodes) and ent[0]._scene and not ent[0]._ctx_gone():
            _renderer_cache.move_to_end(key)
            return ent[0]
        _renderer_ids.pop(ikey)
    ck = _content_key(packed, nodes)
    key = (device, lbvh_mode, ck if ck is not None else ("id", id(packed), id(nodes)))
    ent = _renderer_cache.get(key)
    r = None
    if ent is not None:
        r, pref, nref = ent
        alive = r._scene and not r._ctx_gone()
        same = ck is not None or (pref() is packed and (nodes is None or nref() is nodes))
        if alive and same:
            _renderer_cache.move_to_end(key)
        else:
            _renderer_cache.pop(key)
            r.close()
            r = None
    if r is None:
        for k in [k for k, (_, pr, _) in _renderer_cache.items() if pr() is None]:       # scenes that no longer exist
            _renderer_cache.pop(k)[0].close()
        r = Renderer(packed, nodes, device=device, build_lbvh=lbvh_mode)
        _renderer_cache[key] = (r, weakref.ref(packed), weakref.ref(nodes) if nodes is not None else None)
        while len(_renderer_cache) > RENDERER_CACHE_SIZE:
            _, (old, _, _) = _renderer_cache.popitem(last=False)
            old.close()
    for k in [k for k, (_, pr, nr) in _renderer_ids.items() if pr() is None or (nr is not None and nr() is None)]:
        _renderer_ids.pop(k)
    _renderer_ids[ikey] = (key, weakref.ref(packed), weakref.ref(nodes) if nodes is not None else None)
    return r


def clear_renderer_cache():
    """Close every cached Renderer (device memory of the uploaded scenes is released)."""
    _renderer_ids.clear()
    while _renderer_cache:
        _, (r, _, _) = _renderer_cache.popitem()
        r.close()


# ---- page-locked result arrays ------------------------------------------------------------------------------------------
# render_scene() hands back scene.image like the reference does; registering that array once (cudaHostRegister) turns every
# later read-back into a single DMA (C1: a 1.5 MB image, 0.1 ms of a 0.7 ms call otherwise spent in staging + host copy).
_pinned: dict[int, tuple] = {}          # address -> (bytes, device, finalizer)
PIN_MIN_BYTES, PIN_MAX_BYTES = 1 << 16, 1 << 30


def _unpin(addr):
    ent = _pinned.pop(addr, None)
    if ent is not None and _ctx_cache.get(ent[1]) is not None:
        lib().lt_host_unregister(_ctx_cache[ent[1]], C.c_void_p(addr))


def pin_array(a: np.ndarray, device: int = 0) -> bool:
    """Page-lock ``a``'s memory for the life of the array object (no-op if already done; False if the driver refuses)."""
    if not (isinstance(a, np.ndarray) and a.flags.c_contiguous and PIN_MIN_BYTES <= a.nbytes <= PIN_MAX_BYTES):
        return False
    addr = a.ctypes.data
    ent = _pinned.get(addr)
    if ent is not None and ent[0] >= a.nbytes:
        return True
    if ent is not None:
        ent[2].detach()
        _unpin(addr)
    if lib().lt_host_register(_get_ctx(device), C.c_void_p(addr), C.c_size_t(a.nbytes)) != 0:
        return False
    _pinned[addr] = (a.nbytes, device, weakref.finalize(a, _unpin, addr))
    return True


def _lbvh_mode(traversal):
    """what Renderer(build_lbvh=...) has to build for a traversal: nothing, the 8-wide hierarchy, or that plus the binary nodes"""
    return "binary" if traversal == "lbvh2" else traversal == "lbvh"


def _nodes_from_bvh(bvh, packed):
    if bvh is None:
        return compat_bvh.build_compat_nodes(packed.tri_v0, packed.tri_v1, packed.tri_v2)
    if isinstance(bvh, compat_bvh.CompatNodes):
        return bvh
    return compat_bvh.from_linear_bvh(bvh)


def intersect(primitives, bvh, o, d, *, lights=(), traversal="compat", real="f64", device=0):
    """Batch form of ``intersect_bvh(ray, primitives, linear_bvh)`` (R/src/bvh_new.py:413-482)."""
    packed = primitives if isinstance(primitives, _pack.PackedScene) else _pack.pack_scene(primitives, lights)
    r = cached_renderer(packed, _nodes_from_bvh(bvh, packed) if traversal == "compat" else None, device, _lbvh_mode(traversal))
    r.set_stream(None)
    return r.intersect(o, d, traversal, real)


def render_scene(scene, primitives, bvh=None, *, spp=None, rng=None, seed=0, semantics="reference", traversal="compat",
                 light_idx=None, real=None, return_samples=False, world=None, device=None, return_info=False,
                 scheduler="auto", all_ranks=False, film_blend=None):
    """Drop-in for ``render_scene(scene, primitives, bvh)``.

    ``scene``/``primitives``/``bvh`` are the reference's objects (or the ``api`` mirrors / a
    ``PackedScene`` / ``CompatNodes``).  Returns the image ``clip(mean radiance, 0, 1)`` as float64
    ``[H,W,3]`` and writes it to ``scene.image`` like the reference does.

    rng="replay" (default when ``scene.rand_0`` exists and ``spp`` is not overridden) consumes
    ``scene.rand_0/rand_1`` and optional ``light_idx[H,W,spp,K]``; rng="philox" draws counter-based
    numbers keyed by ``seed`` (needed for configs whose ``Scene`` arrays cannot exist, Q15).
    ``real`` defaults to "f64" for replay and "f32" for philox.
    ``world=(rank, world_size)`` or an initialised ``torch.distributed`` group shards the samples
    ``[r*spp/R, (r+1)*spp/R)`` across ranks and sums the film with ONE reduce to rank 0 (NCCL on GPUs): rank 0
    finalizes, copies the image to the host and returns it, the other ranks return ``None`` for the image
    (``all_ranks=True``: all-reduce, every rank gets the image).
    ``film_blend=b`` reproduces ``path_tracing_fix1.render_scene`` (R/src/path_tracing_fix1.py:165-166): instead of being
    overwritten, ``scene.image`` ACCUMULATES ``b * clip(mean radiance, 0, 1)`` (the reference uses b = 0.25 and calls it
    four times); the accumulated ``scene.image`` is returned.
    """
    H, W, D = int(scene.height), int(scene.width), int(scene.max_depth)
    n_spp = int(scene.number_of_samples) if spp is None else int(spp)
    have_rand = getattr(scene, "rand_0", None) is not None
    if rng is None:
        rng = "replay" if (have_rand and spp is None) else "philox"
    if real is None:
        real = "f64" if rng == "replay" else "f32"
    rand0 = rand1 = None
    if rng == "replay":
        if not have_rand:
            raise ValueError("rng='replay' needs scene.rand_0/rand_1")
        rand0, rand1 = np.asarray(scene.rand_0), np.asarray(scene.rand_1)
        if rand0.shape[2] != n_spp:
            raise ValueError("spp must equal scene.number_of_samples in replay mode")
    elif rng != "philox":
        raise ValueError("rng must be 'replay' or 'philox'")
    packed = primitives if isinstance(primitives, _pack.PackedScene) else _pack.pack_scene(primitives, scene.lights)
    nodes = _nodes_from_bvh(bvh, packed) if traversal == "compat" else None

    from . import parallel
    rank, size, group, use_dist = 0, 1, None, False
    if world is not None:
        if isinstance(world, tuple):
            rank, size = world                                   # virtual rank: renders its share, no reduce
        else:
            import torch.distributed as dist
            group = None if world is True else world
            rank, size, use_dist = dist.get_rank(group), dist.get_world_size(group), True
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if use_dist else 0
    s0, s1 = parallel.sample_range(rank, size, n_spp)

    r = cached_renderer(packed, nodes, device, _lbvh_mode(traversal))
    r.set_stream(None)
    film = None
    if use_dist:
        import torch
        torch.cuda.set_device(device)
        r.set_stream(torch.cuda.current_stream().cuda_stream)   # one stream: zero-fill → walk → reduce → finalize
        film = torch.zeros(H * W * 3, dtype=torch.float64, device=f"cuda:{device}")
    image = None
    if not use_dist:
        # like the reference, hand back scene.image itself when it can take the result (path_tracing_old.py:171 returns scene.image)
        si = getattr(scene, "image", None)
        if (size == 1 and film_blend is None and isinstance(si, np.ndarray) and si.dtype == np.float64 and si.shape == (H, W, 3)
                and si.flags.c_contiguous and si.flags.writeable):
            image = si
            pin_array(si, device)
        else:
            image = np.empty((H, W, 3), np.float64)
    film_out, samples, info = r.render(width=W, height=H, spp=n_spp, max_depth=D, camera=scene.camera,
                                       f_distance=scene.f_distance, s0=s0, s1=s1, semantics=semantics, traversal=traversal,
                                       real=real, rand0=rand0, rand1=rand1, light_idx=light_idx, seed=seed, film=film,
                                       return_samples=return_samples, image=image, scheduler=scheduler)
    if use_dist:
        # the single collective of a pass: lt_film_reduce (ncclReduce on the library's communicator + finalize on the root)
        parallel.reduce_film(film, group, root=None if all_ranks else 0, device=device, spp_total=n_spp)
        if all_ranks or rank == 0:
            si = getattr(scene, "image", None)
            if (film_blend is None and isinstance(si, np.ndarray) and si.dtype == np.float64 and si.shape == (H, W, 3)
                    and si.flags.c_contiguous and si.flags.writeable):
                torch.from_numpy(si.reshape(-1)).copy_(film)          # one D2H, straight into scene.image (403 MB at 4096^2)
                image = si
            else:
                image = film.cpu().numpy().reshape(H, W, 3)
    if image is not None and film_blend is not None:
        if getattr(scene, "image", None) is None:
            raise ValueError("film_blend accumulates into scene.image, which this scene does not have")
        image = np.asarray(scene.image, dtype=np.float64) + float(film_blend) * image      # fix1:166
    if image is not None and getattr(scene, "image", None) is not None and image is not getattr(scene, "image", None) and (size == 1 or use_dist):
        try:
            scene.image[...] = image
        except Exception:
            scene.image = image
    out = (image,)
    if return_samples:
        out += (samples,)
    if return_info:
        out += (info,)
    return out[0] if len(out) == 1 else out


MORE_SAMPLES_PIXELS = ((20, 60), (60, 25), (60, 80), (65, 110))      # more_samples_h / more_samples_w, R/src/path_tracing.py:313-314


def render_more_samples(scene, primitives, bvh=None, *, pixels=MORE_SAMPLES_PIXELS, n=500, rand0=None, rand1=None, light_idx=None,
                        traversal="compat", real="f64", device=0, renderer=None):
    """The "more samples" block of the reference's ``render_scene_samples`` (R/src/path_tracing.py:310-364): ``n`` extra
    samples at a few fixed pixels ``(i, j)``, each with its 4*D finite-difference replays, returned as the reference's
    5-tuple ``(samples_2 [P,n,3], record_log_pdf_2 [P,n,D], record_log_pdf_minus_set_2 [P,n,2D,D], ..._plus_set_2, record_s_set_2
    [P,n,2D])``.  Two quirks of the block are kept: it hands the UNIFORM numbers to ``calculate_gradients`` where that expects
    logits (:347), and it replays with ``calculate_grad=False`` (no gating by ``bounce_record``).  Not reproduced: the block's
    writes into ``scene.bounce_record`` / ``scene.record_log_pdf`` at sample indices up to ``n-1``, which run past
    ``number_of_samples`` in the reference.  ``rand0`` / ``rand1``: ``[P, n, D]`` uniforms (the reference draws them from numba's
    private generator inside the loop, :328-329; default: ``np.random.rand``)."""
    H, W, D = int(scene.height), int(scene.width), int(scene.max_depth)
    Pn = len(pixels)
    rand0 = np.random.rand(Pn, n, D) if rand0 is None else np.ascontiguousarray(rand0, dtype=np.float64)
    rand1 = np.random.rand(Pn, n, D) if rand1 is None else np.ascontiguousarray(rand1, dtype=np.float64)
    if rand0.shape != (Pn, n, D) or rand1.shape != (Pn, n, D):
        raise ValueError(f"rand0/rand1 must be {(Pn, n, D)}")
    own = renderer is None
    if own:
        packed = primitives if isinstance(primitives, _pack.PackedScene) else _pack.pack_scene(primitives, scene.lights)
        nodes = _nodes_from_bvh(bvh, packed) if traversal == "compat" else None
        renderer = Renderer(packed, nodes, device=device, build_lbvh=_lbvh_mode(traversal))
    try:
        parts = []
        for k, (i, j) in enumerate(pixels):
            if not (0 <= i < H and 0 <= j < W):
                raise ValueError(f"pixel {(i, j)} outside the {H} x {W} film")
            pix = i * W + j
            o = renderer.render_records(width=W, height=H, spp=n, max_depth=D, camera=scene.camera, f_distance=scene.f_distance,
                                        rand0=rand0[k:k + 1], rand1=rand1[k:k + 1],
                                        light_idx=None if light_idx is None else np.asarray(light_idx)[k:k + 1], traversal=traversal,
                                        real=real, gradients=True, lists=False, pixel_range=(pix, pix + 1), ungated=True,
                                        logits=(rand0[k:k + 1], rand1[k:k + 1]))
            parts.append(o)
    finally:
        if own:
            renderer.close()
    cat = lambda key: np.concatenate([p_[key] for p_ in parts], axis=0)
    return cat("samples"), cat("record_log_pdf"), cat("rlp_minus"), cat("rlp_plus"), cat("record_s")


def render_scene_full(scene, primitives, bvh=None, *, light_idx=None, traversal="compat", real="f64", device=0, more_samples=None):
    """The reference's CURRENT ``path_tracing.render_scene`` (R/src/path_tracing.py:371-387), whose return value
    feeds the control-variate cells of ``R/examples/LTS.ipynb``: the 10-tuple

        image_ver0, image_ver1, image_ver2, samples, record_log_pdf, record_log_pdf_minus_set,
        record_log_pdf_plus_set, record_s_set, data_more_samples, two_light_list

    and the side effects on ``scene.image`` / ``scene.bounce_record`` / ``scene.record_log_pdf``.  Consumes
    ``scene.rand_0/rand_1`` (replay).  ``data_more_samples`` is the 5-tuple of ``render_more_samples`` (the block :310-364 with
    its fixed pixels of a 150x150 film) when the film contains those pixels — ``more_samples=False`` skips it, a dict of
    keyword arguments (``pixels``, ``n``, ``rand0``, ``rand1``) customises it — else ``None``."""
    H, W, D, S = int(scene.height), int(scene.width), int(scene.max_depth), int(scene.number_of_samples)
    if getattr(scene, "rand_0", None) is None:
        raise ValueError("render_scene_full replays scene.rand_0/rand_1")
    packed = primitives if isinstance(primitives, _pack.PackedScene) else _pack.pack_scene(primitives, scene.lights)
    nodes = _nodes_from_bvh(bvh, packed) if traversal == "compat" else None
    r = Renderer(packed, nodes, device=device, build_lbvh=_lbvh_mode(traversal))
    try:
        o = r.render_records(width=W, height=H, spp=S, max_depth=D, camera=scene.camera, f_distance=scene.f_distance,
                             rand0=np.asarray(scene.rand_0), rand1=np.asarray(scene.rand_1), light_idx=light_idx,
                             traversal=traversal, real=real, gradients=True)
        more = None
        mkw = dict(more_samples) if isinstance(more_samples, dict) else {}
        pixels = mkw.get("pixels", MORE_SAMPLES_PIXELS)
        if more_samples is not False and all(0 <= i < H and 0 <= j < W for i, j in pixels):
            more = render_more_samples(scene, primitives, bvh, traversal=traversal, real=real, renderer=r, **mkw)
    finally:
        r.close()
    image_ver0 = np.array(scene.image, dtype=np.float64, copy=True)        # :373 copies scene.image BEFORE rendering
    image_ver2 = image_ver0.copy()
    image = np.clip(o["samples"].sum(axis=2) / S, 0, 1)                   # :304-305 and :378-384 (same numbers)
    for name, val in (("image", image), ("bounce_record", o["bounce_record"]), ("record_log_pdf", o["scene_record_log_pdf"])):
        try:
            getattr(scene, name)[...] = val
        except Exception:
            setattr(scene, name, val)
    return (image_ver0, image, image_ver2, o["samples"], o["record_log_pdf"], o["rlp_minus"], o["rlp_plus"], o["record_s"], more,
            (o["direct_list"], o["indirect_list"]))
