"""C4b — many-bounce photon walk in a homogeneous scattering medium (BASELINE.json config 4).

The reference's only medium code is the Henyey-Greenstein evaluation ``henyey_greenstein(cosTheta, g)``
(``R/src/medium_samples.py:14-16``; PBRT's sign convention ``1 + g^2 + 2 g cos``); it is mirrored here and
pinned bit-for-bit by a golden table.  The walk itself (``lt_medium_walk``: free flights, analog absorption
or implicit capture + the reference's roulette rule, HG scattering, shared-memory-staged tallies) is beyond the
reference and is checked against the oracle and analytic expectations (``tests/test_gpu_medium.py``).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import host

inv_4_pi = 0.25 * (1 / np.pi)      # R/src/constants.py:9


def henyey_greenstein(cosTheta, g):
    """R/src/medium_samples.py:14-16, verbatim semantics."""
    denom = 1 + g**2 + 2 * g * cosTheta
    return inv_4_pi * (1 - g**2) / (denom * np.sqrt(denom))


class LtMediumParams(C.Structure):
    _fields_ = [("albedo", C.c_double), ("g", C.c_double), ("radius", C.c_double), ("max_depth", C.c_int32), ("mode", C.c_int32),
                ("rr_start", C.c_int32), ("n_rbins", C.c_int32), ("rr_floor", C.c_double), ("seed", C.c_uint64),
                ("first_photon", C.c_uint64), ("real", C.c_int32), ("reserved", C.c_int32)]


class LtMediumTotals(C.Structure):
    _fields_ = [(k, C.c_double) for k in ("escaped", "absorbed", "truncated", "rr_killed", "rr_gained", "sum_mu")] + \
               [(k, C.c_uint64) for k in ("n_collisions", "n_segments", "n_scatters")]


def walk(n_photons, *, albedo=0.99, g=0.0, radius=0.0, max_depth=256, mode="implicit", rr_start=3, rr_floor=0.05, n_rbins=32,
         seed=0, first_photon=0, real="f32", device=0):
    """Run ``n_photons`` photon histories on the GPU.  Returns a dict: ``hist[max_depth+1]`` (photons by collision
    count), ``rbins[n_rbins]`` (absorbed weight by radial shell), the weight totals (``escaped``, ``absorbed``,
    ``truncated``, ``rr_killed``, ``rr_gained``), ``sum_mu``, the counters and ``kernel_ms``.
    ``mode``: "analog" (absorb with probability 1-albedo) or "implicit" (deposit w(1-albedo), roulette on w)."""
    L = host.lib()
    L.lt_medium_walk.restype = C.c_int
    ctx = host._get_ctx(device)
    p = LtMediumParams(float(albedo), float(g), float(radius), int(max_depth), 0 if mode == "analog" else 1, int(rr_start), int(n_rbins),
                       float(rr_floor), int(seed), int(first_photon), host.REAL[real], 0)
    hist = np.zeros(max_depth + 1, np.uint64)
    rbins = np.zeros(max(n_rbins, 1), np.float64)
    tot = LtMediumTotals()
    ms = C.c_float()
    host._check(L.lt_medium_walk(ctx, C.byref(p), C.c_uint64(int(n_photons)), hist.ctypes.data_as(C.c_void_p),
                                 rbins.ctypes.data_as(C.c_void_p), C.byref(tot), C.byref(ms)))
    out = {k: getattr(tot, k) for k, _ in LtMediumTotals._fields_}
    out.update(hist=hist, rbins=rbins[:n_rbins], kernel_ms=float(ms.value), n_photons=int(n_photons))
    return out
