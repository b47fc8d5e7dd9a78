"""C4b — homogeneous scattering medium.  CPU part: the Henyey-Greenstein evaluation against the reference's own
function (golden table), and the oracle's walk against analytic expectations.  GPU part (marked): lt_medium_walk
against the oracle on the same Philox streams, and the same analytic properties at BASELINE's size
(albedo 0.99, max depth 256)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_hg_matches_reference_function(oracle):
    """R/src/medium_samples.py:14-16 evaluated by the reference itself -> golden table"""
    from light_transport_b200 import medium
    t = np.load(os.path.join(GOLDEN, "hg_table.npz"))
    for gi, g in enumerate(t["g"]):
        mine = medium.henyey_greenstein(t["cos_theta"], g)
        assert np.array_equal(mine, t["value"][gi])                              # same expression, same bits
        orc = np.array([oracle.henyey_greenstein(c, g) for c in t["cos_theta"]])
        assert np.allclose(orc, t["value"][gi], rtol=1e-15, atol=0)
    # it is a normalised phase function: integral over the sphere = 1 (mu-quadrature)
    mu = np.linspace(-1, 1, 200001)
    for g in (0.0, 0.3, 0.8):
        assert abs(np.trapezoid(medium.henyey_greenstein(mu, g), mu) * 2 * np.pi - 1) < 1e-6


def _check_physics(r, albedo, g, D, radius, mode, bal_tol=1e-9):
    n = r["n_photons"]
    assert int(r["hist"].sum()) == n
    bal = (r["escaped"] + r["absorbed"] + r["truncated"] + r["rr_killed"] - r["rr_gained"]) / n
    assert abs(bal - 1) < bal_tol                                  # weight bookkeeping per history (float32: rounding of w*a + w*(1-a))
    if mode == "analog":
        assert r["rr_killed"] == 0 and r["rr_gained"] == 0
    assert abs(r["rbins"].sum() - r["absorbed"]) < 1e-6 * max(r["absorbed"], 1)
    if r["n_scatters"] > 1000:
        mean_mu = r["sum_mu"] / r["n_scatters"]                    # <cos theta> of Henyey-Greenstein is g
        assert abs(mean_mu - g) < 5 * np.sqrt((1 - g * g) / r["n_scatters"]) + 1e-3
    if radius <= 0 and mode == "analog":
        # infinite medium: collisions until absorption are geometric, truncated at D
        expect = (1 - albedo**D) / (1 - albedo)
        k = np.arange(D + 1)
        mean = (r["hist"] * k).sum() / n
        var = ((r["hist"] * k**2).sum() / n) - mean**2
        assert abs(mean - expect) < 5 * np.sqrt(var / n)
        assert abs(r["truncated"] / n - albedo**D) < 5 * np.sqrt(albedo**D / n) + 1e-4
        assert r["escaped"] == 0
    if radius <= 0 and mode == "implicit":
        assert r["escaped"] == 0


@pytest.mark.parametrize("mode,radius,g", [("analog", 0.0, 0.8), ("analog", 10.0, 0.0), ("implicit", 10.0, 0.8), ("implicit", 0.0, 0.3)])
def test_oracle_medium_physics(oracle, mode, radius, g):
    r = oracle.medium_walk(60000, albedo=0.99, g=g, radius=radius, max_depth=256, mode=mode, seed=11)
    _check_physics(r, 0.99, g, 256, radius, mode)
    again = oracle.medium_walk(60000, albedo=0.99, g=g, radius=radius, max_depth=256, mode=mode, seed=11)
    assert np.array_equal(again["hist"], r["hist"])               # counter-based: deterministic
    # disjoint photon ranges add up (the multi-GPU decomposition)
    a = oracle.medium_walk(20000, albedo=0.99, g=g, radius=radius, max_depth=256, mode=mode, seed=11, first_photon=0)
    b = oracle.medium_walk(40000, albedo=0.99, g=g, radius=radius, max_depth=256, mode=mode, seed=11, first_photon=20000)
    assert np.array_equal(a["hist"] + b["hist"], r["hist"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode,radius,g", [("analog", 0.0, 0.8), ("analog", 10.0, 0.0), ("implicit", 10.0, 0.8), ("implicit", 0.0, 0.3)])
def test_gpu_medium_f64_replays_oracle(oracle, mode, radius, g):
    from light_transport_b200 import medium
    n = 200000
    kw = dict(albedo=0.99, g=g, radius=radius, max_depth=256, mode=mode, seed=5, n_rbins=32)
    r = medium.walk(n, real="f64", **kw)
    o = oracle.medium_walk(n, **kw)
    _check_physics(r, 0.99, g, 256, radius, mode)
    # same Philox streams: histories coincide except where libm/CUDA ulps flip a boundary decision
    assert np.abs(r["hist"].astype(np.int64) - o["hist"].astype(np.int64)).sum() <= 2e-5 * n + 2
    assert abs(int(r["n_segments"]) - int(o["n_segments"])) <= 1e-5 * o["n_segments"] + 300
    for k in ("escaped", "absorbed", "truncated"):
        assert abs(r[k] - o[k]) <= 1e-5 * n + 1e-9
    assert np.allclose(r["rbins"], o["rbins"], rtol=1e-4, atol=1e-5 * n)


@pytest.mark.gpu
def test_gpu_medium_f32_baseline_config():
    """BASELINE config 4: albedo 0.99, max depth 256 — float32, analytic checks + 3 sigma against float64"""
    from light_transport_b200 import medium
    n = 2_000_000
    for mode, radius, g in (("analog", 0.0, 0.8), ("implicit", 20.0, 0.8)):
        kw = dict(albedo=0.99, g=g, radius=radius, max_depth=256, mode=mode, n_rbins=32)
        r32 = medium.walk(n, real="f32", seed=1, **kw)
        r64 = medium.walk(n, real="f64", seed=2, **kw)
        _check_physics(r32, 0.99, g, 256, radius, mode, bal_tol=2e-6)
        k = np.arange(257)
        m32, m64 = (r32["hist"] * k).sum() / n, (r64["hist"] * k).sum() / n
        var = (r64["hist"] * k**2).sum() / n - m64**2
        assert abs(m32 - m64) < 5 * np.sqrt(2 * var / n)
        assert abs(r32["escaped"] - r64["escaped"]) / n < 5 * np.sqrt(0.25 * 2 / n) + 1e-4
        print(f"{mode} R={radius}: {r32['n_segments']/r32['kernel_ms']/1e3:.0f} Msegments/s f32, {r64['n_segments']/r64['kernel_ms']/1e3:.0f} f64, "
              f"mean collisions {m32:.2f}")
    # photon-range additivity on the GPU (multi-GPU decomposition) and determinism
    a = medium.walk(300000, real="f32", seed=9, first_photon=0, **kw)
    b = medium.walk(700000, real="f32", seed=9, first_photon=300000, **kw)
    c = medium.walk(1000000, real="f32", seed=9, **kw)
    assert np.array_equal(a["hist"] + b["hist"], c["hist"])
    with pytest.raises(Exception):
        medium.walk(10, albedo=1.5)
