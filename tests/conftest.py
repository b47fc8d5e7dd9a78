import os
import sys

import numpy as np
import pytest

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_scene(name):
    """(PackedScene, CompatNodes, camera[3], f_distance) from a golden scene file."""
    from light_transport_b200 import _pack, compat_bvh
    d = np.load(os.path.join(GOLDEN, name))
    ps = _pack.PackedScene.from_npz_dict(d)
    nodes = compat_bvh.CompatNodes(d["node_bmin"], d["node_bmax"], d["node_prim_off"], d["node_second"],
                                   d["node_n_prims"], d["node_axis"])
    return ps, nodes, d["camera"], float(d["f_distance"])


def load_paths(name):
    """golden per-path radiance + the regenerated replay inputs"""
    g = np.load(os.path.join(GOLDEN, name))
    W, H, spp, D = int(g["W"]), int(g["H"]), int(g["spp"]), int(g["D"])
    np.random.seed(int(g["rand_seed"]))          # legacy MT19937 stream: stable across numpy versions
    r0 = np.random.rand(H, W, spp, D)
    r1 = np.random.rand(H, W, spp, D)
    return dict(W=W, H=H, spp=spp, D=D, rand0=r0, rand1=r1, light_idx=g["light_idx"].astype(np.int32),
                radiance=g["radiance"], camera=g["camera"], f_distance=float(g["f_distance"]),
                semantics=str(g["semantics"]), image=g["image"])


PATH_SETS = [("lts_scene.npz", "paths_lts_ref_d8.npz"), ("lts_scene.npz", "paths_lts_ref_d16.npz"),
             ("lts_scene.npz", "paths_lts_iter_d8.npz"), ("mirror_scene.npz", "paths_mirror_ref_d256.npz")]
RAY_SETS = [("lts_scene.npz", "rays_lts.npz"), ("mirror_scene.npz", "rays_mirror.npz")]


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def rel_err(a, b, floor=1e-3):
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
