"""light_transport_b200 — B200-native drop-in for the Monte Carlo path walk of
zhouyifan233/light-transport (``render_scene(scene, primitives, bvh)``,
``R/src/path_tracing.py:371-387``).  See DESIGN.md.

The CUDA library (``csrc/`` → ``liblt_b200.so``) is loaded on first use; there is no CPU
fallback — calls raise if the library or a GPU is missing.
"""
from . import api, scenes, compat_bvh  # noqa: F401
from ._pack import PackedScene, pack_scene, from_arrays  # noqa: F401
from .mesh_io import load_obj  # noqa: F401

__all__ = ["render_scene", "render_scene_full", "render_more_samples", "intersect", "Renderer", "api", "scenes", "compat_bvh", "PackedScene",
           "pack_scene", "from_arrays", "load_obj"]


def __getattr__(name):
    if name in ("render_scene", "render_scene_full", "render_more_samples", "intersect", "Renderer", "lib"):
        from . import host
        return getattr(host, name)
    raise AttributeError(name)
