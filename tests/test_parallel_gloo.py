"""CPU, world_size 2 over gloo: the N>1 path — disjoint sample ranges per rank, ONE film reduce —
gives the single-rank film.  The per-rank renders come from the oracle here (no GPU in this suite);
the GPU twin is tests/test_gpu_render.py::test_virtual_ranks_equal_full."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import REPO, load_scene


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from light_transport_b200 import parallel
    from oracle import oracle as O
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    S = O.OracleScene(ps, nodes)
    spp = 6
    s0, s1 = parallel.sample_range(rank, world, spp)
    film, _, cnt = O.render(S, width=20, height=12, spp=spp, s0=s0, s1=s1, max_depth=8, camera=cam, f_distance=f, seed=3,
                            n_threads=1)
    t = torch.from_numpy(film.reshape(-1).copy())
    parallel.reduce_film(t)                      # all-reduce: every rank ends with the full film
    np.save(os.path.join(out_dir, f"film_{rank}.npy"), t.numpy())
    t2 = torch.from_numpy(film.reshape(-1).copy())
    parallel.reduce_film(t2, root=0, spp_total=spp)     # reduce to rank 0 + fused finalize there (the GPU path: lt_film_reduce)
    np.save(os.path.join(out_dir, f"image_{rank}.npy"), t2.numpy())
    np.save(os.path.join(out_dir, f"paths_{rank}.npy"), np.array([cnt["paths"]]))
    dist.destroy_process_group()


def test_two_rank_film_reduce(tmp_path, oracle):
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ps, nodes, cam, f = load_scene("lts_scene.npz")
    S = oracle.OracleScene(ps, nodes)
    full, _, cnt = oracle.render(S, width=20, height=12, spp=6, max_depth=8, camera=cam, f_distance=f, seed=3)
    f0 = np.load(tmp_path / "film_0.npy")
    f1 = np.load(tmp_path / "film_1.npy")
    assert np.array_equal(f0, f1)
    assert np.allclose(f0.reshape(full.shape), full, rtol=1e-12, atol=1e-12)
    img0 = np.load(tmp_path / "image_0.npy")
    assert np.allclose(img0.reshape(full.shape), np.clip(full / 6, 0, 1), rtol=1e-12, atol=1e-12)     # R/src/path_tracing.py:304-305
    assert int(np.load(tmp_path / "paths_0.npy")[0] + np.load(tmp_path / "paths_1.npy")[0]) == cnt["paths"]
