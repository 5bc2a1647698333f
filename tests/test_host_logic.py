"""Host-side logic without a GPU: planted-map layout, the python mirror's error behaviour, and the
N > 1 sequence sharding + gather over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest


def test_cubes_blob_layout_roundtrip(op, synth):
    """synth.cubes_blob follows laser_mapping.cpp:747-761: the oracle re-exports it unchanged and one
    VoxelGrid pass leaves a planted (one point per voxel, voxel-ordered) cube untouched."""
    rng = np.random.RandomState(1)
    pts = np.zeros((20000, 4), np.float32)
    pts[:, :3] = (rng.rand(20000, 3) - 0.5) * [400, 400, 60]
    blob = synth.cubes_blob(pts, 0.8)
    o = op.Oracle()
    o.set("lm.surfMap", blob)
    assert o.get("lm.surfMap") == blob
    counts = synth.blob_counts(blob)
    cloud = np.frombuffer(blob[len(counts) * 4:], np.float32).reshape(-1, 4)
    # every cube is voxel-sorted and deduplicated: filtering it again is the identity
    off = 0
    checked = 0
    for c in range(len(counts)):
        n = counts[c]
        if n and checked < 60:
            cube = cloud[off:off + n]
            again = op.voxel_grid(cube, 0.8)
            assert again.shape == cube.shape and (again == cube).all()
            checked += 1
        off += n
    assert checked > 10


def test_sequence_sharding(pkg):
    par = __import__("importlib").import_module("vloam-noted_b200.parallel")
    for g in (1, 2, 4, 8):
        got = sorted(s for r in range(g) for s in par.sequences_of_rank(8, r, g))
        assert got == list(range(8))
        assert all(len(par.sequences_of_rank(8, r, g)) == 8 // g for r in range(g))
    assert par.aggregate_throughput([100.0, 200.0], 50) == pytest.approx(2 * 50 / 0.2)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import importlib
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    par = importlib.import_module("vloam-noted_b200.parallel")
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    seqs = par.sequences_of_rank(8, rank, world)
    poses = np.full((5, 14), float(rank)) + np.arange(5)[:, None]
    tim = np.array([10.0 * (rank + 1), len(seqs)])
    gp, gt = par.gather_results(dist, poses, tim)
    q.put((rank, gp.shape, gp[:, 0, 0].tolist(), gt[:, 0].tolist(), seqs))
    dist.destroy_process_group()


def test_gather_over_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs: p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs: p.join(timeout=60)
    for rank, shape, firsts, tims, seqs in res:
        assert shape == (2, 5, 14) and firsts == [0.0, 1.0] and tims == [10.0, 20.0]
        assert seqs == [s for s in range(8) if s % 2 == rank]


def test_context_fails_loudly_without_gpu(pkg):
    """No CPU fallback: on a box without a CUDA device creating a context raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pkg.load_lib(build=True)
    with pytest.raises(pkg.VloamError):
        pkg.Context()
