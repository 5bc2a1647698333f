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


def test_prefetch_ahead_keeps_two_sweeps_registered_with_one_call_per_sweep():
    """bench.prefetch_ahead (the replay loop of bench.py and of the benchmarked-path tests): sweeps k+1 and k+2 are registered
    before sweep k is processed, every sweep exactly once in steady state, on host and device buffers, and a new buffer list
    starts over."""
    import bench

    class Buf:
        def __init__(self, i): self.i, self.shape = i, (100 + i, 4)
        def data_ptr(self): return 0x1000 * (self.i + 1)

    class Ctx:
        def __init__(self): self.calls = []
        def prefetch_device(self, p, n, s): self.calls.append(("dev", p, n, s))
        def prefetch_ptr(self, p, n, s): self.calls.append(("host", p, n, s))

    bufs = [Buf(i) for i in range(6)]
    for device in (True, False):
        c = Ctx()
        seen = []
        for k in range(6):
            before = len(c.calls)
            bench.prefetch_ahead(c, bufs, k, device)
            seen.append([call[1] // 0x1000 - 1 for call in c.calls[before:]])
        assert seen == [[1, 2], [3], [4], [5], [], []], seen          # two at the start, then one per sweep, none past the end
        assert all(call[0] == ("dev" if device else "host") and call[3] == 4 for call in c.calls)
        assert [call[2] for call in c.calls] == [101, 102, 103, 104, 105]
    c = Ctx()
    bench.prefetch_ahead(c, bufs, 0, True)
    other = [Buf(10 + i) for i in range(4)]
    bench.prefetch_ahead(c, other, 0, True)                             # another list (a new chunk of a long replay): starts over
    assert [call[1] for call in c.calls] == [0x2000, 0x3000, 0xc000, 0xd000]
    bench.prefetch_ahead(c, other, 2, True)                             # a skipped sweep: whatever is missing of k+1, k+2
    assert [call[1] for call in c.calls[4:]] == [0xe000]


def test_rank_core_slices_are_disjoint_and_interleaved(monkeypatch):
    """bench.pin_rank_threads: every local rank gets its own slice of the host cores, interleaved (rank r: r, r + N, ...), so
    that hyper-thread siblings i and i + n/2 stay inside one rank; one rank alone is not pinned."""
    import bench
    if not hasattr(os, "sched_setaffinity"):
        pytest.skip("no sched_setaffinity")
    cores = list(range(32))
    pinned = {}
    monkeypatch.setattr(os, "sched_getaffinity", lambda pid: set(cores))
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, s: pinned.__setitem__("last", list(s)))
    assert bench.pin_rank_threads(0, 1) is None
    for n in (2, 4, 8):
        slices = [bench.pin_rank_threads(r, n) for r in range(n)]
        flat = sorted(x for s in slices for x in s)
        assert flat == cores and all(len(s) == 32 // n for s in slices)
        for r, s in enumerate(slices):
            assert s == list(range(r, 32, n))
            assert all(((x + 16) % 32) in s for x in s)               # the sibling of every core belongs to the same rank
    monkeypatch.setenv("VLOAM_PIN", "block")
    assert bench.pin_rank_threads(1, 4) == list(range(8, 16))
