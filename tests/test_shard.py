"""Multi-GPU host logic on CPU: the batch split of SURVEY.md §8e (independent maps, no data-path collective),
exercised with a world_size-2 gloo process group."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _shard_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("lsdb200_shard", os.path.join(ROOT, "linesegmentdetector-slam_b200", "shard.py"))
    sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
    return sh


def test_shard_range_covers_batch_exactly():
    sh = _shard_module()
    for n in (0, 1, 5, 6, 256, 10000):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, cnt = sh.shard_range(n, r, world)
                seen.extend(range(first, first + cnt))
                for i in range(first, first + cnt):
                    assert sh.owner_of(i, n, world) == r
            assert seen == list(range(n))
            sizes = [sh.shard_range(n, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sh.shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    import importlib.util
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    spec = importlib.util.spec_from_file_location("shard", os.path.join(ROOT, "linesegmentdetector-slam_b200", "shard.py"))
    sh = importlib.util.module_from_spec(spec); spec.loader.exec_module(sh)
    n = 7
    first, cnt = sh.shard_range(n, rank, world)
    # stand-in for the per-rank LSD results: map i yields i+1 segments with a recognisable payload
    counts = [i + 1 for i in range(first, first + cnt)]
    tables = [np.full((i + 1, 10), float(i)) for i in range(first, first + cnt)]
    full = sh.gather_counts(counts, n)
    tot = sh.all_reduce_scalar(sum(counts))
    mx = sh.all_reduce_scalar(10.0 + rank, op="max")
    merged = sh.gather_segments(tables, n, dst=0)
    ok = list(full) == [i + 1 for i in range(n)] and tot == sum(range(1, n + 1)) and mx == 10.0 + world - 1
    if rank == 0:
        ok = ok and len(merged) == n and all(merged[i].shape == (i + 1, 10) and merged[i][0, 0] == i for i in range(n))
    else:
        ok = ok and merged is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_gathers_in_batch_order():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    assert res == {0: True, 1: True}


def test_c_abi_shard_rule_equals_shard_range(lsdb):
    """lsdb_multi_shard (the split of the one-process multi-device entry point) is shard.shard_range."""
    import ctypes as C
    from lsdb200 import shard
    L = lsdb.lib()
    for n in (0, 1, 5, 8, 255, 256, 1000):
        for k in (1, 2, 3, 8):
            for d in range(k):
                a, b = C.c_int(-1), C.c_int(-1)
                L.lsdb_multi_shard(n, d, k, C.byref(a), C.byref(b))
                assert (a.value, b.value) == shard.shard_range(n, d, k)
