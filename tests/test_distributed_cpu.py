"""world_size-2 gloo test of the multi-GPU host logic (scenario sharding + the single gather of result records).
The compute itself needs a GPU; here each rank fabricates its shard's records so that the collective path, the
ragged shard sizes and the scenario order are exercised on CPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, n_total, port, out):
    sys.path.insert(0, ROOT)
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat
    from otg_b200.batch import gather_results, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    rec = np.zeros(hi - lo, dtype=np.dtype(nat.RESULT_DTYPE))
    rec["argmin_ctot"] = np.arange(lo, hi)
    rec["argmin_masked"] = np.arange(lo, hi) * 2 - 1
    rec["status"] = rank
    rec["min_ctot"] = np.arange(lo, hi) * 0.5
    allrec = gather_results(rec, n_total)
    ok = (len(allrec) == n_total and np.array_equal(allrec["argmin_ctot"], np.arange(n_total))
          and np.array_equal(allrec["argmin_masked"], np.arange(n_total) * 2 - 1)
          and np.array_equal(allrec["min_ctot"], np.arange(n_total) * 0.5)
          and all(np.all(allrec["status"][slice(*shard_range(n_total, r, world))] == r) for r in range(world)))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [7, 64])
def test_gather_results_world2_gloo(n_total):
    world = 2
    port = 29500 + (os.getpid() + n_total) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, n_total, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _records_worker(rank, world, n_total, port, out):
    """BatchPlanner.gather_records with RAGGED shards: every rank must take the padded path (round-1 advisor finding: the long
    ranks used to take the device path and return world * cap records).  The planner is a stub without an engine - only the
    branch selection and the collective are under test."""
    sys.path.insert(0, ROOT)
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat
    from otg_b200.batch import BatchPlanner, shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, rank, world)
    rec = np.zeros(hi - lo, dtype=np.dtype(nat.RESULT_DTYPE))
    rec["argmin_masked"] = np.arange(lo, hi)
    rec["n_valid"] = rank + 1
    bp = object.__new__(BatchPlanner)
    bp.n_scen, bp.engine = hi - lo, None
    bp._local_records = lambda: rec
    bp._collective_device = lambda group=None: None
    allrec = bp.gather_records(n_total)
    ok = len(allrec) == n_total and np.array_equal(allrec["argmin_masked"], np.arange(n_total)) and \
        all(np.all(allrec["n_valid"][slice(*shard_range(n_total, r, world))] == r + 1) for r in range(world))
    try:
        bp.gather_records(n_total, sync=False)
        ok = False
    except ValueError:
        pass
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 7), (3, 7)])
def test_gather_records_ragged_shards_gloo(world, n_total):
    port = 31500 + (os.getpid() + 7 * world + n_total) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_records_worker, args=(world, n_total, port, out), nprocs=world, join=True)
        assert dict(out) == {r: True for r in range(world)}


def test_gather_without_process_group_is_identity():
    sys.path.insert(0, ROOT)
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat
    from otg_b200.batch import gather_results
    rec = np.zeros(5, dtype=np.dtype(nat.RESULT_DTYPE))
    assert gather_results(rec, 5) is rec


def _sized_worker(rank, world, sizes, port, out):
    sys.path.insert(0, ROOT)
    import otg_b200  # noqa: F401
    from otg_b200 import _native as nat
    from otg_b200.batch import gather_results, shard_range_from_sizes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = sum(sizes)
    lo, hi = shard_range_from_sizes(sizes, rank)
    rec = np.zeros(hi - lo, dtype=np.dtype(nat.RESULT_DTYPE))
    rec["argmin_ctot"] = np.arange(lo, hi)
    rec["status"] = rank
    allrec = gather_results(rec, n_total, sizes=sizes)
    ok = (len(allrec) == n_total and np.array_equal(allrec["argmin_ctot"], np.arange(n_total))
          and all(np.all(allrec["status"][slice(*shard_range_from_sizes(sizes, r))] == r) for r in range(world)))
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_with_speed_balanced_shards_gloo():
    """Shards sized by measured GPU speed (bench.py at N > 1): the padded gather returns every rank's records in scenario order."""
    world, sizes = 2, [9, 5]
    port = 29500 + (os.getpid() + 977) % 2000
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_sized_worker, args=(world, sizes, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_balanced_shard_sizes():
    sys.path.insert(0, ROOT)
    import otg_b200  # noqa: F401
    from otg_b200.batch import balanced_shard_sizes, shard_range_from_sizes
    t = [5.032, 5.21, 5.086, 5.349, 5.262, 5.337, 5.148, 5.112]        # the eight GPUs of one box, equal shards of 512 (profiles/r2_history.md)
    s = balanced_shard_sizes(4096, t)
    assert sum(s) == 4096 and len(s) == 8 and all(abs(v - 512) <= 52 for v in s)
    assert s[0] == max(s) and s[3] == min(s)                             # fastest GPU gets the most, slowest the least
    assert max(np.array(s) * np.array(t)) / 512 < 0.98 * max(t)           # the slowest rank's time drops by > 2 %
    assert balanced_shard_sizes(1024, [5.0, 5.0]) == [512, 512]
    assert balanced_shard_sizes(1024, [5.0, 9.0]) == [563, 461]          # clipped at +-10 % of the equal share
    assert sum(balanced_shard_sizes(7, [1.0, 2.0, 3.0])) == 7
    assert [shard_range_from_sizes(s, r) for r in (0, 1, 7)] == [(0, s[0]), (s[0], s[0] + s[1]), (4096 - s[7], 4096)]
