"""CPU tier: the N>1 host path (static sharding + best-plan all-gather) on gloo, world_size 2."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import qtos_b200 as Q
from qtos_b200 import parallel


def _fake_results(n, seed):
    rng = np.random.default_rng(seed)
    r = np.zeros(n, dtype=Q.RESULT_DTYPE)
    r["status"] = np.where(rng.uniform(size=n) < 0.2, -1, 0)
    r["cost"] = np.round(rng.uniform(1, 2, n), 3)          # ties on purpose
    r["constr_viol"] = rng.uniform(0, 1e-4, n)
    return r


def _worker(rank, world, port, n_total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    allres = _fake_results(n_total, 7)                      # what a single process would have computed
    idx = parallel.shard_indices(n_total, rank, world)
    rec = parallel.make_records(allres[idx], idx, idx // 8)
    winners, gathered = parallel.select_best(rec)
    out[rank] = (winners, len(gathered))
    dist.destroy_process_group()


def test_shard_and_select_best_world2():
    n_total = 64
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(2, port, n_total, out), nprocs=2, join=True)
    allres = _fake_results(n_total, 7)
    want = parallel.argmin_per_group(parallel.make_records(allres, np.arange(n_total), np.arange(n_total) // 8))
    assert out[0][0] == want and out[1][0] == want and out[0][1] == n_total
    # winners are converged whenever the group has a converged candidate, and minimal in cost among them
    for g, w in want.items():
        members = np.arange(8 * g, 8 * g + 8)
        conv = members[allres["status"][members] == 0]
        if len(conv):
            assert allres["status"][w] == 0 and allres["cost"][w] == allres["cost"][conv].min()


def test_shards_partition_the_batch():
    for world in (1, 2, 4, 8):
        parts = [parallel.shard_indices(4096 * world, r, world) for r in range(world)]
        assert all(len(p) == 4096 for p in parts)
        assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(4096 * world))
        for p in parts:                                      # every group of 8 candidates is split evenly
            assert np.all(np.bincount(p // 8) == 8 // world if world <= 8 else True)


def test_segmented_selection_equals_the_sorted_reference():
    """argmin_per_group (radix sort of group ids + segmented minima) against the one-sort reference, with heavy ties on
    every key, ragged groups, singleton groups and one group holding everything."""
    rng = np.random.default_rng(7)
    for n, gs in ((4096, 8), (1000, 7), (64, 1), (500, 500), (1, 1)):
        rec = np.empty((n, 5))
        rec[:, 0] = np.arange(n) // gs
        rec[:, 1] = rng.random(n) < 0.3
        rec[:, 2] = np.round(rng.random(n) * 3)
        rec[:, 3] = np.round(rng.random(n) * 2) * 1e-5
        rec[:, 4] = np.arange(n)
        rec = rec[rng.permutation(n)]
        assert parallel.argmin_per_group(rec) == parallel.argmin_per_group_sorted(rec)
    assert parallel.argmin_per_group(np.empty((0, 5))) == {}


def test_torch_selection_equals_numpy():
    """the tensor version used on the gathered records (GPU under NCCL, CPU under gloo)."""
    rng = np.random.default_rng(11)
    for n, gs in ((4096, 8), (999, 13), (32, 1)):
        rec = np.empty((n, 5))
        rec[:, 0] = 5 + 3 * (np.arange(n) // gs)            # sparse group ids
        rec[:, 1] = rng.random(n) < 0.3
        rec[:, 2] = np.round(rng.random(n) * 3)
        rec[:, 3] = np.round(rng.random(n) * 2) * 1e-5
        rec[:, 4] = np.arange(n)
        rec = rec[rng.permutation(n)]
        grp, win = parallel.argmin_per_group_torch(torch.from_numpy(rec))
        assert dict(zip(grp.tolist(), win.tolist())) == parallel.argmin_per_group(rec)


def test_nan_records_do_not_win_or_poison_a_group():
    """a status -13 window carries NaN cost / violation: every selector must still name a valid member of each group
    (the minimum of a NaN key used to come back as INT64_MIN)."""
    import torch
    from qtos_b200 import parallel
    nan = float("nan")
    rec = np.array([[0, 1, nan, nan, 0], [0, 1, 2.0, 0.1, 1],            # no candidate converged, one is all-NaN
                    [1, 1, nan, nan, 2], [1, 0, 5.0, 0.0, 3],            # NaN beside a converged candidate
                    [2, 1, nan, nan, 4]], dtype=np.float64)             # a group of one NaN window: it still names itself
    want = {0: 1, 1: 3, 2: 4}
    assert parallel.argmin_per_group_sorted(rec) == want
    assert parallel.argmin_per_group(rec) == want
    g, w = parallel.argmin_per_group_torch(torch.from_numpy(rec))
    assert dict(zip(g.tolist(), w.tolist())) == want
    import qtos_b200 as Q
    r = np.zeros(2, dtype=Q.RESULT_DTYPE); r["status"] = (-13, 0); r["cost"] = (nan, 1.0); r["constr_viol"] = (nan, 0.0)
    out = parallel.make_records(r, np.array([7, 8]), np.array([0, 0]))
    assert np.isinf(out[0, 2]) and np.isinf(out[0, 3]) and not np.isnan(out).any()
