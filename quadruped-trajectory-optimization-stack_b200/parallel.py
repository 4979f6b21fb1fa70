"""Multi-GPU plumbing: one process per GPU, independent solves sharded statically, and ONE
all-gather of small per-candidate records where candidates compete (multi-start goals of the same
window).  torch.distributed is used for the rendezvous/collective only (NCCL on GPUs, gloo on CPU).

The reference has no collective at all: it fans solves out over 32 OS processes
(ref: QTOS/generateHeightField.py:18,344-404) and never compares plans; best-plan selection is the
exchange step BASELINE.json's north_star introduces.  Key = (not converged, cost, violation, global id),
lexicographic, so the winner is identical on every rank and for every GPU count."""
import numpy as np
import torch
import torch.distributed as dist


def shard_indices(n_total, rank, world):
    """interleaved static shard: candidate c lives on rank c % world (groups of consecutive
    candidates are spread evenly over the ranks)."""
    return np.arange(rank, n_total, world, dtype=np.int64)


def make_records(results, global_idx, groups):
    """float64 [n, 5] = (group, not_converged, cost, violation, global id)"""
    rec = np.empty((len(results), 5), dtype=np.float64)
    rec[:, 0] = groups
    rec[:, 1] = (results["status"] != 0).astype(np.float64)
    # a window that ended on non-finite values (status -13) carries NaN metrics: +inf keeps every selector's minimum defined
    rec[:, 2] = np.nan_to_num(results["cost"], nan=np.inf)
    rec[:, 3] = np.nan_to_num(results["constr_viol"], nan=np.inf)
    rec[:, 4] = global_idx
    return rec


def argmin_per_group_sorted(rec):
    """reference implementation: one lexicographic sort of all records (O(n log n), ~20 ms for 32768 records)."""
    rec = np.nan_to_num(rec, nan=np.inf, posinf=np.inf)
    order = np.lexsort((rec[:, 4], rec[:, 3], rec[:, 2], rec[:, 1], rec[:, 0]))
    srt = rec[order]
    first = np.ones(len(srt), dtype=bool)
    first[1:] = srt[1:, 0] != srt[:-1, 0]
    return {int(g): int(i) for g, i in zip(srt[first, 0], srt[first, 4])}


def argmin_per_group(rec):
    """rec [n, 5] (any order) -> {group: winning global id}; deterministic lexicographic key
    (not converged, cost, violation, global id).  Segmented minimum, one key after the other over the still-tied
    candidates: a radix sort of the integer group ids plus O(n) passes (the selection sits inside the timed step at
    every GPU count, and the gathered record count grows with the number of ranks)."""
    if len(rec) == 0:
        return {}
    order = np.argsort(rec[:, 0].astype(np.int64), kind="stable")
    srt = np.nan_to_num(rec[order], nan=np.inf, posinf=np.inf)
    grp = srt[:, 0]
    starts = np.flatnonzero(np.concatenate(([True], grp[1:] != grp[:-1])))
    counts = np.diff(np.concatenate((starts, [len(srt)])))
    tied = np.ones(len(srt), dtype=bool)
    best = None
    for col in (1, 2, 3, 4):
        key = np.where(tied, srt[:, col], np.inf)
        best = np.minimum.reduceat(key, starts)
        tied &= key == np.repeat(best, counts)
    return dict(zip(grp[starts].astype(np.int64).tolist(), best.astype(np.int64).tolist()))


def argmin_per_group_torch(rec):
    """the same selection on a torch tensor [n, 5] (device-resident after the all-gather): dense group index by
    torch.unique, then one segmented minimum (scatter_reduce amin) per key over the still-tied candidates.
    Returns (groups int64 [G], winning global ids int64 [G]) on the tensor's device."""
    rec = torch.nan_to_num(rec, nan=float("inf"), posinf=float("inf"))
    grp, inv = torch.unique(rec[:, 0].to(torch.int64), return_inverse=True)
    tied = torch.ones(rec.shape[0], dtype=torch.bool, device=rec.device)
    inf = torch.tensor(float("inf"), dtype=rec.dtype, device=rec.device)
    best = None
    for col in (1, 2, 3, 4):
        key = torch.where(tied, rec[:, col], inf)
        best = torch.full((grp.shape[0],), float("inf"), dtype=rec.dtype, device=rec.device).scatter_reduce(0, inv, key, "amin")
        tied &= key == best[inv]
    return grp, best.to(torch.int64)


def select_best(rec_local, device=None, host_group=None):
    """all-gather the records of every rank (equal counts per rank) and pick the winner of each group
    on every rank identically.  Returns (winners dict, gathered records or None).  With a process group the
    selection runs on the gathered tensor where it lands (on the GPU under NCCL): only the winners come back."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return argmin_per_group(rec_local), rec_local
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(rec_local))
    if device is not None:
        t = t.to(device)
    if host_group is not None:
        dist.barrier(group=host_group)                 # line the ranks up on the host: a waiting NCCL kernel spins on the GPU
    out = torch.empty((world * t.shape[0], t.shape[1]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t)
    grp, win = argmin_per_group_torch(out)
    gw = torch.stack((grp, win)).cpu().numpy()
    return dict(zip(gw[0].tolist(), gw[1].tolist())), out


def select_best_device(solver, d_res, d_group, rank, world, n_groups, host_group=None):
    """The selection without a host round trip: records from the device-resident results (k_records), one all-gather over
    NCCL, winners by the library's hand-written kernel (k_select).  d_res: uint8 [n, sizeof(qtos_result)], d_group: int32 [n]
    with dense group ids in [0, n_groups); candidate i of this rank has global id rank + world * i (shard_indices).
    `host_group` (a gloo process group) lines the ranks up on the host first: an NCCL kernel that waits for a late rank spins ON
    the GPU beside the solver's kernels, a host barrier does not.
    Returns the winning global ids as an int64 tensor [n_groups] on the device (-1 for an empty group)."""
    n = d_res.shape[0]
    dev = d_res.device
    rec = torch.empty((n, 5), dtype=torch.float64, device=dev)
    solver.make_records(d_res.data_ptr(), d_group.data_ptr(), rank, world, n, rec.data_ptr())      # synchronises the solver's stream
    if dist.is_available() and dist.is_initialized() and world > 1:
        if host_group is not None:
            dist.barrier(group=host_group)
        allrec = torch.empty((world * n, 5), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allrec, rec)
        torch.cuda.current_stream(dev).synchronize()           # the solver's kernels run on its own stream
    else:
        allrec = rec
    win = torch.empty(n_groups, dtype=torch.int64, device=dev)
    solver.select_best(allrec.data_ptr(), allrec.shape[0], n_groups, win.data_ptr())
    return win
