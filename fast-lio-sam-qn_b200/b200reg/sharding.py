"""Batch sharding over ranks: keyframe pairs are independent registrations, so the
batch is partitioned across GPUs with NO data-path collective; the only exchange is ONE all-gather of
the fixed-size result records at the end (4x4 transform + score + flags).

Host-side logic only; tested on CPU with the gloo backend (tests/test_sharding.py).
"""
import numpy as np

RECORD_DOUBLES = 24  # T(16) fitness converged valid iterations n_linearize n_error lm_failed pair_index


def shard_pairs(n_pairs, world, rank, costs=None):
    """Indices of the pairs rank `rank` processes.

    costs=None -> contiguous static blocks (64/GPU for 512 pairs on 8 GPUs).
    costs given (e.g. N+M per pair) -> greedy longest-processing-time assignment, deterministic.
    Every rank gets ceil(n/world) slots at most so the gathered buffer has a fixed shape.
    """
    if costs is None:
        per = (n_pairs + world - 1) // world
        lo = min(rank * per, n_pairs)
        return list(range(lo, min(lo + per, n_pairs)))
    costs = np.asarray(costs, dtype=np.float64)
    order = sorted(range(n_pairs), key=lambda i: (-costs[i], i))
    cap = (n_pairs + world - 1) // world
    loads = [0.0] * world
    bins = [[] for _ in range(world)]
    for i in order:
        r = min((r for r in range(world) if len(bins[r]) < cap), key=lambda r: (loads[r], r))
        bins[r].append(i)
        loads[r] += costs[i]
    return sorted(bins[rank])


def pack_records(indices, results, slots):
    """results: list of dicts (native.Result.as_dict()) -> (slots, RECORD_DOUBLES) float64, padded with -1 index."""
    rec = np.zeros((slots, RECORD_DOUBLES), np.float64)
    rec[:, 23] = -1
    for s, (i, r) in enumerate(zip(indices, results)):
        rec[s, :16] = np.asarray(r["T"], np.float64).reshape(16)
        rec[s, 16] = r["fitness"]
        rec[s, 17] = float(r["converged"])
        rec[s, 18] = float(r["valid"])
        rec[s, 19] = r["iterations"]
        rec[s, 20] = r["n_linearize"]
        rec[s, 21] = r["n_error"]
        rec[s, 22] = float(r["lm_failed"])
        rec[s, 23] = i
    return rec


def unpack_records(all_rec, n_pairs):
    """(world*slots, RECORD_DOUBLES) gathered buffer -> list of per-pair dicts in pair order."""
    out = [None] * n_pairs
    for row in np.asarray(all_rec).reshape(-1, RECORD_DOUBLES):
        i = int(row[23])
        if i < 0:
            continue
        out[i] = dict(T=row[:16].reshape(4, 4).copy(), fitness=float(row[16]), converged=bool(row[17]),
                      valid=bool(row[18]), iterations=int(row[19]), n_linearize=int(row[20]), n_error=int(row[21]),
                      lm_failed=bool(row[22]))
    return out


def register_sharded(register_fn, srcs, dsts, dist=None, device="cpu", costs=None):
    """Run register_fn(list_of_src, list_of_dst) -> list of result dicts on this rank's shard and all-gather.

    dist: an initialised torch.distributed module (nccl on GPUs, gloo in the CPU tests) or None (single rank).
    Returns the full per-pair result list on every rank; the gathered bytes are identical for any world size.
    """
    import torch
    n = len(srcs)
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    mine = shard_pairs(n, world, rank, costs)
    slots = (n + world - 1) // world
    res = register_fn([srcs[i] for i in mine], [dsts[i] for i in mine]) if mine else []
    rec = torch.from_numpy(pack_records(mine, res, slots)).to(device)
    if dist is None or world == 1:
        return unpack_records(rec.cpu().numpy(), n)
    allrec = torch.empty((world * slots, RECORD_DOUBLES), dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(allrec, rec)
    return unpack_records(allrec.cpu().numpy(), n)


def register_sharded_native(ctx, register_raw_fn, srcs, dsts, costs=None):
    """The same sharded batch with the collective behind the C ABI: every rank registers its shard, tags each
    b200reg_result with its global pair index, and ONE b200reg_allgather_results (ncclAllGather of the full records on
    the context's communicator, or a copy when the context has none) leaves all results on every rank.

    register_raw_fn(list_of_src, list_of_dst) -> ctypes (Result * n) array.  Returns native.Result records in pair order.
    """
    from .native import Result, lib
    n = len(srcs)
    world = ctx.comm_world
    rank = max(0, int(lib().b200reg_comm_rank(ctx.h)))
    mine = shard_pairs(n, world, rank, costs)
    slots = (n + world - 1) // world
    local = (Result * slots)()
    for s in range(slots):
        local[s].status = -100  # empty slot of a ragged last shard
        local[s].tag = -1
    if mine:
        res = register_raw_fn([srcs[i] for i in mine], [dsts[i] for i in mine])
        for s, (i, r) in enumerate(zip(mine, res)):
            local[s] = r
            local[s].tag = i
    allr = ctx.allgather_results(local)
    out = [None] * n
    for r in allr:
        if r.tag >= 0 and r.status != -100:
            out[r.tag] = r
    return out
