"""Why does the headline not double at N=2?  torchrun --nproc-per-node 2 profiles/n2_diag.py
Runs bench.py's own Runner over the headline jobs with the per-step gather on and off, prints per-RANK region times
(bench.py reports the max) and the host-side duration of every b200reg_allgather_results call."""
import ctypes
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "fast-lio-sam-qn_b200"))
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import b200reg
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dist.all_reduce(torch.zeros(8, device="cuda"))
    ctx = b200reg.Context(local_rank)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    box = [b200reg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)
    depth = int(os.environ.get("DEPTH", "3"))
    batch = b200reg.Batch(local_rank, depth=depth)
    runner = bench.Runner(batch, ctx, dist, stream, depth)
    prm = b200reg.default_params()
    pairs = bench.gen_pairs(bench.primary_seeds(0 if os.environ.get("SAME_SEEDS") else rank), 100000)
    arena = bench.Arena(pairs, bench.JOB_PAIRS)
    gather_ms = []
    real = ctx.allgather_results

    def timed_gather(local):
        t0 = time.perf_counter()
        out = real(local)
        gather_ms.append(1e3 * (time.perf_counter() - t0))
        return out
    ctx.allgather_results = timed_gather
    sub = lambda j: batch.submit_icp(*arena.job(j, True), prm)
    report = {}
    for name, gather, solo in (("warm", True, -1), ("gather", True, -1), ("no_gather", False, -1), ("gather2", True, -1),
                               ("rank0_alone", False, 0), ("rank1_alone", False, 1)):
        gather_ms.clear()
        n_jobs = 10 * bench.JOBS_PER_STEP
        if solo >= 0:  # one rank works, the other idles: is the slowdown cross-process interference on the host / the box?
            dist.barrier()
            t0 = time.perf_counter()
            if rank == solo:
                r2 = bench.Runner(batch, ctx, None, stream, depth)
                r2.run(n_jobs, sub, bench.JOBS_PER_STEP, gather=False)
            own = 1e3 * (time.perf_counter() - t0)
            dist.barrier()
        else:
            t0 = time.perf_counter()
            runner.run(n_jobs, sub, bench.JOBS_PER_STEP, gather=gather)
            own = 1e3 * (time.perf_counter() - t0)
        rec = dict(rank=rank, wall_ms=own, pairs_per_s_this_rank=bench.JOB_PAIRS * n_jobs / (own * 1e-3),
                   gather_ms=[round(g, 2) for g in gather_ms])
        allr = [None] * world
        dist.all_gather_object(allr, rec)
        report[name] = allr
    if rank == 0:
        print(json.dumps(report, indent=1))
    batch.close()
    os._exit(0)


if __name__ == "__main__":
    main()
