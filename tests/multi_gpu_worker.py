"""torchrun worker for tests/test_gpu_multi.py: shard a batch of pairs over the ranks (NCCL all-gather of the result
records) and check on every rank that the gathered result equals the single-rank result bit for bit."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "fast-lio-sam-qn_b200"))


def main():
    import torch
    import torch.distributed as dist
    import b200reg
    from b200reg import synth
    from b200reg.sharding import register_sharded
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_pairs = 7  # odd: the last rank gets a padded slot
    pairs = [synth.make_pair(3000 + i, 5000 + 700 * i, 6000 - 200 * i) for i in range(n_pairs)]
    srcs = [p[0] for p in pairs]
    dsts = [p[1] for p in pairs]
    ctx = b200reg.Context(local)

    def reg(s, d):
        return ctx.icp_alignment(s, d)
    costs = [len(s) + len(d) for s, d in zip(srcs, dsts)]
    sharded = register_sharded(reg, srcs, dsts, dist=dist, device="cuda")
    balanced = register_sharded(reg, srcs, dsts, dist=dist, device="cuda", costs=costs)
    single = register_sharded(reg, srcs, dsts)  # every rank also computes the whole batch alone
    for a, b, c in zip(sharded, balanced, single):
        assert a["T"].tobytes() == c["T"].tobytes() == b["T"].tobytes(), "gathered bytes must not depend on the world size"
        assert a["fitness"] == c["fitness"] and a["converged"] == c["converged"] and a["n_linearize"] == c["n_linearize"]
    for p, r in zip(pairs, single):
        rot, tr = synth.se3_error(r["T"], p[2])
        assert r["converged"] and rot < 2e-2 and tr < 0.5
    # the same batch with the collective BEHIND the C ABI (b200reg_comm_init + b200reg_allgather_results = ncclAllGather of
    # the full result records); the 128-byte unique id travels over the process group that is already up
    from b200reg.sharding import register_sharded_native
    box = [b200reg.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, int(os.environ["WORLD_SIZE"]))
    native = register_sharded_native(ctx, lambda s, d: ctx.icp_alignment(s, d, raw=True), srcs, dsts, costs=costs)
    for a, c in zip(native, single):
        assert np.array(a.T).tobytes() == c["T"].tobytes() and a.fitness == c["fitness"] and a.n_linearize == c["n_linearize"]
        assert bool(a.valid) == c["valid"]
    ctx.comm_destroy()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MULTI_GPU_OK world=%d" % int(os.environ["WORLD_SIZE"]))


if __name__ == "__main__":
    main()
