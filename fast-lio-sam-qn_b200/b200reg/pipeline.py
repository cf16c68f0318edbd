"""Pipelined batch registration = the C ABI's batch driver (b200reg_batch_*, csrc/batch.cu): `depth` engine contexts on
their own C++ host threads take alternate jobs, so the PCIe upload and the LM polling of one job hide behind the kernels
of the others.  This module only keeps the round-1 Python names alive on top of it; there are no Python threads here.
"""
from .native import Batch


class PipelinedRegistrar:
    def __init__(self, device=0, depth=3):
        self.batch = Batch(device, depth)

    def icp_alignment_ptrs(self, *args, **kw):
        return self.batch.submit_icp(*args, **kw)

    def loop_closure_ptrs(self, *args, **kw):
        """LoopClosure::coarseToFineAlignment batches; wait() resolves to the Result array."""
        return self.batch.submit_loop_closure(*args, **kw)

    def wait(self, ticket):
        return self.batch.wait(ticket)

    def synchronize(self):
        pass  # a waited job has synchronised its context's stream

    @property
    def launch_count(self):
        return self.batch.launch_count

    def close(self):
        self.batch.close()
