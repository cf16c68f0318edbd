"""Double-buffered batch driver: two engine contexts (one per host thread, each with its own CUDA streams) take
alternate batches, so the PCIe upload and the host-side polling of one batch hide behind the kernels of the other.

This is the documented threading model of the C ABI (one b200reg_ctx per host thread, include/b200reg.h); ctypes
releases the GIL for the duration of each call, so plain Python threads are enough.
"""
import queue
import threading

from .native import Context


class PipelinedRegistrar:
    def __init__(self, device=0, depth=2):
        self.ctxs = [Context(device) for _ in range(depth)]
        self._q = [queue.Queue() for _ in range(depth)]
        self._threads = [threading.Thread(target=self._run, args=(i,), daemon=True) for i in range(depth)]
        self._rr = 0
        for t in self._threads:
            t.start()

    def _run(self, i):
        ctx = self.ctxs[i]
        while True:
            item = self._q[i].get()
            if item is None:
                return
            fn, args, fut = item
            try:
                fut["result"] = fn(ctx, *args)
            except Exception as e:  # surfaced by wait()
                fut["error"] = e
            fut["done"].set()

    def submit(self, fn, *args):
        """fn(ctx, *args) runs on the next context's thread; returns a future dict (use wait())."""
        fut = {"done": threading.Event()}
        self._q[self._rr].put((fn, args, fut))
        self._rr = (self._rr + 1) % len(self.ctxs)
        return fut

    @staticmethod
    def wait(fut):
        fut["done"].wait()
        if "error" in fut:
            raise fut["error"]
        return fut["result"]

    def icp_alignment_ptrs(self, *args, **kw):
        return self.submit(lambda ctx: ctx.icp_alignment_ptrs(*args, **kw))

    def loop_closure_ptrs(self, *args, **kw):
        """LoopClosure::coarseToFineAlignment batches; the future resolves to the Result array."""
        return self.submit(lambda ctx: ctx.loop_closure_ptrs(*args, **kw)[0])

    def synchronize(self):
        for c in self.ctxs:
            c.synchronize()

    @property
    def launch_count(self):
        return sum(c.launch_count for c in self.ctxs)

    def close(self):
        for q_ in self._q:
            q_.put(None)
        for t in self._threads:
            t.join(timeout=10)
        for c in self.ctxs:
            c.close()
