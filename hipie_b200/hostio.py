"""Host-side staging for the results of HIPIE_IMG.forward: one persistent page-locked arena, so the per-step device->host reads are
plain async copies into memory that is pinned ONCE (a fresh `torch.empty(pin_memory=True)` per tensor costs a cudaHostAlloc -- an
implicit device synchronisation of a millisecond or more each -- whenever the caching host allocator has no block of that size).

The reference hands its results to the evaluator with `.to('cpu')` per tensor (detectron2 evaluators, e.g.
detectron2/evaluation/coco_evaluation.py:  instances = output["instances"].to(self._cpu_device)); this is the same hand-off with the
copies queued back to back on the current stream and ONE synchronisation at the end (`wait`)."""
import torch


class PinnedArena:
    def __init__(self, nbytes=64 << 20):
        self._buf = torch.empty(int(nbytes), dtype=torch.uint8, pin_memory=True)
        self._off = 0
        self._spill = []

    def reset(self):
        """Start a new step: every view handed out before is recycled (the caller has consumed or copied the previous results)."""
        if self._spill:        # the last step did not fit: grow once, outside the copies
            self._buf = torch.empty(int(self._off * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)      # _off counted the spilled bytes too
            self._spill = []
        self._off = 0

    def to_host(self, t: torch.Tensor) -> torch.Tensor:
        """Queue an async device->host copy of `t` on the current stream; returns the host view (valid after `wait`)."""
        t = t.contiguous()
        n = t.numel() * t.element_size()
        off = (self._off + 255) & ~255
        if off + n > self._buf.numel():
            h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)      # rare: arena too small this step; grown at the next reset
            self._spill.append(h)
            self._off = off + n
        else:
            h = self._buf[off:off + n].view(t.dtype).view(t.shape)
            self._off = off + n
        h.copy_(t, non_blocking=True)
        return h

    @staticmethod
    def wait():
        torch.cuda.current_stream().synchronize()
