"""Gradient synchronisation for data-parallel training of the decoder (replaces the DDP reducer that the
reference gets implicitly from Lightning's ``strategy="ddp"``, train.py:99-100).

The decoder shards by utterance batch; the only exchange is the all-reduce of the parameter gradients.  The
library finalises gradients segment by segment (last layers first, ``vcd_num_backward_segments``); each finished
segment is a contiguous slice of one flat fp32 buffer and is all-reduced asynchronously (NCCL over NVLink on
its own stream) while the remaining backward kernels run.  Backend-agnostic (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


class SegmentReducer:
    """Issues one asynchronous all-reduce per finished gradient segment; ``finish()`` waits and averages."""

    def __init__(self, flat: torch.Tensor, ranges: Sequence[Tuple[int, int]], group: Optional[dist.ProcessGroup],
                 prescaled: bool = False):
        """``prescaled``: the gradients were already multiplied by 1/world where they were produced
        (``vcd_set_gradient_scale``), so the SUM all-reduce is the average and ``finish`` only waits."""
        self.flat = flat
        self.prescaled = prescaled
        self.ranges = list(ranges)
        self.group = group
        self.world = dist.get_world_size(group) if group is not None else 1
        self.works: List = []
        self.issued = 0

    def segment_done(self, seg: int) -> None:
        """Gradients of segment ``seg`` are final (enqueued on the current stream)."""
        if self.group is None or self.world == 1:
            return
        assert seg == self.issued, "segments must complete in order"
        lo, hi = self.ranges[seg]
        self.issued += 1
        if hi > lo:
            # SUM + scale instead of AVG: AVG does not exist on every backend (gloo)
            self.works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        if self.group is None or self.world == 1:
            return
        for w in self.works:
            w.wait()
        self.works = []
        if not self.prescaled:
            self.flat.mul_(1.0 / self.world)
