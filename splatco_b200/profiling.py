"""Per-stage CUDA-event timing hooks (used by bench.py; zero cost when inactive).

Events are recorded on torch's current stream, which is the stream every kernel of this package is
launched on (the wrappers pass `torch.cuda.current_stream().cuda_stream` through the C ABI)."""
from __future__ import annotations

from collections import defaultdict
from contextlib import contextmanager, nullcontext

import torch

ACTIVE = None


class StageTimer:
    def __init__(self):
        self.events = defaultdict(list)

    @contextmanager
    def stage(self, name):
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        try:
            yield
        finally:
            e.record()
            self.events[name].append((s, e))

    def summary(self):
        """name -> (calls, total_ms); call after torch.cuda.synchronize()."""
        return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in self.events.items()}


def stage(name):
    return ACTIVE.stage(name) if ACTIVE is not None else nullcontext()


@contextmanager
def collect():
    global ACTIVE
    prev, ACTIVE = ACTIVE, StageTimer()
    try:
        yield ACTIVE
    finally:
        ACTIVE = prev
