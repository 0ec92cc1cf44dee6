"""Per-stage CUDA-event timing hooks (used by bench.py; zero cost when inactive).

Events are recorded on torch's current stream, which is the stream every kernel of this package is
launched on (the wrappers pass `torch.cuda.current_stream().cuda_stream` through the C ABI)."""
from __future__ import annotations

from collections import defaultdict
from contextlib import contextmanager, nullcontext

import torch

import os

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


NVTX = False


def enable_nvtx(on: bool = True):
    """Wrap every stage of the path (visible_filter, decode_fwd, preprocess_fwd, binning, blend_fwd, blend_bwd, ...)
    in an NVTX range so Nsight timelines show the stages; also switched on by SPLATCO_NVTX=1."""
    global NVTX
    NVTX = bool(on)


@contextmanager
def _nvtx_stage(name):
    torch.cuda.nvtx.range_push("splatco/" + name)
    try:
        if ACTIVE is not None:
            with ACTIVE.stage(name):
                yield
        else:
            yield
    finally:
        torch.cuda.nvtx.range_pop()


def stage(name):
    if NVTX:
        return _nvtx_stage(name)
    return ACTIVE.stage(name) if ACTIVE is not None else nullcontext()


@contextmanager
def collect():
    global ACTIVE
    prev, ACTIVE = ACTIVE, StageTimer()
    try:
        yield ACTIVE
    finally:
        ACTIVE = prev


if os.environ.get("SPLATCO_NVTX") == "1":
    NVTX = True
