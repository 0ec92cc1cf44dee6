"""Gradient buffers shared by the decode nodes of ONE backward pass.

The reference sums the losses of the `--mv` views and calls backward once (train.py:199,240); its
autograd then materialises, per view, N-row gradients for `_anchor_feat/_anchor/_offset`, plane-sized
gradients and ~50 small weight gradients, and adds them pairwise (SURVEY.md §7 "dense-grad hidden
cost").  `splatco_decode_bwd` accumulates (+=) into its destinations, so all decode nodes of a
backward pass that were fed the SAME input tensor can write into ONE zero-filled buffer: the first
node that runs hands the buffer to autograd, the later ones add into it in place and return None.

Why this is safe: a consumer of that gradient (AccumulateGrad of a leaf, or the grad_fn of a non-leaf
such as the cached TriPlaneAttention output) has one incoming edge per decode node that used the
tensor, and the engine runs it only after all of them have run — i.e. after the last in-place add.
Buffers are keyed on the identity of the forward input objects (kept alive by the nodes), inputs that
do not require grad are never shared, and the table is dropped by a callback the engine runs at the
end of the backward pass (`queue_callback`), so nothing leaks into the next pass.
"""
from __future__ import annotations

import threading

import torch
from torch.autograd import Variable

_lock = threading.Lock()
_passes = {}          # device index -> {key: (flat buffer, offset in floats)}
_ALIGN = 64           # floats (256 bytes): every carved buffer keeps the alignment of a fresh allocation


_last = {}            # device index -> flat buffers of the most recent finished backward pass
_arena = {}           # (device index, leaf?) -> {cur: flat being carved, used: floats taken from it, total: floats carved this pass}
_arena_size = {}      # (device index, leaf?) -> floats the previous pass carved (sizes the next pass's single allocation)
# Two arenas per pass: gradients of LEAF tensors (parameters: what a multi-GPU caller all-reduces) and gradients of
# intermediate tensors (the cached TriPlaneAttention outputs, the channel-last plane copies), which autograd consumes
# inside the pass and nobody needs afterwards.


_pools = {}           # device index -> torch.cuda.MemPool the arenas are allocated from


def _zeros(numel, device):
    """The arena allocation.  On CUDA it comes from a PRIVATE memory pool: a multi-GPU caller all-reduces the leaf arena on
    NCCL's stream, so after `p.grad = None` its block stays unavailable until that stream's event has passed; the next
    pass's arena then took the best-fitting free block of the shared pool -- the decode backward's workspace -- whose owner
    had to cudaMalloc a new one (10-100 ms with peer access enabled, a stall every rank shares through the all-reduce;
    found with the allocator's history: every new segment came from that one line).  In their own pool the arenas can only
    take each other's blocks and the pool stops growing after a couple of passes."""
    if device.type != "cuda" or not hasattr(torch.cuda, "MemPool"):
        return torch.zeros(numel, dtype=torch.float32, device=device)
    index = device.index if device.index is not None else torch.cuda.current_device()
    pool = _pools.get(index)
    if pool is None:
        pool = _pools[index] = torch.cuda.MemPool()
    with torch.cuda.use_mem_pool(pool, device=device):
        return torch.zeros(numel, dtype=torch.float32, device=device)


def _end_of_pass(index):
    with _lock:
        table = _passes.pop(index, None)
        for leaf in (True, False):
            arena = _arena.pop((index, leaf), None)
            if arena is not None:
                _arena_size[(index, leaf)] = arena["total"]
        if table is not None:
            seen, flats = set(), []
            for flat, _, _, leaf in table.values():
                if leaf and id(flat) not in seen:
                    seen.add(id(flat))
                    flats.append(flat)
            _last[index] = flats


def last_pass_buffers(device):
    """Flat fp32 buffers that hold the shared LEAF gradients of the most recent backward pass on `device` (the
    parameters' .grad are views of them; gradients of intermediate tensors live elsewhere).  Same sizes on every rank of a replicated model, so a multi-GPU caller
    can all-reduce THEM in place instead of packing each .grad into a bucket (multiview.GradBucket)."""
    device = torch.device(device)
    index = device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else -1)
    with _lock:
        return list(_last.get(index, ()))


def acquire(device, requests, want_views=False):
    """requests: list of (key or None, shape) or (key or None, shape, is_leaf) -- is_leaf (default True) says whether the
    gradient belongs to a leaf tensor (a parameter) or to an intermediate one.  Returns a list of (pointer, ret): the address to
    accumulate into and what to return to autograd for that input (None when an earlier node of this
    backward pass already returned the buffer; do not keep `ret`: autograd adopts it as .grad without a
    copy only while nobody else holds it).  key None = private buffer, never shared.  With
    want_views=True the entries are (pointer, ret, buffer-as-tensor) (host-logic tests)."""
    device = torch.device(device)
    index = device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else -1)
    with _lock:
        table = _passes.get(index)
        fresh_pass = table is None
        if fresh_pass:
            table = _passes[index] = {}
    if fresh_pass:
        try:
            Variable._execution_engine.queue_callback(lambda: _end_of_pass(index))
        except RuntimeError:
            # not inside an engine run (direct call of backward in a test): nothing may be shared
            _end_of_pass(index)
            table = {}
    out = [None] * len(requests)
    pending = {True: [], False: []}
    for n, req in enumerate(requests):
        key, shape = req[0], req[1]
        leaf = bool(req[2]) if len(req) > 2 else True
        hit = table.get(key) if key is not None else None
        if hit is not None:
            flat, off, numel, _ = hit
            out[n] = (flat.data_ptr() + 4 * off, None, flat[off:off + numel].view(shape)) if want_views else \
                     (flat.data_ptr() + 4 * off, None)
        else:
            numel = 1
            for s in shape:
                numel *= int(s)
            lst = pending[leaf]
            off = lst[-1][3] + (lst[-1][4] + _ALIGN - 1) // _ALIGN * _ALIGN if lst else 0
            lst.append((n, key, shape, off, numel))
    for leaf, todo in pending.items():
        if not todo:
            continue
        total = todo[-1][3] + (todo[-1][4] + _ALIGN - 1) // _ALIGN * _ALIGN
        akey = (index, leaf)
        _carve(table, out, todo, total, akey, index in _passes, device, want_views, leaf)
    return out


def _carve(table, out, todo, total, akey, in_pass, device, want_views, leaf):
    if True:
        # ONE zero-filled allocation per backward pass (sized from what the previous pass carved): the decode nodes, the
        # TriPlaneAttention backward and the plane unpacking all carve from it, so a multi-GPU caller all-reduces one
        # buffer with one collective (multiview.GradBucket) and the pass pays one fill kernel
        total = max(total, 1)
        arena = _arena.get(akey) if in_pass else None
        if arena is not None and arena["used"] + total <= arena["cur"].numel():
            flat, start = arena["cur"], arena["used"]
        else:
            want = max(total, _arena_size.get(akey, 0)) if (in_pass and arena is None) else total
            flat, start = _zeros(want, device), 0
            if in_pass:
                if arena is None:
                    arena = _arena[akey] = {"cur": flat, "used": 0, "total": 0}
                arena["cur"], arena["used"] = flat, 0
        if arena is not None:
            arena["used"] = start + total
            arena["total"] += total
        base = flat.data_ptr() + 4 * start
        # all views with one split call (a slice + view pair per tensor costs ~5 us, there are ~50 of them)
        bounds = [off for _, _, _, off, _ in todo] + [total]
        pieces = flat[start:start + total].split_with_sizes([b - a for a, b in zip(bounds[:-1], bounds[1:])])
        for (n, key, shape, off, numel), piece in zip(todo, pieces):
            buf = (piece if piece.numel() == numel else piece[:numel]).view(shape)
            out[n] = (base + 4 * off, buf, buf) if want_views else (base + 4 * off, buf)
            if key is not None:
                table[key] = (flat, start + off, numel, leaf)
