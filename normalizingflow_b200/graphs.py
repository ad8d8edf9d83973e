"""CUDA-graph capture of whole steps: small-batch work on this path is bound by
launch and Python-dispatch latency (a 2-D RealNVP training step is ~600 kernel launches of a few
microseconds each), so a step whose launch sequence is fixed — same shapes, same layers — is
captured once and replayed as ONE graph launch.

``GraphedCallable`` wraps an inference call (static input buffers -> static outputs);
``GraphedTrainStep`` wraps ``loss = loss_fn(*inputs); loss.backward(); optimizer.step()`` including
the optimizer (the optimizer must be constructed with ``capturable=True``).  Every libnfk kernel
launches on the capturing stream and never allocates or synchronises, so the C-ABI needs nothing
special; weight images of the tensor-core paths are re-packed inside the graph after each step."""
from __future__ import annotations

import torch

from . import _lib


def _warmup(fn, n=3):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(n):
            fn()
    torch.cuda.current_stream().wait_stream(s)


class GraphedCallable:
    """fn(*static_inputs) -> tensor or tuple of tensors, replayed as one CUDA graph.  Call with new
    inputs of the SAME shapes; the returned tensors are static buffers (clone to keep them)."""

    def __init__(self, fn, *example_inputs):
        self.fn = fn
        self.static_in = [t.clone() for t in example_inputs]
        with torch.no_grad():
            _warmup(lambda: fn(*self.static_in))
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t)
        self.graph.replay()
        return self.static_out


class GraphedTrainStep:
    """One optimisation step as a CUDA graph: ``loss_fn(*inputs)`` -> scalar loss, backward, optimizer
    step (applications/src/train.py:22-29).  ``inputs`` may be empty (e.g. reverse-KL losses that draw
    their own latents: the generator state is graph-safe)."""

    def __init__(self, loss_fn, optimizer, *example_inputs, warmup=3):
        self.loss_fn, self.opt = loss_fn, optimizer
        self.static_in = [t.clone() for t in example_inputs]

        def step():
            optimizer.zero_grad(set_to_none=True)
            loss = loss_fn(*self.static_in)
            loss.backward()
            optimizer.step()
            return loss

        _warmup(step, warmup)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            loss = loss_fn(*self.static_in)
            loss.backward()
            optimizer.step()
            self.static_loss = loss.detach()

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            s.copy_(t)
        self.graph.replay()
        # the replayed optimizer step rewrote the parameters without touching their version counters:
        # weight images / chunk graphs cached outside this graph are stale from here on
        _lib.invalidate_caches()
        return self.static_loss
