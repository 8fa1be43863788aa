"""Per-shape CUDA-graph replay for the stream-ordered forwards of this package.

Every forward behind include/st2_b200.h is a fixed sequence of launches on the caller's stream that allocates nothing and
never synchronises, so it can be captured once per (shape, precision) and replayed as one graph launch: what matters for
one-sentence latency, where a forward is tens to hundreds of kernels of a few microseconds each.  Inputs are copied into
the graph's static buffers, outputs are clones of its static outputs; each graph owns its workspace.
"""
from __future__ import annotations

import collections
from typing import Callable, Tuple

import torch


class GraphReplay:
    """LRU cache of captured graphs: every entry owns a private workspace and static input / output buffers, so a server that
    sees many shapes must not keep them all -- the least recently used graph is dropped beyond `max_graphs`."""

    def __init__(self, max_graphs: int = 8) -> None:
        self._cache: "collections.OrderedDict[tuple, dict]" = collections.OrderedDict()
        self.max_graphs = max_graphs

    def clear(self) -> None:
        """Drop every captured graph (the packed weights they point to are about to be re-allocated)."""
        self._cache.clear()

    def __len__(self) -> int:
        return len(self._cache)

    def run(self, key: tuple, inputs: Tuple[torch.Tensor, ...], make_workspace: Callable[[], torch.Tensor],
            launch: Callable[[Tuple[torch.Tensor, ...], torch.Tensor], Tuple[torch.Tensor, ...]]) -> Tuple[torch.Tensor, ...]:
        """`launch(static_inputs, workspace)` must only enqueue work on the current stream and return its output tensors."""
        dev = inputs[0].device
        g = self._cache.get(key)
        if g is not None:
            self._cache.move_to_end(key)
        if g is None:
            while len(self._cache) >= max(1, self.max_graphs):
                self._cache.popitem(last=False)
            static_in = tuple(torch.empty_like(t) for t in inputs)
            for a, b in zip(static_in, inputs):
                a.copy_(b)
            ws = make_workspace()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                launch(static_in, ws)                    # eager warm-up: function attributes, tensor-map driver entry point
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                outs = launch(static_in, ws)
            g = {"in": static_in, "ws": ws, "graph": graph, "out": tuple(outs)}
            self._cache[key] = g
        for a, b in zip(g["in"], inputs):
            a.copy_(b, non_blocking=True)
        g["graph"].replay()
        return tuple(o.clone() for o in g["out"])
