"""Step-level runtime helpers around the drop-in module (the glue of utils/train.py:30-53,168-170 that is
left once the layer is fused -- per-step host->device copies and launch overhead).

`PipelinedStep` runs a training step as a captured CUDA graph with DOUBLE-BUFFERED device inputs: while step k computes
from input set k % 2, the pinned-host inputs of step k+1 are copied into the other set on a copy stream, so the host->device
transfer (4.8 MB per Water-3D batch, ~0.15 ms over PCIe) is off the critical path the way a prefetching loader takes it off.
The C ABI never allocates or synchronises, so the whole step (graph prep, forward, losses, backward, collectives, optimizer)
is capturable; two graphs are captured, one per input set, over the same parameters and optimizer state."""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch


class PipelinedStep:
    def __init__(self, step_fn: Callable[[Dict[str, torch.Tensor]], torch.Tensor], host_inputs: Dict[str, torch.Tensor],
                 device, warmup: int = 3, use_cuda_graph: bool = True):
        """step_fn(inputs) -> scalar loss tensor, reading ONLY the tensors of `inputs` (same keys / shapes as
        host_inputs).  host_inputs: pinned host tensors of one batch (shapes are static across steps)."""
        self.step_fn, self.device = step_fn, device
        self.sets = [{k: torch.empty_like(v, device=device) for k, v in host_inputs.items()} for _ in range(2)]
        for s in self.sets:
            for k, v in host_inputs.items():
                s[k].copy_(v)
        self.copy_stream = torch.cuda.Stream(device=device)
        self.ready = [torch.cuda.Event() for _ in range(2)]       # input set i holds the batch it was last asked to load
        self.free = [torch.cuda.Event() for _ in range(2)]        # the step that read input set i has finished
        self.graphs, self.losses, self.why = [None, None], [None, None], None
        # the loss of a step lands in pinned host memory through a copy INSIDE the step's graph; `done` is recorded behind it
        self.loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self.done = [torch.cuda.Event() for _ in range(2)]
        self._next = 0
        self._primed = False
        if use_cuda_graph:
            try:
                side = torch.cuda.Stream(device=device)
                side.wait_stream(torch.cuda.current_stream(device))
                with torch.cuda.stream(side):
                    for _ in range(warmup):
                        step_fn(self.sets[0])
                torch.cuda.current_stream(device).wait_stream(side)
                torch.cuda.synchronize(device)
                for i in range(2):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self.losses[i] = step_fn(self.sets[i])
                        self.loss_host[i:i + 1].copy_(self.losses[i].detach().reshape(1), non_blocking=True)
                    self.graphs[i] = g
            except Exception as exc:            # reported by the caller; the eager path below still works
                self.graphs, self.why = [None, None], f"{type(exc).__name__}: {exc}"[:300]
                torch.cuda.synchronize(device)
        cur = torch.cuda.current_stream(device)
        for i in range(2):
            self.free[i].record(cur)
            self.ready[i].record(cur)

    def prefetch(self, host_inputs: Dict[str, torch.Tensor], slot: Optional[int] = None) -> None:
        """Start copying a batch into input set `slot` (default: the set the NEXT run() will read) on the copy stream."""
        i = self._next if slot is None else slot
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[i])             # the step that last read this set is done with it
            for k, v in host_inputs.items():
                self.sets[i][k].copy_(v, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self._primed = True

    def run(self, next_host_inputs: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
        """Run one step on the batch last prefetched into the current set; if next_host_inputs is given, its copy into the
        other set is started first and overlaps this step.  Returns the (device) loss of this step."""
        i = self._next
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[i])
        if self.graphs[i] is not None:
            self.graphs[i].replay()
            loss = self.losses[i]
        else:
            loss = self.step_fn(self.sets[i])
            self.loss_host[i:i + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.free[i].record(cur)
        self.done[i].record(cur)
        self._last = i
        # the next batch's copies are SUBMITTED after this step's launch: the host time of a dozen cudaMemcpyAsync calls
        # (30-40 us) then passes while the GPU is already computing instead of in front of the step.  The copies only wait
        # for the step that last read the other input set -- the previous one, whose `free` event is long recorded.
        if next_host_inputs is not None:
            self.prefetch(next_host_inputs, slot=i ^ 1)
        self._next = i ^ 1
        return loss

    def run_and_read(self, next_host_inputs: Optional[Dict[str, torch.Tensor]] = None) -> float:
        """run() + the device -> host read of the step's loss (utils/train.py:172-173 accumulates loss.item() every batch).
        The value travels by a copy node at the end of the step's graph into pinned host memory and the host waits on the
        event behind it: no separate cudaMemcpy + stream synchronise round trip after the step (25-30 us with `.item()`)."""
        self.run(next_host_inputs)
        self.done[self._last].synchronize()
        return float(self.loss_host[self._last])

