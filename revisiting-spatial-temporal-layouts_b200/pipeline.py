"""Host <-> device pipelining around the drop-in modules.

The reference inference loop (src/inference.py:72-78) is strictly sequential per batch: copy the batch to
the device, run the model, read the logits back (`evaluator.process` calls `.cpu()`). With the forward at
~27 ms for 4096 videos the two PCIe copies and the host synchronisation are a visible fraction of a
step, so this helper overlaps them with compute: inputs of step i+1 travel on a copy stream while step i
computes, and the logits of step i-1 travel back on a third stream. Semantics are unchanged: every batch
is still uploaded from (pinned) host memory, every result is delivered to host memory, in order.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, Iterator, List, Optional

import torch


class HostPipeline:
    def __init__(self, model, output_key: str, depth: int = 2, device: Optional[torch.device] = None):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one batch in flight per direction)")
        self.model = model
        self.output_key = output_key
        self.depth = depth
        self.device = device or next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the module on a CUDA device")
        self.copy_in = torch.cuda.Stream(self.device)
        self.copy_out = torch.cuda.Stream(self.device)
        self._dev: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth
        self._host_out: List[Optional[torch.Tensor]] = [None] * depth
        self._in_ready = [torch.cuda.Event() for _ in range(depth)]
        self._compute_done = [torch.cuda.Event() for _ in range(depth)]
        self._out_done = [torch.cuda.Event() for _ in range(depth)]
        self._used = [False] * depth

    def _upload(self, slot: int, host_batch: Dict[str, torch.Tensor]) -> None:
        dev = self._dev[slot]
        if dev is None or any(k not in dev or dev[k].shape != v.shape or dev[k].dtype != v.dtype
                              for k, v in host_batch.items()):
            dev = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device) for k, v in host_batch.items()}
            self._dev[slot] = dev
        with torch.cuda.stream(self.copy_in):
            if self._used[slot]:
                self.copy_in.wait_event(self._compute_done[slot])  # the previous occupant has been consumed
            for k, v in host_batch.items():
                dev[k].copy_(v, non_blocking=True)
            self._in_ready[slot].record(self.copy_in)

    def run(self, host_batches: Iterable[Dict[str, torch.Tensor]],
            consume: Optional[Callable[[int, torch.Tensor], None]] = None) -> Iterator[torch.Tensor]:
        """Feeds host batches (tensor values; pinned memory makes the copies asynchronous) through the
        model and yields the host copy of ``model(batch)[output_key]`` for every batch, in order. The
        yielded tensor is a pinned staging buffer that is reused ``depth`` batches later."""
        compute = torch.cuda.current_stream(self.device)
        pending: List[int] = []
        index = 0
        with torch.no_grad():
            for host_batch in host_batches:
                slot = index % self.depth
                if self._used[slot] and slot in pending:  # deliver the batch that still owns this slot
                    yield from self._deliver(pending, upto=slot, consume=consume)
                tensors = {k: v for k, v in host_batch.items() if isinstance(v, torch.Tensor)}
                self._upload(slot, tensors)
                compute.wait_event(self._in_ready[slot])
                out = self.model(self._dev[slot])[self.output_key]
                self._compute_done[slot].record(compute)
                host_out = self._host_out[slot]
                if host_out is None or host_out.shape != out.shape or host_out.dtype != out.dtype:
                    host_out = torch.empty(out.shape, dtype=out.dtype).pin_memory()
                    self._host_out[slot] = host_out
                with torch.cuda.stream(self.copy_out):
                    self.copy_out.wait_event(self._compute_done[slot])
                    host_out.copy_(out, non_blocking=True)
                    out.record_stream(self.copy_out)
                    self._out_done[slot].record(self.copy_out)
                self._used[slot] = True
                pending.append(slot)
                index += 1
            yield from self._deliver(pending, upto=None, consume=consume)

    def _deliver(self, pending: List[int], upto: Optional[int], consume) -> Iterator[torch.Tensor]:
        while pending:
            slot = pending.pop(0)
            self._out_done[slot].synchronize()
            if consume is not None:
                consume(slot, self._host_out[slot])
            yield self._host_out[slot]
            if upto is not None and slot == upto:
                return
