"""Host <-> device streaming around the trunk forward: `HostPipeline(model).run(batches)`.

The reference's callers (DNAModel.predict / ism / score_variant, dna_model.py:391-673) feed one host batch after
another through `model(dna, organism_index)` and read the embeddings back.  Doing that naively serialises three things
per step: the H2D copy of the tokens, the forward, and the D2H read-back (235 MB of embeddings at 1 Mbp = ~5 ms of
PCIe).  This helper overlaps them with two CUDA streams and double buffering:

    compute stream:  forward(i) (CUDA-graph replay) -> D2D of the requested outputs into staging[i % 2]
    copy stream   :  H2D tokens(i+1)   ...   D2H staging[i % 2] -> pinned host[i % 2]        (while forward(i) / (i+1) run)

Nothing here touches the arithmetic; every step still copies its own inputs from pinned host memory and its own
results to pinned host memory — the copies just do not sit on the forward's critical path.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence, Tuple

import torch


class HostPipeline:
    def __init__(self, model, outputs: Sequence[str] = ("embeds_128bp", "embeds_pair"), depth: int = 2):
        self.model = model
        self.outputs = tuple(outputs)
        self.depth = depth
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._stage = None
        self._host = None
        self._done = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _ensure(self, emb, dev):
        if self._stage is not None:
            return
        self._stage, self._host, self._done = [], [], []
        for _ in range(self.depth):
            srcs = [getattr(emb, n) for n in self.outputs]
            self._stage.append([torch.empty_like(s) for s in srcs])
            self._host.append([torch.empty(s.shape, dtype=s.dtype, pin_memory=True) for s in srcs])
            self._done.append(torch.cuda.Event())

    def run(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], consume: Optional[Callable] = None):
        """batches: iterable of (tokens int64 [B, L] pinned host, organism_index int64 [B] pinned host).
        consume(i, host_outputs) is called once slot i's read-back has completed (outputs are pinned host tensors
        that the pipeline reuses `depth` steps later).  Returns the number of steps run."""
        main = torch.cuda.current_stream()
        dev = main.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        cs = self._copy_stream
        pending = []  # (step, slot)
        n = 0
        it = iter(batches)

        def upload(batch):
            """H2D of one batch on the copy stream (overlaps the forward running on the compute stream)."""
            tok, org = batch
            with torch.cuda.stream(cs):
                t = tok.to(dev, non_blocking=True)
                o = org.to(dev, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cs)
            self.h2d_bytes += tok.numel() * tok.element_size() + org.numel() * org.element_size()
            return t, o, ev

        nxt = next(it, None)
        staged = upload(nxt) if nxt is not None else None
        i = 0
        while staged is not None:
            slot = i % self.depth
            t, o, ev = staged
            main.wait_event(ev)
            t.record_stream(main)
            o.record_stream(main)
            emb = self.model(t, o)
            nxt = next(it, None)
            staged = upload(nxt) if nxt is not None else None     # the next batch's tokens travel while this forward runs
            self._ensure(emb, dev)
            if len(pending) >= self.depth:       # this slot's previous read-back must be finished before it is reused
                j, s_ = pending.pop(0)
                self._done[s_].synchronize()
                if consume is not None:
                    consume(j, self._host[s_])
            for dst, name in zip(self._stage[slot], self.outputs):
                dst.copy_(getattr(emb, name), non_blocking=True)      # D2D on the compute stream (outputs are graph-static)
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(cs):
                cs.wait_event(ready)
                for h, d in zip(self._host[slot], self._stage[slot]):
                    h.copy_(d, non_blocking=True)
                    self.d2h_bytes += d.numel() * d.element_size()
                self._done[slot].record(cs)
            pending.append((i, slot))
            n += 1
            i += 1
        for j, s_ in pending:
            self._done[s_].synchronize()
            if consume is not None:
                consume(j, self._host[s_])
        # the caller's timing events live on the compute stream: make it wait for the last read-backs
        if self._done is not None:
            for ev in self._done:
                main.wait_event(ev)
        return n
