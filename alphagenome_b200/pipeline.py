"""Host <-> device streaming around the forward: `HostPipeline(model).run(batches)`.

The reference's callers (DNAModel.predict / ism / score_variant, dna_model.py:391-673) feed one host batch after
another through `model(dna, organism_index)` and read the results back.  Doing that naively serialises three things
per step: the H2D copy of the tokens, the forward, and the D2H read-back (6.7 GB for the three embeddings of a 1 Mbp
context = ~120 ms of PCIe, longer than the forward itself).  This helper overlaps them with two CUDA streams and
double buffering:

    compute stream:  forward(i) (CUDA-graph replay) -> D2D of the outputs into staging[i % 2]
    copy stream   :  H2D tokens(i+1)   ...   D2H staging[i % 2] -> pinned host[i % 2]        (while forward(i) / (i+1) run)

Nothing here touches the arithmetic; every step still copies its own inputs from pinned host memory and ALL of its
results (by default everything the call returns) to pinned host memory — the copies just do not sit on the forward's
critical path.  When the read-back is longer than the forward the pipeline runs at PCIe speed.
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple, Union

import torch


def flatten_result(res, prefix: str = "") -> List[Tuple[str, torch.Tensor]]:
    """Every tensor a forward returned, with a path name: Embeds -> embeds_1bp, ...; {organism: {head: tensor | dict}}
    -> 'human/1bp_tracks', 'human/splice_sites_junction/predictions', ..."""
    if torch.is_tensor(res):
        return [(prefix.rstrip("/"), res)]
    if hasattr(res, "_fields"):
        return [x for f in res._fields for x in flatten_result(getattr(res, f), prefix + f + "/")]
    if isinstance(res, dict):
        return [x for k, v in res.items() for x in flatten_result(v, f"{prefix}{k}/")]
    if isinstance(res, (list, tuple)):
        return [x for i, v in enumerate(res) for x in flatten_result(v, f"{prefix}{i}/")]
    return []


class HostPipeline:
    def __init__(self, model, outputs: Union[None, Sequence[str], Callable] = None, depth: int = 2, forward_kwargs: Optional[Dict] = None):
        """outputs: None = every tensor the call returns; a sequence of names (as flatten_result names them) = only
        those; or a callable(result) -> [(name, tensor)].  forward_kwargs: device-resident extra arguments of
        model(tokens, organism_index, **forward_kwargs) (e.g. splice-site indices)."""
        self.model = model
        self.outputs = outputs
        self.depth = depth
        self.forward_kwargs = dict(forward_kwargs or {})
        self._copy_stream: Optional[torch.cuda.Stream] = None
        self._stage = None
        self._host = None
        self._done = None
        self.names: List[str] = []
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _select(self, res) -> List[Tuple[str, torch.Tensor]]:
        if callable(self.outputs):
            return list(self.outputs(res))
        flat = flatten_result(res)
        if self.outputs is None:
            return flat
        by = dict(flat)
        return [(n, by[n]) for n in self.outputs]

    def _ensure(self, picked):
        if self._stage is not None:
            return
        self.names = [n for n, _ in picked]
        self._stage, self._host, self._done = [], [], []
        for _ in range(self.depth):
            self._stage.append([torch.empty(s.shape, dtype=s.dtype, device=s.device) for _, s in picked])
            self._host.append([torch.empty(s.shape, dtype=s.dtype, pin_memory=True) for _, s in picked])
            self._done.append(torch.cuda.Event())

    def run(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]], consume: Optional[Callable] = None):
        """batches: iterable of (tokens int64 [B, L] pinned host, organism_index int64 [B] pinned host).
        consume(i, host_outputs) is called once slot i's read-back has completed (outputs are pinned host tensors, in
        the order of `self.names`, that the pipeline reuses `depth` steps later).  Returns the number of steps run."""
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
            res = self.model(t, o, **self.forward_kwargs)
            nxt = next(it, None)
            staged = upload(nxt) if nxt is not None else None     # the next batch's tokens travel while this forward runs
            picked = self._select(res)
            self._ensure(picked)
            if len(pending) >= self.depth:       # this slot's previous read-back must be finished before it is reused
                j, s_ = pending.pop(0)
                self._done[s_].synchronize()
                if consume is not None:
                    consume(j, self._host[s_])
            for dst, (_, src) in zip(self._stage[slot], picked):
                dst.copy_(src, non_blocking=True)      # D2D on the compute stream (outputs are graph-static)
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
