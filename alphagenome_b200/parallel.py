"""Sequence-parallel plumbing for the trunk forward (one process per GPU, torch.distributed / NCCL over NVLink).

The reference has no sequence parallelism; this is new work.  Partitioning of a length-L
context over G ranks (contiguous, L/G a multiple of 2048 bp so every pool-2 pair, 128 bp token and 16-token pair
bin is rank-local):

  rank r owns bp [r L/G, (r+1) L/G)  ->  tokens [r n/G, ...)  ->  pair ROWS [r P/G, ...) x all P columns.

  conv encoder / decoder   each rank runs the convs on its window EXTENDED by `halo_bp` real bases on each
                           interior side and crops.  A w5 conv contaminates 2 positions per layer from a cut
                           edge: 513 bp through the encoder, +762 bp through the decoder = 1275 bp < 2048 bp, so
                           everything inside the owned range is exact; at the two true sequence ends there is no
                           halo and the window edge IS the reference's zero padding.  Costs 2*2048/(L/G) extra
                           conv work (3 % at 1 Mbp / 8 GPUs) and only ONE exchange (16 tower-output tokens per
                           side, fp32 residual + bf16 operand packed into one collective) instead of 27
                           per-layer halo exchanges.
  tower attention          Q local; [K | V] = [n,128+192] bf16 all-gathered per layer in ONE collective.  Bias rows local.
  pair blocks              pooled/normalised single rows all-gathered (bf16 [P,1536] x2); q.k and rel-pos GEMMs,
                           row attention, FF and the bias projection run on local rows only.
  pair output embedding    symmetrize needs pair[j, i]: one all_to_all of [P/G x P/G] blocks.

Every helper below is device-agnostic (the CPU/gloo tests in tests/test_parallel_cpu.py run them with
world_size 2).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

HALO_BP = 2048  # >= 1275 bp (encoder 513 + decoder 762 receptive-field radius), multiple of 128
TOKEN_BP = 128
PAIR_BP = 2048


@dataclass(frozen=True)
class ShardPlan:
    rank: int
    world: int
    L: int            # global context length (bp)
    start: int        # owned bp range [start, end)
    end: int
    halo_left: int    # bp of real halo on each side of the window (0 at a true sequence end)
    halo_right: int

    @property
    def win_start(self) -> int:
        return self.start - self.halo_left

    @property
    def win_end(self) -> int:
        return self.end + self.halo_right

    @property
    def L_loc(self) -> int:
        return self.end - self.start

    @property
    def n_loc(self) -> int:
        return self.L_loc // TOKEN_BP

    @property
    def n(self) -> int:
        return self.L // TOKEN_BP

    @property
    def tok_off(self) -> int:
        return self.start // TOKEN_BP

    @property
    def halo_tok_left(self) -> int:
        return self.halo_left // TOKEN_BP

    @property
    def halo_tok_right(self) -> int:
        return self.halo_right // TOKEN_BP

    @property
    def P(self) -> int:
        return self.L // PAIR_BP

    @property
    def P_loc(self) -> int:
        return self.L_loc // PAIR_BP

    @property
    def row_off(self) -> int:
        return self.start // PAIR_BP


def shard_plan(L: int, world: int, rank: int, halo_bp: int = HALO_BP) -> ShardPlan:
    if L % (PAIR_BP * world) != 0:
        raise ValueError(f"context length {L} must be a multiple of {PAIR_BP} * world_size ({world})")
    if halo_bp % TOKEN_BP != 0:
        raise ValueError("halo must be a whole number of 128 bp tokens")
    L_loc = L // world
    if world > 1 and halo_bp > L_loc:
        raise ValueError(f"halo {halo_bp} bp exceeds the shard length {L_loc} bp")
    start = rank * L_loc
    return ShardPlan(rank=rank, world=world, L=L, start=start, end=start + L_loc,
                     halo_left=halo_bp if rank > 0 else 0, halo_right=halo_bp if rank < world - 1 else 0)


class SeqParallel:
    """Collectives used by TrunkEngine in sequence-parallel mode."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, halo_bp: int = HALO_BP):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.halo_bp = halo_bp
        self.n_collectives = 0
        # NCCL moves device buffers directly over NVLink.  With the gloo backend (CPU tests; the 2-process
        # single-GPU parity test) device tensors are staged through host memory — test plumbing only.
        self._stage = dist.get_backend(group) == "gloo"
        # Peer-memory path (NVLink / NVSwitch): gathers whose producer is one of our kernels are written by that kernel
        # straight into every rank's symmetric buffer, followed by ONE cross-rank barrier per tower layer, instead of
        # an NCCL all-gather per tensor.  Needs one GPU per rank and NCCL; AGB_SP_PEER=0 keeps the NCCL collectives.
        self._symm = {}
        self.peer = False
        if not self._stage and self.world > 1 and self.world <= 8 and torch.cuda.is_available() and os.environ.get("AGB_SP_PEER", "1") != "0":
            try:
                import torch.distributed._symmetric_memory as symm_mem  # noqa: F401

                self.peer = torch.cuda.device_count() >= self.world
            except Exception:  # noqa: BLE001 — no symmetric-memory support in this torch build
                self.peer = False
        self.n_barriers = 0

    def _all_gather(self, out: torch.Tensor, x: torch.Tensor) -> None:
        if self._stage and x.is_cuda:
            o = torch.empty(out.shape, dtype=out.dtype)
            dist.all_gather_into_tensor(o.view(-1), x.cpu().view(-1), group=self.group)
            out.copy_(o)
        else:
            dist.all_gather_into_tensor(out.view(-1), x.view(-1), group=self.group)
        self.n_collectives += 1

    # ------------------------------------------------------------------------------------------ peer memory
    def symm_buffer(self, key, shape, dtype, device):
        """A symmetric allocation (same shape on every rank, peer-mapped): returns (local tensor, handle).  Cached by
        `key`; the first call is a collective (rendezvous), so every rank must ask for the same keys in the same order."""
        ent = self._symm.get(key)
        if ent is None:
            import torch.distributed._symmetric_memory as symm_mem

            t = symm_mem.empty(*shape, dtype=dtype, device=device)
            hdl = symm_mem.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
            ent = (t, hdl, [int(p) for p in hdl.buffer_ptrs])
            self._symm[key] = ent
        return ent

    def peer_barrier(self, hdl, channel: int = 0):
        """All ranks' earlier stores into the symmetric buffers are visible to every rank's later kernels (stream-ordered,
        device-side; a lost peer trips the timeout instead of hanging the GPU)."""
        hdl.barrier(channel=channel, timeout_ms=20000)
        self.n_barriers += 1

    class _Pending:
        """An all-gather in flight on NCCL's stream; wait() orders the current stream after it and returns the result."""

        def __init__(self, work, out, finish, keep=None):
            self.work, self.out, self.finish = work, out, finish
            self.keep = keep      # the send buffer stays referenced until the collective has been waited for

        def wait(self):
            if self.work is not None:
                self.work.wait()
            return self.finish(self.out)

    def gather_rows_async(self, x: torch.Tensor) -> "SeqParallel._Pending":
        """gather_rows that does not block the current stream: the collective runs on NCCL's own stream while the
        caller keeps launching independent kernels; .wait() returns [B, world*rows_loc, C]."""
        x = x.contiguous()
        B, r, C = x.shape
        out = torch.empty(self.world, B, r, C, dtype=x.dtype, device=x.device)

        def finish(o):
            return o.view(1, self.world * r, C) if B == 1 else o.permute(1, 0, 2, 3).reshape(B, self.world * r, C)

        if self._stage and x.is_cuda:      # gloo test plumbing: host-staged, synchronous
            self._all_gather(out, x)
            return SeqParallel._Pending(None, out, finish)
        work = dist.all_gather_into_tensor(out.view(-1), x.view(-1), group=self.group, async_op=True)
        self.n_collectives += 1
        return SeqParallel._Pending(work, out, finish, keep=x)

    def plan(self, L: int) -> ShardPlan:
        return shard_plan(L, self.world, self.rank, self.halo_bp)

    def gather_rows(self, x: torch.Tensor) -> torch.Tensor:
        """x: [B, rows_loc, C] (same shape on every rank) -> [B, world*rows_loc, C] in rank order."""
        x = x.contiguous()
        B, r, C = x.shape
        out = torch.empty(self.world, B, r, C, dtype=x.dtype, device=x.device)
        self._all_gather(out, x)
        if B == 1:
            return out.view(1, self.world * r, C)
        return out.permute(1, 0, 2, 3).reshape(B, self.world * r, C)

    def exchange_halo(self, x: torch.Tensor, h: int):
        """x: [B, rows, C]; returns (left, right): the last h rows of rank-1 and the first h rows of rank+1
        (None at the two sequence ends)."""
        (left, right), = self.exchange_halos([x], h)
        return left, right

    def exchange_halos(self, xs, h: int):
        """Several [B, rows, C_i] tensors (any dtypes, same B and rows) in ONE collective: their edge rows are packed
        as bytes.  Returns [(left_i, right_i), ...] like exchange_halo."""
        return self.exchange_halos_async(xs, h).wait()

    def exchange_halos_async(self, xs, h: int) -> "SeqParallel._Pending":
        """exchange_halos that leaves the current stream free until .wait()."""
        B, r = xs[0].shape[0], xs[0].shape[1]
        parts = []
        for x in xs:
            edges = torch.stack((x[:, :h], x[:, r - h:]), dim=0).contiguous()          # [2, B, h, C]
            parts.append(edges.view(torch.uint8).view(2, B, h, -1))                     # bytes of each row
        packed = torch.cat(parts, dim=-1).contiguous() if len(parts) > 1 else parts[0]
        nb = packed.shape[-1]
        out = torch.empty(self.world, 2, B, h, nb, dtype=torch.uint8, device=packed.device)
        shapes = [(x.dtype, x.shape[2], x.shape[2] * x.element_size()) for x in xs]

        def finish(out):
            res, off = [], 0
            for dtype, C, w in shapes:
                def take(side_rank, side):
                    return out[side_rank, side, :, :, off:off + w].contiguous().view(dtype).view(B, h, C)

                left = take(self.rank - 1, 1) if self.rank > 0 else None
                right = take(self.rank + 1, 0) if self.rank < self.world - 1 else None
                res.append((left, right))
                off += w
            return res

        if self._stage and packed.is_cuda:
            self._all_gather(out, packed)
            return SeqParallel._Pending(None, out, finish)
        work = dist.all_gather_into_tensor(out.view(-1), packed.view(-1), group=self.group, async_op=True)
        self.n_collectives += 1
        return SeqParallel._Pending(work, out, finish, keep=packed)

    def all_reduce_sum(self, x: torch.Tensor) -> torch.Tensor:
        """Sum over ranks (splice-junction head: each gathered site row is non-zero on exactly one rank)."""
        x = x.contiguous()
        if self._stage and x.is_cuda:
            h = x.cpu()
            dist.all_reduce(h, group=self.group)
            x = h.to(x.device)
        else:
            dist.all_reduce(x, group=self.group)
        self.n_collectives += 1
        return x

    def transpose_blocks_async(self, pair: torch.Tensor) -> "SeqParallel._Pending":
        """transpose_blocks without blocking the current stream (the decoder runs while the blocks travel)."""
        B, Pl, P, C = pair.shape
        G = self.world
        send = pair.view(B, Pl, G, Pl, C).permute(2, 0, 1, 3, 4).contiguous()
        recv = torch.empty_like(send)
        if self._stage and send.is_cuda:
            return SeqParallel._Pending(None, self.transpose_blocks(pair), lambda o: o)
        work = dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group, async_op=True)
        self.n_collectives += 1
        return SeqParallel._Pending(work, recv, lambda o: o, keep=send)

    def transpose_blocks(self, pair: torch.Tensor) -> torch.Tensor:
        """pair: this rank's rows [B, P_loc, P, C] -> pairT [world, B, P_loc, P_loc, C] with
        pairT[s, b, jl, il] = pair_global[b, s*P_loc + jl, row_off + il]  (the layout ag_pair_out_embed_fwd reads)."""
        B, Pl, P, C = pair.shape
        G = self.world
        send = pair.view(B, Pl, G, Pl, C).permute(2, 0, 1, 3, 4).contiguous()  # [dest s, B, il, jl(cols of s), C]
        recv = torch.empty_like(send)                                          # [src s, B, jl (rows of s), il (my cols), C]
        if self._stage and send.is_cuda:
            r = torch.empty(send.shape, dtype=send.dtype)
            dist.all_to_all_single(r.view(-1), send.cpu().view(-1), group=self.group)
            recv.copy_(r)
        else:
            dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        self.n_collectives += 1
        return recv
