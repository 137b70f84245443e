"""CPU: the multi-rank path.  (1) world_size-2 gloo run of the collectives TrunkEngine uses in sequence-parallel
mode (row gather, halo exchange, pair block transpose).  (2) The sharding scheme itself — run the conv encoder /
decoder on a window extended by a 2048 bp halo and crop — is shown to reproduce the unsharded result using the
pinned oracle, so the halo size (>= 513 + 762 bp receptive-field radius) is a tested fact, not an estimate."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alphagenome_b200.parallel import HALO_BP, SeqParallel, shard_plan


def test_shard_plan_partitions_the_context():
    L, G = 1_048_576, 8
    plans = [shard_plan(L, G, r) for r in range(G)]
    assert [p.start for p in plans] == [r * L // G for r in range(G)]
    assert plans[0].halo_left == 0 and plans[-1].halo_right == 0
    assert all(p.halo_left == HALO_BP for p in plans[1:]) and all(p.halo_right == HALO_BP for p in plans[:-1])
    assert sum(p.L_loc for p in plans) == L and all(p.n_loc == 1024 and p.P_loc == 64 for p in plans)
    assert plans[3].tok_off == 3 * 1024 and plans[3].row_off == 3 * 64 and plans[3].P == 512
    with pytest.raises(ValueError):
        shard_plan(2048 * 3, 2, 0)
    one = shard_plan(8192, 1, 0)
    assert one.win_start == 0 and one.win_end == 8192


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sp = SeqParallel()
        B, rows, C = 2, 6, 4
        full = torch.arange(B * world * rows * C, dtype=torch.float32).view(B, world * rows, C)
        mine = full[:, rank * rows:(rank + 1) * rows].contiguous()
        ok = torch.equal(sp.gather_rows(mine), full)
        left, right = sp.exchange_halo(mine, 2)
        if rank > 0:
            ok &= torch.equal(left, full[:, rank * rows - 2:rank * rows])
        else:
            ok &= left is None
        if rank < world - 1:
            ok &= torch.equal(right, full[:, (rank + 1) * rows:(rank + 1) * rows + 2])
        else:
            ok &= right is None
        # pair transpose: pairT[s, b, jl, il] == pair_global[b, s*Pl + jl, row_off + il]
        Pl, Cc = 3, 5
        P = world * Pl
        pair = torch.randn(B, P, P, Cc, generator=torch.Generator().manual_seed(0))
        pT = sp.transpose_blocks(pair[:, rank * Pl:(rank + 1) * Pl].contiguous())
        for s in range(world):
            want = pair[:, s * Pl:(s + 1) * Pl, rank * Pl:(rank + 1) * Pl]
            ok &= torch.equal(pT[s], want)
        # two tensors of different dtype / width packed into ONE halo collective (decoder input: fp32 residual + bf16 operand)
        mine16 = (mine[..., :2] * 0.5).to(torch.bfloat16).contiguous()
        (l32, r32), (l16, r16) = sp.exchange_halos([mine, mine16], 2)
        if rank > 0:
            ok &= torch.equal(l32, full[:, rank * rows - 2:rank * rows])
            ok &= torch.equal(l16, (full[:, rank * rows - 2:rank * rows, :2] * 0.5).to(torch.bfloat16))
        else:
            ok &= l32 is None and l16 is None
        if rank < world - 1:
            ok &= torch.equal(r16, (full[:, (rank + 1) * rows:(rank + 1) * rows + 2, :2] * 0.5).to(torch.bfloat16))
        else:
            ok &= r32 is None and r16 is None
        # masked sum over ranks (splice-junction head: each gathered site row is owned by exactly one rank)
        owned = torch.zeros(B, 4, 3)
        owned[:, rank::world] = float(rank + 1)
        tot = sp.all_reduce_sum(owned)
        want_tot = torch.zeros(B, 4, 3)
        for r in range(world):
            want_tot[:, r::world] = float(r + 1)
        ok &= torch.equal(tot, want_tot)
        q.put((rank, bool(ok), sp.n_collectives))
    finally:
        dist.destroy_process_group()


def test_collectives_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] for r in res), res
    assert all(r[2] == 5 for r in res)   # gather, halo, transpose, packed halos, all-reduce


def test_halo_window_scheme_is_exact_for_the_conv_stacks():
    """Oracle emulation of the sharded conv encoder/decoder on 2 ranks at L = 16384 (default architecture)."""
    from alphagenome_b200.init_weights import synth_state_dict
    from oracle import trunk_oracle as O
    from tests.test_oracle_cpu import reference_shapes

    shapes = {k: v for k, v in reference_shapes().items() if k.startswith("transformer_unet.dna_embed") or
              k.startswith("transformer_unet.downs") or k.startswith("transformer_unet.ups")}
    sd = synth_state_dict(shapes, seed=3)
    L, G = 16384, 2
    p = "transformer_unet."
    seq = torch.from_numpy(np.random.default_rng(5).integers(0, 4, size=(1, L)))

    def encode(tokens):
        x, skip = O.dna_embed(sd, p + "dna_embed.", tokens)
        skips = [skip]
        for i in range(6):
            x, skip = O.down_block(sd, f"{p}downs.{i}.", x)
            skips.append(skip)
        return x, skips

    def decode(x, skips):
        skips = list(skips)
        for j in range(7):
            x = O.up_block(sd, f"{p}ups.{j}.", x, skips.pop())
        return x

    with torch.no_grad():
        x_full, skips_full = encode(seq)
        single_full = torch.tanh(x_full) * 3  # any token-local stand-in for the tower
        y_full = decode(single_full, skips_full)
        for r in range(G):
            pl = shard_plan(L, G, r)
            x_win, skips_win = encode(seq[:, pl.win_start:pl.win_end])
            hl = pl.halo_tok_left
            x_loc = x_win[:, hl:hl + pl.n_loc]
            want = x_full[:, pl.tok_off:pl.tok_off + pl.n_loc]
            assert (x_loc - want).abs().max() <= 1e-5 * want.abs().max()
            # decoder: extended tower output = neighbours' halo tokens + own tokens
            lo, hi = pl.tok_off - pl.halo_tok_left, pl.tok_off + pl.n_loc + pl.halo_tok_right
            y_win = decode(single_full[:, lo:hi], skips_win)
            y_loc = y_win[:, pl.halo_left:pl.halo_left + pl.L_loc]
            want = y_full[:, pl.start:pl.end]
            assert (y_loc - want).abs().max() <= 1e-5 * want.abs().max()
