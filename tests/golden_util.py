"""Shared definition of the golden cases (inputs + which rows of each output are stored)."""
from __future__ import annotations

import numpy as np
import torch

# name, batch, length, weight seed, input pattern
CASES = [
    dict(name="ref_B1_L2048_seed42", B=1, L=2048, weight_seed=0, pattern="rng42"),       # = reference tests/regression_data input
    dict(name="ref_B2_L8192_mixed", B=2, L=8192, weight_seed=1, pattern="rng42_N"),      # BASELINE config 2 shape, N tokens, both organisms
    dict(name="ref_B1_L32768_pair16", B=1, L=32768, weight_seed=2, pattern="rng7"),      # 16x16 pair tensor
    dict(name="ref_B1_L4096_homopolymer", B=1, L=4096, weight_seed=0, pattern="homoA"),  # degenerate input (verification_utils.py:159-185)
]


def make_inputs(case):
    B, L, pat = case["B"], case["L"], case["pattern"]
    if pat == "rng42":
        seq = np.random.default_rng(42).integers(0, 4, size=(B, L))
        org = np.zeros(B, dtype=np.int64)
    elif pat == "rng42_N":
        seq = np.random.default_rng(42).integers(0, 4, size=(B, L))
        seq[0, 5] = -1
        seq[0, L - 1] = -1
        seq[1, 1000:1100] = -1
        org = np.arange(B, dtype=np.int64) % 2
    elif pat == "rng7":
        seq = np.random.default_rng(7).integers(0, 4, size=(B, L))
        org = np.ones(B, dtype=np.int64)
    elif pat == "homoA":
        seq = np.zeros((B, L), dtype=np.int64)
        seq[:, L // 2: L // 2 + 3] = 2
        org = np.zeros(B, dtype=np.int64)
    else:
        raise ValueError(pat)
    return torch.from_numpy(seq.astype(np.int64)), torch.from_numpy(org)


def sample_rows(name: str, shape):
    """Row indices (dim 1) stored in the fixture, or None for 'all'. 64 strided rows + 8 at each end."""
    if name in ("embeds_pair", "pairwise_repr"):
        return None
    n = shape[1]
    if n <= 96:
        return None
    stride = n // 64
    idx = np.unique(np.concatenate([np.arange(0, n, stride), np.arange(8), np.arange(n - 8, n)]))
    return torch.from_numpy(idx)


# prediction heads: publication heads of both organisms on every sample (AG:1823-1850)
HEAD_CASES = [
    dict(name="heads_B2_L4096", B=2, L=4096, weight_seed=3, n_sites=12),
    dict(name="heads_B1_L16384", B=1, L=16384, weight_seed=4, n_sites=40, first_org=1),   # 8x8 contact map, ragged site count
]


def make_head_inputs(case):
    B, L = case["B"], case["L"]
    rng = np.random.default_rng(99)
    seq = rng.integers(0, 4, size=(B, L))
    seq[0, 17] = -1
    org = (np.arange(B, dtype=np.int64) + case.get("first_org", 0)) % 2
    donor = np.sort(rng.choice(L, size=(B, case["n_sites"]), replace=True), axis=1)
    acceptor = np.sort(rng.choice(L, size=(B, case["n_sites"]), replace=True), axis=1)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int64))
    return t(seq), t(org), t(donor), t(acceptor)


# JAX-aligned heads (`add_reference_heads`, AG:1684-1731): genome track heads per resolution, contact maps, splice sites
REF_HEAD_CASES = [
    dict(name="refheads_B2_L4096", B=2, L=4096, weight_seed=6, n_sites=10),
]


def make_ref_head_inputs(case):
    """tokens, organism index and splice_site_positions [B, 4, P] (pos donor, pos acceptor, neg donor, neg acceptor;
    -1 = padding, AG:1503-1508)."""
    B, L, P = case["B"], case["L"], case["n_sites"]
    rng = np.random.default_rng(123)
    seq = rng.integers(0, 4, size=(B, L))
    org = np.arange(B, dtype=np.int64) % 2
    pos = np.sort(rng.choice(L, size=(B, 4, P)), axis=-1)
    pos[0, 0, -2:] = -1
    pos[1, 3, -3:] = -1
    pos[1, 1, -1:] = -1
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a).astype(np.int64))
    return t(seq), t(org), t(pos)


# BatchRMSNorm with updating statistics (the reference's default; AG:324-361): two consecutive forwards with different
# inputs, so the second one normalises with statistics the first one left behind
UPDATE_CASE = dict(name="update_B2_L4096", B=2, L=4096, weight_seed=8)


def make_update_inputs(case):
    B, L = case["B"], case["L"]
    rng = np.random.default_rng(77)
    out = []
    for step in range(2):
        seq = rng.integers(0, 4, size=(B, L))
        seq[0, 11 + step] = -1
        org = (np.arange(B, dtype=np.int64) + step) % 2
        out.append((torch.from_numpy(seq.astype(np.int64)), torch.from_numpy(org)))
    return out


# DNAModel callers: predict with the two-pass splice flow and in-silico mutagenesis
DNA_MODEL_CASE = dict(name="dna_model_L2048", L=2048, weight_seed=10, num_splice_sites=16, threshold=0.0, ism_batch=5)


def make_dna_model_inputs(case):
    """A DNA string with a few N and the positions to mutate (one of them on an N, one at each end)."""
    rng = np.random.default_rng(55)
    s = np.array(list("ACGT"))[rng.integers(0, 4, size=case["L"])]
    s[700:704] = "N"
    return "".join(s), [0, 5, 701, 1023, 1024, case["L"] - 1]
