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
