"""Deterministic synthetic weights for parity tests and benchmarks.

There is no network for checkpoints, and `torch.manual_seed(0); AlphaGenome()` depends on module construction
order and the torch version.  This generator depends only on (seed, parameter name, shape) through numpy's
PCG64, so the same tensors can be produced in this container (to run the real reference and record golden
outputs) and on the GPU box (to run the CUDA path), without shipping 1.6 GB of weights.

Unlike the default init (gamma=1, beta=0, running_var=1, zero biases), every norm scale/shift, running
variance, bias and `qk_rel_pos_bias` is perturbed, so a kernel that ignores one of them fails parity.
"""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Iterable, Tuple

import numpy as np
import torch

# buffers defined by a formula in the reference (AG:541-543, AG:581): never randomised
FORMULA_BUFFERS = ("rotary_emb.inv_freq", "rel_pos_features.dummy", "rope.inv_freq")


def _rng(seed: int, name: str) -> np.random.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    return np.random.default_rng(int.from_bytes(h[:8], "little"))


def synth_tensor(seed: int, name: str, shape: Tuple[int, ...]) -> np.ndarray:
    r = _rng(seed, name)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "running_var":
        return r.uniform(0.5, 1.5, shape).astype(np.float32)
    if leaf == "residual_scale":
        return r.uniform(0.7, 1.0, shape).astype(np.float32)
    if leaf == "qk_rel_pos_bias":
        return r.uniform(-0.2, 0.2, shape).astype(np.float32)
    if leaf == "gamma" or (leaf == "weight" and len(shape) == 1):
        return (1.0 + r.uniform(-0.2, 0.2, shape)).astype(np.float32)
    if leaf in ("beta", "bias"):
        return r.uniform(-0.1, 0.1, shape).astype(np.float32)
    if leaf == "weight" and ".embed." in "." + name:
        return r.standard_normal(shape).astype(np.float32)  # nn.Embedding default N(0, 1)
    if leaf == "scale" and len(shape) == 1:      # TracksScaledPrediction.scale (AG:1386), used as softplus(scale)
        return r.uniform(-0.5, 0.5, shape).astype(np.float32)
    if leaf == "scale":                           # SpliceJunctionHead.scale [contexts, hidden] (AG:1333), init ones
        return (1.0 + r.uniform(-0.2, 0.2, shape)).astype(np.float32)
    if leaf == "offset":                          # SpliceJunctionHead.offset (AG:1334), init zeros
        return r.uniform(-0.1, 0.1, shape).astype(np.float32)
    if leaf.endswith("_embeddings") and len(shape) == 1:   # SpliceSitesJunctionHead: [scale | offset] flat (AG:1473-1476)
        half = shape[0] // 2
        return np.concatenate([1.0 + r.uniform(-0.2, 0.2, half), r.uniform(-0.1, 0.1, shape[0] - half)]).astype(np.float32)
    if leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        return r.uniform(-bound, bound, shape).astype(np.float32)
    raise KeyError(f"no synthetic rule for parameter {name!r} with shape {shape}")


def synth_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    """shapes: name -> shape for every floating-point entry of a state_dict (formula buffers are skipped)."""
    out = {}
    for name, shape in shapes.items():
        if name.endswith(FORMULA_BUFFERS):
            continue
        out[name] = torch.from_numpy(synth_tensor(seed, name, tuple(shape)))
    return out


def apply_synth_weights(model: torch.nn.Module, seed: int = 0, keys: Iterable[str] | None = None) -> None:
    """Overwrite every float parameter/buffer of `model` (except formula buffers) in place."""
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items() if v.is_floating_point() and (keys is None or k in keys)}
    new = synth_state_dict(shapes, seed)
    with torch.no_grad():
        for k, v in new.items():
            sd[k].copy_(v.to(sd[k].device))
