"""A deterministic synthetic JAX/Haiku-layout checkpoint of the published architecture (no network for the real one):
every array the conversion table reads, in the checkpoint's own layout and shape, filled from numpy PCG64 keyed by the
array's name.  Shapes are derived from the PyTorch model's state_dict by inverting each layout rule."""
from __future__ import annotations

import zlib

import numpy as np
import torch

from alphagenome_b200.convert import Rule, build_rules


def _rand(name: str, shape, seed: int, positive: bool = False) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
    a = rng.standard_normal(size=shape).astype(np.float32) * 0.05
    return (np.abs(a) + 0.5).astype(np.float32) if positive else a


def jax_shapes(rule: Rule, torch_shape, bias_shape=None, n_org: int = 2):
    """{checkpoint key: shape} for the arrays `rule` reads, given the shape of the state_dict entry it produces."""
    ts = tuple(torch_shape)
    lead = (n_org,) if rule.index is not None else ()
    op = rule.op
    if op == "copy":
        return {rule.src[0]: lead + ts}
    if op == "linear":
        return {rule.src[0]: lead + (ts[1], ts[0])}
    if op == "conv":
        return {rule.src[0]: (ts[2], ts[1], ts[0])}
    if op == "pointwise":
        return {rule.src[0]: (ts[1], ts[0])}
    if op == "squeeze":
        return {rule.src[0]: (1, 1, ts[0])}
    if op == "scalar":
        return {rule.src[0]: ()}
    if op == "stack":
        return {k: ts[1:] for k in rule.src}
    if op == "standardized":
        ex = dict(rule.extra)
        return {rule.src[0]: (ts[2], ts[1], ts[0]), ex["scale"]: (1, 1, ts[0]), ex["bias"]: (ts[0],)}
    raise ValueError(op)


def concat_split(torch_key: str, rows: int):
    """Row counts of the projections concatenated into one PyTorch weight."""
    if torch_key.endswith("to_qkv.weight"):      # 8 heads x 128 | 128 | 192 (AG:672-676)
        return [rows - 320, 128, 192]
    return [rows // 2, rows // 2]                 # q | k halves


def synth_checkpoint(model, seed: int = 0):
    """(flat params, flat state) as numpy arrays, covering every rule whose state_dict entry `model` has."""
    sd = model.state_dict()
    params, state = {}, {}
    for rule in build_rules():
        if rule.torch_key not in sd:
            continue
        ts = tuple(sd[rule.torch_key].shape)
        dst = state if rule.state else params
        if rule.op == "concat":
            for k, r in zip(rule.src, concat_split(rule.torch_key, ts[0])):
                dst[k] = _rand(k, (ts[1], r), seed)
            continue
        for k, shape in jax_shapes(rule, ts).items():
            if k not in dst:   # multi-organism head arrays are shared by both organisms' rules
                dst[k] = _rand(k, shape, seed, positive=k.endswith("var_ema"))
    return params, state


def checkpoint_from_state_dict(sd, n_org: int = 2):
    """The inverse problem: a JAX-layout checkpoint whose conversion reproduces `sd` (a PyTorch state_dict with the
    checkpoint's heads for both organisms) — used to drive the conversion path with a WELL-BEHAVED weight distribution
    (alphagenome_b200.init_weights).  Exact for every rule except the standardized convs, whose kernels come back
    centred per output channel (the bake always centres): scale is chosen as sqrt(fan_in * var) so that nothing else
    changes."""
    params, state = {}, {}
    per_org = {}
    for rule in build_rules():
        if rule.torch_key not in sd:
            continue
        t = sd[rule.torch_key].detach().float().cpu()
        dst = state if rule.state else params
        op = rule.op
        if op == "concat":
            outs = [x.t().contiguous() for x in torch.split(t, concat_split(rule.torch_key, t.shape[0]), dim=0)]
        elif op == "stack":
            outs = list(t.unbind(0))
        elif op == "copy":
            outs = [t]
        elif op == "linear":
            outs = [t.t().contiguous()]
        elif op == "conv":
            outs = [t.permute(2, 1, 0).contiguous()]
        elif op == "pointwise":
            outs = [t.squeeze(-1).t().contiguous()]
        elif op == "squeeze":
            outs = [t.reshape(1, 1, -1)]
        elif op == "scalar":
            outs = [t.reshape(())]
        elif op == "standardized":
            w = t.permute(2, 1, 0).contiguous()                       # (width, in, out)
            var = (w - w.mean(dim=(0, 1), keepdim=True)).var(dim=(0, 1), keepdim=True, unbiased=False)
            ex = dict(rule.extra)
            dst[ex["scale"]] = torch.sqrt(w.shape[0] * w.shape[1] * var).numpy()
            dst[ex["bias"]] = sd[rule.torch_key[: -len("weight")] + "bias"].detach().float().cpu().numpy()
            outs = [w]
        else:
            raise ValueError(op)
        for k, o in zip(rule.src, outs):
            if rule.index is None:
                dst[k] = o.numpy()
            else:
                per_org.setdefault(k, {})[rule.index] = o
    for k, by in per_org.items():
        params[k] = torch.stack([by[i] for i in range(n_org)], dim=0).numpy()
    return params, state
