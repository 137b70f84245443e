"""Import shim (test infrastructure, this container only): the three einx 0.3.0 elementwise ops the reference
calls, restricted to the 11 patterns that occur in alphagenome.py.  einx is not
installed and there is no network; the arithmetic of the reference itself is all torch."""
import torch


def _parse(pattern):
    if "->" in pattern:
        lhs, out = pattern.split("->")
        out = out.split()
    else:
        lhs, out = pattern, None
    ins = [p.split() for p in lhs.split(",")]
    return ins, out


def _expand_ellipsis(names, ndim):
    if "..." not in names:
        return names
    i = names.index("...")
    n_extra = ndim - (len(names) - 1)
    return names[:i] + [f"_e{k}" for k in range(n_extra)] + names[i + 1:]


def _binary(op, pattern, a, b):
    ins, out = _parse(pattern)
    ts = [torch.as_tensor(a), torch.as_tensor(b)]
    ins = [_expand_ellipsis(n, t.ndim) for n, t in zip(ins, ts)]
    if out is None:
        out = max(ins, key=len)
    else:
        out = _expand_ellipsis(out, max(t.ndim for t in ts))
    views = []
    for names, t in zip(ins, ts):
        assert len(names) == t.ndim, (pattern, names, t.shape)
        perm = [names.index(n) for n in out if n in names]
        t = t.permute(perm)
        shape = []
        it = iter(t.shape)
        for n in out:
            shape.append(next(it) if n in names else 1)
        views.append(t.reshape(shape))
    return op(views[0], views[1])


def add(pattern, a, b):
    return _binary(torch.add, pattern, a, b)


def multiply(pattern, a, b):
    return _binary(torch.mul, pattern, a, b)


def greater(pattern, a, b):
    return _binary(torch.gt, pattern, a, b)
