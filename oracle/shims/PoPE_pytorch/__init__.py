"""Import shim: PoPE is non-default (polar_pos_emb=False) and outside every BASELINE config."""


class PoPE:  # noqa: D101
    def __init__(self, *a, **k):
        raise NotImplementedError("PoPE_pytorch is not available offline; polar_pos_emb=True is out of scope")


def compute_attn_similarity(*a, **k):
    raise NotImplementedError("PoPE_pytorch is not available offline")
