"""Import shim: torch_einops_utils.pack_with_inverse (only reached on the PoPE branch, AG:741)."""
from einops import pack, unpack


def pack_with_inverse(t, pattern):
    packed, ps = pack([t], pattern)

    def inverse(out, inv_pattern=None):
        (res,) = unpack(out, ps, inv_pattern or pattern)
        return res

    return packed, inverse
