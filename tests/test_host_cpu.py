"""CPU: host-side logic that needs no GPU — the C-ABI library loads and exports every declared symbol, the host
model mirrors the reference's parameter contract, the weight generator is deterministic, the header and the
ctypes mirror agree on struct sizes, and the product path refuses to run without CUDA."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

from alphagenome_b200 import _lib
from alphagenome_b200.init_weights import synth_state_dict, synth_tensor
from alphagenome_b200.model import AlphaGenome, Embeds, publication_heads_config

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from alphagenome_b200.csrc.build import build

    build()
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    header = (ROOT / "include" / "agb200.h").read_text()
    declared = set(re.findall(r"\b(ag_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    assert declared <= set(_lib.EXPORTED), declared - set(_lib.EXPORTED)
    for name in declared | set(_lib.EXPORTED):
        assert hasattr(lib, name), f"libagb200.so does not export {name}"
    assert lib.ag_version() == 1
    assert lib.ag_launch_count() == 0  # nothing launched on a CPU-only box


def test_struct_layout_matches_header(lib, tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "agb200.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(AgEpiOut), sizeof(AgGemmArgs), sizeof(AgAttnArgs));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.AgEpiOut), C.sizeof(_lib.AgGemmArgs), C.sizeof(_lib.AgAttnArgs)]
    # field offsets too (equal sizes do not rule out two swapped fields): every field of the two argument structs
    checks = [("AgGemmArgs", _lib.AgGemmArgs), ("AgAttnArgs", _lib.AgAttnArgs)]
    body = "".join(f'printf("%zu ", offsetof({cn}, {fn}));' for cn, cls in checks for fn, _ in cls._fields_)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "agb200.h"\nint main(){{{body}return 0;}}\n')
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    offs = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [getattr(cls, fn).offset for _, cls in checks for fn, _ in cls._fields_]
    assert offs == want


def test_no_gpu_is_a_loud_error(lib):
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    assert lib.ag_check_device(0) != 0
    assert _lib.last_error()
    m = AlphaGenome(dims=(128, 256), transformer_kwargs=dict(depth=1))
    with pytest.raises(_lib.AgbError):
        m(torch.zeros(1, 2048, dtype=torch.long), 0)


def test_state_dict_contract_matches_reference():
    ref = {}
    for line in (ROOT / "tests" / "golden" / "state_dict_keys.txt").read_text().splitlines():
        k, shp = re.match(r"(\S+) \((.*)\)", line).groups()
        ref[k] = tuple(int(x) for x in shp.replace(",", " ").split())
    mine = {k: tuple(v.shape) for k, v in AlphaGenome().state_dict().items()}
    assert mine == ref
    assert Embeds._fields == ("embeds_1bp", "embeds_128bp", "embeds_pair")
    assert publication_heads_config["human"]["num_tracks_1bp"] == 3165


def test_synthetic_weights_are_deterministic_and_perturbed():
    a = synth_tensor(0, "transformer_unet.downs.0.conv.net.0.running_var", (896,))
    b = synth_tensor(0, "transformer_unet.downs.0.conv.net.0.running_var", (896,))
    c = synth_tensor(1, "transformer_unet.downs.0.conv.net.0.running_var", (896,))
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert a.min() >= 0.5 and a.max() <= 1.5
    sd = synth_state_dict({"x.gamma": (8,), "x.beta": (8,), "y.weight": (4, 16), "rotary_emb.inv_freq": (64,)})
    assert "rotary_emb.inv_freq" not in sd  # formula buffer, never randomised
    assert abs(float(sd["x.gamma"].mean()) - 1) < 0.2 and float(sd["y.weight"].abs().max()) <= 0.25


@pytest.mark.skipif(not Path("/root/reference/alphagenome_pytorch").exists(), reason="reference checkout not present (GPU box)")
def test_install_on_the_real_reference_model_packs_identically():
    """install() must find every parameter it needs on an unmodified reference AlphaGenome: pack both the
    reference model and our mirror (same state_dict) on CPU and compare every packed tensor bit for bit."""
    import sys

    sys.path.insert(0, str(ROOT / "oracle" / "shims"))
    sys.path.append("/root/reference")
    from alphagenome_pytorch.alphagenome import AlphaGenome as RefAlphaGenome

    import alphagenome_b200
    from alphagenome_b200.init_weights import apply_synth_weights
    from alphagenome_b200.trunk import engine_for

    kw = dict(dims=(128, 256, 384), transformer_kwargs=dict(depth=2))
    ref = RefAlphaGenome(**kw)
    apply_synth_weights(ref, seed=5)
    mine = alphagenome_b200.AlphaGenome(**kw)
    mine.load_state_dict(ref.state_dict(), strict=True)
    alphagenome_b200.install(ref)
    packs = []
    for m in (ref, mine):
        eng = engine_for(m.transformer_unet, m)
        eng._allow_cpu_pack = True
        packs.append(eng.pack(torch.device("cpu")))

    def flat(x, prefix=""):
        if isinstance(x, dict):
            for k, v in x.items():
                yield from flat(v, f"{prefix}.{k}")
        elif isinstance(x, (list, tuple)):
            for i, v in enumerate(x):
                yield from flat(v, f"{prefix}[{i}]")
        else:
            yield prefix, x

    a, b = dict(flat(packs[0])), dict(flat(packs[1]))
    assert a.keys() == b.keys() and len(a) > 100
    for k in a:
        if torch.is_tensor(a[k]):
            assert torch.equal(a[k], b[k]), k
        else:
            assert a[k] == b[k], k
    feats = ref.transformer_unet.transformer.rel_pos_features.features(8)
    assert torch.equal(feats, mine.transformer_unet.transformer.rel_pos_features.features(8))
    # and the reference's own RelativePosFeatures.forward agrees with ours (AG:587-604)
    want = ref.transformer_unet.transformer.rel_pos_features(torch.zeros(1, 8 * 16, 4))
    assert torch.equal(feats, want)


def test_heads_state_dict_matches_reference():
    """add_heads(**publication_heads_config[org]) registers the reference's head parameters under the reference's names
    and shapes (tests/golden/state_dict_keys_heads.txt is dumped from the reference, tools/make_golden_heads.py)."""
    from alphagenome_b200.model import AlphaGenome, publication_heads_config

    m = AlphaGenome()
    for org in ("human", "mouse"):
        m.add_heads(**publication_heads_config[org])
    sd = m.state_dict()
    want = {}
    for line in (Path(__file__).resolve().parent / "golden" / "state_dict_keys_heads.txt").read_text().splitlines():
        k, shape = line.split(" ", 1)
        want[k] = shape
    got = {k: str(tuple(v.shape)) for k, v in sd.items() if k.startswith("heads.")}
    assert got == want
    assert m.head_forward_arg_names["human"]["splice_juncs"] == ["embeds_1bp", "splice_donor_idx", "splice_acceptor_idx"]
    assert tuple(m.head_forward_arg_names["mouse"]["1bp_tracks"]) == ("embeds_1bp",)
    assert m.head_forward_arg_maps["human"]["1bp_tracks"] == {"embeds_1bp": "x"}


def test_reference_heads_state_dict_matches_reference():
    """add_reference_heads (JAX-aligned heads, AG:1684-1731): same 98 head keys / shapes as the reference."""
    from alphagenome_b200.model import AlphaGenome

    m = AlphaGenome()
    for org in ("human", "mouse"):
        m.add_reference_heads(org)
    want = dict(line.split(" ", 1) for line in (Path(__file__).resolve().parent / "golden" / "state_dict_keys_ref_heads.txt").read_text().splitlines())
    got = {k: str(tuple(v.shape)) for k, v in m.state_dict().items() if k.startswith("heads.")}
    assert got == want and len(got) == 98
    assert m.head_forward_arg_names["human"]["rna_seq"] == ["embeds_1bp", "embeds_128bp"]
    assert m.head_forward_arg_names["mouse"]["splice_sites_junction"] == ["embeds_1bp", "splice_site_positions"]


def test_heads_engine_packs_padded_bf16_weights_on_cpu():
    """HeadsEngine.pack is host logic: every Linear becomes a bf16 [1, N, K] operand plus a bias padded to a multiple of
    16 (the kernels compute 16-padded outputs and return the [..., :N] view); softplus(scale) is folded once."""
    import torch.nn.functional as F

    from alphagenome_b200.heads import HeadsEngine
    from alphagenome_b200.init_weights import apply_synth_weights
    from alphagenome_b200.model import AlphaGenome, publication_heads_config

    m = AlphaGenome()
    m.add_heads(**publication_heads_config["mouse"])
    apply_synth_weights(m, seed=2)
    P = HeadsEngine(m).pack(torch.device("cpu"))
    e = P[("mouse", "1bp_tracks")]
    head = m.heads["mouse"]["1bp_tracks"]
    # 730 tracks run at width 768 (256-wide MMA tiles) instead of 736 = 23 x 32 (heads._padded_width)
    assert e["N"] == 730 and e["N16"] == 768 and tuple(e["W"].shape) == (1, 730, 1536) and e["W"].dtype == torch.bfloat16
    assert torch.equal(e["b"][:730], head.to_pred.bias.detach()) and torch.all(e["b"][730:] == 0)
    assert torch.allclose(e["mul"][:730], F.softplus(head.scale.detach())) and torch.all(e["mul"][730:] == 1)
    sj = P[("mouse", "splice_juncs")]
    assert sj["T"] == 75 and sj["Hd"] == 512 and tuple(sj["cos"].shape) == (75, 256)
    # rotary position = context index (AG:1356): angle[t, j] = t * inv_freq[j]
    inv = m.heads["mouse"]["splice_juncs"].rope.inv_freq
    assert torch.allclose(sj["cos"][3], torch.cos(3.0 * inv)) and torch.allclose(sj["sin"][74], torch.sin(74.0 * inv))
    ch = P[("mouse", "contact_head")]
    assert ch["N"] == 8 and ch["N16"] == 16 and tuple(ch["emb"].shape) == (2, 128)
    from alphagenome_b200.heads import _gemm_n_tile, _padded_width

    assert (_padded_width(3165), _gemm_n_tile(3328)) == (3328, 256) and _padded_width(2733) == 2816 and _padded_width(282) == 288
    assert _padded_width(1152) == 1152 and _padded_width(768) == 768 and _padded_width(5) == 16
    # a second pack with unchanged parameters is the cached object; an in-place update invalidates it
    eng = HeadsEngine(m)
    p1 = eng.pack(torch.device("cpu"))
    assert eng.pack(torch.device("cpu")) is p1
    with torch.no_grad():
        head.scale.add_(1.0)
    assert eng.pack(torch.device("cpu")) is not p1


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours): one JSON line with the contract's keys.
    The sample window is shrunk through AGB_BENCH_CPU_SAMPLE_BP so the test stays short."""
    import json
    import os
    import sys

    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, AGB_BENCH_CPU_SAMPLE_BP="4096")
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "trunk_forward_bp_per_s" and line["unit"] == "bp/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "4096" in line["cpu_baseline"]["sample"]
    assert line["e2e"] == dict(value=line["value"], unit="bp/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in line["config"]
    # ranks other than 0 of a torchrun launch exit 0 without work
    r2 = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                        env=dict(env, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=120)
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_one_mufu_erf_gelu_coefficients():
    """ag_pair_bias_fwd's tensor-core path evaluates erf-GELU as 0.5 x (1 + tanh(x (c1 + c3 x^2 + c5 x^4))) with x^2 clamped to
    64 (elementwise.cu: gelu_erf_tanh).  The coefficients are read from the source, so the documented bound (max |error|
    2.5e-5 against nn.GELU before tanh.approx's own 2^-11) cannot drift from the code."""
    import math
    import re

    src = (ROOT / "alphagenome_b200" / "csrc" / "elementwise.cu").read_text()
    body = src[src.index("float gelu_erf_tanh(float x)"):]
    body = body[: body.index("}")]
    c5, c3 = (float(v) for v in re.search(r"fmaf\(t, (-?[\d.e+-]+)f, ([\d.e+-]+)f\)", body).groups())
    c1 = float(re.search(r"fmaf\(t, q, ([\d.e+-]+)f\)", body).group(1))
    clamp = float(re.search(r"fminf\(x \* x, ([\d.]+)f\)", body).group(1))
    x = torch.linspace(-30, 30, 600001, dtype=torch.float64)
    t = (x * x).clamp(max=clamp)
    approx = 0.5 * x * (1 + torch.tanh(x * (c1 + t * (c3 + t * c5))))
    exact = 0.5 * x * (1 + torch.erf(x / math.sqrt(2)))
    assert (approx - exact).abs().max().item() < 2.6e-5


def test_bench_clock_sampler_keeps_rows_of_the_timed_window(tmp_path, monkeypatch):
    """bench.py's `clocks` key: nvidia-smi is started before the warm-up (its start-up alone can outlast a 0.3 s timed
    region) and only the rows printed between mark_begin() and mark_end() count.  A stand-in nvidia-smi prints a row every
    20 ms: 1200 MHz for the first 8 rows ("warm-up"), 1500 MHz with sw_power_cap active afterwards."""
    import os
    import stat
    import sys
    import time

    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nfor i in $(seq 1 400); do\n  if [ $i -le 8 ]; then echo '1200, 1965, 700.0, Not Active, Not Active, Not Active, Not Active';\n"
                    "  else echo '1500, 1965, 990.0, Not Active, Not Active, Not Active, Active'; fi\n  sleep 0.02\ndone\n")
    fake.chmod(fake.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", f"{tmp_path}:{os.environ['PATH']}")
    sys.path.insert(0, str(ROOT))
    try:
        import bench
    finally:
        sys.path.pop(0)
    s = bench.ClockSampler(0)
    s.start()
    assert s.rows, "the sampler waits for nvidia-smi's first row"
    time.sleep(1.0)          # "warm-up": rows arrive but are outside the window
    s.mark_begin()
    time.sleep(0.3)
    s.mark_end()
    out = s.stop()
    assert out["samples"] >= 3 and out["sm_mhz"] == 1500.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def test_lazy_rescale_online_softmax_is_the_exact_softmax():
    """The arithmetic of row_attention_kernel (attention.cu), restated in fp64: key blocks of 128, a running row maximum
    that is only raised — and O / the row sum only rescaled — when a block's maximum exceeds it by more than 2^8, weights
    taken relative to the (possibly stale) running maximum.  Whatever the stale maximum is, it scales P and the row sum
    alike, so the quotient is the exact softmax; and no weight ever exceeds 2^8."""
    import math

    g = torch.Generator().manual_seed(0)
    n_q, n_k, d, dv, blk = 64, 640, 128, 128, 128
    q = torch.randn(n_q, d, generator=g, dtype=torch.float64) * 1.5
    k = torch.randn(n_k, d, generator=g, dtype=torch.float64) * 1.5
    k[128:] *= 4.0                      # later blocks raise most row maxima by far more than the threshold ...
    k[256:384] *= 0.02                  # ... one block does not ...
    q[::5] *= 0.01                      # ... and some rows never cross it (stale maximum kept for the whole row)
    v = torch.randn(n_k, dv, generator=g, dtype=torch.float64)
    scale = d ** -0.5
    c = scale * math.log2(math.e)       # the kernel works in the log2 domain: p = 2^(s c - m c)
    s = q @ k.T
    want = torch.softmax(s * scale, dim=-1) @ v

    m_run = torch.zeros(n_q, dtype=torch.float64)
    rsum = torch.zeros(n_q, dtype=torch.float64)
    O = torch.zeros(n_q, dv, dtype=torch.float64)
    rescales, pmax = 0, 0.0
    for j in range(0, n_k, blk):
        sb = s[:, j:j + blk]
        bmax = sb.max(dim=1).values
        if j == 0:
            m_run = bmax.clone()
        else:
            grow = (bmax - m_run) * c > 8.0
            fac = torch.where(grow, torch.exp2((m_run - bmax) * c), torch.ones_like(m_run))
            m_run = torch.where(grow, bmax, m_run)
            rsum *= fac
            O *= fac[:, None]
            rescales += int(grow.sum())
        p = torch.exp2((sb - m_run[:, None]) * c)
        pmax = max(pmax, float(p.max()))
        rsum += p.sum(dim=1)
        O += p @ v[j:j + blk]
    got = O / rsum[:, None]
    assert rescales > 0 and rescales < n_q * (n_k // blk - 1), "the case must mix rescaled and non-rescaled (row, block) pairs"
    assert pmax <= 2.0 ** 8 * (1 + 1e-12)
    assert (got - want).abs().max().item() < 1e-12
