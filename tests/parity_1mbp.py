"""End-to-end oracle parity at the HEADLINE configuration: B=1 x 1,048,576 bp (BASELINE config 4), trunk + output
embeddings, optionally + the publication heads of both organisms.  Test infrastructure (not collected by pytest: it
needs ~60 GB of host RAM and minutes of CPU); run on the GPU box and commit the report:

    python -m tests.parity_1mbp [--heads] [--out gpurun_out/parity_1mbp.json]

The B200 forward and the pinned CPU oracle (oracle/trunk_oracle.py) run on the same synthetic weights (seed 0) and the
same tokens as bench.py's workload; every output is compared over ALL of its elements (max|d| / max|ref| and Pearson,
accumulated in fp64 over row chunks).  north_star tolerance: 1e-2 on the three embeddings, Pearson >= 0.999 on heads.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class Acc:
    """Streaming max|d|, max|ref| and Pearson over chunks."""

    def __init__(self):
        self.maxd = self.maxr = 0.0
        self.n = 0
        self.sa = self.sb = self.saa = self.sbb = self.sab = 0.0

    def add(self, got: torch.Tensor, want: torch.Tensor):
        g, w = got.double().flatten(), want.double().flatten()
        self.maxd = max(self.maxd, float((g - w).abs().max()))
        self.maxr = max(self.maxr, float(w.abs().max()))
        self.n += g.numel()
        self.sa += float(g.sum()); self.sb += float(w.sum())
        self.saa += float(g @ g); self.sbb += float(w @ w); self.sab += float(g @ w)

    def report(self):
        n = self.n
        cov = self.sab - self.sa * self.sb / n
        va, vb = self.saa - self.sa ** 2 / n, self.sbb - self.sb ** 2 / n
        return dict(max_rel=self.maxd / max(self.maxr, 1e-30), pearson=cov / max((va * vb) ** 0.5, 1e-300), elements=n)


def compare(got: torch.Tensor, want: torch.Tensor, chunk_rows: int = 16384):
    acc = Acc()
    g = got.reshape(-1, got.shape[-1])
    w = want.reshape(-1, want.shape[-1])
    for r in range(0, g.shape[0], chunk_rows):
        acc.add(g[r:r + chunk_rows].float().cpu(), w[r:r + chunk_rows])
    return acc.report()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=1_048_576)
    ap.add_argument("--heads", action="store_true")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "parity_1mbp.json"))
    args = ap.parse_args()
    import psutil

    ram_gb = psutil.virtual_memory().available / 2 ** 30
    need = 96 * args.length / 1_048_576
    rep = dict(length_bp=args.length, host_cores=os.cpu_count(), host_ram_available_gb=round(ram_gb, 1))
    if ram_gb < need:
        rep["skipped"] = f"needs ~{need:.0f} GB of host RAM"
        print("PARITY_1MBP " + json.dumps(rep))
        return

    from alphagenome_b200.init_weights import apply_synth_weights
    from alphagenome_b200.model import AlphaGenome, publication_heads_config, set_update_running_var
    from bench import N_SITES, synth_sites, synth_tokens
    from oracle import trunk_oracle as O

    L = args.length
    model = AlphaGenome()
    if args.heads:
        for organism in ("human", "mouse"):
            model.add_heads(**publication_heads_config[organism])
    model = model.cuda().eval()
    set_update_running_var(model, False)
    apply_synth_weights(model, seed=0)
    seq = synth_tokens(1, L)
    org = torch.zeros(1, dtype=torch.long)
    sd = {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
    emb = model(seq.cuda(), org.cuda(), return_embeds=True)
    torch.cuda.synchronize()
    heads = None
    if args.heads:
        donor, acceptor = synth_sites(1, L)
        emb = type(emb)(*[t.clone() for t in emb])
        heads = model(seq.cuda(), org.cuda(), splice_donor_idx=donor.cuda(), splice_acceptor_idx=acceptor.cuda())
        torch.cuda.synchronize()

    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.perf_counter()
    ref = O.alphagenome_embeds(sd, seq, org)
    rep["oracle_seconds"] = round(time.perf_counter() - t0, 1)
    rep["oracle_bp_per_s"] = L / (time.perf_counter() - t0)
    for name, got in zip(("embeds_1bp", "embeds_128bp", "embeds_pair"), emb):
        rep[name] = compare(got, ref[name])
    ok = all(rep[k]["max_rel"] <= 1e-2 and rep[k]["pearson"] >= 0.999 for k in ("embeds_1bp", "embeds_128bp", "embeds_pair"))

    if heads is not None:
        hr = {}
        e1, e128, epair = ref["embeds_1bp"], ref["embeds_128bp"], ref["embeds_pair"]
        for organism in ("human", "mouse"):
            hp = f"heads.{organism}."
            # position-wise heads on embeds_1bp, in row chunks (the 1 bp human tracks alone are 13 GB in fp32)
            accs = {n: Acc() for n in ("1bp_tracks", "splice_logits", "splice_usage")}
            for r in range(0, L, 32768):
                x = e1[:, r:r + 32768]
                want = {"1bp_tracks": O.tracks_scaled_prediction(sd, hp + "1bp_tracks.", x),
                        "splice_logits": O.linear(sd, hp + "splice_logits.linear.", x),
                        "splice_usage": torch.sigmoid(O.linear(sd, hp + "splice_usage.linear.", x))}
                for n, w in want.items():
                    accs[n].add(heads[organism][n][:, r:r + 32768].float().cpu(), w)
            for n, a in accs.items():
                hr[f"{organism}/{n}"] = a.report()
            hr[f"{organism}/128bp_tracks"] = compare(heads[organism]["128bp_tracks"], O.tracks_scaled_prediction(sd, hp + "128bp_tracks.", e128))
            hr[f"{organism}/contact_head"] = compare(heads[organism]["contact_head"], O.contact_map_head(sd, hp + "contact_head.", epair, org))
            # splice junctions only touch the gathered site rows: project(x)[idx] == project(x[idx])
            want = O.splice_junction_head(sd, hp + "splice_juncs.", e1, donor, acceptor)
            hr[f"{organism}/splice_juncs"] = compare(heads[organism]["splice_juncs"], want)
        rep["heads"] = hr
        rep["heads_worst_pearson"] = min(v["pearson"] for v in hr.values())
        rep["heads_worst_max_rel"] = max((v["max_rel"], k) for k, v in hr.items())
        ok = ok and rep["heads_worst_pearson"] >= 0.999
        rep["n_sites"] = N_SITES
    rep["ok"] = bool(ok)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(rep, indent=1))
    print("PARITY_1MBP " + json.dumps(rep))
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
