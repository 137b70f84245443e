"""A small pass over every kernel family for `compute-sanitizer --tool memcheck` (one tool per gpurun call):
trunk + publication heads (frozen norms), a default-constructed model (updating BatchRMSNorm kernels), the JAX-aligned
heads through DNAModel.predict (junction head, embedding reuse) and a ragged-size forward (L = 2048 * 3, B = 2)."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from alphagenome_b200 import _lib, dna_model
from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome, publication_heads_config, set_update_running_var

rng = np.random.default_rng(0)


def tokens(B, L):
    return torch.from_numpy(rng.integers(0, 4, size=(B, L)).astype(np.int64)).cuda()


m = AlphaGenome()
for o in ("human", "mouse"):
    m.add_heads(**publication_heads_config[o])
m = m.cuda().eval()
set_update_running_var(m, False)
apply_synth_weights(m, 0)
L = 4096
seq = tokens(1, L)
idx = torch.sort(torch.randint(0, L, (1, 9), device="cuda"))[0]
out = m(seq, 0, splice_donor_idx=idx, splice_acceptor_idx=idx)
seq2 = tokens(2, 6144)
emb = m(seq2, torch.tensor([0, 1]).cuda(), return_embeds=True)
torch.cuda.synchronize()
print("frozen + heads ok", float(out["human"]["1bp_tracks"].sum()), float(emb.embeds_pair.sum()))

u = AlphaGenome().cuda().eval()          # BatchRMSNorm updating (default)
apply_synth_weights(u, 1)
e = u(tokens(2, 2048), 0)
torch.cuda.synchronize()
print("updating ok", float(e.embeds_128bp.sum()))

dm = dna_model.create(add_reference_heads=True, device="cuda", cuda_graphs=False, settings=dna_model.ModelSettings(num_splice_sites=8, splice_site_threshold=0.0))
apply_synth_weights(dm.model, 2)
s = "".join(np.array(list("ACGT"))[rng.integers(0, 4, size=2048)])
p = dm.predict(s, organism="human")
i = dm.ism(s, output_key="chip_tf", organism="human", positions=[3, 1000], batch_size=4)
torch.cuda.synchronize()
print("dna_model ok", float(p["splice_sites_junction"]["predictions"].sum()), i["delta"].shape, "launches", _lib.launch_count())
