#!/bin/bash
# Same-box A/B of the round-3 kernel changes on the 1 Mbp step (CUDA-event ms per graph replay, tools/sp_breakdown.py):
# "old" = fp32 FMA to_attn_bias + one-tile-per-CTA two-pass row attention + fp32 x0 residual; "new" = the defaults.
# Interleaved so that box drift shows up as a difference between the two rounds, not between the arms.
cd "${GRAFT_REPO_ROOT:-.}"
for round in 1 2; do
  AGB_ROWATTN=0 AGB_PAIR_BIAS_MMA=0 AGB_EMBED_FUSE=0 timeout 200 python tools/sp_breakdown.py --out r3_ab_old_$round.json 2>&1 | grep SP_BREAKDOWN | cut -c1-200
  timeout 200 python tools/sp_breakdown.py --out r3_ab_new_$round.json 2>&1 | grep SP_BREAKDOWN | cut -c1-200
done
