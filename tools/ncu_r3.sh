#!/bin/bash
# Round-3 ncu captures on one GPU (run under gpurun): launch list of one eager 1 Mbp forward, then --set full captures
# of the kernels that changed this round.  Reports land in gpurun_out/ncu3/ and are digested by tools/ncu_digest.py.
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/ncu3
O=gpurun_out/ncu3
# launch list of the whole run (one-time weight packing + three eager forwards; the last forward is the last 184 rows)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
  --log-file $O/launches_r03_all.csv python tools/ncu_forward.py > $O/launches.log 2>&1
cap() {  # name, kernel regex, launch-skip
  timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" --launch-skip $3 --launch-count 1 -f \
    -o $O/ncu_r03_$1 python tools/ncu_forward.py > $O/$1.log 2>&1
  tail -1 $O/$1.log
  # the reports are ~30 MB each (imported sources): digest them here, only the CSV / hot-line text travels back
  python tools/ncu_digest.py $O/ncu_r03_$1.ncu-rep $O/ncu_r03_$1 && rm -f $O/ncu_r03_$1.ncu-rep
}
# gemm launches: forward 1 issues 119 + the cached relative-position encodings, forwards 2 and 3 issue 119; the skips below
# are approximate — name a capture after the kernel / duration / DRAM bytes it actually holds (the run that produced
# profiles/ncu_r03_* caught outembed_1bp and the K = 64 embed GEMM as intended and a width-5 conv_out at 2 bp for 357)
cap outembed_1bp gemm_conv 361
cap w1conv_768_1bp gemm_conv 357
cap embed_k64 gemm_conv 243
ls -la $O
