"""SASS opcode histogram per kernel of libagb200.so (what proves tcgen05 / TMEM / TMA in the shipped binary):

    python tools/sass_histogram.py > profiles/sass_r03_histogram.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "alphagenome_b200" / "libagb200.so"
KEY = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "UTCATOMSWS", "SYNCS", "MUFU.TANH", "MUFU.EX2", "MUFU.RCP", "MUFU.RSQ", "HMMA",
       "LDG", "STG", "LDS", "STS", "FFMA", "DADD", "RED", "ATOM", "BAR", "ACQBULK", "UCGABAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    arch = set()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.search(r"arch = (sm_\w+)", line)
        if m:
            arch.add(m.group(1))
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS opcode histogram of alphagenome_b200/libagb200.so ({len(funcs)} functions, arch {sorted(arch)})\n")
    print("Counts of instructions whose mnemonic starts with the column's prefix (`cuobjdump -sass`).  `UTCHMMA` = tcgen05.mma, "
          "`LDTM`/`STTM` = tcgen05.ld/st, `UTMALDG` = TMA tile load, `UTCBAR` = tcgen05.commit, `HMMA` = the legacy warp MMA (`mma.sync`): used by `pair_bias_mma_kernel` only (a 128 -> 8 projection, DESIGN.md section 4), absent from every GEMM / attention kernel.\n")
    print("| kernel | instrs | " + " | ".join(KEY) + " |")
    print("|---|---:|" + "---:|" * len(KEY))
    total = collections.Counter()
    for (name, c), dn in zip(funcs.items(), demangle):
        short = re.sub(r"\(.*", "", dn).replace("void agb::", "").replace("agb::", "")
        row = [sum(v for k, v in c.items() if k.startswith(p)) for p in KEY]
        for p, v in zip(KEY, row):
            total[p] += v
        if sum(c.values()) == 0:
            continue
        print(f"| `{short}` | {sum(c.values())} | " + " | ".join(str(v) if v else "" for v in row) + " |")
    print("| **total** | | " + " | ".join(str(total[p]) for p in KEY) + " |")


if __name__ == "__main__":
    main()
