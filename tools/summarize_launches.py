"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals of ONE step
(the launches between the last two onehot_im2col_kernel launches) and, with --each, every launch of the step.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv [--each]"""
import csv
import sys
from collections import OrderedDict


def load(path):
    rows = list(csv.reader(open(path, errors="replace")))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h = i
            break
    hdr = rows[h]
    kn, mv, gs, un = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Unit")
    out = []
    for r in rows[h + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(r[un], 1e-6)
        out.append((r[kn], v * scale, r[gs]))
    return out


def main():
    L = load(sys.argv[1])
    idx = [i for i, (k, _, _) in enumerate(L) if "onehot_im2col" in k]
    step = L[idx[-2]:idx[-1]] if len(idx) >= 2 else L
    tot = sum(ms for _, ms, _ in step)
    agg = OrderedDict()
    for k, ms, _ in step:
        name = k.split("(")[0].replace("agb::", "").replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    print(f"launches in the step: {len(step)}; sum of kernel durations {tot:.2f} ms")
    print("| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {ms:.3f} | {100 * ms / tot:.1f}% |")
    if "--each" in sys.argv:
        for k, ms, g in step:
            print(f"{ms:9.4f} ms  {g:>16s}  {k[:90]}")


if __name__ == "__main__":
    main()
