"""Digest one `ncu --set full --import-source on` report into small text artefacts (the .ncu-rep files are 4-14 MB each
and do not travel back from the GPU box in bulk):

    python tools/ncu_digest.py gpurun_out/x.ncu-rep profiles/ncu_r02_x      # -> .raw.csv (every metric of the launch),
                                                                             #    .hot.txt (top SASS lines by stall samples)
"""
import csv
import io
import subprocess
import sys


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(out + ".raw.csv", "w").write(raw)
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hdr = rows[hdr_i]
    si, so = hdr.index("# Samples"), hdr.index("Source")
    data = [(int(r[si] or 0), i, r[so].strip()) for i, r in enumerate(rows[hdr_i + 1:]) if len(r) > si]
    tot = sum(d[0] for d in data) or 1
    top = sorted(sorted(data, key=lambda d: -d[0])[:32], key=lambda d: d[1])
    with open(out + ".hot.txt", "w") as f:
        f.write(f"{rows[0][1] if len(rows[0]) > 1 else ''}\ntotal warp-stall samples {tot}; top SASS lines (samples, share, line, instruction)\n")
        for n, i, s in top:
            f.write(f"{n:8d} {100 * n / tot:5.1f}%  {i:5d}  {s[:120]}\n")


if __name__ == "__main__":
    main()
