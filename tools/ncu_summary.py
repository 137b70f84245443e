"""Key metrics of every launch in an .ncu-rep (`ncu --set full` capture) as a markdown table.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--csv profiles/x.csv] [--traffic-json profiles/gemm_traffic.json --note "..."]

--csv            also write the raw metric table (`ncu --page raw --csv`) next to the summary (tracked evidence)
--traffic-json   write {dram_bytes_per_launch, duration_us, kernel, source, note} of launch 0: what bench.py reports as
                 roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant GEMM shape)"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "time"),
    ("sm__cycles_elapsed.avg.per_second", "SM clk"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU(MUFU) %"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue %"),
    ("dram__bytes_read.sum", "DRAM rd"),
    ("dram__bytes_write.sum", "DRAM wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 rd sectors (tex)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_sb %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
]


def main():
    out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    names = [w for w in WANT if w[0] in col]
    print("| # | kernel | " + " | ".join(n for _, n in names) + " |")
    print("|---|---|" + "---|" * len(names))
    for k, r in enumerate(rows[2:]):
        cells = []
        for m, _ in names:
            i = col[m]
            cells.append(f"{r[i]} {units[i]}".strip())
        print(f"| {k} | `{r[col['Kernel Name']][:38]}` | " + " | ".join(cells) + " |")
    def to_bytes(val, unit):
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        return float(val) * mult.get(unit, 1)

    if "--csv" in sys.argv:
        open(sys.argv[sys.argv.index("--csv") + 1], "w").write(out)
    if "--traffic-json" in sys.argv:
        import json

        r = rows[2]
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        dur = r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]
        note = sys.argv[sys.argv.index("--note") + 1] if "--note" in sys.argv else ""
        json.dump(dict(dram_bytes_per_launch=rd + wr, dram_bytes_read=rd, dram_bytes_write=wr, duration=dur, kernel=r[col["Kernel Name"]],
                       source=sys.argv[1].split("/")[-1], note=note), open(sys.argv[sys.argv.index("--traffic-json") + 1], "w"), indent=1)
    if "--list" in sys.argv:
        for h in hdr:
            print(h)


if __name__ == "__main__":
    main()
