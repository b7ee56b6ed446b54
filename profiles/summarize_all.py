#!/usr/bin/env python
"""Per-kernel table from an ncu report that holds many launches: for every distinct kernel the longest launch, with the
numbers the round's brief asks for (issue-slot and integer-pipe utilisation, achieved DRAM GB/s against the measured peak).

    python profiles/summarize_all.py gpurun_out/prof_all.ncu-rep profiles/r01_all_kernels.txt [hbm_peak_gbs]
"""
import csv
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def main():
    rep, out = sys.argv[1], sys.argv[2]
    peak = float(sys.argv[3]) if len(sys.argv) > 3 else 6551.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}
    best = {}
    for r in rows[2:]:
        name = r[ix["Kernel Name"]]
        t = num(r[ix["gpu__time_duration.sum"]]) * scale.get(units[ix["gpu__time_duration.sum"]], 1e-6)
        if name not in best or t > best[name][0]:
            best[name] = (t, r)
    cols = [("grid", "launch__grid_size"), ("block", "launch__block_size"), ("regs", "launch__registers_per_thread"),
            ("warps/SM", "sm__warps_active.avg.per_cycle_active"), ("issue%", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            ("alu%", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fma%", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"), ("L2hit%", "lts__t_sector_hit_rate.pct")]
    with open(out, "w") as f:
        f.write("# longest launch of every kernel in %s; DRAM GB/s = (dram__bytes_read.sum + dram__bytes_write.sum) / gpu__time_duration.sum, peak %.0f GB/s\n" % (rep, peak))
        f.write("# issue% = smsp__issue_active; alu% / fma% / lsu% = sm__inst_executed_pipe_* (pct of peak sustained active; the packed int16x2 ops run on\n")
        f.write("# two 16-lane pipes, see DESIGN 5.0: 100 % of alu / fma = 0.5 warp-instructions per clock and sub-partition)\n")
        f.write("%-62s %9s %8s %8s %7s" % ("kernel", "time us", "rd MB", "wr MB", "GB/s") + " %6s" % "%peak" + "".join(" %8s" % c for c, _ in cols) + "\n")
        for name, (t, r) in sorted(best.items(), key=lambda kv: -kv[1][0]):
            rd = num(r[ix["dram__bytes_read.sum"]]) * scale.get(units[ix["dram__bytes_read.sum"]], 1.0) if "dram__bytes_read.sum" in ix else float("nan")
            wr = num(r[ix["dram__bytes_write.sum"]]) * scale.get(units[ix["dram__bytes_write.sum"]], 1.0) if "dram__bytes_write.sum" in ix else float("nan")
            gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
            line = "%-62s %9.1f %8.1f %8.1f %7.0f %6.1f" % (name[:62], t * 1e6, rd / 1e6, wr / 1e6, gbs, 100.0 * gbs / peak)
            for _, k in cols:
                v = r[ix[k]] if k in ix else ""
                try:
                    line += " %8.1f" % num(v)
                except Exception:
                    line += " %8s" % v
            f.write(line + "\n")


if __name__ == "__main__":
    main()
