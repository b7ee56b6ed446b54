#!/usr/bin/env python
"""Turn an ncu report (.ncu-rep) into the small text summaries committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_map_fast16

writes <out>.metrics.txt (selected raw metrics per profiled launch) and <out>.hotspots.txt (stall samples by
instruction class and the top stalled SASS lines; needs -lineinfo / --import-source on)."""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out + ".metrics.txt", "w") as f:
        for r in rows[2:]:
            f.write("kernel: %s\n" % r[hdr.index("Kernel Name")])
            for i, h in enumerate(hdr):
                if h in KEYS or h.endswith("_per_issue_active.ratio"):
                    try:
                        if float(r[i]) == 0:
                            continue
                    except ValueError:
                        pass
                    f.write("  %-90s %s %s\n" % (h, r[i], units[i]))
            f.write("\n")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) < 3:
        return
    # one section per profiled launch: a title row (kernel name), a header row (starts with "Address"), then data rows
    sections, title = [], ""
    for r in rows:
        if r and r[0] == "Address":
            sections.append({"title": title, "hdr": r, "data": []})
        elif sections and len(r) >= len(sections[-1]["hdr"]):
            sections[-1]["data"].append(r)
        elif r:
            title = r[1] if len(r) > 1 else r[0]
    with open(out + ".hotspots.txt", "w") as f:
        for sec in sections:
            hdr, data = sec["hdr"], sec["data"]
            ix = {h: i for i, h in enumerate(hdr)}
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot, opc, ops = collections.Counter(), collections.Counter(), collections.Counter()
            per = []
            for r in data:
                n = int(r[ix["# Samples"]] or 0)
                ex = int(r[ix["Instructions Executed"]] or 0)
                for st in stalls:
                    tot[st] += int(r[ix[st]] or 0)
                t = r[ix["Source"]].split()
                op = (t[1] if t and t[0].startswith("@") else (t[0] if t else "")).split(".")[0]
                opc[op] += ex
                ops[op] += n
                per.append((n, r[ix["Source"]], ex))
            f.write(sec["title"] + "\n")
            f.write("warp-level instructions executed: %d\n" % sum(opc.values()))
            f.write("stall samples by reason: %s\n\n" % dict(tot.most_common()))
            f.write("instruction class: executed (warp-level), stall samples\n")
            for k, v in opc.most_common(24):
                f.write("  %-12s %12d %8d\n" % (k, v, ops[k]))
            f.write("\ntop stalled SASS lines: samples | executed | instruction\n")
            for n, sline, ex in sorted(per, key=lambda x: -x[0])[:16]:
                f.write("  %6d %10d  %s\n" % (n, ex, sline[:100]))
            f.write("\n")


if __name__ == "__main__":
    main()
