import csv, subprocess, sys
KEYS = ["gpu__time_duration.sum","launch__registers_per_thread","smsp__inst_executed.sum","smsp__issue_active.avg.pct_of_peak_sustained_active",
"sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
"sm__warps_active.avg.pct_of_peak_sustained_active","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"lts__t_sector_hit_rate.pct","l1tex__t_sector_hit_rate.pct","sm__icc_request_hit_rate.pct","l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum","l1tex__data_pipe_lsu_wavefronts.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
"lts__throughput.avg.pct_of_peak_sustained_elapsed","l1tex__m_xbar2l1tex_read_bytes.sum","l1tex__m_l1tex2xbar_write_bytes.sum","l1tex__lsuin_requests.avg.pct_of_peak_sustained_elapsed",
"l1tex__throughput.avg.pct_of_peak_sustained_elapsed","sm__throughput.avg.pct_of_peak_sustained_elapsed","smsp__inst_issued.sum","sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active","smsp__warps_eligible.avg.per_cycle_active","smsp__inst_executed_op_shared_ld.sum","smsp__inst_executed_op_global_st.sum"]
cols = []
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
    rows = list(csv.reader(raw.splitlines())); hdr=rows[0]; r=rows[2]
    d = {h:r[i] for i,h in enumerate(hdr)}
    cols.append(d)
keys = [k for k in cols[0] if k in KEYS or k.endswith("_per_issue_active.ratio")]
print("%-86s" % "metric" + "".join("%16s" % a.split('/')[-1][:15] for a in sys.argv[1:]))
for k in keys:
    vals = []
    for d in cols:
        v = d.get(k,"")
        try: v = "%.3f" % float(v.replace(",",""))
        except: pass
        vals.append(v)
    if all(v in ("0.000","") for v in vals): continue
    print("%-86s" % k[:86] + "".join("%16s" % v for v in vals))
