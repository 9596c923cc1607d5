"""Summarise an .ncu-rep (raw + source pages) into the few numbers we track.  Usage: ncu_summary.py rep [ntop]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("==", d.get("Kernel Name", "")[:90], "grid", d.get("launch__grid_size"), "block", d.get("launch__block_size"), "regs", d.get("launch__registers_per_thread"))
    for k in ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]:
        if k in d: print(f"  {k:75s} {d[k]:>14s} {units[hdr.index(k)]}")
    st = {h.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[h]) for h in hdr if "issue_stalled" in h and "per_issue_active" in h}
    tot = sum(st.values())
    print("  stalls (% of warp time):", ", ".join(f"{k} {v/tot*100:.1f}" for k, v in sorted(st.items(), key=lambda x: -x[1]) if v/tot > 0.01))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; data = rows[2:]
ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
tot_ex = sum(int(r[iex]) for r in data); tot_s = sum(int(r[ismp]) for r in data)
cols = {n: hdr.index(n) for n in hdr if n.startswith("stall_") and "Not Issued" not in n}
print("  top stall sites:")
for r in sorted(data, key=lambda r: -int(r[ismp]))[:ntop]:
    print("   ", r[ia][-5:], f"{int(r[iex])/tot_ex*100:5.2f}% ex {int(r[ismp])/tot_s*100:5.2f}% smp", r[isrc][:58].ljust(58), {k[6:]: int(r[v]) for k, v in cols.items() if int(r[v]) > 200})
