"""Summarise an ncu report: headline metrics (raw page) and the top stalled SASS instructions (source page).
usage: python tools/ncu_src_summary.py report.ncu-rep [pattern ...]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
pats = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "sm__cycles_elapsed.max"]
for r in rows[2:]:
    print("---- launch")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print("  %-70s %s %s" % (w, r[i], units[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# first kernel only
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot, top = {}, []
for r in rows[2:]:
    if len(r) < len(hdr):
        if r and r[0] == "Kernel Name":
            break
        continue
    try:
        n = int(r[ix["# Samples"]])
    except ValueError:
        continue
    d = {s: int(r[ix[s]] or 0) for s in stalls}
    for s, v in d.items():
        tot[s] = tot.get(s, 0) + v
    top.append((n, r[ix["Address"]][-5:], r[ix["Source"]][:80], {k: v for k, v in d.items() if v > 0}, r[ix["Instructions Executed"]]))
print("stall totals:", sorted(tot.items(), key=lambda x: -x[1])[:12])
for t in sorted(top, key=lambda x: -x[0])[:30]:
    print(t)
for p in pats:
    print("---- instructions matching", p)
    for t in top:
        if p in t[2]:
            print(t)
