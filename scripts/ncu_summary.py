"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.  usage: ncu_summary.py rep out.txt"""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max"]
stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h]
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none summary of {rep}\n")
    for r in rows[2:]:
        f.write("---\n")
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                f.write(f"{k} = {r[i]} {units[i]}\n")
        st = sorted(((float(r[hdr.index(h)] or 0), h) for h in stall), reverse=True)[:6]
        f.write("top stalls (per issue): " + ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st) + "\n")
print(open(out).read())
