"""Summarises an .ncu-rep (ncu -i ... --page raw --csv) into a small table:
per captured launch the duration, DRAM bytes, throughput %, occupancy,
registers and top stall reasons.  Usage: python tools/ncu_summary.py rep [out.md]"""
import csv, io, re, subprocess, sys

rep = sys.argv[1]
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("smsp__inst_executed.sum", "inst"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
out = []
for r in data:
    name = re.search(r"(k_[a-z_0-9]+)", r[col["Kernel Name"]]).group(1)
    line = ["%-14s grid=%s" % (name, r[col["Grid Size"]])]
    for m, lab in want:
        if m in col:
            line.append("%s=%s%s" % (lab, r[col[m]], units[col[m]] if lab in ("dur", "dram_rd", "dram_wr") else ""))
    st = sorted(((float(r[col[h]].replace(",", "")) if r[col[h]] else 0.0, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall), reverse=True)[:4]
    line.append("stalls: " + ", ".join("%s=%.2f" % (n, v) for v, n in st))
    out.append("  ".join(line))
text = "\n".join(out)
print(text)
if len(sys.argv) > 2:
    open(sys.argv[2], "w").write("```\n" + text + "\n```\n")
