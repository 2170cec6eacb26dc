"""Summarises an .ncu-rep (ncu -i ... --page raw --csv) into a small table:
per captured launch the duration, DRAM bytes, throughput %, occupancy,
registers and top stall reasons.
Usage: python tools/ncu_summary.py rep [out.md] [--traffic workload]
--traffic also (re)writes profiles/ncu_traffic.json — dram__bytes_read.sum + dram__bytes_write.sum
per launch of every captured kernel class, stamped with the commit it was captured at — which is
what bench.py reports as roofline.traffic for that workload."""
import csv, io, json, os, re, subprocess, sys

argv = [a for a in sys.argv[1:]]
traffic_for = None
if "--traffic" in argv:
    i = argv.index("--traffic")
    traffic_for = argv[i + 1]
    del argv[i:i + 2]
sys.argv = [sys.argv[0]] + argv
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

if traffic_for:
    def num(x):
        return float(x.replace(",", "")) if x else 0.0
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    kern = {}
    for r in data:
        full = r[col["Kernel Name"]]
        name = re.search(r"(k_[a-z_0-9]+)", full).group(1)
        if name == "k_solve_async":
            name = "k_solve"
        if name == "k_mesh_emit_cubes":  # the profile slot both cube passes are timed under
            name = "k_mesh_emit"
        rd = num(r[col["dram__bytes_read.sum"]]) * mult.get(units[col["dram__bytes_read.sum"]], 1)
        wr = num(r[col["dram__bytes_write.sum"]]) * mult.get(units[col["dram__bytes_write.sum"]], 1)
        dur = num(r[col["gpu__time_duration.sum"]])
        k = kern.setdefault(name, {"dram_bytes": 0, "launches": 0, "dur": 0.0})
        k["dram_bytes"] += rd + wr
        k["launches"] += 1
        k["dur"] += dur
    for k in kern.values():
        k["dram_bytes"] = int(k["dram_bytes"] / k["launches"])
        k["dur_" + units[col["gpu__time_duration.sum"]]] = round(k.pop("dur") / k["launches"], 2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    commit = subprocess.check_output(["git", "-C", root, "rev-parse", "--short", "HEAD"], text=True).strip()
    json.dump({"workload": traffic_for, "commit": commit, "report": os.path.basename(rep),
               "what": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, --clock-control none)",
               "kernels": kern}, open(os.path.join(root, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print("wrote profiles/ncu_traffic.json for", traffic_for, "at", commit)
