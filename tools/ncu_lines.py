"""Per-source-line warp-stall samples and executed instructions of one kernel
from an .ncu-rep (needs -lineinfo + --import-source on).
Usage: python tools/ncu_lines.py rep kernel_regex [launch_index] [top]"""
import csv, io, subprocess, sys, os
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                               "--kernel-name", "regex:" + kern], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
# launches are separated by a repeated first "File Path"; detect by function-name rows order
launch, fname, hdr, seen_files = -1, None, None, set()
out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path":
        if r[1] in seen_files or launch < 0:
            launch += 1
            seen_files = set()
        seen_files.add(r[1]); fname = os.path.basename(r[1]); continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if launch != which or not r[0].isdigit(): continue
    c = {h: i for i, h in enumerate(hdr)}
    def num(k):
        v = r[c[k]].replace(",", "")
        return float(v) if v not in ("", "-") else 0.0
    try:
        out.append((num("# Samples"), num("Instructions Executed"), fname, int(r[0]), r[1].strip()[:90]))
    except (ValueError, IndexError):
        pass  # a source line with unescaped quotes
ts = sum(o[0] for o in out); ti = sum(o[1] for o in out)
print("launch %d: %d source lines, %d samples, %.0f warp instructions" % (which, len(out), ts, ti))
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100 * s / max(ts, 1), 100 * i / max(ti, 1), f, ln, src))
