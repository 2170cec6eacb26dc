"""Executed warp instructions per source line of one kernel launch in an .ncu-rep.
Usage: python tools/ncu_inst.py rep kernel_regex [launch_index] [top]"""
import csv, io, os, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                               "--kernel-name", "regex:" + kern], text=True, stderr=subprocess.DEVNULL)
rows = list(csv.reader(io.StringIO(raw)))
launch, hdr, seen, fname, agg, src = -1, None, set(), None, {}, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        if r[1] in seen or launch < 0:
            launch += 1
            seen = set()
        seen.add(r[1]); fname = os.path.basename(r[1]); continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or launch != which or not r[0].isdigit():
        continue
    c = {h: i for i, h in enumerate(hdr)}
    col = "# Warp Instructions Executed" if "# Warp Instructions Executed" in c else "Instructions Executed"
    try:
        v = float(r[c[col]] or 0)
    except (ValueError, IndexError):
        continue
    k = (fname, int(r[0]))
    agg[k] = agg.get(k, 0) + v
tot = sum(agg.values()) or 1
lines = {}
for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
    if f not in lines:
        for base in ("voxelengine-cpp_b200/csrc/", ""):
            try:
                lines[f] = open(base + f).read().split("\n"); break
            except OSError:
                lines[f] = []
    text = lines[f][n - 1].strip()[:100] if 0 < n <= len(lines[f]) else ""
    print("%5.1f%%  %s:%d  %s" % (100 * v / tot, f, n, text))
