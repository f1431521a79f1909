"""Top stall lines of one launch of an ncu report, by CUDA source line and by SASS instruction.
usage: python tools/ncu_hot_lines.py report.ncu-rep [launch_index] [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the export is a sequence of per-file blocks: "File Path", "Function Name", header, lines...
cur_file = None
hdr = None
by_line = defaultdict(lambda: [0, 0, defaultdict(int)])
sass = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        col = {}
        for i, h in enumerate(hdr):
            col.setdefault(h, i)
        continue
    if hdr is None:
        continue
    try:
        line = int(r[col["Line No"]])
    except ValueError:
        continue
    def num(x):
        try:
            return int(float(x.replace(",", "")))
        except ValueError:
            return 0
    samp = num(r[col["# Samples"]])
    inst = num(r[col["Instructions Executed"]])
    key = (cur_file, line)
    by_line[key][0] += samp
    by_line[key][1] += inst
    by_line[key][2]["src"] = r[col["Source"]][:110]
    for h, i in col.items():
        if h.startswith("stall_") and "Not Issued" not in h and r[i] not in ("", "0"):
            by_line[key][2][h] += num(r[i])
    sass.append((samp, cur_file, line, r[3][:90] if len(r) > 3 else ""))
tot = sum(v[0] for v in by_line.values()) or 1
print("total samples", tot)
for (f, l), v in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(((c, h[6:]) for h, c in v[2].items() if h != "src"), reverse=True)[:3]
    print("%5.1f%% %9d inst  %s:%d  %s   | %s" % (100.0 * v[0] / tot, v[1], f, l, ", ".join("%s %d" % (h, c) for c, h in st), v[2]["src"].strip()))
print("---- top SASS instructions")
for samp, f, l, txt in sorted(sass, reverse=True)[:top]:
    print("%5.1f%%  %s:%d  %s" % (100.0 * samp / tot, f, l, txt))
