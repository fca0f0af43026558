"""Per-source-line stall samples and executed instructions of one launch of an ncu report taken with --import-source on
(not part of the product).  usage: python tools/ncu_lines.py report.ncu-rep launch_index [top_n]"""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
fname = ""
lines = {}
hdr = None
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) <= ie or r[0] == "":
        continue
    try:
        lines[(fname, int(r[0]))] = (int(r[si] or 0), int(r[ie] or 0), r[1].strip()[:100])
    except ValueError:
        pass
tot = sum(v[0] for v in lines.values()) or 1
toti = sum(v[1] for v in lines.values()) or 1
print(f"samples {tot}  warp instructions {toti}")
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * v[0] / tot:5.1f}% smp {100 * v[1] / toti:5.1f}% inst  {f}:{ln}  {v[2]}")
