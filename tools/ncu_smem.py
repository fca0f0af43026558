"""Shared-memory wavefronts per source line of one launch (ncu report with --import-source on; not part of the product).
usage: python tools/ncu_smem.py report.ncu-rep launch_index"""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
fname, hdr, rows = "", None, []
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
        wi, xi, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("L1 Wavefronts Shared Ideal")
    elif hdr is not None and len(r) > wi and r[0] != "":
        try:
            rows.append((int(r[wi] or 0), int(r[xi] or 0), int(r[ii] or 0), fname, int(r[0]), r[1].strip()[:90]))
        except ValueError:
            pass
tot = sum(r[0] for r in rows) or 1
print("total wavefronts", tot, "excessive", sum(r[1] for r in rows))
for r in sorted(rows, reverse=True)[:25]:
    if r[0]:
        print(f"{100 * r[0] / tot:5.1f}%  wavefronts {r[0]:>10} excessive {r[1]:>10} ideal {r[2]:>10}  {r[3]}:{r[4]}  {r[5]}")
