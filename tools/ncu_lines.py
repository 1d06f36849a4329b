"""Per-source-line instruction / stall-sample shares of one kernel launch in an .ncu-rep.
Usage: python tools/ncu_lines.py rep kernel_regex launch_skip [top]"""
import csv
import subprocess
import sys

rep, rx, skip = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                      f"regex:{rx}", "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
data, fname = [], ""
hdr = None
for r in rows:
    if r and r[0] in ("File Path", "File Name"):
        fname = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) == len(hdr) and r[0].isdigit():
        ci, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        data.append((int(r[ci]), int(r[si]), fname, r[0], r[1].strip()[:100]))
ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
print(f"# {rep} {rx} skip={skip}: {ti} warp instructions, {ts} samples")
for d in sorted(data, reverse=True)[:top]:
    print(f"{100 * d[0] / ti:5.1f}% inst {100 * d[1] / max(ts, 1):5.1f}% smp  {d[2]}:{d[3]}: {d[4]}")
