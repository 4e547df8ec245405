"""Per-source-line aggregation of an .ncu-rep captured with --import-source on.
usage: ncu_hot.py <report.ncu-rep> [top] [launch-id]
Reads `ncu -i rep --page source --csv --print-source cuda,sass`: each CUDA line is followed by its
SASS rows; samples / executed instructions / stall reasons are summed per CUDA line."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if len(sys.argv) > 3:
    cmd += ["--launch-skip", sys.argv[3], "--launch-count", "1"]
txt = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = collections.defaultdict(lambda: [0] * 8)
text = {}
fname, hdr, cur = None, None, None
STALLS = ["stall_long_sb", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_not_selected"]
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ci = {n: hdr.index(n) for n in ["# Samples", "Instructions Executed"] + STALLS}
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0]:                            # a CUDA line: carries the totals of its SASS rows
        cur = (fname, int(r[0]))
        text[cur] = r[1].strip()
        v = agg[cur]
        for j, n in enumerate(["# Samples", "Instructions Executed"] + STALLS):
            try:
                v[j] += int(r[ci[n]] or 0)
            except ValueError:
                pass
# an instruction inlined from a header is listed under the header line AND under its call sites,
# so percentages are relative to the file with the largest total (the one holding the kernel body)
per_file = collections.defaultdict(lambda: [0, 0])
for k, v in agg.items():
    per_file[k[0]][0] += v[0]
    per_file[k[0]][1] += v[1]
tot = max(v[0] for v in per_file.values()) or 1
totx = max(v[1] for v in per_file.values()) or 1
print(f"kernel total: samples {tot}  warp-instructions {totx}")
for f, v in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:32s} samples {100*v[0]/tot:5.1f}%  instructions {100*v[1]/totx:5.1f}%")
print("file:line                      samp%  inst%   long   wait  short branch   math notsel | source")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0][:22]:22s}:{k[1]:<5d} {100*v[0]/tot:6.1f} {100*v[1]/totx:6.1f} {v[2]:6d} {v[3]:6d} {v[4]:6d} {v[5]:6d} "
          f"{v[6]:6d} {v[7]:6d} | {text.get(k, '')[:70]}")
