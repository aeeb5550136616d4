"""Top warp-stall-sample SASS lines of an .ncu-rep captured with --import-source on (read on the build box).
    python tools/ncu_top_stalls.py rep.ncu-rep [N]"""
import csv, io, subprocess, sys
rep, n = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
i_s, i_src, i_ex = h.index("Warp Stall Sampling (All Samples)"), h.index("Source"), h.index("Instructions Executed")
body = [(int(r[i_s]), k, r) for k, r in enumerate(rows[2:]) if len(r) > i_s and r[i_s].isdigit()]
tot = sum(b[0] for b in body)
print(f"{tot} samples")
for s, k, r in sorted(body, key=lambda b: -b[0])[:n]:
    print(f"{s / tot * 100:5.1f}%  line {k:5d}  exec {r[i_ex]:>8s}  {r[i_src].strip()[:100]}")
