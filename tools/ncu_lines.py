"""Per-source-line stall samples of one kernel in an .ncu-rep (read here, no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof_tc.ncu-rep bwd_tc [min_samples]
"""
import csv
import io
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 150
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = None
fname = ""
tot = 0
out = []
for r in rows:
    if r and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        stalls = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or not r or not r[0].isdigit():
        continue
    try:
        s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    tot += s
    if s >= thr:
        st = {h[6:]: int(r[i]) for h, i in stalls if r[i].isdigit() and int(r[i]) > s * 0.1}
        out.append((s, fname, r[0], r[1].strip()[:100], r[ix["Instructions Executed"]], st))
print("total samples", tot)
for o in sorted(out, key=lambda o: (o[1], int(o[2]))):
    print("%6d  %s:%s  %-100s inst=%s %s" % o)
