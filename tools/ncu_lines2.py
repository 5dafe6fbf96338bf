"""Samples of an ncu report aggregated by CUDA source line (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines2.py report.ncu-rep kernel_regex [units_per_launch] [top]"""
import collections
import csv
import io
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 else None
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr, agg, tot = None, None, collections.OrderedDict(), 0
stall_cols = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
        continue
    if hdr is None or r[0] == "":
        continue
    try:
        s, ie = int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    st = {}
    for i, h in stall_cols:
        try:
            v = int(r[i])
        except ValueError:
            v = 0
        if v:
            st[h[6:]] = v
    agg[(cur, int(r[0]))] = (s, ie, r[1].strip()[:100], st)
    tot += s
print("total samples", tot)
allst = collections.Counter()
for (_k), (s, ie, src, st) in agg.items():
    allst.update(st)
print("stall reasons:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in allst.most_common(10)))
for (f, l), (s, ie, src, st) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    per = f"{ie / units:7.1f}/unit" if units else f"{ie:12d}"
    main = ",".join(f"{k}:{100 * v / max(s, 1):.0f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{l:4d} {100 * s / tot:5.1f}% instr {per}  [{main}]  {src}")
