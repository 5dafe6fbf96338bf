"""profiles/r2_sass_evidence.txt: which of the shipped kernels carry bulk copies (TMA: UBLKCP), mbarrier operations
(SYNCS.*) and register reallocation between warpgroups (USETMAXREG) -- counted in `cuobjdump -sass` of the built library.
usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "opensc2_b200", "libopensc2_b200.so")], capture_output=True, text=True).stdout
cur, cnt, first = None, collections.defaultdict(collections.Counter), collections.defaultdict(dict)
pat = re.compile(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);")
keys = ("UBLKCP", "SYNCS", "USETMAXREG", "UTMALDG", "LDGSTS", "BAR.SYNC", "ELECT", "FENCE.VIEW.ASYNC")
for line in sass.splitlines():
    if "Function :" in line:
        cur = line.split("Function :")[1].strip()
        continue
    m = pat.match(line)
    if not m or cur is None:
        continue
    t = re.sub(r"^@!?U?P\d+\s+", "", m.group(2).strip())
    op = t.split()[0]
    for k in keys:
        if op.startswith(k):
            kk = op if k in ("SYNCS", "USETMAXREG", "UBLKCP") else k
            cnt[cur][kk] += 1
            first[cur].setdefault(kk, f"/*{m.group(1)}*/ {t}")
names = subprocess.run(["c++filt"], input="\n".join(cnt.keys()), capture_output=True, text=True).stdout.split("\n")
print("SASS evidence from the shipped opensc2_b200/libopensc2_b200.so (cuobjdump -sass, sm_100a): bulk copies (TMA; UBLKCP.S.G =\n"
      "global -> shared, UBLKCP.G.S = shared -> global), mbarrier operations (SYNCS.*), register reallocation between warpgroups\n"
      "(USETMAXREG): instruction counts per kernel and the first occurrence of each.  tools/sass_evidence.py\n")
for mangled, dem in zip(cnt.keys(), names):
    if not any(x in dem for x in ("forward_split", "forward_tile", "backward_tile", "forward64")):
        continue
    print(dem[:150])
    for k, v in sorted(cnt[mangled].items()):
        print("    %-34s x%-3d  first: %s" % (k, v, first[mangled][k][:110]))
    print()
