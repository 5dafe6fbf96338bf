"""Attribute ncu SASS-page samples/instructions to CUDA source lines (needs nvdisasm -g output)."""
import collections
import csv
import re
import sys

sass_g, src_csv, src_file, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
off2line, cur = {}, None
for ln in open(sass_g):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = m.group(1).split('/')[-1], int(m.group(2)), m.group(3)
        mm = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
        outer = (mm[-1][0].split('/')[-1], int(mm[-1][1])) if mm else (f, l)
        cur = ((f, l), outer)
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*);', ln)
    if m and cur:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
iA, iSm, iE, iS = hdr.index('Address'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
base = int(data[0][iA], 16)
inner, outer, ops = collections.Counter(), collections.Counter(), collections.Counter()
inner_e, outer_e = collections.Counter(), collections.Counter()
tot = tot_e = 0
for r in data:
    off = int(r[iA], 16) - base
    s, e = int(r[iSm]), int(r[iE])
    (fi, fo) = off2line.get(off, (('?', 0), ('?', 0)))
    inner[fi] += s; inner_e[fi] += e; outer[fo] += s; outer_e[fo] += e
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+(\.MOV)?)', r[iS].strip())
    ops[m.group(2) if m else '?'] += e
    tot += s; tot_e += e
src = open(src_file).read().split('\n')
print(f"instructions per unit: {tot_e / units:.1f}; samples {tot}")
print("top opcodes:", ", ".join(f"{k} {v / units:.0f}" for k, v in ops.most_common(18)))
for title, c, ce in (("by outermost line", outer, outer_e), ("by innermost line", inner, inner_e)):
    print("--", title)
    for (f, l), s in sorted(c.items(), key=lambda kv: -kv[1])[:22]:
        text = src[l - 1].strip()[:84] if src_file.endswith(f) and l > 0 else f
        print(f"{100 * s / tot:5.1f}%  {ce[(f, l)] / units:7.1f}  {f}:{l}: {text}")
