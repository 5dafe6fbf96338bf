"""profiles/ncu_traffic.json: DRAM bytes (read + write) per unit of every kernel family, from one `ncu --set full`
report taken at bench size, stamped with the hash of opensc2_b200/csrc at the time -- bench.py quotes
`roofline.traffic` from it and drops the figure when the kernels have changed since (VERDICT r1 hygiene 10).
usage: python tools/ncu_traffic.py report.ncu-rep <units per launch, e.g. 4096*10001>"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_hash  # noqa: E402

FAMILY = (("osc2_forward", "forward"), ("osc2_backward", "backward"), ("osc2_gauss", "gauss"), ("osc2_opcond", "props"),
          ("osc2_norm", "norms"))
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, units = sys.argv[1], float(eval(sys.argv[2]))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, un, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out, seen = {}, set()
    for r in data:
        name = r[col["Kernel Name"]]
        fam = next((f for key, f in FAMILY if key in name), None)
        short = name.split("(")[0]
        if fam is None or short in seen:          # one launch of every distinct kernel; a family may have several kernels
            continue
        seen.add(short)
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[col[key]].replace(",", "")) * TO_BYTES.get(un[col[key]], 1.0)
        ms = float(r[col["gpu__time_duration.sum"]].replace(",", ""))
        e = out.setdefault(fam, {"name": "", "dram_bytes_per_unit": 0.0, "duration_ms_under_ncu": 0.0})
        e["name"] = (e["name"] + " + " if e["name"] else "") + short
        e["dram_bytes_per_unit"] += tot / units
        e["duration_ms_under_ncu"] += ms
    j = {"report": os.path.basename(rep), "units_per_launch": units, "csrc_sha1": csrc_hash(), "kernels": out}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(j, f, indent=1)
    print(json.dumps(j, indent=1))


if __name__ == "__main__":
    main()
