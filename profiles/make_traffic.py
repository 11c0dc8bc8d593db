#!/usr/bin/env python
"""DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels in one or more
`ncu --set full` reports -> JSON fragment for profiles/r1_traffic.json.
Usage: make_traffic.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import json
import re
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main(paths):
    out = {}
    for path in paths:
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        for r in rows[2:]:
            name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("b200::", "").strip()
            name = name.replace("(bool)", "").replace("(int)", "")
            b = float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
            out.setdefault(name, []).append(b)
    print(json.dumps({k: sum(v) / len(v) for k, v in out.items()}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1:])
