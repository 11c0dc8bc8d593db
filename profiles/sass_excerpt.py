#!/usr/bin/env python
"""Static SASS summary of the hot kernels (no GPU needed): `python profiles/sass_excerpt.py > profiles/r2_sass_excerpt.txt`.
Counts the memory / synchronisation / FP64 mnemonics per kernel from `cuobjdump -sass glomap_b200/libb200sfm.so`."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNELS = ["ba3_linearize_points", "ba2_linearize_cams", "ba3_pass_a", "ba2_pass_b", "ba2_schur_diag", "ba3_cost",
           "ba2_pcg_direction_pack", "pcg_update", "pcg_apply_diag", "bax_pass_b", "ba2k_cross", "gp_schur_pass", "ra_laplacian_csr",
           "ra2_laplacian_dot", "ra2_coarse", "p2p_allreduce_sum", "proc_undistort", "trk_hook"]
KEEP = re.compile(r"^(LDG|STG|LDS|STS|RED|ATOM|BAR|SHFL|DFMA|DMUL|DADD|MUFU|CCTL|UBLKCP|SYNCS|MEMBAR|ERRBAR|LDGSTS)")


def main():
    lib = os.path.join(ROOT, "glomap_b200", "libb200sfm.so")
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
    mangled = re.findall(r"Function : (\S+)", txt)
    demangle = dict(zip(mangled, names))
    cur, counts, totals = None, collections.defaultdict(collections.Counter), collections.Counter()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            totals[cur] += 1
            op = m.group(1)
            if KEEP.match(op):
                counts[cur][op] += 1
    print("# SASS mnemonics of the hot kernels (cuobjdump -sass glomap_b200/libb200sfm.so, sm_100a), round 2, final build")
    print("# per kernel: total instructions, then the memory / synchronisation / FP64 mnemonics with their static counts")
    print("# (profiles/sass_excerpt.py regenerates this file; no tensor-core or TMA mnemonics are expected: FP64 streaming kernels)\n")
    for k in sorted(counts):
        if not any(re.search(r"\b" + re.escape(n) + r"\b", k) for n in KERNELS):
            continue
        ops = ", ".join(f"{o} x{c}" for o, c in counts[k].most_common(14))
        print(k[:200])
        print(f"    total {totals[k]} instructions: {ops}")


if __name__ == "__main__":
    main()
