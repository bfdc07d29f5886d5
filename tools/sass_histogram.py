"""Opcode histogram of the kernels in libhept_sm100.so (cuobjdump -sass, read here, no GPU needed).

    python tools/sass_histogram.py [regex of kernel names] > profiles/r2_sass_histogram.txt

What to look for (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / .st, UTCBAR = tcgen05.commit,
LDGSTS = cp.async, UTMALDG / UTMASTG = TMA (none here: the rows are GATHERED through the sort permutation, 96-byte pieces
that a tensor map cannot describe), STG.E.ENL2.256 = 256-bit stores, RED / ATOMG = the table-ordered vector adds.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "hept_b200", "libhept_sm100.so")
pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "block_attn_(fwd|bwd)_tc_kernel")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, sizes = None, {}, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = name if pat.search(name) else None
        if kern:
            hist[kern] = collections.Counter()
            sizes[kern] = 0
        continue
    if kern:
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Za-z0-9_.]*)", line)
        if m:
            op = m.group(2)
            base = op.split(".")[0]
            key = op if base in ("UTCHMMA", "LDTM", "STTM", "STG", "LDG", "ATOMG", "RED", "LDGSTS", "UTCBAR", "SYNCS", "MUFU") else base
            hist[kern][key] += 1
            sizes[kern] = max(sizes[kern], int(m.group(1), 16) + 16)
for k in hist:
    short = re.sub(r"\(.*", "", k)
    print(f"== {short}   ({sizes[k] / 1024:.1f} KB of SASS, {sum(hist[k].values())} instructions)")
    for op, c in hist[k].most_common():
        print(f"   {op:28s} {c}")
