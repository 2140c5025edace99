"""SASS evidence for a kernel of liblgx.so (no GPU needed): opcode histogram, the TMA / mbarrier / FP64 mnemonics.
    python tools/sass_excerpt.py <substring of the mangled kernel name> [<second substring>]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "cylinder-pose-estimation_b200", "liblgx.so")
want = sys.argv[1:]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
fn, ops = None, collections.Counter()
out = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if fn and m and all(w in fn for w in want):
        out.setdefault(fn, collections.Counter())[m.group(2)] += 1
for fn, ops in out.items():
    total = sum(ops.values())
    print(f"Function : {fn}\n  {total} SASS instructions, arch sm_100a")
    groups = collections.Counter()
    for op, n in ops.items():
        groups[op.split(".")[0]] += n
    for key in ("UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "LDGSTS", "DADD", "DMUL", "DFMA", "MUFU", "LDS", "STS", "LDG", "STG", "SHFL", "ATOMS", "BAR", "NANOSLEEP", "HMMA", "UTCHMMA"):
        print(f"    {key:10s} {groups.get(key, 0)}")
    print("  most frequent: " + ", ".join(f"{op} {n}" for op, n in groups.most_common(12)))
    full = [f"{op} x{n}" for op, n in sorted(ops.items()) if op.startswith(("UTMA", "SYNCS", "UBLKCP", "LDGSTS"))]
    print("  TMA / mbarrier forms: " + ", ".join(full))
