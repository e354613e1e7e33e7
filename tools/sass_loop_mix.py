"""Opcode histogram of the innermost loop that contains a given marker opcode in a SASS dump.
usage: python tools/sass_loop_mix.py <kernel-name-substring> [marker=SHFL] [lib.so]"""
import collections
import os
import re
import subprocess
import sys

name = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "SHFL"
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "3d-point-clouds-autocomplete_b200", "lib", "libhp_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs:
    if name not in f.split("\n", 1)[0]:
        continue
    lines = [l for l in f.splitlines() if re.search(r"/\*[0-9a-f]{4,5}\*/\s+\S", l)]
    addr = lambda l: int(re.search(r"/\*([0-9a-f]{4,5})\*/", l).group(1), 16)
    op = lambda l: re.search(r"/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l).group(1)
    best = None
    for i, l in enumerate(lines):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,?\s*)?(0x[0-9a-f]+)", l)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt >= addr(l):
            continue
        body = [x for x in lines if tgt <= addr(x) <= addr(l)]
        if any(marker in x for x in body) and (best is None or len(body) < len(best)):
            best = body
    print(f.split("\n", 1)[0])
    if best:
        c = collections.Counter(op(x) for x in best)
        print(f"  innermost loop with {marker}: {len(best)} instructions", dict(c.most_common()))
