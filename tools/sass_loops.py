"""Instruction mix of the loops (backward branches) of one kernel in a cuobjdump -sass dump.
usage: python tools/sass_loops.py <dump.sass> <mangled-name-substring>"""
import collections
import re
import sys

text = open(sys.argv[1]).read().split("Function : ")
for fn in text[1:]:
    name = fn.split("\n", 1)[0].strip()
    if sys.argv[2] not in name:
        continue
    ops = []
    for l in fn.splitlines():
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
        if m:
            ops.append((int(m.group(1), 16), m.group(2), m.group(3)))
    print(name, len(ops), "instructions")
    for a, op, rest in ops:
        m = re.search(r"0x([0-9a-f]+)", rest) if op.startswith("BRA") else None
        if m and int(m.group(1), 16) < a:
            t = int(m.group(1), 16)
            body = [x for x in ops if t <= x[0] <= a]
            c = collections.Counter(x[1].split(".")[0] + (".SAT" if ".SAT" in x[1] else "") for x in body)
            print(f"  loop {t:#x}..{a:#x}: {len(body)} instr", dict(c.most_common()))
