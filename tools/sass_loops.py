#!/usr/bin/env python
"""List the loops (backward branches) of a cuobjdump -sass dump with their opcode mix.

    cuobjdump -sass -fun <mangled> lib.so > k.sass;  python tools/sass_loops.py k.sass [min_size]
"""
import collections
import re
import sys

ins = []
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)\s*(.*?);")
for line in open(sys.argv[1]):
    m = pat.match(line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2), m.group(3)))
min_size = int(sys.argv[2]) if len(sys.argv) > 2 else 40
addr_index = {a: i for i, (a, _, _) in enumerate(ins)}
for i, (a, op, rest) in enumerate(ins):
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", rest)
        if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in addr_index:
            j = addr_index[int(m.group(1), 16)]
            body = ins[j:i + 1]
            if len(body) < min_size:
                continue
            mix = collections.Counter(o.split(".")[0] for _, o, _ in body)
            print("loop 0x%x..0x%x  %d instr  %s" % (body[0][0], a, len(body),
                  " ".join("%s:%d" % kv for kv in mix.most_common(14))))
