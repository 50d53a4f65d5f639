#!/usr/bin/env python
"""Static size of the main particle loop of a pass kernel: the smallest backward-branch range of a SASS dump
(cuobjdump -sass, stdin or file) that contains all of the given marker instructions (default ATOMS), with its opcode mix.
    cuobjdump -sass -fun <mangled> x.o | python tools/sass_loop.py [MARKER] [MIN_COUNT]"""
import collections, re, sys
marker = sys.argv[1] if len(sys.argv) > 1 else "ATOMS"
mincount = int(sys.argv[2]) if len(sys.argv) > 2 else 8
ins = []
for l in sys.stdin:
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
best = None
for a, text in ins:
    m = re.search(r'BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', text)
    if m and int(m.group(1), 16) < a:
        t = int(m.group(1), 16)
        body = [x for (ad, x) in ins if t <= ad <= a]
        if sum(marker in x for x in body) >= mincount and (best is None or len(body) < len(best)):
            best = body
if best is None:
    print("no loop with", mincount, marker); sys.exit(1)
ops = collections.Counter()
for x in best:
    x = re.sub(r'^@!?U?P\d+\s+', '', x)
    ops[x.split()[0].split('.')[0]] += 1
print(len(best), "instructions;", ", ".join(f"{k} {v}" for k, v in ops.most_common()))
