#!/usr/bin/env python
"""Size and opcode mix of the innermost loop that contains DMMA instructions (here, no GPU):
   python tools/sass_loop.py OBJ FUNCTION_SUBSTRING"""
import collections
import re
import subprocess
import sys

obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn, ins = None, collections.defaultdict(list)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and fn:
        ins[fn].append((int(m.group(1), 16), m.group(2).strip()))
for f, L in ins.items():
    if pat not in f:
        continue
    dm = [a for a, s in L if "DMMA" in s]
    best = None
    for a, s in L:  # backward branches enclosing the first DMMA
        m = re.search(r"BRA\S*\s+.*?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < a and t <= dm[0] <= a and (best is None or a - t < best[1] - best[0]):
                best = (t, a)
    body = [s for a, s in L if best[0] <= a <= best[1]]
    c = collections.Counter((s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0] for s in body)
    print(f[:60], "loop 0x%x..0x%x: %d instructions" % (best[0], best[1], len(body)))
    print("  ", dict(c.most_common()))
