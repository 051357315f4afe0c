#!/usr/bin/env python
"""Lists the loops of every kernel in an object file with their instruction count and the special-register reads
(S2R / S2UR), local-memory accesses and MUFU / tensor instructions inside -- ptxas likes to re-derive shared-memory window
conversions, warp and lane indices from special registers inside latency-bound loops instead of keeping a register.
    python tools/sass_loops.py svol_b200/csrc/build/attn_tc.o [kernel-name-substring]"""
import re
import subprocess
import sys
from collections import Counter

obj = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kernels[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?)\s*;", line)
    if m and cur:
        kernels[cur].append((int(m.group(1), 16), m.group(2)))
for name, ins in kernels.items():
    if flt not in name:
        continue
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110]
    print(dem, f"({len(ins)} instructions)")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for idx, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                loops.append((addr[tgt], idx))
    for lo, hi in loops:
        inner = any(l2 >= lo and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops)
        body = [x for _, x in ins[lo:hi + 1]]
        op = lambda x: x.split()[1] if x.startswith("@") else x.split()[0]
        c = Counter(op(x) for x in body)
        n = len(body)
        if n < 24:
            continue
        sr = c["S2R"] + c["S2UR"]
        loc = sum(v for k, v in c.items() if k.startswith("LDL") or k.startswith("STL"))
        mufu = sum(v for k, v in c.items() if k.startswith("MUFU"))
        tc = sum(v for k, v in c.items() if k.startswith("UTC") or k.startswith("LDTM") or k.startswith("STTM"))
        print(f"   {'outer' if inner else 'inner'} {ins[lo][0]:#07x}-{ins[hi][0]:#07x} {n:5d} instr  S2R/S2UR {sr:3d}  local {loc:3d}  MUFU {mufu:3d}  tensor/TMEM {tc:3d}")
