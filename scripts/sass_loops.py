#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in a cubin/.so with their static instruction
counts and the notable opcodes inside: python scripts/sass_loops.py LIB KERNEL_SUBSTRING [min_len]"""
import re, subprocess, sys
lib, kern = sys.argv[1], sys.argv[2]
min_len = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", txt)
body = next(b for b in blocks if kern in b.split("\n", 1)[0])
ins = []
for line in body.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr2idx = {a: i for i, (a, _) in enumerate(ins)}
print(f"{len(ins)} instructions")
for i, (a, t) in enumerate(ins):
    m = re.search(r"\bBRA(?:\.U)?(?:\.ANY)?\s+(?:[!\w]+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr2idx:
            j = addr2idx[tgt]
            n = i - j + 1
            if n >= min_len:
                ops = [x[1].split()[0] if not x[1].startswith("@") else x[1].split()[1] for x in ins[j:i + 1]]
                def c(pfx): return sum(1 for o in ops if o.startswith(pfx))
                print(f"loop [{j}:{i}] len {n}: MUFU {c('MUFU')} LDS {c('LDS')} LDG {c('LDG')} STG {c('STG')} "
                      f"UBLKCP {c('UBLKCP')} SYNCS {c('SYNCS')} FFMA2 {c('FFMA2')} S2R {c('S2R')} LDC {c('LDC')} BRA {c('BRA')}")
