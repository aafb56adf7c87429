"""Estimate FP64-pipe cycles of the k_model_chisq inner loop from its SASS with the
register-file rule measured by peakprobe.py (variants 5/6): a DFMA/DADD/DMUL that
reads three different vector registers (none through the reuse cache) takes 3 issue
cycles on the FP64 pipe, otherwise 2.

    cuobjdump -sass -fun <kernel> lib.so > k.sass ; python profiles/sass_rf_model.py k.sass
"""
import re, sys
lines = [l for l in open(sys.argv[1]) if re.search(r'/\*[0-9a-f]{4}\*/', l)]
ins = []
for l in lines:
    m = re.search(r'/\*([0-9a-f]{4})\*/\s+(.*?);', l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
# innermost loop = backward branch with the largest FP64 density
best = None
for i, (a, t) in enumerate(ins):
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        tgt = int(m.group(1), 16)
        body = [x for x in ins if tgt <= x[0] <= a]
        f = sum(1 for x in body if re.match(r'(@!?U?P\d\s+)?D(FMA|ADD|MUL)\b', x[1]))
        if f >= 20 and (best is None or f / len(body) > best[0]): best = (f / len(body), body)
body = best[1]
prev = {}
c2 = c3 = 0; nlds = 0
for a, t in body:
    if t.startswith('LDS'): nlds += 1
    m = re.match(r'(?:@!?U?P\d\s+)?(DFMA|DADD|DMUL)\s+(\S+),\s*(.*)', t)
    if not m: prev = {}; continue            # (another instruction between: reuse slots survive in HW, keep simple)
    ops = [o.strip() for o in m.group(3).split(',')]
    reads = set(); cur = {}
    for slot, o in enumerate(ops):
        r = re.match(r'[-|]*\|?(R\d+)(\.reuse)?', o)
        if not r: continue
        reg = r.group(1)
        if prev.get(slot) != reg: reads.add(reg)
        if r.group(2): cur[slot] = reg
    prev = cur
    if len(reads) >= 3: c3 += 1
    else: c2 += 1
n = c2 + c3
print(f'loop: {len(body)} instr, {n} FP64 ({c3} with 3 register reads), {nlds} LDS; '
      f'modelled pipe cycles {2*c2 + 3*c3} = {(2*c2 + 3*c3)/n:.2f} per FP64 instr')
