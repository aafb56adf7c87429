"""Print the inner loop of a kernel with its scheduling control fields decoded from the
128-bit SASS encoding (Volta-and-later layout, high word bits 41..61: stall count, yield,
write/read barrier index, wait mask, operand-reuse flags).

    cuobjdump -sass -fun <mangled kernel> libmc3b200.so > k.sass ; python profiles/sass_ctrl.py k.sass

What it showed for k_model_chisq<SineGridModel> (round 1): dependent FP64 instructions are spaced
8 cycles (the pipe's latency), independent ones 2; `.reuse` is set inside groups of same-kind FMAs
(recurrence constant, slope/offset) but not on the first FMA after a shared-memory scoreboard
wait; the eight residual FMAs per iteration read three fresh registers each, which no ordering
can change (1/sigma and d/sigma are per-point values held in vector registers)."""
import re
import sys

L = open(sys.argv[1]).read().splitlines()
ins = []
i = 0
while i < len(L):
    m = re.search(r'/\*([0-9a-f]{4})\*/\s+(.*?);\s*/\* (0x[0-9a-f]{16}) \*/', L[i])
    if m and i + 1 < len(L):
        m2 = re.search(r'/\* (0x[0-9a-f]{16}) \*/', L[i + 1])
        ins.append((int(m.group(1), 16), m.group(2).strip(), int(m2.group(1), 16) if m2 else 0))
        i += 2
    else:
        i += 1
best = None
for a, t, h in ins:
    m = re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)', t)
    if m and int(m.group(1), 16) < a:
        body = [x for x in ins if int(m.group(1), 16) <= x[0] <= a]
        f = sum(1 for x in body if re.match(r'(@!?U?P\d\s+)?D(FMA|ADD|MUL)\b', x[1]))
        if f >= 20 and (best is None or f/len(body) > best[0]):
            best = (f/len(body), body)
tot = 0
for a, t, h in best[1]:
    stall, yld = (h >> 41) & 0xf, (h >> 45) & 1
    wb, rb, wait, reuse = (h >> 46) & 7, (h >> 49) & 7, (h >> 52) & 0x3f, (h >> 58) & 0xf
    tot += stall
    print(f'{a:04x} {t:48s} stall={stall:2d} yield={yld} wbar={wb} rbar={rb} wait={wait:06b} reuse={reuse:04b}')
print(f'{len(best[1])} instructions, {tot} stall cycles for one warp alone')
