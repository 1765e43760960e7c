"""Dev helper: instruction mix of the hottest loop(s) of a kernel in `cuobjdump -sass` output."""
import re, sys, collections
lines = [l for l in open(sys.argv[1]) if re.match(r"^\s+/\*[0-9a-f]{4,5}\*/", l)]
ins = []
for l in lines:
    m = re.match(r"^\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
    addr = int(m.group(1), 16); txt = m.group(2)
    txt = re.sub(r"^@!?U?P\w+\s+", "", txt)
    ins.append((addr, txt))
loops = []
for a, t in ins:
    m = re.match(r"BRA(\.\w+)*\s+(?:\w+,\s*)?(0x[0-9a-f]+)", t)
    if m and int(m.group(2), 16) < a:
        loops.append((int(m.group(2), 16), a))
for lo, hi in loops:
    body = [t for a, t in ins if lo <= a <= hi]
    c = collections.Counter(re.split(r"[.\s]", t)[0] for t in body)
    print("loop 0x%x..0x%x: %d instr" % (lo, hi, len(body)), dict(c.most_common(14)))
