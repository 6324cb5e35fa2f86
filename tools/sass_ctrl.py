"""Decode the control fields (stall, yield, write/read scoreboard, wait mask) of sm_90/sm_100 SASS
from `cuobjdump -sass` output.  usage: cuobjdump -sass -fun NAME file.o | python tools/sass_ctrl.py [lo hi]"""
import re, sys
lines = sys.stdin.read().splitlines()
lo = int(sys.argv[1], 16) if len(sys.argv) > 1 else 0
hi = int(sys.argv[2], 16) if len(sys.argv) > 2 else 1 << 30
i = 0
pat = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\* (0x[0-9a-f]+) \*/")
pat2 = re.compile(r"^\s+/\* (0x[0-9a-f]+) \*/")
while i < len(lines):
    m = pat.match(lines[i])
    if m and i + 1 < len(lines):
        m2 = pat2.match(lines[i + 1])
        if m2:
            addr = int(m.group(1), 16)
            h = int(m2.group(1), 16)
            stall = (h >> 41) & 0xf; yld = (h >> 45) & 1; wbar = (h >> 46) & 7; rbar = (h >> 49) & 7; wait = (h >> 52) & 0x3f
            if lo <= addr <= hi:
                w = "".join(str(b) for b in range(6) if wait >> b & 1)
                print("%04x s%-2d %s W%s R%s wait[%-6s] %s" % (addr, stall, "Y" if yld else "-", wbar if wbar != 7 else "-", rbar if rbar != 7 else "-", w, m.group(2).strip()))
            i += 2
            continue
    i += 1
