"""Decode the scheduling control fields of a cuobjdump -sass listing (sm_7x+ 128-bit encoding):
stall count, yield, write/read barrier, wait mask.  usage: sass_ctrl.py file.sass <function-substring> <addr_lo> <addr_hi>"""
import re, sys
path, fn, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
lines = open(path).read().split("\n")
infn = False
i = 0
tot = 0
while i < len(lines):
    l = lines[i]
    if "Function :" in l:
        infn = fn in l
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/", l)
    if infn and m:
        addr = int(m.group(1), 16)
        m2 = re.search(r"/\* (0x[0-9a-f]{16}) \*/", lines[i + 1])
        hi64 = int(m2.group(1), 16)
        ctrl = hi64 >> 41
        stall = ctrl & 0xF
        yld = (ctrl >> 4) & 1
        wbar = (ctrl >> 5) & 7
        rbar = (ctrl >> 8) & 7
        wait = (ctrl >> 11) & 0x3F
        if lo <= addr <= hi:
            tot += stall
            print(f"{addr:05x} st={stall:2d} y={yld} w={wbar if wbar!=7 else '-'} r={rbar if rbar!=7 else '-'} wait={wait:06b}  {m.group(2).strip()}")
        i += 2
        continue
    i += 1
print("sum of stall counts:", tot)
