"""Development aid: print SASS with decoded control codes (stall, yield, write/read barrier, wait mask).
usage: sass_ctl.py file.o function-substring [first] [last]"""
import re, subprocess, sys
obj, sub = sys.argv[1], sys.argv[2]
first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
last = int(sys.argv[4]) if len(sys.argv) > 4 else 10**9
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
on = False
k = 0
ins = None
for ln in out:
    if "Function :" in ln:
        on = sub in ln
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);\s*/\* 0x([0-9a-f]+) \*/", ln)
    if m:
        ins = (m.group(1), m.group(2).strip(), int(m.group(3), 16))
        continue
    m = re.match(r"\s*/\* 0x([0-9a-f]+) \*/", ln)
    if m and ins:
        hi = int(m.group(1), 16)
        ctl = hi >> 41
        stall = ctl & 0xf
        yld = (ctl >> 4) & 1
        wr = (ctl >> 5) & 7
        rd = (ctl >> 8) & 7
        wait = (ctl >> 11) & 0x3f
        if first <= k <= last:
            w = "".join(str(b) if wait >> b & 1 else "-" for b in range(6))
            print(f"{k:6d} st={stall:2d} {'Y' if yld else ' '} W{wr if wr != 7 else '-'} R{rd if rd != 7 else '-'} wait[{w}]  {ins[1][:80]}")
        k += 1
        ins = None
