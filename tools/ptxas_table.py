"""Registers / stack / spills per kernel from `nvcc -Xptxas -v` output on stdin."""
import re, subprocess, sys
name, spill = None, None
for l in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", l)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:90]
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", l)
    if m:
        spill = m.groups()
    m = re.search(r"Used (\d+) registers", l)
    if m and name:
        print("%-92s regs %3s  stack/spill %s" % (name, m.group(1), spill))
        name = None
