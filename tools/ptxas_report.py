"""Registers / spills / stack per kernel from the last build (tarok_b200/build.log, nvcc -Xptxas -v)."""
import os, re, subprocess, sys
log = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tarok_b200", "build.log")).read()
rows = []
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\s*\nptxas info\s*: Function properties for \S+\s*\n\s*(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\nptxas info\s*: Used (\d+) registers", log):
    rows.append(m.groups())
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.split("\n")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for r, n in zip(rows, names):
    n = re.sub(r"\(.*", "", n).replace("void tk::", "")
    if pat in n:
        print("%-34s regs %3s  stack %3s  spill st/ld %s/%s" % (n, r[4], r[1], r[2], r[3]))
