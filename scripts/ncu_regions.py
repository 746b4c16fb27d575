"""Aggregate the output of scripts/ncu_lines.py (NCU_LINES_TOP=100000) by code region: a region starts at every source line
containing `[region: name]` (marker comments in the .cu files) and extends to the next marker.
usage: python scripts/ncu_regions.py lines.txt csrc/file.cu"""
import re, sys, os
src = open(sys.argv[2]).read().splitlines()
marks = [(i + 1, re.search(r"\[region: ([^\]]+)\]", l).group(1)) for i, l in enumerate(src) if "[region:" in l]
base = os.path.basename(sys.argv[2])
def region(f, l):
    if f != base:
        return f
    name = "(top)"
    for ln, nm in marks:
        if ln <= l:
            name = nm
    return name
agg = {}
for line in open(sys.argv[1]):
    m = re.match(r"\s*([\d.]+)% \(smp\s+([\d.]+)%\) (\S+):(\d+)", line)
    if not m:
        if line.startswith(("total", "opcode")):
            print(line.strip())
        continue
    a = agg.setdefault(region(m.group(3), int(m.group(4))), [0.0, 0.0])
    a[0] += float(m.group(1)); a[1] += float(m.group(2))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-34s inst %5.1f%%  samples %5.1f%%" % (k, v[0], v[1]))
