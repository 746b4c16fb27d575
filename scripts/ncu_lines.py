"""Line-level view of one kernel of an .ncu-rep: joins `ncu --page source --csv` (per-SASS-instruction counters, in
program order) with `nvdisasm --print-line-info` of the same cubin (source line of every instruction, same order).

usage: python scripts/ncu_lines.py report.ncu-rep <kernel regex> <cubin name inside the .so, e.g. accumulate> [mangled filter]
(the filter defaults to the template instantiation named in the report; the library must be the build that was profiled)
"""
import csv, io, os, re, subprocess, sys, tempfile, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "larnd-sim-jax_b200", "larndsim_b200", "liblarnd_b200.so")
rep, kre, cub = sys.argv[1], sys.argv[2], sys.argv[3]
filt = sys.argv[4] if len(sys.argv) > 4 else None

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
# several launches may match: keep the first block
blocks = out.split('"Kernel Name"')
body = '"Kernel Name"' + blocks[1]
lines = body.splitlines()
if filt is None:
    # template instantiations share the source lines: pick the .text section of the instantiation that was profiled
    # (k_acc_tiles<4> -> mangled ...k_acc_tilesILi4E...), not the first one in the cubin
    m = re.search(r"(%s)<(?:\(int\))?(\d+)>" % kre, out)
    filt = "%sILi%sE" % (m.group(1), m.group(2)) if m else kre
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
iS, iI, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
prof = [(r[iS].strip(), int(r[iI] or 0), int(r[iN] or 0)) for r in rows[1:] if len(r) > iN]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", cub + ".sm_100a.cubin", SO], cwd=tmp, capture_output=True)
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub + ".sm_100a.cubin")], capture_output=True, text=True).stdout
# walk the .text section of the wanted kernel
sec = None
cur = ("?", 0)
insts = []
for l in dis.splitlines():
    m = re.match(r"//-+ \.text\.(\S+) -+", l)
    if m:
        sec = m.group(1)
        continue
    if sec is None or not re.search(filt, sec):
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m:
        insts.append((cur, m.group(1).strip()))
print("prof", len(prof), "dis", len(insts))
n = min(len(prof), len(insts))
tot = sum(p[1] for p in prof)
tots = sum(p[2] for p in prof) or 1
agg = collections.defaultdict(lambda: [0, 0])
for (src, ni, ns), (loc, txt) in zip(prof[:n], insts[:n]):
    agg[loc][0] += ni
    agg[loc][1] += ns
print("total instr", tot)
srcs = {}
def srcline(loc):
    f, ln = loc
    for d in (os.path.join(ROOT, "larnd-sim-jax_b200", "csrc"),):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in srcs:
                srcs[p] = open(p).read().splitlines()
            return srcs[p][ln - 1].strip()[:100] if ln - 1 < len(srcs[p]) else ""
    return ""
TOP = int(os.environ.get("NCU_LINES_TOP", "45"))
ops = collections.Counter()
for (src, ni, ns), (loc, txt) in zip(prof[:n], insts[:n]):
    t = txt.split()
    op = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "?")
    ops[op.split(".")[0]] += ni
print("opcode mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in ops.most_common(24)))
for loc, (ni, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:TOP]:
    print("%5.1f%% (smp %4.1f%%) %s:%d  %s" % (100.0 * ni / tot, 100.0 * ns / tots, loc[0], loc[1], srcline(loc)))
