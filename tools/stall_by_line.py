"""Joins the per-instruction stall samples of an .ncu-rep (SASS source page) with the CUDA source lines of the SAME build
(nvdisasm -g line info of the cubin inside libmshgnn_b200.so), by instruction order.
usage: python tools/stall_by_line.py <rep> <kernel-name-substring> [launch index] [top N]"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

rep, kname = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[5] if len(sys.argv) > 5 else os.path.join(ROOT, "morphsym-hgnn_b200", "lib", "libmshgnn_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith("api.") and f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cub)], capture_output=True, text=True).stdout
lines = dis.split("\n")
insts, cur, on = [], None, False
for l in lines:
    if l.startswith(".text.") and kname in l:
        on = True
        continue
    if on and (l.startswith("//-----") or l.startswith(".text.")):
        break
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insts.append((int(m.group(1), 16), m.group(2).strip(), cur))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.split("\n")
nname = os.environ.get("NCU_KERNEL", kname)      # demangled name in the report when it differs from the mangled substring
starts = [i for i, l in enumerate(raw) if l.startswith('"Kernel Name"') and nname in l]
a = starts[which]
b = next((i for i in range(a + 1, len(raw)) if raw[i].startswith('"Kernel Name"')), len(raw))
rd = csv.reader(io.StringIO("\n".join(raw[a + 1:b])))
hdr = next(rd)
rows = [r for r in rd if len(r) == len(hdr)]
si, so_ = hdr.index("# Samples"), hdr.index("Source")
ws = hdr.index("Warp Stall Sampling (All Samples)")
assert len(rows) == len(insts), (len(rows), len(insts))
by_line, tot = Counter(), 0
for r, (off, txt, loc) in zip(rows, insts):
    n = int(r[si] or 0)
    tot += n
    by_line[loc] += n
src_cache = {}
def src(loc):
    if not loc: return ""
    fn = os.path.join(ROOT, "morphsym-hgnn_b200", "csrc", loc[0])
    if fn not in src_cache:
        src_cache[fn] = open(fn).read().split("\n") if os.path.exists(fn) else []
    L = src_cache[fn]
    return L[loc[1] - 1].strip()[:110] if 0 < loc[1] <= len(L) else ""
print("total samples", tot)
for loc, n in by_line.most_common(top):
    print(f"{n:6d} {100.0 * n / tot:5.1f}%  {loc}  {src(loc)}")
