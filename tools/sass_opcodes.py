"""SASS opcode histogram per kernel of libmshgnn_b200.so (cuobjdump -sass): the mnemonics that prove tcgen05 / TMEM / TMA / bulk copies
and the absence of global atomics in the contraction kernels.  usage: python tools/sass_opcodes.py > profiles/<round>_sass_opcodes.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "morphsym-hgnn_b200", "lib", "libmshgnn_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "USETMAXREG", "ATOMG", "REDG", "RED", "ATOM",
        "HMMA", "LDG", "STG", "LDS", "STS"]
rows, cur, k = [], None, -1
for line in sass.split("\n"):
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = collections.Counter()
        short = re.sub(r"\(.*", "", names[k]).replace("mshgnn::", "")
        rows.append((short, cur))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["_total"] += 1
        for c in COLS:
            if op == c or op.startswith(c + "."):
                cur[c] += 1
print("# SASS opcode histogram of libmshgnn_b200.so (cuobjdump -sass, sm_100a; tools/sass_opcodes.py)")
print("# tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG / UTMASTG (prefetch UTMAPF), cp.async.bulk -> UBLKCP, mbarrier -> SYNCS, cluster barrier -> UCGABAR_*,")
print("# setmaxnreg -> USETMAXREG; no global ATOM / RED in the contraction kernels (the stack kernels use RED only for the per-item completion counters and")
print("# ATOMG for the work counter, the persistent encoder ATOMG for its work counter)\n")
print(f"{'kernel':58s}" + "".join(f"{c:>9s}" for c in COLS + ["_total"]))
for name, c in rows:
    print(f"{name[:58]:58s}" + "".join(f"{c[x]:9d}" for x in COLS + ["_total"]))
