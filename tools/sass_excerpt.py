#!/usr/bin/env python
"""profiles/{tag}_sass_excerpt.txt: which Blackwell mechanisms the shipped sm_100a kernels actually contain.
For every kernel of libpwr_b200.so: instruction count and the counts of the mnemonics that matter here
(UBLKCP = 1-D bulk TMA copy, SYNCS = mbarrier ops, LDG/STG/LDS, FFMA2 = packed f32x2), plus the first bulk-copy
site of each pipelined kernel with its neighbours.  No GPU needed (cuobjdump on the .so).

    python tools/sass_excerpt.py r2
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "pixelwiseregression_b200", "libpwr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
KEYS = ["UBLKCP", "SYNCS", "LDG", "STG", "LDS", "STS", "SHFL", "MUFU", "DADD", "F2F", "FFMA2", "BAR"]
out = ["# SASS of pixelwiseregression_b200/libpwr_b200.so (sm_100a), `cuobjdump -sass`\n",
       "UBLKCP = cp.async.bulk (1-D bulk TMA), SYNCS = mbarrier arrive / expect_tx / try_wait.  No tensor-core "
       "instructions are expected: nothing on this path is a contraction.\n",
       "%-64s %6s " % ("kernel", "instr") + " ".join("%6s" % k for k in KEYS)]
excerpts = []
cur, lines = None, []


def flush():
    if cur is None:
        return
    ops = collections.Counter()
    for l in lines:
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", l)
        if m:
            ops[m.group(1)] += 1
    total = sum(ops.values())
    name = demangle(cur)
    short = (name[:name.rfind(">(") + 1] if ">(" in name else name.split("(")[0]).replace("void pwr::", "")
    short = short.replace("(int)", "").replace("(bool)", "").replace("__nv_bfloat16", "bf16")
    out.append("%-64s %6d " % (short[:64], total) + " ".join("%6d" % sum(v for k, v in ops.items() if k.startswith(key)) for key in KEYS))
    if any(k.startswith("UBLKCP") for k in ops) and len(excerpts) < 40:
        idx = next(i for i, l in enumerate(lines) if "UBLKCP" in l)
        excerpts.append("\n## %s\n" % short + "".join(lines[max(0, idx - 4):idx + 5]))


for l in sass.splitlines(True):
    m = re.match(r"\s+Function : (\S+)", l)
    if m:
        flush()
        cur, lines = m.group(1), []
    elif cur is not None:
        lines.append(l)
flush()
seen, uniq = set(), []
for e in excerpts:                     # one excerpt per kernel family
    fam = re.sub(r"<.*", "", e.split("\n")[1])
    if fam not in seen:
        seen.add(fam)
        uniq.append(e)
path = os.path.join(ROOT, "profiles", "%s_sass_excerpt.txt" % tag)
with open(path, "w") as f:
    f.write("\n".join(out) + "\n\n# first bulk-copy site of each pipelined kernel family\n" + "".join(uniq))
print(path, len(out) - 3, "kernels")
