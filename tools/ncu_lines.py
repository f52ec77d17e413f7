#!/usr/bin/env python
"""Per-source-line instruction / stall totals of one kernel from an .ncu-rep.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [--top N] [--lib path/to/lib.so]

Joins `ncu --page source --csv` (SASS-level counters) with `nvdisasm -g` line
info of the same function (matched by instruction offset).  Needs -lineinfo.
"""
import argparse
import csv
import glob
import io
import os
import re
import subprocess
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_lines(lib):
    tmp = tempfile.mkdtemp(prefix="cubin_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    out = {}
    for cubin in glob.glob(os.path.join(tmp, "*.cubin")):
        txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
        func, line = None, None
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                func = m.group(1)
                line = None
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m and func:
                out.setdefault(func, {})[int(m.group(1), 16)] = (line, m.group(2).strip())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=30)
    ap.add_argument("--lib", default=os.path.join(ROOT, "pixelwiseregression_b200", "libpwr_b200.so"))
    ap.add_argument("--launch", type=int, default=0, help="which matching launch in the report")
    args = ap.parse_args()
    res = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--kernel-name",
                          "regex:" + args.kernel], capture_output=True, text=True).stdout
    blocks = res.split('"Kernel Name"')[1:]
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blocks[args.launch])))
    name = rows[0][1]
    hdr = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][col["Address"]], 16)
    cands = sass_lines(args.lib)
    short = re.sub(r"^void\s+", "", name).split("<")[0].split("(")[0].split("::")[-1]
    funcs = {f: v for f, v in cands.items() if short in f and len(v) == len(data)}
    if not funcs:
        raise SystemExit("no function with %d instructions matching %s (library rebuilt since the capture?)"
                         % (len(data), short))
    func, lines = sorted(funcs.items())[0]
    per_line = defaultdict(lambda: [0, 0, 0])
    tot_i = tot_s = 0
    for r in data:
        off = int(r[col["Address"]], 16) - base
        line, _ = lines.get(off, (None, ""))
        inst = int(r[col["Instructions Executed"]] or 0)
        stall = int(r[col["Warp Stall Sampling (All Samples)"]] or 0)
        per_line[line][0] += inst
        per_line[line][1] += stall
        per_line[line][2] += 1
        tot_i += inst
        tot_s += stall
    print("%s\n%s: %d SASS instructions, %d warp-instructions executed, %d stall samples" % (
        name, func, len(data), tot_i, tot_s))
    src_cache = {}

    def text(line):
        if line is None:
            return "?"
        f, n = line
        if f not in src_cache:
            path = glob.glob(os.path.join(ROOT, "**", f), recursive=True)
            src_cache[f] = open(path[0]).read().splitlines() if path else []
        s = src_cache[f]
        return s[n - 1].strip()[:110] if 0 < n <= len(s) else ""
    print("  inst%  stall%  sass  line")
    for line, (i, s, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:args.top]:
        print("%6.1f %7.1f %5d  %s:%s  %s" % (100.0 * i / max(tot_i, 1), 100.0 * s / max(tot_s, 1), n,
                                             line[0] if line else "?", line[1] if line else "?", text(line)))


if __name__ == "__main__":
    main()
