#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into profiles/: one markdown table per kernel
and profiles/traffic.json (DRAM bytes per launch, read by bench.py for roofline.traffic).

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep --tag r1 [--batch 4096 --joints 14]
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pixelwiseregression_b200 import roofline  # noqa: E402

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of ncu peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem"),
    ("launch__shared_mem_per_block_static", "static smem"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
ENTRY = {"sfr_build_kernel": "pwr_sfr_build", "decoder_fused_kernel": "pwr_decoder_fwd_bwd_loss",
         "decoder_fwd_kernel": "pwr_decoder_fwd",
         "decoder_bwd_pipe_kernel": "pwr_decoder_bwd_loss", "decoder_bwd_kernel": "pwr_decoder_bwd_loss"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--joints", type=int, default=14)
    ap.add_argument("--print-only", action="store_true", help="exploratory capture: write nothing under profiles/")
    ap.add_argument("--append", default="", metavar="TITLE",
                    help="append the tables under this title to profiles/{tag}_ncu_summary.md; traffic.json is left alone")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    alg = {"pwr_sfr_build": roofline.sfr_build_bytes(args.joints) * args.batch,
           "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(args.joints) * args.batch,
           "pwr_decoder_fwd": roofline.decoder_fwd_bytes(args.joints) * args.batch,
           "pwr_decoder_bwd_loss": roofline.decoder_bwd_bytes(args.joints) * args.batch}
    out_md = ["# ncu summary %s (`%s`, B=%d, J=%d)\n" % (args.tag, os.path.basename(args.report), args.batch, args.joints),
              "Captured with `ncu --set full --clock-control none --import-source on` under gpurun on one B200; "
              "durations under the profiler are not benchmark values.\n"]
    traffic = {}
    for r in data:
        name = r[col["Kernel Name"]]
        short = name.split("<")[0].split("(")[0].replace("void ", "").strip()
        out_md.append("\n## `%s`\n\n| metric | value |\n|---|---|" % name[:100])
        vals = {}
        for m, label in METRICS:
            if m in col:
                v, u = r[col[m]], units[col[m]]
                vals[m] = (v, u)
                out_md.append("| %s (`%s`) | %s %s |" % (label, m, v, u))
        entry = next((e for k, e in ENTRY.items() if short.endswith(k)), None)
        try:
            rd = float(vals["dram__bytes_read.sum"][0]) * UNIT_SCALE[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * UNIT_SCALE[vals["dram__bytes_write.sum"][1]]
        except Exception:
            continue
        if entry and entry not in traffic:
            traffic[entry] = {"kernel": short, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                              "algorithmic_bytes_per_launch": alg[entry], "traffic_over_algorithmic": (rd + wr) / alg[entry],
                              "batch": args.batch, "joints": args.joints, "report": os.path.basename(args.report)}
            out_md.append("| **DRAM traffic / algorithmic bytes** | %.3f GB / %.3f GB = %.2f |" % (
                (rd + wr) / 1e9, alg[entry] / 1e9, (rd + wr) / alg[entry]))
    if args.print_only:
        print("\n".join(out_md))
        return
    if args.append:
        out_md[0] = "\n# %s (`%s`, B=%d, J=%d)\n" % (args.append, os.path.basename(args.report), args.batch, args.joints)
        del out_md[1]
        out_md = [l for l in out_md if "DRAM traffic / algorithmic bytes" not in l]
        with open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % args.tag), "a") as f:
            f.write("\n".join(out_md) + "\n")
        print("\n".join(out_md[-30:]))
        return
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "%s_ncu_summary.md" % args.tag), "w") as f:
        f.write("\n".join(out_md) + "\n")
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(out_md[-40:]))


if __name__ == "__main__":
    main()
