#!/usr/bin/env python
"""Regenerate profiles/README.md from the round's artefacts in profiles/:
{tag}_bench_n1.json (+ n2 / n4 / n8), {tag}_bench_reference.json, {tag}_launches.csv, traffic.json,
{tag}_ncu_summary.md, {tag}_sanitizer_*.log, {tag}_pcie_probe*.json.

    python tools/profiles_readme.py r2
"""
import csv
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"


def load(name):
    path = os.path.join(P, name)
    if not os.path.isfile(path):
        return None
    with open(path) as f:
        txt = f.read().strip()
    return json.loads(txt.splitlines()[-1]) if txt else None


bench = load("%s_bench_n1.json" % tag)
ref = load("%s_bench_reference.json" % tag)
ev, evtot = bench["kernels"], bench["ms_per_step"]
B = bench["config"]["batch_per_gpu"]
L = ["# profiles/ — round %s\n" % tag[1:],
     "All captured under `gpurun` on B200 (sm_100a).  `bench.py` at its default workload: NYU shape, B = %d per GPU, "
     "J = %d, float32 frames resident in HBM, dense targets, last stage in one pass (the product default).  Round-1 files "
     "(`r1_*`) are kept for comparison.\n" % (B, bench["config"]["joints"]),
     "| file | what |", "|---|---|",
     "| `%s_bench_n1.json` (+ `_n2`, `_n4`, `_n8`) | the JSON line of `bench.py` at N = 1 / 2 / 4 / 8 (torchrun, NCCL) |" % tag,
     "| `%s_bench_reference.json` | the JSON line of `bench.py --impl reference` (CPU oracle port) on the same box |" % tag,
     "| `%s_launches.csv` | `ncu --metrics gpu__time_duration.sum --clock-control none` launch list of a short `bench.py` run |" % tag,
     "| `%s_ncu_summary.md`, `traffic.json` | one `ncu --set full` capture of every hot kernel (`tools/prof_step.py`, `tools/ncu_summary.py`); DRAM bytes per launch (read by `bench.py` -> `roofline.traffic`) |" % tag,
     "| `%s_sass_excerpt.txt` | per-kernel instruction mix of the shipped `.so` (`tools/sass_excerpt.py`): which kernels contain `UBLKCP` (bulk TMA) / `SYNCS` (mbarrier) |" % tag,
     "| `%s_sanitizer_{racecheck,memcheck,initcheck,smoke}.log` | `compute-sanitizer` over the one-pass kernel (all 12 instantiations), the six-slot backward, fetch + window mode, the raw-frame window table, the bb loader and `smoke()` (`tools/sanitize.sh`) |" % tag,
     "| `%s_pcie_probe.json`, `%s_pcie_probe_n8.json`, `%s_topo*.txt` | `tools/pcie_probe.cu`: how a GPU can pull frames / crop windows from pinned host memory (design of the e2e feed); the same probe on 8 GPUs at once; `nvidia-smi topo -m` |" % (tag, tag, tag),
     "| `%s_sanitizer_lean_racecheck.log`, `%s_sanitizer_window_table_{racecheck,memcheck}.log`, `%s_ab_fused_lean.txt` | racecheck of `decoder_fused_lean_kernel` (the three-CTA one-pass variant, all 4 instantiations); `tools/ab_fused.py`: lean vs two-CTA one-pass kernel, ms and bit-identity; racecheck + memcheck of the SFR builder with the sentinel window table and of the uneven-shard feed test |" % (tag, tag, tag),
     "| `%s_ab_sfr_window_table.txt`, `%s_ab_sfr_hoisted_taps.txt`, `%s_ab_sfr_packed_taps.txt` | `tools/ab_sfr.py` A/B of the SFR builder: sentinel window-table lookup (kept: 0.244 -> 0.212 ms compact raw frames); per-thread hoisted column taps at 4 / 5 / 6 CTAs per SM and packed 16-byte tap tables (both measured, rejected: register pressure / no gain) |" % (tag, tag, tag),
     "| `%s_bench_n8_balanced.json`, `%s_bench_n4_balanced.json`, `%s_bench_n2_balanced.json` | `bench.py --no-extras` at N = 8 / 4 / 2 with the bandwidth-proportional shards of the e2e leg (`e2e.sharding`, `e2e.equal_shards`; at N = 2 and 4 the links of those boxes were equal and the equal shards stood) |" % (tag, tag, tag),
     "| `%s_sweep_hand17_n1.txt` | `tools/sweep_inference.py`: BASELINE configs[4] inference sweep, batch 256 .. 16384 |" % tag,
     "| `../tools/capture_profiles.sh`, `../tools/sanitize.sh` | the exact commands behind the above |", ""]

# ---- share of the step: ncu launch list vs events
lp = os.path.join(P, "%s_launches.csv" % tag)
if os.path.isfile(lp):
    rows = list(csv.reader(open(lp)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    seq = [(r[ki], float(r[vi].replace(",", "")) / (1e3 if r[ui] == "ns" else 1)) for r in data if len(r) > vi]
    idx = [i for i, (n, _) in enumerate(seq) if "sfr_prep" in n]
    step = seq[idx[-1]:]
    tot = sum(t for _, t in step)
    agg = {}
    for n, t in step:
        k = n.split("(")[0].replace("void ", "")[:70]
        agg[k] = agg.get(k, 0) + t
    ent = lambda k: ("pwr_sfr_build" if "sfr_build_kernel" in k else "pwr_decoder_fwd_bwd_loss" if "decoder_fused" in k else None)
    L += ["## Share of the step: ncu launch list (cold, serialised) vs CUDA events inside bench.py\n",
          "| kernel | ncu us | ncu share | bench events ms (entry point) | bench share |", "|---|---|---|---|---|"]
    for k, t in sorted(agg.items(), key=lambda kv: -kv[1]):
        e = ent(k)
        L.append("| `%s` | %.1f | %.1f %% | %s | %s |" % (k, t, 100 * t / tot, ("%.3f" % ev[e]["avg_ms"]) if e else "-",
                                                     ("%.1f %%" % (100 * ev[e]["avg_ms"] / evtot)) if e else "-"))
    L += ["| **step total** | %.1f | 100 %% | %.3f (`ms_per_step`) | |" % (tot, evtot), ""]

# ---- roofline
tr = json.load(open(os.path.join(P, "traffic.json")))
peak = bench["roofline"]["peak"]
L += ["## Roofline (measured peak %.1f GB/s, MEASURED_PEAKS.json)\n" % peak,
      "| entry point | algorithmic bytes / launch | event-timed ms | achieved GB/s | frac of measured peak | DRAM traffic / algorithmic (ncu) |",
      "|---|---|---|---|---|---|"]
for e in ev:
    L.append("| `%s` | %d | %.3f | %.0f | %.3f | %s |" % (e, ev[e]["algorithmic_bytes"], ev[e]["avg_ms"], ev[e]["achieved_gbs"], ev[e]["frac"],
                                                      ("%.2f" % tr[e]["traffic_over_algorithmic"]) if e in tr else "-"))
L += ["",
      "Whole step: %d B/sample x %d / %.3f ms = %.0f GB/s = %.3f of the measured peak (`step_roofline_frac`), %.2f M samples/s "
      "(median step %.3f ms); replayed as one CUDA graph: %.3f ms, %.1f us of host work per step (call by call: %.0f us)."
      % (bench["config"]["algorithmic_bytes_per_sample"], B, evtot, bench["config"]["algorithmic_bytes_per_sample"] * B / evtot / 1e6,
         bench["step_roofline_frac"], bench["value"] / 1e6, bench["ms_per_step_median"], bench["graph_step"]["ms_per_step"],
         bench["graph_step"]["host_issue_ms_per_step"] * 1e3, bench["graph_step"]["host_issue_ms_per_step_calls"] * 1e3),
      "`pwr_sfr_build` moves %.2fx its algorithmic bytes (the formula counts a 128x128 crop; the non-antialiased taps of a "
      "176-352 px box touch every source pixel): against its real DRAM traffic it runs at %.0f GB/s = %.0f %% of the measured peak."
      % (tr["pwr_sfr_build"]["traffic_over_algorithmic"], tr["pwr_sfr_build"]["dram_bytes_per_launch"] / ev["pwr_sfr_build"]["avg_ms"] / 1e6,
         tr["pwr_sfr_build"]["dram_bytes_per_launch"] / ev["pwr_sfr_build"]["avg_ms"] / 1e6 / peak * 100), ""]

# ---- variants and inner stage
L += ["## Variants of the step and the kernels of an inner stage (`%s_bench_n1.json`)\n" % tag,
      "| what | samples/s | ms/step | step frac | kernels (ms, frac of measured peak on own bytes) |", "|---|---|---|---|---|"]
for k, title in (("two_kernel_step", "SURVEY 8d accounting (forward, then backward+loss)"), ("raw_frames_step", "raw uint16 NYU frames (decode + hand rectangle in the kernel)"),
                 ("no_heat_store_step", "last stage without the heat-map store (what `forward_loss` runs)"),
                 ("dense_tuple_compact_loss", "dense tuple written for the caller, loss evaluated from the taps (`targets='both'`)"),
                 ("sparse_targets", "compact targets (64-byte taps per joint)"),
                 ("raw_frames_sparse_targets", "raw uint16 frames AND compact targets")):
    v = bench.get(k)
    if v:
        L.append("| `%s`: %s | %.2f M | %.3f | %.3f | %s |" % (k, title, v["value"] / 1e6, v["ms_per_step"], v["step_roofline_frac"],
                                                           ", ".join("`%s` %.3f (%.2f)" % (a, b["avg_ms"], b["frac"]) for a, b in v["kernels"].items())))
inner = bench.get("inner_stage")
if inner:
    L += ["", "| inner-stage kernel | ms | B/sample | frac of measured peak |", "|---|---|---|---|"]
    for k, v in inner.items():
        if k != "clocks":
            L.append("| %s | %.3f | %d | %.3f |" % (k, v["avg_ms"], v["algorithmic_bytes_per_sample"], v["frac"]))
    ts = bench["two_stage_decoder"]
    L.append("| decoder kernels of a 2-stage training step (stage-0 forward+loss, last stage in one pass, stage-0 backward) | %.3f | %d | %.3f |"
             % (ts["avg_ms"], ts["algorithmic_bytes_per_sample"], ts["frac"]))

# ---- e2e, baselines
e = bench["e2e"]
L += ["", "## End to end and baselines\n",
      "* `e2e` (public feed API, `feed.HostFeed`): %.0f samples/s at N = 1, %.2f ms per step, %.1f MB H2D per step (%.1f KB per sample "
      "of the %.0f KB a raw frame occupies) = %.1f GB/s on the PCIe link, %.0f KB D2H; host time per step: submit %.2f ms, consume %.2f ms, "
      "waiting for the results %.2f ms.  Round 1's definition (whole float32 frames + logits, serial): %.0f samples/s."
      % (e["value"], e["ms_per_step"], e["h2d_bytes_per_step"] / 1e6, e["h2d_bytes_per_step"] / B / 1e3,
         e["frames_in_host_memory_bytes"] / B / 1e3, e["pcie_gbs_per_gpu"], e["d2h_bytes_per_step"] / 1e3, e["host_ms_per_step"]["submit"],
         e["host_ms_per_step"]["consume"], e["host_ms_per_step"]["wait_for_results"],
         bench["e2e_whole_frames_r1_definition"]["value"] if bench.get("e2e_whole_frames_r1_definition") else float("nan")),
      "* `cpu_baseline` (oracle port, %d host cores): %.0f samples/s; `--impl reference` run: %s samples/s."
      % (bench["cpu_baseline"]["cores"], bench["cpu_baseline"]["value"], ("%.0f" % ref["value"]) if ref else "n/a"),
      "* `gpu_eager_decoder` (the reference's decoder + loss lines as eager PyTorch on the same GPU): %.2f ms vs %.2f ms fused = %.1fx."
      % (bench["gpu_eager_decoder"]["ms"], bench["gpu_eager_decoder"]["fused_ms"], bench["gpu_eager_decoder"]["speedup"])]

# ---- scaling
L += ["", "## Scaling (weak; one process per GPU, NCCL)\n",
      "| N | samples/s (HBM-resident) | ms/step | of N x N=1 | e2e samples/s | e2e of N x N=1 | PCIe GB/s per GPU | train_step fused samples/s (configs[2]) | train_msra samples/s (configs[3]) | sweep @16384 samples/s (configs[4]) |",
      "|---|---|---|---|---|---|---|---|---|---|"]
for n in (1, 2, 4, 8):
    b = bench if n == 1 else load("%s_bench_n%d.json" % (tag, n))
    if not b:
        continue
    sw = b["sweep"]["rows"][-1] if b.get("sweep") else None
    L.append("| %d | %.2f M | %.3f | %.1f %% | %.0f | %.1f %% | %.1f | %.0f | %.0f | %s |" % (
        n, b["value"] / 1e6, b["ms_per_step"], 100 * b["value"] / (n * bench["value"]), b["e2e"]["value"],
        100 * b["e2e"]["value"] / (n * bench["e2e"]["value"]), b["e2e"]["pcie_gbs_per_gpu"],
        b["train_step"]["fused"]["samples_per_s"], b["train_msra"]["fused"]["samples_per_s"],
        ("%.2f M" % (sw["samples_per_s_calls"] / 1e6)) if sw else "-"))
bal = load("%s_bench_n8_balanced.json" % tag)
if bal and bal.get("e2e") and bal["e2e"].get("equal_shards"):
    e = bal["e2e"]
    L += ["", "e2e at N = 8 with bandwidth-proportional shards of the global batch (`%s_bench_n8_balanced.json`, end of r2): "
          "%.0f samples/s against %.0f with equal shards on the same box (+%.0f %%), %.1f ms per step, shares %s of %d; "
          "%.0f GB/s of windows for the whole box - the host side of the box is the bound, unevenly shared between the GPUs."
          % (tag, e["value"], e["equal_shards"]["value"], 100 * (e["value"] / e["equal_shards"]["value"] - 1), e["ms_per_step"],
             e["sharding"]["samples_per_rank"], e["sharding"]["global_batch"],
             e["h2d_bytes_per_step_all_ranks"] / e["ms_per_step"] / 1e6)]
ts = bench["train_step"]
L += ["", "configs[2] at N = 1, batch 128 (ms per step): decoder + loss as the reference writes them (eager) %.2f, drop-in model %.2f, fused "
      "criterion %.2f - the cuDNN hourglass backbone (out of scope, unchanged) is ~98 %% of the step." % (
          ts["eager"]["ms_per_step"], ts["dropin"]["ms_per_step"], ts["fused"]["ms_per_step"])]

# ---- sweep table
L += ["", "## Inference sweep (BASELINE configs[4]: HAND17 shape, J = 21, one B200; `%s_bench_n1.json` -> `sweep`)\n" % tag,
      "| batch/GPU | ms (calls) | ms (graph) | samples/s (calls) | % HBM roofline (calls) | % HBM roofline (graph) |", "|---|---|---|---|---|---|"]
for r in bench["sweep"]["rows"]:
    L.append("| %d | %.3f | %.3f | %.3g | %.1f | %.1f |" % (r["batch_per_gpu"], r["ms_calls"], r["ms_graph"], r["samples_per_s_calls"],
                                                        100 * r["roofline_frac_calls"], 100 * r["roofline_frac_graph"]))

# ---- sanitizer
L += ["", "## compute-sanitizer\n"]
for tool in ("racecheck", "memcheck", "initcheck", "smoke", "lean_racecheck", "window_table_racecheck", "window_table_memcheck"):
    f = os.path.join(P, "%s_sanitizer_%s.log" % (tag, tool))
    if os.path.isfile(f):
        lines = [l.strip() for l in open(f) if "SUMMARY" in l or "passed" in l or "failed" in l]
        L.append("* %s: %s" % (tool, "; ".join(lines)))
# ---- pcie probe
pp = load("%s_pcie_probe.json" % tag)
if pp:
    L += ["", "## PCIe probe (one GPU, B = %d NYU-like windows; `tools/pcie_probe.cu`)\n" % pp["B"],
          "| method | ms | GB/s (useful) | samples/s |", "|---|---|---|---|"]
    for k, v in pp.items():
        if isinstance(v, dict) and "ms" in v:
            L.append("| %s | %.2f | %.1f | %s |" % (k, v["ms"], v.get("gbs_useful", v.get("gbs", 0)), ("%.0f" % v["samples_per_s"]) if "samples_per_s" in v else "-"))
p8 = sorted(glob.glob(os.path.join(P, "%s_pcie_probe_n8.json" % tag)))
if p8:
    d8 = json.load(open(p8[0]))
    L += ["", "Eight GPUs pulling at once (`cudaMemcpyAsync` of whole frames / window kernel), GB/s per GPU: " +
          ", ".join("GPU%s %.1f / %.1f" % (k, v["memcpy_full"], v["kernel_windows"]) for k, v in sorted(d8.items())) + "."]
open(os.path.join(P, "README.md"), "w").write("\n".join(L) + "\n")
print("\n".join(L))
