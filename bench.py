#!/usr/bin/env python
"""bench.py — SFR + decoder micro-benchmark (BASELINE.json configs[1]) plus the records of configs[2-4].

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = one pass of the hot path over one batch of synthetic NYU-shaped input
(B = 4096 samples per GPU, J = 14): SFR target build from 480x640 depth frames ->
last-stage decoder forward + stage loss + backward on N(0,1) logits (one pass) ->
batch reduction of dL/dw and the loss sums.  With N > 1 (torchrun) every rank owns
its own B samples (weak scaling); the only exchange is the all-reduce of dL/dw [J]
and the three logged loss terms, issued on a side stream so it overlaps the next
step's SFR build (what DDP's bucket does for the backbone).

Prints ONE JSON line (rank 0):
  value        whole-job samples/s with inputs resident in HBM, CUDA events, max over ranks
  e2e          the same path fed from pinned HOST memory through the public feed API
               (feed.HostFeed: annotations H2D + pwr_sfr_fetch pulling only the crop windows
               of the raw uint16 frames over PCIe, double-buffered against the compute
               stream), loss + decoded joints read back to the host every step
  roofline     dominant kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline the oracle port on the host cores (bounded sample)
  extras       inner_stage / two_stage_decoder (kernels of a 2-stage training step),
               raw_frames_step, two_kernel_step, sparse_targets, train_step (configs[2]),
               train_msra (configs[3]), sweep (configs[4])

`--impl reference` times the CPU oracle port (the reference is Python; it cannot
travel to the GPU box) on the same workload definition.
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "SFR+decoder samples/s"
UNIT = "samples/s"
NOMINAL_HBM_GBS = 8000.0      # report against the nominal B200 figure as well as the measured copy peak
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback
OUT = sys.stdout


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)       # >= 50 timed iterations after 10 warm-ups
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=4096, help="samples per GPU per step")
    ap.add_argument("--shape", default="NYU")
    ap.add_argument("--alpha", type=float, default=1.0)
    ap.add_argument("--frame-format", default="f32", choices=["f32", "nyu_gb16", "u16"],
                    help="format of the HBM-resident frames of the main timed region: f32 = decoded frames (what "
                         "process_single_data receives; SURVEY 8d's synthetic input); nyu_gb16 / u16 = raw sensor "
                         "samples decoded inside the SFR kernel with the load_from_text hand rectangle (8f-1)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--cpu-samples", type=int, default=512, help="samples in the bounded CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--augment", action="store_true",
                    help="build the SFR targets through the augmented branch (datasets.py:216-299, train.py defaults)")
    ap.add_argument("--watchdog", type=int, default=480,
                    help="seconds after the timed region at which rank 0 prints the line with whatever the extras "
                         "have measured by then and exits (0 = off)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e-balance", action="store_true",
                    help="N > 1: keep equal shards in the e2e leg (default: also try bandwidth-proportional shards)")
    ap.add_argument("--no-sparse", action="store_true", help="skip the variants of the step (two-kernel, compact, raw)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the PyTorch-eager decoder baseline on the GPU")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip inner_stage / two_stage_decoder / train_step / train_msra / sweep")
    ap.add_argument("--train-steps", type=int, default=8, help="timed steps of each training configuration")
    ap.add_argument("--sweep-batches", default="256,1024,4096,16384")
    return ap.parse_args()


def workload_name(shape, batch):
    return ("configs[1]: decoder+SFR microbenchmark alone, %s shape (J=%d, %dx%d frames), batch %d per GPU, "
            "SFR build + decoder fwd + fused bwd/loss" % (shape.name, shape.joints, shape.height, shape.width, batch))


def workload_config(shape, batch, frame_format, augment, alpha, lambda_h, lambda_d):
    """The workload both arms are quoted on: they print this dict unchanged, so the driver's `same_config` check
    compares like with like (arm-specific remarks live in `config_note`, outside it)."""
    from pixelwiseregression_b200 import roofline
    J = shape.joints
    return {"workload": workload_name(shape, batch), "batch_per_gpu": batch, "joints": J,
            "frame_format": frame_format, "augment": bool(augment), "alpha": alpha, "lambda_h": lambda_h,
            "lambda_d": lambda_d,
            "l2": "inputs larger than L2 (frames %.2f GB, logits 2 x %.2f GB per GPU); no flush needed"
                  % (batch * shape.height * shape.width * (4 if frame_format == "f32" else 2) / 1e9,
                     batch * J * 4096 * 4 / 1e9),
            "algorithmic_bytes_per_sample": roofline.step_one_pass_bytes(J),
            "last_stage": "one pass (pwr_decoder_fwd_bwd_loss): forward + loss + backward visit z, D and "
                          "the targets once; the two-kernel route of SURVEY 8d is timed in `two_kernel_step`"}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML every few ms
    while the timed region runs (the recipe's clocks line, in-process)."""
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index, period=0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.reasons, self.power = [], set(), []
        self.sm_max = None
        self.stop_flag = threading.Event()
        self.error = None
        self.marked = 0
        self.sections = {}

    def mark(self):
        """Start of the timed region: samples taken from here on are `samples_timed`."""
        self.marked = len(self.sm)

    def section(self, name):
        """Start of a named extra: the record of that extra carries its own clock summary."""
        self.sections[name] = [len(self.sm), None]
        return name

    def end_section(self, name):
        self.sections[name][1] = len(self.sm)
        lo, hi = self.sections[name]
        sm = self.sm[lo:hi]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": self.sm_max, "samples": len(sm), "reasons": sorted(self.reasons)}

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[0].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            while not self.stop_flag.is_set():
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                for name, bit in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                try:
                    self.power.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                self.stop_flag.wait(self.period)
        except Exception as exc:  # NVML missing: report it, do not fake numbers
            self.error = repr(exc)

    def summary(self, upto=None):
        sm = self.sm[:upto] if upto else self.sm
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["unavailable: %s" % self.error]}
        return {"sm_mhz": statistics.median(sm), "sm_min_mhz": min(sm), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "samples_timed": len(sm) - self.marked,
                "sm_mhz_timed": statistics.median(sm[self.marked:]) if len(sm) > self.marked else None,
                "power_w_max": max(self.power) if self.power else None}

    def stop(self):
        self.stop_flag.set()
        self.join(timeout=5)


# --------------------------------------------------------------------------- #
# reference arm: the CPU oracle port
# --------------------------------------------------------------------------- #
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline
    from pixelwiseregression_b200 import synth
    shape = synth.SHAPES[args.shape]
    n = args.cpu_samples
    cores = os.cpu_count() or 1
    import torch
    torch.set_num_threads(cores)
    d = cpu_baseline.prepare(shape, n, seed=0)                 # inputs are built outside the timed region
    for _ in range(max(1, args.warmup)):                       # warm-up passes (pool start-up, page-in)
        cpu_baseline.run_path(shape, d, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_baseline.run_path(shape, d, cores)
    dt = time.perf_counter() - t0
    cpu_baseline.shutdown()
    res = {"sample": cpu_baseline.describe(shape, n, cores)}
    value = n * args.steps / dt
    cfg = workload_config(shape, args.batch, "f32", False, args.alpha, 1.0, 0.01)
    cfg_note = ("each step is a bounded sample of %d samples of that workload on the host CPU; the port (fork pool over "
                   "all cores, OpenCV resize / blur) is FASTER than the stock reference path, whose process_single_data "
                   "takes 38-63 ms per sample per DataLoader worker (SURVEY section 6; measured in the build container)" % n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg, "config_note": cfg_note,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": res["sample"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=OUT, flush=True)


def eager_decoder_baseline(z, D, w, batch, alpha, lambda_h, lambda_d, iters=5):
    """model.py:83-95, 123-130, 151 and train.py:197-207 written as the reference writes them
    (eager PyTorch, autograd backward), timed with CUDA events on the same GPU."""
    import torch
    from torch.functional import F
    B, J, H, W = z.shape
    idx = torch.arange(W, device=z.device, dtype=torch.float64)
    coord = ((idx - W // 2) / (W - 1)).float()
    U = coord.view(1, 1, 1, W)
    V = coord.view(1, 1, H, 1)

    def one():
        zz = z.detach().requires_grad_(True)
        DD = D.detach().requires_grad_(True)
        ww = w.detach().requires_grad_(True)
        heat = F.softmax(ww * zz.view(B, J, -1), dim=2).view(B, J, H, W)
        u = torch.sum(U * heat, dim=(2, 3)).unsqueeze(-1)
        v = torch.sum(V * heat, dim=(2, 3)).unsqueeze(-1)
        rec = DD + batch.label_img
        mr = batch.mask * rec
        mh = heat * batch.mask
        d = (torch.sum(mh * mr, dim=(2, 3)) / (torch.sum(mh, dim=(2, 3)) + 1e-14)).unsqueeze(-1)
        uvd_out = torch.cat([torch.cat([u, v], dim=2), d], dim=2)
        hl = lambda_h * torch.mean(torch.sum((heat - batch.heatmaps) ** 2, dim=(2, 3)))
        dl = lambda_d * torch.mean(torch.sum((DD - batch.depthmaps) ** 2, dim=(2, 3)))
        ul = torch.mean(torch.sum((uvd_out - batch.uvd) ** 2, dim=2))
        (alpha * ul + (1 - alpha) * (hl + dl)).backward()

    for _ in range(2):
        one()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        one()
    e.record()
    torch.cuda.synchronize()
    return {"ms": s.elapsed_time(e) / iters, "iters": iters,
            "what": "reference decoder + loss lines (model.py:83-95,123-130,151; train.py:197-207) as eager "
                    "PyTorch ops with autograd backward, same GPU, same inputs"}


# --------------------------------------------------------------------------- #
# B200 arm
# --------------------------------------------------------------------------- #
def run_b200(args):
    import torch
    import torch.distributed as dist
    from pixelwiseregression_b200 import _lib, feed, ops, roofline, sfr, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    shape = synth.SHAPES[args.shape]
    B, J = args.batch, shape.joints
    alpha, lambda_h, lambda_d = args.alpha, 1.0, 0.01
    peak, peak_src = measured_peak()
    errors = {}                    # extras that failed (the main line is printed regardless)

    # ---- synthetic inputs, resident in HBM (seed = rank: every rank owns different samples) ----
    d = synth.make_frames_device(shape, B, seed=rank, device=dev)
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    z = torch.randn(B, J, 64, 64, device=dev, generator=g).requires_grad_(True)
    D = torch.randn(B, J, 64, 64, device=dev, generator=g).requires_grad_(True)
    w = (torch.rand(J, 1, device=dev, generator=g) + 0.5).requires_grad_(True)
    frames, com, cube, uvd = d["frames"], d["com"], d["cube"], d["uvd"]
    raw_fmt = "nyu_gb16" if shape.name == "NYU" else "u16"
    raw_frames = None
    if not shape.frame_f64:
        raw_frames = frames.round().clamp_(0, 65535).to(torch.int32).to(torch.uint16)      # sensor counts (mm)
    raw_kw = dict(fx=shape.fx, fy=shape.fy, frame_format=raw_fmt, prefilter=(40.0, shape.halfu, shape.halfv))
    sfr_kw = dict(fx=shape.fx, fy=shape.fy, frame_f64=shape.frame_f64)
    if args.frame_format != "f32":
        frames = raw_frames
        sfr_kw = dict(raw_kw, frame_format=args.frame_format)
    if args.augment:
        import numpy as np
        sfr_kw["augment"] = sfr.draw_augmentation(B, np.random.default_rng(rank))

    # the [J + 3] all-reduce rides on a side stream: it overlaps the NEXT step's SFR build, and the decoder
    # of the next step (which would read the updated temperature w) waits for it - DDP's overlap, in small
    comm_stream = torch.cuda.Stream(dev) if world > 1 else None
    pending = {"work": False}
    arenas = {}

    reduce_mode = {"n_mean": 0, "op": dist.ReduceOp.AVG if world > 1 else None}     # unequal shards: global n_mean + SUM
    use_taps = {"on": False}
    store_heat = {"on": True}      # False: the last stage as model.forward_loss runs it (heat maps never leave the kernel)

    def step(frames_, com_, cube_, uvd_, z_, D_, kw=None, key="main"):
        """The public-API call sequence a training loop makes for this path."""
        kw = kw or sfr_kw
        arena = arenas.setdefault(key, sfr.SfrArena())
        batch = frames_ if isinstance(frames_, sfr.SFRBatch) else sfr.build_sfr(frames_, com_, cube_, uvd_, arena=arena, **kw)
        # dense maps, or compact taps (targets="both" with use_taps: the caller keeps the dense tuple, the loss reads taps)
        heat_t = batch.taps if (batch.taps is not None and (batch.heatmaps is None or use_taps["on"])) else batch.heatmaps
        if world > 1 and pending["work"]:
            torch.cuda.current_stream().wait_stream(comm_stream)                  # w of the previous step is reduced
            pending["work"] = False
        total, terms, uvd_out = ops.fused_decoder_loss(z_, w, D_, batch.label_img, batch.mask, heat_t,
                                                       batch.depthmaps, batch.uvd, method="softmax", alpha=alpha,
                                                       lambda_h=lambda_h, lambda_d=lambda_d, store_heat=store_heat["on"],
                                                       n_mean=reduce_mode["n_mean"])[:3]
        z_.grad = D_.grad = w.grad = None
        total.backward()
        if world > 1:
            # what the DDP bucket carries for this path: dL/dw [J] and the logged loss terms [3]
            comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(comm_stream):
                dist.all_reduce(w.grad, op=reduce_mode["op"])
                dist.all_reduce(terms, op=reduce_mode["op"])
            w.grad.record_stream(comm_stream)
            terms.record_stream(comm_stream)
            pending["work"] = True
        return total, terms, uvd_out

    def barrier():
        if world > 1:
            torch.cuda.current_stream().wait_stream(comm_stream)
            pending["work"] = False
            dist.barrier()
        torch.cuda.synchronize()

    def sum_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(round(float(t.item())))

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # clocks are sampled from the warm-up on (same load as the timed region, which may last
    # only tens of milliseconds); samples inside the timed region are counted separately
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step(frames, com, cube, uvd, z, D)
    barrier()

    # ---- timed region: K steps, device-timed, per-kernel events on the launching stream ----
    sampler.mark()
    launches0 = _lib.launch_count()
    _lib.PROFILE = []
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    marks[0].record()
    t_issue = time.perf_counter()
    for k in range(args.steps):
        step(frames, com, cube, uvd, z, D)
        marks[k + 1].record()
    issue_ms = (time.perf_counter() - t_issue) * 1e3 / args.steps
    barrier()
    elapsed_ms = marks[0].elapsed_time(marks[-1])
    step_ms = [marks[k].elapsed_time(marks[k + 1]) for k in range(args.steps)]
    launches = (_lib.launch_count() - launches0) * world      # every rank launches the same sequence
    prof, _lib.PROFILE = _lib.PROFILE, None
    clocks = sampler.summary()
    elapsed_ms = max_over_ranks(elapsed_ms)
    value = B * world * args.steps / (elapsed_ms * 1e-3)
    median_ms = max_over_ranks(statistics.median(step_ms))

    kernel_ms = {}
    for name, s, e in prof:
        kernel_ms.setdefault(name, []).append(s.elapsed_time(e))
    # device time between the big kernels (small kernels, launch gaps), averaged per step
    gap_ms = sum(prof[i][2].elapsed_time(prof[i + 1][1]) for i in range(len(prof) - 1)) / args.steps
    per_launch_bytes = {"pwr_sfr_build": roofline.sfr_build_bytes(J) * B,
                        "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(J) * B,
                        "pwr_decoder_fwd": roofline.decoder_fwd_bytes(J) * B,
                        "pwr_decoder_bwd_loss": roofline.decoder_bwd_bytes(J) * B}
    step_bytes = roofline.step_one_pass_bytes(J) if "pwr_decoder_fwd_bwd_loss" in kernel_ms else roofline.step_bytes(J)
    kernels = {}
    for name, ms in kernel_ms.items():
        avg = sum(ms) / len(ms)
        gbs = per_launch_bytes[name] / (avg * 1e-3) / 1e9
        kernels[name] = {"avg_ms": avg, "median_ms": statistics.median(ms), "launches": len(ms),
                         "algorithmic_bytes": per_launch_bytes[name], "achieved_gbs": gbs, "frac": gbs / peak}
    dominant = max(kernels, key=lambda k: kernels[k]["avg_ms"])
    step_kernel_ms = sum(k["avg_ms"] for k in kernels.values())
    # Every extra starts as None and the line is built by a closure: if an extra hangs (a rank-local failure inside a
    # collective at N > 1), rank 0's watchdog still prints the main line with what was measured so far.
    e2e = e2e_r1 = two_kernel = sparse = raw_step = raw_sparse = no_heat = both = inner = two_stage = graph_step = None
    train_step = train_msra = sweep = cpu = gpu_eager = None
    def build_line():
        dk = kernels[dominant]
        cfg = workload_config(shape, B, args.frame_format, args.augment, alpha, lambda_h, lambda_d)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "clocks": clocks,
            "e2e": e2e,
            "ms_per_step_median": median_ms, "value_median": B * world / (median_ms * 1e-3),
            "e2e_whole_frames_r1_definition": e2e_r1,
            "two_kernel_step": two_kernel,
            "sparse_targets": sparse,
            "raw_frames_step": raw_step, "raw_frames_sparse_targets": raw_sparse,
            "no_heat_store_step": no_heat, "dense_tuple_compact_loss": both,
            "inner_stage": inner,
            "two_stage_decoder": two_stage,
            "graph_step": graph_step,
            "train_step": train_step,
            "train_msra": train_msra,
            "sweep": sweep,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": dk["achieved_gbs"], "peak": peak,
                         "unit": "GB/s", "frac": dk["frac"], "traffic": None, "peak_source": peak_src,
                         "nominal_peak": NOMINAL_HBM_GBS, "frac_of_nominal": dk["achieved_gbs"] / NOMINAL_HBM_GBS,
                         "avg_launch_ms": dk["avg_ms"], "algorithmic_bytes_per_launch": dk["algorithmic_bytes"],
                         "share_of_step": dk["avg_ms"] / step_kernel_ms if step_kernel_ms else None},
            "kernels": kernels,
            "step_roofline_frac": (step_bytes * B / (elapsed_ms / args.steps * 1e-3) / 1e9) / peak,
            "step_roofline_frac_of_nominal": (step_bytes * B / (elapsed_ms / args.steps * 1e-3) / 1e9) / NOMINAL_HBM_GBS,
            "host_issue_ms_per_step": issue_ms,
            "between_kernels_ms_per_step": gap_ms,
            "cpu_baseline": cpu,
            "gpu_eager_decoder": gpu_eager,
            "errors": errors or None,
        }
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(traffic_file):
            try:
                with open(traffic_file) as f:
                    tr = json.load(f)
                line["roofline"]["traffic"] = tr.get(dominant, {}).get("dram_bytes_per_launch")
            except Exception:
                pass
        return line

    watchdog = None
    if rank == 0 and args.watchdog > 0:
        def fire():
            try:
                line_ = build_line()
                line_["errors"] = dict(errors, watchdog="extras did not finish within %d s after the timed region; "
                                                        "the line holds what was measured until then" % args.watchdog)
                print(json.dumps(line_, default=str), file=OUT, flush=True)
            finally:
                os._exit(0)
        watchdog = threading.Timer(args.watchdog, fire)
        watchdog.daemon = True
        watchdog.start()

    def kernel_table(prof_v, bytes_table):
        kms = {}
        for name, s_, e_ in prof_v:
            kms.setdefault(name, []).append(s_.elapsed_time(e_))
        return {k: {"avg_ms": sum(v) / len(v), "algorithmic_bytes": bytes_table[k] * B,
                    "frac": bytes_table[k] * B / (sum(v) / len(v) * 1e-3) / 1e9 / peak}
                for k, v in kms.items() if k in bytes_table}

    # ---- variants of the same step, each against ITS OWN algorithmic bytes ("report the
    # elided variant separately, never against the larger figure"); identical loss terms and gradients
    # (tests/test_gpu_decoder.py) ----
    def timed_variant(frames_v, kw, one_pass, bytes_table, sample_bytes, note, key):
        ops.ONE_PASS_LAST_STAGE = one_pass
        try:
            for _ in range(3):
                step(frames_v, com, cube, uvd, z, D, kw, key)
            barrier()
            _lib.PROFILE = []
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.steps):
                step(frames_v, com, cube, uvd, z, D, kw, key)
            s1.record()
            barrier()
            prof_v, _lib.PROFILE = _lib.PROFILE, None
        finally:
            ops.ONE_PASS_LAST_STAGE = True
        ms_v = max_over_ranks(s0.elapsed_time(s1)) / args.steps
        arenas.pop(key, None)
        return {"value": B * world / (ms_v * 1e-3), "unit": UNIT, "ms_per_step": ms_v,
                "algorithmic_bytes_per_sample": sample_bytes,
                "step_roofline_frac": sample_bytes * B / (ms_v * 1e-3) / 1e9 / peak,
                "kernels": kernel_table(prof_v, bytes_table), "note": note}

    # ---- the kernels of an INNER stage (north_star's training is 2 stages, model.py:200-210): forward with the
    # loss riding along, and the backward with dense upstream gradients on the heat maps and depth maps ----
    def time_launch(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(iters):
            fn()
        s1.record()
        torch.cuda.synchronize()
        return max_over_ranks(s0.elapsed_time(s1) / iters)

    def entry(ms, bytes_per_sample, what):
        gbs = bytes_per_sample * B / (ms * 1e-3) / 1e9
        return {"avg_ms": ms, "algorithmic_bytes_per_sample": bytes_per_sample, "achieved_gbs": gbs, "frac": gbs / peak,
                "what": what}

    two_kernel = sparse = raw_step = raw_sparse = no_heat = both = None
    try:
        if not args.no_sparse:
            # (1) SURVEY 8d's own accounting: forward kernel, then backward+loss kernel (the logits are read twice)
            two_kernel = timed_variant(
                frames, None, False,
                {"pwr_sfr_build": roofline.sfr_build_bytes(J), "pwr_decoder_fwd": roofline.decoder_fwd_bytes(J),
                 "pwr_decoder_bwd_loss": roofline.decoder_bwd_bytes(J)}, roofline.step_bytes(J),
                "SURVEY 8d accounting: pwr_decoder_fwd then pwr_decoder_bwd_loss (ops.ONE_PASS_LAST_STAGE = False), "
                "622 780 + 721 120 + 1 409 024 B/sample", "two")
            # (2) compact targets: the SFR builder emits 64 B of taps per joint instead of two dense maps and the
            # loss kernel evaluates the heat-map / depth-map targets on the fly
            sparse = timed_variant(
                frames, dict(sfr_kw, targets="sparse"), True,
                {"pwr_sfr_build": roofline.sfr_build_sparse_bytes(J),
                 "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(J, sparse=True)},
                roofline.step_one_pass_bytes(J, sparse=True),
                "same step, same results; targets handed to the loss kernel as 64-byte taps per joint instead of two "
                "dense 16 KiB maps (sfr.build_sfr(targets='sparse'))", "sparse")
            # (2b) the last stage exactly as PixelwiseRegression.forward_loss runs it: nothing but the loss consumes its heat
            # maps, so they are not stored (6J + 2 maps instead of 7J + 2)
            store_heat["on"] = False
            try:
                no_heat = timed_variant(
                    frames, None, True,
                    {"pwr_sfr_build": roofline.sfr_build_bytes(J),
                     "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(J, store_heat=False)},
                    roofline.sfr_build_bytes(J) + roofline.decoder_fused_bytes(J, store_heat=False),
                    "same step without the heat-map store of the last stage (store_heat=False, what model.forward_loss "
                    "does: train.py:192-207 only feeds the last stage's heat maps to the loss)", "noheat")
            finally:
                store_heat["on"] = True
            # (2c) the dense tuple for the caller AND the compact taps for the loss: the reference's 9-tuple is still
            # written in full, but the loss kernel does not read the 2J target maps back
            use_taps["on"] = True
            try:
                both = timed_variant(
                    frames, dict(sfr_kw, targets="both"), True,
                    {"pwr_sfr_build": roofline.sfr_build_bytes(J) + 64 * J,
                     "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(J, sparse=True)},
                    roofline.sfr_build_bytes(J) + 64 * J + roofline.decoder_fused_bytes(J, sparse=True),
                    "sfr.build_sfr(targets='both'): dense heat maps / depth maps written for the caller as the reference "
                    "returns them, the loss evaluated from the 64-byte taps (the 2J dense target maps are never read back)",
                    "both")
            finally:
                use_taps["on"] = False
            # (3) the step fed with the raw 16-bit sensor frames (half the source bytes; PNG decode + hand rectangle of
            # load_from_text inside the SFR kernel), dense targets, against SURVEY 8d's bytes
            if raw_frames is not None and args.frame_format == "f32" and not args.augment:
                raw_step = timed_variant(
                    raw_frames, raw_kw, True,
                    {"pwr_sfr_build": roofline.sfr_build_bytes(J), "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(J)},
                    roofline.step_one_pass_bytes(J),
                    "same step on raw uint16 %s frames resident in HBM, decoded inside the SFR kernel with the "
                    "load_from_text hand rectangle (SURVEY 8f-1)" % raw_fmt, "raw")
                # (3b) both reductions of traffic at once: raw 16-bit frames in, compact targets out
                raw_sparse = timed_variant(
                    raw_frames, dict(raw_kw, targets="sparse"), True,
                    {"pwr_sfr_build": roofline.sfr_build_sparse_bytes(J),
                     "pwr_decoder_fwd_bwd_loss": roofline.decoder_fused_bytes(J, sparse=True)},
                    roofline.step_one_pass_bytes(J, sparse=True),
                    "raw uint16 %s frames (decode + hand rectangle in the SFR kernel) AND compact targets" % raw_fmt,
                    "raw_sparse")

    except Exception as exc:      # an extra must never cost the main JSON line
        errors['variants'] = repr(exc)
        sys.stderr.write("bench.py: variants failed: %r\n" % (exc,))
    # ---- the same step captured once and replayed as a CUDA graph: the library never allocates or synchronises, so
    # SFR build + one-pass last stage + loss / dL/dw reduction + the (unit) upstream scale replay with ~10 us of host
    # work per step; `host_issue_ms_per_step` above is the price of issuing the same launches from Python ----
    graph_step = None
    try:
        if not args.no_extras:
            ones = torch.ones((), device=dev)
            g_arena = sfr.SfrArena()
            zd, Dd, wd = z.detach(), D.detach(), w.detach()

            def graph_body():
                bt = sfr.build_sfr(frames, com, cube, uvd, arena=g_arena, **sfr_kw)
                H, uvd_o, gz, gD, gwp, lp = ops.decoder_fused_raw(zd, wd, Dd, bt.label_img, bt.mask,
                                                                 (bt.heatmaps, bt.depthmaps, bt.uvd), "softmax", alpha,
                                                                 lambda_h, lambda_d)
                out4, gw = ops.stage_loss(lp, lambda_h, lambda_d, alpha, 0, gwp)
                ops.scale_inplace_(gz, ones, gD, gw)
                return out4, gw, uvd_o, gz, gD

            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                graph_body()
            torch.cuda.current_stream().wait_stream(side)
            cuda_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(cuda_graph):
                g_out = graph_body()
            for _ in range(3):
                cuda_graph.replay()
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            t_h = time.perf_counter()
            for _ in range(args.steps):
                cuda_graph.replay()
            host_ms = (time.perf_counter() - t_h) * 1e3 / args.steps
            s1.record()
            barrier()
            ms_g = max_over_ranks(s0.elapsed_time(s1)) / args.steps
            # calls without per-kernel profiling events, for the plain host-issue figure
            for _ in range(3):
                step(frames, com, cube, uvd, z, D)
            barrier()
            t_h = time.perf_counter()
            for _ in range(args.steps):
                step(frames, com, cube, uvd, z, D)
            host_calls_ms = (time.perf_counter() - t_h) * 1e3 / args.steps
            barrier()
            graph_step = {"value": B * world / (ms_g * 1e-3), "unit": UNIT, "ms_per_step": ms_g,
                          "host_issue_ms_per_step": host_ms, "host_issue_ms_per_step_calls": host_calls_ms,
                          "loss": float(g_out[0][3]),
                          "note": "sfr.build_sfr(arena=) -> ops.decoder_fused_raw -> ops.stage_loss(+ dL/dw) -> ops.scale_inplace_ "
                                  "captured once with torch.cuda.graph and replayed; `host_issue_ms_per_step_calls` = the autograd "
                                  "step of the main region issued call by call with the per-kernel profiling events off"}
            del cuda_graph, g_out, g_arena

    except Exception as exc:      # an extra must never cost the main JSON line
        errors['graph_step'] = repr(exc)
        sys.stderr.write("bench.py: graph_step failed: %r\n" % (exc,))
    inner = two_stage = None
    try:
        if not args.no_extras:
            sampler.section("inner_stage")
            batch = sfr.build_sfr(frames, com, cube, uvd, **sfr_kw)
            zd, Dd, wd = z.detach(), D.detach(), w.detach()
            dense_t = (batch.heatmaps, batch.depthmaps, batch.uvd)
            gH_up = torch.randn(B, J, 64, 64, device=dev, generator=g) * 1e-4
            gD_up = torch.randn(B, J, 64, 64, device=dev, generator=g) * 1e-4
            g_uvd = torch.randn(B, J, 3, device=dev, generator=g) * 1e-3
            one = torch.ones((), device=dev)
            it = max(10, args.steps // 2)
            _, uvd_f, stats_f, _ = ops.decoder_forward_raw(zd, wd, Dd, batch.label_img, batch.mask, targets=dense_t)
            ms_fwd = time_launch(lambda: ops.decoder_forward_raw(zd, wd, Dd, batch.label_img, batch.mask, targets=dense_t), it)
            ms_bwd_a1 = time_launch(lambda: ops.decoder_backward_raw(
                zd, wd, Dd, batch.label_img, batch.mask, stats_f, uvd_f, g_uvd, gH_up, gD_up, targets=dense_t, alpha=1.0,
                loss_scale_dev=one), it)
            ms_bwd_a05 = time_launch(lambda: ops.decoder_backward_raw(
                zd, wd, Dd, batch.label_img, batch.mask, stats_f, uvd_f, g_uvd, gH_up, gD_up, targets=dense_t, alpha=0.5,
                loss_scale_dev=one), it)
            inner = {
                "pwr_decoder_fwd+loss": entry(ms_fwd, roofline.decoder_fwd_bytes(J, with_targets=True),
                                              "inner-stage forward, H stored, loss value from dense targets (decoder_fwd_kernel)"),
                "pwr_decoder_bwd_loss alpha=1": entry(ms_bwd_a1, roofline.decoder_bwd_bytes(J, with_targets=False, upstream_maps=True),
                                                      "inner-stage backward, dense gH_up / gD_up, map targets carry no weight "
                                                      "and are not read (train.py default alpha = 1): 6J + 2 maps"),
                "pwr_decoder_bwd_loss alpha=0.5": entry(ms_bwd_a05, roofline.decoder_bwd_bytes(J, with_targets=True, upstream_maps=True),
                                                        "inner-stage backward, dense targets AND dense gH_up / gD_up: SURVEY 8d's "
                                                        "131 072 J + 32 768 B (six-slot pipelined kernel)"),
                "clocks": sampler.end_section("inner_stage"),
            }
            # the decoder work of one 2-stage training step, in forward_loss's order (model.py:200-210, train.py:192-207):
            # stage-0 forward+loss, last stage in one pass, stage-0 backward with the dense gradients the next stage's
            # conv would hand back (synthetic here: the conv backbone is out of scope)
            sampler.section("two_stage_decoder")

            def two_stage_pass():
                ops.decoder_forward_raw(zd, wd, Dd, batch.label_img, batch.mask, targets=dense_t)
                ops.decoder_fused_raw(zd, wd, Dd, batch.label_img, batch.mask, dense_t, "softmax", alpha, store_heat=False)
                ops.decoder_backward_raw(zd, wd, Dd, batch.label_img, batch.mask, stats_f, uvd_f, None, gH_up, gD_up,
                                         targets=dense_t, alpha=alpha, loss_scale_dev=one)

            ms_two = time_launch(two_stage_pass, it)
            bytes_two = (roofline.decoder_fwd_bytes(J, with_targets=True) + roofline.decoder_fused_bytes(J, store_heat=False) +
                         roofline.decoder_bwd_bytes(J, with_targets=(alpha != 1.0), upstream_maps=True))
            two_stage = entry(ms_two, bytes_two, "decoder kernels of a 2-stage training step at alpha = %g: stage-0 forward+loss, "
                              "last stage in one pass (no H store), stage-0 backward with dense upstream maps" % alpha)
            two_stage["samples_per_s"] = B * world / (ms_two * 1e-3)
            two_stage["clocks"] = sampler.end_section("two_stage_decoder")
            del batch, gH_up, gD_up, stats_f, uvd_f
            torch.cuda.empty_cache()

    except Exception as exc:      # an extra must never cost the main JSON line
        errors['inner_stage'] = repr(exc)
        sys.stderr.write("bench.py: inner_stage failed: %r\n" % (exc,))
    # ---- end to end: inputs in pinned host memory, H2D + D2H inside the timed region ----
    e2e_pool = {}

    def e2e_inputs(cap):
        """Pinned host copies of `cap` samples (the rank's B synthetic samples, repeated beyond B) and logits for them."""
        if e2e_pool.get("cap", 0) >= cap:
            return e2e_pool
        e2e_pool.clear()
        torch.cuda.empty_cache()
        fmt = raw_fmt if raw_frames is not None else "f32"
        src = raw_frames if raw_frames is not None else frames
        host_frames = torch.empty((cap,) + tuple(src.shape[1:]), dtype=src.dtype, pin_memory=True)
        for lo in range(0, cap, B):
            n = min(B, cap - lo)
            host_frames[lo:lo + n].copy_(src[:n])
        idx = torch.arange(cap) % B
        e2e_pool.update(cap=cap, fmt=fmt, host_frames=host_frames,
                        com=com.cpu()[idx].numpy(), cube=cube.cpu()[idx].numpy(), uvd=uvd.cpu()[idx].numpy(),
                        z=z.detach() if cap == B else z.detach()[idx.to(dev)],
                        D=D.detach() if cap == B else D.detach()[idx.to(dev)])
        return e2e_pool

    def run_e2e(Br=None, n_steps=None, n_mean=0):
        """Public feed API: raw uint16 frames + annotations live in pinned host memory; every step the annotations are
        copied and pwr_sfr_fetch pulls the crop windows over PCIe on the copy stream while the previous batch is being
        built and decoded; the loss vector and the decoded joints are read back and waited for every step.  The
        logits stand in for the conv backbone's outputs, which only ever exist on the device.
        `Br`: this rank's share of the global batch (default: the equal share B); with unequal shares `n_mean` =
        global B*J keeps every loss the global-batch mean and the cross-rank reduction is a SUM."""
        Br = B if Br is None else int(Br)
        pool = e2e_inputs(max(Br, B))
        fmt, host_frames = pool["fmt"], pool["host_frames"][:Br]
        host = {"com": pool["com"][:Br], "cube": pool["cube"][:Br], "uvd": pool["uvd"][:Br]}
        pf = (40.0, shape.halfu, shape.halfv) if raw_frames is not None else None
        hf = feed.HostFeed(shape, Br, frame_format=fmt, prefilter=pf, device=dev, timing=True)
        out_host = torch.empty(4 + Br * J * 3, pin_memory=True)
        done = torch.cuda.Event()
        zz, DD = pool["z"][:Br].detach().requires_grad_(True), pool["D"][:Br].detach().requires_grad_(True)
        reduce_mode.update(n_mean=int(n_mean), op=dist.ReduceOp.SUM if n_mean else dist.ReduceOp.AVG)

        def submit():
            return hf.submit(host_frames, host["com"], host["cube"], host["uvd"])

        def consume(t):
            batch = hf.build(t)
            total, terms, uvd_out = step(batch, None, None, None, zz, DD, None, "e2e_%d" % Br)
            if world > 1:                                    # the logged terms are being averaged on the side stream
                torch.cuda.current_stream().wait_stream(comm_stream)
                pending["work"] = False
            out_host[:1].copy_(total.detach().reshape(1), non_blocking=True)
            out_host[1:4].copy_(terms, non_blocking=True)
            out_host[4:].copy_(uvd_out.reshape(-1), non_blocking=True)
            done.record()

        try:
            n_e2e = n_steps or args.e2e_steps or min(args.steps, 20)
            t = submit()
            for _ in range(3):                                   # warm-up with the pipeline running
                nxt = submit()
                consume(t)
                done.synchronize()
                t = nxt
            fetched = hf.fetched_bytes(t)
            barrier()
            host_ms = [0.0, 0.0, 0.0]                            # host time in submit / consume / waiting for the results
            fetch_ms = 0.0
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                ta = time.perf_counter()
                nxt = submit()                                   # batch k+1 starts crossing PCIe
                tb = time.perf_counter()
                consume(t)                                       # batch k: build + decode + loss + backward + D2H
                tc = time.perf_counter()
                done.synchronize()                               # the host holds loss and joints of batch k
                td = time.perf_counter()
                host_ms[0] += (tb - ta) * 1e3; host_ms[1] += (tc - tb) * 1e3; host_ms[2] += (td - tc) * 1e3
                fetch_ms += hf.fetch_ms(t)                       # long finished: batch k was built from it
                t = nxt
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            barrier()
            dt = max_over_ranks(dt)
        finally:
            reduce_mode.update(n_mean=0, op=dist.ReduceOp.AVG if world > 1 else None)
            arenas.pop("e2e_%d" % Br, None)
        loss_host = float(out_host[0])
        h2d = fetched + hf.h2d_bytes_small
        total_samples = sum_over_ranks(Br)
        res = {"value": total_samples * n_e2e / dt, "unit": UNIT, "h2d_bytes_per_step": h2d,
               "h2d_bytes_per_step_all_ranks": sum_over_ranks(h2d),
               "d2h_bytes_per_step": out_host.numel() * 4, "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3,
               "pcie_gbs_per_gpu": h2d / (dt / n_e2e) / 1e9, "window_hw": list(hf.win_hw), "loss_read_back": loss_host,
               "host_ms_per_step": {"submit": host_ms[0] / n_e2e, "consume": host_ms[1] / n_e2e,
                                    "wait_for_results": host_ms[2] / n_e2e},
               "fetch_ms_per_step": fetch_ms / n_e2e, "batch_this_rank": Br,
               "frames_in_host_memory_bytes": host_frames.numel() * host_frames.element_size(),
               "note": "per rank and step: com + cube + uvd (%d B) copied from pinned host memory and the crop windows of "
                       "the raw %s frames (%d B of the %d B the frames occupy) pulled over PCIe by pwr_sfr_fetch on a "
                       "copy stream, one batch ahead of the compute stream (feed.HostFeed); loss[4] + uvd[B,J,3] read back "
                       "and waited for every step; the logits z, D stay on the device (they are the conv backbone's outputs); "
                       "wall clock, max over ranks; byte counts are rank 0's" % (hf.h2d_bytes_small, fmt, fetched,
                                                       host_frames.numel() * host_frames.element_size())}
        del hf
        return res

    def run_e2e_balanced(equal):
        """N > 1: the host links of a multi-GPU box are not equal (GPUs behind a shared upstream port halve each other)
        and a lock-step job runs at the pace of the slowest one.  Give every rank a share of the SAME global batch
        proportional to the rate its own fetches sustained (distributed.proportional_shards; three short rounds, each
        measured under the previous round's shares), then time the feed again.  Same total work per step (N x B
        samples), global-batch-mean loss through n_mean.  Returns None when it does not beat equal shares."""
        from pixelwiseregression_b200 import distributed as pdist
        n_mean = world * B * J
        Br, fetch = B, equal["fetch_ms_per_step"]
        history = []
        for _ in range(3):
            rates = pdist.gather_rates(Br / max(fetch, 1e-6), dev)
            shards = pdist.proportional_shards(rates, world * B, 64, B // 2, B + B // 2)
            history.append(shards)
            Br = shards[rank]
            probe = run_e2e(Br, n_steps=4, n_mean=n_mean)
            fetch = probe["fetch_ms_per_step"]
        if history[-1] == [B] * world:           # the links are equal (same list on every rank): nothing to gain
            return None
        res = run_e2e(Br, n_mean=n_mean)
        res["sharding"] = {"mode": "bandwidth-proportional", "samples_per_rank": history[-1], "rounds": history,
                           "global_batch": world * B,
                           "what": "every rank's share of the global batch is proportional to the PCIe rate its own window "
                                   "fetches sustained in the previous round (HostFeed.fetch_ms -> "
                                   "distributed.proportional_shards); loss = global-batch mean (n_mean), gradients SUM-reduced"}
        return res if res["value"] > equal["value"] else None

    def run_e2e_whole_frames(n_steps=3):
        """Round 1's definition, kept for continuity: whole float32 frames AND the logits cross PCIe, serially."""
        src = dict(frames=frames, com=com, cube=cube, uvd=uvd, z=z.detach(), D=D.detach())
        host = {k: torch.empty(v.shape, dtype=v.dtype, pin_memory=True) for k, v in src.items()}
        for k, v in src.items():
            host[k].copy_(v)
        out_host = {"loss": torch.empty(4, pin_memory=True), "uvd": torch.empty(B, J, 3, pin_memory=True)}
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        dev_in = {k: torch.empty_like(v, device=dev) for k, v in host.items()}

        def e2e_step():
            for k in host:
                dev_in[k].copy_(host[k], non_blocking=True)
            z_ = dev_in["z"].requires_grad_(True)
            D_ = dev_in["D"].requires_grad_(True)
            total, terms, uvd_out = step(dev_in["frames"], dev_in["com"], dev_in["cube"], dev_in["uvd"], z_, D_, None, "e2e_r1")
            out_host["loss"][:1].copy_(total.detach().reshape(1), non_blocking=True)
            out_host["loss"][1:].copy_(terms, non_blocking=True)
            out_host["uvd"].copy_(uvd_out, non_blocking=True)
            torch.cuda.synchronize()
            dev_in["z"].requires_grad_(False)
            dev_in["D"].requires_grad_(False)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_steps):
            e2e_step()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        arenas.pop("e2e_r1", None)
        return {"value": B * world * n_steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "steps": n_steps,
                "note": "round 1's e2e definition: whole %s frames + com + cube + uvd + the logits z, D copied serially "
                        "from pinned host memory every step (no overlap, no windows)" % args.frame_format}

    e2e = e2e_r1 = None
    try:
        if not args.no_e2e:
            sampler.section("e2e")
            if world > 1 and not args.no_e2e_balance:
                e2e_inputs(B + B // 2)                       # one pinned pool for both legs
            e2e = run_e2e()
            e2e["sharding"] = {"mode": "equal", "samples_per_rank": [B] * world, "global_batch": world * B}
            if world > 1 and not args.no_e2e_balance:
                try:
                    bal = run_e2e_balanced(e2e)
                    if bal is not None:
                        bal["equal_shards"] = {k: e2e[k] for k in ("value", "ms_per_step", "pcie_gbs_per_gpu", "fetch_ms_per_step")}
                        e2e = bal
                except Exception as exc:      # keep the equal-shard figure
                    errors['e2e_balanced'] = repr(exc)
                    sys.stderr.write("bench.py: e2e_balanced failed: %r\n" % (exc,))
                    barrier()
            e2e["clocks"] = sampler.end_section("e2e")
            e2e_pool.clear()
            torch.cuda.synchronize()
            if hasattr(torch._C, "_host_emptyCache"):        # hand the pinned pool back before the next leg pins its own
                torch._C._host_emptyCache()
            if not args.no_extras:
                e2e_r1 = run_e2e_whole_frames()

    except Exception as exc:      # an extra must never cost the main JSON line
        errors['e2e'] = repr(exc)
        sys.stderr.write("bench.py: e2e failed: %r\n" % (exc,))
    # ---- GPU baseline of configs[1] ("vs reference PyTorch path"): the reference's decoder + loss
    # lines as plain eager PyTorch ops on the same GPU and inputs (the SFR builder has no GPU
    # reference: upstream it is CPU-only) ----
    gpu_eager = None
    try:
        if rank == 0 and world == 1 and not args.no_gpu_eager:
            batch = sfr.build_sfr(frames, com, cube, uvd, **sfr_kw)
            gpu_eager = eager_decoder_baseline(z, D, w, batch, alpha, lambda_h, lambda_d)
            gpu_eager["fused_ms"] = sum(v["avg_ms"] for k_, v in kernels.items() if k_.startswith("pwr_decoder"))
            gpu_eager["speedup"] = gpu_eager["ms"] / gpu_eager["fused_ms"]
            del batch
            torch.cuda.empty_cache()

    except Exception as exc:      # an extra must never cost the main JSON line
        errors['gpu_eager_decoder'] = repr(exc)
        sys.stderr.write("bench.py: gpu_eager_decoder failed: %r\n" % (exc,))
    # ---- the records of BASELINE configs[2], [3], [4] ----
    train_step = train_msra = sweep = None
    try:
        if not args.no_extras:
            del frames, raw_frames, z, D, d
            arenas.clear()
            torch.cuda.empty_cache()
            from examples import train_synthetic as ts
            from tools import sweep_inference
            # configs[2]: NYU-shape end-to-end training step (train.py:158-208), batch 128 per GPU, DDP over NCCL when
            # N > 1; the same step with the decoder + loss as the reference writes them (eager), with the drop-in
            # model (fused decoder kernels, reference loss lines), and with the fused criterion
            sampler.section("train_step")
            train_step = {"config": "configs[2]: NYU shape (J=14), batch 128 per GPU, 2 stages, features 128, level 4, "
                                    "InstanceNorm, AdamW; on-GPU SFR build + cuDNN backbone + decoder + loss; DDP over NCCL "
                                    "when n_gpus > 1", "n_gpus": world, "batch_per_gpu": 128}
            for mode in ("eager", "dropin", "fused"):
                train_step[mode] = ts.run_training(synth.NYU, 128, args.train_steps, 3, mode, world=world, rank=rank,
                                                   local=local_rank)
            train_step["speedup_fused_over_eager"] = train_step["eager"]["ms_per_step"] / train_step["fused"]["ms_per_step"]
            train_step["clocks"] = sampler.end_section("train_step")
            # configs[3]: MSRA shape (J = 21, 240x320 float64-semantics frames, centre-of-mass fallback, cube 125)
            sampler.section("train_msra")
            train_msra = {"config": "configs[3]: MSRA shape (J=21, float64 frame semantics, CoM from the frame), batch 128 per "
                                    "GPU, fused criterion, DDP over NCCL when n_gpus > 1", "n_gpus": world, "batch_per_gpu": 128}
            train_msra["fused"] = ts.run_training(synth.MSRA, 128, args.train_steps, 3, "fused", world=world, rank=rank,
                                                  local=local_rank)
            train_msra["clocks"] = sampler.end_section("train_msra")
            # configs[4]: HAND17-shape inference sweep (test.py:93-124 around the backbone), every rank its own replica
            sampler.section("sweep")
            rows = sweep_inference.sweep(synth.HAND17, [int(b) for b in args.sweep_batches.split(",")], 10, 3, world, rank,
                                         local_rank, peak)
            sweep = {"config": "configs[4]: HAND17 shape (J=21), test-only SFR + decoder forward without the heat-map store + "
                               "recover_uvd, per-replica batch swept, %d independent replica(s)" % world,
                     "rows": [{k: r[k] for k in ("batch_per_gpu", "ms_calls", "ms_graph", "samples_per_s_calls",
                                                 "samples_per_s_graph", "roofline_frac_calls", "roofline_frac_graph",
                                                 "graph_equals_calls")} for r in rows],
                     "clocks": sampler.end_section("sweep")}

    except Exception as exc:      # an extra must never cost the main JSON line
        errors['configs_2_3_4'] = repr(exc)
        sys.stderr.write("bench.py: configs_2_3_4 failed: %r\n" % (exc,))
    # ---- CPU baseline: oracle port on the host cores (rank 0, N = 1 only) ----
    cpu = None
    try:
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            from oracle import cpu_baseline
            res = cpu_baseline.time_path(shape, args.cpu_samples, seed=0, repeats=2)
            cpu = {"value": res["samples_per_s"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                   "sample": res["sample"] + "; best of 2 passes"}
    except Exception as exc:      # an extra must never cost the main JSON line
        errors['cpu_baseline'] = repr(exc)
        sys.stderr.write("bench.py: cpu_baseline failed: %r\n" % (exc,))
    sampler.stop()

    if rank == 0:
        if watchdog is not None:
            watchdog.cancel()
        line = build_line()
        print(json.dumps(line), file=OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def claim_stdout():
    """Keep stdout for the ONE JSON line: everything else a library may print to fd 1
    (e.g. NCCL's version banner) is routed to stderr."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse_args()
    global OUT
    OUT = claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
