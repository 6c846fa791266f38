#!/usr/bin/env python
"""bench.py - calibrated images/sec for the scoring + DAC + softmax-confidence + ECE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[3], the one the >=60 %-of-roofline target is quoted on): per GPU
1,000,000 synthetic L2-normalised image features x a 49,408-word vocabulary, 512-d, bf16 operands,
1,000 base classes, k=5, logit scale 100, 10 ECE bins.  Images are sharded across ranks (weak
scaling: every rank owns 1M images), text features replicated, the only collective is ONE
all-reduce of the 33-integer bin table per step.

One step = DAC fit (4 text matrices -> per-class multipliers) + fused scoring of the rank's image shard
(logits never reach HBM) with binning in the epilogue + table all-reduce + reading the table back.
`value` has the features resident in HBM; `e2e` runs the same step through the public API from pinned
HOST buffers, host->device copies inside the timed region.

Beyond the headline line the same JSON object carries (unless --no-extras):
  configs.in21k   BASELINE.json configs[4]: the per-GPU shard of 14M x 21,841 x 768 (1.75M images per rank,
                  10,000 base classes) - value / e2e / ms_per_step / dac_fit_ms / roofline, at every N;
  strong          N > 1: configs[3] at a FIXED 1M images in total (1M / N per rank, DAC fit sharded over the
                  ranks + all-gather of the multipliers) - value, ms_per_step, efficiency against the 1M-per-rank step;
  check.dist      N > 1: a fixed-seed 30k-image case scored sharded over the N ranks and alone on rank 0 -
                  bin tables, ACE, PIECE and macro-F1 must be identical; the run exits non-zero otherwise.
"""
from __future__ import annotations

import argparse
import datetime
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from dataclasses import dataclass

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "calibrated images/sec (GEMM+DAC+softmax+ECE)"
UNIT = "images/s"
LOGIT_SCALE, N_BINS = 100.0, 10
CLOCK_SAMPLE_MS = 200          # nvidia-smi sampling period during the timed region (B200_PROFILING.md's clocks line)


@dataclass(frozen=True)
class Workload:
    name: str
    n_images: int          # per GPU
    n_classes: int
    n_base: int
    dim: int
    k: int
    signal: float
    label: str

    def describe(self, n_images=None) -> str:
        n = self.n_images if n_images is None else n_images
        return (f"{self.label}: {n} images/GPU x {self.n_classes}-word vocabulary, {self.dim}-d bf16 features, "
                f"{self.n_base} base classes, DAC k={self.k}, {N_BINS}-bin ECE")


# BASELINE.json configs.  The default (and the headline the driver reads) is configs[3].
WORKLOADS = {w.name: w for w in [
    Workload("openvocab", 1_000_000, 49408, 1000, 512, 5, 0.50, "open-vocabulary"),
    Workload("imagenet", 50_000, 1000, 500, 512, 5, 0.25, "ImageNet-shaped base2new"),
    Workload("sun397", 19_850, 397, 199, 768, 5, 0.15, "SUN397-shaped ViT-L/14"),
    Workload("in21k", 1_750_000, 21841, 10000, 768, 5, 0.45, "ImageNet-21k-shaped (14M / 8 per GPU)"),
    Workload("eurosat", 8_100, 10, 5, 512, 5, 0.15, "EuroSAT-shaped base2new"),
]}


def config_dict(w: Workload, n_gpus: int, n_images=None):
    n = w.n_images if n_images is None else n_images
    big = n * w.dim * 2 > 126e6
    return {"workload": w.describe(n), "images_per_gpu": n, "classes": w.n_classes, "dim": w.dim,
            "base_classes": w.n_base, "k": w.k, "logit_scale": LOGIT_SCALE, "ece_bins": N_BINS, "signal": w.signal,
            "sharding": f"images sharded over {n_gpus} rank(s), text replicated, one bin-table all-reduce",
            "l2": (f"inputs ({n * w.dim * 2 / 1e9:.2f} GB of image features per step) are larger than the 126 MB L2; "
                   "no explicit flush") if big else
                  "inputs fit in L2: a 256 MB buffer is written between timed steps to flush it"}


# ----------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, the B200_PROFILING.md query line)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 200):
        self.gpu_index = gpu_index
        self.period_ms = int(period_ms)
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", str(self.period_ms)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------
# synthetic data (SURVEY.md 8(d) recipe), generated on the device, bf16-rounded
# ----------------------------------------------------------------------------------------
def make_device_data(w: Workload, n: int, seed: int):
    c, d, signal = w.n_classes, w.dim, w.signal
    g = torch.Generator(device="cuda").manual_seed(seed)
    unit = lambda x: x / x.norm(dim=-1, keepdim=True)
    # text features are identical on every rank (seeded apart from the images)
    gt = torch.Generator(device="cuda").manual_seed(12345)
    u = unit(torch.randn(d, device="cuda", generator=gt))
    txt_zs = unit(u[None, :] + 0.6 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=gt))
    txt_tuned = unit(txt_zs + 0.1 / d ** 0.5 * torch.randn(c, d, device="cuda", generator=gt))
    txt_zs = txt_zs.to(torch.bfloat16).float()
    txt_tuned = txt_tuned.to(torch.bfloat16).float()
    labels = torch.randint(0, c, (n,), device="cuda", generator=g)
    img = torch.empty((n, d), dtype=torch.bfloat16, device="cuda")
    for lo in range(0, n, 131072):
        hi = min(n, lo + 131072)
        raw = signal * txt_tuned[labels[lo:hi]] + torch.randn(hi - lo, d, device="cuda", generator=g) / d ** 0.5
        img[lo:hi] = unit(raw).to(torch.bfloat16)
    return img, labels, txt_zs, txt_tuned


# ----------------------------------------------------------------------------------------
# CPU leg: the reference's own functions (oracle/_ref, fetched by oracle/fetch_ref.py) where they import on their
# own, the oracle's restatement for the inline glue; the oracle port alone when oracle/_ref is absent.
# ----------------------------------------------------------------------------------------
def _host_sample(w: Workload, rows: int):
    """Host copy of `rows` images of the workload (same recipe, numpy generator)."""
    from clip_calibration_b200 import synth
    txt_zs, txt_tuned, rng = synth.make_text(w.n_classes, w.dim, 0)
    labels = rng.integers(0, w.n_classes, size=rows, dtype=np.int64)
    g = rng.standard_normal((rows, w.dim)).astype(np.float32)
    raw = np.float32(w.signal) * txt_tuned[labels] + g * np.float32(1.0 / np.sqrt(w.dim))
    img = synth.round_to_bf16(raw / np.linalg.norm(raw, axis=-1, keepdims=True))
    return img, labels, txt_zs, txt_tuned


def cpu_reference_step(w: Workload, sample, rows: int, fit_classes: int, threads: int):
    """One bounded sample of the workload through the reference chain.  Returns (extrapolated images/s for the
    full per-GPU job, stage seconds, kind).  kind "reference": DistanseAwareCalibration.fit/.predict and
    tools.metrics.ECE/MCE are the reference's own code (oracle/_ref); the contraction, softmax and argmax lines,
    which the reference performs inline in modules that need dassl / CLIP, are the oracle's restatement of them."""
    from oracle import cpu_oracle as orc
    from oracle import ref_loader
    img_np, labels_np, txt_zs_np, txt_tuned_np = sample
    ref = ref_loader.load()
    torch.set_num_threads(threads)
    c = txt_zs_np.shape[0]
    sel = np.linspace(0, c - 1, fit_classes).astype(int)
    t0 = time.perf_counter()
    if ref is not None:
        dac = ref.DistanseAwareCalibration()
        dac.fit(txt_zs_np[:w.n_base], txt_zs_np[sel], txt_tuned_np[:w.n_base], txt_tuned_np[sel], w.k)
        cc_sub = np.asarray(dac.class_confidence)
    else:
        cc_sub, *_ = orc.dac_fit(txt_zs_np[:w.n_base], txt_zs_np[sel], txt_tuned_np[:w.n_base], txt_tuned_np[sel], w.k)
    t_fit = time.perf_counter() - t0
    cc = np.ones(c)
    cc[sel] = cc_sub
    t0 = time.perf_counter()
    if ref is not None:
        dac = ref.DistanseAwareCalibration()
        dac.class_confidence = cc
        preds, confs = [], []
        with ref_loader.cpu_only():
            for lo in range(0, rows, 2048):
                lg = orc.logits_fp32(img_np[lo:lo + 2048], txt_tuned_np, LOGIT_SCALE, threads)     # zsclip.py:97-102
                lg = dac.predict(lg.astype(np.float64))                                            # reference code
                p, cf = orc.pred_and_conf(orc.softmax_lastaxis(lg))                                # vl_calibrator.py:91, vl_evaluator.py:68,:83
                preds.append(p); confs.append(cf)
        pred, conf = np.concatenate(preds), np.concatenate(confs)
    else:
        pred, conf, _ = orc.score_chain(img_np[:rows], txt_tuned_np, cc, LOGIT_SCALE, chunk=2048, threads=threads)
    t_chain = time.perf_counter() - t0
    t0 = time.perf_counter()
    if ref is not None:
        ref.metrics.ECE(conf, pred, labels_np[:rows], N_BINS)
        ref.metrics.MCE(conf, pred, labels_np[:rows], N_BINS)
    else:
        orc.ece(conf, pred, labels_np[:rows], N_BINS)
        orc.mce(conf, pred, labels_np[:rows], N_BINS)
    t_metrics = time.perf_counter() - t0
    full_job_s = t_fit * (c / fit_classes) + (t_chain + t_metrics) * (w.n_images / rows)
    return (w.n_images / full_job_s, {"fit_s": t_fit, "chain_s": t_chain, "metrics_s": t_metrics},
            "reference" if ref is not None else "port")


def sample_text(w: Workload, rows, fit_classes, kind):
    who = ("the reference's own DistanseAwareCalibration.fit/.predict and tools.metrics.ECE/MCE (oracle/_ref) around "
           "the oracle's restatement of the inline contraction / softmax / argmax" if kind == "reference"
           else "the oracle port of the reference chain (oracle/_ref absent)")
    return (f"{rows} of {w.n_images} image rows at the full {w.n_classes}-class vocabulary through "
            f"(100*img)@txt.T fp32 -> DAC.predict -> scipy-style softmax -> argmax/gather -> ECE+MCE, plus DAC.fit on "
            f"{fit_classes} of {w.n_classes} classes x {w.n_base} base; both extrapolated linearly to the full job; {who}")


def run_reference_arm(args, w: Workload, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows, fit_classes = min(1024, w.n_images), min(128, w.n_classes)
    sample = _host_sample(w, rows)
    vals, kind = [], "port"
    for i in range(args.warmup + args.steps):
        v, stages, kind = cpu_reference_step(w, sample, rows, fit_classes, threads)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * w.n_images / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(w, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                             "sample": sample_text(w, rows, fit_classes, kind), "stages_s": stages},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference CPU path on this box's host cores; ms_per_step extrapolated to the full per-GPU job"}
    out.emit(json.dumps(line))


# ----------------------------------------------------------------------------------------
# the CUDA arm
# ----------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end path are allocated on the socket the GPU's PCIe root hangs off.  Best effort: returns
    {"cpus": bound CPU count, "distinct": whether the set differs from the process's previous affinity} or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = [64 * wd + b for wd, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * wd + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"cpus": len(cpus), "narrowed": len(cpus) < len(allowed)}
    except Exception:  # noqa: BLE001
        pass
    return None


class _StdoutToStderr:
    """Route everything libraries print to fd 1 (e.g. NCCL's version banner) to stderr so that the
    ONE JSON line is the only thing on stdout; `emit()` writes that line to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self._real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str) -> None:
        sys.stdout.flush()
        os.write(self._real, (line + "\n").encode())


class StallGuard:
    """The driver must get its line even if a multi-rank step stops making progress (a collective that never
    completes would otherwise sit in NCCL's watchdog for minutes and end in SIGABRT with nothing printed).  Timed loops
    arm the guard and `beat()` once per step; if no beat arrives for `limit` seconds the guard reports where it
    stalled and which streams are busy, lets rank 0 print the line from what HAS been measured (marked "stalled"),
    and ends the process.  Every rank runs its own guard, so all of them leave."""

    def __init__(self):
        self.deadline = None
        self.where = ""
        self.limit = 0.0
        self.on_stall = None                 # set by main(): callable(where) -> None, must not return normally
        self.context = {}
        self._lock = threading.Lock()
        threading.Thread(target=self._run, daemon=True).start()

    def arm(self, where: str, limit: float = 75.0):
        with self._lock:
            self.where, self.limit, self.deadline = where, float(limit), time.time() + float(limit)

    def beat(self):
        with self._lock:
            if self.deadline is not None:
                self.deadline = time.time() + self.limit

    def disarm(self):
        with self._lock:
            self.deadline = None

    def _run(self):
        while True:
            time.sleep(1.0)
            with self._lock:
                late = self.deadline is not None and time.time() > self.deadline
                where = self.where
            if late and self.on_stall is not None:
                self.on_stall(where)


GUARD = None                                 # created by main() for the CUDA arm


class Ctx:
    """Process-wide handles of the CUDA arm."""

    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p.get("bf16_tflops_sustained", 1400.0)), float(p.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def static_traffic(w: Workload):
    """DRAM bytes per scoring call from the committed `ncu --set full` captures (a static figure: ncu cannot run
    inside the timed region), or None when the workload has no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            t = json.load(fh)
        e = t.get(w.name)
        return (e["dram_bytes_per_call"], e["source"]) if e else (None, None)
    except Exception:  # noqa: BLE001
        return None, None


def measure(ctx: Ctx, w: Workload, n_images: int, steps: int, warmup: int, do_e2e: bool, want_clocks: bool,
            shard_fit: bool = False, publish=None):
    """Device-resident (and optionally end-to-end) timing of one workload with `n_images` per rank."""
    from clip_calibration_b200 import native, pipeline
    from clip_calibration_b200 import table_math as tm
    dist, world, rank = ctx.dist, ctx.world, ctx.rank
    img, labels, txt_zs, txt_tuned = make_device_data(w, n_images, seed=1000 + rank)
    base_zs, base_tuned = txt_zs[:w.n_base].contiguous(), txt_tuned[:w.n_base].contiguous()
    txt_op = txt_tuned.to(torch.bfloat16).contiguous()
    thr = tm.uniform_thresholds(N_BINS)
    table = native.new_table(N_BINS)
    host_table = torch.empty_like(table, device="cpu").pin_memory()
    call_events = []
    need_flush = n_images * w.dim * 2 <= 126e6          # inputs that fit in L2 would otherwise be re-read from it
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if need_flush else None
    use_graph = need_flush                                # small shapes are launch-bound: replay the step as ONE CUDA graph

    def fit():
        if shard_fit and world > 1:
            return pipeline.dac_fit_sharded(base_zs, txt_zs, base_tuned, txt_tuned, w.k)
        return native.dac_fit(base_zs, txt_zs, base_tuned, txt_tuned, w.k)[0]

    side = pipeline._side_stream(img.device)
    held = {}

    def compute(record=False):
        """DAC fit + fused scoring/binning (per-image pred / conf are written too: 8 B/image)"""
        if use_graph:
            # launch-bound shapes: the fit (which pass 1 does not need) runs on a side stream underneath pass 1,
            # pass 2 joins them - the two-launch form of the same scoring call, bit-identical to the fused one
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                held["cc"] = fit()
            table.zero_()
            dotmax, pred = native.score_pass1(img, txt_op)
            main.wait_stream(side)
            held["out"] = (pred, native.score_pass2(img, txt_op, dotmax, pred, held["cc"], LOGIT_SCALE, labels, thr, table))
            return
        cc = fit()
        table.zero_()
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        native.score_fused(img, txt_op, cc, LOGIT_SCALE, labels, thr, table, want_pred=True, want_conf=True)
        if record:
            e1.record()
            call_events.append((e0, e1))

    fit_identical = None
    if shard_fit and world > 1:       # the sharded fit must reproduce the plain one bit for bit
        fit_identical = bool(torch.equal(fit(), native.dac_fit(base_zs, txt_zs, base_tuned, txt_tuned, w.k)[0]))

    graph = None
    if use_graph:
        # warm every lazy initialisation (function attributes, memory pools) before capturing
        warm = torch.cuda.Stream()
        warm.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(warm):
            for _ in range(3):
                compute()
        torch.cuda.current_stream().wait_stream(warm)
        torch.cuda.synchronize()
        launches_before = native.launch_count()
        try:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="relaxed"):
                compute()
            graph_launches = native.launch_count() - launches_before
        except Exception as exc:  # noqa: BLE001  (capture is an optimisation: fall back to plain launches)
            print(f"bench: CUDA graph capture failed, timing plain launches instead: {exc!r}", file=sys.stderr)
            graph = None
            torch.cuda.synchronize()

    def step(record=False):
        """device-resident step: DAC fit + fused scoring/binning + table all-reduce + table D2H"""
        if graph is not None:
            graph.replay()
        else:
            compute(record)
        if world > 1:
            dist.all_reduce(table)
        host_table.copy_(table, non_blocking=True)

    def timed(fn, n_steps, finish=None):
        if GUARD is not None:
            GUARD.arm(f"{w.name} ({n_images} images/rank): timed loop of {getattr(fn, '__name__', 'step')}")
        ctx.barrier()
        if not need_flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n_steps):
                fn()
                if GUARD is not None:
                    GUARD.beat()
            if finish is not None:
                finish()
            e1.record()
            ctx.barrier()
            total = e0.elapsed_time(e1)
        else:                                          # flush L2 between steps, time each step on its own
            pairs = []
            for _ in range(n_steps):
                flush_buf.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                if finish is not None:
                    finish()
                e1.record()
                pairs.append((e0, e1))
                if GUARD is not None:
                    GUARD.beat()
            ctx.barrier()
            total = sum(a.elapsed_time(b) for a, b in pairs)
        total = ctx.max_over_ranks(total)
        if GUARD is not None:
            GUARD.disarm()
        return total

    # ---------------- device-resident number
    for _ in range(warmup):
        step()
    sampler = ClockSampler(ctx.local_rank, CLOCK_SAMPLE_MS) if (want_clocks and rank == 0 and CLOCK_SAMPLE_MS > 0) else None
    if sampler:
        sampler.start()
    native.score_trace(reset=True)
    native.score_guess_stats(reset=True)
    launches0 = native.launch_count()
    ms_total = timed(lambda: step(record=True), steps)
    launches = (native.launch_count() - launches0) if graph is None else graph_launches * steps
    clocks = sampler.stop() if sampler else None
    trace = native.score_trace(reset=True)
    guess_rows, guess_redone = native.score_guess_stats(reset=True)
    call_ms = [a.elapsed_time(b) for a, b in call_events]
    tab = host_table.numpy().view(np.uint64)
    summary = {"n": tm.total_count(tab), "ece": float(tm.ece_from_table(tab)), "accuracy": tm.accuracy(tab)}
    assert summary["n"] == n_images * world, summary

    # DAC fit alone (reported separately).  The workload's text features are bf16 values (the SURVEY 8(d) recipe rounds
    # every operand to bf16), which the kNN filter detects and serves with ONE MMA per K step; features with full
    # fp32 mantissas take the hi/lo split (three MMAs) - timed too, on a perturbed copy, so the line shows both.
    fit_ms = timed(lambda: fit(), 3) / 3
    gq = torch.Generator(device="cuda").manual_seed(777)
    jitter = lambda t: torch.nn.functional.normalize(t + 1e-4 * torch.randn(t.shape, device="cuda", generator=gq), dim=-1)
    zs_g, tu_g = jitter(txt_zs), jitter(txt_tuned)
    bz_g, bt_g = zs_g[:w.n_base].contiguous(), tu_g[:w.n_base].contiguous()
    native.dac_fit(bz_g, zs_g, bt_g, tu_g, w.k)
    fit_fp32_ms = timed(lambda: native.dac_fit(bz_g, zs_g, bt_g, tu_g, w.k), 3) / 3
    del zs_g, tu_g, bz_g, bt_g
    res = {"ms_total": ms_total, "steps": steps, "launches": int(launches), "clocks": clocks, "trace": trace,
           "call_ms": call_ms, "fit_ms": fit_ms, "fit_fp32_ms": fit_fp32_ms, "summary": summary, "graph": graph is not None,
           "redo_rate": (guess_redone / guess_rows) if guess_rows else None, "n_images": n_images,
           "fit_identical": fit_identical}

    # ---------------- end-to-end number: host buffers through the public API
    if do_e2e:
        host_img = img.cpu().pin_memory()
        host_labels = labels.cpu().pin_memory()
        # text features sit on the host in the feature dtype of the workload (bf16; the synthetic values are
        # bf16-representable, so this is lossless) - DAC fit widens them to fp32 on the device
        host_txt = {k: v.to(torch.bfloat16).cpu().pin_memory() for k, v in
                    {"bz": base_zs, "cz": txt_zs, "bt": base_tuned, "ct": txt_tuned}.items()}
        del img
        torch.cuda.empty_cache()
        e2e_table = {}
        chunk_rows = 262144 if n_images > 2 * 262144 else max(1024, -(-n_images // 4 // 128) * 128)

        pending = []

        def e2e_queue():
            # H2D of the four text matrices + DAC fit (class_confidence stays on the device); with several ranks the
            # text side is uploaded and fitted by rank 0 and broadcast over NVLink (the text features are replicated)
            scorer = pipeline.CalibratedScorer.from_dac(host_txt["bz"], host_txt["cz"], host_txt["bt"], host_txt["ct"],
                                                        k=w.k, logit_scale=LOGIT_SCALE, n_bins=N_BINS,
                                                        operand_dtype=torch.bfloat16, share_text=world > 1,
                                                        overlap_fit=True)
            if GUARD is not None:
                GUARD.context["scorers"] = (GUARD.context.get("scorers", []) + [scorer])[-3:]
            scorer.accumulate_host(host_img, host_labels, chunk_rows=chunk_rows)                 # chunked H2D + scoring
            pending.append(scorer.reduced_table_async())                                         # all-reduce + D2H, queued

        def e2e_drain(keep=0):
            while len(pending) > keep:
                e2e_table["t"] = pending.pop(0).result()                                         # host waits for the D2H
                assert tm.total_count(e2e_table["t"]) == n_images * world

        def e2e_step_blocking():
            e2e_queue()
            e2e_drain()

        def e2e_step_pipelined():
            # an evaluation LOOP (several test sets / shards in a row): step i's table is read on the host after step
            # i+1 has been queued, so that step's uploads and first launches run underneath step i's last kernels.
            # Every step's H2D copies, kernels, all-reduce and D2H read still lie inside the timed region (the last
            # step is drained before the closing event).  Small (L2-flushed) shapes are timed one step at a time.
            e2e_queue()
            e2e_drain(keep=0 if need_flush else 1)

        for _ in range(2):
            e2e_step_blocking()
        e2e_steps = max(3, steps // 2)
        blocking_ms = timed(e2e_step_blocking, e2e_steps)
        h2d = host_img.numel() * 2 + host_labels.numel() * 8 + sum(v.numel() * v.element_size() for v in host_txt.values())
        # what is known so far, should the pipelined loop below stall (StallGuard): the blocking form IS an end-to-end number
        res["e2e"] = {"value": n_images * world * e2e_steps / (blocking_ms * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(3 * (N_BINS + 1) * 8),
                      "ms_per_step": blocking_ms / e2e_steps, "steps": e2e_steps, "chunk_rows": chunk_rows,
                      "mode": "each step's table is read back before the next step is queued (the pipelined loop did not finish)"}
        if publish is not None:
            publish(res)
        if GUARD is not None:
            GUARD.arm(f"{w.name}: warm-up of the pipelined end-to-end loop")
        for _ in range(3):            # two steps in flight need more device / pinned blocks: let the allocators reach
            e2e_step_pipelined()      # their steady state (cudaMalloc / cudaHostAlloc synchronise the device)
            if GUARD is not None:
                GUARD.beat()
        e2e_drain()
        e2e_ms = timed(e2e_step_pipelined, e2e_steps, finish=e2e_drain)
        res["e2e"] = {"value": n_images * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(3 * (N_BINS + 1) * 8),
                      "ms_per_step": e2e_ms / e2e_steps, "steps": e2e_steps, "chunk_rows": chunk_rows,
                      "mode": ("one step at a time (L2 flushed between steps)" if need_flush else
                               "evaluation loop, software-pipelined by one step: the host reads step i's table after "
                               "queueing step i+1; all copies of all steps are inside the timed region"),
                      "blocking": {"value": n_images * world * e2e_steps / (blocking_ms * 1e-3),
                                   "ms_per_step": blocking_ms / e2e_steps,
                                   "mode": "each step's table is read back before the next step is queued"}}
        del host_img, host_labels
    return res


def roofline_of(w: Workload, res: dict, peaks):
    """Roofline of the scoring call (the kernels of ONE ccal_score_fused call, timed with CUDA events on the launching
    stream inside the timed region) + the in-kernel timings of its tensor kernels."""
    peak_sust, peak_burst, which = peaks
    n = res["n_images"]
    flops = 2.0 * n * w.n_classes * w.dim
    call_ms = statistics.mean(res["call_ms"]) if res["call_ms"] else None
    trace = res["trace"]
    pipeline_used = "verify_bf16" in trace
    kernels = {}
    exec_flops = 0.0
    for name, t in trace.items():
        per_launch = t["ms_per_launch"]
        launches_per_step = t["launches"] / res["steps"]
        # executed tensor flops per launch: one pass over N x C x D per guess / verify launch, two for the two-pass kernel
        passes = {"guess_fp8": 1.0, "verify_bf16": 1.0, "two_pass": 2.0}.get(name)
        entry = {"launches_per_step": launches_per_step, "ms_per_launch": per_launch, "sm_mhz_in_kernel": t["sm_mhz"]}
        if passes and launches_per_step > 0:
            fl = passes * flops / launches_per_step          # chunked calls split the rows over several launches
            entry["executed_tflops"] = fl / (per_launch * 1e-3) / 1e12
            if t["sm_mhz"]:
                rate = 16384 if name == "guess_fp8" else 8192   # dense flop / cycle / SM: e4m3 vs bf16
                hw = 148 * rate * t["sm_mhz"] * 1e6 / 1e12
                entry["hw_peak_at_kernel_clock_tflops"] = hw
                entry["frac_of_hw_peak_at_kernel_clock"] = entry["executed_tflops"] / hw
            exec_flops += passes * flops
        kernels[name] = entry
    dominant = "verify_bf16" if pipeline_used else "two_pass"
    traffic, traffic_src = static_traffic(w)
    if call_ms is None:                                       # CUDA-graph replay: no events inside; use the in-kernel spans
        call_ms = sum(t["ms_per_launch"] * t["launches"] / res["steps"] for t in trace.values())
    algo_tf = flops / (call_ms * 1e-3) / 1e12
    roof = {"kernel": ("ccal_score_fused call = FP8 guess pass (tcgen05 kind::f8f6f4) + exact guessed-class logit + bf16 "
                       "verify pass (tcgen05 kind::f16, softmax/argmax/bin epilogue) + redo of mis-guessed rows"
                       if pipeline_used else "score_fused_kernel (tcgen05 two-pass GEMM + softmax/bin epilogue)"),
            "dominant_kernel": dominant, "bound": "tensor", "achieved": algo_tf, "peak": peak_sust, "unit": "TFLOP/s",
            "frac": algo_tf / peak_sust, "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": which + ", sustained bf16 (timed inside a back-to-back loop)",
            "achieved_is": "ALGORITHMIC flops 2*N*C*D per scoring call (one contraction, what the reference computes) / "
                           "CUDA-event time of the call's kernels",
            "call_ms": call_ms, "call_ms_min": min(res["call_ms"]) if res["call_ms"] else None,
            "call_share_of_step": call_ms * res["steps"] / res["ms_total"],
            "executed": {"tensor_flops_per_call": exec_flops, "achieved": exec_flops / (call_ms * 1e-3) / 1e12,
                         "note": "guess pass = N*C*D*2 e4m3 flops (kind::f8f6f4, 2x the bf16 rate), verify pass = N*C*D*2 "
                                 "bf16 flops; redo < 2 % extra" if pipeline_used else
                                 "two passes of N*C*D*2 bf16 flops (pass 1 max/argmax, pass 2 sum-exp)"},
            "kernels": kernels, "redo_rate": res["redo_rate"], "frac_of_burst": algo_tf / peak_burst}
    return roof


def dist_check(ctx: Ctx):
    """N ranks == 1 rank, on the driver's box: a fixed-seed 30,011-image case scored sharded (NCCL all-reduce of the
    bin table, all-reduced radix histograms for the quantile edges, all-reduced class counts) and alone on rank 0."""
    from clip_calibration_b200 import native, pipeline, synth
    from clip_calibration_b200.tools import metrics
    dist, rank, world = ctx.dist, ctx.rank, ctx.world
    case = synth.make_case("dist", 30011, 1000, 500, 512, 5, 0.25, seed=0)       # same on every rank
    lo, hi = pipeline.shard_bounds(len(case.labels), rank, world)
    scorer = pipeline.CalibratedScorer.from_dac(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k=5,
                                                logit_scale=100.0, n_bins=10, share_text=True, overlap_fit=True,
                                                keep_outputs=True)
    scorer.score(case.img[lo:hi], case.labels[lo:hi])
    prox_all = np.random.default_rng(1).random(len(case.labels)).astype(np.float32)
    ev = scorer.evaluate(proximity=torch.from_numpy(prox_all[lo:hi]).cuda())
    out = None
    if rank == 0:
        solo = pipeline.CalibratedScorer(case.txt_tuned, scorer.class_conf, 100.0, 10, keep_outputs=True, group=False)
        solo.score(case.img, case.labels)
        ev1 = solo.evaluate(proximity=torch.from_numpy(prox_all).cuda())
        same = lambda k: bool(ev[k] == ev1[k])
        out = {"world": world, "rows": len(case.labels),
               "tables_identical": bool(np.array_equal(ev["bin_table"], ev1["bin_table"])),
               "ace_identical": same("ace"), "piece_identical": same("piece"), "macro_f1_identical": same("macro_f1"),
               "ece_identical": same("ece"), "mce_identical": same("mce"), "accuracy_identical": same("accuracy"),
               "ece": ev["ece"], "ace": ev["ace"], "piece": ev["piece"], "macro_f1": ev["macro_f1"]}
        out["ok"] = all(v for k, v in out.items() if k.endswith("_identical"))
    flag = torch.tensor([1 if (out is None or out["ok"]) else 0], device="cuda")
    dist.broadcast(flag, 0)
    return out, int(flag.item()) == 1


def main():
    global CLOCK_SAMPLE_MS, GUARD
    out = _StdoutToStderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline workload only (no in21k / strong / dist check)")
    ap.add_argument("--clock-sample-ms", type=int, default=CLOCK_SAMPLE_MS,
                    help="nvidia-smi sampling period inside the timed region (0 = no sampling; development aid)")
    ap.add_argument("--workload", default="openvocab", choices=sorted(WORKLOADS),
                    help="BASELINE.json config shape (default: the headline open-vocabulary workload)")
    args = ap.parse_args()
    CLOCK_SAMPLE_MS = args.clock_sample_ms
    w = WORKLOADS[args.workload]
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args, w, out)
        return

    from clip_calibration_b200 import build as _build
    _build.build()                         # no-op when libccal.so matches the sources (it normally travels pre-built)
    from clip_calibration_b200 import _lib, pipeline

    ctx = Ctx()
    dist, world, rank = ctx.dist, ctx.world, ctx.rank
    torch.cuda.set_device(ctx.local_rank)
    numa = bind_to_gpu_numa_node(ctx.local_rank) if world > 1 else None
    lib = _lib.load()
    _lib.check(lib.ccal_check_device(), "ccal_check_device")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # a collective that cannot complete is reported by the StallGuard within ~75 s; NCCL's own watchdog is the backstop
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank),
                                timeout=datetime.timedelta(seconds=240))
    GUARD = StallGuard()

    peaks = load_peaks()
    state = {"head": None, "configs": {}, "strong": None, "check_dist": None, "cpu": None}

    def build_line(stalled=None):
        head = state["head"]
        roofline = roofline_of(w, head, peaks)
        ms_total, steps = head["ms_total"], head["steps"]
        check = dict(head["summary"])
        if state["check_dist"] is not None:
            check["dist"] = state["check_dist"]
        line = {"metric": METRIC, "value": w.n_images * world * steps / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": steps, "warmup": args.warmup, "ms_per_step": ms_total / steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": config_dict(w, world), "clocks": head["clocks"], "e2e": head["e2e"],
                "gpu_launches": head["launches"], "roofline": roofline, "cpu_baseline": state["cpu"],
                "dac_fit_ms": head["fit_ms"], "dac_fit_ms_fp32_features": head["fit_fp32_ms"],
                "cuda_graph_step": head["graph"], "check": check, "numa": numa}
        if state["configs"]:
            line["configs"] = state["configs"]
        if state["strong"] is not None:
            line["strong"] = state["strong"]
        if stalled is not None:
            line["stalled"] = stalled
        return line

    def on_stall(where):
        # runs on the guard's thread while the main thread sits in a wait that will not return
        dev = torch.device("cuda", ctx.local_rank)
        busy = {}
        try:
            busy = {"compute": not torch.cuda.current_stream(dev).query(), "side": not pipeline._side_stream(dev).query(),
                    "copy": not pipeline._copy_stream(dev).query()}
            for sc in GUARD.context.get("scorers", []):
                d = getattr(sc, "_dbg", None)
                if d:
                    busy.setdefault("from_dac", []).append({k: ([e.query() for e in v] if isinstance(v, list) else v.query())
                                                            for k, v in d.items()})
            if os.environ.get("CCAL_TRACE_MARKS"):
                import ctypes
                buf = ctypes.create_string_buffer(8192)
                lib.ccal_trace_marks_report(buf, 8192)
                busy["fit_marks"] = buf.value.decode()
        except Exception as exc:  # noqa: BLE001
            busy["error"] = repr(exc)
        print(f"bench[rank {rank}]: no progress for {GUARD.limit:.0f} s in {where}; busy streams: {busy}", file=sys.stderr, flush=True)
        if rank == 0 and state["head"] is not None:
            try:
                out.emit(json.dumps(build_line({"where": where, "busy_streams": busy, "note":
                                                "measurements taken before the stall are reported; the rest is missing"})))
            except Exception as exc:  # noqa: BLE001
                print(f"bench: could not assemble the line after the stall: {exc!r}", file=sys.stderr, flush=True)
        os._exit(0 if state["head"] is not None else 4)

    GUARD.on_stall = on_stall
    state["head"] = None
    head = measure(ctx, w, w.n_images, args.steps, args.warmup, do_e2e=True, want_clocks=True,
                   publish=lambda r: state.__setitem__("head", r))
    state["head"] = head
    extras = not args.no_extras and args.workload == "openvocab"
    dist_ok = True
    if extras:
        torch.cuda.empty_cache()
        w5 = WORKLOADS["in21k"]
        r5 = measure(ctx, w5, w5.n_images, max(3, args.steps // 2), 3, do_e2e=True, want_clocks=False)
        roof5 = roofline_of(w5, r5, peaks)
        state["configs"]["in21k"] = {
            "workload": w5.describe(), "total_images": w5.n_images * world,
            "value": w5.n_images * world * r5["steps"] / (r5["ms_total"] * 1e-3), "unit": UNIT,
            "ms_per_step": r5["ms_total"] / r5["steps"], "steps": r5["steps"], "dac_fit_ms": r5["fit_ms"],
            "dac_fit_ms_fp32_features": r5["fit_fp32_ms"],
            "e2e": r5["e2e"], "roofline": roof5, "check": r5["summary"], "gpu_launches": r5["launches"]}
        torch.cuda.empty_cache()
        if world > 1:
            n_strong = w.n_images // world
            rs = measure(ctx, w, n_strong, args.steps, 3, do_e2e=False, want_clocks=False, shard_fit=True)
            ms_strong, ms_weak = rs["ms_total"] / rs["steps"], head["ms_total"] / head["steps"]
            state["strong"] = {
                "workload": f"{w.label}: {n_strong * world} images in TOTAL ({n_strong} per rank) x {w.n_classes} classes, "
                            "DAC fit sharded over the ranks (classes / N each) + all-gather of the multipliers",
                "value": n_strong * world * rs["steps"] / (rs["ms_total"] * 1e-3), "unit": UNIT,
                "ms_per_step": ms_strong, "steps": rs["steps"], "dac_fit_ms": rs["fit_ms"],
                "efficiency_vs_1m_per_rank_step": ms_weak / (world * ms_strong),
                "roofline_frac": roofline_of(w, rs, peaks)["frac"], "check": rs["summary"],
                "sharded_fit_identical_to_plain_fit": rs["fit_identical"]}
            GUARD.arm("N ranks == 1 rank identity check", 120.0)
            state["check_dist"], dist_ok = dist_check(ctx)
            GUARD.disarm()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        sys.exit(0 if dist_ok else 1)

    # ---------------- CPU baseline (bounded sample, this box's host cores)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rows, fit_classes = min(4096, w.n_images), min(512, w.n_classes)
        cpu_val, stages, kind = cpu_reference_step(w, _host_sample(w, rows), rows, fit_classes, threads)
        state["cpu"] = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": kind,
                        "sample": sample_text(w, rows, fit_classes, kind), "stages_s": stages}

    out.emit(json.dumps(build_line()))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if dist_ok else 1)


if __name__ == "__main__":
    main()
