#!/usr/bin/env python
"""bench.py - calibrated images/sec for the scoring + DAC + softmax-confidence + ECE hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[3], the one the >=60 %-of-roofline target is quoted on): per GPU
1,000,000 synthetic L2-normalised image features x a 49,408-word vocabulary, 512-d, bf16 operands,
1,000 base classes, k=5, logit scale 100, 10 ECE bins.  Images are sharded across ranks (weak
scaling: every rank owns 1M images), text features replicated, the only collective is ONE
all-reduce of the 33-integer bin table per step.

One step = DAC fit (4 text matrices -> 49,408 per-class multipliers) + fused two-pass scoring of
the rank's image shard (logits never reach HBM) with binning in the epilogue + table all-reduce
+ reading the table back.  `value` has the features resident in HBM; `e2e` runs the same step
through the public API from pinned HOST buffers, host->device copies inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "calibrated images/sec (GEMM+DAC+softmax+ECE)"
UNIT = "images/s"
LOGIT_SCALE, N_BINS = 100.0, 10
# name -> (images per GPU, classes, base classes, feature width, k, signal, label); BASELINE.json configs.
# The default (and the only one the driver runs) is configs[3], the one the roofline target is quoted on.
WORKLOADS = {
    "openvocab": (1_000_000, 49408, 1000, 512, 5, 0.50, "open-vocabulary"),
    "imagenet": (50_000, 1000, 500, 512, 5, 0.25, "ImageNet-shaped base2new"),
    "sun397": (19_850, 397, 199, 768, 5, 0.15, "SUN397-shaped ViT-L/14"),
    "in21k": (1_750_000, 21841, 10000, 768, 5, 0.45, "ImageNet-21k-shaped (14M / 8 per GPU)"),
    "eurosat": (8_100, 10, 5, 512, 5, 0.15, "EuroSAT-shaped base2new"),
}


def set_workload(name: str) -> None:
    global N_IMAGES, N_CLASSES, N_BASE, DIM, K_DAC, SIGNAL, WORKLOAD
    N_IMAGES, N_CLASSES, N_BASE, DIM, K_DAC, SIGNAL, label = WORKLOADS[name]
    WORKLOAD = (f"{label}: {N_IMAGES} images/GPU x {N_CLASSES}-word vocabulary, {DIM}-d bf16 features, "
                f"{N_BASE} base classes, DAC k={K_DAC}, {N_BINS}-bin ECE")


set_workload("openvocab")


def config_dict(n_gpus):
    return {"workload": WORKLOAD, "images_per_gpu": N_IMAGES, "classes": N_CLASSES, "dim": DIM, "base_classes": N_BASE,
            "k": K_DAC, "logit_scale": LOGIT_SCALE, "ece_bins": N_BINS, "signal": SIGNAL,
            "sharding": f"images sharded over {n_gpus} rank(s), text replicated, one bin-table all-reduce",
            "l2": (f"inputs ({N_IMAGES * DIM * 2 / 1e9:.2f} GB of image features per step) are larger than the 126 MB L2; "
                   "no explicit flush") if N_IMAGES * DIM * 2 > 126e6 else
                  "inputs fit in L2: a 256 MB buffer is written between timed steps to flush it"}


# ----------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, the B200_PROFILING.md query line)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "100"],
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
def make_device_data(seed: int):
    n, c, d, signal = N_IMAGES, N_CLASSES, DIM, SIGNAL
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
# CPU leg: the reference's own path (oracle restatement; /root/reference is absent on the box)
# ----------------------------------------------------------------------------------------
def cpu_reference_step(img_np, labels_np, txt_zs_np, txt_tuned_np, rows: int, fit_classes: int, threads: int):
    """One bounded sample of the workload through the reference chain.  Returns
    (extrapolated images/s for the full job, dict of stage seconds)."""
    from oracle import cpu_oracle as orc
    torch.set_num_threads(threads)
    c = txt_zs_np.shape[0]
    sel = np.linspace(0, c - 1, fit_classes).astype(int)
    t0 = time.perf_counter()
    cc_sub, *_ = orc.dac_fit(txt_zs_np[:N_BASE], txt_zs_np[sel], txt_tuned_np[:N_BASE], txt_tuned_np[sel], K_DAC)
    t_fit = time.perf_counter() - t0
    cc = np.ones(c)
    cc[sel] = cc_sub
    t0 = time.perf_counter()
    pred, conf, _ = orc.score_chain(img_np[:rows], txt_tuned_np, cc, LOGIT_SCALE, chunk=2048, threads=threads)
    t_chain = time.perf_counter() - t0
    t0 = time.perf_counter()
    orc.ece(conf, pred, labels_np[:rows], N_BINS)
    orc.mce(conf, pred, labels_np[:rows], N_BINS)
    t_metrics = time.perf_counter() - t0
    full_job_s = t_fit * (c / fit_classes) + (t_chain + t_metrics) * (N_IMAGES / rows)
    return N_IMAGES / full_job_s, {"fit_s": t_fit, "chain_s": t_chain, "metrics_s": t_metrics}


def sample_text(rows, fit_classes):
    return (f"{rows} of {N_IMAGES} image rows at the full {N_CLASSES}-class vocabulary through "
            f"(100*img)@txt.T fp32 -> DAC.predict -> scipy-style softmax -> argmax/gather -> ECE+MCE, plus DAC.fit on "
            f"{fit_classes} of {N_CLASSES} classes x {N_BASE} base; both extrapolated linearly to the full job")


def run_reference_arm(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows, fit_classes = min(1024, N_IMAGES), min(128, N_CLASSES)
    rng_case = _host_sample(rows)
    vals = []
    for i in range(args.warmup + args.steps):
        v, stages = cpu_reference_step(*rng_case, rows, fit_classes, threads)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * N_IMAGES / value, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": sample_text(rows, fit_classes)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference torch/numpy CPU path (oracle port; the Python reference tree is absent on the GPU box), "
                    "ms_per_step extrapolated to the full 1M-image job"}
    out.emit(json.dumps(line))


def _host_sample(rows):
    """Host copy of the first `rows` images of rank 0's data (same recipe, numpy generator)."""
    from clip_calibration_b200 import synth
    txt_zs, txt_tuned, rng = synth.make_text(N_CLASSES, DIM, 0)
    labels = rng.integers(0, N_CLASSES, size=rows, dtype=np.int64)
    g = rng.standard_normal((rows, DIM)).astype(np.float32)
    raw = np.float32(SIGNAL) * txt_tuned[labels] + g * np.float32(1.0 / np.sqrt(DIM))
    img = synth.round_to_bf16(raw / np.linalg.norm(raw, axis=-1, keepdims=True))
    return img, labels, txt_zs, txt_tuned


# ----------------------------------------------------------------------------------------
# the CUDA arm
# ----------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(gpu_index: int):
    """Pin this rank to the CPUs NVML reports as local to its GPU, so that the pinned host buffers of the
    end-to-end path are allocated on the socket the GPU's PCIe root hangs off (8 ranks x 1.2 GB per step
    otherwise cross the inter-socket link).  Best effort: returns the CPU count bound, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1 and 64 * w + b < n_cpu]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
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


def main():
    out = _StdoutToStderr()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="openvocab", choices=sorted(WORKLOADS),
                    help="BASELINE.json config shape (default: the headline open-vocabulary workload)")
    args = ap.parse_args()
    set_workload(args.workload)
    if args.warmup < 3 and args.impl == "cuda":
        args.warmup = 3

    if args.impl == "reference":
        run_reference_arm(args, out)
        return

    import torch.distributed as dist
    from clip_calibration_b200 import build as _build
    _build.build()                         # no-op when libccal.so matches the sources (it normally travels pre-built)
    from clip_calibration_b200 import _lib, native, pipeline
    from clip_calibration_b200 import table_math as tm

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    lib = _lib.load()
    _lib.check(lib.ccal_check_device(), "ccal_check_device")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    img, labels, txt_zs, txt_tuned = make_device_data(seed=1000 + rank)
    base_zs, base_tuned = txt_zs[:N_BASE].contiguous(), txt_tuned[:N_BASE].contiguous()
    txt_op = txt_tuned.to(torch.bfloat16).contiguous()
    thr = tm.uniform_thresholds(N_BINS)
    table = native.new_table(N_BINS)
    host_table = torch.empty_like(table, device="cpu").pin_memory()
    kern_events = []

    def step(record=False):
        """device-resident step: DAC fit + fused scoring/binning + table all-reduce + table D2H"""
        cc, *_ = native.dac_fit(base_zs, txt_zs, base_tuned, txt_tuned, K_DAC)
        table.zero_()
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        # per-image (pred, conf) are written too (8 B/image), although only the bin table is needed for the metric
        native.score_fused(img, txt_op, cc, LOGIT_SCALE, labels, thr, table, want_pred=True, want_conf=True)
        if record:
            e1.record()
            kern_events.append((e0, e1))
        if world > 1:
            dist.all_reduce(table)
        host_table.copy_(table, non_blocking=True)

    need_flush = N_IMAGES * DIM * 2 <= 126e6          # inputs that fit in L2 would otherwise be re-read from it
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if need_flush else None

    def timed(fn, steps):
        barrier()
        if not need_flush:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            total = e0.elapsed_time(e1)
        else:                                          # flush L2 between steps, time each step on its own
            pairs = []
            for _ in range(steps):
                flush_buf.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                pairs.append((e0, e1))
            barrier()
            total = sum(a.elapsed_time(b) for a, b in pairs)
        ms = torch.tensor([total], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---------------- device-resident number
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = native.launch_count()
    ms_total = timed(lambda: step(record=True), args.steps)
    launches = native.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    kern_ms = [a.elapsed_time(b) for a, b in kern_events]
    summary = {"n": tm.total_count(host_table.numpy().view(np.uint64)),
               "ece": float(tm.ece_from_table(host_table.numpy().view(np.uint64))),
               "accuracy": tm.accuracy(host_table.numpy().view(np.uint64))}
    assert summary["n"] == N_IMAGES * world, summary

    # DAC fit alone (reported separately)
    def fit_only():
        native.dac_fit(base_zs, txt_zs, base_tuned, txt_tuned, K_DAC)
    fit_ms = timed(fit_only, 3) / 3

    # ---------------- end-to-end number: host buffers through the public API
    host_img = img.cpu().pin_memory()
    host_labels = labels.cpu().pin_memory()
    # text features sit on the host in the feature dtype of the workload (bf16; the synthetic values are
    # bf16-representable, so this is lossless) - DAC fit widens them to fp32 on the device
    host_txt = {k: v.to(torch.bfloat16).cpu().pin_memory() for k, v in
                {"bz": base_zs, "cz": txt_zs, "bt": base_tuned, "ct": txt_tuned}.items()}
    del img
    torch.cuda.empty_cache()
    e2e_table = {}

    def e2e_step():
        # H2D of the four text matrices + DAC fit (class_confidence stays on the device); with several ranks the text
        # side is uploaded and fitted by rank 0 and broadcast over NVLink (the text features are replicated)
        scorer = pipeline.CalibratedScorer.from_dac(host_txt["bz"], host_txt["cz"], host_txt["bt"], host_txt["ct"],
                                                    k=K_DAC, logit_scale=LOGIT_SCALE, n_bins=N_BINS,
                                                    operand_dtype=torch.bfloat16, share_text=world > 1,
                                                    overlap_fit=True)
        scorer.accumulate_host(host_img, host_labels, chunk_rows=131072)                     # chunked H2D + scoring
        e2e_table["t"] = scorer.reduced_table()                                              # all-reduce + D2H

    for _ in range(2):
        e2e_step()
    e2e_ms = timed(e2e_step, max(3, args.steps // 2))
    e2e_steps = max(3, args.steps // 2)
    assert tm.total_count(e2e_table["t"]) == N_IMAGES * world
    h2d = host_img.numel() * 2 + host_labels.numel() * 8 + sum(v.numel() * v.element_size() for v in host_txt.values())
    d2h = 3 * (N_BINS + 1) * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except Exception:  # noqa: BLE001
        pass
    peak_sust = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_burst = float(peaks.get("bf16_tflops", 1590.0))
    which = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    k_ms = statistics.mean(kern_ms)
    algo_tf = 2.0 * N_IMAGES * N_CLASSES * DIM / (k_ms * 1e-3) / 1e12
    exec_tf = 2.0 * algo_tf
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as fh:
            traffic = json.load(fh).get("score_fused_dram_bytes_per_launch")
    except Exception:  # noqa: BLE001
        pass
    roofline = {"kernel": "score_fused_kernel<resident, DAC> (tcgen05 two-pass GEMM + softmax/bin epilogue)",
                "bound": "tensor", "achieved": algo_tf, "peak": peak_sust, "unit": "TFLOP/s", "frac": algo_tf / peak_sust,
                "traffic": traffic, "peak_source": which + ", sustained bf16 (kernel timed inside a back-to-back loop)",
                "achieved_is": "ALGORITHMIC flops 2*N*C*D per launch (one contraction, what the reference computes)",
                "executed": {"achieved": exec_tf, "frac_of_sustained": exec_tf / peak_sust,
                             "frac_of_burst": exec_tf / peak_burst, "peak_burst": peak_burst,
                             "note": "two-pass algorithm executes 4*N*C*D tensor flops (pass 1 max/argmax, pass 2 "
                                     "sum-exp); the north-star >=60 % target is read against this figure"},
                "kernel_ms": k_ms, "kernel_ms_min": min(kern_ms), "kernel_share_of_step": k_ms * args.steps / ms_total}
    if clocks and clocks.get("sm_mhz"):
        # the hardware ceiling at the clock the power cap allowed: 148 SMs x 8192 dense bf16 flop/cycle/SM
        hw = 148 * 8192 * clocks["sm_mhz"] * 1e6 / 1e12
        roofline["executed"]["hw_peak_at_observed_clock"] = hw
        roofline["executed"]["frac_of_hw_peak_at_observed_clock"] = exec_tf / hw
        roofline["executed"]["why_above_cublas"] = (
            "the measured peaks are cuBLAS bf16 GEMM rates under the same 1000 W cap, not the tensor pipe's limit; "
            "ncu shows sm__pipe_tensor_cycles_active 99.8 % for this kernel (profiles/r01b_score_fused_ncu_raw.csv)")

    # ---------------- CPU baseline (bounded sample, this box's host cores)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rows, fit_classes = min(4096, N_IMAGES), min(512, N_CLASSES)
        sample = _host_sample(rows)
        cpu_val, stages = cpu_reference_step(*sample, rows, fit_classes, threads)
        cpu = {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample_text(rows, fit_classes),
               "stages_s": stages}

    value = N_IMAGES * world * args.steps / (ms_total * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "config": config_dict(world), "clocks": clocks,
            "e2e": {"value": N_IMAGES * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / e2e_steps,
                    "steps": e2e_steps},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
            "dac_fit_ms": fit_ms, "check": summary, "numa_bound_cpus": numa}
    out.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
