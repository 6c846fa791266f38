"""One-call scoring + calibration + ECE over image shards (additive API).

What the reference does across VLBaseLearner.test (trainers/classification/base_learner.py:84-144),
VLCalibration.fit/predict (trainers/calibration/vl_calibrator.py:71-109) and
VLClassification.evaluate (evaluators/vl_evaluator.py:59-92) - contraction, DAC, softmax, argmax,
confidence gather, ECE/MCE/accuracy - is here one fused kernel per image shard plus, across the
GPUs of a box, ONE all-reduce of a 33-integer bin table.

Sharding: images (rows) are independent, so each rank owns a contiguous slice of the images;
the text features, the per-class multipliers and the bin edges are replicated (the DAC fit is
recomputed on every rank: it is tiny and avoids a broadcast).  The table is integer / fixed
point, so the 1-GPU and the N-GPU results are bit-identical.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import native
from . import table_math as tm


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) slice of n images for `rank` of `world`."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


MAX_IMAGES_PER_TABLE = 1 << 24         # 64-bit sum of round(conf * 2^40), conf <= 1

def dac_fit_sharded(base_zs, cur_zs, base_tuned, cur_tuned, k: int, group=None) -> torch.Tensor:
    """DAC multipliers [C] (float32, on this rank's GPU) with the fit itself sharded over the ranks of `group`:
    the classes of the test vocabulary are independent queries against the (replicated) base rows
    (distanse_aware_calibration.py:25-42 is a loop over them), so rank r fits classes [r*C/W, (r+1)*C/W) and ONE
    all-gather of C floats hands every rank the full vector.  All four matrices are float32 CUDA tensors and are
    the same on every rank; the result is identical to the unsharded fit (per-class arithmetic does not depend on
    which other classes are in the call).  Strong scaling: the fit's share of a step stays constant instead of
    growing with the rank count."""
    dist = torch.distributed
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return native.dac_fit(base_zs, cur_zs, base_tuned, cur_tuned, k)[0]
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    c = int(cur_zs.shape[0])
    per = -(-c // world)
    lo, hi = min(c, rank * per), min(c, (rank + 1) * per)
    part = torch.ones(per, dtype=torch.float32, device=cur_zs.device)
    if hi > lo:
        part[: hi - lo] = native.dac_fit(base_zs, cur_zs[lo:hi].contiguous(), base_tuned, cur_tuned[lo:hi].contiguous(), k)[0]
    full = torch.empty(per * world, dtype=torch.float32, device=cur_zs.device)
    dist.all_gather_into_tensor(full, part, group=group)
    return full[:c].contiguous()


_SIDE_STREAMS = {}


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    """One long-lived side stream per device: the caching allocator keeps a block pool per stream, so a fresh stream
    per call would go back to cudaMalloc (slow, device-synchronising) for every temporary."""
    key = (device.type, device.index)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


_STAGING = {}
_PIPE_DEBUG = bool(os.environ.get("CCAL_PIPE_DEBUG"))      # development aid: keep per-stage events of from_dac(overlap_fit=True)


def _staging(device: torch.device, rows: int, d: int, dtype, tag: str = "ring") -> dict:
    """Device staging buffers of accumulate_host, kept per (device, shape): three [rows, d] feature buffers + label
    buffers with their `copied` / `consumed` events.  Because they persist, a new call only has to wait for the
    previous user of the SAME buffer (its `consumed` event) instead of ordering its copies after everything on the
    compute stream."""
    key = (device.type, device.index, int(rows), int(d), dtype, tag)
    st = _STAGING.get(key)
    if st is None:
        if len(_STAGING) > 8:                      # shapes changed: drop the old buffers
            _STAGING.clear()
        n_buf = 3 if tag == "ring" else 1
        st = {"bufs": [torch.empty((rows, d), dtype=dtype, device=device) for _ in range(n_buf)],
              "lbufs": [torch.empty(rows, dtype=torch.int64, device=device) for _ in range(n_buf)],
              "copied": [torch.cuda.Event() for _ in range(n_buf)],
              "consumed": [torch.cuda.Event() for _ in range(n_buf)],
              "next": 0}                               # ring position, carried from call to call
        cur = torch.cuda.current_stream(device)
        for ev in st["consumed"]:
            ev.record(cur)                         # allocation (and any earlier use of that memory) precedes the first copy
        _STAGING[key] = st
    return st


_COPY_STREAMS = {}


def _copy_stream(device: torch.device) -> torch.cuda.Stream:
    key = (device.type, device.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return _COPY_STREAMS[key]


class PendingTable:
    """A reduced bin table on its way to the host (CalibratedScorer.reduced_table_async)."""

    def __init__(self, host: torch.Tensor, done: torch.cuda.Event, keep_alive=None):
        self._host, self._done, self._keep = host, done, keep_alive
        self._out = None

    def ready(self) -> bool:
        return self._out is not None or self._done.query()

    def result(self) -> np.ndarray:
        if self._out is None:
            self._done.synchronize()
            out = self._host.numpy().view(np.uint64).copy()
            self._host = self._keep = None
            # the per-bin confidence sums are 64-bit fixed point (2^-40 units): 2^24 images of confidence ~1 fill them
            if tm.total_count(out) > MAX_IMAGES_PER_TABLE:
                raise OverflowError(f"{tm.total_count(out)} images in one bin table: the 2^-40 fixed-point confidence "
                                    f"sums hold at most {MAX_IMAGES_PER_TABLE}; evaluate in shards of <= 2^24 images "
                                    "(reset() between them) and add the float results")
            self._out = out
        return self._out


class CalibratedScorer:
    """Text side of the problem, resident on this rank's GPU, plus the running bin table."""

    def __init__(self, text_features, class_conf=None, logit_scale: float = 100.0, n_bins: int = 10,
                 operand_dtype=None, group=None, device=None, keep_outputs: bool = False):
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        probe = text_features if isinstance(text_features, torch.Tensor) else torch.from_numpy(np.asarray(text_features)[:1])
        self.operand_dtype = native.operand_dtype_for(probe, operand_dtype)
        self.txt = self._features(text_features)
        self.class_conf = None
        if class_conf is not None:
            cc = class_conf if isinstance(class_conf, torch.Tensor) else torch.from_numpy(np.asarray(class_conf))
            self.class_conf = cc.to(device=self.device, dtype=torch.float32).contiguous()
        self.logit_scale = float(logit_scale)
        self.n_bins = int(n_bins)
        self.thresholds = tm.uniform_thresholds(self.n_bins)
        self.group = group          # None = default process group (if initialised); False = never reduce
        self.table = native.new_table(self.n_bins, device=self.device)
        self._copy_stream = None
        self._pending_img, self._pending_lab, self._pending_rows = [], [], 0
        # evaluator mode: per-image (pred, conf, label) stay on the device (16 B/image) and per-class {tp, fp, fn}
        # are counted, so that evaluate() can report every key of the reference's evaluator (macro-F1, ACE, PIECE)
        self._fit_done = None                 # event of a DAC fit still running on a side stream (from_dac(overlap_fit=True))
        self._device_was_busy = None          # from_dac(overlap_fit=True): was the compute stream busy when it was called?
        self.keep_outputs = bool(keep_outputs)
        self._kept = []                       # [(pred int32, conf float32, labels int64)] per scored shard
        self.class_counts = None

    # ------------------------------------------------------------------ construction helpers
    def _features(self, x) -> torch.Tensor:
        t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
        return t.detach().to(self.device).to(self.operand_dtype).contiguous()

    @classmethod
    def from_dac(cls, base_zs, cur_zs, base_tuned, cur_tuned, k: int = 5, share_text: bool = False,
                 overlap_fit: bool = False, **kw):
        """Fit DAC on the four text matrices and score against the tuned test-vocabulary features - what
        VLBaseLearner.test + build_dac_calibrator set up (base_learner.py:117-119, vl_calibrator.py:155-180).
        Host inputs are uploaded once.  By default every rank does this redundantly.  With `share_text` (all ranks
        hold the SAME text features and call together) only rank 0 of the group uploads and fits; the scoring
        operand and the per-class multipliers reach the other ranks by NCCL broadcast over NVLink (51 MB + 0.2 MB
        at 49,408 x 512) instead of N uploads competing for the host's memory bandwidth.
        With `overlap_fit` only the scoring operand is uploaded (and, with share_text, broadcast) on the compute
        stream; the other three matrices, the DAC fit and the broadcast of the multipliers go to a side stream, and
        `accumulate_host` runs the first chunks' pass 1 (which needs no multipliers) underneath them - the fit
        leaves the critical path on every rank."""
        from .trainers.calibration.distanse_aware_calibration import DistanseAwareCalibration, _to_cuda_f32
        group = kw.get("group")
        shared = (share_text and group is not False and torch.distributed.is_available()
                  and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1)
        is_root = not shared or torch.distributed.get_rank(group) == 0
        src = (torch.distributed.get_global_rank(group, 0) if group is not None else 0) if shared else None

        # The scoring operand's dtype is decided ONCE, from the features as the caller holds them (fp16 / bf16 stay
        # 16-bit, anything else is scored in the fp32 split mode), and handed to every rank explicitly: the root's
        # device copy is widened to fp32 for the DAC fit, and inferring the dtype from THAT would make the root
        # hold an fp32 operand while the other ranks allocate a 16-bit one for the same broadcast.
        probe = cur_tuned if isinstance(cur_tuned, torch.Tensor) else torch.from_numpy(np.asarray(cur_tuned)[:1])
        kw = dict(kw, operand_dtype=native.operand_dtype_for(probe, kw.get("operand_dtype")))

        def empty_text():
            rows, dim = (int(x) for x in cur_tuned.shape)
            dev = torch.device(kw["device"]) if kw.get("device") is not None else torch.device("cuda", torch.cuda.current_device())
            return torch.empty((rows, dim), dtype=kw["operand_dtype"], device=dev)

        if overlap_fit:
            # Host inputs go up on the COPY stream, scoring operand first: the copy stream is a FIFO shared with
            # accumulate_host's image chunks, so a call queued while the previous evaluation is still being scored
            # uploads its text side (and then its first image chunks) underneath that evaluation's last kernels
            # instead of behind them.  The buffers are allocated under the copy stream (a block of the compute
            # stream's pool could still be in use by kernels queued there) and handed to their readers by event.
            dev = torch.device(kw["device"]) if kw.get("device") is not None else torch.device("cuda", torch.cuda.current_device())
            comp = torch.cuda.current_stream(dev)
            busy = not comp.query()          # earlier work (the previous evaluation) is still running on the device
            side = _side_stream(dev)
            copy = _copy_stream(dev)
            names = ("current_text_features_tuned", "base_text_features_zs", "current_text_features_zs",
                     "base_text_features_tuned")
            staged, uploaded = [], False
            dbg = {}
            if is_root:
                for x, nm in zip((cur_tuned, base_zs, cur_zs, base_tuned), names):
                    t = x.detach() if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
                    if t.dim() != 2:
                        raise ValueError(f"{nm}: expected a [rows, D] matrix, got shape {tuple(t.shape)}")
                    ev = None
                    if not t.is_cuda:
                        with torch.cuda.stream(copy):
                            t = t.contiguous().to(device=dev, non_blocking=t.is_pinned())
                            ev = torch.cuda.Event()
                            ev.record(copy)
                        uploaded = True
                    staged.append((t, ev))
                op_src, op_ev = staged[0]
                if op_ev is not None:
                    comp.wait_event(op_ev)
                    op_src.record_stream(comp)
            obj = cls(op_src if is_root else empty_text(), None, **kw)       # narrows / widens on the compute stream
            if shared:
                torch.distributed.broadcast(obj.txt, src=src, group=group)
            if is_root and staged[0][1] is None:
                side.wait_stream(comp)                               # device inputs were produced on the compute stream
            dac = DistanseAwareCalibration() if is_root else None
            with torch.cuda.stream(side):
                if is_root:
                    for t, ev in staged:
                        if ev is not None:
                            side.wait_event(ev)
                            t.record_stream(side)
                    f32 = [_to_cuda_f32(t, nm) for (t, _), nm in zip(staged, names)]
                    dac.fit(f32[1], f32[2], f32[3], f32[0], k, sync_host_copy=False)
                    cc = dac.class_confidence_device
                    if _PIPE_DEBUG:
                        dbg["fit_end"] = torch.cuda.Event()
                        dbg["fit_end"].record(side)
                else:
                    cc = torch.empty(obj.txt.shape[0], dtype=torch.float32, device=obj.device)
                if shared:
                    torch.distributed.broadcast(cc, src=src, group=group)
                obj._fit_done = torch.cuda.Event()
                obj._fit_done.record(side)
            if _PIPE_DEBUG:
                dbg["staged"] = [ev for _, ev in staged if ev is not None]
                dbg["fit_done"] = obj._fit_done
                obj._dbg = dbg
            obj.class_conf = cc
            obj.class_conf.record_stream(comp)
            obj.dac = dac
            obj._device_was_busy = busy
            return obj

        if not is_root:
            obj = cls(empty_text(), torch.empty(int(cur_tuned.shape[0]), dtype=torch.float32,
                                                device=torch.device(kw["device"]) if kw.get("device") is not None
                                                else torch.device("cuda", torch.cuda.current_device())), **kw)
            obj.dac = None
        else:
            cur_tuned_dev = _to_cuda_f32(cur_tuned, "current_text_features_tuned")
            dac = DistanseAwareCalibration()
            dac.fit(base_zs, cur_zs, base_tuned, cur_tuned_dev, k, sync_host_copy=False)
            obj = cls(cur_tuned_dev, dac.class_confidence_device, **kw)
            obj.dac = dac
        if shared:
            torch.distributed.broadcast(obj.txt, src=src, group=group)
            torch.distributed.broadcast(obj.class_conf, src=src, group=group)
        return obj

    # ------------------------------------------------------------------ scoring
    def reset(self):
        self.table.zero_()
        self._pending_img, self._pending_lab, self._pending_rows = [], [], 0
        self._kept, self.class_counts = [], None

    def _await_fit(self) -> None:
        """Order the compute stream after a DAC fit that from_dac(overlap_fit=True) left running on a side stream."""
        if self._fit_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._fit_done)
            self._fit_done = None

    def _keep(self, pred, conf, labels) -> None:
        if self.class_counts is None:
            self.class_counts = torch.zeros((self.txt.shape[0], 3), dtype=torch.int64, device=self.device)
        native.class_counts(pred, labels, self.txt.shape[0], self.class_counts)
        self._kept.append((pred, conf, labels))

    def score(self, image_features, labels=None, accumulate: bool = True):
        """Device-resident shard -> (pred, conf); with labels the shard is also binned into the
        running table.  No logits are materialised."""
        img = self._features(image_features)
        if labels is not None:
            labels = (labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels)))
            labels = labels.to(device=self.device, dtype=torch.int64)
        use_table = labels is not None and accumulate
        if self._fit_done is not None and self.operand_dtype in (torch.float16, torch.bfloat16) and img.shape[0] \
                and not native.guess_pipeline_applies(img.shape[0], self.txt.shape[0], img.shape[1], img.dtype):
            # the DAC fit is still running on its side stream (from_dac(overlap_fit=True)): pass 1 needs no
            # multipliers and runs underneath it; pass 2 waits for the fit.  Bit-identical to the fused launch.
            # (Large shards stay on the one-call path: its FP8 guess pass costs half a bf16 pass 1.)
            dotmax, pred = native.score_pass1(img, self.txt)
            self._await_fit()
            conf = native.score_pass2(img, self.txt, dotmax, pred, self.class_conf, self.logit_scale, labels,
                                      self.thresholds if use_table else None, self.table if use_table else None)
        else:
            self._await_fit()
            pred, conf, _ = native.score_fused(img, self.txt, self.class_conf, self.logit_scale, labels,
                                               self.thresholds if use_table else None, self.table if use_table else None)
        if use_table and self.keep_outputs:
            self._keep(pred, conf, labels)
        return pred, conf

    def add(self, image_features, labels, flush_rows: int = 32768) -> None:
        """Per-batch entry point for an evaluation loop (the reference feeds 100 images at a time,
        base_learner.py:84-88).  A 100-row launch would occupy ONE of the 148 SMs, so batches are parked on
        the device (operand dtype, no host copies) and scored together once `flush_rows` have arrived;
        `summary()` / `reduced_table()` flush the remainder.  Only the bin table is produced on this path."""
        img = self._features(image_features)
        lab = (labels if isinstance(labels, torch.Tensor) else torch.from_numpy(np.asarray(labels)))
        self._pending_img.append(img)
        self._pending_lab.append(lab.to(device=self.device, dtype=torch.int64))
        self._pending_rows += int(img.shape[0])
        if self._pending_rows >= flush_rows:
            self.flush()

    def flush(self) -> None:
        if self._pending_rows:
            img = torch.cat(self._pending_img) if len(self._pending_img) > 1 else self._pending_img[0]
            lab = torch.cat(self._pending_lab) if len(self._pending_lab) > 1 else self._pending_lab[0]
            self._pending_img, self._pending_lab, self._pending_rows = [], [], 0
            self._await_fit()
            pred, conf, _ = native.score_fused(img, self.txt, self.class_conf, self.logit_scale, lab, self.thresholds,
                                               self.table, want_pred=self.keep_outputs, want_conf=self.keep_outputs)
            if self.keep_outputs:
                self._keep(pred, conf, lab)

    def accumulate_host(self, image_features: torch.Tensor, labels: torch.Tensor, chunk_rows: int = 131072,
                        keep_outputs: bool = False, ramp: bool = True):
        """End-to-end path for HOST inputs (ideally pinned): the shard is cut into row chunks,
        chunk i+1 is copied host->device on a side stream while chunk i is being scored, and
        only the bin table (and optionally pred/conf) ever comes back.  With `ramp` the first chunks
        are small (16k rows, doubling up to chunk_rows) so scoring starts after a few MB have arrived instead
        of after a full chunk - the un-overlapped head of the pipeline is what 8 ranks sharing one
        host's memory bandwidth feel most.  The device staging buffers are kept per (device, shape) and are ordered
        by their own events, so the first copies do not wait for whatever else is queued on the compute stream
        (another rank's text upload + broadcast, the DAC fit)."""
        if image_features.is_cuda:
            raise ValueError("accumulate_host expects host tensors; use score() for device tensors")
        if image_features.dtype != self.operand_dtype:
            raise ValueError(f"host features must already be {self.operand_dtype} (convert once, outside the hot loop)")
        n, d = image_features.shape
        labels = labels.to(torch.int64)
        comp = torch.cuda.current_stream(self.device)
        # Latency mode or throughput mode?  If the device is idle, the first launch waits for the first bytes: small
        # first chunks (ramp) and pass 1 underneath the DAC fit shorten that head.  If earlier work is still running
        # (an evaluation loop that queues the next evaluation before reading the last result), the uploads and the
        # fit are hidden underneath it anyway, and full-size chunks through the one-call scoring path are cheaper.
        busy = self._device_was_busy if self._device_was_busy is not None else not comp.query()
        self._device_was_busy = None
        if self._copy_stream is None:
            self._copy_stream = _copy_stream(self.device)
        copy = self._copy_stream
        chunk_rows = max(128, min(int(chunk_rows), n))
        # the scoring kernels are persistent over 256-row tiles, one CTA pair per two SMs: a chunk that is a whole number
        # of waves leaves no SM idle in the last wave of each of its launches
        wave = 256 * max(1, torch.cuda.get_device_properties(self.device).multi_processor_count // 2)
        if n > chunk_rows >= 4 * wave:
            per = -(-n // -(-n // chunk_rows))                       # balanced: ceil(n / number of chunks)
            chunk_rows = -(-per // wave) * wave
        bounds, lo = [], 0
        sizes = []
        if ramp and n > chunk_rows and not busy:
            step = min(16384, chunk_rows // 2)
            while step < chunk_rows and sum(sizes) + step < n - chunk_rows // 2:
                sizes.append(step)
                step *= 2
        while lo < n:
            step = max(128, sizes.pop(0)) if sizes else chunk_rows
            bounds.append((lo, min(n, lo + step)))
            lo += step
        want_out = keep_outputs or self.keep_outputs
        preds, confs = [], []
        stage = _staging(self.device, chunk_rows, d, self.operand_dtype)
        bufs, lbufs, copied, consumed = stage["bufs"], stage["lbufs"], stage["copied"], stage["consumed"]
        if self._fit_done is not None and not busy and len(bounds) > 3 and self.operand_dtype in (torch.float16, torch.bfloat16):
            # The DAC fit is still running on its side stream (from_dac(overlap_fit=True)).  Pass 1 needs no
            # multipliers: run it on the head chunks as they arrive, underneath the fit and its uploads, then wait
            # for the fit and finish the head with ONE pass-2 launch.  Same results as the fused launch, bit for bit.
            n_head = 3
            head, bounds = bounds[:n_head], bounds[n_head:]
            head_rows = head[-1][1]
            hstage = _staging(self.device, head_rows, d, self.operand_dtype, tag="head")
            hbuf, hlab = hstage["bufs"][0], hstage["lbufs"][0]
            parts = []
            for j, (lo, hi) in enumerate(head):
                arrived = torch.cuda.Event()
                with torch.cuda.stream(copy):
                    if j == 0:
                        copy.wait_event(hstage["consumed"][0])       # the previous call's pass 2 has read the buffer
                    hbuf[lo:hi].copy_(image_features[lo:hi], non_blocking=True)
                    hlab[lo:hi].copy_(labels[lo:hi], non_blocking=True)
                    arrived.record(copy)
                comp.wait_event(arrived)
                parts.append(native.score_pass1(hbuf[lo:hi], self.txt))
            dotmax = torch.cat([q[0] for q in parts])
            pred = torch.cat([q[1] for q in parts])
            self._await_fit()
            conf = native.score_pass2(hbuf[:head_rows], self.txt, dotmax, pred, self.class_conf, self.logit_scale,
                                      hlab[:head_rows], self.thresholds, self.table, want_conf=want_out)
            if self.keep_outputs:
                self._keep(pred, conf, hlab[:head_rows].clone())
            hstage["consumed"][0].record(comp)
            if keep_outputs:
                preds.append(pred)
                confs.append(conf)
        self._await_fit()
        for lo, hi in bounds:
            # the ring position persists across calls: the first chunk of a call queued behind another evaluation goes
            # to the buffer that frees up first, not to the one that evaluation's last chunk is still being scored from
            b = stage["next"]
            stage["next"] = (b + 1) % len(bufs)
            with torch.cuda.stream(copy):
                copy.wait_event(consumed[b])                         # (recorded at creation / by the previous user)
                bufs[b][: hi - lo].copy_(image_features[lo:hi], non_blocking=True)
                lbufs[b][: hi - lo].copy_(labels[lo:hi], non_blocking=True)
                copied[b].record(copy)
            comp.wait_event(copied[b])
            pred, conf, _ = native.score_fused(bufs[b][: hi - lo], self.txt, self.class_conf, self.logit_scale,
                                               lbufs[b][: hi - lo], self.thresholds, self.table,
                                               want_pred=want_out, want_conf=want_out)
            if self.keep_outputs:
                self._keep(pred, conf, lbufs[b][: hi - lo].clone())
            consumed[b].record(comp)
            if keep_outputs:
                preds.append(pred)
                confs.append(conf)
        if keep_outputs:
            return torch.cat(preds), torch.cat(confs)
        return None

    # ------------------------------------------------------------------ results
    def reduced_table_async(self) -> "PendingTable":
        """Queue the reduction of the bin table (one NCCL all-reduce of 3*(n_bins+1) int64 on the compute stream) and
        its device->host copy into pinned memory, and return at once: `.result()` waits for exactly that copy.  An
        evaluation loop over several shards / datasets calls this, queues the NEXT evaluation, and only then reads the
        result - the next evaluation's uploads and first launches then run underneath this one's last kernels."""
        self.flush()
        t = self.table
        if self.group is not False and torch.distributed.is_available() and torch.distributed.is_initialized() \
                and torch.distributed.get_world_size(self.group) > 1:
            t = t.clone()
            torch.distributed.all_reduce(t, group=self.group)
        host = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        host.copy_(t, non_blocking=True)
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream(self.device))
        return PendingTable(host, done, t)

    def reduced_table(self) -> np.ndarray:
        """The bin table summed over all ranks of `group`, as a host uint64 array (blocking)."""
        return self.reduced_table_async().result()

    def _reduce_group(self):
        """The process group the tables are reduced over, or None when there is nothing to reduce."""
        if self.group is False or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return None
        group = self.group if self.group is not None else torch.distributed.group.WORLD
        return group if torch.distributed.get_world_size(group) > 1 else None

    def evaluate(self, proximity=None, piece_bins: int = 10) -> dict:
        """Every key of the reference's `VLClassification.evaluate` result (evaluators/vl_evaluator.py:59-116), as
        percentages, from the fused path: accuracy / error_rate / confidence / ece / mce from the bin table, macro_f1
        from the per-class counts, ace (and piece when `proximity` [N] for this rank's images, in scoring order, is
        given) from the kept per-image confidences.  Needs keep_outputs=True.  Multi-GPU: all ranks call it together."""
        if not self.keep_outputs:
            raise RuntimeError("evaluate() needs CalibratedScorer(..., keep_outputs=True); summary() works without")
        from .evaluators.vl_evaluator import results_from_tables
        from .tools import metrics
        table = self.reduced_table()                      # flushes pending batches
        group = self._reduce_group()
        if self.class_counts is None:
            self.class_counts = torch.zeros((self.txt.shape[0], 3), dtype=torch.int64, device=self.device)
        counts = self.class_counts.clone()
        if group is not None:
            torch.distributed.all_reduce(counts, group=group)
        cat = lambda i, dt: (torch.cat([k[i] for k in self._kept]) if self._kept
                             else torch.empty(0, dtype=dt, device=self.device))
        pred, conf, labels = cat(0, torch.int32), cat(1, torch.float32), cat(2, torch.int64)
        results = results_from_tables(table, counts.cpu().numpy())
        results["ace"] = 100.0 * float(metrics.AdaptiveECE(conf, pred, labels, self.n_bins, group=group))
        if proximity is not None:
            results["piece"] = 100.0 * float(metrics.PIECE(conf, proximity, pred, labels, piece_bins, self.n_bins,
                                                           group=group))
        results["bin_table"] = table
        return results

    def summary(self) -> dict:
        table = self.reduced_table()
        return {"n": tm.total_count(table), "accuracy": tm.accuracy(table), "confidence": tm.mean_confidence(table),
                "ece": float(tm.ece_from_table(table)), "mce": float(tm.mce_from_table(table)), "table": table}


def score_and_ece(image_features, text_features, labels, class_conf=None, logit_scale: float = 100.0,
                  n_bins: int = 10, operand_dtype=None, group=None) -> dict:
    """features -> {accuracy, confidence, ece, mce, table, pred, conf} in one fused pass."""
    scorer = CalibratedScorer(text_features, class_conf, logit_scale, n_bins, operand_dtype, group)
    pred, conf = scorer.score(image_features, labels)
    out = scorer.summary()
    out["pred"], out["conf"] = pred, conf
    return out
