"""Run under torchrun on N GPUs: the sharded path over real NCCL.
  - every rank scores its image shard (fused kernel) into an integer bin table; ONE all-reduce must
    reproduce, bit for bit, the table rank 0 gets by scoring all images alone;
  - AdaptiveECE / PIECE with `group=` (all-reduced radix histograms -> global quantile edges) must equal
    the single-GPU values on the gathered arrays;
  - CalibratedScorer(keep_outputs=True).evaluate() (bin table + per-class counts all-reduced, global quantile edges)
    must report on N ranks what one rank reports alone, macro-F1 included.
Prints one JSON line from rank 0; exits non-zero on any mismatch."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
from clip_calibration_b200 import native, pipeline, synth
from clip_calibration_b200 import table_math as tm
from clip_calibration_b200.tools import metrics

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))

case = synth.make_case("dist", 30011, 1000, 500, 512, 5, 0.25, seed=0)       # same on every rank
lo, hi = pipeline.shard_bounds(len(case.labels), rank, world)
scorer = pipeline.CalibratedScorer.from_dac(case.base_zs, case.txt_zs, case.base_tuned, case.txt_tuned, k=5,
                                            logit_scale=100.0, n_bins=10, share_text=True, overlap_fit=True)   # rank 0 fits on a side stream, NCCL broadcasts
pred, conf = scorer.score(case.img[lo:hi], case.labels[lo:hi])
reduced = scorer.reduced_table()                                               # NCCL all-reduce of 33 int64
labels_dev = torch.from_numpy(case.labels[lo:hi]).cuda()
prox_all = np.random.default_rng(1).random(len(case.labels)).astype(np.float32)
ace = float(metrics.AdaptiveECE(conf, pred, labels_dev, 10, group=dist.group.WORLD))
piece = float(metrics.PIECE(conf, torch.from_numpy(prox_all[lo:hi]).cuda(), pred, labels_dev, 10, 10, group=dist.group.WORLD))
ece = float(metrics.ECE(conf, pred, labels_dev, 10, group=dist.group.WORLD))

ev_scorer = pipeline.CalibratedScorer(case.txt_tuned, scorer.class_conf, 100.0, 10, keep_outputs=True)
ev_scorer.score(case.img[lo:hi], case.labels[lo:hi])
ev = ev_scorer.evaluate(proximity=torch.from_numpy(prox_all[lo:hi]).cuda())
f1 = float(metrics.macro_f1(pred, labels_dev, n_classes=1000, group=dist.group.WORLD))

ok = True
if rank == 0:
    solo = pipeline.CalibratedScorer(case.txt_tuned, scorer.class_conf, 100.0, 10)
    solo.group = False
    p_all, c_all = solo.score(case.img, case.labels)
    full = native.table_to_numpy(solo.table)
    ok &= bool(np.array_equal(full, reduced))
    lab_all = torch.from_numpy(case.labels).cuda()
    ace1 = float(metrics.AdaptiveECE(c_all, p_all, lab_all, 10))
    piece1 = float(metrics.PIECE(c_all, torch.from_numpy(prox_all).cuda(), p_all, lab_all, 10, 10))
    ece1 = float(tm.ece_from_table(full))
    ok &= abs(ace - ace1) < 1e-12 and abs(piece - piece1) < 1e-12 and abs(ece - ece1) < 1e-12
    solo_ev = pipeline.CalibratedScorer(case.txt_tuned, scorer.class_conf, 100.0, 10, keep_outputs=True, group=False)
    solo_ev.score(case.img, case.labels)
    ev1 = solo_ev.evaluate(proximity=torch.from_numpy(prox_all).cuda())
    keys = ("accuracy", "error_rate", "macro_f1", "confidence", "ece", "mce", "ace", "piece")
    ev_ok = all(abs(ev[k] - ev1[k]) < 1e-9 for k in keys) and abs(100.0 * f1 - ev1["macro_f1"]) < 1e-9
    ok &= ev_ok
    print(json.dumps({"evaluate_identical": bool(ev_ok), "macro_f1": [ev["macro_f1"], ev1["macro_f1"]]}), file=sys.stderr)
    print(json.dumps({"world": world, "tables_identical": bool(np.array_equal(full, reduced)), "ece": [ece, ece1],
                      "ace": [ace, ace1], "piece": [piece, piece1], "ok": bool(ok)}), flush=True)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.broadcast(flag, 0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
