"""CPU, world_size 2, gloo: the multi-GPU decomposition.  Each rank bins its own image shard
into an integer table; ONE all-reduce(sum) of the tables must equal the table of the whole set
bit for bit, and the metrics composed from it must equal the single-process ones."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clip_calibration_b200 import pipeline
from clip_calibration_b200 import table_math as tm
from oracle import cpu_oracle as orc


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, conf, pred, gt, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pipeline.shard_bounds(len(conf), rank, world)
    tab = orc.bin_table(conf[lo:hi], pred[lo:hi], gt[lo:hi], tm.uniform_thresholds(10))
    t = torch.from_numpy(tab.view(np.int64).copy())
    dist.all_reduce(t)                          # the only collective on the data path
    if rank == 0:
        np.save(out, t.numpy().view(np.uint64))
    dist.destroy_process_group()


def test_two_rank_table_allreduce_is_exact(tmp_path):
    rng = np.random.default_rng(3)
    n = 20001
    conf = np.where(rng.random(n) < 0.1, 1.0, rng.random(n)).astype(np.float32)
    pred = rng.integers(0, 5, n)
    gt = rng.integers(0, 5, n)
    out = str(tmp_path / "table.npy")
    mp.spawn(_worker, args=(2, _free_port(), conf, pred, gt, out), nprocs=2, join=True)
    reduced = np.load(out)
    full = orc.bin_table(conf, pred, gt, tm.uniform_thresholds(10))
    assert np.array_equal(reduced, full)
    assert abs(tm.ece_from_table(reduced) - orc.ece(conf, pred, gt, 10)) < 1e-7
    assert abs(tm.mce_from_table(reduced) - orc.mce(conf, pred, gt, 10)) < 1e-7


def _class_counts(pred, gt, c):
    counts = np.zeros((c, 3), np.int64)
    np.add.at(counts[:, 0], gt[pred == gt], 1)
    np.add.at(counts[:, 1], pred[pred != gt], 1)
    np.add.at(counts[:, 2], gt[pred != gt], 1)
    return counts


def _worker_counts(rank, world, port, pred, gt, c, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = pipeline.shard_bounds(len(pred), rank, world)
    t = torch.from_numpy(_class_counts(pred[lo:hi], gt[lo:hi], c))
    dist.all_reduce(t)                          # per-class {tp, fp, fn}: 3*C integers
    if rank == 0:
        np.save(out, t.numpy())
    dist.destroy_process_group()


def test_two_rank_macro_f1_from_allreduced_class_counts(tmp_path):
    """macro-F1 (evaluators/vl_evaluator.py:74-79) of the whole set from per-shard class counts."""
    import warnings
    from sklearn.metrics import f1_score
    rng = np.random.default_rng(5)
    n, c = 15001, 37
    gt = rng.integers(0, c, n)
    pred = np.where(rng.random(n) < 0.55, gt, rng.integers(0, c, n))
    out = str(tmp_path / "counts.npy")
    mp.spawn(_worker_counts, args=(2, _free_port(), pred, gt, c, out), nprocs=2, join=True)
    reduced = np.load(out)
    assert np.array_equal(reduced, _class_counts(pred, gt, c))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = f1_score(gt, pred, average="macro", labels=np.unique(gt))
    assert abs(tm.macro_f1_from_counts(reduced) - want) < 1e-12
