"""Development aid (GPU box, torchrun): ONLY the pipelined end-to-end evaluation loop of bench.py, many steps, with a
watchdog that reports which stream / stage is stuck if a step makes no progress (CCAL_PIPE_DEBUG=1 events).

    CCAL_PIPE_DEBUG=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 scripts/gpu_e2e_loop.py --steps 40
"""
import argparse, datetime, faulthandler, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from clip_calibration_b200 import pipeline, table_math as tm

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--rows", type=int, default=1_000_000)
ap.add_argument("--stall", type=float, default=25.0)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=90))
w = bench.WORKLOADS["openvocab"]
img, labels, txt_zs, txt_tuned = bench.make_device_data(w, args.rows, seed=1000 + rank)
host_img, host_labels = img.cpu().pin_memory(), labels.cpu().pin_memory()
host_txt = {k: v.to(torch.bfloat16).cpu().pin_memory() for k, v in
            {"bz": txt_zs[:w.n_base], "cz": txt_zs, "bt": txt_tuned[:w.n_base], "ct": txt_tuned}.items()}
del img
torch.cuda.empty_cache()

progress = {"t": time.time(), "step": -1, "scorers": []}


def watchdog():
    while True:
        time.sleep(1.0)
        if time.time() - progress["t"] > args.stall:
            dev = torch.device("cuda", local)
            lines = [f"[rank {rank}] STALL at step {progress['step']}: comp idle={torch.cuda.current_stream(dev).query()} "
                     f"side idle={pipeline._side_stream(dev).query()} copy idle={pipeline._copy_stream(dev).query()}"]
            for i, sc in progress["scorers"][-3:]:
                d = getattr(sc, "_dbg", {})
                lines.append(f"[rank {rank}]   step {i}: staged={[e.query() for e in d.get('staged', [])]} "
                             f"fit_end={d['fit_end'].query() if 'fit_end' in d else None} "
                             f"fit_done={d['fit_done'].query() if 'fit_done' in d else None}")
            print("\n".join(lines), file=sys.stderr, flush=True)
            faulthandler.dump_traceback(file=sys.stderr)
            os._exit(3)


threading.Thread(target=watchdog, daemon=True).start()
pending = []


def queue(i):
    sc = pipeline.CalibratedScorer.from_dac(host_txt["bz"], host_txt["cz"], host_txt["bt"], host_txt["ct"], k=w.k,
                                            logit_scale=100.0, n_bins=10, operand_dtype=torch.bfloat16,
                                            share_text=world > 1, overlap_fit=True)
    progress["scorers"].append((i, sc))
    del progress["scorers"][:-4]
    sc.accumulate_host(host_img, host_labels, chunk_rows=262144)
    pending.append(sc.reduced_table_async())


def drain(keep):
    while len(pending) > keep:
        t = pending.pop(0).result()
        assert tm.total_count(t) == args.rows * world


t0 = time.time()
for rep in range(4):                       # like bench: barrier, then a burst of pipelined steps, drained at the end
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    for i in range(args.steps // 4):
        progress.update(t=time.time(), step=rep * 100 + i)
        queue(rep * 100 + i)
        drain(1)
    drain(0)
    if rank == 0:
        print(f"burst {rep} done at {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
progress["t"] = time.time() + 1e9
print(f"[rank {rank}] ok: {args.steps} pipelined steps in {time.time() - t0:.1f}s", flush=True)
if world > 1:
    dist.destroy_process_group()
