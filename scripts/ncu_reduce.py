"""Reduce an `ncu --csv --page raw` log to one line per distinct (kernel, grid): mean time, DRAM bytes, DRAM %, tensor-pipe %,
issue %, achieved warps %, registers.   python scripts/ncu_reduce.py raw.csv > summary.csv"""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr, units = rows[0], rows[1]
col = lambda name: next((i for i, h in enumerate(hdr) if h == name), None)
want = {"time_ms": "gpu__time_duration.sum", "dram_rd_GB": "dram__bytes_read.sum", "dram_wr_GB": "dram__bytes_write.sum",
        "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "tensor_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "tensor_inst_pct": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
        "issue_pct": "sm__inst_issued.avg.pct_of_peak_sustained_elapsed", "warps_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "sm_ghz": "sm__cycles_elapsed.avg.per_second"}
idx = {k: col(v) for k, v in want.items()}
scale = {"usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6, "second": 1e3, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3,
         "byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
agg = collections.OrderedDict()
ki = col("Kernel Name")
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("ccal::", "")
    key = (name, r[idx["grid"]] if idx["grid"] is not None else "")
    vals = {}
    for k, i in idx.items():
        if i is None or r[i] == "":
            continue
        try:
            v = float(r[i].replace(",", ""))
        except ValueError:
            continue
        vals[k] = v * scale.get(units[i], 1.0) if k in ("time_ms", "dram_rd_GB", "dram_wr_GB") else v
    a = agg.setdefault(key, [0, collections.defaultdict(float)])
    a[0] += 1
    for k, v in vals.items():
        a[1][k] += v
cols = [k for k in want if k != "grid"]
print("kernel,grid,launches," + ",".join(cols))
for (name, grid), (n, s) in agg.items():
    print(f'"{name}",{grid},{n},' + ",".join(f"{s[k] / n:.4g}" if k in s else "" for k in cols))
