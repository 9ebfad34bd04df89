"""Extract the per-kernel counters we quote from an `ncu --set full` report into profiles/<round>_ncu_full_summary.json.
Usage: python scripts/ncu_extract.py <report.ncu-rep> <capture-tag> [round-prefix, default r2]   (needs `ncu` on PATH; CPU only)"""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__cycles_active.avg", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
name_col = col.get("Kernel Name")
res = []
for r in data:
    name = r[name_col].split("(")[0].replace("csb::", "")
    e = {"capture": tag, "kernel": name}
    for m in METRICS:
        if m in col:
            e[m] = f"{r[col[m]]} {units[col[m]]}".strip()
    res.append(e)
path = os.path.join(ROOT, "profiles", (sys.argv[3] if len(sys.argv) > 3 else "r2") + "_ncu_full_summary.json")
cur = json.load(open(path)) if os.path.isfile(path) else []
cur = [c for c in cur if c.get("capture") != tag] + res
json.dump(cur, open(path, "w"), indent=1)
for e in res:
    print(e["kernel"], e.get("gpu__time_duration.sum"), e.get("dram__bytes_read.sum"), e.get("dram__bytes_write.sum"),
          e.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))
