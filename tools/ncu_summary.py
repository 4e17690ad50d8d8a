"""Summarise `ncu -i rep --page raw --csv` into a small JSON for profiles/ (one entry per captured launch).
usage: ncu_summary.py raw.csv out.json 'command line that was profiled' [layer names, comma separated, in launch order]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_active",
        "lts__t_sectors.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_active.avg"]
ix = {h: i for i, h in enumerate(hdr)}
names = sys.argv[4].split(",") if len(sys.argv) > 4 else []
out = {"command": sys.argv[3], "units": {w: units[ix[w]] for w in want if w in ix}, "kernels": []}
for n, r in enumerate(data):
    e = {"layer": names[n] if n < len(names) else None, "kernel": r[ix["Kernel Name"]][:120]}
    for w in want:
        if w in ix:
            try: e[w] = float(r[ix[w]].replace(",", ""))
            except ValueError: e[w] = r[ix[w]]
    out["kernels"].append(e)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print("wrote", sys.argv[2], len(out["kernels"]), "launches")
