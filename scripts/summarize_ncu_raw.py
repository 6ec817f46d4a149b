"""Summarise `ncu --page raw --csv` output: one row per launch with time, DRAM bytes, achieved GB/s, pipe activity.
python scripts/summarize_ncu_raw.py raw.csv "title" > profiles/x.md"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}; units = rows[1]
def g(r, k, d=0.0):
    try: return float(r[idx[k]].replace(",", ""))
    except Exception: return d
def unit(k): return units[idx[k]] if k in idx else ""
print("# %s" % sys.argv[2]); print()
print("Source: `%s` (ncu --set full --clock-control none; cold-cache serialised replays: compare shares and ratios)." % sys.argv[1].split("/")[-1]); print()
print("| kernel | grid | us | DRAM read | DRAM write | DRAM TB/s | DRAM %% of peak | L2 %% | tensor pipe %% | SM busy %% | regs |".replace("%%", "%"))
print("|---|---|---|---|---|---|---|---|---|---|---|")
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    name = name.split("(")[0].replace("void ", "").replace("stx::<unnamed>::", "")
    t = g(r, "gpu__time_duration.sum"); tu = unit("gpu__time_duration.sum")
    t_us = t / 1000.0 if tu.startswith("ns") or tu == "nsecond" else (t if tu.startswith("us") or tu == "usecond" else t * 1000.0)
    def b2mb(k):
        v = g(r, k); u = unit(k)
        return v if u.startswith("M") else (v / 1e6 if u in ("byte", "B") else (v / 1e3 if u.startswith("K") else v * 1e3 if u.startswith("G") else v))
    rd, wr = b2mb("dram__bytes_read.sum"), b2mb("dram__bytes_write.sum")
    print("| `%s` | %d | %.1f | %.1f MB | %.1f MB | %.2f | %.1f | %.1f | %.1f | %.1f | %d |" % (
        name, g(r, "launch__grid_size"), t_us, rd, wr, (rd + wr) / t_us * 1e3 / 1e3 if t_us else 0,
        g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"), g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
        g(r, "launch__registers_per_thread")))
