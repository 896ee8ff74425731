"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the counters the roofline needs."""
import csv
import sys

WANT = [("gpu__time_duration.sum", "us"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"), ("dram__bytes_read.sum", "dramR"),
        ("dram__bytes_write.sum", "dramW"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts%"), ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"), ("smsp__inst_executed.sum", "inst"),
        ("sm__inst_executed.avg.per_cycle_elapsed", "ipc"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__registers_per_thread", "regs"), ("lts__t_sector_hit_rate.pct", "l2hit%")]


def main(path):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units = rows[hi], rows[hi + 1]
    idx = [(hdr.index(k) if k in hdr else None, short) for k, short in WANT]
    kn = hdr.index("Kernel Name")
    print("kernel | " + " | ".join(s for _, s in WANT))
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        vals = []
        for i, short in idx:
            v = r[i] if i is not None else "-"
            u = units[i] if i is not None else ""
            try:
                f = float(v.replace(",", ""))
                v = f"{f:.4g}" + (u if short in ("dramR", "dramW", "us") else "")
            except ValueError:
                pass
            vals.append(v)
        print(r[kn][:70] + " | " + " | ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
