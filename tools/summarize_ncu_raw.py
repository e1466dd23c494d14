"""Per-launch table (markdown) from an `ncu --page raw --csv` export of a `--set full` capture.

    python tools/summarize_ncu_raw.py gpurun_out/conv_TAG.raw.csv
"""
import csv
import sys

COLS = [
    ("time us", "gpu__time_duration.sum", 1.0),
    ("DRAM rd MB", "dram__bytes_read.sum", 1.0),
    ("DRAM wr MB", "dram__bytes_write.sum", 1.0),
    ("DRAM %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("L2 hit %", "lts__t_sector_hit_rate.pct", 1.0),
    ("tensor pipe %", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("TC smem rd %", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", 1.0),
    ("SM %", "sm__throughput.avg.pct_of_peak_sustained_elapsed", 1.0),
    ("regs", "launch__registers_per_thread", 1.0),
]
TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def find(hdr, key):
    for i, h in enumerate(hdr):
        if h == key or h.endswith("." + key):
            return i
    return None


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki, gi = hdr.index("Kernel Name"), hdr.index("Grid Size")
    idx = [(title, find(hdr, key)) for title, key, _ in COLS]
    print("| # | kernel | grid | " + " | ".join(t for t, _ in idx) + " |")
    print("|---|---|---|" + "---|" * len(idx))
    tot_t = tot_r = tot_w = 0.0
    for n, r in enumerate(data):
        cells = []
        for title, i in idx:
            if i is None or r[i] in ("", "n/a", "no data"):
                cells.append("-")
                continue
            v = float(r[i].replace(",", ""))
            u = units[i]
            if "MB" in title:
                v *= TO_MB.get(u, 1.0)
            if "us" in title:
                v *= TO_US.get(u, 1.0)
            cells.append(f"{v:.1f}" if title != "regs" else f"{v:.0f}")
        name = r[ki].split("(")[0].replace("void ", "").replace("dram::", "")
        print(f"| {n} | {name} | {r[gi]} | " + " | ".join(cells) + " |")
        try:
            tot_t += float(cells[0]); tot_r += float(cells[1]); tot_w += float(cells[2])
        except ValueError:
            pass
    print(f"\ntotal: {tot_t:.1f} us, DRAM read {tot_r:.1f} MB, DRAM write {tot_w:.1f} MB")


if __name__ == "__main__":
    main()
