"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (markdown on stdout): launches,
total ms and share over the WHOLE run, ours (namespace dram::) vs everything else.
    python tools/summarize_launches_by_kernel.py gpurun_out/launches_train_TAG.csv [steps-in-the-run]
"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        name = re.sub(r"<.*", "", name) if not name.startswith("dram::") else name
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[mi].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if k.startswith("dram::"))
    print(f"{sum(v[0] for v in agg.values())} launches, {tot:.2f} ms under ncu over {steps} step(s) "
          f"({tot / steps:.2f} ms/step); hand-written kernels (dram::) {100 * ours / tot:.1f} % of the time\n")
    print("| kernel | launches/step | ms/step | share |")
    print("|---|---|---|---|")
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| {k[:90]} | {n / steps:.1f} | {ms / steps:.3f} | {100 * ms / tot:.1f} % |")


if __name__ == "__main__":
    main()
