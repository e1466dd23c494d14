"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps K --warmup W
--batch B --no-cpu-baseline` into a per-kernel table of the K timed steps (markdown on stdout).

    python tools/summarize_launches.py gpurun_out/launches_TAG.csv [K] [W]
A step of the device-resident arm starts at `window_stats_kernel` (K8, once per volume) and ends at
`dram_finalize_kernel` (K7); framework kernels (torch copies) inside a step are listed as `(torch)`.
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    ki, mi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    launches = [(r[ki], float(r[mi].replace(",", "")) / 1e6) for r in rows[hi + 1:] if len(r) > mi]  # ms
    ends = [i for i, (k, _) in enumerate(launches) if "dram_finalize_kernel" in k]
    starts = [i for i, (k, _) in enumerate(launches) if "window_stats_kernel" in k]
    # volumes per step = window_stats launches between two dram_finalize launches
    first = ends[W - 1] + 1
    last = ends[W + K - 1]
    seg = launches[first:last + 1]
    tot = collections.OrderedDict()
    for k, ms in seg:
        name = k.split("(")[0].replace("void ", "").replace("dram::", "")
        if name.startswith("at::") or "elementwise" in name or "vectorized" in name:
            name = "(torch)"
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + ms)
    total = sum(t for _, t in tot.values())
    print(f"`{path}`: launches {first}..{last} = the {K} timed steps ({len(seg)} launches, {len(seg) // K} per step)\n")
    print("| kernel | launches/step | ms/step | share |")
    print("|---|---|---|---|")
    for name, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| {name} | {n / K:g} | {t / K:.3f} | {100 * t / total:.1f} % |")
    print(f"| total | {len(seg) / K:g} | {total / K:.3f} | 100 % |")
    conv = sum(t for k, (n, t) in tot.items() if k.startswith("conv3d_"))
    print(f"\nconv3d_* share of the step by ncu: {100 * conv / total:.1f} %")


if __name__ == "__main__":
    main()
