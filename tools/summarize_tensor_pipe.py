"""Per-launch and per-family tensor-pipe table from tools/gpu_profile_pipe.sh (single-pass ncu metrics).

    python tools/summarize_tensor_pipe.py gpurun_out/pipe_TAG.csv gpurun_out/steps_TAG.json [--json profiles/tensor_pipe.json]
Markdown on stdout; with --json the per-family medians are merged into that file under "<arch>:<DxHxW>" (read by
bench.py for roofline.tensor_pipe_pct).  Two utilisation figures per launch:
  counter  = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed as ncu reports it;
  analytic = executed FLOPs / (sm__cycles_elapsed.max x 148 SMs x 8192 FLOP/clk/SM), executed FLOPs from the plan
             (dram_conv3d_plan_executed_flops: skipped taps removed, padded tiles included; M128 x N x K16 UTCHMMA =
             N/2 clocks at the dense 16-bit rate).
Identical launches (same layer shape) must agree to a few per cent — the check the round-1 capture failed.
"""
import collections
import csv
import json
import statistics
import sys

SM, FLOP_PER_CLK = 148, 8192.0


def family(name):
    if name == "conv1":
        return "stem"
    if name.startswith("layer"):
        layer, blk, conv = name.split(".")
        return f"{layer}.0.{conv}" if blk == "0" else f"{layer}.x"
    return name


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    steps = json.load(open(sys.argv[2]))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hi]
    col = {k: hdr.index(k) for k in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value")}
    launches = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= col["Metric Value"]:
            continue
        d = launches.setdefault(r[col["ID"]], {"kernel": r[col["Kernel Name"]].split("(")[0].replace("void ", "").replace("dram::", "")})
        try:
            v = float(r[col["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[col["Metric Unit"]]
        v *= {"us": 1e3, "ms": 1e6, "s": 1e9, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        d[r[col["Metric Name"]]] = v
    conv = [d for d in launches.values() if d["kernel"].startswith("conv3d_")]
    seq = [s for s in steps["steps"] if s["flops"] > 0]
    per_pass = len(seq)
    if len(conv) % per_pass:
        print(f"warning: {len(conv)} conv launches captured, {per_pass} per pass", file=sys.stderr)
    conv = conv[-2 * per_pass:] if len(conv) >= 2 * per_pass else conv[-per_pass:]
    print("| # | layer | kernel | us | tensor pipe % (counter) | tensor pipe % (analytic, executed) | TFLOP/s alg | DRAM MB | TC smem rd % |")
    print("|---|---|---|---|---|---|---|---|---|")
    fam = collections.OrderedDict()
    for i, d in enumerate(conv):
        s = seq[i % per_pass]
        ns = d.get("gpu__time_duration.sum", float("nan"))
        cyc = d.get("sm__cycles_elapsed.max", float("nan"))
        counter = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", float("nan"))
        analytic = 100.0 * s["executed_flops"] / (cyc * SM * FLOP_PER_CLK)
        mb = (d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)) / 1e6
        tc = d.get("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", float("nan"))
        if i >= len(conv) - per_pass:
            print(f"| {i % per_pass} | {s['name']} | {d['kernel']} | {ns / 1e3:.1f} | {counter:.1f} | {analytic:.1f} | "
                  f"{s['flops'] / ns / 1e3:.0f} | {mb:.1f} | {tc:.1f} |")
        f = fam.setdefault(family(s["name"]), {"counter": [], "analytic": [], "ns": [], "flops": 0.0, "n": 0})
        f["counter"].append(counter)
        f["analytic"].append(analytic)
        f["ns"].append(ns)
    print("\n| family | launches (2 passes) | us each (median) | tensor pipe % counter: median [min, max] | analytic: median [min, max] |")
    print("|---|---|---|---|---|")
    out = {}
    tot_ns = sum(sum(f["ns"]) for f in fam.values())
    w_counter = w_analytic = 0.0
    for name, f in fam.items():
        c, a = f["counter"], f["analytic"]
        print(f"| {name} | {len(c)} | {statistics.median(f['ns']) / 1e3:.1f} | {statistics.median(c):.1f} [{min(c):.1f}, {max(c):.1f}] | "
              f"{statistics.median(a):.1f} [{min(a):.1f}, {max(a):.1f}] |")
        out[name] = {"counter_pct": round(statistics.median(c), 2), "analytic_executed_pct": round(statistics.median(a), 2),
                     "spread_pct_points": round(max(c) - min(c), 2), "launches": len(c)}
        w_counter += sum(x * t for x, t in zip(c, f["ns"]))
        w_analytic += sum(x * t for x, t in zip(a, f["ns"]))
    out["all_conv_time_weighted"] = {"counter_pct": round(w_counter / tot_ns, 2), "analytic_executed_pct": round(w_analytic / tot_ns, 2)}
    print(f"\ntime-weighted over all conv launches: counter {w_counter / tot_ns:.1f} %, analytic (executed FLOPs) {w_analytic / tot_ns:.1f} % "
          f"(ncu serialises launches with cold caches and the clock it meets: compare shares, not absolutes)")
    if "--json" in sys.argv:
        path = sys.argv[sys.argv.index("--json") + 1]
        try:
            allp = json.load(open(path))
        except (OSError, ValueError):
            allp = {}
        out["source"] = f"{sys.argv[1]} (ncu single-pass metric list, tools/gpu_profile_pipe.sh)"
        allp[f"{steps['arch']}:{'x'.join(str(v) for v in steps['dims'])}"] = out
        json.dump(allp, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
