"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python profiles/summarize.py launches gpurun_out/launches_r01.csv            > profiles/<name>.txt
    python profiles/summarize.py raw gpurun_out/prof_find_r01.ncu-rep            > profiles/<name>.txt
"""
import collections
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.avg.per_second",
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(r[ki][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {len(data)} launches, {tot:.1f} us total (ncu: cold-cache, serialised -- compare SHARES)")
    print(f"{'n':>6} {'total_us':>11} {'avg_us':>10} {'share':>7}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:6d} {t:11.1f} {t / n:10.1f} {100 * t / tot:6.1f}%  {k}")


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# {path}: ncu --set full, per launch")
    for r in rows[2:]:
        print("kernel:", r[hdr.index("Kernel Name")][:110])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print(f"  {m:64s} {r[i]:>18s} {units[i]}")
        print()


def traffic(path):
    """Merge dram bytes per launch (read + write, averaged over the captured launches) into profiles/traffic.json,
    keyed by the kernel names bench.py uses."""
    import json
    import os
    import re
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = {}
    for r in rows[2:]:
        m = re.match(r"(?:void )?([A-Za-z_0-9]+(?:<[^>]*>)?)", r[ik])
        name = m.group(1)
        if name.startswith("k_find<"):                     # k_find<FILL, PROBE>: bench.py prints FILL as a bool
            name = name.replace("<0, ", "<false, ").replace("<1, ", "<true, ")
        b = float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
        acc.setdefault(name, []).append(b)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    cur = json.load(open(dst)) if os.path.exists(dst) else {}
    for k, v in acc.items():
        cur[k] = int(sum(v) / len(v))
    json.dump(cur, open(dst, "w"), indent=1, sort_keys=True)
    print(json.dumps(cur, indent=1))


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "traffic": traffic}[sys.argv[1]](sys.argv[2])
