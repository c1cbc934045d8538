"""Turns ncu outputs (gpurun_out/) into the small text summaries kept under profiles/.
  python profiles/summarize.py launches <launches.csv> > profiles/<name>.md
  python profiles/summarize.py kernels  <report.ncu-rep> > profiles/<name>.md
  python profiles/summarize.py traffic  <solve report.ncu-rep> > profiles/<name>.json   (DRAM bytes per primal colour sweep, for bench.py's roofline.traffic)
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[h], rows[h + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    if "--last-step" in sys.argv:       # one whole step: from the last launch of the step's first kernel to the end
        starts = [i for i, r in enumerate(data) if "bp_cells" in r[ki]]
        data = data[starts[-1]:]
    agg = collections.OrderedDict()
    for r in data:
        name = r[ki].split("(")[0].replace("void ", "")[:70]
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.1f} | {v[1] / tot:.3f} |")
    print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches (ncu serialises and runs cold: compare SHARES, not absolutes)")


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    idx = [(k, H.index(k)) for k in KEEP if k in H]
    ki = H.index("Kernel Name")
    for r in rows[2:]:
        print(f"### `{r[ki][:90]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k, i in idx:
            print(f"| {k} | {r[i]} | {U[i]} |")
        print()


def traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of the primal kernels of a `capture.sh solve` report, per colour sweep
    (= one primal_sweep_warp launch, the unit bench.py's roofline is quoted per)."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    ki, ri, wi, ti = H.index("Kernel Name"), H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    visit = solve = 0.0; nv = ns = 0; dual = 0.0; nd = 0
    gi, lastGrid, done = H.index("launch__grid_size"), None, False
    for r in rows[2:]:
        bytes_ = float(r[ri].replace(",", "")) * scale[U[ri]] + float(r[wi].replace(",", "")) * scale[U[wi]]
        if "primal_visit" in r[ki] or "primal_sweep" in r[ki]:
            # ONE whole iteration (every colour once): colours come largest first, so a grid that grows again starts the next iteration
            grid = int(r[gi].replace(",", ""))
            if lastGrid is not None and grid > lastGrid: done = True
            lastGrid = grid
            if not done: visit += bytes_; nv += 1
        elif "primal_solve" in r[ki]:
            if not done: solve += bytes_; ns += 1
        elif "dual_contacts" in r[ki]: dual += bytes_; nd += 1
    sweeps = max(ns, nv, 1)          # the fused sweep kernel (round 2) has no separate solve launch: one launch = one colour sweep
    print(json.dumps({"source": path.split("/")[-1], "how": "ncu --set full --clock-control none (cold caches, serialised launches), 1M-box pre-stacked grid, the first iteration (every colour once) of one step",
                      "colour_sweeps_captured": sweeps, "visit_kernel_launches": nv, "dram_bytes_per_colour_sweep": (visit + solve) / sweeps,
                      "visit_kernel_bytes_per_colour_sweep": visit / sweeps, "solve_kernel_bytes_per_colour_sweep": solve / sweeps,
                      "dual_pass_bytes": dual / nd if nd else None}, indent=1))


if __name__ == "__main__":
    {"launches": launches, "kernels": kernels, "traffic": traffic}[sys.argv[1]](sys.argv[2])
