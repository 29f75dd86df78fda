"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/ncu_summary.py TAG launches.csv [full.ncu-rep ...]"""
import csv, collections, os, subprocess, sys
tag, launches = sys.argv[1], sys.argv[2]
reps = sys.argv[3:]
os.makedirs("profiles", exist_ok=True)
rows = [r for r in csv.reader(open(launches)) if len(r) > 10 and r[0].isdigit()]
per = collections.OrderedDict()
for r in rows:
    name, t = r[4], float(r[-1])
    grid, block = r[8], r[7]
    per.setdefault((name, grid, block), []).append(t)
total = sum(sum(v) for v in per.values())
with open("profiles/%s_launches.md" % tag, "w") as fh:
    fh.write("# ncu launch list (%s): `ncu --metrics gpu__time_duration.sum --clock-control none` over `python bench.py --steps 20 --warmup 5`\n\n" % tag)
    fh.write("Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes, with bench.py's CUDA-event timings.\n\n")
    fh.write("| kernel | grid | block | launches | mean us | min us | share of GPU time |\n|---|---|---|---:|---:|---:|---:|\n")
    for (name, grid, block), v in per.items():
        fh.write("| `%s` | %s | %s | %d | %.2f | %.2f | %.1f%% |\n" % (name, grid, block, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, 100 * sum(v) / total))
    fh.write("\nFirst %d launches captured; full list: `%s_launches.csv`.\n" % (len(rows), tag))
import shutil
shutil.copy(launches, "profiles/%s_launches.csv" % tag)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__issue_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "local_load", "local_store", "derived__smsp_sass_thread_inst_executed_op_local"]
for rep in reps:
    base = os.path.splitext(os.path.basename(rep))[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    hdr, units = rr[0], rr[1]
    with open("profiles/%s_key_metrics.md" % base, "w") as fh:
        fh.write("# %s: key metrics of `ncu --set full --clock-control none --import-source on` (one row per captured launch)\n\n" % base)
        for r in rr[2:]:
            fh.write("## %s  grid %s block %s\n\n| metric | unit | value |\n|---|---|---:|\n" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for i, h in enumerate(hdr):
                if any(h == k or (k in h and "local" in k) for k in KEYS):
                    fh.write("| %s | %s | %s |\n" % (h, units[i], r[i]))
            fh.write("\n")
    det = subprocess.run(["ncu", "-i", rep, "--page", "details"], stdout=subprocess.PIPE, text=True).stdout
    open("profiles/%s_details.txt" % base, "w").write(det)
print("profiles written")
