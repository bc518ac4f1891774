"""Turn ncu outputs (gpurun_out/) into the tracked text summaries under profiles/.

    python scripts/summarize_profiles.py <launches.csv> <prof.ncu-rep> <tag>
"""
import csv, re, subprocess, sys, collections, io, os

KEY = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r["Metric Name"] == "gpu__time_duration.sum"]
    names = [(re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", ""), float(r["Metric Value"].replace(",", "")) / 1e3,
              r["Grid Size"], r["Block Size"]) for r in rows]
    idx = [i for i, n in enumerate(names) if n[0] == "k_motion"]
    out.write("ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
    out.write("launches captured: %d; k_motion launches at %s\n\n" % (len(names), idx))
    if len(idx) >= 2:
        s, e = idx[-2], idx[-1]
        out.write("One full filter step (between the last two k_motion launches):\n")
        tot = sum(d for _, d, _, _ in names[s:e] if not _.startswith("at::"))
        for n, d, g, b in names[s:e]:
            tag = "  (L2 flush of the benchmark, not part of the step)" if n.startswith("at::") else ""
            out.write("  %-28s %9.1f us  %5.1f %%   grid %-14s block %s%s\n" % (n[:28], d, 100 * d / tot if not tag else 0, g, b, tag))
        out.write("  %-28s %9.1f us\n\n" % ("sum of kernels", tot))
    agg = collections.OrderedDict()
    for n, d, _, _ in names:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += d
    out.write("All captured launches, by kernel:\n")
    for n, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.write("  %-40s n=%4d  total %10.1f us  mean %8.1f us\n" % (n[:40], c, d, d / c))


def full(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: j for j, h in enumerate(hdr)}
    out.write("\nncu --set full --clock-control none --import-source on  (%s)\n" % os.path.basename(rep))
    for d in data:
        out.write("\n== %s   grid %s block %s\n" % (d[col["Kernel Name"]][:60], d[col.get("Grid Size", 0)], d[col.get("Block Size", 0)]))
        for k in KEY:
            if k in col:
                out.write("  %-88s %-12s %s\n" % (k, units[col[k]], d[col[k]]))


def traffic_json(rep, tag, particles, sets):
    """profiles/likelihood_traffic.json: DRAM bytes per k_likelihood_g1 launch from the --set full capture (bench.py
    reads it for roofline.traffic instead of a literal)."""
    import json
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: j for j, h in enumerate(hdr)}
    vals = []
    for d in data:
        if "k_likelihood_g1" in d[col["Kernel Name"]]:
            def num(k):
                v = float(d[col[k]].replace(",", ""))
                u = units[col[k]].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            vals.append(num("dram__bytes_read.sum") + num("dram__bytes_write.sum"))
    if not vals:
        return
    out = {"kernel": "k_likelihood_g1", "dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
           "particles_per_launch": particles, "sets_per_launch": sets,
           "source": "profiles/%s_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum per k_likelihood_g1 launch, ncu --set full "
                     "--clock-control none, mean of %d launches (%d particles x %d sets per launch)" % (tag, len(vals), particles, sets)}
    json.dump(out, open("profiles/likelihood_traffic.json", "w"), indent=1)
    print(out)


if __name__ == "__main__":
    lpath, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
    if len(sys.argv) > 4:
        traffic_json(rep, tag, int(sys.argv[4]), int(sys.argv[5]) if len(sys.argv) > 5 else 2)
    os.makedirs("profiles", exist_ok=True)
    with open("profiles/%s_summary.txt" % tag, "w") as out:
        out.write("profile summary %s  (command: python bench.py --steps 30 --warmup 10 --quick, 1M particles x 360 beams)\n\n" % tag)
        launches(lpath, out)
        full(rep, out)
    print(open("profiles/%s_summary.txt" % tag).read())
