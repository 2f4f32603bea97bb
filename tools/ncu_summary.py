"""Summarise an .ncu-rep (read here, without a GPU): headline metrics + hottest source lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md"""
import collections, csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print("# ncu summary of `%s`\n" % rep)
WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_lsu.sum"]
for k, line in enumerate(rows[2:]):
    name = line[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("## launch %d: `%s`\n" % (k, name))
    print("| metric | value | unit |\n|---|---|---|")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("| %s | %s | %s |" % (w, line[i], units[i]))
    print()
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None
agg = collections.defaultdict(lambda: [0, 0])
files = collections.defaultdict(lambda: [0, 0])
for r in csv.reader(io.StringIO(src)):
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] and len(r) >= 8 and r[2] == "-":
        try:
            inst, smp = int(r[7]), int(r[4])
        except ValueError:
            continue
        key = (cur, r[0], r[1].strip()[:100])
        agg[key][0] += inst
        agg[key][1] += smp
        files[cur][0] += inst
        files[cur][1] += smp
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print("## where the warp-instructions and the stall samples go (source files)\n")
print("| file | instructions | samples |\n|---|---|---|")
for f, v in sorted(files.items(), key=lambda x: -x[1][1]):
    print("| %s | %.1f%% | %.1f%% |" % (f, 100 * v[0] / ti, 100 * v[1] / ts))
print("\n## hottest source lines by stall samples\n")
print("| samples | instructions | location | source |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    print("| %.1f%% | %.1f%% | %s:%s | `%s` |" % (100 * v[1] / ts, 100 * v[0] / ti, k[0], k[1], k[2].replace("|", "\\|")))
