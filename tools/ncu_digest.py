"""Digest of an ncu report for one kernel launch: key raw metrics, stall reasons, and a basic-block view (instructions executed and
stall samples per block of equally often executed SASS instructions).  usage: ncu_digest.py <report> [launch index] [min share %]"""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
g = {h: (r[i], units[i]) for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
for w in want:
    if w in g: print("%-72s %s %s" % (w, g[w][0], g[w][1]))
st = sorted(((float(v[0].replace(",", "")), h) for h, v in g.items() if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")), reverse=True)
print("stalls per issue: " + ", ".join("%s %.2f" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v) for v, h in st[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", str(idx), "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
isrc, ismp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
def I(x):
    try: return int(x)
    except Exception: return 0
data = [(x[isrc], I(x[ismp]), I(x[iex])) for x in rows[2:] if len(x) > iex]
tot = sum(d[2] for d in data); ts = sum(d[1] for d in data)
print("SASS lines %d, warp instructions %d, stall samples %d" % (len(data), tot, ts))
def opn(s):
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)', s)
    return m.group(2) if m else '?'
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and data[j + 1][2] == data[i][2]: j += 1
    cnt = data[i][2]; k = j - i + 1; share = 100.0 * cnt * k / max(tot, 1); smp = 100.0 * sum(d[1] for d in data[i:j + 1]) / max(ts, 1)
    if share > thr or smp > thr:
        ops = collections.Counter(opn(d[0]) for d in data[i:j + 1])
        print("[%4d..%4d] n=%3d exec=%9d instr%%=%5.1f stall%%=%5.1f  %s" % (i, j, k, cnt, share, smp, dict(ops.most_common(7))))
    i = j + 1
top = sorted(range(len(data)), key=lambda t: -data[t][1])[:12]
print("top stall lines:")
for t in top: print("   %4d samples=%5d exec=%9d %s" % (t, data[t][1], data[t][2], data[t][0][:70]))
