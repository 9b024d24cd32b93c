"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/ (run in the build container).

  gpurun_out/launches_full.csv            -> profiles/<tag>_launches.csv (one line per launch: kernel, us, DRAM MB)
                                             profiles/<tag>_families.json (per kernel family: launches/step, ms, share,
                                             DRAM bytes per step) -- bench.py reads `traffic` from it
  gpurun_out/prof_full_<kernel>.ncu-rep   -> profiles/<tag>_full_<kernel>.txt (key raw metrics + top stall lines)
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
LAUNCHES_PER_STEP = int(os.environ.get("LAUNCHES_PER_STEP", "44"))   # kernels of one forward + decode (memset excluded; fused groups are one launch each)


def fam(name):
    n = name.split("(")[0].replace("void ", "").split("<")[0]
    if n in ("dw3x3_v2_kernel", "dw3x3_tma_kernel"):
        return "dw3x3_kernels"                                        # bench.py's family names
    if n in ("deform_tile_int_kernel", "deform_int_v3_kernel"):
        return "deform_int_kernel"
    if n in ("deform_tile_bil_kernel", "deform_dw_v2_kernel"):
        return "deform_bilinear_kernel"
    if n in ("unit_fused_kernel", "unit_s2_fused_kernel"):
        return "unit_fused_kernel"
    if n in ("stem_kernel", "stem_fast_kernel"):
        return "stem_kernel"
    if n in ("ctdet_peaks_rows_kernel", "ctdet_peaks_kernel", "ctdet_decode_kernel"):
        return "ctdet_decode"
    return n


def launch_list():
    rows = [r for r in csv.reader(open(os.path.join(GO, "launches_full.csv"))) if len(r) > 10]
    hdr = rows[0]
    iid, ik, im, iu, iv = (hdr.index(x) for x in ("ID", "Kernel Name", "Metric Name", "Metric Unit", "Metric Value"))
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[iid]), {"kernel": r[ik].split("(")[0].replace("void ", "")})
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        if r[im] == "gpu__time_duration.sum":
            d["us"] = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        else:
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            d["dram_rd" if "read" in r[im] else "dram_wr"] = v * mult
    L = list(launches.values())
    # the bench runs 3 warm-up + 1 timed step on the device-resident fp32 batch first: take the LAST full step of those
    first_e2e = next((i for i, d in enumerate(L) if d["kernel"].startswith(("stem_kernel<1>", "stem_fast_kernel<1"))), len(L))
    step = L[first_e2e - LAUNCHES_PER_STEP:first_e2e]
    assert step and step[0]["kernel"].startswith(("stem_kernel", "stem_fast_kernel")), (first_e2e, step[:2])
    with open(os.path.join(OUT, TAG + "_launches.csv"), "w") as f:
        f.write("# one timed step of bench.py (batch 256), ncu --metrics gpu__time_duration.sum,dram__bytes_*.sum "
                "--clock-control none: cold-cache, serialised -> compare SHARES\nidx,kernel,us,dram_read_MB,dram_write_MB\n")
        for i, d in enumerate(step):
            f.write("%d,%s,%.2f,%.2f,%.2f\n" % (i, d["kernel"], d["us"], d.get("dram_rd", 0) / 1e6, d.get("dram_wr", 0) / 1e6))
    fams = collections.OrderedDict()
    for d in step:
        a = fams.setdefault(fam(d["kernel"]), {"launches": 0, "ms": 0.0, "dram_bytes": 0.0})
        a["launches"] += 1; a["ms"] += d["us"] / 1e3; a["dram_bytes"] += d.get("dram_rd", 0) + d.get("dram_wr", 0)
    tot = sum(a["ms"] for a in fams.values())
    for a in fams.values():
        a["share"] = round(a["ms"] / tot, 4); a["ms"] = round(a["ms"], 4); a["dram_bytes"] = int(a["dram_bytes"])
    json.dump({"batch": 256, "step_ms_under_ncu": round(tot, 3), "families": fams}, open(os.path.join(OUT, TAG + "_families.json"), "w"), indent=1)
    print(json.dumps(fams, indent=1))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"]


def full(kernel):
    rep = os.path.join(GO, "prof_full_%s.ncu-rep" % kernel)
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# ncu --set full --clock-control none --import-source on, bench.py --steps 1 --warmup 3 (batch 256), kernel regex %s" % kernel]
    for r in rows[2:]:
        out.append("launch %s  %s" % (r[hdr.index("ID")], r[hdr.index("Kernel Name")][:90]))
        for w in WANT:
            if w in hdr:
                out.append("    %-70s %s %s" % (w, r[hdr.index(w)], units[hdr.index(w)]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", "0", "--launch-count", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    if len(rows) > 3:
        hdr = rows[1]
        iS, iI, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[iS].isdigit()]
        data = data[:len(data) // 2] if len(data) > 1 and data[0][isrc] == data[len(data) // 2][isrc] else data
        agg = collections.Counter()
        for r in data:
            for c in stall:
                if c < len(r) and r[c].isdigit():
                    agg[hdr[c]] += int(r[c])
        out.append("first launch: warp-state samples by reason: " + ", ".join("%s %d" % kv for kv in agg.most_common(9)))
        out.append("first launch: SASS lines with most samples (samples, executions, instruction, top reasons):")
        for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:16]):
            r = data[i]
            st = sorted(((hdr[c][6:], int(r[c])) for c in stall if c < len(r) and r[c].isdigit() and int(r[c]) > 0), key=lambda kv: -kv[1])[:2]
            out.append("    %5s %9s  %-60s %s" % (r[iS], r[iI], r[isrc].strip()[:60], st))
    open(os.path.join(OUT, "%s_full_%s.txt" % (TAG, kernel)), "w").write("\n".join(out) + "\n")
    print("wrote", kernel)


if __name__ == "__main__":
    launch_list()
    for k in ("unit_fused", "stem_fast", "pw_gemm_tc", "deform_tile_int", "deform_tile_bil", "dw3x3_tma", "heads_fused"):
        full(k)
