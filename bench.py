#!/usr/bin/env python
"""Benchmark of the hot path: CoDeNet1x 512x512 W4A8 forward + ctdet decode (BASELINE.json config c), images/s.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU forward (oracle port) on the host

One JSON line on stdout (rank 0).  `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same metric
through Engine.run_host with HOST buffers (H2D of the images and D2H of the detections inside the timed region);
`roofline` = algorithmic bytes / device time of the dominant kernel family against the measured HBM peak;
`cpu_baseline` = the oracle port timed on this box's host cores on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "CoDeNet1x 512x512 W4A8 forward + ctdet decode throughput"
UNIT = "images/s"
FALLBACK_HBM_GBS = 6650.0


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ---- algorithmic bytes per op (DESIGN.md / SURVEY.md 8(d)) ----------------------------------------------------------
def op_bytes(plan, op, batch):
    """Bytes one launch must move: logical input elements read once + logical output elements written once."""
    a = op.a
    T = plan.tensors
    if op.kind == "stem":
        to = T[a["out_t"]]
        return batch * (3 * a["H"] * a["W"] * 4 + to.C * to.H * to.W)
    if op.kind == "dw" and a.get("fused_into_next"):
        # heads tail as one kernel: stored int8 input read once, fp32 planes written once; nothing in between touches HBM
        ti = T[a["in_t"]]
        return batch * (ti.C * ti.H * ti.W + 4 * a["fused_n_f32"] * 4 * ti.H * ti.W) + 9 * ti.C
    if op.kind == "pw" and a.get("fused_with_prev"):
        return 0
    if op.kind in ("dw", "deform"):
        ti, to = T[a["in_t"]], T[a["out_t"]]
        return batch * (ti.C * ti.H * ti.W + to.C * to.H * to.W) + 9 * ti.C
    ti = T[a["in_t"]]
    k_real = int((np.abs(a["wq"]).sum(0) > 0).sum())
    px = batch * ti.H * ti.W
    if a["n_f32"]:
        return px * (k_real + 4 * a["n_f32"])
    to = T[a["out_t"]]
    n_pass = sum(int(c[1]) for c in a["chunks"] if c[2] >= 0)
    n_new = sum(int(c[1]) for c in a["chunks"])
    return px * (k_real + n_pass + n_new + n_pass) + a["wq"].size


def family(op):
    if op.a.get("fused_into_next") or op.a.get("fused_with_prev"):
        return "heads_fused_kernel"
    if op.kind == "deform":                      # integer offsets: v3 kernel; bilinear: v2 kernel with the guarded fp32 blend
        return "deform_int_v3_kernel" if op.a.get("mode", 0) == 0 else "deform_dw_v2_kernel"
    # depthwise: dw3x3_tma_kernel (input tile staged by TMA; every depthwise layer of config c) or dw3x3_v2_kernel (LDG rows)
    return {"pw": "pw_gemm_tc_kernel", "dw": "dw3x3_kernels", "stem": "stem_kernel"}[op.kind]


def ncu_traffic():
    """DRAM bytes per step and kernel family from the latest committed ncu launch list (profiles/*_families.json)."""
    d = os.path.join(ROOT, "profiles")
    try:
        f = sorted(x for x in os.listdir(d) if x.endswith("_families.json"))[-1]
        j = json.load(open(os.path.join(d, f)))
        return {k: (v["dram_bytes"], v["launches"]) for k, v in j["families"].items()}, f
    except Exception:
        return {}, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from codenet_b200.arch import NetConfig
    from codenet_b200.engine import Engine
    from codenet_b200.synth import make_quant_state, make_images

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = NetConfig(num_classes=20)
    calib = np.load(os.path.join(ROOT, "tests", "golden", "codenet1x_calib.npz"))
    st = make_quant_state(cfg, calib, args.offset_mode, 512)
    B, R = args.batch, 512
    eng = Engine.from_state_dict(cfg, st, R, R, B, offset_mode=args.offset_mode, device=local)
    if args.no_fuse_heads:
        eng.set_option("fuse_heads", 0)
    if eng.heads_fused:                          # heads.dw2 + heads.out run as one kernel: account them as one launch
        ops = eng.plan.ops
        i = next(k for k, o in enumerate(ops) if o.name == "heads.dw2")
        ops[i].a["fused_into_next"], ops[i].a["fused_n_f32"], ops[i + 1].a["fused_with_prev"] = True, ops[i + 1].a["n_f32"], True
    eng.set_option("host_chunk", args.host_chunk)
    # synthetic images: 16 distinct ones per rank, tiled to the batch (805 MB fp32 at B=256: larger than L2)
    base = make_images(min(16, B), R, seed=100 + rank)
    reps = (B + base.shape[0] - 1) // base.shape[0]
    host = torch.from_numpy(np.concatenate([base] * reps)[:B].copy()).pin_memory()
    dev = host.cuda(non_blocking=True)
    torch.cuda.synchronize()
    out = {}

    def step():
        eng.run(dev, maps=False, dets=True, out=out)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through host buffers --------------------------------------------------------------------------
    # (1) detector-level input: uint8 HWC images at the input size, as cv2 hands them to the reference's run()
    #     (pre_process is then normalisation only, applied inside the stem kernel);
    # (2) model-level input: the normalised fp32 NCHW tensor the reference's model.forward takes.
    mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
    eng.set_normalization(mean, std)
    hnp = host.numpy()
    u8 = np.clip(np.rint((hnp.transpose(0, 2, 3, 1) * std + mean) * 255.0), 0, 255).astype(np.uint8)
    u8_t = torch.from_numpy(np.ascontiguousarray(u8)).pin_memory()
    u8 = u8_t.numpy()
    dets_t = torch.empty((B, eng.K, 6), dtype=torch.float32).pin_memory()
    inds_t = torch.empty((B, eng.K), dtype=torch.int32).pin_memory()
    dets_h, inds_h = dets_t.numpy(), inds_t.numpy()
    e2e_steps = max(2, min(args.steps, 10))

    def time_host(arr):
        eng.run_host(arr, dets_h, inds_h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            eng.run_host(arr, dets_h, inds_h)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * B * e2e_steps / float(tt.item())

    e2e_val = time_host(u8)
    e2e_f32 = time_host(hnp)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- per-kernel times (eager, CUDA events between launches on the launching stream) -----------------------------
    prof_runs = 3
    acc = None
    for _ in range(prof_runs):
        pr = eng.profile(dev)
        acc = [x[2] for x in pr] if acc is None else [a + x[2] for a, x in zip(acc, pr)]
    per_op = [a / prof_runs for a in acc]
    if args.dump_ops:
        rows = [{"op": o.name, "kind": o.kind, "ms": round(m, 4), "MB": round(op_bytes(eng.plan, o, B) / 1e6, 1),
                 "GBps": round(op_bytes(eng.plan, o, B) / m / 1e6, 1) if m > 0 else None} for o, m in zip(eng.plan.ops, per_op)]
        rows.append({"op": "ctdet_decode", "kind": "decode", "ms": round(per_op[-1], 4)})
        with open(args.dump_ops, "w") as f:
            json.dump(rows, f, indent=1)
    fam_ms, fam_bytes, deform_layers = {}, {}, []
    for op, msop in zip(eng.plan.ops, per_op):
        f = family(op)
        by = op_bytes(eng.plan, op, B)
        fam_ms[f] = fam_ms.get(f, 0.0) + msop
        fam_bytes[f] = fam_bytes.get(f, 0) + by
        if op.kind == "deform":
            deform_layers.append({"layer": op.name, "C": int(op.a["C"]), "H": int(eng.plan.tensors[op.a["out_t"]].H),
                                  "ms": round(msop, 4), "GBps": round(by / msop / 1e6, 1)})
    fam_ms["ctdet_decode"] = per_op[-1]
    fam_bytes["ctdet_decode"] = B * (eng.plan.cat + 4) * eng.plan.out_H * eng.plan.out_W * 4
    fam_launches = {"ctdet_decode": 2}
    for op in eng.plan.ops:
        if not op.a.get("fused_with_prev"):
            fam_launches[family(op)] = fam_launches.get(family(op), 0) + 1
    total_ms = sum(fam_ms.values())
    dom = max(fam_ms, key=fam_ms.get)
    peak, peak_src = peaks()
    ach = fam_bytes[dom] / fam_ms[dom] / 1e6
    traffic, traffic_src = ncu_traffic()
    tr = traffic.get(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(ach, 1), "peak": peak, "unit": "GB/s",
                "frac": round(ach / peak, 4),
                "traffic": (round(tr[0] / tr[1]) if tr and B == 256 else None),
                "traffic_note": "mean DRAM bytes per launch of this family (ncu dram__bytes_read+write, %s); algorithmic bytes "
                                "per launch: %d" % (traffic_src, fam_bytes[dom] // max(fam_launches.get(dom, 1), 1)),
                "launches_per_step": fam_launches.get(dom), "peak_source": peak_src,
                "share_of_step": round(fam_ms[dom] / total_ms, 3),
                "families": {k: {"ms": round(v, 4), "GBps": round(fam_bytes[k] / v / 1e6, 1), "share": round(v / total_ms, 3)}
                             for k, v in sorted(fam_ms.items(), key=lambda kv: -kv[1])}}
    dby = sum(op_bytes(eng.plan, op, B) for op in eng.plan.ops if op.kind == "deform")
    dms = sum(v for k, v in fam_ms.items() if k.startswith("deform_"))
    deform = {"GBps": round(dby / dms / 1e6, 1), "frac_of_hbm_peak": round(dby / dms / 1e6 / peak, 4), "layers": deform_layers}
    cpu = None if (args.no_cpu or world > 1) else cpu_baseline(args, sample_images=1)     # rank 0 at N = 1 only
    line = {
        "metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int8",
        "data": "synthetic", "impl": "codenet_b200",
        "config": {"workload": "BASELINE config c: CoDeNet1x 512x512 stride-4 W4A8, 20 classes, K=100, batch %d per GPU" % B,
                   "batch_per_gpu": B, "offset_mode": args.offset_mode, "arithmetic": "s8 x s8 -> s32 (4-bit weights, 8-bit activations), exact integer requantisation", "parallelism": "batch-sharded, no collective",
                   "l2": "inputs larger than L2 (%.0f MB fp32 images per step)" % (B * 3 * R * R * 4 / 1e6),
                   "outputs": "detections [B,100,6] (+ heat-map/wh/reg maps on request)"},
        "e2e": {"value": round(e2e_val, 1), "unit": UNIT, "h2d_bytes_per_step": int(B * 3 * R * R),
                "d2h_bytes_per_step": int(B * eng.K * (6 * 4 + 4)), "steps": e2e_steps, "host_chunk": args.host_chunk,
                "input": "uint8 HWC images at the input size in pinned host memory (the detector's run() input; "
                         "normalisation inside the stem kernel), detections [B,100,6] + indices back to the host",
                "api": "Engine.run_host -> cdn_engine_run_host_u8"},
        "e2e_fp32_input": {"value": round(e2e_f32, 1), "unit": UNIT, "h2d_bytes_per_step": int(B * 3 * R * R * 4),
                           "d2h_bytes_per_step": int(B * eng.K * (6 * 4 + 4)),
                           "input": "normalised fp32 NCHW tensor (the model.forward input), PCIe-bound",
                           "api": "Engine.run_host -> cdn_engine_run_host"},
        "gpu_launches": int(eng.num_launches * args.steps),
        "requant": dict(zip(("int_layers", "guarded_fp32_layers"), eng.requant_stats)),
        "heads_fused": bool(eng.heads_fused),
        "clocks": clocks, "roofline": roofline, "deform": deform, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ---- CPU arm: the oracle port of the reference's forward + decode -------------------------------------------------
def cpu_baseline(args, sample_images=1):
    """The oracle port (the reference path restated in numpy) on all host cores: one worker process per core, each looping
    over forward + decode of its own 512x512 image for ~10 s (oracle/cpu_bench.py)."""
    from oracle import cpu_bench
    value, workers, n = cpu_bench.measure(args.offset_mode, seconds=10.0)
    return {"value": round(value, 3), "unit": UNIT, "cores": workers, "kind": "port",
            "sample": "%d x (forward + decode) of 1 image at 512x512 in 10 s on %d worker processes, integer-exact numpy oracle "
                      "(oracle/int_oracle.py via oracle/cpu_bench.py)" % (n, workers)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cpu = cpu_baseline(args, sample_images=1)
    line = {"metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 / max(cpu["value"], 1e-9), 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp64/int64 (numpy)",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": "BASELINE config c: CoDeNet1x 512x512 stride-4 W4A8, 20 classes, K=100 (bounded sample on host cores)",
                       "offset_mode": args.offset_mode},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--offset-mode", default="round", choices=["round", "bilinear"])
    ap.add_argument("--host-chunk", type=int, default=64)
    ap.add_argument("--no-fuse-heads", action="store_true", help="run heads.dw2 and heads.out as separate kernels (A/B)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs only)")
    ap.add_argument("--dump-ops", default="", help="write the per-op device times (JSON) to this file")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
