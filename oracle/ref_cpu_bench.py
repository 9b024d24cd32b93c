"""TEST / MEASUREMENT INFRASTRUCTURE: the reference arm of bench.py (`--impl reference`, `cpu_baseline.kind = "reference"`).

Times the UNMODIFIED reference -- PoseShuffleNetV2 (lib/models/networks/shufflenetv2_dcn.py:189-330) rewritten by
quantize_shufflenetv2_dcn (portable_quantizer/quantization_utils/quantize_model.py:7-82), called as `model(x)[-1]`, then
`hm.sigmoid_()` and `ctdet_decode` (lib/models/decode.py:474-505), i.e. the body of CtdetDetector.process
(lib/detectors/ctdet.py:29-46) without its unconditional torch.cuda.synchronize() -- on this host's CPU cores, as SURVEY.md
8(d) specifies: fp32, torch.set_num_threads(os.cpu_count()), batch 8, QuantAct ranges frozen, weights re-quantised on every
call as the reference does, the reference's CUDA-only deformable op replaced by torchvision's CPU kernel (ref_harness).

The reference tree is /root/reference in the build container and the copy installed by oracle/build_ref.build_py() into the
git-ignored oracle/_ref/py on the GPU box.  Never imported by the product.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def available():
    from oracle import ref_harness as H
    return H.available()


def build_model(cfg, offset_mode, res, calib_name="codenet1x_calib.npz"):
    """The reference's quantised model (fp32, eval, frozen ranges) carrying the synthetic weights of codenet_b200.synth and
    the calibrated BatchNorm statistics / QuantAct ranges of tests/golden/<calib_name> -- the network bench.py's GPU arm runs."""
    import torch
    from oracle import ref_harness as H
    from codenet_b200.synth import make_quant_state, make_raw_state
    calib = np.load(os.path.join(ROOT, "tests", "golden", calib_name))
    raw = make_raw_state(cfg, 0)
    for k in list(raw):
        if "bn/" + k in calib:
            raw[k] = np.asarray(calib["bn/" + k], dtype=np.float32)
    m = H.build_reference_model({k: torch.from_numpy(v) for k, v in raw.items()}, dict(cfg.head_list()), cfg.w2, cfg.maxpool,
                                dtype=torch.float32)
    m.eval()
    H.quantize_reference_model(m, cfg.w2, cfg.maxpool, cfg.w_bit, cfg.a_bit)
    st = make_quant_state(cfg, calib, offset_mode, res)
    sd = m.state_dict()
    for k, v in st.items():                             # frozen ranges (x_min / x_max buffers) in the reference's key space
        if k.endswith(".x_min") or k.endswith(".x_max"):
            sd[k].copy_(torch.from_numpy(np.asarray(v, np.float32)).reshape(sd[k].shape))
    H.freeze_ranges(m)
    H.set_integer_offsets(m, offset_mode == "round")
    return m.float()


def measure(cfg, offset_mode="round", res=512, batch=8, steps=5, warmup=2, threads=None, budget_s=120.0, K=100,
            calib_name="codenet1x_calib.npz"):
    """Returns dict(value images/s, cores, batch, steps, ms_per_step (median), times).  A step = forward + sigmoid + decode of
    `batch` images.  Stops early (never below 2 timed steps) once `budget_s` of timed work has been spent."""
    import torch
    from oracle import ref_harness as H
    from codenet_b200.synth import make_images
    threads = threads or (os.cpu_count() or 1)
    torch.set_num_threads(threads)
    R = H.load_reference()
    m = build_model(cfg, offset_mode, res, calib_name)
    x = torch.from_numpy(make_images(batch, res, seed=100))

    def step():
        with torch.no_grad():
            o = m(x)[-1]
            hm = o["hm"].sigmoid_()
            return R.decode.ctdet_decode(hm, o["wh"], reg=o.get("reg"), K=K)

    for _ in range(warmup):
        step()
    times = []
    t_all = time.perf_counter()
    for i in range(steps):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
        if i >= 1 and time.perf_counter() - t_all > budget_s:
            break
    med = float(np.median(times))
    return {"value": batch / med, "cores": threads, "batch": batch, "steps": len(times), "ms_per_step": med * 1e3,
            "times_s": [round(t, 4) for t in times]}


if __name__ == "__main__":
    from codenet_b200.arch import NetConfig
    r = measure(NetConfig(num_classes=20), sys.argv[1] if len(sys.argv) > 1 else "round",
                res=int(sys.argv[2]) if len(sys.argv) > 2 else 512)
    print(r)
