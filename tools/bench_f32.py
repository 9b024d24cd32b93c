"""GPU box only: throughput of the float (fp32) CoDeNet2x / COCO path (BASELINE config 5 geometry, 512x512) on one GPU."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from codenet_b200.arch import NetConfig
from codenet_b200.synth import make_raw_state, make_images
from codenet_b200.engine_f32 import EngineF32
B = int(os.environ.get("B", "32"))
for tag, cfg in (("1x_voc", NetConfig(num_classes=20)), ("2x_coco", NetConfig(num_classes=80, w2=True))):
    eng = EngineF32(cfg, make_raw_state(cfg, 0))
    x = torch.from_numpy(np.concatenate([make_images(8, 512, seed=5)] * (B // 8))).cuda()
    for _ in range(2):
        eng.detect(x)
    torch.cuda.synchronize()
    n, best = 5, None                       # eager launches: the host can be the bottleneck on a busy box -> best of 4 repeats
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.detect(x)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / n
        best = t if best is None else min(best, t)
    ms = best
    print(json.dumps({"config": "float %s 512x512 batch %d, eager fp32 SIMT path" % (tag, B), "ms_per_step": round(ms, 3),
                      "images_per_s": round(B / ms * 1e3, 1)}))
