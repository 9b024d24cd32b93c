"""Per-step device times of the float path (config 5, batch 128, eager launches and graph replay) with clocks/power beside."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tools.bench_f32_config5 import _cfg, _state
from codenet_b200.engine_f32 import EngineF32
from codenet_b200.synth import make_images
raw, g = _state()
eng = EngineF32(_cfg(), raw, device=0, gemm=os.environ.get("CODENET_F32_GEMM", "tf32x3"))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = torch.from_numpy(np.concatenate([make_images(8, 512, seed=100)] * (B // 8)).copy()).cuda()
def smi():
    return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
for mode in ("forward", "detect"):
    ts = []
    for i in range(40):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if mode == "forward":
            eng.forward(x)
        else:
            eng.detect(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        if i % 10 == 9:
            print(mode, i, smi(), "mem_alloc_GB %.1f" % (torch.cuda.memory_allocated() / 1e9), flush=True)
    print(mode, " ".join("%.1f" % t for t in ts), flush=True)
