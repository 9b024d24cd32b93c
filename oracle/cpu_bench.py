"""TEST / MEASUREMENT INFRASTRUCTURE: the CPU arm of bench.py (`cpu_baseline`, `--impl reference`).

Times the integer-exact numpy port of the reference path (oracle/int_oracle.py: W4A8 forward + ctdet decode of one 512x512
image per call) on ALL host cores: one worker process per core, each with its own copy of the network and its own image,
all looping until a common deadline; throughput = images finished / wall time.  The port is single-threaded per image
(integer matmuls do not go through BLAS), so processes are the only way it can use the box.  Spawned (not forked): the
parent has usually initialised CUDA.  Never imported by the product."""
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, offset_mode, t_start, seconds, q):
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    import numpy as np
    from codenet_b200.arch import NetConfig
    from codenet_b200.synth import make_quant_state, make_images
    from oracle import int_oracle as io
    cfg = NetConfig(num_classes=20)
    calib = np.load(os.path.join(ROOT, "tests", "golden", "codenet1x_calib.npz"))
    st = make_quant_state(cfg, calib, offset_mode, 512)
    x = make_images(1, 512, seed=100 + rank)
    o = io.IntOracle(cfg, st, offset_mode)

    def one():
        out = o.forward(x)
        io.ctdet_decode(out["hm"], out["wh"], out["reg"], 100)

    one()                                           # warm-up (also: every worker is ready before the window opens)
    while time.time() < t_start:
        time.sleep(0.01)
    n, t0 = 0, time.time()
    while time.time() - t_start < seconds:
        one()
        n += 1
    q.put((n, t0, time.time()))


def measure(offset_mode="round", seconds=10.0, workers=None):
    """Returns (images_per_second, workers, images_done)."""
    workers = workers or (os.cpu_count() or 1)
    env_before = os.environ.get("OMP_NUM_THREADS")
    os.environ["OMP_NUM_THREADS"] = "1"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    t_start = time.time() + 8.0 + 0.05 * workers       # import + network construction + one warm-up image per worker
    procs = [ctx.Process(target=_worker, args=(r, offset_mode, t_start, seconds, q), daemon=True) for r in range(workers)]
    for p in procs:
        p.start()
    res = [q.get(timeout=seconds + 120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    if env_before is None:
        os.environ.pop("OMP_NUM_THREADS", None)
    else:
        os.environ["OMP_NUM_THREADS"] = env_before
    n = sum(r[0] for r in res)
    wall = max(r[2] for r in res) - min(r[1] for r in res)
    return n / wall, workers, n


if __name__ == "__main__":
    v, w, n = measure(sys.argv[1] if len(sys.argv) > 1 else "round")
    print("%.2f images/s on %d worker processes (%d images)" % (v, w, n))
