"""The multi-GPU host logic on CPU: two gloo ranks shard a batch, run the (oracle) decode on their slice and gather
the detections; the result must equal the single-process answer.  The data path itself has no collective."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, q):
    sys.path.insert(0, ROOT)
    from codenet_b200.shard import my_slice, gather_detections
    from oracle import int_oracle as io
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                               # same data on every rank; each decodes its slice
    hm = rng.standard_normal((n_items, 3, 8, 8)); wh = rng.uniform(1, 9, (n_items, 2, 8, 8)); reg = rng.uniform(0, 1, (n_items, 2, 8, 8))
    b, e = my_slice(n_items, rank, world)
    dets, _ = io.ctdet_decode(hm[b:e], wh[b:e], reg[b:e], 10) if e > b else (np.zeros((0, 10, 6)), None)
    full = gather_detections(torch.from_numpy(np.asarray(dets, np.float32).reshape(e - b, 10, 6)), n_items)
    if rank == 0:
        ref, _ = io.ctdet_decode(hm, wh, reg, 10)
        q.put(bool(np.array_equal(full.numpy(), np.asarray(ref, np.float32))))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _run(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ok


def test_two_ranks_even_shards():
    _run(6)


def test_two_ranks_uneven_shards():
    _run(5)
